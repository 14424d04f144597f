"""SASS / resource evidence of the built library (no GPU needed): per kernel, the counts of the tcgen05 / TMA / TMEM /
mbarrier instructions (`cuobjdump -sass`) and registers / shared memory / spills (`cuobjdump -res-usage`).

    python tools/sass_evidence.py profiles/r02
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "openseq2seq_b200", "lib", "libos2s_b200.so")
PAT = re.compile(r"\b(UTC[A-Z0-9_.]*|UTMA[A-Z0-9_.]*|LDTM[A-Z0-9_.x]*|STTM[A-Z0-9_.x]*|SYNCS[A-Z0-9_.]*|UBLKCP[A-Z0-9_.]*|"
                 r"REDG[A-Z0-9_.]*|MULTIMEM[A-Z0-9_.]*|ATOMG?\.[A-Z0-9_.]*SYS[A-Z0-9_.]*)")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return [re.sub(r"\(.*$", "", o.replace("(anonymous namespace)", "{anon}")) for o in out]


def main():
    prefix = sys.argv[1]
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur:
            for op in PAT.findall(line.split("/*")[1] if line.count("/*") > 1 else line):
                counts[cur][op] += 1
    names = demangle(list(counts))
    with open(prefix + "_sass_mnemonics.txt", "w") as f:
        f.write("# SASS evidence (cuobjdump -sass openseq2seq_b200/lib/libos2s_b200.so, sm_100a): tcgen05 (UTC*) / TMA "
                "(UTMA*, UBLKCP) / TMEM (LDTM) / mbarrier (SYNCS) / system-scope atomics per kernel\n\n")
        for (mangled, c), name in zip(counts.items(), names):
            for op, n in c.most_common():
                f.write("%d\t %s %s\n" % (n, name, op))
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    rows, fn = [], None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            fn = m.group(1)
        elif fn and "REG:" in line:
            rows.append((fn, line.strip()))
            fn = None
    dn = demangle([r[0] for r in rows])
    with open(prefix + "_kernel_resources.txt", "w") as f:
        f.write("# Resource usage per kernel (cuobjdump -res-usage, sm_100a)\n")
        for (_, usage), name in zip(rows, dn):
            f.write("%s %s\n" % (name, usage))
    print(len(counts), "kernels,", len(rows), "resource rows")


if __name__ == "__main__":
    main()
