"""Summarise an `ncu --csv --metrics gpu__time_duration.sum` launch list into per-kernel totals and
shares of the step (markdown + json).  usage: summarize_ncu.py launches.csv out_prefix"""
import csv
import json
import re
import sys
from collections import OrderedDict


def main():
    path, out = sys.argv[1], sys.argv[2]
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1, "s": 1e9,
                  "second": 1e9}.get(unit, 1)
        rows.append((r["Kernel Name"], ns))
    tot = sum(ns for _, ns in rows)
    agg = OrderedDict()
    for name, ns in rows:
        key = re.sub(r"\(.*", "", name)
        key = re.sub(r"^void ", "", key)
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += ns
    items = sorted(agg.items(), key=lambda kv: -kv[1][1])
    with open(out + ".md", "w") as f:
        f.write("# ncu launch list summary (%s)\n\n%d launches, %.3f ms total kernel time (cold-cache, serialised: "
                "compare SHARES, not absolutes)\n\n| kernel | launches | total ms | share |\n|---|---|---|---|\n"
                % (path, len(rows), tot / 1e6))
        for k, (n, ns) in items:
            f.write("| `%s` | %d | %.3f | %.1f%% |\n" % (k[:110], n, ns / 1e6, 100 * ns / tot))
    json.dump({"launches": len(rows), "total_ms": tot / 1e6,
               "kernels": [{"name": k, "launches": n, "ms": ns / 1e6, "share": ns / tot} for k, (n, ns) in items]},
              open(out + ".json", "w"), indent=1)
    print(open(out + ".md").read())


if __name__ == "__main__":
    main()
