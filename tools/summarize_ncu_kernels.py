"""Per-kernel evidence table from one ncu pass over a training step:

  ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/kernels.csv \
      --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,\
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed python tools/profile_step.py
  python tools/summarize_ncu_kernels.py gpurun_out/kernels.csv profiles/r01_kernel_evidence

For every kernel family: launches, total time, DRAM bytes moved, achieved DRAM GB/s (bytes / time) and
the time-weighted tensor-pipe activity.  ncu serialises the launches (cold L2, no power throttling):
compare shares and per-kernel rates, not the step total.
"""
import csv
import json
import re
import sys
from collections import OrderedDict

UNIT = {"ns": 1e-9, "nsecond": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3, "s": 1.0, "second": 1.0,
        "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "%": 1.0}


def main():
    path, out = sys.argv[1], sys.argv[2]
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    per_launch = OrderedDict()
    for r in csv.DictReader(lines):
        key = (r["ID"], r["Kernel Name"])
        d = per_launch.setdefault(key, {})
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        d[r["Metric Name"]] = v * UNIT.get(r.get("Metric Unit", ""), 1.0)
    agg = OrderedDict()
    for (_, name), d in per_launch.items():
        k = re.sub(r"\(.*", "", name)
        k = re.sub(r"^void ", "", k)
        a = agg.setdefault(k, {"launches": 0, "s": 0.0, "rd": 0.0, "wr": 0.0, "tensor_w": 0.0})
        t = d.get("gpu__time_duration.sum", 0.0)
        a["launches"] += 1
        a["s"] += t
        a["rd"] += d.get("dram__bytes_read.sum", 0.0)
        a["wr"] += d.get("dram__bytes_write.sum", 0.0)
        a["tensor_w"] += t * d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0)
    tot = sum(a["s"] for a in agg.values())
    rows = []
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["s"]):
        rows.append({"kernel": k, "launches": a["launches"], "ms": round(a["s"] * 1e3, 3),
                     "share": round(a["s"] / tot, 4), "dram_read_MB": round(a["rd"] / 1e6, 1),
                     "dram_write_MB": round(a["wr"] / 1e6, 1),
                     "dram_GBps": round((a["rd"] + a["wr"]) / a["s"] / 1e9, 1) if a["s"] else 0.0,
                     "tensor_pipe_active_pct": round(a["tensor_w"] / a["s"], 1) if a["s"] else 0.0})
    json.dump({"source": path, "total_ms": round(tot * 1e3, 3), "kernels": rows}, open(out + ".json", "w"), indent=1)
    with open(out + ".md", "w") as f:
        f.write("# Per-kernel ncu evidence (%s)\n\n%d launches, %.3f ms of kernel time (serialised, cold L2, "
                "unthrottled clocks: compare shares and rates)\n\n" % (path, sum(r["launches"] for r in rows), tot * 1e3))
        f.write("| kernel | launches | ms | share | DRAM read MB | DRAM write MB | DRAM GB/s | tensor pipe active % |\n")
        f.write("|---|---|---|---|---|---|---|---|\n")
        for r in rows:
            f.write("| `%s` | %d | %.3f | %.1f%% | %.1f | %.1f | %.1f | %.1f |\n" % (
                r["kernel"][:90], r["launches"], r["ms"], 100 * r["share"], r["dram_read_MB"], r["dram_write_MB"],
                r["dram_GBps"], r["tensor_pipe_active_pct"]))
    print(open(out + ".md").read())


if __name__ == "__main__":
    main()
