"""Pull the headline metrics of an `ncu --set full` report (read here, no GPU needed):
   python tools/summarize_ncu_full.py gpurun_out/conv_full.ncu-rep profiles/ncu_conv_summary.json"""
import csv
import io
import json
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__cycles_active.avg",
        "sm__inst_executed_pipe_tensor.sum", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    res = []
    for r in data:
        d = {"kernel": r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"}
        for k in hdr:
            if any(k == w or k.startswith(w) for w in WANT) or "tensor" in k:
                try:
                    d[k + " [" + units[hdr.index(k)] + "]"] = float(r[hdr.index(k)].replace(",", ""))
                except ValueError:
                    pass
        res.append(d)
    summary = {"report": rep, "launches": res}
    if res:
        rd = [v for k, v in res[0].items() if k.startswith("dram__bytes_read.sum")]
        wr = [v for k, v in res[0].items() if k.startswith("dram__bytes_write.sum")]
        if rd and wr:
            # unit may be Mbyte / Gbyte: normalise via the unit string
            def to_bytes(key_prefix):
                for k, v in res[0].items():
                    if k.startswith(key_prefix):
                        u = k.split("[")[-1].rstrip("]").lower()
                        mult = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
                        return v * mult
                return None
            summary["dram_bytes_per_launch"] = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
    json.dump(summary, open(out, "w"), indent=1)
    print(json.dumps(summary, indent=1)[:3000])


if __name__ == "__main__":
    main()
