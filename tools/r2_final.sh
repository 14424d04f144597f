# final single-GPU confirmation of the tree: the driver's own sequence (pytest -m gpu, smoke(), default bench.py)
export OMP_NUM_THREADS=16
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/r02_final_gputests.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02_final_gputests.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_final_smoke.log | cut -c1-200
timeout 900 python bench.py > gpurun_out/r02_final_bench.log 2>&1; echo "bench rc=$?"
grep '"metric"' gpurun_out/r02_final_bench.log | tail -1 > gpurun_out/r02_final_bench.json
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02_final_bench.json").read())
    r = d["roofline"]
    print("value %.0f ms %.2f e2e %.0f (%.2f ms) conv_tflops %.0f frac %.3f hbm %s cpu %s clocks %s launches %s" % (
        d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], r["achieved"], r["frac"],
        r["hbm_family"].get("gbs"), d.get("cpu_baseline", {}).get("value"), d["clocks"], d["gpu_launches"]))
except Exception as e:
    print("bench FAILED", e)
PY
grep -q '"metric"' gpurun_out/r02_final_bench.log || tail -25 gpurun_out/r02_final_bench.log
