# default bench line (with the calibrated CPU arm) and the reference arm, as the driver runs them
export OMP_NUM_THREADS=16
mkdir -p gpurun_out
( time timeout 900 python bench.py > gpurun_out/r02_final2_bench.log 2>&1 ) 2> gpurun_out/r02_final2_time.txt; echo "bench rc=$?"; grep real gpurun_out/r02_final2_time.txt
grep '"metric"' gpurun_out/r02_final2_bench.log | tail -1 > gpurun_out/r02_final2_bench.json
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02_final2_bench.json").read())
    print("value %.0f ms %.2f e2e %.0f cpu %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("cpu_baseline")))
except Exception as e:
    print("bench FAILED", e)
PY
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_final2_ref.log 2>&1 ) 2> gpurun_out/r02_final2_reftime.txt; echo "ref rc=$?"; grep real gpurun_out/r02_final2_reftime.txt
tail -1 gpurun_out/r02_final2_ref.log | cut -c1-300
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02_final2_ref.log").read().strip().splitlines()[-1])
    print("ref value", d["value"], d["cpu_baseline"])
except Exception as e:
    print("ref FAILED", e)
PY
