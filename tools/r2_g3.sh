export OMP_NUM_THREADS=16
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -q -m gpu -x --deselect tests/test_parity_full_gpu.py > gpurun_out/r2_gputests3.log 2>&1
timeout 1500 python -m pytest tests/test_parity_full_gpu.py -q -m gpu -s > gpurun_out/r2_parity3.log 2>&1
b() { tag=$1; shift; env "$@" timeout 400 python bench.py --steps 20 --warmup 5 --no_cpu_baseline $EXTRA > gpurun_out/r2c_bench_$tag.log 2>&1; grep '"metric"' gpurun_out/r2c_bench_$tag.log | tail -1 > gpurun_out/r2c_bench_$tag.json; }
EXTRA="" b aug_bf16 A=1
EXTRA="--no_augmentation" b noaug_bf16 A=1
EXTRA="" b aug_bf16_noskip OS2S_SKIP_TILES=0
for tool in racecheck; do
  timeout 600 compute-sanitizer --tool $tool --log-file gpurun_out/r02_sanitizer_$tool.log python tools/sanitize_small.py > gpurun_out/r02_san_$tool.out 2>&1
done
for tool in memcheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --log-file gpurun_out/r02_sanitizer_$tool.log python tools/sanitize_small.py > gpurun_out/r02_san_$tool.out 2>&1
done
echo "=== gpu tests"; grep "^FAILED\|passed\|failed\|^E   " gpurun_out/r2_gputests3.log | head -40
echo "=== parity"; grep "full 10x5\|passed\|failed\|toy-speech\|trained\|^FAILED\|^E  " gpurun_out/r2_parity3.log | head -40
echo "=== bench"; for f in gpurun_out/r2c_bench_*.json; do echo $f; python - $f <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    r=d["roofline"]
    print("  value %.0f  ms %.2f  e2e %.0f (%.2f ms)  conv_tflops %.0f conv_ms %.2f clocks %s  by_kind %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], r["achieved"], r["conv_ms_per_step"], d["clocks"]["sm_mhz"], {k:v["ms_per_step"] for k,v in r["by_kind"].items()}))
except Exception as e:
    print("  FAILED", e)
PY
done
for f in gpurun_out/r2c_bench_*.log; do if ! grep -q '"metric"' $f; then echo "--- $f"; tail -25 $f; fi; done
for tool in memcheck synccheck racecheck; do tail -2 gpurun_out/r02_sanitizer_$tool.log; done
