# round-2 GPU session 1: parity on the full config, new kernels, bench A/B (storage modes, augmentation, grid waves)
export OMP_NUM_THREADS=16
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_parity_full_gpu.py -q -m gpu -s > gpurun_out/r2_parity1.log 2>&1
timeout 900 python -m pytest tests/ -q -m gpu --deselect tests/test_parity_full_gpu.py > gpurun_out/r2_gputests1.log 2>&1
b() { tag=$1; shift; env "$@" timeout 400 python bench.py --steps 20 --warmup 5 --no_cpu_baseline $EXTRA > gpurun_out/r2_bench_$tag.log 2>&1; grep '"metric"' gpurun_out/r2_bench_$tag.log | tail -1 > gpurun_out/r2_bench_$tag.json; }
EXTRA="--no_augmentation" b noaug_bf16 A=1
EXTRA="--no_augmentation" b noaug_fp16_fp16 OS2S_ACT_DTYPE=fp16
EXTRA="--no_augmentation" b noaug_fp16_fp32 OS2S_ACT_DTYPE=fp16 OS2S_CONV_DTYPE=fp32
EXTRA="--no_augmentation" b noaug_bf16_w2 OS2S_CONV_WAVES=2
EXTRA="" b aug_bf16 A=1
EXTRA="" b aug_bf16_w2 OS2S_CONV_WAVES=2
EXTRA="" b aug_bf16_w3 OS2S_CONV_WAVES=3
EXTRA="" b aug_fp16_fp32 OS2S_ACT_DTYPE=fp16 OS2S_CONV_DTYPE=fp32
echo "=== parity"; grep "full 10x5\|passed\|failed\|Error\|toy-speech\|^FAILED\|^E  " gpurun_out/r2_parity1.log | head -60
echo "=== gpu tests"; tail -25 gpurun_out/r2_gputests1.log
echo "=== bench"; for f in gpurun_out/r2_bench_*.json; do echo $f; python - $f <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    r=d["roofline"]
    print("  value %.0f  ms %.2f  e2e %.0f  conv_tflops %.0f conv_ms %.2f  hbm %s  clocks %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], r["achieved"], r["conv_ms_per_step"], json.dumps(r.get("hbm_family")), d["clocks"]["sm_mhz"]))
except Exception as e:
    print("  FAILED", e)
PY
done
for f in gpurun_out/r2_bench_*.log; do if ! grep -q '"metric"' $f; then echo "--- $f"; tail -15 $f; fi; done
