set -x
timeout 1200 python -m pytest tests/test_parity_full_gpu.py -q -m gpu -s > gpurun_out/r2_parity1.log 2>&1
timeout 600 python -m pytest tests/ -x -q -m gpu --deselect tests/test_parity_full_gpu.py 2>&1 | tail -8 > gpurun_out/r2_gputests1.log
for m in "bf16 fp16" "fp16 fp16" "fp16 fp32"; do set -- $m; OS2S_ACT_DTYPE=$1 OS2S_CONV_DTYPE=$2 timeout 300 python bench.py --steps 20 --warmup 5 --no_cpu_baseline > gpurun_out/r2_bench_$1_$2.log 2>&1; grep '"metric"' gpurun_out/r2_bench_$1_$2.log | tail -1 > gpurun_out/r2_bench_$1_$2.json; done
grep -v "^$" gpurun_out/r2_parity1.log | grep "full 10x5\|passed\|failed\|Error\|error\|assert\|toy-speech" | head -60; cat gpurun_out/r2_gputests1.log; cut -c1-300 gpurun_out/r2_bench_*.json
