# N=2 A/B of NCCL CTA limits and graph capture with communication (run under gpurun --gpus 2)
export OMP_NUM_THREADS=8
run() { tag=$1; shift; env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no_cpu_baseline > gpurun_out/n2_$tag.log 2>&1; grep '"metric"' gpurun_out/n2_$tag.log | tail -1 | python -c "
import sys, json
l=[x for x in sys.stdin if x.startswith('{')]
d=json.loads(l[-1]) if l else {}
print(json.dumps({'tag': '$tag', 'value': d.get('value'), 'ms': d.get('ms_per_step'), 'sm_mhz': (d.get('clocks') or {}).get('sm_mhz')}))
" | tee -a gpurun_out/n2sweep.jsonl; tail -3 gpurun_out/n2_$tag.log | cut -c1-300; }
rm -f gpurun_out/n2sweep.jsonl
run default A=1
run ctas8 NCCL_MAX_CTAS=8
run ctas16 NCCL_MAX_CTAS=16
run graph OS2S_GRAPH_DIST=1
