# N = 2 A/B: NCCL CTA limits, graph capture with the collectives, conv grid waves (run under gpurun --gpus 2)
export OMP_NUM_THREADS=8
mkdir -p gpurun_out
rm -f gpurun_out/r02_n2sweep.jsonl
run() { tag=$1; shift; env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 12 --warmup 4 --no_cpu_baseline > gpurun_out/r02_n2_$tag.log 2>&1; grep '"metric"' gpurun_out/r02_n2_$tag.log | tail -1 | python -c "
import sys, json
l=[x for x in sys.stdin if x.startswith('{')]
d=json.loads(l[-1]) if l else {}
r=d.get('roofline') or {}
print(json.dumps({'tag': '$tag', 'value': d.get('value'), 'ms': d.get('ms_per_step'), 'e2e_ms': (d.get('e2e') or {}).get('ms_per_step'), 'conv_tflops': r.get('achieved'), 'sm_mhz': (d.get('clocks') or {}).get('sm_mhz')}))
" | tee -a gpurun_out/r02_n2sweep.jsonl; tail -2 gpurun_out/r02_n2_$tag.log | cut -c1-200; }
timeout 300 python bench.py --steps 12 --warmup 4 --no_cpu_baseline > gpurun_out/r02_n1_ref.log 2>&1; grep '"metric"' gpurun_out/r02_n1_ref.log | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print(json.dumps({'tag':'n1','value':d['value'],'ms':d['ms_per_step'],'e2e_ms':d['e2e']['ms_per_step'],'conv_tflops':d['roofline']['achieved']}))" | tee -a gpurun_out/r02_n2sweep.jsonl
run default A=1
run graph OS2S_GRAPH_DIST=1
run waves2 OS2S_CONV_WAVES=2
run graph_waves2 OS2S_GRAPH_DIST=1 OS2S_CONV_WAVES=2
run ctas8 NCCL_MAX_CTAS=8
run ctas32 NCCL_MAX_CTAS=32
run graph_ctas8 OS2S_GRAPH_DIST=1 NCCL_MAX_CTAS=8
