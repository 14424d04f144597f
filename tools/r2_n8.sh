export OMP_NUM_THREADS=8
mkdir -p gpurun_out
rm -f gpurun_out/r02_n8.jsonl
run() { tag=$1; shift; env "$@" timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 12 --warmup 4 --no_cpu_baseline > gpurun_out/r02_n8_$tag.log 2>&1; grep '"metric"' gpurun_out/r02_n8_$tag.log | tail -1 | python -c "
import sys, json
l=[x for x in sys.stdin if x.startswith('{')]
d=json.loads(l[-1]) if l else {}
r=d.get('roofline') or {}
print(json.dumps({'tag': '$tag', 'value': d.get('value'), 'ms': d.get('ms_per_step'), 'e2e_ms': (d.get('e2e') or {}).get('ms_per_step'), 'conv_tflops': r.get('achieved'), 'sm_mhz': (d.get('clocks') or {}).get('sm_mhz')}))
" | tee -a gpurun_out/r02_n8.jsonl; tail -3 gpurun_out/r02_n8_$tag.log | cut -c1-160; }
timeout 300 python bench.py --steps 12 --warmup 4 --no_cpu_baseline > gpurun_out/r02_n8_n1.log 2>&1; grep '"metric"' gpurun_out/r02_n8_n1.log | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print(json.dumps({'tag':'n1','value':d['value'],'ms':d['ms_per_step'],'e2e_ms':d['e2e']['ms_per_step'],'conv_tflops':d['roofline']['achieved']}))" | tee -a gpurun_out/r02_n8.jsonl
run default A=1
run graph OS2S_GRAPH_DIST=1
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL,TUNING timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 2 --warmup 3 --no_cpu_baseline 2>&1 | grep -i "nvls\|algo\|AllReduce" | head -40 > gpurun_out/r02_n8_nccl_info.log
tail -8 gpurun_out/r02_n8_nccl_info.log | cut -c1-200
