# N = 2 diagnosis: where do the milliseconds of the data-parallel step go?
export OMP_NUM_THREADS=8
mkdir -p gpurun_out
rm -f gpurun_out/r02_n2diag.jsonl
run() { tag=$1; shift; env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 12 --warmup 4 --no_cpu_baseline > gpurun_out/r02_n2d_$tag.log 2>&1; grep '"metric"' gpurun_out/r02_n2d_$tag.log | tail -1 | python -c "
import sys, json
l=[x for x in sys.stdin if x.startswith('{')]
d=json.loads(l[-1]) if l else {}
r=d.get('roofline') or {}
print(json.dumps({'tag': '$tag', 'value': d.get('value'), 'ms': d.get('ms_per_step'), 'e2e_ms': (d.get('e2e') or {}).get('ms_per_step'), 'conv_tflops': r.get('achieved'), 'sm_mhz': (d.get('clocks') or {}).get('sm_mhz')}))
" | tee -a gpurun_out/r02_n2diag.jsonl; }
timeout 300 python bench.py --steps 12 --warmup 4 --no_cpu_baseline > gpurun_out/r02_n1_ref2.log 2>&1; grep '"metric"' gpurun_out/r02_n1_ref2.log | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print(json.dumps({'tag':'n1','value':d['value'],'ms':d['ms_per_step'],'e2e_ms':d['e2e']['ms_per_step'],'conv_tflops':d['roofline']['achieved']}))" | tee -a gpurun_out/r02_n2diag.jsonl
run nocomm OS2S_NO_COMM=1
run graph OS2S_GRAPH_DIST=1
run graph_b32 OS2S_GRAPH_DIST=1 OS2S_BUCKET_MB=32
run graph_b512 OS2S_GRAPH_DIST=1 OS2S_BUCKET_MB=512
run graph_simple OS2S_GRAPH_DIST=1 NCCL_PROTO=Simple
run graph_ctas4 OS2S_GRAPH_DIST=1 NCCL_MAX_CTAS=4
run graph_nvls OS2S_GRAPH_DIST=1 NCCL_ALGO=NVLS
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 --no_cpu_baseline 2>&1 | grep -i "nvls\|algo\|channels\|Connected all\|via" | head -30 > gpurun_out/r02_n2_nccl_info.log
tail -12 gpurun_out/r02_n2_nccl_info.log
