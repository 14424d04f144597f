# N = $NG: peer-memory exchange replayed from the step's CUDA graph vs launched from the plan (same box)
NG=${NG:-2}
export OMP_NUM_THREADS=8
mkdir -p gpurun_out
rm -f gpurun_out/r02_peer_graph_n${NG}.jsonl
run() { tag=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 16 --warmup 5 --no_cpu_baseline > gpurun_out/r02_peer_graph_n${NG}_$tag.log 2>&1; grep '"metric"' gpurun_out/r02_peer_graph_n${NG}_$tag.log | tail -1 | python -c "
import sys, json
l=[x for x in sys.stdin if x.startswith('{')]
d=json.loads(l[-1]) if l else {}
r=d.get('roofline') or {}
print(json.dumps({'tag': '$tag', 'n_gpus': d.get('n_gpus'), 'value': d.get('value'), 'ms': d.get('ms_per_step'), 'e2e_ms': (d.get('e2e') or {}).get('ms_per_step'), 'conv_tflops': r.get('achieved'), 'exchange': (d.get('grad_exchange') or '')[:11], 'loss': d.get('loss')}))
" | tee -a gpurun_out/r02_peer_graph_n${NG}.jsonl; tail -2 gpurun_out/r02_peer_graph_n${NG}_$tag.log | cut -c1-200; }
run graph OS2S_PEER_GRAPH=1
run plan OS2S_PEER_GRAPH=0
