# N = 2 (or $NG): correctness of the peer-memory exchange vs NCCL, then same-box A/B of the training step
NG=${NG:-2}
export OMP_NUM_THREADS=8
mkdir -p gpurun_out
rm -f gpurun_out/r02_peer_n${NG}.jsonl
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533 \
    tools/peer_check.py --floats 333000000 --iters 3 --time > gpurun_out/r02_peer_check_n${NG}.log 2>&1
echo "peer_check rc=$?"; grep '^{' gpurun_out/r02_peer_check_n${NG}.log | tee -a gpurun_out/r02_peer_n${NG}.jsonl; tail -5 gpurun_out/r02_peer_check_n${NG}.log | cut -c1-300
run() { tag=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 12 --warmup 4 --no_cpu_baseline > gpurun_out/r02_peer_n${NG}_$tag.log 2>&1; grep '"metric"' gpurun_out/r02_peer_n${NG}_$tag.log | tail -1 | python -c "
import sys, json
l=[x for x in sys.stdin if x.startswith('{')]
d=json.loads(l[-1]) if l else {}
r=d.get('roofline') or {}
print(json.dumps({'tag': '$tag', 'n_gpus': d.get('n_gpus'), 'value': d.get('value'), 'ms': d.get('ms_per_step'), 'e2e_ms': (d.get('e2e') or {}).get('ms_per_step'), 'conv_tflops': r.get('achieved'), 'exchange': d.get('grad_exchange'), 'loss': d.get('loss')}))
" | tee -a gpurun_out/r02_peer_n${NG}.jsonl; tail -2 gpurun_out/r02_peer_n${NG}_$tag.log | cut -c1-200; }
run peer OS2S_GRAD_EXCHANGE=auto
run nccl OS2S_GRAD_EXCHANGE=nccl
