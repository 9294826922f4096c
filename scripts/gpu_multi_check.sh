# usage: bash scripts/gpu_multi_check.sh N   (N GPUs: correctness check + bench arms for configs 2..5)
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29533 scripts/mp_check.py > gpurun_out/mp_check_${N}gpu.txt 2>&1; echo "mp_check rc=$?"; tail -2 gpurun_out/mp_check_${N}gpu.txt
port=29540
for c in 2 3 4 5; do
  for impl in ours nccl; do
    port=$((port+1))
    timeout 300 $TR --master-port $port bench.py --gpus $N --config $c --impl $impl --steps 10 --warmup 3 > gpurun_out/bench_${impl}_c${c}_n${N}.json 2> gpurun_out/bench_${impl}_c${c}_n${N}.err; rc=$?
    python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_${impl}_c${c}_n${N}.json').read().strip().splitlines()[-1]); print('$impl', $c, $N, round(d['value'],2), round(d['e2e']['value'],2), d.get('phase_ms'), d.get('exchange_aggregate'))
except Exception as e: print('ERR $impl $c rc=$rc', e)
PY
  done
done
timeout 300 $TR --master-port 29599 scripts/p2p_bench.py > gpurun_out/p2p_bench_${N}gpu.log 2>&1; echo "p2p rc=$?"; grep -c kernel gpurun_out/p2p_bench_${N}gpu.log
