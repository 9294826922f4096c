# usage: bash scripts/gpu_final_multi.sh N "<impls for c2>" "<impls for c3..c5>" [check]
N=${1:-8}; I2=${2:-"ours nccl"}; IX=${3:-"ours"}; CHECK=${4:-}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ -n "$CHECK" ]; then
  timeout 300 $TR --master-port 29533 scripts/mp_check.py > gpurun_out/mp_check_${N}gpu.txt 2>&1; echo "mp_check rc=$?"; grep -c "^OK" gpurun_out/mp_check_${N}gpu.txt; tail -1 gpurun_out/mp_check_${N}gpu.txt
fi
port=29540
for c in 2 3 4 5; do
  if [ $c = 2 ]; then IM="$I2"; else IM="$IX"; fi
  for impl in $IM; do
    port=$((port+1))
    timeout 300 $TR --master-port $port bench.py --gpus $N --config $c --impl $impl --steps 10 --warmup 3 > gpurun_out/bench_${impl}_c${c}_n${N}.json 2> gpurun_out/bench_${impl}_c${c}_n${N}.err; rc=$?
    python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_${impl}_c${c}_n${N}.json').read().strip().splitlines()[-1]); print('$impl', $c, $N, round(d['value'],2), round(d['e2e']['value'],2), d.get('phase_ms'), {k: d['exchange_aggregate'][k] for k in ('ms_per_round','roofline_frac')})
except Exception as e: print('ERR $impl $c rc=$rc', e)
PY
  done
done
timeout 300 $TR --master-port 29599 scripts/p2p_bench.py > gpurun_out/p2p_bench_${N}gpu.log 2>&1; echo "p2p rc=$?"; grep kernel gpurun_out/p2p_bench_${N}gpu.log | python -c "
import sys, json
for l in sys.stdin:
    try:
        r = json.loads(l); print(r['kernel'], r['ms'], r['nvlink_GBps'], r['frac_of_measured_nvlink'])
    except Exception: pass"
