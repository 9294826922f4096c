mkdir -p gpurun_out
timeout 60 python __graft_entry__.py > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; grep "\[smoke\]\|\[build\]" gpurun_out/smoke.log
for c in 2 3 4 5; do timeout 40 python bench.py --gpus 1 --config $c --steps 10 --warmup 3 > gpurun_out/bench_ours_c${c}_n1.json 2> gpurun_out/bench_ours_c${c}_n1.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_ours_c${c}_n1.json').read().strip().splitlines()[-1]); print($c, round(d['value'],2), round(d['e2e']['value'],2), d.get('phase_ms'), d['exchange_aggregate']['roofline_frac'], d.get('gpu_launches'), d['clocks']['sm_mhz'])
except Exception as e: print('ERR', e)
PY
done
timeout 40 python bench.py --gpus 1 --config 2 --impl nccl --steps 10 --warmup 3 > gpurun_out/bench_nccl_c2_n1.json 2>/dev/null; tail -c 400 gpurun_out/bench_nccl_c2_n1.json | head -c 200
