mkdir -p gpurun_out
timeout 300 python __graft_entry__.py > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
tail -4 gpurun_out/smoke.log
for c in 2 3 4 5; do timeout 300 python bench.py --gpus 1 --config $c --steps 10 --warmup 3 > gpurun_out/bench_c$c.json 2> gpurun_out/bench_c$c.err; echo "c$c rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_c$c.json').read().strip().splitlines()[-1]); print($c, d['value'], d['e2e']['value'], d.get('phase_ms'))
except Exception as e: print('ERR', e)
PY
done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/gpu_tests.log
