#!/usr/bin/env bash
# compute-sanitizer passes over the hand-written kernels: memcheck, racecheck (shared-memory hazards),
# synccheck and initcheck on the single-GPU kernel tests.  Run on a GPU box:  bash scripts/sanitize.sh [outdir]
set -u
OUT=${1:-gpurun_out/sanitizer}
mkdir -p "$OUT"
SEL='publish or weighted_gather or edge_distances or pairwise or count_sketch or filters or sgd or evidential or mobility or tail_blend'
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 --log-file "$OUT/$tool.log" \
      python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "$SEL" > "$OUT/$tool.pytest.log" 2>&1
  echo "$tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' "$OUT/$tool.log" | tail -1)"
done
# the tcgen05/TMA kernel: memcheck only (racecheck does not model async-proxy / tensor-memory traffic)
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file "$OUT/memcheck_gram.log" \
    python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "gram" > "$OUT/memcheck_gram.pytest.log" 2>&1
echo "memcheck(gram) rc=$? $(grep -E 'ERROR SUMMARY' "$OUT/memcheck_gram.log" | tail -1)"
