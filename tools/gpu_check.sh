#!/bin/bash
# One gpurun call: parity suite on the default (tile) schedule, then on the round-1 schedule, the gated tests, a short bench.
# Usage: gpurun --timeout 1200 -- 'bash tools/gpu_check.sh [tag]'.  Outputs in gpurun_out/<tag>/.
set -u
tag=${1:-check}
out=gpurun_out/$tag
mkdir -p "$out"
echo "== kernel-by-kernel prefixes (tile schedule)" | tee "$out/summary.txt"
DSM_GRAPHS=0 timeout 120 python tools/debug_one.py 640 480 2>&1 | tail -14 | tee -a "$out/summary.txt"
if ! grep -q "label mismatches: 0" "$out/summary.txt"; then echo "tile schedule broken: stopping here" | tee -a "$out/summary.txt"; exit 1; fi
echo "== default (tile) schedule" | tee -a "$out/summary.txt"
timeout 420 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 | tee -a "$out/summary.txt"
if [ "${QUICK:-0}" = "1" ]; then
  echo "== bench (tile schedule)" | tee -a "$out/summary.txt"
  timeout 420 python bench.py --steps 20 --warmup 5 --no-cpu > "$out/bench.json" 2> "$out/bench.err"
  tail -c 2500 "$out/bench.json" | tee -a "$out/summary.txt"
  exit 0
fi
echo "== round-1 schedule (DSM_EXPERIMENTAL_VARIANTS=256)" | tee -a "$out/summary.txt"
DSM_EXPERIMENTAL_VARIANTS=256 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -5 | tee -a "$out/summary.txt"
echo "== gated tests (DSM_TEST_UNVERIFIED=1)" | tee -a "$out/summary.txt"
DSM_TEST_UNVERIFIED=1 timeout 300 python -m pytest tests/test_gpu_resident.py tests/test_cpp_adapter.py tests/test_refmap.py -m gpu -q 2>&1 | tail -25 | tee -a "$out/summary.txt"
echo "== shared-kernel variants on the tile schedule (16 = init multiblock, 64 = seed init wide)" | tee -a "$out/summary.txt"
for mask in 16 64; do
  DSM_EXPERIMENTAL_VARIANTS=$mask timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3 | tee -a "$out/summary.txt"
done
echo "== bench (tile schedule)" | tee -a "$out/summary.txt"
timeout 420 python bench.py --steps 20 --warmup 5 > "$out/bench.json" 2> "$out/bench.err"
tail -c 2500 "$out/bench.json" | tee -a "$out/summary.txt"
echo "== bench (round-1 schedule)" | tee -a "$out/summary.txt"
DSM_EXPERIMENTAL_VARIANTS=256 timeout 420 python bench.py --steps 20 --warmup 5 > "$out/bench_legacy.json" 2> "$out/bench_legacy.err"
tail -c 1500 "$out/bench_legacy.json" | tee -a "$out/summary.txt"
