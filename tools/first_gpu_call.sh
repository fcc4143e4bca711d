#!/bin/bash
# One gpurun call that checks everything written without hardware at the end of round 1 and A/B-times it.
# Usage (from the repo root):  /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/first_gpu_call.sh'
# Outputs land in gpurun_out/first_call/ (merged back by gpurun).  Each step has its own timeout so that a hang in an
# experimental kernel cannot eat the whole call.
set -u
out=gpurun_out/first_call
mkdir -p "$out"
echo "== default suite" | tee "$out/summary.txt"
timeout 240 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee -a "$out/summary.txt"
echo "== code written without hardware (chunked stream, C++ resident helper, reference node over the product)" | tee -a "$out/summary.txt"
DSM_TEST_UNVERIFIED=1 timeout 240 python -m pytest tests/test_gpu_resident.py tests/test_cpp_adapter.py tests/test_refmap.py -m gpu -q 2>&1 | tail -15 | tee -a "$out/summary.txt"
echo "== variants, one mask at a time (a hang or a fault in one does not hide the others)" | tee -a "$out/summary.txt"
for mask in 1 2 4 8 16 32 64 128; do
  echo "-- mask $mask" | tee -a "$out/summary.txt"
  DSM_EXPERIMENTAL_VARIANTS=$mask timeout 180 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3 | tee -a "$out/summary.txt"
done
DSM_TEST_VARIANTS=1 timeout 300 python -m pytest tests/test_gpu_variants.py -m gpu -q 2>&1 | tail -8 | tee -a "$out/summary.txt"
echo "== A/B timing" | tee -a "$out/summary.txt"
timeout 300 python tools/ab_variants.py 32 30 0 1 2 4 8 16 32 64 128 255 > "$out/ab_variants.jsonl" 2> "$out/ab_variants.err"
cat "$out/ab_variants.jsonl" | tee -a "$out/summary.txt"
echo "== bench with the optional legs" | tee -a "$out/summary.txt"
DSM_BENCH_STREAM_CHUNK=8 DSM_BENCH_NODE=1 timeout 420 python bench.py --steps 20 --warmup 3 > "$out/bench_optional.json" 2> "$out/bench_optional.err"
tail -c 3000 "$out/bench_optional.json" | tee -a "$out/summary.txt"
