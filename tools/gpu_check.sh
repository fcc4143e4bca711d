#!/bin/bash
# One gpurun call: kernel-by-kernel smoke, the whole GPU parity suite, the bench (QUICK=1: bench without the CPU leg).
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
echo "== bench" | tee -a "$out/summary.txt"
timeout 600 python bench.py --steps 20 --warmup 5 > "$out/bench.json" 2> "$out/bench.err"
tail -c 4000 "$out/bench.json" | tee -a "$out/summary.txt"
