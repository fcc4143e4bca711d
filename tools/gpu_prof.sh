#!/bin/bash
# ncu evidence for one resident 32-frame step: launch list of the whole step, then --set full of one launch of each kernel.
# Usage: gpurun --timeout 900 -- 'bash tools/gpu_prof.sh <tag>'   -> gpurun_out/<tag>/{launches.csv,full.ncu-rep}
set -u
tag=${1:-prof}
out=gpurun_out/$tag
mkdir -p "$out"
# launch list: 1 untimed warm step + 2 steps (cold-cache serialised times: shares, not absolutes)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file "$out/launches.csv" \
    python tools/prof_step.py 32 3 > "$out/launches.log" 2>&1
# full set: skip the setup batch (pool seeding: one schedule) and the first step, capture the second step's kernels
nk=${2:-12}
timeout 600 ncu --set full --clock-control none --import-source on --launch-skip ${3:-29} --launch-count $nk -f -o "$out/full" \
    python tools/prof_step.py 32 3 > "$out/full.log" 2>&1
ls -la "$out"
tail -3 "$out/full.log"
