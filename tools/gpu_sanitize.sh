#!/bin/bash
# compute-sanitizer over a small subset of the GPU tests (memcheck, then racecheck for the shared-memory tile kernels).
# Usage: gpurun --timeout 1500 -- 'bash tools/gpu_sanitize.sh [tag]'
set -u
tag=${1:-sanitize}
out=gpurun_out/$tag
mkdir -p "$out"
T="tests/test_gpu_parity.py::test_single_frame_vga_identity tests/test_gpu_parity.py::test_concurrent_sub_batches_do_not_change_results tests/test_gpu_comm.py::test_single_rank_gather_equals_batch_download tests/test_gpu_resident.py"
timeout 700 compute-sanitizer --tool memcheck --error-exitcode 0 python -m pytest $T -m gpu -q -x 2>&1 | grep -v "^$" | tail -25 > "$out/memcheck.txt"
tail -4 "$out/memcheck.txt"
timeout 600 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 0 python -m pytest tests/test_gpu_parity.py::test_single_frame_vga_identity "tests/test_gpu_parity.py::test_adversarial_small_frames" -m gpu -q -x 2>&1 | grep -v "^$" | tail -25 > "$out/racecheck.txt"
tail -4 "$out/racecheck.txt"
