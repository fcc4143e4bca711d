#!/bin/bash
# Multi-GPU check: the two-rank NCCL gather test and bench.py at N ranks (as the driver launches it).
# Usage: gpurun --gpus N --timeout 900 -- 'bash tools/gpu_multi.sh N <tag>'
set -u
n=${1:-2}
tag=${2:-multi$n}
out=gpurun_out/$tag
mkdir -p "$out"
nvidia-smi --query-gpu=index,name --format=csv,noheader | tee "$out/summary.txt"
echo "== NCCL gather tests" | tee -a "$out/summary.txt"
[ "${SKIPTESTS:-0}" = 1 ] || timeout 300 python -m pytest tests/test_gpu_comm.py -m gpu -q 2>&1 | tail -5 | tee -a "$out/summary.txt"
echo "== bench at N=$n" | tee -a "$out/summary.txt"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus "$n" --steps 20 --warmup 5 > "$out/bench_n$n.json" 2> "$out/bench_n$n.err"
tail -c 1800 "$out/bench_n$n.json" | tee -a "$out/summary.txt"
tail -5 "$out/bench_n$n.err"
[ "${REFARM:-0}" = 1 ] || exit 0
echo "== reference arm under torchrun (rank 0 alone works)" | tee -a "$out/summary.txt"
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port 29518 \
    bench.py --impl reference --gpus "$n" --steps 2 --warmup 1 > "$out/ref_n$n.json" 2> "$out/ref_n$n.err"
echo "rc=$?" | tee -a "$out/summary.txt"
tail -c 600 "$out/ref_n$n.json" | tee -a "$out/summary.txt"
