#!/bin/bash
# bench.py at N ranks in several configurations: "<contexts>[:nccl]" ...   (gpurun --gpus N -- 'bash tools/gpu_multi_ab.sh N tag 2 3 2:nccl')
set -u
n=$1; tag=$2; shift; shift
out=gpurun_out/$tag
mkdir -p "$out"
for spec in "$@"; do
    c=${spec%%:*}
    env=""
    [[ "$spec" == *:nccl ]] && env="DSM_GATHER_NCCL=1"
    echo "== contexts=$c $env" | tee -a "$out/summary.txt"
    env $env DSM_BENCH_NO_EXTRAS=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port 29517 \
        bench.py --gpus "$n" --steps 30 --warmup 5 --contexts "$c" > "$out/bench_$spec.json" 2> "$out/bench_$spec.err"
    python - "$out/bench_$spec.json" <<'PY' | tee -a "$out/summary.txt"
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("value %.0f ms_per_step %.4f e2e %.0f (%.4f ms) parity %s" % (j["value"], j["ms_per_step"], j["e2e"]["value"], j["e2e"]["ms_per_step"], j["parity"]["ok"]))
except Exception as e:
    print("FAILED", e)
PY
done
