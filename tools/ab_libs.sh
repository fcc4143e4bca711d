#!/bin/bash
# A/B of prebuilt variant libraries (ab/lib_<name>.so, git-ignored) on one box: each is copied over the product
# library in the box's scratch copy of the repo, checked with the parity tests and timed with the resident bench.
# Usage: gpurun -- 'bash tools/ab_libs.sh <tag> name[:sub_batches[:contexts]] ...'   ("base" = the library as built)
set -u
tag=$1; shift
out=gpurun_out/$tag
mkdir -p "$out"
L=densesurfelmapping_b200/libdsm_b200.so
cp $L /tmp/lib_base.so
last=""
for spec in "$@"; do
    IFS=: read -r v sb nc <<< "$spec"
    sb=${sb:-2}; nc=${nc:-2}
    if [ "$v" = base ]; then cp /tmp/lib_base.so $L; else cp ab/lib_$v.so $L; fi
    echo "== $v sub_batches=$sb contexts=$nc" | tee -a "$out/summary.txt"
    if [ "$v" != "$last" ]; then
        timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_resident.py -m gpu -x -q 2>&1 | tail -1 | tee -a "$out/summary.txt"
        last=$v
    fi
    for rep in 1 2; do
        DSM_BENCH_NO_EXTRAS=1 timeout 300 python bench.py --no-cpu --steps 30 --warmup 5 --sub-batches $sb --contexts $nc 2> "$out/$v.$sb.$nc.err" | tail -1 > "$out/$v.$sb.$nc.$rep.json"
        python - "$out/$v.$sb.$nc.$rep.json" <<'PY' | tee -a "$out/summary.txt"
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read())
    print("ms_per_step %.4f  e2e %.4f  parity %s  " % (j["ms_per_step"], j["e2e"]["ms_per_step"], j["parity"]["ok"]) + " ".join("%s=%.1f" % (k, 1e3 * v) for k, v in j["kernel_ms_per_launch"].items()))
except Exception as e:
    print("FAILED", e)
PY
    done
done
cp /tmp/lib_base.so $L
if [ "${EXTRAS:-0}" = 1 ]; then
    echo "== base, full bench with extras" | tee -a "$out/summary.txt"
    timeout 600 python bench.py --no-cpu > "$out/bench_full.json" 2> "$out/bench_full.err"
    python - "$out/bench_full.json" <<'PY' | tee -a "$out/summary.txt"
import json, sys
j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("ms_per_step %.4f e2e %.4f" % (j["ms_per_step"], j["e2e"]["ms_per_step"]))
for k, v in j["extras"].items():
    print(k, json.dumps(v)[:400])
PY
fi
