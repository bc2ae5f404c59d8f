#!/bin/bash
# Lean perf iteration on ONE GPU: fused-kernel trace + the default bench line without the side legs, for a list of
# environment variants.  Usage: bash tools/gpu_perf.sh <tag> ["VAR=val VAR2=val" ...]
TAG=${1:-p}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== trace (default)"; timeout 120 python tools/fused_trace.py 2>&1 | grep -v "MUSIC DOA" > $OUT/fused_trace.txt; cat $OUT/fused_trace.txt
run() {
  tag=$(echo "$1" | tr ' =-' '___'); [ -z "$tag" ] && tag=default
  env $1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-next-rows --no-other-configs > $OUT/bench_$tag.json 2> $OUT/bench_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_$tag.json")); print("[$1]", "value=%.4e"%d["value"], "ms/step=%.4f"%d["ms_per_step"], "roof=%.3f"%d["roofline"]["frac"])
except Exception as e:
    print("[$1] FAILED", e); print(open("$OUT/bench_$tag.err").read()[-1500:])
PY
}
run ""
for v in "$@"; do run "$v"; done
