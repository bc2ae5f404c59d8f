#!/bin/bash
# Round-2 evidence run on ONE GPU: tests, bench (default + A/B variants), ncu launch list + full captures, clock64 trace.
# Usage (repo root on the GPU box): bash tools/gpu_round2.sh <tag> [notests]
TAG=${1:-r02}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== smoke"; timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log; tail -2 $OUT/smoke.log
if [ "$2" != "notests" ]; then
echo "== pytest gpu"; timeout 1500 python -m pytest tests -x -q -m gpu --timeout 300 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log; tail -8 $OUT/pytest_gpu.log
fi
echo "== bench (default)"; timeout 600 python bench.py --steps 20 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cut -c1-3000 $OUT/bench.json; tail -5 $OUT/bench.err
echo "== bench reference arm"; timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; cut -c1-1500 $OUT/bench_ref.json; tail -3 $OUT/bench_ref.err
for var in "MUSIC_B200_EIG=jacobi" "MUSIC_B200_MMA_FIN=8" "MUSIC_B200_MMA_FIN=-1" "MUSIC_B200_MMA_FIN=0" "MUSIC_B200_MMA_FIN=4" "MUSIC_B200_FUSED=0"; do
  tag=$(echo "$var" | tr ' =-' '___')
  env $var timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-next-rows --no-other-configs > $OUT/bench_$tag.json 2> $OUT/bench_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_$tag.json")); print("$var", "value=%.4e"%d["value"], "ms/step=%.4f"%d["ms_per_step"], "roof=%.3f"%d["roofline"]["frac"])
except Exception as e:
    print("$var FAILED", e); print(open("$OUT/bench_$tag.err").read()[-1500:])
PY
done
echo "== fused-kernel clock64 trace"; timeout 120 python tools/fused_trace.py 2>&1 | grep -v "MUSIC DOA" > $OUT/fused_trace.txt; cat $OUT/fused_trace.txt
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"cov|eig|scan|topn|prep_table|fused|gather" -c 80 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-next-rows > $OUT/ncu_launch_bench.log 2>&1; echo "ncu rc=$?"
echo "== ncu full: fused M=4, fused M=8, covN<16>"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"music4_fused" -s 3 -c 1 -f -o $OUT/prof_fused4 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-next-rows --no-other-configs > $OUT/ncu_full_fused4.log 2>&1; echo "ncu fused4 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"music8_fused" -s 3 -c 1 -f -o $OUT/prof_fused8 python bench.py --config 4 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-next-rows > $OUT/ncu_full_fused8.log 2>&1; echo "ncu fused8 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"covN|eig_coop" -s 6 -c 2 -f -o $OUT/prof_c5 python bench.py --config 5 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-next-rows > $OUT/ncu_full_c5.log 2>&1; echo "ncu c5 rc=$?"
ls -la $OUT
