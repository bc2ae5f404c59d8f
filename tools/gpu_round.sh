#!/bin/bash
# One gpurun call: tests, smoke, bench, ncu launch list + full capture of the top kernels.
# Usage (from the repo root on the GPU box): bash tools/gpu_round.sh [tag] [quick]
TAG=${1:-r01}
MODE=${2:-full}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== smoke" ; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log; tail -3 $OUT/smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log; tail -15 $OUT/pytest_gpu.log
for var in "MUSIC_B200_FUSED=0" "MUSIC_B200_FUSED=1"; do
  tag=$(echo "$var" | tr ' =' '__')
  echo "== bench $var"; env $var timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-next-rows > $OUT/bench_$tag.json 2> $OUT/bench_$tag.err; python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_$tag.json")); print("$var", "value=%.3e"%d["value"], "ms/step=%.4f"%d["ms_per_step"], {k:(round(v,4) if isinstance(v,float) else v) for k,v in d["stages"].items()}, "roof=%.3f whole=%.3f"%(d["roofline"]["frac"], d["roofline"]["whole_step_frac"]), d["clocks"])
except Exception as e:
    print("$var FAILED", e); print(open("$OUT/bench_$tag.err").read()[-2000:])
PY
done
echo "== bench (default)"; timeout 900 python bench.py --steps 20 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json; tail -5 $OUT/bench.err
if [ "$MODE" = "full" ]; then
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; cat $OUT/bench_ref.json
fi
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"cov|eig_kernel|scan|topn_kernel|prep_table|fused" -c 60 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-next-rows > $OUT/ncu_launch_bench.log 2>&1; echo "ncu rc=$?"
echo "== ncu full (fused kernel, then the three unfused kernels)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"fused" -s 3 -c 1 -o $OUT/prof_fused python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-next-rows > $OUT/ncu_full_fused.log 2>&1; echo "ncu fused rc=$?"
MUSIC_B200_FUSED=0 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"cov|scan|eig_kernel" -s 9 -c 3 -o $OUT/prof python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-next-rows > $OUT/ncu_full_bench.log 2>&1; echo "ncu full rc=$?"
if [ "$MODE" = "full" ]; then
echo "== configs 3-5 (parity-test shapes; reported, not the headline)"
for c in 3 4 5; do timeout 400 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-next-rows > $OUT/bench_c$c.json 2> $OUT/bench_c$c.err; python -c "
import json
d=json.load(open('$OUT/bench_c$c.json')); print('config $c', '%.3e'%d['value'], 'ms/step=%.4f'%d['ms_per_step'], {k:(round(v,4) if isinstance(v,float) else v) for k,v in d['stages'].items()}, 'cov frac=%.3f'%d['roofline']['unfused_cov_kernel_frac'])"; done
echo "== ncu full (TMA-tiled covariance, M = 8 and M = 16)"
for c in 4 5; do timeout 600 ncu --set full --clock-control none --import-source on -k regex:covN -s 3 -c 1 -f -o $OUT/prof_covn_c$c python bench.py --config $c --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-next-rows > $OUT/ncu_covn_c$c.log 2>&1; echo "ncu covN c$c rc=$?"; done
fi
echo "== fused-kernel clock64 trace"; timeout 300 python tools/fused_trace.py 2>&1 | grep -v "MUSIC DOA" > $OUT/fused_trace.txt; cat $OUT/fused_trace.txt
if [ -x tools/microbench ]; then (cd tools && timeout 300 ./microbench) > $OUT/microbench.log 2>&1; head -8 $OUT/microbench.log; fi
ls -la $OUT
