#!/bin/bash
# One gpurun call: tests, smoke, bench, microbench, ncu launch list + full capture of the top kernel.
# Usage (from the repo root on the GPU box): bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== smoke" ; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log; tail -3 $OUT/smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log; tail -15 $OUT/pytest_gpu.log
echo "== microbench"; (cd tools && timeout 300 ./microbench) > $OUT/microbench.log 2>&1; tail -30 $OUT/microbench.log
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json; tail -5 $OUT/bench.err
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; cat $OUT/bench_ref.json
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/ncu_launch_bench.log 2>&1; echo "ncu rc=$?"
echo "== ncu full (top kernels)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"cov_tile|scan_kernel|eig_kernel" -s 9 -c 3 -o $OUT/prof python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/ncu_full_bench.log 2>&1; echo "ncu full rc=$?"
ls -la $OUT
