#!/bin/bash
# Round-end evidence run on ONE GPU: smoke, GPU tests, bench (default line + reference arm), fused-kernel trace, ncu launch
# list and ncu --set full captures of the dominant kernels.  Usage (repo root on the GPU box): bash tools/gpu_final.sh <tag> <commit>
TAG=${1:-r02}; COMMIT=${2:-unknown}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "$COMMIT" > $OUT/commit.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== smoke"; timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log; tail -2 $OUT/smoke.log
echo "== pytest gpu"; timeout 900 python -m pytest tests -x -q -m gpu --timeout 300 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log; tail -4 $OUT/pytest_gpu.log
echo "== bench (default)"; timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cut -c1-1200 $OUT/bench.json; tail -3 $OUT/bench.err
echo "== bench reference arm"; timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; cut -c1-600 $OUT/bench_ref.json
echo "== fused-kernel clock64 trace"; timeout 120 python tools/fused_trace.py 2>&1 | grep -v "MUSIC DOA" > $OUT/fused_trace.txt; cat $OUT/fused_trace.txt
echo "== ncu launch list"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"cov|eig|scan|topn|prep_table|fused|gather" -c 60 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-next-rows --no-other-configs > $OUT/ncu_launch_bench.log 2>&1; echo "ncu rc=$?"
echo "== ncu full: fused M=4, fused M=8, covN<16> + eig_coop<16>"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"music4_fused" -s 3 -c 1 -f -o $OUT/prof_fused4 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-next-rows --no-other-configs > $OUT/ncu_full_fused4.log 2>&1; echo "ncu fused4 rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"music8_fused" -s 3 -c 1 -f -o $OUT/prof_fused8 python bench.py --config 4 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-next-rows > $OUT/ncu_full_fused8.log 2>&1; echo "ncu fused8 rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"covN|eig_coop" -s 6 -c 2 -f -o $OUT/prof_c5 python bench.py --config 5 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-next-rows > $OUT/ncu_full_c5.log 2>&1; echo "ncu c5 rc=$?"
ls -la $OUT
