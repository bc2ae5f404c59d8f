#!/bin/bash
# Short GPU check: smoke, GPU tests, default bench line, fused-kernel trace.  Usage: bash tools/gpu_quick.sh <tag> [pytest -k expr]
TAG=${1:-q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== smoke"; timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log; tail -3 $OUT/smoke.log
echo "== pytest gpu"; if [ -n "$2" ]; then timeout 1200 python -m pytest tests -x -q -m gpu --timeout 240 -k "$2" > $OUT/pytest_gpu.log 2>&1; else timeout 1200 python -m pytest tests -x -q -m gpu --timeout 240 > $OUT/pytest_gpu.log 2>&1; fi; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log; tail -15 $OUT/pytest_gpu.log
echo "== bench (default)"; timeout 420 python bench.py --steps 20 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json; tail -5 $OUT/bench.err
echo "== fused-kernel clock64 trace"; timeout 120 python tools/fused_trace.py 2>&1 | grep -v "MUSIC DOA" > $OUT/fused_trace.txt; cat $OUT/fused_trace.txt
