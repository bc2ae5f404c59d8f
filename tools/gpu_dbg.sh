mkdir -p gpurun_out/r02c
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02c/memcheck_smoke.log 2>&1
echo "rc=$?"; grep -v "^=========     at\|^=========     Host\|^=========         in" gpurun_out/r02c/memcheck_smoke.log | head -60
