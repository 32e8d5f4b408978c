mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_smoke.py > gpurun_out/r02o_memcheck.log 2>&1; echo "memcheck exit $?"
tail -16 gpurun_out/r02o_memcheck.log
