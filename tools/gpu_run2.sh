set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_metric_parity.py -x -q -k "x3" -s > gpurun_out/r02b_x3_tests.log 2>&1; echo "pytest exit $?"
grep -v "^$" gpurun_out/r02b_x3_tests.log | tail -60
timeout 300 python bench.py --precision tc_f16x3 --no-train --no-cpu-baseline > gpurun_out/r02b_bench_x3.json 2> gpurun_out/r02b_bench_x3.err; echo "bench exit $?"
cat gpurun_out/r02b_bench_x3.json; tail -5 gpurun_out/r02b_bench_x3.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8
