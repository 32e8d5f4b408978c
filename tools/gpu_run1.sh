set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
nproc
timeout 900 python -m pytest tests -m gpu -x -q -k "not x3" -s > gpurun_out/r02a_gputests.log 2>&1; echo "pytest exit $?"
tail -5 gpurun_out/r02a_gputests.log
grep -A16 "parity," gpurun_out/r02a_gputests.log | head -120
timeout 600 python bench.py > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench exit $?"
cat gpurun_out/r02a_bench.json
tail -5 gpurun_out/r02a_bench.err
