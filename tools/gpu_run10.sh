mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02i_gputests.log 2>&1; echo "pytest exit $?"
tail -4 gpurun_out/r02i_gputests.log
timeout 300 python bench.py --workload train --steps 30 --warmup 5 | tee gpurun_out/r02i_train_1gpu.json | cut -c1-330
