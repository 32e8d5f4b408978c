mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02k_gputests.log 2>&1; echo "pytest exit $?"
tail -4 gpurun_out/r02k_gputests.log
for r in 4096 512; do timeout 300 python bench.py --workload train --steps 30 --warmup 5 --train-rays $r | cut -c1-250; done
