mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02j_gputests.log 2>&1; echo "pytest exit $?"
tail -4 gpurun_out/r02j_gputests.log
for r in 4096 512; do timeout 300 python bench.py --workload train --steps 30 --warmup 5 --train-rays $r | cut -c1-250; done
ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 110 --csv --log-file gpurun_out/r02j_train512_launches.csv python bench.py --workload train --steps 3 --warmup 3 --train-graph 0 --train-rays 512 > /dev/null 2>&1
