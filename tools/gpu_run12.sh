mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 120 --csv --log-file gpurun_out/r02j_train512_launches.csv python bench.py --workload train --steps 3 --warmup 3 --train-graph 0 --train-rays 512 > /dev/null 2>&1
grep -c nerf_mlp gpurun_out/r02j_train512_launches.csv
