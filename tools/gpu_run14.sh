mkdir -p gpurun_out
run() { n=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700+n)) "$@"; }
for n in 8 4 2 1; do run $n bench.py --gpus $n --workload train --steps 40 --warmup 5 2>/dev/null | grep '^{' > gpurun_out/r02l_train_${n}gpu.json; done
run 8 bench.py --gpus 8 --workload train --steps 40 --warmup 5 --train-overlap 0 2>/dev/null | grep '^{' > gpurun_out/r02l_train_8gpu_nooverlap.json
run 8 tools/dist_check.py > gpurun_out/r02l_dist_check_8gpu.log 2>&1; echo "dist_check exit $?"; grep -v "^$\|\*\*\*\|OMP_NUM" gpurun_out/r02l_dist_check_8gpu.log | tail -6
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02l_train*.json")):
    d=json.loads(open(f).read()); print(f.split("/")[-1], "%.4g rays/s  %.4g ms/step  launches/step %.1f" % (d["value"], d["ms_per_step"], d["gpu_launches"]/d["steps"]))
PY
