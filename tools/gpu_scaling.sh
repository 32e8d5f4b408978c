#!/bin/bash
# 1 / 2 / 4 / 8-GPU numbers of ONE box (run with `gpurun --gpus 8`): the default bench line (weak-scaling render + strong-scaling
# config-3 train record), the full-frame workloads, and the multi-GPU correctness check.
mkdir -p gpurun_out
run() { n=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) "$@"; }
for n in 8 4 2 1; do run $n bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | grep '^{' > gpurun_out/scale_${n}gpu.json; done
run 8 bench.py --gpus 8 --workload image --steps 10 --warmup 3 2>/dev/null | grep '^{' > gpurun_out/image_8gpu.json
run 8 bench.py --gpus 8 --workload video --steps 10 --warmup 3 2>/dev/null | grep '^{' > gpurun_out/video_8gpu.json
run 8 tools/dist_check.py > gpurun_out/dist_check_8gpu.log 2>&1; echo "dist_check exit $?"
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/scale_*gpu.json")) + ["gpurun_out/image_8gpu.json", "gpurun_out/video_8gpu.json"]:
    d = json.loads(open(f).read()); t = d.get("train") or {}
    print(f, "%.4g rays/s  %.4g ms" % (d["value"], d["ms_per_step"]), "| train %s rays/s %s ms" % (t.get("value"), t.get("ms_per_step")))
PY
