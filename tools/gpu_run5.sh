set -x
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/r02e_dist_check.log 2>&1; echo "dist_check exit $?"
grep -v "^$" gpurun_out/r02e_dist_check.log | tail -12
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02e_bench_2gpu.json 2> gpurun_out/r02e_bench_2gpu.err; echo "bench exit $?"
python - <<'PY'
import json
for l in open("gpurun_out/r02e_bench_2gpu.json"):
    if l.startswith("{"):
        d=json.loads(l)
        print("render", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"])
        print("train", d["train"]["value"], d["train"]["ms_per_step"], d["train"]["gpu_launches"], d["train"]["config"]["collective"])
        print("cpu", d.get("cpu_baseline"))
PY
tail -5 gpurun_out/r02e_bench_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 | tail -1 | cut -c1-400
