mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29811 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02p_bench_2gpu.json 2> gpurun_out/r02p_bench_2gpu.err; echo "bench exit $?"
tail -3 gpurun_out/r02p_bench_2gpu.err
wc -l gpurun_out/r02p_bench_2gpu.json
python - <<'PY'
import json
for l in open("gpurun_out/r02p_bench_2gpu.json"):
    if l.startswith("{"):
        d=json.loads(l)
        print("render", d["value"], d["ms_per_step"], d["roofline"]["frac"], "e2e", d["e2e"]["value"], d["e2e"]["pipelined_value"])
        print("train", d["train"]["value"], d["train"]["ms_per_step"])
        print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], d["clocks"])
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29812 bench.py --gpus 2 --impl reference --steps 20 --warmup 5 2>/dev/null | cut -c1-200
