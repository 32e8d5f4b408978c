mkdir -p gpurun_out
export NCCL_DEBUG=WARN
run() { n=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n "$@"; }
run 8 --steps 20 --warmup 5 > gpurun_out/r02h_bench_8gpu.json 2> gpurun_out/r02h_bench_8gpu.err; echo "8gpu exit $?"
run 4 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02h_bench_4gpu.json 2> gpurun_out/r02h_bench_4gpu.err; echo "4gpu exit $?"
run 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02h_bench_2gpu.json 2> gpurun_out/r02h_bench_2gpu.err; echo "2gpu exit $?"
run 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02h_bench_1gpu.json 2> gpurun_out/r02h_bench_1gpu.err; echo "1gpu exit $?"
run 8 --workload image --steps 10 --warmup 3 > gpurun_out/r02h_image_8gpu.json 2> gpurun_out/r02h_image_8gpu.err; echo "image exit $?"
run 8 --workload video --steps 10 --warmup 3 > gpurun_out/r02h_video_8gpu.json 2> gpurun_out/r02h_video_8gpu.err; echo "video exit $?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02h_*.json")):
    for l in open(f):
        if l.startswith("{"):
            d=json.loads(l); t=d.get("train") or {}
            print(f.split("/")[-1], "value %.4g ms %.4g e2e %s | train %s rays/s %s ms | psnr %s | cpu %s" % (d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value"), t.get("value"), t.get("ms_per_step"), d.get("psnr_vs_oracle_db"), (d.get("cpu_baseline") or {}).get("value")))
PY
tail -3 gpurun_out/r02h_bench_8gpu.err
