set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_metric_parity.py tests/test_gpu_tc_train.py -m gpu -x -q -s -k "fused_compositing or joint_sharded or per_image" > gpurun_out/r02d_new.log 2>&1; echo "pytest exit $?"
grep -v "^$" gpurun_out/r02d_new.log | tail -30
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02d_gputests.log 2>&1; echo "pytest exit $?"
tail -15 gpurun_out/r02d_gputests.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02d_bench.json"))
print("render", d["value"], d["ms_per_step"], d["roofline"]["frac"], "e2e", d["e2e"]["value"], d["e2e"]["eager_value"], "launches", d["gpu_launches"])
print("train", d["train"]["value"], d["train"]["ms_per_step"], d["train"]["gpu_launches"])
PY
tail -3 gpurun_out/r02d_bench.err
