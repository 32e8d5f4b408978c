set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -s > gpurun_out/r02c_gputests.log 2>&1; echo "pytest exit $?"
grep -v "^$" gpurun_out/r02c_gputests.log | tail -40
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02c_bench.json"))
print("render", d["value"], d["roofline"]["frac"], "e2e", d["e2e"]["value"])
print("train", d["train"]["value"], d["train"]["ms_per_step"], d["train"]["gpu_launches"])
PY
tail -3 gpurun_out/r02c_bench.err
