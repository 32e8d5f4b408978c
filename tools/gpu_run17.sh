mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02n_gputests.log 2>&1; echo "pytest exit $?"
tail -4 gpurun_out/r02n_gputests.log
timeout 600 python bench.py > gpurun_out/r02n_bench.json 2> gpurun_out/r02n_bench.err; echo "bench exit $?"; tail -3 gpurun_out/r02n_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02n_bench.json"))
print("render", d["value"], d["ms_per_step"], d["roofline"]["frac"], "e2e", d["e2e"]["value"], "pipelined", d["e2e"]["pipelined_value"], "eager", d["e2e"]["eager_value"])
print("train", d["train"]["value"], d["train"]["ms_per_step"], d["train"]["gpu_launches"])
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
PY
