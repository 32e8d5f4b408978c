set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02g_gputests.log 2>&1; echo "pytest exit $?"
tail -5 gpurun_out/r02g_gputests.log
timeout 600 python bench.py > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err; echo "bench exit $?"; tail -2 gpurun_out/r02g_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02g_bench.json"))
print("render", d["value"], d["ms_per_step"], d["roofline"]["frac"], "e2e", d["e2e"]["value"], d["e2e"]["eager_value"])
print("train", d["train"]["value"], d["train"]["ms_per_step"], d["train"]["gpu_launches"])
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
PY
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 120 --csv --log-file gpurun_out/r02g_train_launches.csv python bench.py --workload train --steps 3 --warmup 3 --train-graph 0 > /dev/null 2>&1
