#!/bin/bash
# One-GPU acceptance run on a B200 box (what every round-2 change went through):
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/gpu_check.sh'
# GPU parity tests, smoke, the default bench line (render + train record + reference CPU baseline), the tight mode, and the
# launch list of the step.  Outputs land in gpurun_out/.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/gputests.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/gputests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
timeout 600 python bench.py --precision tc_f16x3 --no-cpu-baseline --no-train > gpurun_out/bench_x3.json 2>> gpurun_out/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-train > /dev/null 2>&1
python - <<'PY'
import json
for f in ("gpurun_out/bench.json", "gpurun_out/bench_x3.json"):
    d = json.load(open(f))
    print(f, "%.4g rays/s, %.4g ms/step, roofline %.3f, e2e %.4g" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"]),
          "| train %.4g ms" % d["train"]["ms_per_step"] if d.get("train") else "")
PY
