set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02f_gputests.log 2>&1; echo "pytest exit $?"
tail -6 gpurun_out/r02f_gputests.log
grep -A14 "parity," gpurun_out/r02f_gputests.log > gpurun_out/r02f_parity_report.txt
grep "mlp x3" gpurun_out/r02f_gputests.log >> gpurun_out/r02f_parity_report.txt
timeout 600 python bench.py > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; echo "bench exit $?"; tail -2 gpurun_out/r02f_bench.err
timeout 600 python bench.py --precision tc_f16x3 --no-cpu-baseline --no-train > gpurun_out/r02f_bench_x3.json 2>> gpurun_out/r02f_bench.err; echo "bench x3 exit $?"
timeout 600 python bench.py --workload render_c2 --no-cpu-baseline > gpurun_out/r02f_bench_c2.json 2>> gpurun_out/r02f_bench.err; echo "bench c2 exit $?"
timeout 600 python bench.py --workload image --steps 5 --warmup 3 > gpurun_out/r02f_bench_c4_1gpu.json 2>> gpurun_out/r02f_bench.err; echo "bench image exit $?"
timeout 900 python bench.py --workload video --steps 5 --warmup 3 > gpurun_out/r02f_bench_c5_1gpu.json 2>> gpurun_out/r02f_bench.err; echo "bench video exit $?"
python tools/bench_perray.py > gpurun_out/r02f_perray.txt 2>&1; tail -12 gpurun_out/r02f_perray.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02f_bench*.json")):
    for l in open(f):
        if l.startswith("{"):
            d=json.loads(l)
            print(f, d["value"], d["ms_per_step"], d.get("roofline",{}).get("frac"), d.get("e2e",{}).get("value"), d.get("psnr_vs_oracle_db"), (d.get("train") or {}).get("ms_per_step"))
PY
# profiles: launch list of the default step, full captures of the fused-composite kernel and the x3 kernel
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02f_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-train > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:nerf_mlp_tc_pp -s 2 -c 1 -o gpurun_out/r02f_prof_comp python tools/profile_mlp.py tc_f16 4 comp > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:nerf_mlp_tc_x3 -s 2 -c 1 -o gpurun_out/r02f_prof_x3 python tools/profile_mlp.py tc_f16x3 4 > /dev/null 2>&1
ls -la gpurun_out | tail -20
