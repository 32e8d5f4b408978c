mkdir -p gpurun_out
timeout 120 python tools/profile_mlp.py tc_f16 6 | tail -3; echo "pp2 plain exit $?"
timeout 120 python tools/profile_mlp.py tc_f16 6 comp | tail -3; echo "pp2 comp exit $?"
SCADE_TC_PP2=0 timeout 120 python tools/profile_mlp.py tc_f16 6 | tail -3
SCADE_TC_PP2=0 timeout 120 python tools/profile_mlp.py tc_f16 6 comp | tail -3
timeout 120 python tools/profile_mlp.py tc_f16 6 | tail -2
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02m_gputests.log 2>&1; echo "pytest exit $?"
tail -5 gpurun_out/r02m_gputests.log
timeout 300 python bench.py --no-cpu-baseline --no-train | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('render', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['ms_per_launch'], 'e2e', d['e2e']['value'])"
