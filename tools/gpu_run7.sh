mkdir -p gpurun_out
for i in 1 2; do
python tools/profile_mlp.py tc_f16 8 | tail -4
python tools/profile_mlp.py tc_f16 8 comp | tail -4
done
python -m pytest tests/test_next_rows.py tests/test_gpu_edge_cases.py tests/test_gpu_parity.py -m gpu -q -k "hypothesis or space_carving" 2>&1 | tail -3
python tools/bench_perray.py 307200 2>&1 | tail -4
