mkdir -p gpurun_out/r02g
( timeout 900 python -m pytest tests/test_gpu_ohem.py tests/test_gpu_parity_at_size.py tests/test_gpu_fullsize.py -q ) > gpurun_out/r02g/pytest.log 2>&1; tail -3 gpurun_out/r02g/pytest.log
timeout 600 python bench.py --no-cpu-baseline --no-aux-workload --logits confident > gpurun_out/r02g/bench_confident.json 2> gpurun_out/r02g/bench_confident.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02g/bench_confident.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["ohem"], d["kernels"]["mdseg_ohem_select"], d["kernels"]["group_A_loss_fwd_select_bwd"])
PY
