# final build of the round: GPU tests, smoke, headline bench + reference arm, 16-bit dense-graph step
mkdir -p gpurun_out/r02y
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/r02y/pytest.log 2>&1; tail -4 gpurun_out/r02y/pytest.log
python __graft_entry__.py smoke > gpurun_out/r02y/smoke.log 2>&1; tail -1 gpurun_out/r02y/smoke.log
timeout 600 python bench.py > gpurun_out/r02y/bench.json 2> gpurun_out/r02y/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02y/bench_ref.json 2> gpurun_out/r02y/bench_ref.err; echo "ref rc=$?"
timeout 600 python bench.py --no-cpu-baseline --no-aux-workload --bi-graphs dense --logits-dtype bf16 > gpurun_out/r02y/bench_dense_bf16.json 2> gpurun_out/r02y/bench_dense_bf16.err; echo "dense bf16 rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02y/bench.json").read().strip().splitlines()[-1])
print("value %.3f Gpx/s" % (d["value"] / 1e9), "ms %.4f" % d["ms_per_step"], "e2e %.3f" % (d["e2e"]["value"] / 1e9), d["roofline"]["frac"], d["kernels"]["group_A_loss_fwd_select_bwd"], d["workloads"][0]["ms_per_step"], d["cpu_baseline"]["value"])
d = json.loads(open("gpurun_out/r02y/bench_dense_bf16.json").read().strip().splitlines()[-1])
print("dense bf16: ms %.4f" % d["ms_per_step"], {k: v["ms_per_step"] for k, v in d["kernels"].items()})
PY
