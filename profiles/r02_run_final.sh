# the evidence of the final build of the round: GPU tests, smoke, headline bench + reference arm, launch list, full-set capture
mkdir -p gpurun_out/r02z
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/r02z/pytest.log 2>&1; tail -4 gpurun_out/r02z/pytest.log
python __graft_entry__.py smoke > gpurun_out/r02z/smoke.log 2>&1; tail -1 gpurun_out/r02z/smoke.log
timeout 600 python bench.py > gpurun_out/r02z/bench.json 2> gpurun_out/r02z/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02z/bench_ref.json 2> gpurun_out/r02z/bench_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02z/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-times --no-aux-workload > gpurun_out/r02z/ncu_launch.log 2>&1; echo "launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'mds_bwd_kernel|up_ce_fwd_warp_kernel|proj_fwd_sparse' -c 3 -o gpurun_out/r02z/full_top3 -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-kernel-times --no-aux-workload > gpurun_out/r02z/ncu_full.log 2>&1; echo "full rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02z/bench.json").read().strip().splitlines()[-1])
print("value %.3f Gpx/s" % (d["value"] / 1e9), "ms %.4f" % d["ms_per_step"], "e2e %.3f" % (d["e2e"]["value"] / 1e9), d["roofline"], d["kernels"]["group_A_loss_fwd_select_bwd"], d["cpu_baseline"]["value"], d["gpu_launches"])
PY
