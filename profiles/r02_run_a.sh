set -x
mkdir -p gpurun_out/r02a
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02a/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a/pytest.log
timeout 600 python bench.py > gpurun_out/r02a/bench.json 2> gpurun_out/r02a/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02a/bench_ref.json 2> gpurun_out/r02a/bench_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02a/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-times --no-aux-workload > gpurun_out/r02a/ncu_launch.log 2>&1; echo "launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'mds_bwd_kernel|up_ce_fwd_warp_kernel|proj_fwd_sparse' -c 3 -o gpurun_out/r02a/full_top3 -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-kernel-times --no-aux-workload > gpurun_out/r02a/ncu_full.log 2>&1; echo "full rc=$?"
tail -3 gpurun_out/r02a/pytest.log
cat gpurun_out/r02a/bench.json | cut -c1-1500
