set -x
mkdir -p gpurun_out/r02d
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02d/pytest.log 2>&1; tail -6 gpurun_out/r02d/pytest.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02d/bench.json 2> gpurun_out/r02d/bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r02d/bench.err
timeout 600 python bench.py --no-cpu-baseline --logits confident --no-aux-workload > gpurun_out/r02d/bench_confident.json 2> gpurun_out/r02d/bench_confident.err; echo "confident rc=$?"
timeout 600 python bench.py --no-cpu-baseline --workload cfg1 --no-aux-workload > gpurun_out/r02d/bench_cfg1.json 2> gpurun_out/r02d/bench_cfg1.err; echo "cfg1 rc=$?"
timeout 600 python profiles/bench_fullres.py > gpurun_out/r02d/fullres.jsonl 2> gpurun_out/r02d/fullres.err; head -2 gpurun_out/r02d/fullres.jsonl
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'head_tc16' -c 6 -o gpurun_out/r02d/full_head_tc16 -f python profiles/bench_gnn_stage_feats.py > gpurun_out/r02d/ncu_head.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import json
for n in ("bench", "bench_confident", "bench_cfg1"):
    try:
        d = json.loads(open("gpurun_out/r02d/%s.json" % n).read().strip().splitlines()[-1])
        k = d["kernels"]
        print(n, "ms %.4f" % d["ms_per_step"], d["ohem"]["branch"], {x: k[x]["ms_per_step"] for x in k})
    except Exception as e:
        print(n, "failed", e)
PY
