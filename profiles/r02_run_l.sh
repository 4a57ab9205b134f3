mkdir -p gpurun_out/r02l
( timeout 900 python -m pytest tests/test_gpu_proto_head.py tests/test_gpu_proj_tc.py tests/test_gpu_dropin.py tests/test_gpu_eval.py -q ) > gpurun_out/r02l/pytest.log 2>&1; tail -3 gpurun_out/r02l/pytest.log
for i in 1 2; do
timeout 600 python bench.py --no-cpu-baseline --no-kernel-times > gpurun_out/r02l/bench_$i.json 2> gpurun_out/r02l/bench_$i.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r02l/bench_$i.json").read().strip().splitlines()[-1])
print("run $i ms %.4f" % d["ms_per_step"], d["workloads"])
PY
done
timeout 300 python profiles/bench_gnn_stage_feats.py > gpurun_out/r02l/gnn_stage_feats.jsonl 2> gpurun_out/r02l/gnn.err; cut -c1-140 gpurun_out/r02l/gnn_stage_feats.jsonl
timeout 300 python profiles/bench_proto_head.py > gpurun_out/r02l/proto_head.jsonl 2> gpurun_out/r02l/ph.err; cut -c1-200 gpurun_out/r02l/proto_head.jsonl
timeout 300 python profiles/bench_proj_dense.py > gpurun_out/r02l/proj_dense.jsonl 2> gpurun_out/r02l/pd.err; tail -1 gpurun_out/r02l/proj_dense.jsonl | cut -c1-200
