set -x
mkdir -p gpurun_out/r02c
( timeout 900 python -m pytest tests/test_gpu_proto_head.py tests/test_gpu_proj_tc.py tests/test_gpu_dropin.py -x -q ) > gpurun_out/r02c/pytest.log 2>&1; tail -15 gpurun_out/r02c/pytest.log
timeout 900 python profiles/bench_gnn_stage_feats.py > gpurun_out/r02c/gnn_stage_feats.jsonl 2> gpurun_out/r02c/gnn_stage_feats.err; tail -3 gpurun_out/r02c/gnn_stage_feats.err; cat gpurun_out/r02c/gnn_stage_feats.jsonl
for b in fullres confusion eval label_pipeline proj_dense proto_head gnn_stage; do
  timeout 600 python profiles/bench_$b.py > gpurun_out/r02c/$b.jsonl 2> gpurun_out/r02c/$b.err; echo "$b rc=$?"; tail -2 gpurun_out/r02c/$b.err
done
