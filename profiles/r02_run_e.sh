set -x
mkdir -p gpurun_out/r02e
python profiles/_dbg_fp16.py 2>&1 | tail -8
( timeout 900 python -m pytest tests/test_gpu_proto_head.py tests/test_gpu_dropin.py -q ) > gpurun_out/r02e/pytest.log 2>&1; tail -8 gpurun_out/r02e/pytest.log
GNN_ROUTES=pair,folded timeout 900 python profiles/bench_gnn_stage_feats.py > gpurun_out/r02e/gnn_stage_feats.jsonl 2> gpurun_out/r02e/gnn_stage_feats.err; tail -3 gpurun_out/r02e/gnn_stage_feats.err; cut -c1-700 gpurun_out/r02e/gnn_stage_feats.jsonl
