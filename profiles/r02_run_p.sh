mkdir -p gpurun_out/r02p
( timeout 900 python -m pytest tests/test_gpu_proto_head.py tests/test_gpu_dropin.py -q ) > gpurun_out/r02p/pytest.log 2>&1; tail -3 gpurun_out/r02p/pytest.log
GNN_ROUTES="pair,folded" timeout 300 python profiles/bench_gnn_stage_feats.py > gpurun_out/r02p/gnn.jsonl 2> gpurun_out/r02p/gnn.err; cut -c1-110 gpurun_out/r02p/gnn.jsonl
