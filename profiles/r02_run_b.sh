set -x
mkdir -p gpurun_out/r02b
( timeout 900 python -m pytest tests/test_gpu_dropin.py -x -q ) > gpurun_out/r02b/pytest_dropin.log 2>&1; tail -5 gpurun_out/r02b/pytest_dropin.log
timeout 900 python profiles/bench_gnn_stage_feats.py > gpurun_out/r02b/gnn_stage_feats.jsonl 2> gpurun_out/r02b/gnn_stage_feats.err; tail -3 gpurun_out/r02b/gnn_stage_feats.err; cat gpurun_out/r02b/gnn_stage_feats.jsonl
bash profiles/abso.sh base recur base recur 2>&1 | tee gpurun_out/r02b/ab_recur.txt
