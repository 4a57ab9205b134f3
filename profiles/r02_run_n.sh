mkdir -p gpurun_out/r02n2
GNN_ROUTES=folded timeout 600 ncu --set full --clock-control none --import-source on -k regex:'head_tc16' -c 4 -o gpurun_out/r02n2/full_folded_tc16 -f python profiles/bench_gnn_stage_feats.py > gpurun_out/r02n2/ncu1.log 2>&1; echo "ncu1 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'head_tc16_kernel' -c 1 -o gpurun_out/r02n2/full_head_fwd -f python profiles/bench_proto_head.py > gpurun_out/r02n2/ncu2.log 2>&1; echo "ncu2 rc=$?"
