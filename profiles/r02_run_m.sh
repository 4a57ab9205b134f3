mkdir -p gpurun_out
BENCH_ARGS="--no-kernel-times-skip" true
bash profiles/abso.sh base fwd_occ16_st2 fwd_occ12 fwd_seg4 fwd_seg8 bwd_seg24 bwd_seg8 base 2>&1 | tee gpurun_out/ab_tuning.txt
