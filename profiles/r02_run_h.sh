mkdir -p gpurun_out/r02h
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ohem_select' -s 3 -c 1 -o gpurun_out/r02h/full_select -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-kernel-times --no-aux-workload --logits confident > gpurun_out/r02h/ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r02h/ncu.log | cut -c1-300
