#!/bin/bash
# A/B of compile-time switches on the GPU box: profiles/ab.sh <file.cu> "<flags A>" "<flags B>" ...
# rebuilds csrc/<file.cu> with each flag set (MDSEG_CFLAGS) and prints the per-call kernel times of bench.py.
f=$1; shift
for flags in "$@"; do
  touch mul-datasets-semantic-segmentation_b200/csrc/$f
  MDSEG_CFLAGS="$flags" python mul-datasets-semantic-segmentation_b200/build.py > /dev/null || { echo "build failed: $flags"; continue; }
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/ab.json 2> gpurun_out/ab.err || { echo "bench failed: $flags"; tail -3 gpurun_out/ab.err; continue; }
  python - "$flags" <<'PY'
import json, sys
d = json.load(open("gpurun_out/ab.json"))
k = d["kernels"]
print("%-60s step %.3f  fwd %.4f  bwd %.4f  proj %.4f  A %.4f" % (sys.argv[1] or "(default)", d["ms_per_step"], k["mdseg_up_ce_fwd"]["ms_per_step"], k["mdseg_mds_bwd"]["ms_per_step"], k["mdseg_proj_fwd"]["ms_per_step"], k["group_A_loss_fwd_select_bwd"]["ms_per_step"]))
PY
done
touch mul-datasets-semantic-segmentation_b200/csrc/$f
python mul-datasets-semantic-segmentation_b200/build.py > /dev/null
