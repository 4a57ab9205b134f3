#!/bin/bash
# A/B of prebuilt library variants on the GPU box (no nvcc time there): profiles/abso.sh <name> [<name> ...]
# Each profiles/_variants/<name>.so (built in the dev container with MDSEG_CFLAGS=..., see profiles/ab.sh for the
# in-place variant) is copied over libmdseg_b200.so and bench.py's per-call kernel times are printed; "base" is the
# default build and is restored at the end.
LIB=mul-datasets-semantic-segmentation_b200/libmdseg_b200.so
for v in "$@"; do
  cp profiles/_variants/$v.so $LIB || continue
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-aux-workload ${BENCH_ARGS} > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err || { echo "bench failed: $v"; tail -3 gpurun_out/ab_$v.err; continue; }
  python - "$v" <<'PY'
import json, sys
d = json.load(open("gpurun_out/ab_%s.json" % sys.argv[1]))
k = d["kernels"]
g = lambda n: k.get(n, {}).get("ms_per_step", float("nan"))
print("%-28s step %.3f  fwd %.4f  bwd %.4f  proj %.4f  A %.4f  loss %.7f" % (sys.argv[1], d["ms_per_step"], g("mdseg_up_ce_fwd"), g("mdseg_mds_bwd"), g("mdseg_proj_fwd"), g("group_A_loss_fwd_select_bwd"), d["loss"]))
PY
done
cp profiles/_variants/base.so $LIB
