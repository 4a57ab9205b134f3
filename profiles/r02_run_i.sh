mkdir -p gpurun_out/r02i
for a in "--logits-dtype bf16" "--with-aux" "--workload cfg3_native" "--logits mixed" "--bi-graphs dense" "--eager-gpu"; do
  n=$(echo "$a" | tr -d ' -')
  timeout 600 python bench.py --no-cpu-baseline --no-aux-workload $a > gpurun_out/r02i/bench_$n.json 2> gpurun_out/r02i/bench_$n.err; echo "$n rc=$?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02i/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        k = d["kernels"]
        print(f.split("/")[-1], "ms %.4f" % d["ms_per_step"], "Gpx/s %.2f"%(d["value"]/1e9), {x: k[x]["ms_per_step"] for x in k}, k["group_A_loss_fwd_select_bwd"], d.get("reference_eager_gpu"))
    except Exception as e:
        print(f, "failed", e)
PY
# (a compute-sanitizer memcheck pass was attempted here: the tool is closed on this GPU pool)
