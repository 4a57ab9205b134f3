mkdir -p gpurun_out/r02j
for a in "--workload cfg1" "--workload cfg1 --cuda-graph" "--cuda-graph" ; do
  n=$(echo "$a" | tr -d ' -')
  timeout 600 python bench.py --no-cpu-baseline --no-aux-workload --no-kernel-times $a > gpurun_out/r02j/bench_$n.json 2> gpurun_out/r02j/bench_$n.err; echo "$n rc=$?"; tail -3 gpurun_out/r02j/bench_$n.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02j/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "ms %.4f" % d["ms_per_step"], "Gpx/s %.2f"%(d["value"]/1e9), d["gpu_launches"], d["loss"], d["hist_check"]["ok"])
    except Exception as e:
        print(f, "failed", e)
PY
