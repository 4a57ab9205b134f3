set -x
mkdir -p gpurun_out/r02f
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/r02f/pytest.log 2>&1; tail -6 gpurun_out/r02f/pytest.log
for a in "" "--logits confident" "--workload cfg1" "--workload cfg2"; do
  n=$(echo "$a" | tr -d ' -'); n=${n:-default}
  timeout 600 python bench.py --no-cpu-baseline --no-aux-workload $a > gpurun_out/r02f/bench_$n.json 2> gpurun_out/r02f/bench_$n.err; echo "$n rc=$?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02f/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        k = d["kernels"]
        print(f.split("/")[-1], "ms %.4f" % d["ms_per_step"], d["ohem"]["branch"], {x: k[x]["ms_per_step"] for x in k}, k.get("mdseg_ohem_select"))
    except Exception as e:
        print(f, "failed", e)
PY
