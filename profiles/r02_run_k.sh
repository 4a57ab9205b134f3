mkdir -p gpurun_out/r02k
for i in 1 2; do
timeout 600 python bench.py --no-cpu-baseline --no-kernel-times > gpurun_out/r02k/bench_$i.json 2> gpurun_out/r02k/bench_$i.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r02k/bench_$i.json").read().strip().splitlines()[-1])
print("run $i ms %.4f" % d["ms_per_step"], [(w["name"], round(w["ms_per_step"], 3)) for w in d["workloads"]], d["e2e"]["ms_per_step"])
PY
done
timeout 600 python bench.py --no-cpu-baseline --no-kernel-times --with-aux --no-aux-workload > gpurun_out/r02k/bench_aux.json 2> gpurun_out/r02k/bench_aux.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r02k/bench_aux.json").read().strip().splitlines()[-1])
print("with-aux ms %.4f" % d["ms_per_step"])
PY
timeout 300 python profiles/bench_gnn_stage_feats.py 2>/dev/null | grep -v eager | cut -c1-420
