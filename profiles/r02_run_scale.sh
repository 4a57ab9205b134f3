# usage: bash profiles/r02_run_scale.sh N   (under gpurun --gpus N)
N=$1
mkdir -p gpurun_out/r02s
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-times "$@"; }
run > gpurun_out/r02s/weak_n$N.json 2> gpurun_out/r02s/weak_n$N.err; echo "weak rc=$?"
run --scaling strong --no-aux-workload > gpurun_out/r02s/strong_n$N.json 2> gpurun_out/r02s/strong_n$N.err; echo "strong rc=$?"
python - <<PY
import json
for k in ("weak", "strong"):
    try:
        d = json.load(open("gpurun_out/r02s/%s_n$N.json" % k))
        print(k, "n", d["n_gpus"], "value %.3f Gpx/s" % (d["value"] / 1e9), "ms %.3f" % d["ms_per_step"], "e2e %.3f" % (d["e2e"]["value"] / 1e9), d["hist_check"]["ok"], d["config"]["pixels_per_step_per_gpu"])
    except Exception as e:
        print(k, "failed", e)
PY
tail -3 gpurun_out/r02s/*_n$N.err
