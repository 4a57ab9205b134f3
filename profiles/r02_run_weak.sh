# usage: bash profiles/r02_run_weak.sh N  (under gpurun --gpus N): the default bench line at N GPUs
N=$1
mkdir -p gpurun_out/r02s
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02s/weak_final_n$N.json 2> gpurun_out/r02s/weak_final_n$N.err; echo "rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/r02s/weak_final_n$N.json").read().strip().splitlines()[-1])
print("n", d["n_gpus"], "value %.3f Gpx/s" % (d["value"] / 1e9), "ms %.4f" % d["ms_per_step"], "e2e %.3f" % (d["e2e"]["value"] / 1e9), d["e2e"]["h2d_gbs_per_rank"], d["hist_check"]["ok"], d["workloads"][0]["ms_per_step"])
PY
