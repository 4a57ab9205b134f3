# usage: bash profiles/r02_run_scale_graph.sh N   (under gpurun --gpus N): strong scaling with and without the CUDA graph
N=$1
mkdir -p gpurun_out/r02s
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 50 --warmup 5 --no-cpu-baseline --no-kernel-times --no-aux-workload --scaling strong "$@"; }
run --shard contiguous > gpurun_out/r02s/strong_n$N.json 2> gpurun_out/r02s/strong_n$N.err; echo "strong rc=$?"
run --shard contiguous --cuda-graph > gpurun_out/r02s/strong_graph_n$N.json 2> gpurun_out/r02s/strong_graph_n$N.err; echo "strong graph rc=$?"
run --shard balanced > gpurun_out/r02s/strong_balanced_n$N.json 2> gpurun_out/r02s/strong_balanced_n$N.err; echo "strong balanced rc=$?"
run --shard balanced --cuda-graph > gpurun_out/r02s/strong_balanced_graph_n$N.json 2> gpurun_out/r02s/strong_balanced_graph_n$N.err; echo "strong balanced graph rc=$?"
python - <<PY
import json
for k in ("strong", "strong_graph", "strong_balanced", "strong_balanced_graph"):
    try:
        d = json.loads(open("gpurun_out/r02s/%s_n$N.json" % k).read().strip().splitlines()[-1])
        print(k, "n", d["n_gpus"], "value %.3f Gpx/s" % (d["value"] / 1e9), "ms %.3f" % d["ms_per_step"], d["hist_check"]["ok"], d["config"]["pixels_per_step_per_gpu"], d["loss"])
    except Exception as e:
        print(k, "failed", e)
PY
grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/r02s/strong_graph_n$N.err | tail -5
