# usage: bash profiles/r02_run_numa.sh N  (under gpurun --gpus N): e2e arm with and without NUMA binding
N=$1
mkdir -p gpurun_out/r02n
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --no-kernel-times --no-aux-workload "$@"; }
run > gpurun_out/r02n/bind_n$N.json 2> gpurun_out/r02n/bind_n$N.err; echo "bind rc=$?"
run --no-numa-bind > gpurun_out/r02n/nobind_n$N.json 2> gpurun_out/r02n/nobind_n$N.err; echo "nobind rc=$?"
python - <<PY
import json
for k in ("bind", "nobind"):
    try:
        d = json.loads(open("gpurun_out/r02n/%s_n$N.json" % k).read().strip().splitlines()[-1])
        print(k, "value %.2f" % (d["value"] / 1e9), "e2e %.3f Gpx/s" % (d["e2e"]["value"] / 1e9), d["e2e"]["h2d_gbs_per_rank"], d["e2e"]["numa"])
    except Exception as e:
        print(k, "failed", e)
PY
nvidia-smi topo -m 2>/dev/null | head -14
