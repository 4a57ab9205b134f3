"""Side benchmark: the multi-dataset loss with DENSE, trainable bi_graphs (GNN stage, loss_cross_datasets.py:997-1007)
on a cfg3 batch — per C-ABI call, so that the next kernel to move to the tensor cores is visible."""
import collections, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mdseg_b200 import ops, native as N
dev = "cuda:0"
n_cats = [19, 64, 37, 19, 26, 150, 133]; ids = [0, 0, 0, 1, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6]
B, Cu, h, w, H, W = 16, 358, 256, 512, 1024, 2048
g = torch.Generator(device=dev).manual_seed(1)
ids_t = torch.tensor(ids, dtype=torch.int32, device=dev)
x = (torch.randn(B, Cu, h, w, generator=g, device=dev)).requires_grad_(True)
graphs = [torch.softmax(torch.randn(c, Cu, generator=g, device=dev) * 4, dim=0).requires_grad_(True) for c in n_cats]
labels = torch.stack([torch.randint(0, n_cats[d], (H, W), generator=g, device=dev) for d in ids])
labels[torch.rand(B, H, W, generator=g, device=dev) < 0.05] = 255
times = collections.defaultdict(list)
orig = N.call
on = {"v": False}
def timed(name, *a):
    if not on["v"]:
        return orig(name, *a)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); orig(name, *a); e1.record()
    times[name].append((e0, e1))
N.call = timed; ops.N.call = timed
def step():
    x.grad = None
    for m in graphs: m.grad = None
    loss = ops.mds_proj_ohem_ce(x, labels, ids_t, graphs, ops.neg_log(0.4))
    loss.backward()
    return loss
for _ in range(2): step()
torch.cuda.synchronize()
on["v"] = True
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 3
t0.record()
for _ in range(n): loss = step()
t1.record(); torch.cuda.synchronize()
out = {"step_ms": round(t0.elapsed_time(t1) / n, 3), "loss": float(loss.detach()), "calls_ms": {}}
for k, v in times.items():
    out["calls_ms"][k] = round(sum(a.elapsed_time(b) for a, b in v) / n, 3)
print(json.dumps(out))
