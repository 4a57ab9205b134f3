"""Side benchmark: the GNN-stage segmentation loss of a cfg3 batch FROM THE FEATURES — what
CrossDatasetsCELoss_AdvGNN / _GNN compute from preds['seg'] [16, 512, 256, 512], unify_prototype [358, 512] and trainable
dense bi_graphs (lib/loss/loss_cross_datasets.py:971, :996-1007, :1074), forward + backward to the features, the
prototypes and every graph:

  unfolded  ops.prototype_head (tcgen05 GEMM, N = C_uni) -> ops.mds_proj_ohem_ce with dense graphs (the reference's
            operation order: the [16, 358, 256, 512] unified logits and their gradient round-trip HBM)
  folded    ops.mds_head_proj_ohem_ce: one contraction with bi_graph @ unify_prototype (N = C_ds), no unified logits
  eager     the reference's torch ops on the same GPU (einsum, einsum, F.interpolate, CrossEntropyLoss(none), OHEM)

Prints one JSON line per (route, dtype) with the step time and the per-C-ABI-call times."""
import collections, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from mdseg_b200 import ops, native as N
dev = "cuda:0"
torch.backends.cuda.matmul.allow_tf32 = False
n_cats = [19, 64, 37, 19, 26, 150, 133]; ids = [0, 0, 0, 1, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6]
B, K, Cu, h, w, H, W = 16, 512, 358, 256, 512, 1024, 2048
g = torch.Generator(device=dev).manual_seed(1)
ids_t = torch.tensor(ids, dtype=torch.int32, device=dev)
labels = torch.stack([torch.randint(0, n_cats[d], (H, W), generator=g, device=dev) for d in ids])
labels[torch.rand(B, H, W, generator=g, device=dev) < 0.05] = 255
thresh = ops.neg_log(0.4)
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


def eager(feats, proto, graphs):
    logits = torch.einsum("bchw,nc->bnhw", feats, proto.to(feats.dtype))
    losses = []
    for d in range(len(n_cats)):
        rows = [i for i, v in enumerate(ids) if v == d]
        y = torch.einsum("bchw,nc->bnhw", logits[rows], graphs[d].to(feats.dtype))
        y = F.interpolate(y, size=(H, W), mode="bilinear", align_corners=True)
        losses.append(F.cross_entropy(y.float(), labels[rows], ignore_index=255, reduction="none").view(-1))
    loss = torch.cat(losses)
    n_min = int((labels != 255).sum()) // 16
    hard = loss[loss > thresh]
    if hard.numel() < n_min:
        hard, _ = loss.topk(n_min)
    return hard.mean()


routes = {
    "unfolded": lambda f, p, gs: ops.mds_proj_ohem_ce(ops.prototype_head(f, p), labels, ids_t, gs, thresh),
    "folded": lambda f, p, gs: ops.mds_head_proj_ohem_ce(f, p, labels, ids_t, gs, thresh),
    "eager": eager,
    # the (hard, soft) graph pair of the GNN stage, blended 0.5 / 0.5 (:1063-1071): two folded losses, or one stacked
    # projection per direction (ops.mds_head_proj_ohem_ce_heads)
    "pair, two folded losses": lambda f, p, gs: 0.5 * ops.mds_head_proj_ohem_ce(f, p, labels, ids_t, gs, thresh)
    + 0.5 * ops.mds_head_proj_ohem_ce(f, p, labels, ids_t, gs2, thresh),
    "pair, stacked heads": lambda f, p, gs: ops.mds_head_proj_ohem_ce_heads(f, p, labels, ids_t, [gs, gs2], thresh).mean(),
}
if os.environ.get("GNN_ROUTES"):
    routes = {k: v for k, v in routes.items() if any(t in k for t in os.environ["GNN_ROUTES"].split(","))}
for dt in (torch.float32, torch.bfloat16):
    feats = torch.randn(B, K, h, w, generator=g, device=dev).to(dt).requires_grad_(True)
    proto = (torch.randn(Cu, K, generator=g, device=dev) * 0.1).requires_grad_(True)
    graphs = [torch.softmax(torch.randn(c, Cu, generator=g, device=dev) * 4, dim=0).requires_grad_(True) for c in n_cats]
    gs2 = [torch.softmax(torch.randn(c, Cu, generator=g, device=dev), dim=0).requires_grad_(True) for c in n_cats]
    for name, fn in routes.items():
        def step():
            feats.grad = None; proto.grad = None
            for m in graphs + gs2: m.grad = None
            loss = fn(feats, proto, graphs)
            loss.backward()
            return loss
        try:
            for _ in range(2): step()
            torch.cuda.synchronize()
            times.clear(); on["v"] = name != "eager"
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 3
            t0.record()
            for _ in range(n): loss = step()
            t1.record(); torch.cuda.synchronize()
            on["v"] = False
            out = {"route": name, "dtype": str(dt).split(".")[-1], "step_ms": round(t0.elapsed_time(t1) / n, 3),
                   "loss": float(loss.detach()), "dfeats_absmax": float(feats.grad.abs().max()),
                   "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 1e9, 2), "calls_ms": {}}
            for k, v in times.items():
                out["calls_ms"][k] = round(sum(a.elapsed_time(b) for a, b in v) / n, 3)
        except torch.OutOfMemoryError as e:
            out = {"route": name, "dtype": str(dt).split(".")[-1], "error": "out of memory"}
        print(json.dumps(out), flush=True)
        torch.cuda.empty_cache(); torch.cuda.reset_peak_memory_stats()
    del feats, proto, graphs
    torch.cuda.empty_cache()
