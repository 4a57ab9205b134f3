"""Side benchmark: per-dataset confusion matrices of a cfg3 batch (16 x 1024 x 2048) for uniformly random
(label, pred) pairs — every pixel is its own shared-memory atomic — and for piecewise-constant maps
(32 x 32 blocks, what segmentation maps look like: the per-thread run-length aggregation removes most atomics)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mdseg_b200 import ops
dev = "cuda:0"
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
n_cats = [19, 64, 37, 19, 26, 150, 133]; ids = [0, 0, 0, 1, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6]
B, H, W = 16, 1024, 2048
g = torch.Generator(device=dev).manual_seed(1)
ids_t = torch.tensor(ids, dtype=torch.int32, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def blocky(d, blk):
    small = torch.randint(0, n_cats[d], (H // blk, W // blk), generator=g, device=dev)
    return small.repeat_interleave(blk, 0).repeat_interleave(blk, 1)
for name, blk in (("uniform random pixels", 1), ("piecewise constant 32x32 blocks", 32)):
    for dt, L in ((torch.int64, 8), (torch.uint8, 1)):
        lab = torch.stack([blocky(d, blk) for d in ids]).to(dt)
        pred = torch.stack([blocky(d, blk) for d in ids])
        hist, _ = ops.confusion_images(lab, pred, ids, n_cats)
        ts = []
        for _ in range(7):
            flush.fill_(1); hist.zero_()
            torch.cuda._sleep(2_000_000)  # ~1 ms of GPU idle so that the host-side call overhead is off the clock
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ops.confusion_images(lab, pred, ids_t, n_cats, hist=hist); e1.record()
            torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        px = B * H * W
        gbs = (L + 8) * px / ms / 1e6
        print(json.dumps({"case": name, "labels": str(dt).split(".")[-1], "ms": round(ms, 4), "gpx_s": round(px / ms / 1e6, 1),
                          "alg_B_per_px": L + 8, "achieved_gbs": round(gbs, 1), "frac_of_measured_peak": round(gbs / peak, 3)}))
