"""Side benchmark (not the headline): evaluator accumulation a11 + a12 for one 1024x2048 image
(6 scales x flip, evaluate.py:136-181) per dataset class count."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mdseg_b200 import ops
dev = "cuda:0"
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
H, W = 1024, 2048
for C in (19, 64, 150):
    g = torch.Generator(device=dev).manual_seed(C)
    label = torch.randint(0, C, (H, W), generator=g, device=dev)
    passes = []
    for s in (0.5, 0.75, 1.0, 1.25, 1.5, 1.75):
        h, w = (int(s * H) + 31) // 32 * 32 // 4, (int(s * W) + 31) // 32 * 32 // 4
        for flip in (False, True):
            passes.append((torch.randn(C, h, w, generator=g, device=dev), flip))
    probs = torch.empty(C, H, W, device=dev)
    px = H * W
    hist = torch.zeros(C, C, dtype=torch.int64, device=dev)
    def run():
        first = True
        for lg, flip in passes:
            ops.eval_accum(lg, probs, flip=flip, first=first); first = False
        ops.argmax_hist(probs, label=label, hist=hist)
    for _ in range(2): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    byts = sum(lg.numel() * 4 for lg, _ in passes) + (2 * len(passes) - 1) * C * px * 4 + C * px * 4 + px * 16
    hist2 = torch.zeros(C, C, dtype=torch.int64, device=dev)
    for _ in range(2): ops.eval_fused(passes, (H, W), label=label, hist=hist2, want_pred=False)
    torch.cuda.synchronize()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(3): ops.eval_fused(passes, (H, W), label=label, hist=hist2, want_pred=False)
    f1.record(); torch.cuda.synchronize()
    fms = f0.elapsed_time(f1) / 3
    fbytes = sum(lg.numel() * 4 for lg, _ in passes) + px * 8 + 12 * px * 8 * (1 + (C + 15) // 16)
    print(json.dumps({"case": f"eval 12 passes C={C} 1024x2048, no probability tensor (mdseg_eval_fused)", "ms": round(fms, 3),
                      "Mpx_s": round(px / fms / 1e3, 1), "alg_GB": round(fbytes / 1e9, 3),
                      "exps_G": round(2 * 12 * C * px / 1e9, 2), "speedup_vs_pass_by_pass": round(ms / fms, 2)}))
    print(json.dumps({"case": f"eval 12 passes C={C} 1024x2048", "ms": round(ms, 3), "Mpx_s": round(px / ms / 1e3, 1),
                      "alg_GB": round(byts / 1e9, 2), "achieved_gbs": round(byts / ms / 1e6, 1),
                      "frac_of_measured_peak": round(byts / ms / 1e6 / peak, 3)}))
