"""Side benchmark: the prototype head of a cfg3 batch — features [16, 512, 256, 512] x unify_prototype
[358, 512] -> unified logits [16, 358, 256, 512] (0.77 TFLOP) — forward, d features and d prototype on the tcgen05
kernels (ops.prototype_head) next to torch.einsum (cuBLAS, TF32 off for fp32) on the same GPU.  Tensor-pipe figure:
useful FLOP / time against MEASURED_PEAKS.json bf16_tflops_sustained (fp32 inputs issue six bf16 products per useful
one, so their ceiling is a sixth of it)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mdseg_b200 import ops
dev = "cuda:0"
torch.backends.cuda.matmul.allow_tf32 = False
peaks = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {}
tf_peak = peaks.get("bf16_tflops_sustained", 1406.7)
hbm = peaks.get("hbm_gbs", 6460.2)
B, K, Nn, h, w = 16, 512, 358, 256, 512
flop = 2.0 * B * K * Nn * h * w
g = torch.Generator(device=dev).manual_seed(1)


def timed(fn, n=5):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(n):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


for dt in (torch.float32, torch.bfloat16):
    feats = torch.randn(B, K, h, w, generator=g, device=dev).to(dt)
    proto = torch.randn(Nn, K, generator=g, device=dev) * 0.1
    dy = torch.randn(B, Nn, h, w, generator=g, device=dev)
    esz = feats.element_size()

    def ours_fwd():
        with torch.no_grad():
            return ops.prototype_head(feats, proto)

    def ours_fwd_bwd():
        f, p = feats.detach().requires_grad_(True), proto.detach().requires_grad_(True)
        ops.prototype_head(f, p).backward(dy)

    def torch_fwd():
        with torch.no_grad():
            return torch.einsum("bchw,nc->bnhw", feats, proto.to(dt))

    def torch_fwd_bwd():
        f, p = feats.detach().requires_grad_(True), proto.detach().requires_grad_(True)
        torch.einsum("bchw,nc->bnhw", f, p.to(dt)).backward(dy.to(dt))

    bytes_fwd = feats.numel() * esz + B * Nn * h * w * 4
    for name, fn, k in (("tcgen05 fwd", ours_fwd, 1), ("torch.einsum fwd", torch_fwd, 1),
                        ("tcgen05 fwd+bwd", ours_fwd_bwd, 3), ("torch.einsum fwd+bwd", torch_fwd_bwd, 3)):
        ms = timed(fn)
        tfl = k * flop / ms / 1e9
        print(json.dumps({"what": name, "dtype": str(dt).split(".")[-1], "ms": round(ms, 3), "useful_tflops": round(tfl, 1),
                          "frac_of_bf16_sustained": round(tfl / tf_peak, 4),
                          "fwd_alg_gbs": round(bytes_fwd / ms / 1e6, 1) if k == 1 else None,
                          "fwd_frac_of_hbm": round(bytes_fwd / ms / 1e6 / hbm, 3) if k == 1 else None}))
    del feats, dy
    torch.cuda.empty_cache()
