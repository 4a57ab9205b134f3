"""Side benchmark (not the headline): OhemCELoss on full-resolution logits (BASELINE cfg1 / cfg4)."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mdseg_b200 import ops
dev = "cuda:0"
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
def run(name, N, C, H, W, dt, cl=False, lab_dt=torch.int64, graph=False):
    g = torch.Generator(device=dev).manual_seed(1)
    x = (torch.randn(N, C, H, W, generator=g, device=dev) * 3).to(dt)
    if cl: x = x.contiguous(memory_format=torch.channels_last)
    x.requires_grad_(True)
    lb = torch.randint(0, C, (N, H, W), generator=g, device=dev)
    lb[torch.rand(N, H, W, generator=g, device=dev) < 0.05] = 255
    lb = lb.to(lab_dt)
    th = ops.neg_log(0.7)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    def step():
        x.grad = None
        l = ops.ohem_ce(x, lb, th); l.backward(); return l
    for _ in range(3): step()
    torch.cuda.synchronize()
    run_step = step
    if graph:  # forward + selection + autograd backward as ONE CUDA graph: what is left when the launches cost nothing
        cg = torch.cuda.CUDAGraph()
        cap = torch.cuda.Stream()
        cap.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(cap):
            with torch.cuda.graph(cg, stream=cap, capture_error_mode="thread_local"):
                step()
        torch.cuda.current_stream().wait_stream(cap)
        run_step = cg.replay
        for _ in range(3): run_step()
        torch.cuda.synchronize()
    ts = []
    for _ in range(9 if graph else 5):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run_step(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    e = x.element_size(); L = lb.element_size()
    bpp = 3 * C * e + 2 * L + 20
    px = N * H * W
    gbs = bpp * px / (ms * 1e-3) / 1e9
    print(json.dumps({"case": name, "ms": round(ms, 4), "gpx_s": round(px / ms / 1e6, 3), "alg_B_per_px": bpp,
                      "achieved_gbs": round(gbs, 1), "frac_of_measured_peak": round(gbs / peak, 4)}))
run("cfg1 2x19x512x1024 f32", 2, 19, 512, 1024, torch.float32)
run("cfg1 2x19x512x1024 f32, step as one CUDA graph", 2, 19, 512, 1024, torch.float32, graph=True)
run("cfg1x8 16x19x512x1024 f32", 16, 19, 512, 1024, torch.float32)
run("cfg4 16x171x640x640 f32 NCHW", 16, 171, 640, 640, torch.float32)
run("cfg4 16x171x640x640 f32 NHWC", 16, 171, 640, 640, torch.float32, cl=True)
run("cfg4 16x171x640x640 bf16 NCHW", 16, 171, 640, 640, torch.bfloat16)
run("cfg4 16x171x640x640 bf16 NCHW u8 labels", 16, 171, 640, 640, torch.bfloat16, lab_dt=torch.uint8)
