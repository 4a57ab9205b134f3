"""Side benchmark: the DENSE bipartite projection of a cfg3 batch (GNN stage: 16 x 358 x 256 x 512 logits, 7 soft
graphs [C_ds, 358]) — tcgen05 kernel (mdseg_proj_fwd_tc) vs the shared-memory FFMA kernel (mdseg_proj_fwd) vs the
reference's per-dataset torch.einsum on the same GPU."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mdseg_b200 import ops, native as N
dev = "cuda:0"
peaks = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6650.0, "bf16_tflops": 1667.0}
n_cats = [19, 64, 37, 19, 26, 150, 133]; ids = [0, 0, 0, 1, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6]
B, Cu, h, w = 16, 358, 256, 512
g = torch.Generator(device=dev).manual_seed(1)
ids_t = torch.tensor(ids, dtype=torch.int32, device=dev)
flops = sum(2.0 * Cu * n_cats[d] * h * w for d in ids)
for dt in (torch.float32, torch.bfloat16):
    x = (torch.randn(B, Cu, h, w, generator=g, device=dev) * 3).to(dt)
    graphs = [torch.softmax(torch.randn(c, Cu, generator=g, device=dev) * 4, dim=0).requires_grad_(True) for c in n_cats]
    tab, keep = ops._default_graphs.table(graphs)
    cmax = max(n_cats)
    y = torch.empty(B, cmax, h, w, dtype=torch.float32, device=dev)
    ef = ops.err_flag(x.device)
    nbytes = N.lib.mdseg_proj_fwd_tc_workspace_bytes(C.byref(tab), ops._DT[dt])
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    def tc():
        N.call("mdseg_proj_fwd_tc", x.data_ptr(), ops._DT[dt], C.byref(tab), ids_t.data_ptr(), B, h, w, y.data_ptr(), cmax,
               None, ws.data_ptr(), nbytes, ef.data_ptr(), ops._stream())
    def ffma():
        N.call("mdseg_proj_fwd", x.data_ptr(), ops._DT[dt], C.byref(tab), ids_t.data_ptr(), B, h, w, y.data_ptr(), cmax,
               None, ef.data_ptr(), ops._stream())
    def eager():
        with torch.no_grad():
            return [torch.einsum("bchw,nc->bnhw", x[ids_t == d], graphs[d].to(dt)) for d in range(len(n_cats))]
    alg_bytes = x.numel() * x.element_size() + sum(n_cats[d] for d in ids) * h * w * 4
    def tc16():
        return ops.project(x, graphs, ids_t)
    for name, fn in (("tcgen05", tc), ("ffma", ffma), ("torch.einsum", eager)) + ((("tcgen05 TMA (ops.project)", tc16),) if dt != torch.float32 else ()):
        for _ in range(3): fn()
        ts = []
        for _ in range(7):
            torch.cuda._sleep(1_000_000)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        print(json.dumps({"kernel": name, "dtype": str(dt).split(".")[-1], "ms": round(ms, 4),
                          "useful_tflops": round(flops / ms / 1e9, 1),
                          "frac_of_bf16_peak": round(flops / ms / 1e9 / peaks["bf16_tflops"], 4),
                          "alg_gbs": round(alg_bytes / ms / 1e6, 1), "frac_of_hbm_peak": round(alg_bytes / ms / 1e6 / peaks["hbm_gbs"], 3)}))
