"""GPU parity of the tcgen05 dense bipartite projection (mdseg_proj_fwd_tc) against a float64 einsum
(loss_cross_datasets.py:1006 with the dense, trainable bi_graphs of the GNN stage)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ops():
    from mdseg_b200 import ops
    return ops


def rel_err(a, b):
    return np.abs(np.asarray(a, dtype=np.float64) - b).max() / max(np.abs(b).max(), 1e-30)


def dense_graph(g, c_ds, c_uni):
    return torch.softmax(torch.randn(c_ds, c_uni, generator=g) * 4, dim=0)


def onehot_graph(g, c_ds, c_uni):
    idx = torch.randint(0, c_ds, (c_uni,), generator=g)
    m = torch.zeros(c_ds, c_uni)
    m[idx, torch.arange(c_uni)] = 1
    return m


CASES = [
    # n_cats, kinds (d = dense, s = one-hot sparse), C_uni, ids, h, w
    ([19, 150, 133], "ddd", 358, [0, 1, 2, 1], 16, 32),      # cfg3-like widths; hw multiple of 128
    ([19, 12, 36], "ddd", 67, [2, 0, 1, 2, 0], 9, 13),       # cfg2 widths; ragged hw (117), K tail (67 = 2*32 + 3)
    ([64, 37, 26], "dsd", 124, [0, 1, 2, 2, 1, 0], 12, 20),  # sparse and dense datasets in one call, shuffled ids
    ([256, 8, 7], "ddd", 96, [0, 1, 2], 8, 16),              # N limit, smallest N, and C_ds 7 < 8 -> FFMA kernel
    ([150], "d", 32, [0, 0], 16, 16),                        # a single K chunk
]


@pytest.mark.parametrize("case", range(len(CASES)))
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16, torch.float16])
def test_proj_tc_matches_float64_einsum(ops, case, dt):
    n_cats, kinds, c_uni, ids, h, w = CASES[case]
    g = torch.Generator().manual_seed(100 + case)
    x = (torch.randn(len(ids), c_uni, h, w, generator=g) * 3).to(dt)
    graphs = [dense_graph(g, c, c_uni) if k == "d" else onehot_graph(g, c, c_uni) for c, k in zip(n_cats, kinds)]
    dev_graphs = [m.to(DEV).requires_grad_(k == "d") for m, k in zip(graphs, kinds)]  # requires_grad -> dense route
    y = ops.project(x.to(DEV), dev_graphs, torch.tensor(ids, device=DEV))
    ops.check_errors(DEV)
    # fp32: the 3 x 3-term bf16 split keeps ~2^-21; 16-bit inputs: G is rounded to the input type (autocast)
    tol = 1e-5 if dt == torch.float32 else 2e-2
    for b, d in enumerate(ids):
        want = torch.einsum("chw,nc->nhw", x[b].double(), graphs[d].double()).numpy()
        got = y[b, :n_cats[d]].cpu().numpy()
        assert rel_err(got, want) <= tol, (b, d, rel_err(got, want))


def test_proj_tc_fp32_is_tighter_than_one_bf16_product(ops):
    """The split really carries fp32: the error is orders of magnitude below a single bf16 product's (~4e-3)."""
    g = torch.Generator().manual_seed(7)
    x = torch.randn(2, 358, 16, 64, generator=g) * 5
    G = dense_graph(g, 150, 358)
    y = ops.project(x.to(DEV), [G.to(DEV).requires_grad_(True)], torch.zeros(2, dtype=torch.int32, device=DEV))
    want = torch.einsum("bchw,nc->bnhw", x.double(), G.double()).numpy()
    assert rel_err(y.cpu().numpy(), want) <= 5e-6


def test_proj_tc_loss_and_gradients_match_reference_ops(ops):
    """The whole fused loss with dense graphs (tensor-core forward, FFMA adjoints) against the reference op
    sequence on the CPU in float64."""
    from oracle import f64
    n_cats, c_uni, ids = [19, 40], 64, [0, 1, 1]
    h, w, H, W = 8, 16, 32, 64
    g = torch.Generator().manual_seed(3)
    x = torch.randn(len(ids), c_uni, h, w, generator=g) * 2
    graphs = [dense_graph(g, c, c_uni) for c in n_cats]
    labels = torch.stack([torch.randint(0, n_cats[d], (H, W), generator=g) for d in ids])
    labels[torch.rand(len(ids), H, W, generator=g) < 0.05] = 255
    xd = x.to(DEV).requires_grad_(True)
    gd = [m.to(DEV).requires_grad_(True) for m in graphs]
    loss = ops.mds_proj_ohem_ce(xd, labels.to(DEV), torch.tensor(ids, device=DEV), gd, ops.neg_log(0.4))
    loss.backward()
    ops.check_errors(DEV)
    want = f64.multi_dataset(x.numpy(), labels.numpy(), np.array(ids), [m.numpy() for m in graphs], ops.neg_log(0.4),
                             want_graph_grads=True)
    assert abs(float(loss) - want["loss"]) <= 1e-5 * abs(want["loss"])
    assert rel_err(xd.grad.cpu().numpy(), want["dlogits_uni"]) <= 1e-5
    for i in range(len(n_cats)):
        assert rel_err(gd[i].grad.cpu().numpy(), want["dgraphs"][i]) <= 2e-5


@pytest.mark.parametrize("case", range(len(CASES)))
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("two_planes", [False, True])
def test_proj_bwd_tc_matches_float64_einsum(ops, case, dt, two_planes):
    """mdseg_proj_bwd_tc: dx[b, u] = sum_n G_d[n, u] (dyA + dyB)[b, n] — dense graphs on the tensor cores, sparse
    ones and an image without a dataset (zero gradient) through the proj.cu kernels of the same call."""
    import ctypes as C
    from mdseg_b200 import native as N
    n_cats, kinds, c_uni, ids, h, w = CASES[case]
    ids = list(ids) + [-1]  # last image: not part of the loss
    B, cmax = len(ids), max(n_cats)
    g = torch.Generator().manual_seed(300 + case)
    graphs = [dense_graph(g, c, c_uni) if k == "d" else onehot_graph(g, c, c_uni) for c, k in zip(n_cats, kinds)]
    dev_graphs = [m.to(DEV).requires_grad_(k == "d") for m, k in zip(graphs, kinds)]
    dyA = torch.randn(B, cmax, h, w, generator=g)
    dyB = torch.randn(B, cmax, h, w, generator=g) if two_planes else None
    tab, keep = ops._default_graphs.table(dev_graphs)
    dx = torch.full((B, c_uni, h, w), 7.0, dtype=dt, device=DEV)
    ids_t = torch.tensor(ids, dtype=torch.int32, device=DEV)
    a, b = dyA.to(DEV), (dyB.to(DEV) if two_planes else None)
    nb = N.lib.mdseg_proj_bwd_tc_workspace_bytes(C.byref(tab), ops._DT[dt])
    ws = torch.empty(nb, dtype=torch.uint8, device=DEV)
    N.call("mdseg_proj_bwd_tc", a.data_ptr(), b.data_ptr() if two_planes else None, cmax, C.byref(tab),
           ids_t.data_ptr(), B, h, w, dx.data_ptr(), ops._DT[dt], ws.data_ptr(), nb, ops._stream())
    torch.cuda.synchronize()
    tol = 1e-5 if dt == torch.float32 else 2e-2
    for i, d in enumerate(ids):
        got = dx[i].float().cpu().numpy()
        if d < 0:
            assert not got.any()
            continue
        dy = dyA[i, :n_cats[d]].double() + (dyB[i, :n_cats[d]].double() if two_planes else 0)
        want = torch.einsum("nhw,nc->chw", dy, graphs[d].double()).numpy()
        assert rel_err(got, want) <= tol, (i, d, rel_err(got, want))


@pytest.mark.parametrize("case", range(len(CASES)))
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_proj_bwd_graph_tc_matches_float64_einsum(ops, case, dt):
    """mdseg_proj_bwd_graph_tc: dG_d[n, c] = sum over the images of d and the pixels of (dyA + dyB)[n] * x[c] — split-K
    UMMA with a fixed-order reduction; datasets outside the envelope (sparse, C_ds < 8) take the FFMA kernel."""
    import ctypes as C
    from mdseg_b200 import native as N
    n_cats, kinds, c_uni, ids, h, w = CASES[case]
    ids = list(ids) + [-1]
    B, cmax = len(ids), max(n_cats)
    g = torch.Generator().manual_seed(500 + case)
    graphs = [dense_graph(g, c, c_uni) if k == "d" else onehot_graph(g, c, c_uni) for c, k in zip(n_cats, kinds)]
    dev_graphs = [m.to(DEV).requires_grad_(k == "d") for m, k in zip(graphs, kinds)]
    x = (torch.randn(B, c_uni, h, w, generator=g) * 2).to(dt)
    dyA = torch.randn(B, cmax, h, w, generator=g)
    dyB = torch.randn(B, cmax, h, w, generator=g)
    tab, keep = ops._default_graphs.table(dev_graphs)
    stride = cmax * c_uni
    ids_t = torch.tensor(ids, dtype=torch.int32, device=DEV)
    xd, a, b = x.to(DEV), dyA.to(DEV), dyB.to(DEV)
    nb = N.lib.mdseg_proj_bwd_graph_tc_workspace_bytes(C.byref(tab), B, h, w)
    ws = torch.empty(nb, dtype=torch.uint8, device=DEV)
    outs = []
    for _ in range(2):
        dG = torch.zeros(len(n_cats), stride, dtype=torch.float32, device=DEV)
        N.call("mdseg_proj_bwd_graph_tc", xd.data_ptr(), ops._DT[dt], a.data_ptr(), b.data_ptr(), cmax, C.byref(tab),
               ids_t.data_ptr(), B, h, w, dG.data_ptr(), stride, ws.data_ptr(), nb, ops._stream())
        outs.append(dG.clone())
    tol = 2e-5 if dt == torch.float32 else 2e-2
    for d, c in enumerate(n_cats):
        sel = [i for i, v in enumerate(ids) if v == d]
        dy = (dyA[sel, :c].double() + dyB[sel, :c].double())
        want = torch.einsum("bnhw,bchw->nc", dy, x[sel].double()).numpy()
        got = outs[0][d, :c * c_uni].view(c, c_uni).cpu().numpy()
        assert rel_err(got, want) <= tol, (d, rel_err(got, want))
        if kinds[d] == "d" and c >= 8:  # tensor-core datasets: no atomics, run-to-run identical
            assert torch.equal(outs[0][d], outs[1][d])
