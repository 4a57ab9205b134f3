"""GPU parity: the fused multi-dataset path (projection -> bilinear upsample -> CE -> one OHEM selection ->
adjoint) and the per-dataset aux heads, against the real reference's golden outputs and the float64 oracle."""
import numpy as np
import pytest
import torch

from oracle import f64, torch_ref as tr

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
RTOL32 = 1e-5


@pytest.fixture(scope="module")
def ops():
    from mdseg_b200 import ops
    return ops


def rel_err(a, b):
    return np.abs(np.asarray(a, dtype=np.float64) - b).max() / max(np.abs(b).max(), 1e-30)


def onehot_graph(g, c_ds, c_uni):
    idx = torch.randint(0, c_ds, (c_uni,), generator=g)
    idx[:c_ds] = torch.arange(c_ds)
    m = torch.zeros(c_ds, c_uni)
    m[idx, torch.arange(c_uni)] = 1
    return m


def make_mds(seed, n_cats, c_uni, ids, h, w, H, W, scale=2.5, p_ign=0.05, graph="onehot"):
    g = torch.Generator().manual_seed(seed)
    B = len(ids)
    x = torch.randn(B, c_uni, h, w, generator=g) * scale
    if graph == "onehot":
        graphs = [onehot_graph(g, c, c_uni) for c in n_cats]
    elif graph == "sparse01":  # general 0/1 remap matrix: a unified id may serve several dataset classes
        graphs = [(torch.rand(c, c_uni, generator=g) < 0.15).float() for c in n_cats]
    else:
        graphs = [torch.softmax(torch.randn(c, c_uni, generator=g) * 4, dim=0) for c in n_cats]
    labels = torch.full((B, H, W), 255, dtype=torch.long)
    for b, d in enumerate(ids):
        if 0 <= d < len(n_cats):
            labels[b] = torch.randint(0, n_cats[d], (H, W), generator=g)
    labels[torch.rand(B, H, W, generator=g) < p_ign] = 255
    return x, graphs, labels


@pytest.mark.parametrize("name", ["sorted", "shuffled", "absent"])
def test_golden_reference_outputs(ops, golden, name):
    z = golden("mds.npz")
    x = torch.from_numpy(z[f"mds_{name}_x"]).to(DEV).requires_grad_(True)
    graphs = [torch.from_numpy(z[f"mds_{name}_graph{i}"]).to(DEV) for i in range(3)]
    loss = ops.mds_proj_ohem_ce(x, torch.from_numpy(z[f"mds_{name}_labels"]).to(DEV),
                                torch.from_numpy(z[f"mds_{name}_ids"]).to(DEV), graphs, ops.neg_log(0.4))
    (loss * 2.0).backward()
    ops.check_errors(DEV)
    want = float(z[f"mds_{name}_loss"])
    assert abs(float(loss) - want) <= RTOL32 * abs(want)
    assert rel_err(x.grad.cpu().numpy(), z[f"mds_{name}_dx"]) <= RTOL32


def test_golden_dense_graphs_with_grad(ops, golden):
    z = golden("mds.npz")
    x = torch.from_numpy(z["mds_dense_x"]).to(DEV).requires_grad_(True)
    graphs = [torch.from_numpy(z[f"mds_dense_graph{i}"]).to(DEV).requires_grad_(True) for i in range(3)]
    loss = ops.mds_proj_ohem_ce(x, torch.from_numpy(z["mds_dense_labels"]).to(DEV),
                                torch.from_numpy(z["mds_dense_ids"]).to(DEV), graphs, ops.neg_log(0.4))
    loss.backward()
    ops.check_errors(DEV)
    want = float(z["mds_dense_loss"])
    assert abs(float(loss) - want) <= RTOL32 * abs(want)
    assert rel_err(x.grad.cpu().numpy(), z["mds_dense_dx"]) <= RTOL32
    for i in range(3):
        assert rel_err(graphs[i].grad.cpu().numpy(), z[f"mds_dense_dgraph{i}"]) <= 2e-5


GEOMS = [
    # (h, w, H, W): stride-4 crops, odd sizes, non-integer ratios, up to identity and down-sampling
    (16, 32, 64, 128), (24, 24, 96, 96), (7, 9, 25, 33), (33, 45, 130, 177), (8, 300, 29, 1200),
    (5, 6, 5, 6), (12, 10, 7, 9), (1, 1, 4, 4), (3, 140, 9, 520), (40, 40, 70, 70), (6, 8, 96, 128),
    # w % 32 == 0: the backward runs as row CTAs (one warp per 32 columns, neighbours exchange the seam column):
    # 2 / 3 / 5 / 16 warps, several 16-row segments, ratios with 4- and 5-pixel cells, a ratio below 4
    (20, 64, 80, 256), (35, 96, 140, 384), (6, 160, 21, 608), (18, 512, 69, 2048), (9, 64, 45, 320),
]


@pytest.mark.parametrize("geom", GEOMS)
@pytest.mark.parametrize("graph", ["onehot", "sparse01", "dense"])
def test_fp32_vs_f64_geometries(ops, geom, graph):
    h, w, H, W = geom
    n_cats, c_uni, ids = [5, 3, 7], 11, [2, 0, 1, 2, 0]
    x, graphs, labels = make_mds(h * 131 + W, n_cats, c_uni, ids, h, w, H, W, graph=graph)
    thresh = ops.neg_log(0.4)
    ref = f64.multi_dataset(x.numpy(), labels.numpy(), np.array(ids), [m.numpy() for m in graphs], thresh)
    xd = x.to(DEV).requires_grad_(True)
    loss = ops.mds_proj_ohem_ce(xd, labels.to(DEV), torch.tensor(ids, dtype=torch.int32, device=DEV),
                                [m.to(DEV) for m in graphs], thresh)
    loss.backward()
    ops.check_errors(DEV)
    assert abs(float(loss) - ref["loss"]) <= RTOL32 * abs(ref["loss"])
    assert rel_err(xd.grad.cpu().numpy(), ref["dlogits_uni"]) <= RTOL32


@pytest.mark.parametrize("c_case", ["wide", "chunked", "wide_rows"])
def test_many_classes_and_class_chunking(ops, c_case):
    """C_ds = 150 / 133 as in the 7-dataset config: exercises the staged class chunks of the fused kernels."""
    n_cats, c_uni = ([150, 19, 133], 358) if c_case.startswith("wide") else ([64, 37, 26], 127)
    ids = [0, 2, 1, 0]
    h, w, H, W = (10, 14, 37, 53) if c_case != "wide_rows" else (19, 64, 76, 256)  # row CTAs, 2 warps, 2 segments
    x, graphs, labels = make_mds(77, n_cats, c_uni, ids, h, w, H, W, scale=2.0)
    thresh = ops.neg_log(0.4)
    ref = f64.multi_dataset(x.numpy(), labels.numpy(), np.array(ids), [m.numpy() for m in graphs], thresh)
    xd = x.to(DEV).requires_grad_(True)
    loss = ops.mds_proj_ohem_ce(xd, labels.to(DEV), torch.tensor(ids, device=DEV), [m.to(DEV) for m in graphs], thresh)
    loss.backward()
    ops.check_errors(DEV)
    assert abs(float(loss) - ref["loss"]) <= RTOL32 * abs(ref["loss"])
    assert rel_err(xd.grad.cpu().numpy(), ref["dlogits_uni"]) <= RTOL32


@pytest.mark.parametrize("geom", [(12, 16, 48, 64), (19, 64, 76, 256)])  # one-warp CTAs / row CTAs
def test_unified_classes_without_a_dataset_class_get_zero_gradient(ops, geom):
    """UOT / pretrain graphs leave some unified classes unmapped for a dataset (empty CSC columns): their channels
    of dlogits_uni must be written as zeros by the fused backward (the output buffer is uninitialised)."""
    h, w, H, W = geom
    n_cats, c_uni, ids = [5, 3, 7], 13, [2, 0, 1, 2]
    x, graphs, labels = make_mds(9, n_cats, c_uni, ids, h, w, H, W)
    for d, m in enumerate(graphs):
        m[:, [1 + d, 8, 12]] = 0  # unmapped unified classes (every dataset class keeps a channel: :c_ds are identity)
        m[:, 0] = 0
        m[0, 0] = 1
    thresh = ops.neg_log(0.4)
    ref = f64.multi_dataset(x.numpy(), labels.numpy(), np.array(ids), [m.numpy() for m in graphs], thresh)
    xd = x.to(DEV).requires_grad_(True)
    torch.empty(1 << 22, device=DEV).fill_(float("nan"))  # poison the allocator's free blocks
    loss = ops.mds_proj_ohem_ce(xd, labels.to(DEV), torch.tensor(ids, device=DEV), [m.to(DEV) for m in graphs], thresh)
    loss.backward()
    ops.check_errors(DEV)
    assert abs(float(loss) - ref["loss"]) <= RTOL32 * abs(ref["loss"])
    assert rel_err(xd.grad.cpu().numpy(), ref["dlogits_uni"]) <= RTOL32
    assert float(xd.grad[:, 8].abs().max()) == 0.0 and float(xd.grad[:, 12].abs().max()) == 0.0


@pytest.mark.parametrize("geom", [(12, 16, 48, 64), (19, 64, 76, 256)])  # one-warp CTAs / row CTAs
def test_topk_branch_of_the_fused_path(ops, geom):
    """Confident predictions: fewer hard pixels than n_min -> device-side top-k fallback."""
    n_cats, c_uni, ids = [5, 3, 7], 15, [0, 1, 2, 2]
    g = torch.Generator().manual_seed(1)
    B, (h, w, H, W) = 4, geom
    graphs = [onehot_graph(g, c, c_uni) for c in n_cats]
    x = torch.randn(B, c_uni, h, w, generator=g) * 0.3
    labels = torch.empty(B, H, W, dtype=torch.long)
    for b, d in enumerate(ids):  # one confident class per image, 2 % of the pixels labelled otherwise (hard)
        c = int(torch.randint(0, n_cats[d], (1,), generator=g))
        x[b, int(graphs[d][c].nonzero()[0, 0])] += 14.0
        labels[b] = c
        other = torch.randint(0, n_cats[d], (H, W), generator=g)
        flip = torch.rand(H, W, generator=g) < 0.02
        labels[b][flip] = other[flip]
    thresh = ops.neg_log(0.4)
    ref = f64.multi_dataset(x.numpy(), labels.numpy(), np.array(ids), [m.numpy() for m in graphs], thresh)
    assert ref["mode"] == 1
    xd = x.to(DEV).requires_grad_(True)
    loss = ops.mds_proj_ohem_ce(xd, labels.to(DEV), torch.tensor(ids, device=DEV), [m.to(DEV) for m in graphs], thresh)
    loss.backward()
    assert abs(float(loss) - ref["loss"]) <= RTOL32 * abs(ref["loss"])
    # pixels tied at the k-th value carry a gradient of ~1e-6 of the hard pixels': tie order is invisible at 5e-5
    assert rel_err(xd.grad.cpu().numpy(), ref["dlogits_uni"]) <= 5e-5


@pytest.mark.parametrize("geom", [(8, 8, 32, 32), (8, 64, 32, 256)])  # one-warp CTAs / row CTAs
def test_invalid_dataset_id_is_skipped_and_flagged(ops, geom):
    """An image whose dataset id matches no dataset takes no part in the loss (ohem_ce_loss.py:58-59) but its
    labels still count in n_min (:52); it gets a zero gradient."""
    n_cats, c_uni, ids = [5, 3], 9, [0, 7, 1]
    x, graphs, labels = make_mds(5, n_cats, c_uni, ids, *geom)
    labels[1] = torch.randint(0, 3, geom[2:])
    thresh = ops.neg_log(0.4)
    ref = f64.multi_dataset(x.numpy(), labels.numpy(), np.array(ids), [m.numpy() for m in graphs], thresh)
    xd = x.to(DEV).requires_grad_(True)
    loss = ops.mds_proj_ohem_ce(xd, labels.to(DEV), torch.tensor(ids, device=DEV), [m.to(DEV) for m in graphs], thresh)
    loss.backward()
    assert abs(float(loss) - ref["loss"]) <= RTOL32 * abs(ref["loss"])
    assert float(xd.grad[1].abs().max()) == 0.0 and rel_err(xd.grad.cpu().numpy(), ref["dlogits_uni"]) <= RTOL32
    with pytest.raises(RuntimeError, match="dataset id"):
        ops.check_errors(DEV)


@pytest.mark.parametrize("geom", [(16, 24, 64, 96), (18, 64, 72, 256)])  # one-warp CTAs / row CTAs
@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
def test_half_precision_logits(ops, dt, geom):
    n_cats, c_uni, ids = [5, 3, 7], 11, [0, 1, 2, 2]
    x, graphs, labels = make_mds(3, n_cats, c_uni, ids, *geom)
    thresh = ops.neg_log(0.4)
    xq = x.to(dt)
    ref_q = f64.multi_dataset(xq.float().numpy(), labels.numpy(), np.array(ids), [m.numpy() for m in graphs], thresh)
    ref = f64.multi_dataset(x.numpy(), labels.numpy(), np.array(ids), [m.numpy() for m in graphs], thresh)
    xd = xq.to(DEV).requires_grad_(True)
    loss = ops.mds_proj_ohem_ce(xd, labels.to(DEV), torch.tensor(ids, device=DEV), [m.to(DEV) for m in graphs], thresh)
    loss.backward()
    assert abs(float(loss) - ref_q["loss"]) <= 1e-5 * abs(ref_q["loss"])
    assert abs(float(loss) - ref["loss"]) <= 2e-2 * abs(ref["loss"])
    # gradient against the oracle on the same rounded inputs (what the reference's AMP path sees)
    assert xd.grad.dtype == dt and rel_err(xd.grad.float().cpu().numpy(), ref_q["dlogits_uni"]) <= 2e-2


def test_uint8_labels_match_int64(ops):
    n_cats, c_uni, ids = [5, 3, 7], 11, [0, 1, 2]
    x, graphs, labels = make_mds(4, n_cats, c_uni, ids, 8, 12, 30, 41)
    thresh = ops.neg_log(0.4)
    gs = [m.to(DEV) for m in graphs]
    a = x.to(DEV).requires_grad_(True)
    b = x.to(DEV).requires_grad_(True)
    la = ops.mds_proj_ohem_ce(a, labels.to(DEV), torch.tensor(ids, device=DEV), gs, thresh)
    lb = ops.mds_proj_ohem_ce(b, labels.to(torch.uint8).to(DEV), torch.tensor(ids, device=DEV), gs, thresh)
    la.backward(); lb.backward()
    assert float(la) == float(lb) and torch.equal(a.grad, b.grad)


def test_advgnn_seg_stage_golden(ops, golden):
    """7 datasets + per-dataset aux heads (loss_cross_datasets.py:1044-1056,1074,1129-1130) vs the real reference."""
    z = golden("advgnn_seg_stage.npz")
    n = len(z["n_cats"])
    x = torch.from_numpy(z["x"]).to(DEV).requires_grad_(True)
    aux = [torch.from_numpy(z[f"aux{i}"]).to(DEV).requires_grad_(True) for i in range(n)]
    graphs = [torch.from_numpy(z[f"graph{i}"]).to(DEV) for i in range(n)]
    labels = torch.from_numpy(z["labels"]).to(DEV)
    ids = torch.from_numpy(z["ids"]).to(DEV)
    main = ops.mds_proj_ohem_ce(x, labels, ids, graphs, ops.neg_log(0.4))
    per_ds = ops.up_ohem_ce(aux, labels, ids, ops.neg_log(0.7), seg_per_dataset=True)
    aux_loss = per_ds.sum()  # every dataset is present in this fixture
    total = main + float(z["aux_weight"]) * aux_loss
    total.backward()
    ops.check_errors(DEV)
    assert abs(float(total) - float(z["loss"])) <= RTOL32 * abs(float(z["loss"]))
    assert abs(float(aux_loss) - float(z["aux_loss"])) <= RTOL32 * abs(float(z["aux_loss"]))
    assert rel_err(x.grad.cpu().numpy(), z["dx"]) <= RTOL32
    for i in range(n):
        assert rel_err(aux[i].grad.cpu().numpy(), z[f"daux{i}"]) <= RTOL32, i


def test_backward_is_deterministic(ops):
    n_cats, c_uni, ids = [19, 12, 36], 67, [0, 1, 2, 2]
    x, graphs, labels = make_mds(6, n_cats, c_uni, ids, 32, 64, 128, 256)
    gs = [m.to(DEV) for m in graphs]
    outs = []
    for _ in range(2):
        xd = x.to(DEV).requires_grad_(True)
        ops.mds_proj_ohem_ce(xd, labels.to(DEV), torch.tensor(ids, device=DEV), gs, ops.neg_log(0.4)).backward()
        outs.append(xd.grad.clone())
    assert torch.equal(outs[0], outs[1])


def test_projection_alone_matches_einsum(ops):
    """Model-side eval projection (lib/models/semseg.py:342-345)."""
    n_cats, c_uni, ids = [19, 12, 36], 67, [0, 1, 2, 1]
    x, graphs, _ = make_mds(8, n_cats, c_uni, ids, 16, 20, 16, 20)
    y = ops.project(x.to(DEV), [m.to(DEV) for m in graphs], torch.tensor(ids, device=DEV))
    for b, d in enumerate(ids):
        want = tr.project(x[b:b + 1], graphs[d])[0]
        assert torch.equal(y[b, :n_cats[d]].cpu(), want)  # 0/1 graphs: exact


@pytest.mark.parametrize("geom", [(16, 32, 64, 128), (9, 12, 36, 48), (7, 9, 25, 33), (20, 64, 80, 256)])
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_aux_heads_per_dataset_selection(ops, geom, dt):
    """Per-dataset aux heads (loss_cross_datasets.py:1044-1056): one OHEM selection per dataset, rows of other
    datasets' images get a zero gradient.  The first geometry takes the fused warp-private kernels (uint8 staged
    labels, direct backward), the others the general route."""
    h, w, H, W = geom
    n_cats, ids = [5, 3, 7], [2, 0, 1, 2, 0, 1]
    B = len(ids)
    g = torch.Generator().manual_seed(h * 17 + W)
    aux = [(torch.randn(B, c, h, w, generator=g) * 2.0).to(dt) for c in n_cats]
    labels = torch.stack([torch.randint(0, n_cats[d], (H, W), generator=g) for d in ids])
    labels[torch.rand(B, H, W, generator=g) < 0.05] = 255
    thresh = ops.neg_log(0.7)
    auxd = [a.to(DEV).requires_grad_(True) for a in aux]
    per_ds = ops.up_ohem_ce(auxd, labels.to(DEV), torch.tensor(ids, device=DEV), thresh, seg_per_dataset=True)
    (per_ds * torch.tensor([1.0, 2.0, 0.5], device=DEV)).sum().backward()
    ops.check_errors(DEV)
    tol = RTOL32 if dt == torch.float32 else 2e-2
    ids_np = np.array(ids)
    for d, c in enumerate(n_cats):
        sel = ids_np == d
        ref_loss, ref_dsrc, _, _ = f64.up_ohem_ce(aux[d].float().numpy()[sel], labels.numpy()[sel], thresh)
        assert abs(float(per_ds[d]) - ref_loss) <= tol * abs(ref_loss), d
        got = auxd[d].grad.float().cpu().numpy()
        scale = [1.0, 2.0, 0.5][d]
        assert rel_err(got[sel], scale * ref_dsrc) <= tol, d
        assert not got[~sel].any(), d
