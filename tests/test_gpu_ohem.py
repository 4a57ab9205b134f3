"""GPU parity: OhemCELoss path (full-resolution CE fwd, selection, bwd) against the real reference's golden
outputs, the torch restatement and the float64 oracle.  Tolerances: loss and dlogits 1e-5 relative (fp32),
2e-2 (bf16) — BASELINE.json north_star; selected set exact outside the near-threshold band."""
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


def make_case(seed, N, C, H, W, scale=3.0, conf=0.0, p_ign=0.05):
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(N, C, H, W, generator=g) * scale
    labels = torch.randint(0, C, (N, H, W), generator=g)
    labels[torch.rand(N, H, W, generator=g) < p_ign] = 255
    if conf:
        lab0 = labels.clone(); lab0[lab0 == 255] = 0
        boost = torch.zeros_like(logits).scatter_(1, lab0.unsqueeze(1), conf)
        logits = logits + boost * (torch.rand(N, 1, H, W, generator=g) < 0.97)
    return logits, labels


@pytest.mark.parametrize("name", ["thresh", "topk"])
def test_golden_reference_outputs(ops, golden, name):
    z = golden("ohem_ce.npz")
    logits = torch.from_numpy(z[f"ohem_{name}_logits"]).to(DEV).requires_grad_(True)
    labels = torch.from_numpy(z[f"ohem_{name}_labels"]).to(DEV)
    loss = ops.ohem_ce(logits, labels, ops.neg_log(0.7))
    (loss * 3.0).backward()
    ops.check_errors(DEV)
    want = float(z[f"ohem_{name}_loss"])
    assert abs(float(loss) - want) <= RTOL32 * abs(want)
    assert rel_err(logits.grad.cpu().numpy(), z[f"ohem_{name}_dlogits"]) <= RTOL32


def test_all_ignored_gives_nan_and_zero_grad(ops, golden):
    z = golden("ohem_ce.npz")
    logits = torch.from_numpy(z["ohem_allign_logits"]).to(DEV).requires_grad_(True)
    loss = ops.ohem_ce(logits, torch.from_numpy(z["ohem_allign_labels"]).to(DEV), ops.neg_log(0.7))
    assert torch.isnan(loss) and np.isnan(z["ohem_allign_loss"])  # torch.mean of an empty selection
    loss.backward()
    assert float(logits.grad.abs().max()) == 0.0


@pytest.mark.parametrize("shape", [(2, 19, 64, 128), (1, 171, 40, 40), (3, 5, 17, 23), (1, 2, 1, 1), (2, 33, 8, 12)])
@pytest.mark.parametrize("mode", ["thresh", "topk"])
@pytest.mark.parametrize("lab_dt", [torch.int64, torch.uint8])
def test_fp32_vs_f64(ops, shape, mode, lab_dt):
    N, C, H, W = shape
    logits, labels = make_case(17 + C, N, C, H, W, scale=3.0 if mode == "thresh" else 1.0,
                               conf=0.0 if mode == "thresh" else 12.0)
    thresh = ops.neg_log(0.7)
    mean, dl, loss_px, mask = f64.ohem_ce(logits.numpy(), labels.numpy(), thresh)
    xd = logits.to(DEV).requires_grad_(True)
    loss = ops.ohem_ce(xd, labels.to(lab_dt).to(DEV), thresh)
    loss.backward()
    ops.check_errors(DEV)
    if np.isnan(mean):
        assert torch.isnan(loss)
        return
    assert abs(float(loss) - mean) <= RTOL32 * abs(mean)
    got = xd.grad.cpu().numpy()
    if mode == "topk":
        # fp32 losses near the k-th value tie (they are quantised to ulp(lse)); which of the tied pixels enter
        # the top-k set is unspecified (torch.topk too), so pixels within a few ulp of the k-th value are
        # compared separately: selected-or-not, nothing else.
        kth = np.sort(loss_px.reshape(-1))[::-1][int((labels.numpy() != 255).sum()) // 16 - 1]
        band = np.abs(loss_px - kth) <= 2e-5
        assert band.mean() < 0.05
        out = ~np.broadcast_to(band[:, None], dl.shape)
        assert np.abs(got - dl)[out].max() <= RTOL32 * np.abs(dl).max()
        n_sel = max(int(mask.sum()), 1)
        w_full = np.abs(dl).max(axis=1)  # |dlogit| of a band pixel is either ~0 or its selected value
        _, p = f64.ce_per_pixel(logits.numpy(), labels.numpy())
        lab0 = np.where(labels.numpy() == 255, 0, labels.numpy())
        sel_val = (1.0 - np.take_along_axis(p, lab0[:, None], 1)[:, 0]) / n_sel
        g_lab = np.abs(np.take_along_axis(got, lab0[:, None], 1)[:, 0])
        ok = (np.abs(g_lab - sel_val) <= RTOL32 / n_sel) | (g_lab == 0)  # tolerance relative to max |dlogit| = 1/n_sel
        assert ok[band & (labels.numpy() != 255)].all()
    else:
        assert rel_err(got, dl) <= RTOL32


def test_selected_set_exact_outside_band(ops):
    """The OHEM set {p: loss_p > thresh}: bit-exact vs the float64 oracle for every pixel whose loss is not within
    a few fp32 ulps of the threshold; in-band disagreements are counted and must be rare."""
    N, C, H, W = 2, 19, 256, 512
    logits, labels = make_case(3, N, C, H, W, scale=1.5)
    thresh = ops.neg_log(0.7)
    loss, loss_px, st = ops.ohem_ce_with_state(logits.to(DEV), labels.to(DEV), thresh)
    got = loss_px.cpu().numpy().reshape(N, H, W)
    lp64, _ = f64.ce_per_pixel(logits.numpy().astype(np.float64), labels.numpy())
    band = f64.threshold_band(lp64, thresh)
    sel_dev, sel_ref = got > np.float32(thresh), lp64 > float(thresh)
    assert np.array_equal(sel_dev[~band], sel_ref[~band])
    assert int((sel_dev != sel_ref).sum()) <= int(band.sum()) and band.mean() < 1e-4
    # per-pixel loss itself: a few ulp
    assert np.abs(got - lp64).max() <= 4e-6
    # the counters the selection is built from are exact functions of the device's own loss vector
    assert st.n_valid == int((labels != 255).sum()) and st.n_px == N * H * W
    assert st.n_hard == int(sel_dev.sum()) and st.mode == 0 and st.n_sel == st.n_hard
    assert abs(st.sum_hard - got[sel_dev].astype(np.float64).sum()) <= 1e-9 * st.sum_hard


@pytest.mark.parametrize("conf,expect_ties", [(12.0, None), (60.0, True)])
def test_topk_selection_is_exact_on_device_values(ops, conf, expect_ties):
    """Top-k fallback: k-th value, #selected, Σ over the selected multiset are exact functions of the loss
    vector (ties at the k-th value included: with conf=60 most losses are exactly 0)."""
    N, C, H, W = 2, 19, 128, 256
    logits, labels = make_case(9, N, C, H, W, scale=1.0, conf=conf, p_ign=0.3)
    thresh = ops.neg_log(0.7)
    loss_val, loss_px, st = ops.ohem_ce_with_state(logits.to(DEV), labels.to(DEV), thresh)
    v = loss_px.cpu().numpy()  # ties that missed the quota were demoted to just below kth, in place
    k = int((labels != 255).sum()) // 16
    assert st.mode == 1 and st.n_min == k and st.n_sel == k
    kth = np.float32(st.kth)
    n_gt, n_eq = int((v > kth).sum()), int((v == kth).sum())
    assert n_gt == st.n_gt and n_gt < k <= n_gt + st.n_ties
    assert st.tie_quota == k - n_gt and n_eq == st.tie_quota
    if expect_ties:  # conf=60: the k-th value is exactly 0 and thousands of pixels tie with it
        assert kth == 0 and st.n_ties > 10 * max(st.tie_quota, 1)
    want_sum = v[v > kth].astype(np.float64).sum() + float(kth) * st.tie_quota
    assert abs(st.sum_sel - want_sum) <= 1e-9 * want_sum + 1e-12
    assert abs(float(loss_val) - want_sum / k) <= 1e-6 * (want_sum / k) + 1e-12
    # the reference (torch.topk on its own loss values) agrees on the mean of the selected multiset
    ref = tr.ohem_ce_loss(logits, labels, 0.7)
    assert abs(float(loss_val) - float(ref)) <= RTOL32 * abs(float(ref)) + 1e-9
    # backward: exactly the selected pixels (k of them, ignored ones contribute nothing) carry gradient
    xd = logits.to(DEV).requires_grad_(True)
    ops.ohem_ce(xd, labels.to(DEV), thresh).backward()
    assert torch.isfinite(xd.grad).all()
    touched = int((xd.grad.abs().amax(dim=1) > 0).sum())
    assert 0 < touched <= k


@pytest.mark.parametrize("dt,tol", [(torch.bfloat16, 2e-2), (torch.float16, 2e-2)])
def test_half_precision_inputs(ops, dt, tol):
    """bf16 / fp16 logits: computed in fp32 from the rounded inputs; compared with the oracle on the SAME rounded
    inputs (tight) and on the fp32 inputs (north_star's 2e-2)."""
    N, C, H, W = 2, 19, 64, 128
    logits, labels = make_case(21, N, C, H, W)
    thresh = ops.neg_log(0.7)
    lq = logits.to(dt)
    mean_q, dl_q, _, _ = f64.ohem_ce(lq.float().numpy(), labels.numpy(), thresh)
    mean, dl, _, _ = f64.ohem_ce(logits.numpy(), labels.numpy(), thresh)
    xd = lq.to(DEV).requires_grad_(True)
    loss = ops.ohem_ce(xd, labels.to(DEV), thresh)
    loss.backward()
    assert abs(float(loss) - mean_q) <= 1e-5 * abs(mean_q)
    assert abs(float(loss) - mean) <= tol * abs(mean)
    # dlogits: against the oracle on the SAME (rounded) inputs, as the reference's bf16/AMP path sees them;
    # the only error left is the rounding of the gradient tensor to its 16-bit dtype.
    assert xd.grad.dtype == dt
    assert rel_err(xd.grad.float().cpu().numpy(), dl_q) <= tol


def test_channels_last_layout(ops):
    N, C, H, W = 2, 19, 32, 48
    logits, labels = make_case(4, N, C, H, W)
    thresh = ops.neg_log(0.7)
    mean, dl, _, _ = f64.ohem_ce(logits.numpy(), labels.numpy(), thresh)
    xd = logits.to(DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    loss = ops.ohem_ce(xd, labels.to(DEV), thresh)
    loss.backward()
    assert abs(float(loss) - mean) <= RTOL32 * abs(mean)
    assert rel_err(xd.grad.cpu().numpy(), dl) <= RTOL32


def test_unaligned_width_uses_scalar_path(ops):
    N, C, H, W = 1, 7, 5, 3  # H*W not a multiple of 4
    logits, labels = make_case(8, N, C, H, W, p_ign=0.0)
    thresh = ops.neg_log(0.7)
    mean, dl, _, _ = f64.ohem_ce(logits.numpy(), labels.numpy(), thresh)
    xd = logits.to(DEV).requires_grad_(True)
    loss = ops.ohem_ce(xd, labels.to(DEV), thresh)
    loss.backward()
    assert abs(float(loss) - mean) <= RTOL32 * abs(mean) and rel_err(xd.grad.cpu().numpy(), dl) <= RTOL32


def test_grad_scale_like_amp(ops):
    """backward receives grad_out = loss scale (GradScaler) as a device scalar."""
    logits, labels = make_case(2, 1, 19, 16, 16)
    thresh = ops.neg_log(0.7)
    a = logits.to(DEV).requires_grad_(True)
    b = logits.to(DEV).requires_grad_(True)
    ops.ohem_ce(a, labels.to(DEV), thresh).backward()
    (ops.ohem_ce(b, labels.to(DEV), thresh) * 1024.0).backward()
    assert torch.allclose(b.grad, a.grad * 1024.0, rtol=1e-6, atol=0)


def test_bad_label_is_flagged(ops):
    logits, labels = make_case(2, 1, 19, 16, 16)
    labels[0, 0, 0] = 19
    ops.ohem_ce(logits.to(DEV), labels.to(DEV), ops.neg_log(0.7))
    with pytest.raises(RuntimeError, match="label out of range"):
        ops.check_errors(DEV)
