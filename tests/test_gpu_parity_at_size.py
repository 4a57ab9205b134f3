"""Parity at BASELINE.json's stated sizes against the REFERENCE'S OWN OP SEQUENCE executed by torch in fp32 on the same
GPU (the pattern of test_gpu_fullsize.py::test_full_resolution_slice_vs_reference_ops_on_device):

  cfg1  2 x 19 x 512 x 1024 OhemCELoss(0.7): loss, dlogits and the OHEM-SELECTED SET against torch's own fp32 per-pixel
        losses, with the number of disagreeing pixels printed (expected 0) and every disagreement required to sit
        within a few fp32 ulps of the threshold; the same for the FUSED up-sampling path on a cfg2 slice;
  cfg4  one 171 x 640 x 640 image: fp32 NCHW, fp32 channels_last, bf16;
  cfg5  one ADE image (150 classes) at 1024 x 2048, 6 scales x flip, against probs += softmax(interpolate(...));
  top-k fallback at the full 16 x 1024 x 2048 batch (confident logits): k-th value, tie quota, loss, gradient.

Bars: loss / gradients 1e-5 relative in fp32 (2e-2 bf16); the selected set is compared bit for bit and the mismatches
are bounded by the band |loss_ref - thresh| <= BAND where the two fp32 evaluations may legitimately round apart.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
BAND = 4e-6  # absolute: ~ 8 fp32 ulps at the thresholds 0.357 / 0.916; the per-pixel loss bar of test_gpu_ohem.py


@pytest.fixture(scope="module")
def ops():
    from mdseg_b200 import ops
    torch.backends.cuda.matmul.allow_tf32 = False
    return ops


def onehot_graphs(gen, n_cats, c_uni):
    out = []
    for c in n_cats:
        idx = torch.randint(0, c, (c_uni,), generator=gen)
        idx[:c] = torch.arange(c)
        m = torch.zeros(c, c_uni)
        m[idx, torch.arange(c_uni)] = 1
        out.append(m.to(DEV))
    return out


def report_set(tag, loss_dev, loss_ref, thresh):
    """Selected-set comparison on the device: returns (#mismatches, #mismatches outside the band)."""
    sel_dev, sel_ref = loss_dev > thresh, loss_ref > thresh
    diff = sel_dev != sel_ref
    n_diff = int(diff.sum())
    outside = int((diff & ((loss_ref - thresh).abs() > BAND)).sum())
    in_band = int(((loss_ref - thresh).abs() <= BAND).sum())
    print(f"[selected set] {tag}: {loss_ref.numel()} px, {int(sel_ref.sum())} selected by torch fp32, "
          f"{n_diff} disagree (expected 0), {outside} of them outside the +-{BAND:g} band, {in_band} px inside the band")
    return n_diff, outside


def ref_ohem(losses, labels, thresh):
    n_min = int((labels != 255).sum()) // 16
    hard = losses[losses > thresh]
    if hard.numel() < n_min:
        hard, _ = losses.topk(n_min)
    return hard.mean()


@pytest.mark.parametrize("scale", [3.0, 0.6])
def test_cfg1_full_resolution_selected_set_vs_torch_fp32(ops, scale):
    """BASELINE config 1 (2 x 19 x 512 x 1024).  scale 3.0 is SURVEY 8d's batch; 0.6 puts ~40 % of the pixels within
    0.2 of the threshold so that the set comparison has teeth."""
    g = torch.Generator(device=DEV).manual_seed(1234)
    N, C, H, W = 2, 19, 512, 1024
    logits = torch.randn(N, C, H, W, generator=g, device=DEV) * scale
    labels = torch.randint(0, C, (N, H, W), generator=g, device=DEV)
    labels[torch.rand(N, H, W, generator=g, device=DEV) < 0.05] = 255
    labels[:, :16] = 255
    thresh = ops.neg_log(0.7)
    xd = logits.clone().requires_grad_(True)
    loss = ops.ohem_ce(xd, labels, thresh)
    loss_px = loss.grad_fn.saved_tensors[2].clone()  # before backward frees the saved tensors
    loss.backward()
    xr = logits.clone().requires_grad_(True)
    lr = F.cross_entropy(xr, labels, ignore_index=255, reduction="none").view(-1)
    ref = ref_ohem(lr, labels, thresh)
    ref.backward()
    n_diff, outside = report_set(f"cfg1 full-res scale {scale}", loss_px, lr.detach(), thresh)
    assert outside == 0 and n_diff <= 1e-5 * lr.numel()
    assert float((loss_px - lr.detach()).abs().max()) <= BAND
    assert abs(float(loss) - float(ref)) <= 1e-5 * abs(float(ref))
    assert float((xd.grad - xr.grad).abs().max() / xr.grad.abs().max()) <= 1e-5
    ops.check_errors(DEV)


@pytest.mark.parametrize("scale", [1.0, 0.25])
def test_cfg2_slice_fused_path_selected_set_vs_torch_fp32(ops, scale):
    """The FUSED projection + up-sampling + CE path (ex2.approx / lg2.approx on max-shifted interpolated corners — a
    different operation order from ATen) on four 1024 x 2048 images of BASELINE config 2 (3 datasets, C_uni 67)
    against einsum -> F.interpolate(align_corners=True) -> CrossEntropyLoss(none) in fp32 on the same GPU."""
    n_cats, c_uni = [19, 12, 36], 67
    ids = [0, 1, 2, 2]
    gen = torch.Generator().manual_seed(5)
    graphs = onehot_graphs(gen, n_cats, c_uni)
    g = torch.Generator(device=DEV).manual_seed(99)
    B, h, w, H, W = len(ids), 256, 512, 1024, 2048
    x = torch.randn(B, c_uni, h, w, generator=g, device=DEV) * scale
    labels = torch.empty(B, H, W, dtype=torch.int64, device=DEV)
    for b, d in enumerate(ids):
        labels[b] = torch.randint(0, n_cats[d], (H, W), generator=g, device=DEV)
    labels[torch.rand(B, H, W, generator=g, device=DEV) < 0.05] = 255
    ids_t = torch.tensor(ids, dtype=torch.int32, device=DEV)
    thresh = ops.neg_log(0.4)
    xd = x.clone().requires_grad_(True)
    loss = ops.mds_proj_ohem_ce(xd, labels, ids_t, graphs, thresh)
    loss_px = loss.grad_fn.saved_tensors[4].clone()
    loss.backward()
    xr = x.clone().requires_grad_(True)
    parts = []
    for i in range(3):
        sel = ids_t == i
        r = F.interpolate(torch.einsum("bchw, nc -> bnhw", xr[sel], graphs[i]), size=(H, W), mode="bilinear",
                          align_corners=True)
        parts.append(F.cross_entropy(r, labels[sel], ignore_index=255, reduction="none").view(-1))
    lr = torch.cat(parts)
    ref = ref_ohem(lr, labels, thresh)
    ref.backward()
    n_diff, outside = report_set(f"cfg2 fused path scale {scale}", loss_px, lr.detach(), thresh)
    assert outside == 0 and n_diff <= 1e-5 * lr.numel()
    assert float((loss_px - lr.detach()).abs().max()) <= BAND * max(1.0, float(lr.max()) / 4)
    assert abs(float(loss) - float(ref)) <= 1e-5 * abs(float(ref))
    assert float((xd.grad - xr.grad).abs().max() / xr.grad.abs().max()) <= 1e-5
    ops.check_errors(DEV)


@pytest.mark.parametrize("variant,tol", [("f32_nchw", 1e-5), ("f32_nhwc", 1e-5), ("bf16_nchw", 2e-2)])
def test_cfg4_one_image_171_classes(ops, variant, tol):
    """BASELINE config 4 (bisenetv2_coco: 171 classes, 640 x 640), one image per layout / dtype."""
    g = torch.Generator(device=DEV).manual_seed(4)
    N, C, H, W = 1, 171, 640, 640
    logits = torch.randn(N, C, H, W, generator=g, device=DEV) * 2.0
    labels = torch.randint(0, C, (N, H, W), generator=g, device=DEV)
    labels[torch.rand(N, H, W, generator=g, device=DEV) < 0.05] = 255
    if variant == "bf16_nchw":
        logits = logits.to(torch.bfloat16)
    if variant == "f32_nhwc":
        logits = logits.contiguous(memory_format=torch.channels_last)
    thresh = ops.neg_log(0.7)
    xd = logits.clone(memory_format=torch.preserve_format).requires_grad_(True)
    loss = ops.ohem_ce(xd, labels, thresh)
    loss_px = loss.grad_fn.saved_tensors[2].clone()
    loss.backward()
    xr = logits.float().contiguous().requires_grad_(True)  # the reference computes the CE in fp32 (autocast)
    lr = F.cross_entropy(xr, labels, ignore_index=255, reduction="none").view(-1)
    ref = ref_ohem(lr, labels, thresh)
    ref.backward()
    assert abs(float(loss) - float(ref)) <= tol * abs(float(ref))
    assert float((xd.grad.float() - xr.grad).abs().max() / xr.grad.abs().max()) <= tol
    if variant != "bf16_nchw":
        n_diff, outside = report_set(f"cfg4 {variant}", loss_px, lr.detach(), thresh)
        assert outside == 0
    ops.check_errors(DEV)


def test_cfg5_one_ade_image_six_scales_and_flip(ops):
    """BASELINE config 5: one ADE image (150 classes) at 1024 x 2048, scales 0.5 ... 1.75 with and without flip
    (12 passes), against probs += softmax(F.interpolate(logits, size, bilinear, align_corners=True)) and argmax
    (evaluate.py:136-181) run by torch on the same GPU; histogram against torch.bincount of the device's own
    predictions."""
    g = torch.Generator(device=DEV).manual_seed(55)
    C, H, W = 150, 1024, 2048
    passes, flips = [], []
    for s in (0.5, 0.75, 1.0, 1.25, 1.5, 1.75):
        hs, ws = int(round(s * H / 32)) * 32 // 4, int(round(s * W / 32)) * 32 // 4
        for fl in (False, True):
            passes.append(torch.randn(1, C, hs, ws, generator=g, device=DEV) * 2.0)
            flips.append(fl)
    label = torch.randint(0, C, (H, W), generator=g, device=DEV)
    label[torch.rand(H, W, generator=g, device=DEV) < 0.05] = 255
    pred, hist = ops.eval_fused(list(zip(passes, flips)), (H, W), label=label)
    probs = torch.zeros(1, C, H, W, device=DEV)
    for lg, fl in zip(passes, flips):
        if fl:
            lg = torch.flip(lg, dims=(3,))
        probs += torch.softmax(F.interpolate(lg, size=(H, W), mode="bilinear", align_corners=True), dim=1)
    ref_pred = probs.argmax(1)[0]
    top2 = probs[0].topk(2, dim=0).values
    clear = (top2[0] - top2[1]) > 1e-5
    n_diff = int((pred != ref_pred).sum())
    print(f"[cfg5] {H * W} px, {n_diff} predictions differ from torch's, {int((~clear).sum())} px have a top-2 gap "
          "below 1e-5")
    assert torch.equal(pred[clear], ref_pred[clear]) and float(clear.float().mean()) > 0.999
    keep = label != 255
    want = torch.bincount(label[keep] * C + pred[keep], minlength=C * C).view(C, C)
    assert torch.equal(hist, want)
    ops.check_errors(DEV)


def test_topk_fallback_at_the_full_cfg3_batch(ops):
    """SURVEY 8d variant 3 at size: 16 x 1024 x 2048 labels, 7 datasets, block-constant labels that the logits
    predict, so fewer than n_min pixels are hard and the radix select runs over all 33.5 M losses: mode, k-th value,
    tie quota and the sum over the selected multiset are exact functions of the device's loss vector; the loss equals
    torch.topk's mean; exactly the selected pixels carry gradient."""
    import bench
    bt = bench.make_batch("cfg3", DEV, 1234, logits="confident")
    n_cats, ids = bt["n_cats"], bt["ids"]
    luts = torch.from_numpy(np.stack(bt["luts"])).to(DEV)
    ids_t = torch.tensor(ids, dtype=torch.int32, device=DEV)
    labels = ops.lut_remap_images(bt["raw"], luts, ids_t, out_dtype=torch.uint8)
    graphs = [gph.to(DEV) for gph in bt["graphs"]]
    thresh = ops.neg_log(0.4)
    x = bt["x"].requires_grad_(True)
    loss = ops.mds_proj_ohem_ce(x, labels, ids_t, graphs, thresh)
    v = loss.grad_fn.saved_tensors[4].clone()  # ties that missed the quota were demoted to just below kth, in place
    loss.backward()
    st = ops.read_states(loss.grad_fn.states)[0]
    n_valid = int((labels != 255).sum())
    k = n_valid // 16
    assert st.mode == 1 and st.n_valid == n_valid and st.n_min == k and st.n_sel == k and st.n_hard < k
    kth = torch.tensor(st.kth, dtype=torch.float32, device=DEV)
    n_gt, n_eq = int((v > kth).sum()), int((v == kth).sum())
    assert n_gt == st.n_gt and n_gt < k <= n_gt + st.n_ties
    assert st.tie_quota == k - n_gt and n_eq == st.tie_quota
    want_sum = float(v[v > kth].double().sum()) + float(kth) * st.tie_quota
    assert abs(st.sum_sel - want_sum) <= 1e-9 * want_sum
    assert abs(float(loss) - want_sum / k) <= 1e-6 * want_sum / k
    ref_topk = float(v.topk(k).values.double().mean())  # torch's own top-k over the same vector
    assert abs(float(loss) - ref_topk) <= 1e-6 * ref_topk
    # gradient: only selected pixels contribute.  Unified channels of a class get the adjoint of the up-sampling of
    # w * (softmax - onehot): zero wherever no selected pixel lies under the tent of the low-res cell
    sel = (v >= kth).view(len(ids), bt["H"], bt["W"]) & (labels != 255)
    assert int(sel.sum()) == k
    touched = F.max_pool2d(sel.float()[:, None], kernel_size=16, stride=4, padding=6)[:, 0] > 0  # generous tent cover
    g_any = x.grad.abs().amax(dim=1) > 0
    assert not bool((g_any & ~touched[:, :bt["h"], :bt["w"]]).any())
    assert torch.isfinite(x.grad).all() and float(x.grad.abs().max()) > 0
    print(f"[top-k at size] n_valid {n_valid}, n_min {k}, n_hard {st.n_hard}, kth {st.kth:.6f}, ties {st.n_ties}, "
          f"quota {st.tie_quota}")
    ops.check_errors(DEV)
