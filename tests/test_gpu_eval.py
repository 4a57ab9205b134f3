"""GPU parity: evaluator tail — softmax-accumulate over scales/flips, argmax, fused histogram (evaluate.py:136-192)."""
import numpy as np
import pytest
import torch

from oracle import label_space as ls, torch_ref as tr

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ops():
    from mdseg_b200 import ops
    return ops


def test_multi_scale_flip_accumulate_argmax_hist(ops):
    g = torch.Generator().manual_seed(0)
    C, H, W = 19, 96, 160
    sizes = [(12, 20), (24, 40), (32, 56), (24, 40)]
    flips = [False, True, False, True]
    passes = [torch.randn(1, C, h, w, generator=g) * 3 for (h, w) in sizes]
    label = torch.randint(0, C, (1, H, W), generator=g)
    label[torch.rand(1, H, W, generator=g) < 0.1] = 255
    want_probs = tr.eval_probs(passes, (H, W), flips)
    probs = torch.empty(C, H, W, device=DEV)
    for i, (lg, fl) in enumerate(zip(passes, flips)):
        ops.eval_accum(lg.to(DEV), probs, flip=fl, first=(i == 0))
    assert torch.allclose(probs.cpu(), want_probs[0], rtol=0, atol=2e-6)
    hist = torch.zeros(C, C, dtype=torch.int64, device=DEV)
    pred = ops.argmax_hist(probs, label.to(DEV), hist)
    # argmax / histogram are exact functions of the device's own probabilities
    assert torch.equal(pred.cpu(), torch.argmax(probs.cpu(), dim=0))
    assert np.array_equal(hist.cpu().numpy(), ls.confusion(label.numpy(), pred.cpu().numpy(), C))
    # against the reference's own probabilities: identical except where the top-2 gap is below fp32 noise
    ref_pred = tr.eval_preds(want_probs)[0]
    top2 = want_probs[0].topk(2, dim=0).values
    clear = (top2[0] - top2[1]) > 1e-5
    assert torch.equal(pred.cpu()[clear], ref_pred[clear]) and clear.float().mean() > 0.999
    ops.check_errors(DEV)


def test_low_res_eval_variant(ops):
    """ori_scales=False (evaluate.py:156-164): logits stay at their resolution, the label is nearest-resized."""
    g = torch.Generator().manual_seed(1)
    C, lH, lW, H, W = 12, 32, 64, 256, 512
    logits = torch.randn(1, C, lH, lW, generator=g) * 3
    label = torch.randint(0, C, (1, H, W), generator=g)
    label[torch.rand(1, H, W, generator=g) < 0.1] = 255
    small = ops.label_nearest(label.to(DEV), (lH, lW))
    assert torch.equal(small.cpu(), tr.nearest_label(label, (lH, lW)))
    probs = torch.empty(C, lH, lW, device=DEV)
    ops.eval_accum(logits.to(DEV), probs, first=True)
    assert torch.allclose(probs.cpu(), torch.softmax(logits, 1)[0], rtol=0, atol=1e-6)
    hist = torch.zeros(C, C, dtype=torch.int64, device=DEV)
    pred = ops.argmax_hist(probs, small, hist)
    assert np.array_equal(hist.cpu().numpy(), ls.confusion(small.cpu().numpy(), pred.cpu().numpy(), C))
    iou, miou = ops.miou(hist)
    assert abs(float(miou) - ls.ious_miou(hist.cpu().numpy())[1]) <= 1e-6


def test_wide_class_count_uses_global_histogram(ops):
    g = torch.Generator().manual_seed(2)
    C, H, W = 171, 40, 56
    probs = torch.rand(C, H, W, generator=g).to(DEV)
    label = torch.randint(0, C, (H, W), generator=g).to(DEV)
    hist = torch.zeros(C, C, dtype=torch.int64, device=DEV)
    pred = ops.argmax_hist(probs, label, hist)
    assert np.array_equal(hist.cpu().numpy(), ls.confusion(label.cpu().numpy(), pred.cpu().numpy(), C))


@pytest.mark.parametrize("C,H,W", [(19, 96, 160), (150, 70, 90), (7, 33, 17)])
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("lab_dt", [torch.int64, torch.uint8])
def test_eval_fused_equals_pass_by_pass(ops, C, H, W, dt, lab_dt):
    """mdseg_eval_fused (all passes in one kernel, accumulator in shared memory) == mdseg_eval_accum pass by pass +
    mdseg_argmax_hist, bit for bit: predictions and confusion matrix; ragged tiles, flips, a same-size pass."""
    g = torch.Generator().manual_seed(C + H)
    sizes = [(max(2, H // 8), max(2, W // 8)), (H // 4, W // 4), (H // 3 + 1, W // 3 + 2), (H // 4, W // 4), (H, W)]
    flips = [False, True, False, True, True]
    passes = [((torch.randn(C, h, w, generator=g) * 3).to(dt).to(DEV), f) for (h, w), f in zip(sizes, flips)]
    label = torch.randint(0, C, (H, W), generator=g)
    label[torch.rand(H, W, generator=g) < 0.1] = 255
    label = label.to(lab_dt).to(DEV)
    probs = torch.empty(C, H, W, device=DEV)
    for i, (lg, fl) in enumerate(passes):
        ops.eval_accum(lg, probs, flip=fl, first=(i == 0))
    hist_ref = torch.zeros(C, C, dtype=torch.int64, device=DEV)
    pred_ref = ops.argmax_hist(probs, label, hist_ref)
    hist = torch.zeros(C, C, dtype=torch.int64, device=DEV)
    hist[0, 0] = 3  # accumulate-into semantics
    pred, _ = ops.eval_fused(passes, (H, W), label=label, hist=hist)
    hist[0, 0] -= 3
    assert torch.equal(pred, pred_ref)
    assert torch.equal(hist, hist_ref)
    assert np.array_equal(hist.cpu().numpy(), ls.confusion(label.cpu().numpy().astype(np.int64), pred.cpu().numpy(), C))
    only_pred, none_hist = ops.eval_fused(passes, (H, W))
    assert none_hist is None and torch.equal(only_pred, pred_ref)
    ops.check_errors(DEV)


def test_eval_fused_against_reference_probabilities(ops):
    """Against the reference op sequence (interpolate -> softmax -> sum -> argmax, evaluate.py:149-172) on the CPU:
    identical predictions wherever the top-2 gap is above fp32 noise; LUT-fused raw labels."""
    g = torch.Generator().manual_seed(5)
    C, H, W = 19, 96, 160
    sizes = [(12, 20), (24, 40), (32, 56), (24, 40)]
    flips = [False, True, False, True]
    passes = [torch.randn(1, C, h, w, generator=g) * 3 for (h, w) in sizes]
    want_probs = tr.eval_probs(passes, (H, W), flips)
    raw = torch.randint(0, 34, (H, W), generator=g).to(torch.uint8)
    lut = np.full(256, 255, dtype=np.uint8)
    lut[:34] = np.random.default_rng(0).integers(0, C, 34)
    lut[5] = 255
    pred, hist = ops.eval_fused([(p.to(DEV), f) for p, f in zip(passes, flips)], (H, W), label=raw.to(DEV), lut=lut)
    ref_pred = tr.eval_preds(want_probs)[0]
    top2 = want_probs[0].topk(2, dim=0).values
    clear = (top2[0] - top2[1]) > 1e-5
    assert torch.equal(pred.cpu()[clear], ref_pred[clear]) and clear.float().mean() > 0.999
    assert np.array_equal(hist.cpu().numpy(), ls.confusion(ls.lut_gather(raw.numpy(), lut), pred.cpu().numpy(), C))
    ops.check_errors(DEV)


def test_dropin_evaluator_takes_the_fused_route(ops):
    from mdseg_b200.dropin.evaluate import SegHist
    g = torch.Generator().manual_seed(9)
    C, H, W = 12, 64, 96
    passes = [(torch.randn(C, 16, 24, generator=g).to(DEV), False), (torch.randn(C, 24, 40, generator=g).to(DEV), True)]
    label = torch.randint(0, C, (H, W), generator=g).to(DEV)
    acc = SegHist(C, torch.device(DEV))
    pred = acc.update_from_passes(label, passes)
    probs = torch.empty(C, H, W, device=DEV)
    for i, (lg, fl) in enumerate(passes):
        ops.eval_accum(lg, probs, flip=fl, first=(i == 0))
    want = torch.zeros(C, C, dtype=torch.int64, device=DEV)
    want_pred = ops.argmax_hist(probs, label, want)
    assert torch.equal(pred, want_pred) and torch.equal(acc.hist, want)


def test_eval_fused_with_a_downsampling_pass_takes_the_direct_kernels(ops):
    """A pass larger than the label (scale > stride) cannot use the staged 17 x 17 patches: direct-load kernels."""
    g = torch.Generator().manual_seed(21)
    C, H, W = 23, 40, 56
    passes = [(torch.randn(C, 10, 14, generator=g).to(DEV), False), (torch.randn(C, 50, 70, generator=g).to(DEV), True)]
    label = torch.randint(0, C, (H, W), generator=g).to(DEV)
    probs = torch.empty(C, H, W, device=DEV)
    for i, (lg, fl) in enumerate(passes):
        ops.eval_accum(lg, probs, flip=fl, first=(i == 0))
    want = torch.zeros(C, C, dtype=torch.int64, device=DEV)
    want_pred = ops.argmax_hist(probs, label, want)
    pred, hist = ops.eval_fused(passes, (H, W), label=label)
    assert torch.equal(pred, want_pred) and torch.equal(hist, want)


def test_msc_eval_crop_against_the_reference_op_sequence(ops):
    """MscEvalCrop (evaluate.py:650-753): chips, flip + the reference's exp() of the flipped sum, per-scale bilinear
    resize of the probability map, arg-max, confusion matrix — against the same op sequence in torch on the GPU."""
    import math
    import torch.nn.functional as F
    from mdseg_b200.dropin.evaluate import MscEvalCrop
    g = torch.Generator(device=DEV).manual_seed(8)
    C, H, W = 7, 72, 104
    wgt = torch.randn(C, 3, 3, 3, generator=g, device=DEV)

    def net(x):  # a fixed "network" with full-resolution logits, like the BiSeNet heads (bisenetv2.py:519)
        return (F.conv2d(x, wgt, padding=1) * 2.0,)

    imgs = torch.randn(2, 3, H, W, generator=g, device=DEV)
    label = torch.randint(0, C, (2, 1, H, W), generator=g, device=DEV)
    label[torch.rand(2, 1, H, W, generator=g, device=DEV) < 0.1] = 255
    for flip in (True, False):
        ev = MscEvalCrop(cropsize=48, cropstride=2. / 3, flip=flip, scales=(0.5, 1, 1.5))
        got = ev(net, [(imgs, label)], C)
        # the reference's sequence, verbatim in torch
        hist = torch.zeros(C, C, device=DEV)
        probs = torch.zeros(2, C, H, W, device=DEV)
        for sc in ev.scales:
            im = F.interpolate(imgs, [int(H * sc), int(W * sc)], mode='bilinear', align_corners=True)
            im, (hst, hed, wst, wed) = ev.pad_tensor(im)
            PH, PW = im.shape[-2:]
            prob = torch.zeros(2, C, PH, PW, device=DEV)
            strd = math.ceil(48 * 2. / 3)
            n_h, n_w = math.ceil((PH - 48) / strd) + 1, math.ceil((PW - 48) / strd) + 1
            for i in range(n_h):
                for j in range(n_w):
                    endH, endW = min(PH, strd * i + 48), min(PW, strd * j + 48)
                    stH, stW = endH - 48, endW - 48
                    chip = im[:, :, stH:endH, stW:endW]
                    p = net(chip)[0].softmax(dim=1)
                    if flip:
                        p = p + net(torch.flip(chip, dims=(3,)))[0].flip(dims=(3,)).softmax(dim=1)
                        p = torch.exp(p)
                    prob[:, :, stH:endH, stW:endW] += p
            probs += F.interpolate(prob[:, :, hst:hed, wst:wed], (H, W), mode='bilinear', align_corners=True)
        preds = probs.argmax(1)
        lab = label.squeeze(1)
        keep = lab != 255
        hist += torch.bincount(lab[keep] * C + preds[keep], minlength=C * C).view(C, C)
        ious = hist.diag() / (hist.sum(0) + hist.sum(1) - hist.diag())
        want = float(np.nanmean(ious.cpu().numpy()))
        assert abs(got - want) <= 2e-3, (flip, got, want)   # a handful of near-tie pixels may flip their arg-max
    ops.check_errors(DEV)
