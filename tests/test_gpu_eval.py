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
