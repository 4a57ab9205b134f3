"""Drop-in for the evaluator tail of evaluate.py (``MscEvalV0`` :45-98 and its multi-scale siblings :100-192).

``MscEvalV0(scales, flip, ignore_label)(net, dl, n_classes, dataset_id) -> float`` keeps the reference
signature.  Per image the reference up-samples every (scale, flip) pass to label size, soft-maxes, sums,
arg-maxes, copies label and prediction to the host and calls ``np.bincount``; here all passes of an image go
through two kernels that keep the probability accumulators in registers, take the arg-max and update an
exact int64 confusion matrix on the device (``ops.eval_fused``; more than 16 passes or mixed dtypes fall
back to one upsample+softmax+accumulate kernel per pass), and the only collective is one all-reduce of that matrix.
"""
import math

import numpy as np
import torch
import torch.distributed as dist
import torch.nn.functional as F

from .. import ops


def get_round_size(size, divisor=32):
    return [math.ceil(el / divisor) * divisor for el in size]


class SegHist:
    """Device-side confusion matrix of one dataset: update(label, logits passes) / all_reduce / miou."""

    def __init__(self, n_classes, device, ignore_label=255, lb_map=None):
        self.n_classes, self.ignore_label, self.lb_map = n_classes, ignore_label, lb_map
        self.hist = torch.zeros(n_classes, n_classes, dtype=torch.int64, device=device)

    @torch.no_grad()
    def update_from_passes(self, label, passes):
        """label [H, W]; passes: iterable of (logits [C, h, w], flip) — evaluate.py:64-93 for one image."""
        H, W = label.shape[-2:]
        passes = list(passes)
        if ops.eval_fused_fits(self.n_classes, len(passes)) and len({p[0].dtype for p in passes}) == 1:
            # every pass of the image at once, no [C, H, W] probability tensor
            return ops.eval_fused(passes, (H, W), label=label.reshape(H, W), hist=self.hist, lut=self.lb_map,
                                  ignore=self.ignore_label)[0]
        probs = torch.empty(self.n_classes, H, W, dtype=torch.float32, device=label.device)
        first = True
        for logits, flip in passes:
            ops.eval_accum(logits, probs, flip=flip, first=first)
            first = False
        return ops.argmax_hist(probs, label=label.reshape(H, W), hist=self.hist, lut=self.lb_map,
                               ignore=self.ignore_label)

    @torch.no_grad()
    def update(self, label, pred):
        """hist += bincount(label[keep] * C + pred[keep]) (evaluate.py:89-93)."""
        ops.confusion(label, pred, self.n_classes, lut=self.lb_map, ignore=self.ignore_label, hist=self.hist)

    def all_reduce(self):
        if dist.is_available() and dist.is_initialized():
            dist.all_reduce(self.hist, dist.ReduceOp.SUM)  # evaluate.py:94-95, exact in int64

    def ious(self):
        return ops.miou(self.hist)

    def miou(self):
        ops.check_errors(self.hist.device)
        return float(self.ious()[1])


class MscEvalV0:
    def __init__(self, scales=(0.5,), flip=False, ignore_label=255):
        self.scales, self.flip, self.ignore_label = scales, flip, ignore_label

    @torch.no_grad()
    def __call__(self, net, dl, n_classes, dataset_id):
        dev = torch.device("cuda", torch.cuda.current_device())
        acc = SegHist(n_classes, dev, self.ignore_label)
        for imgs, label in dl:
            N_, _, H, W = label.shape
            label = label.squeeze(1).to(dev, non_blocking=True)
            per_image = [[] for _ in range(N_)]
            for scale in self.scales:
                sH, sW = get_round_size((int(scale * H), int(scale * W)))
                im_sc = F.interpolate(imgs, size=(sH, sW), mode='bilinear', align_corners=True).to(dev)
                logits = net(im_sc, dataset=dataset_id)[0]
                for b in range(N_):
                    per_image[b].append((logits[b], False))
                if self.flip:
                    logits = net(torch.flip(im_sc, dims=(3,)), dataset=dataset_id)[0]
                    for b in range(N_):
                        per_image[b].append((logits[b], True))  # un-flipped inside the kernel
            for b in range(N_):
                acc.update_from_passes(label[b], per_image[b])
        acc.all_reduce()
        return acc.miou()
