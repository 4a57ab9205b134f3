"""Host side of the GPU label pipeline: the random draws of the reference's training transforms,
turned into the integers ``ops.label_pipeline`` consumes.

The reference transforms one (image, label) pair at a time on a DataLoader worker (lib/transform_cv2.py): the label
goes through ``lb_map`` (lib/base_dataset.py:81-82), ``RandomResizedCrop`` (nearest resize, 255-padding, crop,
:14-62) and ``RandomHorizontalFlip`` (:66-77).  Only the *numbers* are decided on the host here — with the same
``np.random`` calls in the same order, so a seeded run crops exactly where the reference would — and the pixels are
produced by one kernel for the whole batch, uint8 in HBM end to end.
"""
import math

import numpy as np

from .. import ops


class RandomResizedCropPlan:
    """``RandomResizedCrop(scales, size)`` of lib/transform_cv2.py:14-62, label branch: ``plan(shape)`` draws
    ``np.random.uniform`` (scale) and ``np.random.random(2)`` (crop origin) like ``__call__`` does."""

    def __init__(self, scales=(0.5, 1.), size=(384, 384)):
        self.scales, self.size = scales, size

    def plan(self, shape, rng=np.random):
        H, W = shape
        crop_h, crop_w = self.size
        scale = rng.uniform(min(self.scales), max(self.scales))
        if np.min([H, W]) < 1080:  # :36-37
            scale = scale * (1080 / np.min([H, W]))
        im_h, im_w = [math.ceil(el * scale) for el in (H, W)]
        if (im_h, im_w) == (crop_h, crop_w):  # :45, no crop-origin draw in this case
            return dict(im_h=im_h, im_w=im_w, pad_top=0, pad_left=0, crop_y=0, crop_x=0)
        pad_h = (crop_h - im_h) // 2 + 1 if im_h < crop_h else 0
        pad_w = (crop_w - im_w) // 2 + 1 if im_w < crop_w else 0
        sh, sw = rng.random(2)
        sh, sw = int(sh * (im_h + 2 * pad_h - crop_h)), int(sw * (im_w + 2 * pad_w - crop_w))
        return dict(im_h=im_h, im_w=im_w, pad_top=pad_h, pad_left=pad_w, crop_y=sh, crop_x=sw)


class RandomHorizontalFlipPlan:
    """``RandomHorizontalFlip(p)`` of lib/transform_cv2.py:66-77: one ``np.random.random()`` draw; the pair is
    flipped when the draw is NOT below p (the reference returns the input unchanged when it is)."""

    def __init__(self, p=0.5):
        self.p = p

    def plan(self, rng=np.random):
        return not (rng.random() < self.p)


class LabelPipeline:
    """Batch-level label branch: ``LabelPipeline(scales, size, p)(raw_labels, lut_ids)`` -> ``[B, crop_h, crop_w]``.

    raw_labels: list of uint8 CUDA tensors as decoded from disk (``cv2.imread(lbpth, 0)``); luts: uint8
    ``[n_datasets, 256]`` holding every dataset's ``lb_map``; lut_ids[b] = dataset of image b.  The per-image order
    of random draws is the reference's per-sample order: crop draws, then the flip draw (``TransformationTrain``,
    lib/transform_cv2.py:257-270, skipping ColorJitter which only draws for the image)."""

    def __init__(self, scales, size, p=0.5, luts=None, out_dtype=None, image_draws=3):
        import torch
        self.crop, self.flip, self.size = RandomResizedCropPlan(scales, size), RandomHorizontalFlipPlan(p), size
        self.luts, self.out_dtype = luts, out_dtype or torch.int64
        # ColorJitter(brightness, contrast, saturation) draws one np.random.uniform per enabled term after the flip
        # (lib/transform_cv2.py:94-103); they are drawn and dropped here so that the stream stays aligned with a
        # reference worker that also transforms the image
        self.image_draws = image_draws

    def plans(self, shapes, rng=np.random):
        out = []
        for shape in shapes:
            pl = self.crop.plan(shape, rng)
            pl["flip"] = self.flip.plan(rng)
            for _ in range(self.image_draws):
                rng.uniform(0.0, 1.0)
            out.append(pl)
        return out

    def __call__(self, raw_labels, lut_ids=None, rng=np.random, plans=None):
        plans = plans or self.plans([tuple(t.shape) for t in raw_labels], rng)
        return ops.label_pipeline(raw_labels, plans, self.size, luts=self.luts, lut_ids=lut_ids,
                                  out_dtype=self.out_dtype)
