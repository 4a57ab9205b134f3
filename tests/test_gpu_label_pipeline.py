"""GPU parity: the label branch of the data pipeline in one gather — lb_map LUT, cv2 nearest resize,
255-padding, crop, flip, int64 — against the REAL lib/transform_cv2.py output (tests/golden/label_pipeline.npz) and
against the numpy restatement on random geometry."""
import numpy as np
import pytest
import torch

from oracle import label_space as ls

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ops():
    from mdseg_b200 import ops
    return ops


def test_label_pipeline_against_real_transform_cv2(ops, golden):
    """Same seed, same np.random stream as a reference DataLoader worker -> identical label tensors, bit for bit."""
    from mdseg_b200.dropin.label_transform import LabelPipeline
    z = golden("label_pipeline.npz")
    for tag in ("crop", "pad", "wide"):
        n = int(z[f"{tag}_n"])
        scales = tuple(float(v) for v in z[f"{tag}_scales"])
        size = tuple(int(v) for v in z[f"{tag}_size"])
        raws = [torch.from_numpy(z[f"{tag}_raw{k}"]).to(DEV) for k in range(n)]
        pipe = LabelPipeline(scales, size, p=0.5, luts=torch.from_numpy(z[f"{tag}_lut"]).to(DEV))
        out = pipe(raws, rng=np.random.RandomState(int(z[f"{tag}_seed"])))
        assert out.dtype == torch.int64 and tuple(out.shape) == (n,) + size
        for k in range(n):
            assert np.array_equal(out[k].cpu().numpy(), z[f"{tag}_lb{k}"].astype(np.int64)), (tag, k)


@pytest.mark.parametrize("out_dtype", [torch.int64, torch.uint8])
def test_label_pipeline_random_geometry(ops, out_dtype):
    """Ragged sources, per-image LUTs (and identity), down- and up-scaling, padding on either axis, crops that touch
    the borders, flips — against oracle.label_space.label_transform_chain."""
    rng = np.random.RandomState(11)
    luts = rng.randint(0, 256, (3, 256)).astype(np.uint8)
    size = (48, 80)
    srcs, plans, ids, wants = [], [], [], []
    for b in range(24):
        H, W = int(rng.randint(5, 200)), int(rng.randint(5, 300))
        raw = rng.randint(0, 256, (H, W)).astype(np.uint8)
        im_h, im_w = int(rng.randint(1, 260)), int(rng.randint(1, 400))
        pad_h = (size[0] - im_h) // 2 + 1 if im_h < size[0] else 0
        pad_w = (size[1] - im_w) // 2 + 1 if im_w < size[1] else 0
        hh, ww = im_h + 2 * pad_h - size[0], im_w + 2 * pad_w - size[1]
        edge = b % 3
        plan = dict(im_h=im_h, im_w=im_w, pad_top=pad_h, pad_left=pad_w, flip=bool(b & 1),
                    crop_y=[0, hh, int(rng.randint(0, hh + 1))][edge], crop_x=[0, ww, int(rng.randint(0, ww + 1))][edge])
        lid = b % 4 - 1  # -1: identity
        srcs.append(torch.from_numpy(raw).to(DEV))
        plans.append(plan)
        ids.append(lid)
        wants.append(ls.label_transform_chain(raw, luts[lid] if lid >= 0 else None, plan, size))
    out = ops.label_pipeline(srcs, plans, size, luts=torch.from_numpy(luts).to(DEV), lut_ids=ids, out_dtype=out_dtype)
    assert out.dtype == out_dtype
    for b in range(len(srcs)):
        assert np.array_equal(out[b].cpu().numpy().astype(np.int64), wants[b]), b


def test_label_pipeline_strided_source_and_errors(ops):
    rng = np.random.RandomState(5)
    big = torch.from_numpy(rng.randint(0, 256, (70, 128)).astype(np.uint8)).to(DEV)
    view = big[3:60, 10:100]  # row stride 128, 57 x 90
    plan = dict(im_h=90, im_w=140, pad_top=0, pad_left=0, crop_y=7, crop_x=11, flip=True)
    out = ops.label_pipeline([view], [plan], (64, 112), out_dtype=torch.uint8)
    want = ls.label_transform_chain(view.cpu().numpy(), None, plan, (64, 112))
    assert np.array_equal(out[0].cpu().numpy().astype(np.int64), want)
    from mdseg_b200.native import MdsegError
    with pytest.raises(MdsegError):
        ops.label_pipeline([view], [plan], (64, 100))  # out_w not a multiple of 16
    with pytest.raises(TypeError):
        ops.label_pipeline([view.long()], [plan], (64, 112))


def test_label_pipeline_full_size_batch(ops):
    """16 Cityscapes-sized sources (1024 x 2048) -> 768 x 768 crops (the ltbgnn_7_datasets_snp crop): two images are
    checked against the oracle in full, the rest through the histogram of the output against a LUT-of-histogram bound
    (every output value is either 255 or lut[some source value])."""
    from mdseg_b200.dropin.label_transform import LabelPipeline
    rng = np.random.RandomState(3)
    lut = np.arange(256, dtype=np.uint8)
    lut[34:] = 255
    lut[:34] = rng.randint(0, 19, 34)
    raws = [rng.randint(0, 34, (1024, 2048)).astype(np.uint8) for _ in range(4)]
    srcs = [torch.from_numpy(raws[b % 4]).to(DEV) for b in range(16)]
    pipe = LabelPipeline((0.5, 2.0), (768, 768), luts=torch.from_numpy(lut).to(DEV), out_dtype=torch.uint8)
    plans = pipe.plans([tuple(t.shape) for t in srcs], np.random.RandomState(9))
    out = pipe(srcs, plans=plans)
    for b in (0, 15):
        assert np.array_equal(out[b].cpu().numpy().astype(np.int64), ls.label_transform_chain(raws[b % 4], lut, plans[b], (768, 768)))
    vals = torch.unique(out).cpu().numpy()
    assert set(vals.tolist()) <= set(lut[:34].tolist()) | {255}
