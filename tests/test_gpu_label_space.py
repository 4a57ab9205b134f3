"""GPU parity (bit-exact): LUT remap, ClassRemap-as-LUT, confusion matrix, mIoU, nearest label resize."""
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


@pytest.mark.parametrize("n", [0, 1, 15, 16, 17, 1000, 4096, 123457])
@pytest.mark.parametrize("in_dt,out_dt", [(torch.uint8, torch.uint8), (torch.uint8, torch.int64),
                                          (torch.int64, torch.int64), (torch.int32, torch.uint8),
                                          (torch.int64, torch.int32)])
def test_lut_remap(ops, n, in_dt, out_dt):
    rng = np.random.default_rng(n + 1)
    lut = rng.integers(0, 256, 256, dtype=np.uint8)
    raw = rng.integers(0, 256, n)
    x = torch.from_numpy(raw).to(in_dt).to(DEV)
    if in_dt != torch.uint8 and n > 3:  # values outside [0,255] -> oob (class_remap.py: unmatched -> ignore)
        x[0], x[1], x[2] = -1, 256, 70000
    out = ops.lut_remap(x, lut, out_dtype=out_dt, oob=255)
    xv = x.cpu().numpy().astype(np.int64)
    want = np.where((xv >= 0) & (xv < 256), lut[np.clip(xv, 0, 255)], 255)
    assert out.dtype == out_dt and np.array_equal(out.cpu().numpy().astype(np.int64), want)


def test_lut_remap_unaligned_view(ops):
    rng = np.random.default_rng(0)
    lut = rng.integers(0, 256, 256, dtype=np.uint8)
    base = torch.from_numpy(rng.integers(0, 256, 5001, dtype=np.uint8)).to(DEV)
    x = base[1:]  # 1-byte offset: scalar path
    assert np.array_equal(ops.lut_remap(x.contiguous(), lut).cpu().numpy(), lut[x.cpu().numpy()])


def test_class_remap_luts_on_device(ops, golden):
    """ClassRemap.SingleSegRemapping / SegRemapping / ReverseSegRemap of the REAL reference (golden) through
    the LUT kernel, on the int64 maps the reference uses."""
    import json, os
    z = golden("class_remap.npz")
    root = os.path.dirname(os.path.abspath(__file__))
    for tag, n_ds in (("test", 2), ("cca", 3)):
        raw = json.load(open(os.path.join(root, "golden", f"test_{tag}.json")))
        remaps, max_nums = ls.parse_class_remap(raw, n_ds)
        for d in range(n_ds):
            lb = torch.from_numpy(z[f"{tag}_d{d}_labels"]).to(DEV)
            got = ops.lut_remap(lb, ls.single_seg_lut(remaps[d]), oob=255)
            assert np.array_equal(got.cpu().numpy(), z[f"{tag}_d{d}_single"])
            for j, lut in enumerate(ls.seg_luts(remaps[d], max_nums[d])):
                assert np.array_equal(ops.lut_remap(lb, lut, oob=255).cpu().numpy(), z[f"{tag}_d{d}_seg"][j])
            pr = torch.from_numpy(z[f"{tag}_d{d}_preds"]).to(DEV)
            got = ops.lut_remap(pr, ls.reverse_seg_lut(remaps[d], d), oob=0)
            assert np.array_equal(got.cpu().numpy(), z[f"{tag}_d{d}_reverse"])


def test_multihot_remap_golden(ops, golden):
    """ClassRemapOneHotLabel.SegRemapping / SingleSegRemappingOneHot of the REAL reference (golden) through
    mdseg_multihot_remap."""
    import json, os
    z = golden("multihot.npz")
    root = os.path.dirname(os.path.abspath(__file__))
    for tag, n_ds in (("test", 2), ("cca", 3)):
        raw = json.load(open(os.path.join(root, "golden", f"test_{tag}.json")))
        remaps, _ = ls.parse_class_remap(raw, n_ds)
        cu = raw["num_unify_classes"]
        for d in range(n_ds):
            lb = torch.from_numpy(z[f"{tag}_d{d}_labels"]).to(DEV)
            got = ops.multihot_remap(lb, ls.multihot_table(remaps[d], cu))
            assert got.dtype == torch.bool and np.array_equal(got.cpu().numpy(), z[f"{tag}_d{d}_multi"])
            got = ops.multihot_remap(lb, ls.multihot_table(remaps[d], cu, single_only=True))
            assert np.array_equal(got.cpu().numpy(), z[f"{tag}_d{d}_single"])


@pytest.mark.parametrize("shape", [(1, 1, 1), (2, 5, 7), (3, 64, 96)])
@pytest.mark.parametrize("cu", [4, 46, 358, 500])
@pytest.mark.parametrize("dt", [torch.int64, torch.uint8, torch.int32])
def test_multihot_remap_random_tables(ops, shape, cu, dt):
    """Random 0/1 tables (shared-memory and global-table routes, ragged 16-byte tails), labels with 255 and, for
    the wide types, values outside [0, 255] (-> zero row)."""
    rng = np.random.default_rng(cu + shape[1])
    table = (rng.random((256, cu)) < 0.1).astype(np.uint8)
    lab = rng.integers(0, 256, shape)
    x = torch.from_numpy(lab).to(dt).to(DEV)
    want = table[lab].astype(bool)
    if dt != torch.uint8 and lab.size > 2:
        x.view(-1)[0], x.view(-1)[1] = -3, 4000
        want.reshape(-1, cu)[0] = False
        want.reshape(-1, cu)[1] = False
    got = ops.multihot_remap(x, table)
    assert got.shape == tuple(shape) + (cu,) and np.array_equal(got.cpu().numpy(), want)


@pytest.mark.parametrize("n", [0, 1, 7, 8, 9, 1000, 262144 + 5])
@pytest.mark.parametrize("Ca,Cb", [(19, 19), (150, 150), (150, 358), (2, 3), (300, 300)])
@pytest.mark.parametrize("lab_dt", [torch.int64, torch.uint8])
def test_confusion(ops, n, Ca, Cb, lab_dt):
    if lab_dt == torch.uint8 and Ca > 255:
        pytest.skip("uint8 labels cannot hold this many classes")
    rng = np.random.default_rng(n * 7 + Ca)
    lab = rng.integers(0, min(Ca, 255), n)
    lab[rng.random(n) < 0.07] = 255
    pred = rng.integers(0, Cb, n)
    hist = torch.zeros(Ca, Cb, dtype=torch.int64, device=DEV)
    hist[0, 0] = 5  # accumulate-into semantics
    ops.confusion(torch.from_numpy(lab).to(lab_dt).to(DEV), torch.from_numpy(pred).to(DEV), Ca, Cb, hist=hist)
    want = ls.confusion(lab, pred, Ca, Cb)
    want[0, 0] += 5
    assert np.array_equal(hist.cpu().numpy(), want)
    ops.check_errors(DEV)


def test_confusion_piecewise_constant_and_lut(ops):
    """Segmentation-like maps (long runs) through the run-length path, raw ids + LUT fused."""
    rng = np.random.default_rng(11)
    H, W = 257, 1031
    raw = np.repeat(rng.integers(0, 34, (H, W // 20 + 1)), 20, axis=1)[:, :W].astype(np.uint8)
    lut = np.full(256, 255, dtype=np.uint8)
    lut[:34] = rng.integers(0, 19, 34)
    lut[3] = 255
    pred = np.repeat(rng.integers(0, 19, (H, W // 13 + 1)), 13, axis=1)[:, :W]
    hist = ops.confusion(torch.from_numpy(raw).to(DEV), torch.from_numpy(pred).to(DEV), 19, lut=lut)
    assert np.array_equal(hist.cpu().numpy(), ls.confusion(ls.lut_gather(raw, lut), pred, 19))


@pytest.mark.parametrize("lab_dt,pred_dt", [(torch.int64, torch.int64), (torch.uint8, torch.int64),
                                            (torch.uint8, torch.uint8), (torch.int32, torch.int32),
                                            (torch.int64, torch.uint8), (torch.uint8, torch.int32)])
@pytest.mark.parametrize("mean_run", [1.5, 7, 40, 700])
@pytest.mark.parametrize("offset", [0, 3])
def test_confusion_runs_across_lanes(ops, lab_dt, pred_dt, mean_run, offset):
    """Runs of random length (shorter than a lane's pixels up to several warp chunks) with ignored stretches:
    the warp-level run merge must count every pixel exactly once for every element-width pairing; offset 3
    takes the unaligned (scalar) route."""
    rng = np.random.default_rng(int(mean_run * 10) + offset)
    n = 200003
    def runs(C, p_ignore):
        lens = rng.geometric(1.0 / mean_run, size=int(1.3 * n / mean_run) + 64)
        vals = rng.integers(0, C, lens.size)
        vals[rng.random(lens.size) < p_ignore] = 255
        return np.repeat(vals, lens)[:n + offset]
    lab = runs(37, 0.1)
    pred = np.minimum(runs(37, 0.0), 36)
    assert lab.size == n + offset and pred.size == n + offset
    l = torch.from_numpy(lab).to(lab_dt).to(DEV)[offset:]
    p = torch.from_numpy(pred).to(pred_dt).to(DEV)[offset:]
    hist = ops.confusion(l, p, 37)
    assert np.array_equal(hist.cpu().numpy(), ls.confusion(lab[offset:], pred[offset:], 37))
    ops.check_errors(DEV)


def test_confusion_flags_bad_labels(ops):
    lab = torch.tensor([0, 1, 19, 255, 2], device=DEV)  # 19 is out of range for C=19
    pred = torch.tensor([0, 1, 1, 1, 2], device=DEV)
    hist = ops.confusion(lab, pred, 19)
    assert int(hist.sum()) == 3
    with pytest.raises(RuntimeError, match="label out of range"):
        ops.check_errors(DEV)
    ops.check_errors(DEV)  # flag was cleared


def test_confusion_int32_pred_and_checksum(ops):
    """Linearity: hist(a ++ b) == hist(a) + hist(b); total == #kept pixels."""
    rng = np.random.default_rng(5)
    n = 300001
    lab = rng.integers(0, 19, n); lab[rng.random(n) < 0.1] = 255
    pred = rng.integers(0, 19, n).astype(np.int32)
    l, p = torch.from_numpy(lab).to(DEV), torch.from_numpy(pred).to(DEV)
    full = ops.confusion(l, p, 19)
    parts = ops.confusion(l[:100000], p[:100000], 19)
    ops.confusion(l[100000:].contiguous(), p[100000:].contiguous(), 19, hist=parts)
    assert torch.equal(full, parts) and int(full.sum()) == int((lab != 255).sum())


def test_miou(ops):
    rng = np.random.default_rng(2)
    h = rng.integers(0, 10 ** 7, (19, 19)).astype(np.int64)
    h[:, 7] = 0; h[7, :] = 0  # absent class -> NaN IoU, skipped by nanmean
    iou, m = ops.miou(torch.from_numpy(h).to(DEV))
    want_iou, want_m = ls.ious_miou(h)
    assert np.allclose(iou.cpu().numpy(), want_iou, rtol=1e-6, equal_nan=True)
    assert abs(float(m) - want_m) <= 1e-6


@pytest.mark.parametrize("shape,size", [((2, 64, 96), (8, 12)), ((1, 1024, 2048), (64, 128)), ((1, 37, 53), (9, 14)),
                                        ((1, 50, 70), (13, 17))])
def test_label_nearest(ops, shape, size):
    rng = np.random.default_rng(1)
    lab = rng.integers(0, 256, shape)
    got = ops.label_nearest(torch.from_numpy(lab).to(DEV), size)
    assert np.array_equal(got.cpu().numpy(), ls.nearest_resize(lab, size))


@pytest.mark.gpu
@pytest.mark.parametrize("lab_dt", [torch.int64, torch.uint8])
@pytest.mark.parametrize("shape", [(5, 16, 48), (4, 37, 53), (3, 64, 128)])
def test_batched_label_space_ops(ops, lab_dt, shape):
    """lut_remap_images / confusion_images / miou_images == the per-dataset calls == the numpy oracle."""
    g = torch.Generator().manual_seed(11)
    B, H, W = shape
    n_cats = [19, 7, 150]
    ids = torch.randint(0, 3, (B,), generator=g).tolist()
    luts = np.stack([np.where(np.arange(256) < 250, np.arange(256) % c, 255).astype(np.uint8) for c in n_cats])
    raw = torch.randint(0, 256, (B, H, W), generator=g, dtype=torch.uint8)
    pred = torch.stack([torch.randint(0, n_cats[d], (H, W), generator=g) for d in ids])
    lab = ops.lut_remap_images(raw.to(DEV), luts, ids, out_dtype=lab_dt)
    want_lab = np.stack([ls.lut_gather(raw[b].numpy(), luts[d]) for b, d in enumerate(ids)])
    assert np.array_equal(lab.cpu().numpy().astype(np.int64), want_lab.astype(np.int64))
    hist, views = ops.confusion_images(lab, pred.to(DEV), ids, n_cats)
    iou, miou = ops.miou_images(hist, n_cats)
    for d, c in enumerate(n_cats):
        sel = [b for b in range(B) if ids[b] == d]
        want = np.zeros((c, c), dtype=np.int64)
        for b in sel:
            want += ls.confusion(want_lab[b], pred[b].numpy(), c)
        assert np.array_equal(views[d].cpu().numpy(), want), d
        w_iou, w_miou = ls.ious_miou(want)
        got = iou[d, :c].cpu().numpy()
        assert np.allclose(got, w_iou, rtol=1e-6, atol=0, equal_nan=True)
        assert (np.isnan(w_miou) and np.isnan(float(miou[d]))) or abs(float(miou[d]) - w_miou) <= 1e-6
    # the LUT can also be fused into the histogram pass (raw ids in, evaluate-time remap)
    hist2, _ = ops.confusion_images(raw.to(DEV), pred.to(DEV), ids, n_cats, luts=luts)
    assert torch.equal(hist, hist2)
    ops.check_errors(DEV)


@pytest.mark.gpu
def test_batched_ops_flag_bad_dataset_id(ops):
    raw = torch.zeros(2, 16, 16, dtype=torch.uint8, device=DEV)
    luts = np.zeros((2, 256), dtype=np.uint8)
    out = ops.lut_remap_images(raw, luts, [0, 5], oob=255)
    assert int(out[1].min()) == 255 and int(out[0].max()) == 0
    with pytest.raises(RuntimeError, match="dataset id"):
        ops.check_errors(DEV)


def test_lb_map_gather_against_real_dataset_readers(ops, golden):
    """a1 pinned: the lb_map LUTs of the seven REAL reader classes (CityScapes, Mapi, Sunrgbd, Bdd100k, Idd, ade2016,
    Coco_data) and their __getitem__ label output (tests/golden/make_golden_lb_maps.py), one image at a time and as a
    mixed-dataset batch in one launch."""
    z = golden("lb_maps.npz")
    raw = torch.from_numpy(z["raw"]).to(DEV)
    for i in range(7):
        for out_dt in (torch.uint8, torch.int64):  # int64: what ToTensor makes of it (lib/transform_cv2.py:300)
            out = ops.lut_remap(raw, z[f"lb_map{i}"], out_dtype=out_dt)
            assert np.array_equal(out.cpu().numpy().astype(np.uint8), z[f"label{i}"]), str(z["names"][i])
    ids = [6, 0, 3, 3, 5, 1, 2, 4]
    luts = np.stack([z[f"lb_map{i}"] for i in range(7)])
    out = ops.lut_remap_images(raw.unsqueeze(0).repeat(len(ids), 1, 1), luts, ids)
    for b, i in enumerate(ids):
        assert np.array_equal(out[b].cpu().numpy(), z[f"label{i}"]), b
