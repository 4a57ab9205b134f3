"""numpy restatement of the integer half of the path (test infrastructure, see oracle/__init__.py).

Every function cites the reference lines it follows (paths relative to the reference root).
"""
import numpy as np

IGNORE = 255


def build_lb_map(labels_info, mode="eval"):
    """lib/cityscapes_cv2.py:156-164: identity uint8[256], then lb_map[id] = trainId
    (train mode maps trainId 255 to class 19)."""
    lb_map = np.arange(256).astype(np.uint8)
    for el in labels_info:
        if mode == "train" and el["trainId"] == 255:
            lb_map[el["id"]] = 19
        else:
            lb_map[el["id"]] = el["trainId"]
    return lb_map


def lut_gather(label, lb_map):
    """lib/base_dataset.py:81-82: label = self.lb_map[label]."""
    return np.asarray(lb_map)[np.asarray(label)]


def parse_class_remap(cfg, n_datasets):
    """lib/class_remap.py:160-172: read class_remap{i} dicts in key order 0,1,2,... until a key is missing."""
    remap_list, max_map_nums = [], []
    for i in range(1, n_datasets + 1):
        raw = cfg["class_remap" + str(i)]
        class_id, mx, d = 0, 0, {}
        while str(class_id) in raw:
            d[class_id] = list(raw[str(class_id)])
            mx = max(mx, len(d[class_id]))
            class_id += 1
        remap_list.append(d)
        max_map_nums.append(mx)
    return remap_list, max_map_nums


def remap_matrix(remap, n_cats, num_unify_classes):
    """lib/class_remap.py:176-183: M[k, v] = 1."""
    m = np.zeros((n_cats, num_unify_classes), dtype=np.float32)
    for k, v in remap.items():
        m[k, v] = 1
    return m


def single_seg_remapping(labels, remap, ignore_index=IGNORE):
    """lib/class_remap.py:34-48: only classes with exactly one target are mapped, everything else -> ignore."""
    labels = np.asarray(labels)
    mask = np.ones_like(labels) * ignore_index
    for k, v in remap.items():
        if len(v) > 1:
            continue
        mask[labels == int(k)] = v[0]
    return mask


def seg_remapping(labels, remap, max_map_num, ignore_index=IGNORE):
    """lib/class_remap.py:50-66: list of max_map_num maps; the i-th holds v[i] where len(v) > i."""
    labels = np.asarray(labels)
    out = []
    for i in range(max_map_num):
        mask = np.ones_like(labels) * ignore_index
        for k, v in remap.items():
            if len(v) <= i:
                continue
            mask[labels == int(k)] = v[i]
        out.append(mask)
    return out


def reverse_seg_remap(preds, remap, dataset_id):
    """lib/class_remap.py:189-203: out = 0; for k in dict order (stop at 19 / 12 for datasets 0 / 1):
    out[preds == lb] = k for lb in v — later keys overwrite."""
    preds = np.asarray(preds)
    out = np.zeros_like(preds)
    for k, v in remap.items():
        if dataset_id == 0 and k == 19:
            break
        if dataset_id == 1 and k == 12:
            break
        for lb in v:
            out[preds == int(lb)] = int(k)
    return out


def is_single_remap_lb(remap_list, lb):
    """lib/class_remap.py:26-32."""
    for remap in remap_list:
        for _, v in remap.items():
            if len(v) == 1 and v[0] == lb:
                return True
    return False


# --- the same remaps expressed as the 256-entry LUTs the kernels consume -----------------
def single_seg_lut(remap, ignore_index=IGNORE):
    lut = np.full(256, ignore_index, dtype=np.uint8)
    for k, v in remap.items():
        if len(v) == 1 and 0 <= int(k) < 256:
            lut[int(k)] = v[0]
    return lut


def seg_luts(remap, max_map_num, ignore_index=IGNORE):
    luts = []
    for i in range(max_map_num):
        lut = np.full(256, ignore_index, dtype=np.uint8)
        for k, v in remap.items():
            if len(v) > i and 0 <= int(k) < 256:
                lut[int(k)] = v[i]
        luts.append(lut)
    return luts


def reverse_seg_lut(remap, dataset_id):
    lut = np.zeros(256, dtype=np.uint8)
    for k, v in remap.items():
        if dataset_id == 0 and k == 19:
            break
        if dataset_id == 1 and k == 12:
            break
        for lb in v:
            if 0 <= int(lb) < 256:
                lut[int(lb)] = int(k)
    return lut


# --- evaluator tail -------------------------------------------------------------------------
def multihot_seg_remapping(labels, remap, num_unify_classes, single_only=False):
    """lib/class_remap.py:260-276 (SegRemapping of ClassRemapOneHotLabel) and :239-258
    (SingleSegRemappingOneHot, single_only=True): bool [b, h, w, C_uni], out[labels == k, v] = 1."""
    labels = np.asarray(labels)
    out = np.zeros(labels.shape + (num_unify_classes,), dtype=bool)
    for k, v in remap.items():
        if single_only and len(v) > 1:
            continue
        sel = labels == int(k)
        for u in v:
            out[sel, u] = True
    return out


def multihot_table(remap, num_unify_classes, single_only=False):
    """The same remap as a uint8 [256, C_uni] table: out[p, :] = table[labels[p], :]."""
    t = np.zeros((256, num_unify_classes), dtype=np.uint8)
    for k, v in remap.items():
        if 0 <= int(k) < 256 and not (single_only and len(v) > 1):
            t[int(k), v] = 1
    return t


def confusion(label, pred, n_a, n_b=None, ignore_label=IGNORE):
    """evaluate.py:89-93 / :174-181 (square) and :631-634, :1738-1741 (rectangular):
    keep = label != ignore; bincount(label[keep]*n_b + pred[keep], minlength=n_a*n_b).reshape(n_a, n_b).
    Exact int64 (the reference accumulates into a float32 tensor)."""
    n_b = n_b or n_a
    label = np.asarray(label).reshape(-1).astype(np.int64)
    pred = np.asarray(pred).reshape(-1).astype(np.int64)
    keep = label != ignore_label
    idx = label[keep] * n_b + pred[keep]
    h = np.bincount(idx, minlength=n_a * n_b)
    if h.size != n_a * n_b:
        raise ValueError("label out of range (the reference's .view(n, n) would raise)")
    return h.reshape(n_a, n_b).astype(np.int64)


def ious_miou(hist):
    """evaluate.py:94-98: ious = diag / (sum0 + sum1 - diag) in float32; miou = nanmean."""
    h = np.asarray(hist).astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        ious = (np.diag(h) / (h.sum(axis=0) + h.sum(axis=1) - np.diag(h))).astype(np.float32)
    return ious, float(np.nanmean(ious)) if np.any(~np.isnan(ious)) else float("nan")


def nearest_resize(label, size):
    """evaluate.py:156-157: F.interpolate(label.float(), size, mode='nearest').long();
    ATen legacy nearest: src = min(floor(dst * (in/out)), in-1) with the scale in fp32."""
    label = np.asarray(label)
    Hin, Win = label.shape[-2:]
    Ho, Wo = size
    sy, sx = np.float32(Hin) / np.float32(Ho), np.float32(Win) / np.float32(Wo)
    ys = np.minimum(np.floor(np.arange(Ho, dtype=np.float32) * sy).astype(np.int64), Hin - 1)
    xs = np.minimum(np.floor(np.arange(Wo, dtype=np.float32) * sx).astype(np.int64), Win - 1)
    return label[..., ys[:, None], xs[None, :]]


def batch_layout(labels_per_dataset):
    """tools/train_ltbgnn_all_datasets_snp.py:708-750: lb = cat(lbs).squeeze(1), dataset_lbs = cat(j*ones(B_j))."""
    lb = np.concatenate(labels_per_dataset, axis=0)
    ids = np.concatenate([np.full(len(l), j, dtype=np.int32) for j, l in enumerate(labels_per_dataset)])
    return lb, ids


# ---- label branch of the training transforms (lib/transform_cv2.py) ---------------------------------------------
def cv2_nearest_resize(label, size):
    """cv2.resize(lb, (im_w, im_h), interpolation=cv2.INTER_NEAREST) (lib/transform_cv2.py:43).  Third-party
    arithmetic (opencv-python, 4.13.0 in the build container; the reference pins none): restated from OpenCV 4.x
    imgproc/resize.cpp — resize() sets inv_scale = (double)dsize / ssize and scale = 1. / inv_scale, resizeNN() takes
    src = min(cvFloor(dst * scale), ssize - 1) on both axes.  Pinned bit-exact against the real cv2 through
    tests/golden/label_pipeline.npz (tests/golden/make_golden_label_pipeline.py)."""
    label = np.asarray(label)
    Hs, Ws = label.shape[:2]
    im_h, im_w = size
    ify = 1.0 / (float(im_h) / float(Hs))
    ifx = 1.0 / (float(im_w) / float(Ws))
    ys = np.minimum(np.floor(np.arange(im_h, dtype=np.float64) * ify).astype(np.int64), Hs - 1)
    xs = np.minimum(np.floor(np.arange(im_w, dtype=np.float64) * ifx).astype(np.int64), Ws - 1)
    return label[ys[:, None], xs[None, :]]


def plan_random_resized_crop(shape, scales, size, rng=np.random):
    """The integers RandomResizedCrop.__call__ derives from its random draws (lib/transform_cv2.py:22-62), drawing
    from `rng` in the same order: uniform(min, max) for the scale, then random(2) for the crop origin."""
    H, W = shape
    crop_h, crop_w = size
    scale = rng.uniform(min(scales), max(scales))
    if np.min([H, W]) < 1080:
        scale = scale * (1080 / np.min([H, W]))
    im_h, im_w = [int(np.ceil(el * scale)) for el in (H, W)]
    if (im_h, im_w) == (crop_h, crop_w):
        return dict(im_h=im_h, im_w=im_w, pad_top=0, pad_left=0, crop_y=0, crop_x=0)
    pad_h = (crop_h - im_h) // 2 + 1 if im_h < crop_h else 0
    pad_w = (crop_w - im_w) // 2 + 1 if im_w < crop_w else 0
    sh, sw = rng.random(2)
    sh, sw = int(sh * (im_h + 2 * pad_h - crop_h)), int(sw * (im_w + 2 * pad_w - crop_w))
    return dict(im_h=im_h, im_w=im_w, pad_top=pad_h, pad_left=pad_w, crop_y=sh, crop_x=sw)


def label_transform_chain(label, lb_map, plan, size):
    """lb_map gather (lib/base_dataset.py:81-82) -> RandomResizedCrop's label branch (lib/transform_cv2.py:43-61) ->
    RandomHorizontalFlip (:71-77) -> int64 (:300), step by step with numpy as the reference does."""
    lb = np.asarray(label)
    if lb_map is not None:
        lb = np.asarray(lb_map)[lb]
    lb = cv2_nearest_resize(lb, (plan["im_h"], plan["im_w"]))
    ph, pw = plan.get("pad_top", 0), plan.get("pad_left", 0)
    if ph > 0 or pw > 0:
        lb = np.pad(lb, ((ph, ph), (pw, pw)), 'constant', constant_values=IGNORE)
    sh, sw = plan.get("crop_y", 0), plan.get("crop_x", 0)
    lb = lb[sh:sh + size[0], sw:sw + size[1]]
    if plan.get("flip", False):
        lb = lb[:, ::-1]
    return lb.astype(np.int64)
