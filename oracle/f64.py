"""Independent float64 numpy implementation of SURVEY.md §8(a') items 3-9 (test infrastructure).

It shares no code with torch: interpolation geometry follows ATen's fp32 index/weight
arithmetic (so the same pixels use the same corners), everything else is float64.
Used to (i) cross-check torch_ref.py, (ii) give the ground truth that the fp32/bf16
tolerances are measured against and (iii) define the band of pixels whose loss is so
close to the OHEM threshold that fp32 implementations may legitimately disagree.
"""
import numpy as np

IGNORE = 255


def axis_weights(n_in, n_out):
    """§8(a') item 4 / ATen area_pixel_compute_source_index(align_corners=True), fp32:
    r=(n_in-1)/(n_out-1); s=r*o; i0=int(s); i1=i0+(i0<n_in-1); l1=s-i0; l0=1-l1."""
    r = np.float32(n_in - 1) / np.float32(n_out - 1) if n_out > 1 else np.float32(0)
    s = (r * np.arange(n_out, dtype=np.float32)).astype(np.float32)
    i0 = np.minimum(s.astype(np.int64), n_in - 1)
    i1 = i0 + (i0 < n_in - 1)
    l1 = (s - i0.astype(np.float32)).astype(np.float32)
    l0 = (np.float32(1) - l1).astype(np.float32)
    return i0, i1, l0.astype(np.float64), l1.astype(np.float64)


def upsample(x, H, W):
    """out = l0h*(l0w*a + l1w*b) + l1h*(l0w*c + l1w*d) (loss_cross_datasets.py:1007)."""
    x = np.asarray(x, dtype=np.float64)
    y0, y1, l0h, l1h = axis_weights(x.shape[-2], H)
    x0, x1, l0w, l1w = axis_weights(x.shape[-1], W)
    top = x[..., y0, :][..., :, x0] * l0w + x[..., y0, :][..., :, x1] * l1w
    bot = x[..., y1, :][..., :, x0] * l0w + x[..., y1, :][..., :, x1] * l1w
    return top * l0h[:, None] + bot * l1h[:, None]


def upsample_adjoint(g, h, w):
    """Adjoint of `upsample`: g [..., H, W] -> [..., h, w]."""
    g = np.asarray(g, dtype=np.float64)
    H, W = g.shape[-2:]
    y0, y1, l0h, l1h = axis_weights(h, H)
    x0, x1, l0w, l1w = axis_weights(w, W)
    lead = g.shape[:-2]
    tmp = np.zeros(lead + (H, w))
    np.add.at(tmp, (..., slice(None), x0), g * l0w)
    np.add.at(tmp, (..., slice(None), x1), g * l1w)
    out = np.zeros(lead + (h, w))
    np.add.at(out, (..., y0, slice(None)), tmp * l0h[:, None])
    np.add.at(out, (..., y1, slice(None)), tmp * l1h[:, None])
    return out


def ce_per_pixel(z, labels, ignore=IGNORE):
    """§8(a') item 5.  z [B,C,H,W] f64, labels [B,H,W] -> (loss [B,H,W], softmax [B,C,H,W])."""
    z = np.asarray(z, dtype=np.float64)
    m = z.max(axis=1, keepdims=True)
    e = np.exp(z - m)
    s = e.sum(axis=1, keepdims=True)
    lse = (m + np.log(s))[:, 0]
    valid = labels != ignore
    lab = np.where(valid, labels, 0).astype(np.int64)
    zl = np.take_along_axis(z, lab[:, None], axis=1)[:, 0]
    return np.where(valid, lse - zl, 0.0), e / s


def ohem_select(loss_vec, n_valid, thresh):
    """§8(a') item 6: returns (mean, mask, mode, kth).  thresh is the fp32 value of -log(p)."""
    n_min = int(n_valid) // 16
    mask = loss_vec > float(thresh)
    if int(mask.sum()) >= n_min:
        sel = loss_vec[mask]
        return (float(sel.mean()) if sel.size else float("nan")), mask, 0, None
    order = np.argsort(-loss_vec, kind="stable")[:n_min]
    mask = np.zeros(loss_vec.shape, dtype=bool)
    mask[order] = True
    return float(loss_vec[order].mean()), mask, 1, float(loss_vec[order[-1]]) if n_min else None


def ohem_ce(logits, labels, thresh, ignore=IGNORE):
    """OhemCELoss on full-resolution logits: (loss, dlogits, per-pixel loss, selection mask)."""
    loss_px, p = ce_per_pixel(logits, labels, ignore)
    n_valid = int((labels != ignore).sum())
    mean, mask, mode, kth = ohem_select(loss_px.reshape(-1), n_valid, thresh)
    mask = mask.reshape(loss_px.shape)
    n_sel = int(mask.sum())
    valid = labels != ignore
    onehot = np.zeros_like(p)
    np.put_along_axis(onehot, np.where(valid, labels, 0).astype(np.int64)[:, None], 1.0, axis=1)
    w = (mask & valid)[:, None] / max(n_sel, 1)
    return mean, w * (p - onehot), loss_px, mask


def multi_dataset(logits_uni, labels, dataset_ids, graphs, thresh, ignore=IGNORE, want_graph_grads=False):
    """§8(a') items 3,4,5,7,8: project -> upsample -> CE -> one selection -> backward.

    Returns dict(loss, dlogits_uni, loss_px [B,H,W] (nan for images of absent datasets), mask, dgraphs)."""
    x = np.asarray(logits_uni, dtype=np.float64)
    B, Cu, h, w = x.shape
    H, W = labels.shape[1:]
    ids = np.asarray(dataset_ids)
    n_valid = int((labels != ignore).sum())
    loss_px = np.full((B, H, W), np.nan)
    soft = {}
    order = []
    for i, G in enumerate(graphs):
        idx = np.nonzero(ids == i)[0]
        if idx.size == 0:
            continue
        y = np.einsum("bchw,nc->bnhw", x[idx], np.asarray(G, dtype=np.float64))
        z = upsample(y, H, W)
        lp, p = ce_per_pixel(z, labels[idx], ignore)
        loss_px[idx] = lp
        soft[i] = (idx, p)
        order.append(idx)
    present = np.concatenate(order) if order else np.zeros(0, dtype=np.int64)
    vec = loss_px[present].reshape(-1)
    mean, mask_vec, mode, kth = ohem_select(vec, n_valid, thresh)
    mask = np.zeros((B, H, W), dtype=bool)
    mask[present] = mask_vec.reshape(len(present), H, W)
    n_sel = int(mask.sum())
    dx = np.zeros_like(x)
    dgraphs = [np.zeros(np.asarray(G).shape) for G in graphs]
    for i, (idx, p) in soft.items():
        lab = labels[idx]
        valid = lab != ignore
        onehot = np.zeros_like(p)
        np.put_along_axis(onehot, np.where(valid, lab, 0).astype(np.int64)[:, None], 1.0, axis=1)
        dz = (mask[idx] & valid)[:, None] / max(n_sel, 1) * (p - onehot)
        dy = upsample_adjoint(dz, h, w)
        G = np.asarray(graphs[i], dtype=np.float64)
        dx[idx] = np.einsum("bnhw,nc->bchw", dy, G)
        if want_graph_grads:
            dgraphs[i] = np.einsum("bnhw,bchw->nc", dy, x[idx])
    return dict(loss=mean, dlogits_uni=dx, loss_px=loss_px, mask=mask, mode=mode, kth=kth, dgraphs=dgraphs,
                n_sel=n_sel)


def up_ohem_ce(src, labels, thresh, ignore=IGNORE):
    """One aux head: OhemCE(upsample(src), labels) with gradient w.r.t. the low-res src."""
    H, W = labels.shape[1:]
    z = upsample(src, H, W)
    mean, dz, loss_px, mask = ohem_ce(z, labels, thresh, ignore)
    return mean, upsample_adjoint(dz, src.shape[-2], src.shape[-1]), loss_px, mask


def threshold_band(loss_px64, thresh, rel=4e-6):
    """Pixels whose float64 loss lies within `rel` (relative, plus the same absolute) of the threshold:
    an fp32 evaluation order may put them on either side (SURVEY.md §7 'hard parts')."""
    t = float(thresh)
    return np.abs(loss_px64 - t) <= rel * max(1.0, abs(t)) + rel * np.abs(loss_px64)
