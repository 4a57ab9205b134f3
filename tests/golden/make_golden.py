"""Generate tests/golden/*.npz by running the REAL reference modules (build container only).

    python tests/golden/make_golden.py            # needs /root/reference

The reference cannot travel to the GPU box, so its outputs on small seeded inputs are
committed as fixtures next to this script.  Modules used, unmodified, from /root/reference:
  lib.loss.ohem_ce_loss.{OhemCELoss, MdsOhemCELoss}
  lib.class_remap.ClassRemap                      (with tools.configer.Configer)
  lib.loss.loss_cross_datasets.{CrossDatasetsCELoss, CrossDatasetsCELoss_AdvGNN}
      (behind a shim for the missing `timm.models.layers.trunc_normal_`, which the loss never calls)
plus the two hot lines loss_cross_datasets.py:1006-1007 replayed verbatim for the multi-dataset case.
"""
import json
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    tm, tmm, tml = types.ModuleType("timm"), types.ModuleType("timm.models"), types.ModuleType("timm.models.layers")
    tml.trunc_normal_ = torch.nn.init.trunc_normal_
    sys.modules.update({"timm": tm, "timm.models": tmm, "timm.models.layers": tml})
    from lib.loss.ohem_ce_loss import OhemCELoss, MdsOhemCELoss
    from lib.class_remap import ClassRemap
    from lib.loss.loss_cross_datasets import CrossDatasetsCELoss, CrossDatasetsCELoss_AdvGNN
    from tools.configer import Configer
    return OhemCELoss, MdsOhemCELoss, ClassRemap, CrossDatasetsCELoss, CrossDatasetsCELoss_AdvGNN, Configer


def make_labels(g, shape, n_cls, p_ignore=0.05):
    lb = torch.randint(0, n_cls, shape, generator=g)
    lb[torch.rand(shape, generator=g) < p_ignore] = 255
    return lb


def onehot_graph(g, c_ds, c_uni):
    idx = torch.randint(0, c_ds, (c_uni,), generator=g)
    idx[:c_ds] = torch.arange(c_ds)  # every dataset class non-empty
    m = torch.zeros(c_ds, c_uni)
    m[idx, torch.arange(c_uni)] = 1
    return m


def main():
    OhemCELoss, MdsOhemCELoss, ClassRemap, CrossDatasetsCELoss, AdvGNN, Configer = import_reference()
    os.chdir(REF)  # Configer / configs use relative paths
    g = torch.Generator().manual_seed(20261018)

    # ---- 1. OhemCELoss: threshold branch, top-k branch, all-ignore -------------------------------
    cases = {}
    for name, scale, conf in (("thresh", 3.0, 0.0), ("topk", 1.0, 9.0), ("allign", 1.0, 0.0)):
        N_, C_, H_, W_ = 2, 19, 24, 40
        logits = torch.randn(N_, C_, H_, W_, generator=g) * scale
        labels = make_labels(g, (N_, H_, W_), C_)
        labels[:, :3] = 255
        if name == "allign":
            labels[:] = 255
        if conf:
            boost = torch.zeros_like(logits)
            lab0 = labels.clone()
            lab0[lab0 == 255] = 0
            boost.scatter_(1, lab0.unsqueeze(1), conf)
            keep = torch.rand(N_, 1, H_, W_, generator=g) < 0.97
            logits = logits + boost * keep
        logits.requires_grad_(True)
        loss = OhemCELoss(0.7)(logits, labels)
        if name != "allign":
            (loss * 3.0).backward()
        cases[f"ohem_{name}_logits"] = logits.detach().numpy()
        cases[f"ohem_{name}_labels"] = labels.numpy()
        cases[f"ohem_{name}_loss"] = loss.detach().numpy()
        cases[f"ohem_{name}_dlogits"] = (logits.grad if logits.grad is not None else torch.zeros_like(logits)).numpy()
    np.savez_compressed(os.path.join(OUT, "ohem_ce.npz"), **cases)

    # ---- 2. multi-dataset: einsum + interpolate (the verbatim hot lines) + MdsOhemCELoss -----------
    class Cfg:  # MdsOhemCELoss only reads configer.get('n_datasets') (ohem_ce_loss.py:41)
        def __init__(self, n):
            self.n = n

        def get(self, *k):
            assert k == ("n_datasets",)
            return self.n

    cases = {}
    n_cats, c_uni, B, h, w, H, W = [5, 3, 7], 11, 6, 7, 9, 25, 33
    for name, ids in (("sorted", [0, 0, 1, 2, 2, 2]), ("shuffled", [2, 0, 2, 1, 0, 2]), ("absent", [0, 0, 2, 2, 0, 2])):
        ids_t = torch.tensor(ids, dtype=torch.int32)
        x = (torch.randn(B, c_uni, h, w, generator=g) * 2.5).requires_grad_(True)
        graphs = [onehot_graph(g, c, c_uni) for c in n_cats]
        labels = torch.full((B, H, W), 255, dtype=torch.long)
        for b, d in enumerate(ids):
            labels[b] = make_labels(g, (H, W), n_cats[d])
        remap_logits = []
        for i in range(len(n_cats)):
            if not (ids_t == i).any():
                continue
            r = torch.einsum('bchw, nc -> bnhw', x[ids_t == i], graphs[i])                       # :1006
            r = F.interpolate(r, size=(labels.size(1), labels.size(2)), mode="bilinear", align_corners=True)  # :1007
            remap_logits.append(r)
        loss = MdsOhemCELoss(Cfg(len(n_cats)), 0.4)(remap_logits, labels, ids_t)
        (loss * 2.0).backward()
        cases.update({f"mds_{name}_x": x.detach().numpy(), f"mds_{name}_labels": labels.numpy(),
                      f"mds_{name}_ids": np.array(ids, dtype=np.int32), f"mds_{name}_loss": loss.detach().numpy(),
                      f"mds_{name}_dx": x.grad.numpy()})
        for i, gr in enumerate(graphs):
            cases[f"mds_{name}_graph{i}"] = gr.numpy()
    # dense graphs with grad (GNN stage)
    ids = [0, 1, 1, 2, 0, 2]
    ids_t = torch.tensor(ids, dtype=torch.int32)
    x = (torch.randn(B, c_uni, h, w, generator=g) * 2.5).requires_grad_(True)
    graphs = [torch.softmax(torch.randn(c, c_uni, generator=g) * 4, dim=0).requires_grad_(True) for c in n_cats]
    labels = torch.full((B, H, W), 255, dtype=torch.long)
    for b, d in enumerate(ids):
        labels[b] = make_labels(g, (H, W), n_cats[d])
    remap_logits = []
    for i in range(len(n_cats)):
        r = torch.einsum('bchw, nc -> bnhw', x[ids_t == i], graphs[i])
        remap_logits.append(F.interpolate(r, size=(H, W), mode="bilinear", align_corners=True))
    loss = MdsOhemCELoss(Cfg(len(n_cats)), 0.4)(remap_logits, labels, ids_t)
    loss.backward()
    cases.update({"mds_dense_x": x.detach().numpy(), "mds_dense_labels": labels.numpy(),
                  "mds_dense_ids": np.array(ids, dtype=np.int32), "mds_dense_loss": loss.detach().numpy(),
                  "mds_dense_dx": x.grad.numpy()})
    for i, gr in enumerate(graphs):
        cases[f"mds_dense_graph{i}"] = gr.detach().numpy()
        cases[f"mds_dense_dgraph{i}"] = gr.grad.numpy()
    np.savez_compressed(os.path.join(OUT, "mds.npz"), **cases)

    # ---- 3. ClassRemap on the reference's own configs ------------------------------------------------
    cases = {}
    for tag, cfgfile, n_ds in (("test", "configs/test/test.json", 2),
                               ("cca", "configs/bisenetv2_city_cam_a2d2.json", 3)):
        configer = Configer(configs=cfgfile)
        cr = ClassRemap(configer)
        raw = json.load(open(cfgfile))
        for d in range(n_ds):
            n_cats_d = raw[f"dataset{d + 1}"]["n_cats"]
            lb = torch.randint(0, n_cats_d + 2, (2, 9, 13), generator=g)  # includes values that are not keys
            lb[0, 0, :4] = 255
            cases[f"{tag}_d{d}_labels"] = lb.numpy()
            cases[f"{tag}_d{d}_single"] = cr.SingleSegRemapping(lb, d).numpy()
            segs = cr.SegRemapping(lb, d)
            cases[f"{tag}_d{d}_seg"] = np.stack([s.numpy() for s in segs]) if segs else np.zeros((0,) + lb.shape)
            cases[f"{tag}_d{d}_matrix"] = cr.getRemapMatrix(d).numpy()
            preds = torch.randint(0, cr.num_unify_classes, (2, 9, 13), generator=g)
            cases[f"{tag}_d{d}_preds"] = preds.numpy()
            cases[f"{tag}_d{d}_reverse"] = cr.ReverseSegRemap(preds, d).numpy()
        cases[f"{tag}_single_lbs"] = np.array([cr.IsSingleRemaplb(u) for u in range(cr.num_unify_classes)])
    np.savez_compressed(os.path.join(OUT, "class_remap.npz"), **cases)

    # ---- 4. the reference's own known-answer test (test_loss_cross_datasets.py:118-145) ------------------
    configer = Configer(configs="configs/test/test.json")
    loss_fuc = CrossDatasetsCELoss(configer)
    lb = torch.tensor([[[2, 1], [0, 1]], [[2, 1], [1, 2]]])
    logits = torch.tensor([[[[1, 2, 3, 4], [0, 1, 2, 3]], [[2, 3, 4, 1], [3, 0, 1, 2]]],
                           [[[3, 1, 2, 0], [2, 4, 1, 0]], [[3, 1, 0, 2], [2, 4, 3, 1]]]], dtype=torch.float)
    logits = logits.permute(0, 3, 1, 2).contiguous()
    val = float(loss_fuc({"seg": logits}, lb, torch.tensor([0, 1])))
    assert val == 5.106813430786133, val
    np.savez_compressed(os.path.join(OUT, "kat_crossdatasets_celoss.npz"), logits=logits.numpy(), labels=lb.numpy(),
                        ids=np.array([0, 1]), loss=np.float64(val),
                        matrix0=loss_fuc.classRemapper.getRemapMatrix(0).numpy(),
                        matrix1=loss_fuc.classRemapper.getRemapMatrix(1).numpy())

    # ---- 5. CrossDatasetsCELoss_AdvGNN, SEG stage, with dataset aux heads (ltbgnn_7_datasets_snp.json) ---
    configer = Configer(configs="configs/ltbgnn_7_datasets_snp.json")
    crit = AdvGNN(configer)
    n_cats = crit.n_cats
    c_uni = crit.max_num_unify_class
    ids = [0, 1, 1, 2, 3, 4, 5, 6, 6]
    B, h, w, H, W = len(ids), 6, 8, 21, 29
    ids_t = torch.tensor(ids, dtype=torch.int32)
    x = (torch.randn(B, c_uni, h, w, generator=g) * 2.0).requires_grad_(True)
    aux = [(torch.randn(B, c, h, w, generator=g) * 2.0).requires_grad_(True) for c in n_cats]
    graphs = [onehot_graph(g, c, c_uni) for c in n_cats]
    labels = torch.full((B, H, W), 255, dtype=torch.long)
    for b, d in enumerate(ids):
        labels[b] = make_labels(g, (H, W), n_cats[d])
    preds = {"seg": x, "aux": aux, "unify_prototype": None, "bi_graphs": graphs, "adv_out": None}
    loss, orth, aux_loss, adj = crit(preds, labels, ids_t, False, False)
    loss.backward()
    cases = {"x": x.detach().numpy(), "labels": labels.numpy(), "ids": np.array(ids, dtype=np.int32),
             "loss": loss.detach().numpy(), "aux_loss": aux_loss.detach().numpy(), "dx": x.grad.numpy(),
             "n_cats": np.array(n_cats), "c_uni": np.array(c_uni), "aux_weight": np.array(crit.aux_weight)}
    for i in range(len(n_cats)):
        cases[f"graph{i}"] = graphs[i].numpy()
        cases[f"aux{i}"] = aux[i].detach().numpy()
        cases[f"daux{i}"] = (aux[i].grad if aux[i].grad is not None else torch.zeros_like(aux[i])).numpy()
    np.savez_compressed(os.path.join(OUT, "advgnn_seg_stage.npz"), **cases)
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
