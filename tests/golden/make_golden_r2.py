"""Round-2 fixtures from the REAL reference classes (build container only; needs /root/reference):

    python tests/golden/make_golden_r2.py      ->  tests/golden/r2_losses.npz

  lib.loss.ohem_ce_loss.MdsOhemNLLPlusLoss     (:92-146, with lib.loss.loss_helper.AdjNLLPlusLoss :647-668)
  lib.loss.loss_cross_datasets.CrossDatasetsCELoss_GNN   (:714-776)
  lib.loss.loss_cross_datasets.CrossDatasetsCELoss_CLIP  (:662-712, both with_unify_label settings)
  lib.loss.loss_cross_datasets.CrossDatasetsCELoss       (:303-347) on a 3-dataset batch with an absent dataset
  lib.class_remap.ClassRemap.{getReweightMatrix, GetEqWeightMask, get_class_weight}
All unmodified, behind the same `timm.models.layers.trunc_normal_` shim as make_golden.py.
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def shim():
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    tm, tmm, tml = types.ModuleType("timm"), types.ModuleType("timm.models"), types.ModuleType("timm.models.layers")
    tml.trunc_normal_ = torch.nn.init.trunc_normal_
    sys.modules.update({"timm": tm, "timm.models": tmm, "timm.models.layers": tml})


def make_labels(g, shape, n_cls, p_ignore=0.05):
    lb = torch.randint(0, n_cls, shape, generator=g)
    lb[torch.rand(shape, generator=g) < p_ignore] = 255
    return lb


def onehot_graph(g, c_ds, c_uni):
    idx = torch.randint(0, c_ds, (c_uni,), generator=g)
    idx[:c_ds] = torch.arange(c_ds)
    m = torch.zeros(c_ds, c_uni)
    m[idx, torch.arange(c_uni)] = 1
    return m


# the label-space keys of the three-dataset config used below (tests rebuild the same dict)
N_CATS = [5, 3, 7]
C_UNI = 11
REMAP = [
    {"0": [0], "1": [1, 2], "2": [3], "3": [4, 5, 6], "4": [7]},
    {"0": [8], "1": [0, 9], "2": [10]},
    {"0": [0], "1": [1], "2": [2], "3": [3, 8], "4": [4], "5": [9], "6": [10, 5]},
]
CLASS_WEIGHT = [{str(j): (2.0 if j % 3 == 0 else (0.5 if j % 3 == 1 else 1.0)) for j in range(C_UNI)} for _ in range(3)]


def config_dict(reweight=False, with_unify_label=True, with_spa=False, with_max_enc=False):
    c = {"n_datasets": 3, "num_unify_classes": C_UNI, "class_remaper": "ClassRemap",
         "contrast": {"num_prototype": 1, "temperature": 0.07, "with_mulbn": False, "update_sim_thresh": 0.5},
         "network": {"stride": 4}, "lr": {"max_iter": 100},
         "loss": {"ignore_index": 255, "reweight": reweight, "with_aux": False, "with_unify_label": with_unify_label,
                  "with_spa": with_spa, "spa_loss_weight": 0.01, "with_max_enc": with_max_enc, "max_enc_weight": 0.1}}
    for i in range(3):
        c[f"dataset{i + 1}"] = {"n_cats": N_CATS[i]}
        c[f"class_remap{i + 1}"] = REMAP[i]
        c[f"class_weight{i + 1}"] = CLASS_WEIGHT[i]
    return c


def main():
    shim()
    from lib.loss.ohem_ce_loss import MdsOhemNLLPlusLoss
    from lib.loss.loss_cross_datasets import CrossDatasetsCELoss, CrossDatasetsCELoss_CLIP, CrossDatasetsCELoss_GNN
    from lib.class_remap import ClassRemap
    from tools.configer import Configer
    os.chdir(REF)
    g = torch.Generator().manual_seed(20261019)
    cases = {}

    # ---- 1. MdsOhemNLLPlusLoss -----------------------------------------------------------------------------
    ids = [0, 2, 1, 2, 0]
    B, h, w, H, W = len(ids), 7, 9, 25, 33
    ids_t = torch.tensor(ids)
    for name, thresh, dense, scale in (("thresh", 0.4, False, 2.0), ("topk", 0.02, False, 1.0),
                                       ("dense", 0.4, True, 2.0), ("absent", 0.4, False, 2.0)):
        this_ids = ids_t if name != "absent" else torch.tensor([0, 2, 2, 2, 0])
        x = (torch.randn(B, C_UNI, h, w, generator=g) * scale).requires_grad_(True)
        if dense:  # soft graphs whose columns sum to 1 (projected probabilities stay a distribution)
            graphs = [torch.softmax(torch.randn(c, C_UNI, generator=g) * 2.0, dim=0).requires_grad_(True) for c in N_CATS]
        else:
            graphs = [onehot_graph(g, c, C_UNI) for c in N_CATS]
        labels = torch.full((B, H, W), 255, dtype=torch.long)
        for b, d in enumerate(this_ids.tolist()):
            labels[b] = make_labels(g, (H, W), N_CATS[d])
        crit = MdsOhemNLLPlusLoss(Configer(config_dict=config_dict()), thresh)
        loss = crit(x, labels.clone(), graphs, this_ids)
        (loss * 1.5).backward()
        cases[f"nll_{name}_x"] = x.detach().numpy()
        cases[f"nll_{name}_labels"] = labels.numpy()
        cases[f"nll_{name}_ids"] = this_ids.numpy()
        cases[f"nll_{name}_thresh"] = np.float64(thresh)
        cases[f"nll_{name}_loss"] = loss.detach().numpy()
        cases[f"nll_{name}_dx"] = x.grad.numpy()
        for i in range(3):
            cases[f"nll_{name}_graph{i}"] = graphs[i].detach().numpy()
            if dense:
                cases[f"nll_{name}_dgraph{i}"] = graphs[i].grad.numpy()

    # ---- 2. CrossDatasetsCELoss_GNN (prototype head + projection + upsample + plain CE, regularisers on) -----
    D = 6
    feats = (torch.randn(B, D, h, w, generator=g)).requires_grad_(True)
    proto = (torch.randn(C_UNI, D, generator=g)).requires_grad_(True)
    graphs = [torch.softmax(torch.randn(c, C_UNI, generator=g) * 2.0, dim=0).requires_grad_(True) for c in N_CATS]
    labels = torch.full((B, H, W), 255, dtype=torch.long)
    for b, d in enumerate(ids):
        labels[b] = make_labels(g, (H, W), N_CATS[d])
    crit = CrossDatasetsCELoss_GNN(Configer(config_dict=config_dict(with_spa=True, with_max_enc=False)))
    loss = crit({"seg": feats, "unify_prototype": proto, "bi_graphs": graphs}, labels, ids_t)
    loss.backward()
    cases.update({"gnn_feats": feats.detach().numpy(), "gnn_proto": proto.detach().numpy(), "gnn_labels": labels.numpy(),
                  "gnn_ids": ids_t.numpy(), "gnn_loss": loss.detach().numpy(), "gnn_dfeats": feats.grad.numpy(),
                  "gnn_dproto": proto.grad.numpy()})
    for i in range(3):
        cases[f"gnn_graph{i}"] = graphs[i].detach().numpy()
        cases[f"gnn_dgraph{i}"] = graphs[i].grad.numpy()

    # ---- 3. CrossDatasetsCELoss_CLIP, both settings ----------------------------------------------------------
    for name, wul in (("unify", True), ("perds", False)):
        feats = torch.randn(B, D, h, w, generator=g).requires_grad_(True)
        text = [torch.randn(c, D, generator=g) for c in N_CATS] + [torch.randn(C_UNI, D, generator=g)]
        crit = CrossDatasetsCELoss_CLIP(Configer(config_dict=config_dict(with_unify_label=wul)))
        loss = crit({"seg": feats, "prototypes": text}, labels, ids_t)
        loss.backward()
        cases[f"clip_{name}_feats"] = feats.detach().numpy()
        cases[f"clip_{name}_loss"] = loss.detach().numpy()
        cases[f"clip_{name}_dfeats"] = feats.grad.numpy()
        for i, t in enumerate(text):
            cases[f"clip_{name}_text{i}"] = t.numpy()
    cases["clip_labels"] = labels.numpy()
    cases["clip_ids"] = ids_t.numpy()

    # ---- 4. CrossDatasetsCELoss with an absent dataset (full-resolution logits, no upsample) -------------------
    ids2 = torch.tensor([2, 0, 2, 0])
    x = (torch.randn(4, C_UNI, 12, 16, generator=g) * 2.0).requires_grad_(True)
    lb = torch.full((4, 12, 16), 255, dtype=torch.long)
    for b, d in enumerate(ids2.tolist()):
        lb[b] = make_labels(g, (12, 16), N_CATS[d])
    crit = CrossDatasetsCELoss(Configer(config_dict=config_dict()))
    loss = crit({"seg": x}, lb, ids2)
    loss.backward()
    cases.update({"ce_x": x.detach().numpy(), "ce_labels": lb.numpy(), "ce_ids": ids2.numpy(),
                  "ce_loss": loss.detach().numpy(), "ce_dx": x.grad.numpy()})

    # ---- 5. ClassRemap: reweight matrix, equal-weight mask, class weights ---------------------------------------
    cr = ClassRemap(Configer(config_dict=config_dict(reweight=True)))
    for d in range(3):
        lbd = torch.randint(0, N_CATS[d] + 2, (2, 9, 13), generator=g)
        lbd[0, 0, :4] = 255
        cases[f"cr_d{d}_labels"] = lbd.numpy()
        cases[f"cr_d{d}_reweight"] = cr.getReweightMatrix(lbd, d).numpy()
        cases[f"cr_d{d}_eqmask"] = cr.GetEqWeightMask(lbd, d).numpy()
        cases[f"cr_d{d}_cw"] = np.concatenate([cr.get_class_weight(k, d).numpy() for k in range(N_CATS[d])])
    np.savez_compressed(os.path.join(OUT, "r2_losses.npz"), **cases)
    print("written", os.path.join(OUT, "r2_losses.npz"), {k: float(v) for k, v in cases.items() if k.endswith("_loss")})


if __name__ == "__main__":
    main()
