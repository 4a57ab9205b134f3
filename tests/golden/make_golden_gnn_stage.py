"""Generate tests/golden/advgnn_gnn_stage.npz: the REAL lib.loss.loss_cross_datasets.CrossDatasetsCELoss_AdvGNN of
/root/reference in its GNN stage (is_adv=True) and its init stage, on configs/ltbgnn_7_datasets_snp.json.

    python tests/golden/make_golden_gnn_stage.py      # build container only (needs /root/reference)

GNN stage as the config drives it: 2 x 7 trainable graphs (hard, soft) blended by max_rate = (iter % ...) / gnn_iters,
prototype head einsum, dataset aux heads from the dataset prototypes, orth loss, adj MSE against target_bi_graph.
Recorded: every input, the four returned values, and the gradients w.r.t. the features, the prototypes and all graphs.
"""
import os

import numpy as np
import torch

from make_golden import REF, OUT, import_reference, make_labels


def main():
    *_, AdvGNN, Configer = import_reference()
    os.chdir(REF)
    g = torch.Generator().manual_seed(20261020)
    configer = Configer(configs="configs/ltbgnn_7_datasets_snp.json")
    configer.update(("iter",), 21000)  # max_rate = 21000 / 60000 = 0.35
    crit = AdvGNN(configer)
    n_cats, total, c_uni = crit.n_cats, crit.total_cats, crit.max_num_unify_class
    ids = [0, 1, 1, 3, 4, 5, 5, 6]  # dataset 2 absent
    B, D, h, w, H, W = len(ids), 12, 6, 8, 21, 29
    ids_t = torch.tensor(ids, dtype=torch.int32)
    feats = (torch.randn(B, D, h, w, generator=g)).requires_grad_(True)
    proto = (torch.randn(total + c_uni, D, generator=g) * 0.6).requires_grad_(True)
    graphs = []
    for c in n_cats:
        soft = torch.softmax(torch.randn(c, c_uni, generator=g) * 3, dim=0)
        hard = torch.zeros(c, c_uni).scatter_(0, soft.argmax(0, keepdim=True), 1.0) * 0.9 + 0.1 * soft
        graphs += [hard.clone().requires_grad_(True), soft.clone().requires_grad_(True)]
    tgt = []
    for c in n_cats:
        t = torch.full((c, c_uni), 255.0)
        m = torch.rand(c, c_uni, generator=g)
        t[m < 0.2] = 0.0
        t[m > 0.93] = 1.0
        tgt.append(t)
    labels = torch.full((B, H, W), 255, dtype=torch.long)
    for b, d in enumerate(ids):
        labels[b] = make_labels(g, (H, W), n_cats[d])
    preds = {"seg": feats, "unify_prototype": proto, "bi_graphs": graphs, "adv_out": None, "target_bi_graph": tgt}
    loss, orth, aux, adj = crit(preds, labels, ids_t, True, False)
    loss.backward()
    out = {"feats": feats.detach().numpy(), "proto": proto.detach().numpy(), "labels": labels.numpy().astype(np.uint8),
           "ids": np.array(ids, dtype=np.int32), "iter": np.int64(21000),
           "loss": loss.detach().numpy(), "orth": orth.detach().numpy(), "aux": aux.detach().numpy(),
           "adj": adj.detach().numpy(), "dfeats": feats.grad.numpy(), "dproto": proto.grad.numpy()}
    for i, gr in enumerate(graphs):
        out[f"graph{i}"] = gr.detach().numpy()
        out[f"dgraph{i}"] = (gr.grad if gr.grad is not None else torch.zeros_like(gr)).numpy()
    for i, t in enumerate(tgt):
        out[f"target{i}"] = t.numpy()

    # init stage (:1090-1113): graph MSE against the pretrained bipartite graphs + prototype MSE
    # only rows [0, total_cats) are read (:1097); values on the fp16 grid so that the fixture stores them as fp16
    adjm = torch.rand(total, total + c_uni, generator=g).half().float().requires_grad_(True)
    pre = [torch.rand(c, c_uni, generator=g) for c in n_cats]
    p0 = torch.randn(total + c_uni, D, generator=g).requires_grad_(True)
    seg0 = torch.randn(total + c_uni, D, generator=g)
    preds0 = {"seg": seg0, "unify_prototype": p0, "bi_graphs": graphs, "adv_out": None, "adj": adjm,
              "pretrain_bipart_graph": pre}
    configer.update(("loss", "with_datasets_aux"), False)  # the init stage with aux heads is dead code (:1047 fails)
    loss0, *_ = AdvGNN(configer)(preds0, labels, ids_t, True, True)
    loss0.backward()
    out.update({"init_adj": adjm.detach().numpy().astype(np.float16), "init_proto": p0.detach().numpy(), "init_seg": seg0.numpy(),
                "init_loss": loss0.detach().numpy(), "init_dproto": p0.grad.numpy()})
    for i, t in enumerate(pre):
        out[f"init_pre{i}"] = t.numpy()
    path = os.path.join(OUT, "advgnn_gnn_stage.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), float(loss), float(orth), float(aux), float(adj), float(loss0))


if __name__ == "__main__":
    main()
