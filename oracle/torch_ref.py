"""The reference's loss / evaluator call sequence restated with the same torch ops
(test infrastructure and CPU baseline, see oracle/__init__.py).

Runs on whatever device its inputs live on; on the CPU box it is the
``cpu_baseline`` that bench.py times, on the GPU box the same functions on CUDA
tensors are "the reference's eager path on the same B200".
"""
import torch
import torch.nn.functional as F

IGNORE = 255


def neg_log_thresh(p):
    """lib/loss/ohem_ce_loss.py:17 — thresh = -log(tensor(p, dtype=float))."""
    return -torch.log(torch.tensor(p, dtype=torch.float))


def ce_none(logits, labels, ignore=IGNORE):
    """lib/loss/ohem_ce_loss.py:19,27 — CrossEntropyLoss(ignore_index, reduction='none'), fp32 like autocast does."""
    return F.cross_entropy(logits.float(), labels.long(), ignore_index=ignore, reduction="none")


def ohem_select_mean(loss_vec, n_min, thresh):
    """lib/loss/ohem_ce_loss.py:30-34 (and :74-90): keep loss > thresh; fewer than n_min -> topk(n_min); mean."""
    hard = loss_vec[loss_vec > thresh]
    if hard.numel() < n_min:
        hard, _ = loss_vec.topk(n_min)
    return torch.mean(hard)


def ohem_ce_loss(logits, labels, thresh_p, ignore=IGNORE):
    """OhemCELoss.forward, lib/loss/ohem_ce_loss.py:21-34."""
    thresh = neg_log_thresh(thresh_p).to(logits.device)
    n_min = labels[labels != ignore].numel() // 16
    loss = ce_none(logits, labels, ignore).view(-1)
    return ohem_select_mean(loss, n_min, thresh)


def mds_ohem_ce_loss(logits_list, labels, dataset_ids, n_datasets, thresh_p, ignore=IGNORE):
    """MdsOhemCELoss.forward, lib/loss/ohem_ce_loss.py:48-90: per-dataset CE vectors concatenated in dataset
    order, ONE selection, n_min from all labels of the batch."""
    thresh = neg_log_thresh(thresh_p).to(labels.device)
    n_min = labels[labels != ignore].numel() // 16
    losses, cur = [], 0
    for i in range(n_datasets):
        if not (dataset_ids == i).any():
            continue
        losses.append(ce_none(logits_list[cur], labels[dataset_ids == i], ignore).view(-1))
        cur += 1
    return ohem_select_mean(torch.cat(losses, dim=0), n_min, thresh)


def project(logits, graph):
    """lib/loss/loss_cross_datasets.py:1006 — einsum('bchw, nc -> bnhw')."""
    return torch.einsum("bchw, nc -> bnhw", logits, graph)


def upsample(x, size):
    """lib/loss/loss_cross_datasets.py:1007 — bilinear, align_corners=True."""
    return F.interpolate(x, size=size, mode="bilinear", align_corners=True)


def multi_dataset_seg_loss(logits_uni, labels, dataset_ids, bi_graphs, thresh_p=0.4, ignore=IGNORE):
    """The SEG-stage core of CrossDatasetsCELoss_AdvGNN.forward (loss_cross_datasets.py:988-1009,1074):
    per present dataset project + upsample, then MdsOhemCELoss(0.4)."""
    n_datasets = len(bi_graphs)
    size = (labels.size(1), labels.size(2))
    remap_logits = []
    for i in range(n_datasets):
        if not (dataset_ids == i).any():
            continue
        remap_logits.append(upsample(project(logits_uni[dataset_ids == i], bi_graphs[i]), size))
    return mds_ohem_ce_loss(remap_logits, labels, dataset_ids, n_datasets, thresh_p, ignore)


def aux_heads_loss(aux_logits, labels, dataset_ids, thresh_p=0.7, ignore=IGNORE):
    """loss_cross_datasets.py:1044-1056 (is_adv=False branch): Σ_i OhemCELoss(0.7)(upsample(aux[i][ids==i]), target[ids==i])."""
    size = (labels.size(1), labels.size(2))
    total = None
    for i in range(len(aux_logits)):
        if not (dataset_ids == i).any():
            continue
        li = ohem_ce_loss(upsample(aux_logits[i][dataset_ids == i], size), labels[dataset_ids == i], thresh_p, ignore)
        total = li if total is None else total + li
    return total


def seg_stage_total_loss(logits_uni, aux_logits, labels, dataset_ids, bi_graphs, aux_weight):
    """loss = MdsOhemCE(0.4) + aux_weight * Σ aux (loss_cross_datasets.py:1074-1080,1129-1130)."""
    main = multi_dataset_seg_loss(logits_uni, labels, dataset_ids, bi_graphs)
    aux = aux_heads_loss(aux_logits, labels, dataset_ids)
    return main + aux_weight * aux, main, aux


def gnn_stage_total_loss(feats, unify_prototype, bi_graphs, labels, dataset_ids, n_cats, max_rate,
                         target_bi_graph=None, temperature=0.07, orth_weight=1.0, aux_weight=0.2, adj_loss_weight=1.0):
    """CrossDatasetsCELoss_AdvGNN.forward in the GNN stage the ltbgnn configs drive (is_adv=True, dataset aux heads,
    2*n graphs = (hard, soft) pairs, with_orth, mse_or_adv "None"; loss_cross_datasets.py:941-1136):
      aux_i   = OhemCE(0.7)(upsample(einsum(feats[ids==i], proto[cur:cur+n_i])), target[ids==i])        (:941-951,:1047)
      logits  = einsum(feats, proto[total:])                                                              (:961)
      orth    = orth_weight * (-mean(softmax(P P^T / T) * log_softmax(P P^T / T))), P = proto[total:]     (:977-980)
      adj     = sum_i (1/C_uni) * sum-squared-error(soft_i, target_i) where target_i != 255                (:1030-1043)
      ce      = max_rate * MdsOhemCE(0.4)(hard graphs) + (1 - max_rate) * MdsOhemCE(0.4)(soft graphs)     (:1063-1071)
      loss    = ce + aux_weight * sum aux_i + orth + adj_loss_weight * adj                                 (:1129-1136)
    Returns (loss, orth, aux, adj)."""
    n = len(n_cats)
    total = sum(n_cats)
    size = (labels.size(1), labels.size(2))
    aux, adj, cur = None, None, 0
    for i in range(n):
        sel = dataset_ids == i
        if sel.any():
            head = upsample(project(feats[sel], unify_prototype[cur:cur + n_cats[i]]), size)
            li = ohem_ce_loss(head, labels[sel], 0.7)
            aux = li if aux is None else aux + li
            if target_bi_graph is not None:
                keep = target_bi_graph[i] != 255
                soft = bi_graphs[2 * i + 1]
                li = (1 / soft.shape[1]) * F.mse_loss(soft[keep], target_bi_graph[i][keep], reduction="sum")
                adj = li if adj is None else adj + li
        cur += n_cats[i]
    proto = unify_prototype[total:]
    logits = project(feats, proto)
    z = torch.mm(proto, proto.t()) / temperature
    orth = orth_weight * (-1 * torch.mean(F.softmax(z, dim=1) * F.log_softmax(z, dim=1)))
    ce = (max_rate * multi_dataset_seg_loss(logits, labels, dataset_ids, bi_graphs[0::2])
          + (1 - max_rate) * multi_dataset_seg_loss(logits, labels, dataset_ids, bi_graphs[1::2]))
    loss = ce + aux_weight * aux + orth
    if adj is not None:
        loss = loss + adj_loss_weight * adj
    return loss, orth, aux, adj


def remap_matrix_ce_loss(logits, labels, dataset_ids, remap_matrices, ignore=IGNORE):
    """CrossDatasetsCELoss.forward (loss_cross_datasets.py:323-347): per dataset plain mean CE of the
    remap-matrix projection, summed — the path the reference's golden value 5.106813430786133 pins."""
    loss = None
    for i, m in enumerate(remap_matrices):
        if not (dataset_ids == i).any():
            continue
        li = F.cross_entropy(project(logits[dataset_ids == i], m), labels[dataset_ids == i], ignore_index=ignore)
        loss = li if loss is None else loss + li
    return loss


# ---- evaluator (evaluate.py:46-99,101-192) ------------------------------------------------
def eval_probs(logits_passes, size, flips=None):
    """probs = Σ softmax(upsample(logits)) over passes (evaluate.py:149-171).  `flips[i]` marks passes whose
    logits come from the mirrored image and are flipped back before the interpolation (:165-170)."""
    probs = None
    for i, lg in enumerate(logits_passes):
        if flips is not None and flips[i]:
            lg = torch.flip(lg, dims=(3,))
        if tuple(lg.shape[-2:]) != tuple(size):
            lg = upsample(lg, size)
        p = torch.softmax(lg.float(), dim=1)
        probs = p if probs is None else probs + p
    return probs


def eval_preds(probs):
    """evaluate.py:172."""
    return torch.argmax(probs, dim=1)


def nearest_label(label, size):
    """evaluate.py:156-157."""
    return F.interpolate(label.float().unsqueeze(1), size=size, mode="nearest").squeeze(1).long()


def adj_nll_plus(x, adj, lb, ignore=IGNORE):
    """AdjNLLPlusLoss.forward with reduction='none', lib/loss/loss_helper.py:654-668: softmax over the unified
    classes, projection of the PROBABILITIES, bilinear up-sampling, -log at the label class, valid pixels only."""
    pred = torch.softmax(x.float(), dim=1)
    probs = torch.einsum("bchw, nc -> bnhw", pred, adj.float())
    probs = F.interpolate(probs, size=(lb.size(1), lb.size(2)), mode="bilinear", align_corners=True)
    probs = -torch.log(probs)
    keep = lb != ignore
    lb = lb.clone()
    lb[lb == ignore] = 0
    loss = torch.gather(probs, 1, lb.long().unsqueeze(1)).squeeze(1)
    return loss[keep]


def mds_ohem_nll_plus_loss(logits, labels, bi_graphs, dataset_ids, n_datasets, thresh_p, ignore=IGNORE):
    """MdsOhemNLLPlusLoss.forward, lib/loss/ohem_ce_loss.py:104-146."""
    thresh = neg_log_thresh(thresh_p).to(labels.device)
    n_min = labels[labels != ignore].numel() // 16
    losses = []
    for i in range(n_datasets):
        if not (dataset_ids == i).any():
            continue
        losses.append(adj_nll_plus(logits[dataset_ids == i], bi_graphs[i], labels[dataset_ids == i], ignore).view(-1))
    return ohem_select_mean(torch.cat(losses, dim=0), n_min, thresh)


def cross_datasets_ce_mean(logits_uni, labels, dataset_ids, graphs, upsample_to_labels=True, ignore=IGNORE):
    """Per-dataset plain CE of the projected (and up-sampled) logits, summed over the datasets present:
    CrossDatasetsCELoss.forward (lib/loss/loss_cross_datasets.py:329-346, no up-sampling) and
    CrossDatasetsCELoss_GNN.forward (:749-768, with it)."""
    loss = None
    for i in range(len(graphs)):
        if not (dataset_ids == i).any():
            continue
        r = project(logits_uni[dataset_ids == i], graphs[i])
        if upsample_to_labels:
            r = upsample(r, labels.shape[1:])
        ce = F.cross_entropy(r.float(), labels[dataset_ids == i].long(), ignore_index=ignore)
        loss = ce if loss is None else loss + ce
    return loss
