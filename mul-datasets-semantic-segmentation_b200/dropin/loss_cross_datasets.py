"""Drop-in for the SEG-stage path of lib/loss/loss_cross_datasets.py ``CrossDatasetsCELoss_AdvGNN`` (:826-1135).

Covered (the hot path of SURVEY.md §8, rows a5-a10): ``forward(preds, target, dataset_ids, is_adv=False,
init_gnn_stage=False)`` with ``preds = {'seg', 'bi_graphs', 'aux'?, 'unify_prototype': None}`` —
    loss = MdsOhemCELoss(0.4)(upsample(einsum(seg[ids==i], bi_graphs[i])) for i ...)          (:1006-1007, :1074)
         + aux_weight * sum_i OhemCELoss(0.7)(upsample(aux[i][ids==i]), target[ids==i])        (:1044-1056, :1129-1130)
returned as ``(loss, orth_loss, aux_loss, adj_loss)`` with the two unused terms None, exactly like the
reference in that stage.  The GNN / adversarial stage (is_adv=True, prototype einsum, orth / adj / adv terms)
is outside the accelerated path and raises NotImplementedError: keep the reference class for it.
Unlike the reference this forward never synchronises with the host (no `.any()` on dataset_ids): a dataset
without images in the batch contributes NaN-free zeros to the aux sum through its empty OHEM segment.
"""
import torch
import torch.nn as nn

from .. import ops
from .ohem_ce_loss import MdsOhemCELoss, OhemCELoss


class CrossDatasetsCELoss_AdvGNN(nn.Module):
    def __init__(self, configer=None):
        super().__init__()
        self.configer = configer
        self.n_datasets = self.configer.get('n_datasets')
        self.with_datasets_aux = self.configer.get('loss', 'with_datasets_aux')
        self.n_cats = [self.configer.get('dataset' + str(i), 'n_cats') for i in range(1, self.n_datasets + 1)]
        self.total_cats = sum(self.n_cats)
        self.max_num_unify_class = int(self.configer.get('GNN', 'unify_ratio') * self.total_cats)
        self.OhemCELoss = OhemCELoss(0.7, ignore_lb=255)
        self.mdsOhemCELoss = MdsOhemCELoss(self.configer, 0.4, ignore_lb=255)
        if self.with_datasets_aux:
            self.aux_weight = self.configer.get('loss', 'aux_weight')

    def forward(self, preds, target, dataset_ids, is_adv=True, init_gnn_stage=False):
        if is_adv or init_gnn_stage or preds.get('unify_prototype') is not None:
            raise NotImplementedError("mdseg_b200 accelerates the SEG stage (is_adv=False, unify_prototype=None); "
                                      "use the reference CrossDatasetsCELoss_AdvGNN for the GNN stage")
        logits, bi_graphs = preds['seg'], preds['bi_graphs']
        if len(bi_graphs) != self.n_datasets:
            raise NotImplementedError("soft/max graph pairs (2 * n_datasets graphs) belong to the GNN stage")
        loss = self.mdsOhemCELoss.forward_fused(logits, target, dataset_ids, bi_graphs)
        aux_loss = None
        if self.with_datasets_aux:
            per_ds = ops.up_ohem_ce(list(preds['aux']), target, dataset_ids, float(self.OhemCELoss.thresh),
                                    self.OhemCELoss.ignore_lb, seg_per_dataset=True)
            # a dataset without images has an empty OHEM segment (mean of nothing = NaN): the reference skips it
            aux_loss = torch.nan_to_num(per_ds, nan=0.0).sum()
            loss = loss + self.aux_weight * aux_loss
        return loss, None, aux_loss, None
