"""Drop-in for lib/class_remap.py ``ClassRemap`` (:8-232): label-space remaps as 256-entry LUT gathers.

The reference applies one compare + masked write per class (19-150 full-tensor passes per call); every
one of those remaps is ``out = lut[labels]`` with a uint8[256] table built once from the same
``class_remap{i}`` config dicts.  Return types match the reference: tensors of the input's dtype / shape
(lists of them for SegRemapping), fp32 remap matrices.
"""
import numpy as np
import torch
import torch.nn as nn

from .. import ops
from . import _reference

CITY_ID = 0
CAM_ID = 1
_REF = "lib.class_remap"


def _opt(configer, *key):
    """configer.get(*key) for keys the hot path does not need: None when absent (Configer.get exits on a missing
    key, tools/configer.py:157-177, so existence is tested first)."""
    try:
        if hasattr(configer, "exists") and not configer.exists(*key):
            return None
        return configer.get(*key)
    except (KeyError, TypeError):
        return None


class ClassRemap:
    """Every method of the reference class (lib/class_remap.py:8-231).  Methods this file does not define are taken
    from the reference's own class and run on this instance (`__getattr__`), which keeps all attributes the
    reference's ``__init__`` sets."""
    _ref_class = "ClassRemap"

    def __init__(self, configer=None):
        self.configer = configer
        self.ignore_index = self.configer.get('loss', 'ignore_index')
        self.num_unify_classes = self.configer.get('num_unify_classes')
        self.remapList = []
        self.maxMapNums = []
        self.class_weight = []
        # attributes of the reference's __init__ (:12-23) that only the passed-through methods read
        self.temperature = _opt(configer, 'contrast', 'temperature')
        self.network_stride = _opt(configer, 'network', 'stride')
        self.num_prototype = _opt(configer, 'contrast', 'num_prototype')
        self.max_iter = _opt(configer, 'lr', 'max_iter')
        self.reweight = _opt(configer, 'loss', 'reweight')
        self.softmax = nn.Softmax(dim=1)
        if self.network_stride:
            self.Upsample = nn.Upsample(scale_factor=self.network_stride, mode='nearest')
        self._unpack()

    def __getattr__(self, name):  # only reached when normal lookup fails
        if name.startswith("_"):
            raise AttributeError(name)
        return _reference.graft_method(self, _REF, type(self)._ref_class, name)

    # ---- config parsing (class_remap.py:146-183) ----------------------------------------------
    def _unpack(self):
        if not self.configer.exists('n_datasets'):
            raise NotImplementedError("read json errror! no  n_datasets")
        self.n_datasets = self.configer.get('n_datasets')
        if self.reweight:  # :153-159
            for i in range(1, self.n_datasets + 1):
                cw = self.configer.get('class_weight' + str(i))
                self.class_weight.append(torch.tensor([cw[str(j)] for j in range(self.num_unify_classes)]))
        for i in range(1, self.n_datasets + 1):
            if not self.configer.exists('class_remap' + str(i)):
                raise NotImplementedError("read json errror! no class_remap" + str(i))
            raw = self.configer.get('class_remap' + str(i))
            remap, class_id, mx = {}, 0, 0
            while str(class_id) in raw:
                remap[class_id] = list(raw[str(class_id)])
                mx = max(mx, len(remap[class_id]))
                class_id += 1
            self.remapList.append(remap)
            self.maxMapNums.append(mx)
        self.class_remap_matrixs = []
        for i in range(self.n_datasets):
            n_cats = self.configer.get('dataset' + str(i + 1), 'n_cats')
            m = torch.zeros([n_cats, self.num_unify_classes], dtype=torch.float32)
            for k, v in self.remapList[i].items():
                m[k, v] = 1
            self.class_remap_matrixs.append(m)
        # the LUT form of every remap
        ign = self.ignore_index
        self._single_luts, self._seg_luts, self._reverse_luts = [], [], []
        for d, remap in enumerate(self.remapList):
            single = np.full(256, ign, dtype=np.uint8)
            for k, v in remap.items():
                if len(v) == 1 and 0 <= int(k) < 256:
                    single[int(k)] = v[0]
            self._single_luts.append(single)
            segs = []
            for j in range(self.maxMapNums[d]):
                lut = np.full(256, ign, dtype=np.uint8)
                for k, v in remap.items():
                    if len(v) > j and 0 <= int(k) < 256:
                        lut[int(k)] = v[j]
                segs.append(lut)
            self._seg_luts.append(segs)
            rev = np.zeros(256, dtype=np.uint8)
            for k, v in remap.items():  # dict order; later keys overwrite (class_remap.py:189-203)
                if (d == CITY_ID and k == 19) or (d == CAM_ID and k == 12):
                    break
                for lb in v:
                    if 0 <= int(lb) < 256:
                        rev[int(lb)] = int(k)
            self._reverse_luts.append(rev)

    # ---- queries ------------------------------------------------------------------------------
    def IsSingleRemaplb(self, lb):
        return any(len(v) == 1 and v[0] == lb for remap in self.remapList for v in remap.values())

    def getAnyClassRemap(self, lb_id, dataset_id):
        return self.remapList[dataset_id][lb_id]

    def getRemapMatrix(self, dataset_id):
        return self.class_remap_matrixs[dataset_id]

    # ---- remaps (device LUT gathers) ----------------------------------------------------------------
    def SingleSegRemapping(self, labels, dataset_id):
        """Only classes with exactly one unified target are mapped; everything else -> ignore_index (:34-48)."""
        return ops.lut_remap(labels, self._single_luts[dataset_id], oob=self.ignore_index)

    def SegRemapping(self, labels, dataset_id):
        """List of maxMapNums[dataset_id] maps; the j-th holds v[j] where len(v) > j, else ignore_index (:50-66)."""
        return [ops.lut_remap(labels, lut, oob=self.ignore_index) for lut in self._seg_luts[dataset_id]]

    def ReverseSegRemap(self, preds, dataset_id):
        """Unified-space predictions -> dataset classes; unmapped ids -> 0 (:189-203)."""
        return ops.lut_remap(preds, self._reverse_luts[dataset_id], oob=0)

    def ExpendRemapByPrototypeNum(self, v):
        """:205-212 — unified ids -> their prototype rows."""
        out = torch.Tensor()
        for i in v:
            out = torch.cat((out, torch.arange(i * self.num_prototype, (i + 1) * self.num_prototype))).long()
        return out

    def get_class_weight(self, cur_class_id, dataset_id):
        """:214-218"""
        return self.class_weight[dataset_id][self.remapList[dataset_id][cur_class_id]]

    def getReweightMatrix(self, lb, dataset_id):
        """:220-228 — ones_like(lb) with the class weight written where a singly-mapped class has a weight != 1.
        The reference writes a float weight into a tensor of lb's (integer) dtype, i.e. truncates it; so does the
        table.  One LUT gather instead of one masked write per class."""
        table, fits = np.ones(256, dtype=np.int64), True
        pairs = []
        for k, v in self.remapList[dataset_id].items():
            if len(v) == 1 and self.class_weight[dataset_id][v[0]] != 1:
                w = self.class_weight[dataset_id][v[0]]
                pairs.append((int(k), w))
                wi = int(torch.ones((), dtype=lb.dtype).fill_(w)) if not lb.dtype.is_floating_point else None
                if wi is None or not (0 <= wi <= 255 and 0 <= int(k) < 256):
                    fits = False
                else:
                    table[int(k)] = wi
        if fits and lb.dtype in (torch.uint8, torch.int32, torch.int64):
            return ops.lut_remap(lb, table.astype(np.uint8), oob=1)
        out = torch.ones_like(lb)  # weights outside a byte, or float labels: the reference's masked writes (CUDA ops)
        for k, w in pairs:
            out[lb == k] = w
        return out

    def GetEqWeightMask(self, labels, dataset_id):
        """:125-143 — float32 [B, H, W, C_uni]: 1 at every unified target of the pixel's class (one table gather)."""
        if not hasattr(self, "_eq_tables"):
            self._eq_tables = []
            for remap in self.remapList:
                t = np.zeros((256, self.num_unify_classes), dtype=np.uint8)
                for k, v in remap.items():
                    if 0 <= int(k) < 256:
                        t[int(k), v] = 1
                self._eq_tables.append(t)
        return ops.multihot_remap(labels, self._eq_tables[dataset_id]).float()


class ClassRemapOneHotLabel(ClassRemap):
    """Multi-hot variants (class_remap.py:232-276): bool [b, h, w, num_unify_classes], one table gather per call.
    ContrastRemapping / KMeansRemapping / MultiProtoRemapping (:278-594, prototype bookkeeping of the legacy contrast
    trainers) are the reference's own methods, run on this instance."""
    _ref_class = "ClassRemapOneHotLabel"

    def __init__(self, configer=None):
        super().__init__(configer)
        self.update_sim_thresh = _opt(configer, 'contrast', 'update_sim_thresh')
        self._multi_tables, self._single_tables = [], []
        for remap in self.remapList:
            multi = np.zeros((256, self.num_unify_classes), dtype=np.uint8)
            single = np.zeros((256, self.num_unify_classes), dtype=np.uint8)
            for k, v in remap.items():
                if 0 <= int(k) < 256:
                    multi[int(k), v] = 1
                    if len(v) == 1:
                        single[int(k), v[0]] = 1
            self._multi_tables.append(multi)
            self._single_tables.append(single)

    def SingleSegRemappingOneHot(self, labels, dataset_id):
        """Only classes with exactly one unified target are set (:239-258)."""
        return ops.multihot_remap(labels, self._single_tables[dataset_id])

    def SegRemapping(self, labels, dataset_id):
        """mask[p, u] = 1 iff u is a target of labels[p] (:260-276)."""
        return ops.multihot_remap(labels, self._multi_tables[dataset_id])


__getattr__ = _reference.module_getattr(_REF, ("ClassRemap", "ClassRemapOneHotLabel", "CITY_ID", "CAM_ID"))
