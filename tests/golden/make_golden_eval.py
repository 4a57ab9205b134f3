"""Generate tests/golden/evaluator.npz by running the REAL evaluators of /root/reference/evaluate.py on CPU.

    python tests/golden/make_golden_eval.py        # build container only (needs /root/reference)

evaluate.py imports nine packages this image does not have (timm, yacs, munkres, clip, nvidia.dali, cvcuda,
torchnvjpeg, torchvision, ot, matplotlib ...) and hard-codes ``.cuda()``.  None of that is on the evaluator
tail, so the generator
  * serves those top-level names from an import hook that fabricates empty modules (attribute access returns a
    dummy class, nothing of them is ever executed on the path we run),
  * makes ``Tensor.cuda`` the identity for the duration of the run,
and then calls the unmodified classes ``MscEvalV0`` (evaluate.py:46-99), ``MscEvalV0_Contrast`` (:101-192, both
``ori_scales`` settings) and ``MscEvalV0_AutoLink`` (:582-640) with a small seeded network and data loader.
Recorded per case: the label batches, every logits tensor the network returned (in call order — the GPU test
replays them through the drop-in evaluator), every ``np.bincount`` result the evaluator produced (= the
per-batch confusion matrices) and the returned mIoU / arg-max vectors.
"""
import importlib.abc
import importlib.machinery
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
STUBS = ("timm", "yacs", "munkres", "clip", "nvidia", "cvcuda", "torchnvjpeg", "torchvision", "ot", "nvcv",
         "pycuda", "tensorrt", "matplotlib", "seaborn", "PIL", "skimage", "tensorboardX", "thop")


class _Dummy:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Dummy()

    def __getattr__(self, n):
        return _Dummy()


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return type(name, (_Dummy,), {})


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path, target=None):
        if fullname.split(".")[0] in STUBS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


def import_reference_evaluate():
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    os.chdir(REF)
    sys.meta_path.insert(0, _StubFinder())
    import evaluate  # the reference's own file, unmodified
    evaluate.tqdm = lambda x: x
    return evaluate


class RecordingNet:
    """A seeded stride-4 convolution standing in for the segmentation net; keeps every output it returns."""

    def __init__(self, g, heads, as_list, gain=4.0):
        self.w = [torch.randn(c, 3, 3, 3, generator=g) * gain for c in heads]
        self.b = [torch.randn(c, generator=g) for c in heads]
        self.as_list, self.calls = as_list, []

    def _head(self, im, k):
        return F.conv2d(im, self.w[k], self.b[k], stride=4, padding=1)

    def __call__(self, im, dataset=None):
        if dataset is None:  # AutoLink: one logits tensor per dataset head
            outs = [self._head(im, k) for k in range(len(self.w))]
            self.calls.append(outs)
            return outs
        out = self._head(im, 0)
        self.calls.append([out])
        return [out] if self.as_list else out


class FakeConfiger:
    def __init__(self, n_cats):
        self.n_cats = n_cats

    def get(self, *key):
        if key == ("n_datasets",):
            return len(self.n_cats)
        if len(key) == 2 and key[1] == "n_cats":
            return self.n_cats[int(key[0][len("dataset"):]) - 1]
        raise KeyError(key)


def make_loader(g, n_batches, N, C, H, W, p_ignore=0.08):
    dl = []
    for _ in range(n_batches):
        im = torch.randn(N, 3, H, W, generator=g)
        # piecewise-constant labels with noise, like a segmentation map
        coarse = torch.randint(0, C, (N, 1, H // 8, W // 8), generator=g).float()
        lb = F.interpolate(coarse, size=(H, W), mode="nearest").long()
        noise = torch.rand(N, 1, H, W, generator=g)
        lb[noise < 0.1] = torch.randint(0, C, (int((noise < 0.1).sum()),), generator=g)
        lb[torch.rand(N, 1, H, W, generator=g) < p_ignore] = 255
        dl.append((im, lb))
    return dl


def run_case(ev, out, tag, evaluator, net, dl, n_classes, dataset_id):
    hists = []
    real_bincount = np.bincount

    def spy(x, *a, **k):
        r = real_bincount(x, *a, **k)
        hists.append(r.copy())
        return r

    real_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    np.bincount = spy
    try:
        with torch.no_grad():
            result = evaluator(net, dl, n_classes, dataset_id)
    finally:
        np.bincount = real_bincount
        torch.Tensor.cuda = real_cuda
    out[f"{tag}_n_batches"] = np.int64(len(dl))
    for i, (_, lb) in enumerate(dl):
        out[f"{tag}_label{i}"] = lb.numpy().astype(np.uint8)
        out[f"{tag}_imshape{i}"] = np.array(dl[i][0].shape, dtype=np.int64)
    out[f"{tag}_n_calls"] = np.int64(len(net.calls))
    for i, outs in enumerate(net.calls):
        for k, t in enumerate(outs):
            out[f"{tag}_call{i}_head{k}"] = t.numpy()
    out[f"{tag}_n_hists"] = np.int64(len(hists))
    for i, h in enumerate(hists):
        out[f"{tag}_hist{i}"] = h.astype(np.int64)
    return result


def main():
    ev = import_reference_evaluate()
    g = torch.Generator().manual_seed(20261019)
    out = {}

    # 1. MscEvalV0: six scales x flip, 19 classes (evaluate.py:46-99; the ms_flip setting of eval_model)
    C, H, W = 19, 64, 96
    net = RecordingNet(g, [C], as_list=True)
    dl = make_loader(g, 2, 2, C, H, W)
    miou = run_case(ev, out, "v0", ev.MscEvalV0((0.5, 0.75, 1, 1.25, 1.5, 1.75), True), net, dl, C, 0)
    out["v0_miou"] = np.float64(miou)

    # 2. MscEvalV0_Contrast, ori_scales=True, two scales x flip, 37 classes (evaluate.py:101-192)
    C = 37
    net = RecordingNet(g, [C], as_list=False)
    dl = make_loader(g, 2, 1, C, H, W)
    miou = run_case(ev, out, "c_ori", ev.MscEvalV0_Contrast(None, (0.5, 1.0), True, ori_scales=True), net, dl, C, 3)
    out["c_ori_miou"] = np.float64(miou)

    # 3. MscEvalV0_Contrast as eval_model_contrast builds it (evaluate.py:1127): (0.5,), no flip, ori_scales=False
    #    -> the label is nearest-resized to the logits (legacy torch 'nearest'), probabilities stay low-res
    C, H3, W3 = 26, 128, 192
    net = RecordingNet(g, [C], as_list=False)
    dl = make_loader(g, 3, 2, C, H3, W3)
    miou = run_case(ev, out, "c_low", ev.MscEvalV0_Contrast(None, (0.5,), False), net, dl, C, 4)
    out["c_low_miou"] = np.float64(miou)

    # 4. a class that never occurs and is never predicted -> NaN IoU skipped by nanmean (evaluate.py:96-98)
    C = 12
    net = RecordingNet(g, [C], as_list=True)
    net.b[0][C - 1] = -1e4
    dl = make_loader(g, 1, 2, C - 1, H, W)
    miou = run_case(ev, out, "absent", ev.MscEvalV0((1.0,), False), net, dl, C, 0)
    out["absent_miou"] = np.float64(miou)

    # 5. MscEvalV0_AutoLink: rectangular [n_classes, n_cats_k] histograms against the other datasets' heads, row arg-max
    n_cats = [19, 12, 36]
    net = RecordingNet(g, n_cats, as_list=False)
    dl = make_loader(g, 2, 2, n_cats[1], H, W)
    res = run_case(ev, out, "autolink", ev.MscEvalV0_AutoLink(FakeConfiger(n_cats), (1.0,), False), net, dl, n_cats[1], 1)
    for k, r in enumerate(res):
        out[f"autolink_argmax{k}"] = r.numpy().astype(np.int64)

    golden_find_use_and_unuse(ev, out)

    path = os.path.join(OUT, "evaluator.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", {k: float(out[k]) for k in out if k.endswith("_miou")})




# ---- eval_find_use_and_unuse_label (evaluate.py:1788-1930): [n_cats, C_uni] histograms -> target_bi_graph ----------
class _ProtoNet:
    """Stands in for the segmentation net: emits seeded embeddings; carries unify_prototype / bipartite_graphs."""

    def __init__(self, g, n_cats, c_uni, dim):
        self.aux_mode = 'eval'
        self.g, self.dim = g, dim
        self.unify_prototype = torch.randn(c_uni, dim, generator=g)
        self.bipartite_graphs = []
        for c in n_cats:  # 0/1 column-one-hot graphs with a few empty columns (max_value == 0 -> skipped, :1894)
            idx = torch.randint(0, c, (c_uni,), generator=g)
            m = torch.zeros(c, c_uni)
            m[idx, torch.arange(c_uni)] = 1
            m[:, torch.rand(c_uni, generator=g) < 0.1] = 0
            self.bipartite_graphs.append(m)
        self.calls = []

    def eval(self):
        return self

    def __call__(self, im, dataset=None):
        # embeddings correlated with the label through the image's first channel (set by the loader below)
        N, _, H, W = im.shape
        emb = torch.randn(N, self.dim, H // 4, W // 4, generator=self.g)
        emb = emb + 2.0 * self.unify_prototype[im[:, 0, ::4, ::4].long().clamp(0, self.unify_prototype.shape[0] - 1)].permute(0, 3, 1, 2)
        self.calls.append(emb)
        return {'seg': emb}


def golden_find_use_and_unuse(ev, out):
    g = torch.Generator().manual_seed(20261021)
    n_cats, c_uni, dim, H, W = [5, 7], 16, 8, 32, 48
    net = _ProtoNet(g, n_cats, c_uni, dim)
    dls = []
    for i, c in enumerate(n_cats):
        dl = []
        for _ in range(2):
            coarse = torch.randint(0, c, (2, 1, H // 8, W // 8), generator=g).float()
            lb = F.interpolate(coarse, size=(H, W), mode="nearest").long()
            lb[torch.rand(2, 1, H, W, generator=g) < 0.08] = 255
            # the "image" carries, in channel 0, a unified class the label's class maps to most of the time
            cols = [torch.nonzero(net.bipartite_graphs[i][k]).flatten() for k in range(c)]
            im = torch.zeros(2, 3, H, W)
            lab0 = lb.squeeze(1).clone()
            lab0[lab0 == 255] = 0
            pick = torch.zeros_like(lab0)
            for k in range(c):
                if len(cols[k]):
                    sel = lab0 == k
                    choice = cols[k][torch.randint(0, len(cols[k]), (int(sel.sum()),), generator=g)]
                    # skew: the first column of a class gets ~70 % of its pixels
                    first = torch.rand(int(sel.sum()), generator=g) < 0.7
                    choice[first] = cols[k][0]
                    pick[sel] = choice
            im[:, 0] = pick.float()
            dl.append((im, lb))
        dls.append(dl)

    class Cfg:
        def get(self, *k):
            if k == ("n_datasets",):
                return len(n_cats)
            if k == ("loss", "ignore_index"):
                return 255
            if k == ("GNN", "unify_ratio"):
                return c_uni / float(sum(n_cats)) + 1e-9
            if len(k) == 2 and k[1] == "n_cats":
                return n_cats[int(k[0][len("dataset"):]) - 1]
            raise KeyError(k)

    ev.get_data_loader = lambda *a, **k: dls
    hists = []
    real_bincount = np.bincount

    def spy(x, *a, **k):
        r = real_bincount(x, *a, **k)
        hists.append(r.copy())
        return r

    real_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    np.bincount = spy
    try:
        heads, mious, target = ev.eval_find_use_and_unuse_label(Cfg(), net)
    finally:
        np.bincount = real_bincount
        torch.Tensor.cuda = real_cuda
    out["fuu_n_cats"] = np.array(n_cats)
    out["fuu_c_uni"] = np.int64(c_uni)
    out["fuu_proto"] = net.unify_prototype.numpy()
    k = 0
    for i, c in enumerate(n_cats):
        out[f"fuu_graph{i}"] = net.bipartite_graphs[i].numpy()
        out[f"fuu_target{i}"] = target[i].numpy()
        for b, (_, lb) in enumerate(dls[i]):
            out[f"fuu_label{i}_{b}"] = lb.numpy().astype(np.uint8)
            out[f"fuu_emb{i}_{b}"] = net.calls[k].numpy()
            out[f"fuu_hist{i}_{b}"] = hists[k].astype(np.int64)
            k += 1
    assert k == len(hists) == len(net.calls)
    print("find_use_and_unuse:", [(int((t == 0).sum()), int((t == 1).sum()), int((t == 255).sum())) for t in target])


if __name__ == "__main__":
    main()
