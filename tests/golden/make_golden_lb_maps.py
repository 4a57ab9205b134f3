"""Generate tests/golden/lb_maps.npz: the lb_map LUTs of the seven dataset readers of ltbgnn_7_datasets_snp and the
output of their REAL ``__getitem__`` label path (cv2.imread(lbpth, 0) -> ``label = self.lb_map[label]``,
lib/base_dataset.py:78-84) on a synthetic label PNG.

    python tests/golden/make_golden_lb_maps.py      # build container only (needs /root/reference and cv2)

The reader classes are instantiated unmodified (CityScapes, Mapi, Sunrgbd, Bdd100k, Idd, ade2016, Coco_data) on a
temporary one-line annotation file, in mode 'ret_path' (returns the remapped label before any transform).
Missing third-party imports of sibling modules are served by the import hook of make_golden_eval.py.
"""
import importlib
import os
import sys
import tempfile

import numpy as np

sys.dont_write_bytecode = True
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden_eval import _StubFinder, REF, OUT  # noqa: E402

READERS = [("lib.cityscapes_cv2", "CityScapes"), ("lib.Mapi", "Mapi"), ("lib.sunrgbd", "Sunrgbd"),
           ("lib.bdd100k_data", "Bdd100k"), ("lib.idd_cv2", "Idd"), ("lib.ade2016_data", "ade2016"),
           ("lib.coco_data", "Coco_data")]  # dataset order of configs/ltbgnn_7_datasets_snp.json:40-109


def main():
    import cv2
    sys.path.insert(0, REF)
    os.chdir(REF)
    sys.meta_path.insert(0, _StubFinder())
    tmp = tempfile.mkdtemp()
    rng = np.random.default_rng(5)
    raw = np.concatenate([np.arange(256, dtype=np.uint8).reshape(4, 64),
                          rng.integers(0, 256, (60, 64)).astype(np.uint8)])  # every value at least once
    cv2.imwrite(os.path.join(tmp, "lb.png"), raw)
    cv2.imwrite(os.path.join(tmp, "im.png"), np.zeros((64, 64, 3), np.uint8))
    ann = os.path.join(tmp, "ann.txt")
    with open(ann, "w") as f:
        f.write("im.png,lb.png\n")
    out = {"raw": raw, "names": np.array([c for _, c in READERS])}
    for i, (mod, cls) in enumerate(READERS):
        ds = getattr(importlib.import_module(mod), cls)(tmp, ann, trans_func=None, mode='ret_path')
        _, label, _ = ds[0]
        out[f"lb_map{i}"] = np.asarray(ds.lb_map)
        out[f"label{i}"] = np.asarray(label)
        out[f"n_cats{i}"] = np.int64(ds.n_cats)
        assert out[f"lb_map{i}"].dtype == np.uint8 and out[f"label{i}"].dtype == np.uint8
    path = os.path.join(OUT, "lb_maps.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), [int(out[f"n_cats{i}"]) for i in range(len(READERS))])


if __name__ == "__main__":
    main()
