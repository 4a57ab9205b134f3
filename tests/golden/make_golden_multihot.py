"""Generate tests/golden/multihot.npz by running the REAL reference ClassRemapOneHotLabel
(lib/class_remap.py:232-276) in the build container.  Run from anywhere:  python tests/golden/make_golden_multihot.py"""
import json
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    tm, tmm, tml = types.ModuleType("timm"), types.ModuleType("timm.models"), types.ModuleType("timm.models.layers")
    tml.trunc_normal_ = torch.nn.init.trunc_normal_
    sys.modules.update({"timm": tm, "timm.models": tmm, "timm.models.layers": tml})
    os.chdir(REF)
    from lib.class_remap import ClassRemapOneHotLabel
    from tools.configer import Configer
    g = torch.Generator().manual_seed(77)
    cases = {}
    for tag, cfgfile, n_ds in (("test", "configs/test/test.json", 2), ("cca", "configs/bisenetv2_city_cam_a2d2.json", 3)):
        configer = Configer(configs=cfgfile)
        if not configer.exists('contrast', 'update_sim_thresh'):
            configer.add(('contrast', 'update_sim_thresh'), 0.5) if configer.exists('contrast') else None
        cr = ClassRemapOneHotLabel(configer)
        raw = json.load(open(cfgfile))
        for d in range(n_ds):
            n_cats_d = raw[f"dataset{d + 1}"]["n_cats"]
            lb = torch.randint(0, n_cats_d + 2, (2, 7, 11), generator=g)  # includes values that are not keys
            lb[0, 0, :3] = 255
            cases[f"{tag}_d{d}_labels"] = lb.numpy()
            cases[f"{tag}_d{d}_multi"] = cr.SegRemapping(lb, d).numpy()
            cases[f"{tag}_d{d}_single"] = cr.SingleSegRemappingOneHot(lb, d).numpy()
    np.savez_compressed(os.path.join(OUT, "multihot.npz"), **cases)
    print({k: v.shape for k, v in cases.items()})


if __name__ == "__main__":
    main()
