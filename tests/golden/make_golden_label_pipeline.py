"""Generate tests/golden/label_pipeline.npz with the REAL lib/transform_cv2.py of /root/reference (and real cv2).

    python tests/golden/make_golden_label_pipeline.py     # build container only (needs /root/reference and cv2)

Per case: a raw uint8 label image, a lb_map LUT, np.random.seed(seed), then — exactly what a DataLoader worker does
(lib/base_dataset.py:78-95 with get_dataloader.TransformationTrain, lib/get_dataloader.py:44-60) —
    label = lb_map[label];  im_lb = Compose([RandomResizedCrop(scales, size), RandomHorizontalFlip(), ColorJitter(.4,.4,.4)])(im_lb)
    lb = ToTensor()(im_lb)['lb']
for several samples in a row from ONE random stream.  Recorded: inputs, seed, and the int64 label tensors.
"""
import os
import sys

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    import lib.transform_cv2 as T  # the reference's own transforms (needs cv2)
    out = {}
    gen = np.random.default_rng(7)
    mapi = np.load(os.path.join(REF, "mapi_relabel.npy"))  # a real uint8[256] lb_map
    cases = [
        # tag, source sizes, scales, crop size: up-scaling + crop, padding (small sources), identity-size early return
        ("crop", [(96, 160), (120, 200), (1100, 1200)], (0.5, 1.0), (64, 96)),
        ("pad", [(1090, 1100), (1200, 1085), (1085, 1090)], (0.03, 0.06), (64, 96)),
        ("wide", [(300, 700), (250, 650)], (0.75, 2.0), (256, 512)),
    ]
    for ci, (tag, shapes, scales, size) in enumerate(cases):
        seed = 1000 + ci
        lut = mapi if ci != 1 else gen.integers(0, 256, 256).astype(np.uint8)
        trans = T.Compose([T.RandomResizedCrop(scales, size), T.RandomHorizontalFlip(),
                           T.ColorJitter(brightness=0.4, contrast=0.4, saturation=0.4)])
        to_tensor = T.ToTensor()
        np.random.seed(seed)
        out[f"{tag}_seed"], out[f"{tag}_scales"], out[f"{tag}_size"] = np.int64(seed), np.array(scales), np.array(size)
        out[f"{tag}_lut"] = lut
        out[f"{tag}_n"] = np.int64(len(shapes))
        for k, (H, W) in enumerate(shapes):
            coarse = gen.integers(0, 66, (H // 8 + 1, W // 8 + 1)).astype(np.uint8)
            raw = np.kron(coarse, np.ones((8, 8), dtype=np.uint8))[:H, :W].copy()
            noise = gen.random((H, W)) < 0.05
            raw[noise] = gen.integers(0, 256, int(noise.sum())).astype(np.uint8)
            im = gen.integers(0, 256, (H, W, 3)).astype(np.uint8)
            label = lut[raw]                                    # base_dataset.py:81-82
            res = to_tensor(trans(dict(im=im, lb=label)))       # base_dataset.py:90-93
            out[f"{tag}_raw{k}"] = raw
            out[f"{tag}_lb{k}"] = res['lb'].numpy().astype(np.uint8)
            assert res['lb'].dtype.__str__() == "torch.int64" and tuple(res['lb'].shape) == tuple(size), res['lb'].shape
    path = os.path.join(OUT, "label_pipeline.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
