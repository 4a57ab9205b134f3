import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mdseg_b200 import ops
from oracle import f64
dev = "cuda:0"
g = torch.Generator().manual_seed(0)
n_cats, c_uni, ids = [5, 3, 7], 11, [2, 0, 1, 2, 0]
h, w, H, W = 16, 32, 64, 128
B = len(ids)
x = torch.randn(B, c_uni, h, w, generator=g) * 2.5
graphs = []
for c in n_cats:
    idx = torch.randint(0, c, (c_uni,), generator=g); idx[:c] = torch.arange(c)
    m = torch.zeros(c, c_uni); m[idx, torch.arange(c_uni)] = 1; graphs.append(m)
labels = torch.stack([torch.randint(0, n_cats[d], (H, W), generator=g) for d in ids])
labels[torch.rand(B, H, W, generator=g) < 0.05] = 255
thresh = ops.neg_log(0.4)
ref = f64.multi_dataset(x.numpy(), labels.numpy(), np.array(ids), [m.numpy() for m in graphs], thresh)
xd = x.to(dev).requires_grad_(True)
loss = ops.mds_proj_ohem_ce(xd, labels.to(dev), torch.tensor(ids, dtype=torch.int32, device=dev), [m.to(dev) for m in graphs], thresh)
torch.cuda.synchronize()
print("fwd ok", float(loss), ref["loss"])
loss.backward()
torch.cuda.synchronize()
err = np.abs(xd.grad.cpu().numpy() - ref["dlogits_uni"]).max() / np.abs(ref["dlogits_uni"]).max()
print("bwd ok rel err", err)
