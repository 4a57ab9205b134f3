import sys, os
sys.path.insert(0, "/root/repo")
import torch
from mdseg_b200 import ops
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(1)
x = (torch.randn(16, 171, 640, 640, generator=g, device=dev) * 3).to(torch.bfloat16).requires_grad_(True)
lb = torch.randint(0, 171, (16, 640, 640), generator=g, device=dev)
lb[torch.rand(16, 640, 640, generator=g, device=dev) < 0.05] = 255
for _ in range(2):
    x.grad = None
    l = ops.ohem_ce(x, lb, ops.neg_log(0.7)); l.backward()
torch.cuda.synchronize()
