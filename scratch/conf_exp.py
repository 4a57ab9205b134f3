import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdseg_b200 import ops
dev="cuda:0"
n_cats=[19,64,37,19,26,150,133]; ids=[0,0,0,1,1,1,2,2,3,3,4,4,5,5,6,6]
g=torch.Generator(device=dev).manual_seed(1)
B,H,W=16,1024,2048
lab=torch.stack([torch.randint(0,n_cats[d],(H,W),generator=g,device=dev) for d in ids]); lab[torch.rand(B,H,W,generator=g,device=dev)<0.05]=255
pred=torch.stack([torch.randint(0,n_cats[d],(H,W),generator=g,device=dev) for d in ids])
flush=torch.empty(256<<20,dtype=torch.uint8,device=dev)
hist,_=ops.confusion_images(lab,pred,ids,n_cats)
ref=torch.cat([torch.bincount(lab[torch.tensor(ids,device=dev)==d][lab[torch.tensor(ids,device=dev)==d]!=255]*c+pred[torch.tensor(ids,device=dev)==d][lab[torch.tensor(ids,device=dev)==d]!=255],minlength=c*c) for d,c in enumerate(n_cats)])
ts=[]
for _ in range(6):
    flush.fill_(1); hist.zero_()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record(); ops.confusion_images(lab,pred,ids,n_cats,hist=hist); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print(os.environ.get("MDSEG_CONF_RED_EVERY"), sorted(ts)[len(ts)//2], bool(torch.equal(hist,ref)), int(hist.sum()))
