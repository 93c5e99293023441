#!/usr/bin/env python
"""Why the C1 loss curve cannot be compared step by step during its first ~30 steps: with the config's own
hyperparameters (lr 2e-2, AdamW without bias correction, no warm-up) the loss blows up to ~9-10 around steps 3-8 before
it settles, and that transient is CHAOTIC -- the oracle run in fp64 instead of fp32, or with initial weights perturbed
by 1e-6 relative, lands O(1) away at steps 6-8 (output committed as profiles/r02_c1_transient_chaos.txt).
CPU only:  python tools/c1_transient_chaos.py"""
import sys, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from oracle import reference_math as R
torch.set_num_threads(4)
C=dict(hidden=128, inter=256, layers=4, heads=8, kv_heads=8, head_dim=16, vocab_normal=256, batch=16, seq=1024)
d=R.LlamaDims(C['hidden'],C['inter'],C['layers'],C['heads'],C['kv_heads'],C['head_dim'],259)
sched=R.make_schedule({"type":"cosine","min_lr_ratio":0.01},2e-2,1000)
def run(dtype, n, perturb=0.0):
    params={k:v.to(dtype) for k,v in R.init_params(d,42).items()}
    if perturb:
        g=torch.Generator().manual_seed(1)
        params={k:v*(1+perturb*torch.randn(v.shape,generator=g,dtype=torch.float32).to(dtype)) for k,v in params.items()}
    opt=R.AdamWOracle(sched,betas=(0.9,0.999),eps=1e-8,weight_decay=0.01)
    out=[]
    for s in range(n):
        b=R.synthetic_batch(s,0,16,1024,256)
        loss,_,g=R.loss_and_grads(params,b,d,pad_token=256)
        opt.update(params,g); out.append(float(loss))
    return out
n=12
a=run(torch.float32,n); b=run(torch.float64,n); c=run(torch.float32,n,1e-6)
for i in range(n): print(i, round(a[i],4), round(b[i],4), round(c[i],4), 'f32-f64', round(abs(a[i]-b[i]),4), 'f32-perturbed(1e-6)', round(abs(a[i]-c[i]),4))
