"""A/B of the two forward-attention kernels (B200_ATTN_FWD_PF is read once per process: run twice)."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from mlx_cuda_distributed_pretraining_b200 import ops
for (B, S, H, Hk, D) in ((16, 1024, 16, 8, 64), (16, 2048, 16, 8, 64), (64, 2048, 16, 16, 64), (32, 2048, 16, 16, 128)):
    q = torch.randn(B, S, H, D, device="cuda").to(torch.bfloat16); k = torch.randn(B, S, Hk, D, device="cuda").to(torch.bfloat16)
    v = torch.randn(B, S, Hk, D, device="cuda").to(torch.bfloat16)
    for _ in range(3):
        ops.attention_fwd_raw(q, k, v, D ** -0.5, True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        o, lse = ops.attention_fwd_raw(q, k, v, D ** -0.5, True)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    print(f"PF={os.environ.get('B200_ATTN_FWD_PF','-')} B{B} S{S} H{H}/{Hk} D{D}: {us:.1f} us  {4.0*B*H*S*S*D/us/1e6:.0f} TF/s full-count  checksum {float(o.float().sum()):.3f}")
