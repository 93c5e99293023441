import sys; sys.path.insert(0, '.')
import torch
from mlx_cuda_distributed_pretraining_b200 import ops
def t(fn, warm=5, iters=50):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
for (M, N, K, b, bn) in [(256,256,64,1,256),(256,256,1024,1,256),(1024,1024,64,24,256),(1024,1024,256,24,256),(1024,1024,1024,24,256),(1024,1024,1024,24,1256),(512,512,1024,24,1256),(512,512,1024,24,128),(512,1024,512,24,256),(512,1024,512,24,128),(1024,1024,1024,74*1,256)]:
    a = torch.randn(b, M, K, device='cuda').to(torch.bfloat16); bb = torch.randn(b, N, K, device='cuda').to(torch.bfloat16)
    out = torch.empty(b, M, N, device='cuda', dtype=torch.bfloat16)
    us = t(lambda: ops.gemm(a, bb, out=out, force_bn=bn))
    print(f"M{M} N{N} K{K} b{b} bn{bn}: {us:.1f} us  {2*M*N*K*b/us/1e6:.0f} TF (full-count)")
# launch overhead of an empty-ish kernel for reference
x = torch.zeros(8, device='cuda')
print("torch tiny add:", t(lambda: x.add_(1)), "us")
