"""Stand-alone timing of the fused MLP GEMMs against the library GEMM + elementwise pairs they replace (C2 / C3 shapes)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from mlx_cuda_distributed_pretraining_b200 import ops

def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3  # us

for (M, K, I) in ((16384, 1024, 2816), (32768, 1024, 4096)):
    x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    w2 = (torch.randn(2 * I, K, device="cuda") * K ** -0.5).to(torch.bfloat16)
    wd = (torch.randn(K, I, device="cuda") * I ** -0.5).to(torch.bfloat16)
    gu = torch.empty(M, 2 * I, device="cuda", dtype=torch.bfloat16)
    y = torch.empty(M, I, device="cuda", dtype=torch.bfloat16)
    dgu = torch.empty_like(gu)
    dy = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    L = ops.lib()
    st = ops._stream()
    fl = 2.0 * M * 2 * I * K
    t_f = timeit(lambda: L.b200_mlp_gateup_glu_fwd(x.data_ptr(), w2.data_ptr(), gu.data_ptr(), y.data_ptr(), M, K, I, st))
    t_g1 = timeit(lambda: torch.matmul(x, w2.t(), out=gu))
    g, u = torch.empty(M, I, device="cuda", dtype=torch.bfloat16), torch.empty(M, I, device="cuda", dtype=torch.bfloat16)
    t_g2 = timeit(lambda: (torch.matmul(x, w2[:I].t(), out=g), torch.matmul(x, w2[I:].t(), out=u)))
    t_glu = timeit(lambda: L.b200_glu_fwd(g.data_ptr(), u.data_ptr(), y.data_ptr(), g.numel(), st))
    t_plain = timeit(lambda: ops.gemm(x, w2, out=gu.view(1, M, 2 * I)))
    t_b = timeit(lambda: L.b200_mlp_down_glu_bwd(dy.data_ptr(), wd.data_ptr(), gu.data_ptr(), dgu.data_ptr(), M, K, I, st))
    dmid = torch.empty(M, I, device="cuda", dtype=torch.bfloat16)
    t_bd = timeit(lambda: torch.matmul(dy, wd, out=dmid))
    dg, du = torch.empty_like(g), torch.empty_like(u)
    t_bglu = timeit(lambda: L.b200_glu_bwd(dmid.data_ptr(), g.data_ptr(), u.data_ptr(), dg.data_ptr(), du.data_ptr(), g.numel(), st))
    print(f"M={M} K={K} I={I}: fused fwd {t_f:.0f} us ({fl / t_f / 1e6:.0f} TF/s) | in-tree plain GEMM N=2I {t_plain:.0f} | cuBLAS one N=2I {t_g1:.0f}, two N=I {t_g2:.0f}"
          f" + glu_fwd {t_glu:.0f} || fused bwd {t_b:.0f} us ({fl / 2 / t_b / 1e6:.0f} TF/s) | cuBLAS dgrad {t_bd:.0f} + glu_bwd {t_bglu:.0f}")
