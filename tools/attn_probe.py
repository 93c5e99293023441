import sys, torch
sys.path.insert(0, "/root/repo")
from mlx_cuda_distributed_pretraining_b200 import ops
torch.manual_seed(0)
def ref(q, k, v, scale, causal):
    B, S, H, D = q.shape
    Hk = k.shape[2]
    qf, kf, vf = q.float(), k.float().repeat_interleave(H // Hk, 2), v.float().repeat_interleave(H // Hk, 2)
    s = torch.einsum("bqhd,bkhd->bhqk", qf, kf) * scale
    if causal:
        s = s + torch.triu(torch.full((S, S), float("-inf"), device=q.device), 1)
    p = torch.softmax(s, -1)
    return torch.einsum("bhqk,bkhd->bqhd", p, vf)
for (B, S, H, Hk, D) in [(2, 128, 4, 2, 32), (2, 128, 4, 2, 64), (2, 256, 4, 2, 64), (1, 128, 1, 1, 64)]:
    for mag in (1.0, 4.0, 16.0):
        for trial in range(3):
            q = (torch.randn(B, S, H, D, device="cuda") * mag).to(torch.bfloat16)
            k = (torch.randn(B, S, Hk, D, device="cuda") * mag).to(torch.bfloat16)
            v = torch.randn(B, S, Hk, D, device="cuda").to(torch.bfloat16)
            o = ops.attention(q, k, v, D ** -0.5, True)
            r = ref(q, k, v, D ** -0.5, True)
            err = ((o.float() - r).norm() / r.norm()).item()
            bad = (o.float() - r).abs().amax(dim=(0, 2, 3))
            print(B, S, H, Hk, D, "mag", mag, "rel", f"{err:.2e}", "worst row", int(bad.argmax()), f"{bad.max().item():.3f}",
                  "nan" if torch.isnan(o.float()).any() else "")
