"""Replays the fp32 AdamW tiny-model steps of tests/test_gpu_training.py and checks every attention call's
forward output and input gradients against a dense fp32 reference on the exact tensors of that step."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
from mlx_cuda_distributed_pretraining_b200 import ops
from mlx_cuda_distributed_pretraining_b200.core.training import Config, Trainer
from tests.smoke_check import tiny_config

orig = ops.attention
step_no = [0]

def ref(q, k, v, scale, causal):
    B, S, H, D = q.shape
    Hk = k.shape[2]
    kf, vf = k.repeat_interleave(H // Hk, 2), v.repeat_interleave(H // Hk, 2)
    s = torch.einsum("bqhd,bkhd->bhqk", q, kf) * scale
    if causal:
        s = s + torch.triu(torch.full((S, S), float("-inf"), device=q.device), 1)
    return torch.einsum("bhqk,bkhd->bqhd", torch.softmax(s, -1), vf), s

def checked(q, k, v, scale, causal):
    o = orig(q, k, v, scale, causal)
    qr, kr, vr = (t.detach().float().requires_grad_(True) for t in (q, k, v))
    r, s = ref(qr, kr, vr, scale, causal)
    fe = ((o.float() - r).norm() / r.norm()).item()
    g = torch.randn_like(r)
    q2, k2, v2 = (t.detach().clone().requires_grad_(True) for t in (q, k, v))
    o2 = orig(q2, k2, v2, scale, causal)
    o2.backward(g.to(o2.dtype))
    r.backward(g)
    errs = [((a.grad.float() - b.grad).norm() / b.grad.norm()).item() for a, b in ((q2, qr), (k2, kr), (v2, vr))]
    print(f"step {step_no[0]} attn shape {tuple(q.shape)} smax {s[torch.isfinite(s)].abs().max().item():.1f} "
          f"fwd {fe:.2e} dq {errs[0]:.2e} dk {errs[1]:.2e} dv {errs[2]:.2e} nan_o {bool(torch.isnan(o.float()).any())}",
          flush=True)
    return o

ops.attention = checked
import mlx_cuda_distributed_pretraining_b200.arch.flash_attention as fa
fa.ops.attention = checked
tr = Trainer(Config.from_dict(tiny_config(optimizer="adamw", mixed=False)), synthetic=True, quiet=True,
             run_root=str(ROOT / "gpurun_out" / "dbg_runs"))
tr._accum_step, tr._accum_tokens = 0, 0
for step in range(5):
    step_no[0] = step
    loss, _, _ = tr.train_step(step, tr.data_manager.generate_batch(step))
    print("step", step, "loss", float(loss), flush=True)
