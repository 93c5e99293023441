"""Debug: per-phase clock64 stamps of attention-backward CTA 0 (kv tile 0, kv head 0, batch 0)."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from mlx_cuda_distributed_pretraining_b200 import ops
D = int(sys.argv[1]) if len(sys.argv) > 1 else 64
B, S, H, Hk = (16, 1024, 16, 8) if D == 64 else (8, 2048, 16, 16)
q = torch.randn(B, S, H, D, device="cuda").to(torch.bfloat16); k = torch.randn(B, S, Hk, D, device="cuda").to(torch.bfloat16)
v = torch.randn(B, S, Hk, D, device="cuda").to(torch.bfloat16)
o, lse = ops.attention_fwd_raw(q, k, v, D ** -0.5, True)
do = torch.randn_like(o)
trace = torch.zeros(16 * 16, dtype=torch.int64, device="cuda")
for _ in range(3):
    ops.attention_bwd_raw(q, k, v, o, do, lse, D ** -0.5, True)
os.environ["B200_ATTN_TRACE"] = hex(trace.data_ptr())
ops.attention_bwd_raw(q, k, v, o, do, lse, D ** -0.5, True)
torch.cuda.synchronize()
t = trace.cpu().view(16, 16)
if D == 64:
    names = {0: "sm.top", 1: "sm.s_full", 4: "sm.phaseA_done", 6: "sm.dp_full", 2: "sm.pds_arrive", 3: "sm.lse_put", 5: "sm.bar",
             8: "mma.top", 9: "mma.S(it+1) issued", 10: "mma.pds_full", 11: "mma.dq_empty", 12: "mma.dVdKdQ issued",
             13: "drain.dq_full", 14: "drain.reds issued"}
else:
    names = {0: "sm.top", 1: "sm.s_full", 4: "sm.phaseA_done", 5: "sm.dp_full", 6: "sm.tiles_free", 2: "sm.pds_arrive",
             8: "mma.top", 9: "mma.S(it+1) issued", 10: "mma.pds_full", 11: "mma.dVdQdK issued", 12: "mma.dq_empty",
             13: "drain.dq_full", 14: "drain.done"}
base = int(t[0, 8])
for it in range(6):
    row = {names[s]: int(t[it, s]) - base for s in sorted(names) if int(t[it, s])}
    print(it, row)
