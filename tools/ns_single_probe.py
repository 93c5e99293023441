"""Per-kernel timing of a Newton-Schulz chain on ONE shape group (the embedding matrix a rank owns alone at 8 GPUs).

    python tools/ns_single_probe.py [rows cols batch]     (B200_NS_GROUPED=0 for the per-group chain)
"""
import os
import sys
from collections import defaultdict
from pathlib import Path

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from mlx_cuda_distributed_pretraining_b200 import ops  # noqa: E402


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 32003
    cols = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    batch = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    torch.manual_seed(0)
    g = torch.randn(batch, rows, cols, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        ops.zeropower_groups([g])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for _ in range(n):
        ops.zeropower_groups([g])
    e1.record()
    torch.cuda.synchronize()
    m, k = min(rows, cols), max(rows, cols)
    flops = batch * 5 * (4.0 * m * m * k + 2.0 * m ** 3)
    ms = e0.elapsed_time(e1) / n
    print(f"grouped={os.environ.get('B200_NS_GROUPED', '1')} [{batch},{rows},{cols}] whole call (incl. sumsq/scales): "
          f"{ms:.3f} ms, {flops / ms / 1e9:.0f} TF/s")
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        ops.zeropower_groups([g])
        torch.cuda.synchronize()
    ev = sorted((e.time_range.start, e.time_range.end, e.name) for e in prof.events()
                if e.device_type == torch.autograd.DeviceType.CUDA)
    per = defaultdict(lambda: [0, 0.0])
    for s, e, name in ev:
        p = per[name[:90]]
        p[0] += 1
        p[1] += e - s
    print(f"  traced: span {ev[-1][1] - ev[0][0]:.0f} us, busy {sum(e - s for s, e, _ in ev):.0f} us, {len(ev)} kernels")
    for name, (cnt, us) in sorted(per.items(), key=lambda kv: -kv[1][1]):
        print(f"  {us:8.1f} us  x{cnt:<3d} {name}")
    print("  sequence:", " ".join(f"{e - s:.0f}" for s, e, _ in ev))


if __name__ == "__main__":
    main()
