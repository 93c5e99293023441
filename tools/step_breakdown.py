"""torch.profiler breakdown of one C2 bench step: GPU-busy time vs wall, top kernels (real timings,
not ncu-serialised)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
from torch.profiler import ProfilerActivity, profile
import bench
from mlx_cuda_distributed_pretraining_b200.core.training import Config, Trainer

cfg = Config.from_dict(bench.c2_config("breakdown", False))
tr = Trainer(cfg, synthetic=True, quiet=True, run_root=str(ROOT / "gpurun_out" / "bench_runs"))
tr._accum_step, tr._accum_tokens = 0, 0
batches = [tr.data_manager.generate_batch(s).cuda() for s in range(8)]
def step(s):
    b = batches[s]
    loss, _ = tr.compute_loss(tr.model, b[:, :-1], b[:, 1:])
    loss.backward()
    tr.optimizer.update(tr.model)
    tr.store.zero_grad()
for s in range(4):
    step(s)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for s in range(4, 8):
        step(s)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
tot = sum(e.device_time for e in ev) / 4 / 1e3 if hasattr(ev[0], "device_time") else sum(e.cuda_time for e in ev) / 4 / 1e3
t0 = min(e.time_range.start for e in ev); t1 = max(e.time_range.end for e in ev)
print(f"GPU kernel time per step: {tot:.2f} ms; GPU span per step: {(t1 - t0) / 4 / 1e3:.2f} ms; kernels/step: {len(ev) // 4}")
import collections, re
agg = collections.defaultdict(lambda: [0, 0.0])
for e in ev:
    name = re.sub(r"void |b200::|\(anonymous namespace\)::|at::native::|<unnamed>::", "", e.name)
    name = re.sub(r"\(.*", "", name)[:90]
    agg[name][0] += 1
    agg[name][1] += (e.device_time if hasattr(e, "device_time") else e.cuda_time)
lines = []
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    lines.append(f"{t / 4 / 1e3:8.3f} ms  {n // 4:5d}x  {name}")
out = "\n".join(lines)
print(out)
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "step_breakdown.txt").write_text(
    f"GPU kernel time per step: {tot:.2f} ms; span {(t1 - t0) / 4 / 1e3:.2f} ms; kernels/step {len(ev) // 4}\n" + out + "\n")
