"""Device-side idle time inside a training step, from a kineto (CUPTI) trace of a few steps.

    python tools/gpu_gaps.py [c2] [steps]   -> gpurun_out/gpu_gaps_<tag>.json + a one-line summary

Busy = union of kernel/memcpy/memset intervals on the device; gap = wall - busy, attributed to the kernel that
FOLLOWS each idle interval.  Tracing adds host overhead (so gaps are an upper bound on the untraced run's)."""
import json
import sys
from collections import defaultdict
from pathlib import Path

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "c2"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    from mlx_cuda_distributed_pretraining_b200.core.training import Config, Trainer
    d = bench.load_config(tag, distributed=False)
    c = bench.dims_of(d)
    k = c["accum"]
    d["training"]["hyperparameters"]["iters"] = 1000
    tr = Trainer(Config.from_dict(d), synthetic=True, quiet=True, run_root=str(ROOT / "gpurun_out" / "bench_runs"))
    tr._accum_step, tr._accum_tokens = 0, 0
    n = (steps + 4) * k
    batches = [tr.data_manager.generate_batch(s).to(tr.device) for s in range(n)]
    it = iter(range(n))

    def run(nsteps):
        for _ in range(nsteps * k):
            i = next(it)
            tr.micro_step(i, batches[i])

    run(3)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        run(steps)
        torch.cuda.synchronize()
    ev = [(e.time_range.start, e.time_range.end, e.name) for e in prof.events()
          if e.device_type == torch.autograd.DeviceType.CUDA]
    ev.sort()
    t0, t1 = ev[0][0], max(e[1] for e in ev)
    busy, cur_end = 0.0, ev[0][0]
    gaps = defaultdict(lambda: [0, 0.0])
    for s, e, name in ev:
        if s > cur_end:
            g = gaps[name.split("<")[0][:60]]
            g[0] += 1
            g[1] += s - cur_end
            cur_s = s
        else:
            cur_s = cur_end
        if e > cur_end:
            busy += e - max(cur_s, s) if s > cur_end else e - cur_end
            cur_end = e
    per = defaultdict(lambda: [0, 0.0])
    for s_, e_, name in ev:
        short = name.replace("(anonymous namespace)::", "").replace("void ", "").replace("b200::", "")
        q = per[short.split("(")[0][:70]]
        q[0] += 1
        q[1] += e_ - s_
    wall = t1 - t0
    top = sorted(gaps.items(), key=lambda kv: -kv[1][1])[:25]
    out = {"tag": tag, "steps": steps, "wall_us_per_step": wall / steps, "busy_us_per_step": busy / steps,
           "idle_us_per_step": (wall - busy) / steps, "kernels_per_step": len(ev) / steps,
           "kernel_time_in_situ_top": [{"kernel": kname, "count_per_step": v[0] / steps, "us_per_launch": v[1] / v[0],
                                        "us_per_step": v[1] / steps}
                                       for kname, v in sorted(per.items(), key=lambda kv: -kv[1][1])[:40]],
           "idle_before_kernel_top": [{"kernel": kname, "count_per_step": v[0] / steps, "idle_us_per_step": v[1] / steps}
                                      for kname, v in top]}
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / f"gpu_gaps_{tag}.json").write_text(json.dumps(out, indent=1))
    print({k2: (round(v, 1) if isinstance(v, float) else v) for k2, v in out.items()
           if k2 not in ("idle_before_kernel_top", "kernel_time_in_situ_top")})
    for r in out["kernel_time_in_situ_top"][:30]:
        print(f"    {r['us_per_step']:8.1f} us/step  {r['count_per_step']:5.1f} x {r['us_per_launch']:7.1f} us  {r['kernel']}")
    for r in out["idle_before_kernel_top"][:12]:
        print("   ", round(r["idle_us_per_step"], 1), "us idle before", r["count_per_step"], "x", r["kernel"])


if __name__ == "__main__":
    main()
