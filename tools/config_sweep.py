#!/usr/bin/env python
"""Runs a few optimizer steps of every BASELINE config (configs/c1..c5) at FULL per-GPU size on one
B200 through the public Trainer: checks that each workload fits, steps without error and that the
loss is finite; prints step time, tokens/s and peak memory.  For Shampoo the preconditioner start
and period are pulled forward so the matrix-root path runs inside the sample.

Usage (under gpurun):  python tools/config_sweep.py [c1 c2 ...] [--steps N]
"""
import argparse
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
import yaml

from mlx_cuda_distributed_pretraining_b200.core.training import Config, Trainer


def run(path: Path, steps: int) -> dict:
    d = yaml.safe_load(path.read_text())
    d["name"] = f"sweep-{path.stem}"
    d["overwrite"] = True
    d["system"]["distributed"] = False
    opt = d["training"]["optimization"]
    if opt.get("optimizer") == "shampoo":
        opt["start_preconditioning_step"] = 2
        opt["update_period"] = 3
    k = int(d["training"]["hyperparameters"].get("gradient_accumulation_steps") or 1)
    d["training"]["hyperparameters"]["iters"] = steps * k
    d["logging"]["steps"]["checkpoint_interval"] = 0
    cfg = Config.from_dict(d)
    torch.cuda.reset_peak_memory_stats()
    tr = Trainer(cfg, synthetic=True, quiet=True, run_root=str(ROOT / "gpurun_out" / "sweep_runs"))
    tr._accum_step, tr._accum_tokens = 0, 0
    losses, times = [], []
    tokens_per_micro = None
    for s in range(steps * k):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = tr.train_step(s)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
        loss = out[0] if isinstance(out, (tuple, list)) else out
        losses.append(float(loss))
        if tokens_per_micro is None:
            b = tr.data_manager.generate_batch(0)
            tokens_per_micro = b.shape[0] * (b.shape[1] - 1)
    warm = times[k * 2:] if len(times) > k * 3 else times[-k:]
    per_update = sum(warm) / (len(warm) / k)
    res = {"config": path.stem, "optimizer": opt.get("optimizer"), "micro_steps": len(times), "accum": k,
           "first_loss": losses[0], "last_loss": losses[-1], "finite": all(map(lambda v: v == v and abs(v) < 1e9, losses)),
           "ms_per_update": 1e3 * per_update, "tokens_per_s": tokens_per_micro * k / per_update,
           "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
    del tr
    torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("which", nargs="*", default=[])
    ap.add_argument("--steps", type=int, default=6)
    a = ap.parse_args()
    out = []
    for p in sorted((ROOT / "configs").glob("c*.yaml")):
        if a.which and not any(p.stem.startswith(w) for w in a.which):
            continue
        try:
            r = run(p, a.steps)
        except Exception as e:  # report and continue with the next config
            r = {"config": p.stem, "error": f"{type(e).__name__}: {e}"[:400]}
        print(json.dumps(r), flush=True)
        out.append(r)
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "config_sweep.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
