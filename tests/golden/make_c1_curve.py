#!/usr/bin/env python
"""Generates tests/golden/c1_curve_<data>.json: the ORACLE's loss curve for BASELINE configs[0]
("Llama 2M AdamW (configs/model-config-sample.yaml), 1k iters synthetic tokens ... plumbing + loss
parity"; dims/hyperparameters of configs/c1-llama2m-adamw.yaml), CPU fp32, step by step:

    params = init_params(seed 42); for step: batch = synthetic(step) -> loss_and_grads -> AdamWOracle.update

Run in the build container (CPU only; ~1.5 s/step on 8 cores):
    python tests/golden/make_c1_curve.py --data uniform --steps 1000
    python tests/golden/make_c1_curve.py --data markov  --steps 500 --total-steps 500
    python tests/golden/make_c1_curve.py --data markov  --steps 300 --total-steps 300 --lr 1e-3   (non-chaotic control)
The -m gpu test tests/test_gpu_training.py::test_c1_loss_curve_matches_oracle runs the product Trainer
on the same config and compares per-step losses with these files (tolerance stated there).
"""
import argparse
import json
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import reference_math as R  # noqa: E402

C1 = dict(hidden=128, inter=256, layers=4, heads=8, kv_heads=8, head_dim=16, vocab_normal=256, batch=16, seq=1024,
          lr=2e-2, min_lr_ratio=0.01, weight_decay=0.01, betas=(0.9, 0.999), eps=1e-8)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--data", choices=["uniform", "markov"], default="uniform")
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--total-steps", type=int, default=1000, help="schedule length (hyperparameters.iters)")
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--lr", type=float, default=None, help="override the config's learning rate (2e-2); the file tag gets _lr<value>")
    a = ap.parse_args()
    if a.threads:
        torch.set_num_threads(a.threads)
    c = dict(C1)
    if a.lr is not None:
        c["lr"] = a.lr
    d = R.LlamaDims(c["hidden"], c["inter"], c["layers"], c["heads"], c["kv_heads"], c["head_dim"], c["vocab_normal"] + 3)
    params = R.init_params(d, seed=42)
    sched = R.make_schedule({"type": "cosine", "min_lr_ratio": c["min_lr_ratio"]}, c["lr"], a.total_steps)
    opt = R.AdamWOracle(sched, betas=c["betas"], eps=c["eps"], weight_decay=c["weight_decay"])
    gen = R.synthetic_batch if a.data == "uniform" else R.synthetic_batch_markov
    losses = []
    t0 = time.time()
    tag = a.data + (f"_lr{a.lr:g}" if a.lr is not None else "")
    out_path = ROOT / "tests" / "golden" / f"c1_curve_{tag}.json"
    for step in range(a.steps):
        batch = gen(step, 0, c["batch"], c["seq"], c["vocab_normal"])
        loss, ntoks, grads = R.loss_and_grads(params, batch, d, pad_token=c["vocab_normal"])
        opt.update(params, grads)
        losses.append(float(loss))
        if step % 25 == 0 or step == a.steps - 1:
            print(f"step {step} loss {losses[-1]:.5f}  ({time.time() - t0:.0f}s)", flush=True)
            out_path.write_text(json.dumps({
                "config": "configs/c1-llama2m-adamw.yaml", "data": a.data, "lr": c["lr"], "steps": len(losses),
                "total_steps": a.total_steps, "generator": "tests/golden/make_c1_curve.py (oracle/reference_math.py, CPU fp32)",
                "param_checksum": float(sum(float(p.double().abs().sum()) for p in params.values())),
                "loss": losses}))


if __name__ == "__main__":
    main()
