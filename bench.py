#!/usr/bin/env python
"""bench.py -- headline benchmark: tokens/sec (device-timed) of the Llama + Muon training step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1], "C2"): Llama "80M" (hidden 1024, inter 2816, 12 layers, 16 q /
8 kv heads, head_dim 64, tied embeddings, V = 32000+3), Muon (Newton-Schulz, 5 steps) lr 3e-4 with
cosine+warmup, bf16 compute / fp32 masters, batch 16 x seq 1024 per GPU, synthetic tokens.
A "step" = fwd + bwd + gradient all-reduce (N>1) + full Muon update.  Weak scaling: per-GPU batch
fixed.  One JSON line on stdout (rank 0).

Keys beyond the base contract:
  roofline      Newton-Schulz GEMM chain, measured live with CUDA events on the launching stream
  kernels       the other hand-written kernel families timed the same way (attention fwd / bwd)
  cpu_baseline  CPU restatement of the reference step (oracle/), bounded sample, host cores
  e2e           same step through the public Trainer API with per-step pinned H2D + loss D2H
--impl reference times the reference's CPU path (the oracle port: MLX is not installable here).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

C2 = dict(hidden=1024, inter=2816, layers=12, heads=16, kv_heads=8, head_dim=64, vocab_normal=32000,
          batch=16, seq=1024, lr=3e-4, warmup=1000, min_lr_ratio=0.1, iters=10000)


def c2_config(name: str, distributed: bool) -> dict:
    c = C2
    return {
        "name": name, "overwrite": True,
        "data": {"input_file": "synthetic", "preprocessing": {"max_context_size": c["seq"], "chunk_overlap": 0},
                 "tokenizer": {"normal_vocab_size": c["vocab_normal"],
                               "special_tokens": {"pad": "<pad>", "bos": "<bos>", "eos": "<eos>"}}},
        "model": {"architecture": "llama",
                  "dimensions": {"hidden_size": c["hidden"], "intermediate_size": c["inter"], "num_layers": c["layers"]},
                  "attention": {"num_heads": c["heads"], "num_kv_heads": c["kv_heads"], "head_dim": c["head_dim"],
                                "max_position_embeddings": 4096, "use_flash_attention": True},
                  "normalization": {"rms_norm_eps": 1e-5},
                  "rope": {"theta": 10000, "traditional": False, "scaling": None},
                  "misc": {"attention_bias": False, "mlp_bias": False, "tie_word_embeddings": True}},
        "training": {"epochs": None,
                     "hyperparameters": {"batch_size": c["batch"], "learning_rate": c["lr"], "weight_decay": 0.01,
                                         "iters": c["iters"]},
                     "scheduler": {"type": "cosine_with_warmup", "min_lr_ratio": c["min_lr_ratio"],
                                   "warmup_steps": c["warmup"]},
                     "optimization": {"optimizer": "muon", "betas": [0.9, 0.95], "eps": 1e-8}},
        "logging": {"log_dir": "logs", "checkpoint_dir": "checkpoints", "steps": {"logging_interval": 10 ** 9,
                    "checkpoint_interval": 0, "validation_interval": 0}, "metrics": {}},
        "system": {"seed": 42, "device": "gpu", "distributed": distributed, "mixed_precision": True,
                   "precision": "bfloat16"},
    }


def ns_flops_per_step() -> float:
    """5*(4 m^2 n + 2 m^3) over every 2-D parameter: 4.129 TF for C2."""
    c = C2
    V = c["vocab_normal"] + 3
    mats = [(V, c["hidden"])]
    for _ in range(c["layers"]):
        mats += [(c["heads"] * c["head_dim"], c["hidden"]), (c["kv_heads"] * c["head_dim"], c["hidden"]),
                 (c["kv_heads"] * c["head_dim"], c["hidden"]), (c["hidden"], c["heads"] * c["head_dim"]),
                 (c["inter"], c["hidden"]), (c["inter"], c["hidden"]), (c["hidden"], c["inter"])]
    tot = 0.0
    for r, cc in mats:
        m, n = min(r, cc), max(r, cc)
        tot += 5.0 * (4.0 * m * m * n + 2.0 * m ** 3)
    return tot


def attn_flops_fwd_per_step() -> float:
    c = C2
    return 4.0 * c["batch"] * c["heads"] * c["seq"] ** 2 * c["head_dim"] * c["layers"]


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.idx = gpu_index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:  # noqa: BLE001
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def load_peaks() -> dict:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"bf16_tflops": d.get("bf16_tflops", 1590.0), "bf16_tflops_sustained": d.get("bf16_tflops_sustained", 1400.0),
                "hbm_gbs": d.get("hbm_gbs", 6650.0), "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0,
            "source": "fallback (B200_PROFILING.md)"}


# ------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port of the reference step on the host cores (bounded sample)
# ------------------------------------------------------------------------------------------------
def cpu_reference_step(sample_seqs: int = 4, ns_matrix_stride: int = 2) -> dict:
    """Times the CPU restatement of the reference step (oracle/reference_math.py, fp32) on the C2
    workload with a bounded sample: fwd+bwd on `sample_seqs` of the 16 sequences (extrapolated
    linearly in sequences) + the Muon update on every `ns_matrix_stride`-th transformer matrix and
    the embedding (extrapolated by Newton-Schulz flops)."""
    from oracle import reference_math as R
    # more than ~32 threads only adds oversubscription at these matrix sizes and blows the time bound
    cores = min(os.cpu_count() or 1, 32)
    torch.set_num_threads(cores)
    c = C2
    d = R.LlamaDims(c["hidden"], c["inter"], c["layers"], c["heads"], c["kv_heads"], c["head_dim"],
                    c["vocab_normal"] + 3)
    params = R.init_params(d, seed=42)
    batch = R.synthetic_batch(0, 0, sample_seqs, c["seq"], c["vocab_normal"])
    t0 = time.perf_counter()
    loss, ntoks, grads = R.loss_and_grads(params, batch, d, pad_token=c["vocab_normal"])
    t_fb = time.perf_counter() - t0
    names = [n for n, p in params.items() if p.dim() == 2]
    picked = [n for i, n in enumerate(names) if n == "embed_tokens.weight" or i % ns_matrix_stride == 0]

    def fl(shape):
        m, n = min(shape), max(shape)
        return 5.0 * (4.0 * m * m * n + 2.0 * m ** 3)
    t0 = time.perf_counter()
    opt = R.MuonOracle(c["lr"])
    opt.update({n: params[n] for n in picked}, {n: grads[n] for n in picked})
    t_opt_s = time.perf_counter() - t0
    frac = sum(fl(params[n].shape) for n in picked) / sum(fl(params[n].shape) for n in names)
    t_opt = t_opt_s / frac
    step_s = t_fb * (c["batch"] / sample_seqs) + t_opt
    return {"value": c["batch"] * c["seq"] / step_s, "unit": "tokens/s", "cores": cores, "kind": "port",
            "sample": (f"oracle fp32 CPU: fwd+bwd on {sample_seqs}/{c['batch']} sequences ({t_fb:.1f}s, scaled x"
                       f"{c['batch'] // sample_seqs}) + Muon/NS5 on {len(picked)}/{len(names)} matrices "
                       f"({t_opt_s:.1f}s, scaled by NS flops 1/{frac:.3f}); est. step {step_s:.1f}s; "
                       "MLX (the reference runtime) is not installable here"),
            "est_step_s": step_s, "loss": float(loss)}


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    last = None
    for _ in range(max(args.warmup, 0) and 1):
        cpu_reference_step(1, 12)
    for _ in range(max(min(args.steps, 3), 1)):
        last = cpu_reference_step(4, 2)
        vals.append(last["value"])
    v = statistics.median(vals)
    c = C2
    line = {"impl": "reference", "metric": "tokens/sec (device-timed) Llama Muon step", "value": v, "unit": "tokens/s",
            "n_gpus": args.gpus, "steps": len(vals), "warmup": args.warmup,
            "ms_per_step": 1e3 * c["batch"] * c["seq"] / v, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C2: Llama-80M (h1024 i2816 L12 H16/8 D64 V32003) Muon, batch 16 x seq 1024"},
            "cpu_baseline": {**{k: last[k] for k in ("unit", "cores", "kind", "sample")}, "value": v},
            "e2e": {"value": v, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", type=str, default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    args.warmup = max(args.warmup, 3)

    from mlx_cuda_distributed_pretraining_b200 import ops
    from mlx_cuda_distributed_pretraining_b200._lib import lib
    from mlx_cuda_distributed_pretraining_b200.core.training import Config, Trainer
    from mlx_cuda_distributed_pretraining_b200.distributed import dp

    rank, world, local_rank = dp.env_rank_world()
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    cfg = Config.from_dict(c2_config("bench-c2", distributed=world > 1))
    tr = Trainer(cfg, synthetic=True, quiet=True, run_root=str(ROOT / "gpurun_out" / "bench_runs"))
    tr._accum_step, tr._accum_tokens = 0, 0
    c = C2
    tokens_per_step = c["batch"] * c["seq"] * world
    K, W = args.steps, args.warmup
    dev = tr.device

    def sync_all():
        if world > 1:
            dp.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item())

    # ---------------- leg 1: device-resident inputs (the `value`) --------------------------------
    dev_batches = [tr.data_manager.generate_batch(s).to(dev) for s in range(K + W)]

    def step_resident(s: int):
        b = dev_batches[s]
        loss, ntoks = tr.compute_loss(tr.model, b[:, :-1], b[:, 1:])
        loss.backward()
        if tr.distributed:
            dp.all_reduce_sum_(tr.store.grad)
        tr.optimizer.update(tr.model)
        tr.store.zero_grad()
        return loss

    for s in range(W):
        step_resident(s)
    sync_all()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ops.TIMER = ops.KernelTimer()
    launches0 = lib().b200_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(W, W + K):
        loss = step_resident(s)
    e1.record()
    sync_all()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = int(lib().b200_launch_count() - launches0)
    kt = ops.TIMER.totals_ms()
    ops.TIMER = None
    clocks = sampler.stop() if rank == 0 else {}
    final_loss = float(loss.item())
    ms_per_step = ms_total / K
    value = tokens_per_step / (ms_per_step / 1e3)

    # ---------------- leg 2: end to end through the public API (pinned H2D + loss D2H) -----------
    host_batches = [tr.data_manager.generate_batch(s) for s in range(K + W)]
    tr._accum_step, tr._accum_tokens = 0, 0
    for s in range(2):
        l, _, _ = tr.train_step(W + K + s, host_batches[s])
        float(l.item())
    sync_all()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for s in range(K):
        l, _, _ = tr.train_step(W + K + 2 + s, host_batches[W + s])
        float(l.item())  # device -> host read of the step's loss
    f1.record()
    sync_all()
    e2e_ms = max_over_ranks(f0.elapsed_time(f1)) / K
    e2e_value = tokens_per_step / (e2e_ms / 1e3)
    h2d = host_batches[0].numel() * host_batches[0].element_size()

    if world > 1:
        dp.barrier()
    if rank != 0:
        dp.destroy()
        return

    peaks = load_peaks()
    traffic, traffic_src = None, None
    for tf_path in sorted(ROOT.glob("profiles/*_gemm_traffic.json")):   # written by tools/summarize_profiles.py
        try:
            traffic = float(json.loads(tf_path.read_text())["mean_dram_bytes_per_launch"])
            traffic_src = f"{tf_path.relative_to(ROOT)} (ncu --set full, mean DRAM read+write bytes per launch)"
        except Exception:
            pass
    ns_ms, ns_calls = kt.get("newton_schulz", (0.0, 0))
    ns_ms_step = ns_ms / K if K else 0.0
    # flops of the matrices THIS rank orthogonalised (all of them unless the owner-computes mode is on)
    sharded = bool(getattr(tr.optimizer, "shard_ns", False)) and world > 1
    ns_fl = 0.0
    for gi, g in enumerate(tr.store.mat_groups):
        m, n = min(g.rows, g.cols), max(g.rows, g.cols)
        owned = sum(hi - lo for lo, hi in tr.optimizer.owned_ranges_of(gi, world if sharded else 1, rank))
        ns_fl += owned * 5.0 * (4.0 * m * m * n + 2.0 * m ** 3)
    ns_tf = ns_fl / (ns_ms_step * 1e-3) / 1e12 if ns_ms_step > 0 else None
    peak_tf = peaks["bf16_tflops_sustained"]
    roofline = {"kernel": "gemm2_bf16_tc_kernel (Newton-Schulz chain, 15 batched tcgen05 cta_group::2 GEMMs x 5 shape groups)",
                "bound": "tensor", "achieved": ns_tf, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": (ns_tf / peak_tf) if ns_tf else None, "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": peaks["source"] + ", sustained bf16 (kernel timed inside a long step)",
                "frac_of_burst": (ns_tf / peaks["bf16_tflops"]) if ns_tf else None,
                "algorithmic_flops_per_step": ns_fl, "algorithmic_flops_all_ranks": ns_flops_per_step(),
                "ns_sharded_over_ranks": sharded, "ns_exchange": getattr(tr.optimizer, "exchange_mode", None),
                "ms_per_step": ns_ms_step,
                "share_of_step": ns_ms_step / ms_per_step if ms_per_step else None}
    kernels = {}
    afl = attn_flops_fwd_per_step()
    for nm, mult in (("attn_fwd", 1.0), ("attn_bwd", 2.5)):
        ms, calls = kt.get(nm, (0.0, 0))
        if ms > 0:
            kernels[nm] = {"ms_per_step": ms / K, "calls_per_step": calls / K,
                           "tflops_full_count": afl * mult / (ms / K * 1e-3) / 1e12,
                           "frac_of_tensor_peak": afl * mult / (ms / K * 1e-3) / 1e12 / peak_tf}

    line = {
        "metric": "tokens/sec (device-timed) Llama Muon step", "value": value, "unit": "tokens/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "C2: Llama-80M (h1024 i2816 L12 H16/8 D64 V32003) Muon NS5, batch 16 x seq 1024 per GPU",
                   "global_batch": c["batch"] * world, "seq_len": c["seq"],
                   "parallelism": f"dp{world}", "l2": "working set (params+grads+activations, GBs) >> 126 MB L2"},
        "e2e": {"value": e2e_value, "unit": "tokens/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 4},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "kernels": kernels,
        "final_loss": final_loss,
    }
    if world == 1 and not args.no_cpu_baseline:
        try:
            line["cpu_baseline"] = cpu_reference_step(4, 2)
        except Exception as e:  # noqa: BLE001
            line["cpu_baseline"] = {"error": repr(e)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dp.destroy()


if __name__ == "__main__":
    main()
