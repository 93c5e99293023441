#!/usr/bin/env python
"""Turns the scratch outputs of tools/profile_gpu.sh (gpurun_out/<R>_*) into tracked evidence under profiles/:

  <R>_bench_launches.md   launch list of one C2 bench step (every kernel, share of summed kernel time)
  <R>_ncu_kernels.md      ncu --set full of EVERY hand-written kernel launch of one step, aggregated per kernel:
                          duration, tensor-pipe activity, DRAM throughput vs the measured HBM peak, DRAM bytes,
                          achieved occupancy -- tensor-bound and HBM-bound kernels alike
  <R>_gemm_traffic.json   mean DRAM read+write bytes per launch of the dominant kernel (bench.py roofline.traffic)
Usage: python tools/summarize_profiles.py r02"""
import collections
import csv
import json
import re
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
R = sys.argv[1] if len(sys.argv) > 1 else "r02"
OUT = ROOT / "profiles"
GP = ROOT / "gpurun_out"
OUT.mkdir(exist_ok=True)
PEAKS = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
HBM = float(PEAKS.get("hbm_gbs", 6575.1))

MINE = ("gemm_bf16_tc", "gemm2_", "splitk_", "delta_kernel", "attn_", "rmsnorm", "rope_kernel", "muon_", "axpy", "adamw",
        "clip_accum", "sumsq", "ns_scales", "sgd_momentum", "f32_to_bf16", "ema_split", "graft_", "split_bf16", "split4",
        "root_", "glu_", "ce_", "adam_direction", "embedding_")


def short(name: str) -> str:
    name = re.sub(r"\((int|bool)\)", "", name)
    name = re.sub(r"\(.*", "", name)
    return re.sub(r"void |b200::<unnamed>::|b200::\(anonymous namespace\)::|at::native::|native::|<unnamed>::", "", name)[:100]


def launch_list():
    p = GP / f"{R}_bench_launches.csv"
    if not p.exists():
        return
    lines = [l for l in p.read_text().splitlines(True) if not l.startswith("==")]
    rows = [(r["Kernel Name"], float(r["Metric Value"])) for r in csv.DictReader(lines) if r.get("Metric Value")]
    # step boundaries: ce_fwd_kernel runs exactly once per step; the span between the 3rd-last and the last
    # occurrence is two whole steps (a cyclic shift of the step does not change per-kernel totals)
    marks = [i for i, (k, _) in enumerate(rows) if "ce_fwd_kernel" in k]
    if len(marks) >= 3:
        last = rows[marks[-3]:marks[-1]]
    else:
        per = len(rows) // 5
        last = rows[-2 * per:]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for k, v in last:
        key = short(k)
        agg[key][0] += 1
        agg[key][1] += v
    tot = sum(v[1] for v in agg.values())
    mine = sum(v[1] for k, v in agg.items() if any(s in k for s in MINE) and "vectorized" not in k)
    md = [f"# {R}: kernel launch list of one bench step (C2, 1xB200)", "",
          "`ncu --metrics gpu__time_duration.sum --clock-control none` around `python bench.py --steps 2 --warmup 3 --no-e2e`;",
          "last two steps averaged. Times under ncu are serialised and cold-cache: compare SHARES, not absolutes.", "",
          f"- kernels per step: {len(last) // 2}; summed kernel time per step: {tot / 2 / 1e6:.2f} ms",
          f"- hand-written sm_100a kernels: {mine / tot * 100:.1f}% of summed kernel time, "
          f"{sum(v[0] for k, v in agg.items() if any(s in k for s in MINE) and 'vectorized' not in k) // 2} launches per step", "",
          "| share | ms/step | launches/step | kernel |", "|---:|---:|---:|---|"]
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        md.append(f"| {t / tot * 100:.1f}% | {t / 2 / 1e6:.3f} | {c // 2} | `{k}` |")
    (OUT / f"{R}_bench_launches.md").write_text("\n".join(md) + "\n")
    print("wrote", OUT / f"{R}_bench_launches.md")


UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3,
        "second": 1.0, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}


def load_raw(path: Path):
    rows = list(csv.reader([l for l in path.read_text().splitlines() if not l.startswith("==")]))
    if len(rows) < 3:
        return [], [], []
    return rows[0], rows[1], rows[2:]


def kernel_tables():
    md = [f"# {R}: ncu --set full of the hand-written kernels", "",
          "`ncu --set full --clock-control none` (tools/profile_gpu.sh): every launch of a hand-written kernel in ONE C2 bench step",
          "(plus AdamW from a C5 step and the Shampoo elementwise kernels from a C4 step), aggregated per kernel name.",
          f"`hbm frac` = (dram read + write bytes) / duration / {HBM:.0f} GB/s (MEASURED_PEAKS.json); `tensor` =",
          "`sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active`; per-launch times under ncu are cold-cache and",
          "clock-unlocked (compare ratios, not absolutes).", ""]
    traffic = None
    for tag, title in (("step_full", "one C2 step"), ("adamw_full", "AdamW (C5, 1B parameters, one flat launch)"),
                       ("shampoo_full", "Shampoo elementwise kernels (C4)"),
                       ("attn_bwd128_full", "attention backward, head dim 128 (one C5 layer call)")):
        p = GP / f"{R}_{tag}.csv"
        if not p.exists():
            continue
        hdr, units, rows = load_raw(p)
        if not hdr:
            continue

        def col(r, name):
            if name not in hdr:
                return None
            i = hdr.index(name)
            try:
                return float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0)
            except ValueError:
                return None
        agg = collections.OrderedDict()
        for r in rows:
            k = short(r[hdr.index("Kernel Name")])
            a = agg.setdefault(k, collections.defaultdict(float))
            a["n"] += 1
            for key, name in (("t", "gpu__time_duration.sum"), ("rd", "dram__bytes_read.sum"), ("wr", "dram__bytes_write.sum"),
                              ("tensor", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                              ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                              ("occ", "sm__warps_active.avg.pct_of_peak_sustained_active"),
                              ("l2hit", "lts__t_sector_hit_rate.pct"), ("regs", "launch__registers_per_thread"),
                              ("smthru", "sm__throughput.avg.pct_of_peak_sustained_elapsed")):
                v = col(r, name)
                if v is not None:
                    a[key] += v
        md += [f"## {title}", "", "| kernel | launches | us/launch | DRAM MB/launch | achieved GB/s | hbm frac | ncu dram % | tensor pipe % | sm thru % | warps active % | L2 hit % | regs |",
               "|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|"]
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["t"]):
            n = a["n"]
            t = a["t"] / n
            by = (a["rd"] + a["wr"]) / n
            gbs = by / t / 1e9 if t > 0 else 0.0
            md.append(f"| `{k}` | {int(n)} | {t * 1e6:.1f} | {by / 1e6:.1f} | {gbs:.0f} | {gbs / HBM:.2f} | {a['dram_pct'] / n:.1f} | "
                      f"{a['tensor'] / n:.1f} | {a['smthru'] / n:.1f} | {a['occ'] / n:.1f} | {a['l2hit'] / n:.1f} | {a['regs'] / n:.0f} |")
            if tag == "step_full" and "gemm2_grouped" in k:
                traffic = {"kernel": k, "launches": int(n), "mean_dram_bytes_per_launch": by,
                           "mean_us_per_launch": t * 1e6,
                           "note": "ncu --set full, the 15 grouped Newton-Schulz launches of one C2 bench step "
                                   "(algorithmic bytes of a stage: every operand and output of all 85 matrices once)"}
        md.append("")
    (OUT / f"{R}_ncu_kernels.md").write_text("\n".join(md) + "\n")
    print("wrote", OUT / f"{R}_ncu_kernels.md")
    if traffic:
        (OUT / f"{R}_gemm_traffic.json").write_text(json.dumps(traffic, indent=1))
        print("wrote", OUT / f"{R}_gemm_traffic.json")


launch_list()
kernel_tables()
