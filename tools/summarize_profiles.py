#!/usr/bin/env python
"""Turns gpurun_out/<R>_bench_launches.csv and <R>_prof_*.ncu-rep into profiles/<R>_*.md
(the tracked evidence).  Usage: python tools/summarize_profiles.py r01"""
import collections
import csv
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
R = sys.argv[1] if len(sys.argv) > 1 else "r01"
OUT = ROOT / "profiles"
GP = ROOT / "gpurun_out"
OUT.mkdir(exist_ok=True)

KEYS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__cycles_elapsed.avg",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct"]
MINE = ("gemm_bf16_tc", "gemm2_bf16_tc", "splitk_", "delta_kernel", "attn_", "rmsnorm", "rope_kernel", "muon_", "axpy", "adamw", "clip_accum", "sumsq",
        "ns_scales", "sgd_momentum", "f32_to_bf16", "ema_split", "graft_update", "split_bf16", "glu_", "ce_")


def launch_list():
    p = GP / f"{R}_bench_launches.csv"
    if not p.exists():
        return
    lines = [l for l in p.read_text().splitlines(True) if not l.startswith("==")]
    rows = [(r["Kernel Name"], float(r["Metric Value"])) for r in csv.DictReader(lines) if r.get("Metric Value")]
    per = len(rows) // 5          # bench ran 3 warm-up + 2 timed steps
    last = rows[-2 * per:]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for k, v in last:
        key = re.sub(r"\(.*", "", k)
        key = re.sub(r"void |b200::<unnamed>::|at::native::|native::|<unnamed>::", "", key)[:110]
        agg[key][0] += 1
        agg[key][1] += v
    tot = sum(v[1] for v in agg.values())
    mine = sum(v[1] for k, v in agg.items() if any(s in k for s in MINE) and "vectorized" not in k)
    md = [f"# {R}: kernel launch list of one bench step (C2, 1xB200)", "",
          "`ncu --metrics gpu__time_duration.sum --clock-control none` around `python bench.py --steps 2 --warmup 3`;",
          "last two steps averaged. Times under ncu are serialised and cold-cache: compare SHARES, not absolutes.", "",
          f"- kernels per step: {len(last) // 2}; summed kernel time per step: {tot / 2 / 1e6:.2f} ms",
          f"- hand-written sm_100a kernels: {mine / tot * 100:.1f}% of summed kernel time", "",
          "| share | ms/step | launches/step | kernel |", "|---:|---:|---:|---|"]
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
        md.append(f"| {t / tot * 100:.1f}% | {t / 2 / 1e6:.3f} | {c // 2} | `{k}` |")
    (OUT / f"{R}_bench_launches.md").write_text("\n".join(md) + "\n")
    print("wrote", OUT / f"{R}_bench_launches.md")


def full_captures():
    md = [f"# {R}: ncu --set full captures of the hand-written kernels", "",
          "`ncu --set full --clock-control none --import-source on` (tools/profile_gpu.sh); values per launch.", ""]
    for rep in sorted(GP.glob(f"{R}_prof_*.ncu-rep")):
        raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        keys = [k for k in KEYS if k in hdr]
        md += [f"## {rep.stem}", "", "| kernel | " + " | ".join(keys) + " |", "|---|" + "---:|" * len(keys)]
        for r in rows[2:]:
            name = re.sub(r"\(.*", "", r[hdr.index("Kernel Name")]).replace("void ", "").replace("b200::<unnamed>::", "")
            md.append(f"| `{name}` | " + " | ".join(f"{r[hdr.index(k)]} {units[hdr.index(k)]}" for k in keys) + " |")
        md.append("")
        if rep.stem.endswith("prof_gemm") and "dram__bytes_read.sum" in hdr:
            # DRAM traffic per launch of the dominant kernel -> bench.py's roofline.traffic
            def to_bytes(v, u):
                return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            per = [to_bytes(r[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_read.sum")]) +
                   to_bytes(r[hdr.index("dram__bytes_write.sum")], units[hdr.index("dram__bytes_write.sum")])
                   for r in rows[2:]]
            import json
            (OUT / f"{R}_gemm_traffic.json").write_text(json.dumps({
                "kernel": "gemm2_bf16_tc_kernel", "launches": len(per), "dram_bytes_per_launch": per,
                "mean_dram_bytes_per_launch": sum(per) / len(per),
                "note": "ncu --set full, launches 60..63 of tests/gpu_first_light.py ns_perf (batch-24 1024x1024 "
                        "group: G1 reads X = 50.3 MB algorithmic; G3 reads B + X and writes X' = 151 MB)"}, indent=1))
    (OUT / f"{R}_ncu_full.md").write_text("\n".join(md) + "\n")
    print("wrote", OUT / f"{R}_ncu_full.md")


launch_list()
full_captures()
