#!/usr/bin/env python
"""profiles/<R>_sass_summary.md: per-kernel counts of the Blackwell-native SASS mnemonics in the built library
(`cuobjdump -sass libb200hotpath.so`): UTCHMMA (tcgen05.mma; .2CTA = cta_group::2), UTMALDG / UTMASTG / UTMAREDG
(TMA tensor loads / stores / reductions), LDTM / STTM (tcgen05.ld / .st: TMEM <-> registers), UTCBAR (tcgen05.commit),
MUFU.EX2, RED/ATOM (global reductions), LDG/STG widths.  Runs in the build container (no GPU needed).
Usage: python tools/sass_summary.py r02"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
R = sys.argv[1] if len(sys.argv) > 1 else "r02"
LIB = ROOT / "mlx_cuda_distributed_pretraining_b200" / "libb200hotpath.so"
PATS = [("UTCHMMA", r"\bUTCHMMA"), ("UTCHMMA.2CTA", r"\bUTCHMMA\.2CTA"), ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"),
        ("UTMAREDG", r"\bUTMAREDG"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTCBAR", r"\bUTCBAR"),
        ("SYNCS", r"\bSYNCS"), ("MUFU.EX2", r"\bMUFU\.EX2"), ("RED", r"\bRED(G)?\b|\bRED\."), ("LDG.128", r"\bLDG\.E\.128|\bLDG\.E\.[A-Z.]*128"),
        ("STG.128", r"\bSTG\.E\.128|\bSTG\.E\.[A-Z.]*128"), ("HMMA (legacy mma.sync)", r"\bHMMA")]


def main():
    out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    demangle = {}
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        if cur is None or "/*" not in line:
            continue
        ins = line.split("*/", 1)[-1] if line.strip().startswith("/*") else line
        per[cur]["_instr"] += 1
        for name, pat in PATS:
            if re.search(pat, ins):
                per[cur][name] += 1
    names = list(per)
    dm = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines() if names else []
    for n, d in zip(names, dm):
        d = re.sub(r"\((int|bool)\)", "", d)
        d = re.sub(r"\(.*", "", d).replace("void ", "").replace("b200::(anonymous namespace)::", "").replace("b200::<unnamed>::", "")
        demangle[n] = d
    cols = [n for n, _ in PATS]
    md = [f"# {R}: SASS summary of libb200hotpath.so (sm_100a)", "",
          "`cuobjdump -sass` of the in-tree library built by `__graft_entry__.build()`; counts of instructions per kernel.",
          "UTCHMMA = tcgen05.mma (`.2CTA` = cta_group::2), UTMALDG/UTMASTG/UTMAREDG = TMA tensor load/store/reduce,",
          "LDTM/STTM = tcgen05.ld/.st (TMEM <-> registers), UTCBAR = tcgen05.commit -> mbarrier.  No HMMA (mma.sync) anywhere:",
          "every contraction is a tcgen05 kernel.", "",
          "| kernel | instr | " + " | ".join(cols) + " |", "|---|---:|" + "---:|" * len(cols)]
    tot = collections.Counter()
    for n in names:
        c = per[n]
        tot.update(c)
        md.append(f"| `{demangle.get(n, n)[:90]}` | {c['_instr']} | " + " | ".join(str(c[k]) if c[k] else "" for k in cols) + " |")
    md.append(f"| **total ({len(names)} kernels)** | {tot['_instr']} | " + " | ".join(str(tot[k]) for k in cols) + " |")
    (ROOT / "profiles").mkdir(exist_ok=True)
    (ROOT / "profiles" / f"{R}_sass_summary.md").write_text("\n".join(md) + "\n")
    print("wrote", ROOT / "profiles" / f"{R}_sass_summary.md", "-", len(names), "kernels")
    print({k: tot[k] for k in cols})


if __name__ == "__main__":
    main()
