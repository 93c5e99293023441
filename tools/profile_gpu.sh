#!/bin/bash
# Run under gpurun on ONE GPU: evidence for profiles/ (tools/summarize_profiles.py turns it into .md).
#   1. launch list of the bench step (every kernel, gpu__time_duration)
#   2. ncu --set full of EVERY hand-written kernel launch of one C2 bench step (tensor-core and HBM-bound alike)
#   3. ncu --set full of the optimizer kernels the C2 step does not run: AdamW (C5), Shampoo elementwise (C4)
# Only CSV exports leave the box (the .ncu-rep of 150 kernels would exceed the 64 MiB pull limit); the two
# attention kernels are additionally captured with source pages into small .ncu-rep files.
set -u
R=${1:-r02}
MODE=${2:-all}   # launches | step | extra | attn | all
mkdir -p gpurun_out
B="python bench.py --warmup 3 --no-cpu-baseline --no-configs-block --no-e2e"
MINE='regex:gemm2_|gemm_bf16_tc|splitk_|attn_|f32_to_bf16|muon_momentum|ns_scales|axpy_update|sgd_momentum|rmsnorm|glu_|ce_fwd|ce_bwd|clip_accum|embedding_'
if [ "$MODE" = launches ] || [ "$MODE" = all ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv \
      --log-file gpurun_out/${R}_bench_launches.csv $B --steps 2 > gpurun_out/${R}_bench_under_ncu.log 2>&1
  echo "launch list rc=$?"
fi
if [ "$MODE" = step ] || [ "$MODE" = all ]; then
  # matched launches per C2 step: count them from the launch list when present, else assume 149
  PER=149
  if [ -f gpurun_out/${R}_bench_launches.csv ]; then
    PER=$(python - <<PY
import csv,re
rows=[r for r in csv.DictReader(l for l in open("gpurun_out/${R}_bench_launches.csv") if not l.startswith("=="))]
pat=re.compile("${MINE#regex:}")
n=sum(1 for r in rows if pat.search(r["Kernel Name"]))
print(n//5)
PY
)
  fi
  echo "matched launches per step: $PER"
  ncu --set full --clock-control none -k "$MINE" -s $((PER*3)) -c $PER -f -o /tmp/${R}_step_full \
      $B --steps 1 > gpurun_out/${R}_step_full.log 2>&1
  echo "step capture rc=$?"
  ncu -i /tmp/${R}_step_full.ncu-rep --page raw --csv > gpurun_out/${R}_step_full.csv 2>/dev/null
  wc -l gpurun_out/${R}_step_full.csv
fi
if [ "$MODE" = extra ] || [ "$MODE" = all ]; then
  ncu --set full --clock-control none -k regex:adamw_kernel -s 3 -c 1 -f -o /tmp/${R}_adamw \
      python bench.py --config c5 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${R}_adamw.log 2>&1
  ncu -i /tmp/${R}_adamw.ncu-rep --page raw --csv > gpurun_out/${R}_adamw_full.csv 2>/dev/null
  ncu --set full --clock-control none -k 'regex:ema_split|adam_direction|graft_update|sumsq_kernel|split4|root_init|split_bf16' \
      -s 60 -c 24 -f -o /tmp/${R}_shampoo python bench.py --config c4 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e \
      > gpurun_out/${R}_shampoo.log 2>&1
  ncu -i /tmp/${R}_shampoo.ncu-rep --page raw --csv > gpurun_out/${R}_shampoo_full.csv 2>/dev/null
  echo "extra captures done"
fi
if [ "$MODE" = attn ] || [ "$MODE" = all ]; then
  ncu --set full --clock-control none --import-source on -k regex:attn_fwd_kernel -s 36 -c 1 -f \
      -o gpurun_out/${R}_prof_attn_fwd $B --steps 1 > gpurun_out/${R}_prof_attn_fwd.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:attn_bwd64_kernel -s 36 -c 1 -f \
      -o gpurun_out/${R}_prof_attn_bwd $B --steps 1 > gpurun_out/${R}_prof_attn_bwd.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:attn_bwd128_kernel -s 48 -c 1 -f \
      -o gpurun_out/${R}_prof_attn_bwd128 python bench.py --config c5 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e \
      > gpurun_out/${R}_prof_attn_bwd128.log 2>&1
  ncu -i gpurun_out/${R}_prof_attn_bwd128.ncu-rep --page raw --csv > gpurun_out/${R}_attn_bwd128_full.csv 2>/dev/null
  echo "attention source captures done"
fi
ls -la gpurun_out | tail -12
