#!/bin/bash
# Run under gpurun on ONE GPU: launch list of the bench step + full ncu captures of the top kernels.
# Outputs land in gpurun_out/ (scratch); tools/summarize_profiles.py turns them into profiles/*.md.
set -u
R=${1:-r01}
MODE=${2:-all}   # "launches" = launch list only
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv \
    --log-file gpurun_out/${R}_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline \
    > gpurun_out/${R}_bench_under_ncu.log 2>&1
echo "launch list rc=$?"
[ "$MODE" = launches ] && exit 0
ncu --set full --clock-control none --import-source on -k regex:gemm2_bf16_tc -s 60 -c 4 -f \
    -o gpurun_out/${R}_prof_gemm python tests/gpu_first_light.py ns_perf > gpurun_out/${R}_prof_gemm.log 2>&1
echo "gemm capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:attn_fwd_kernel -s 3 -c 2 -f \
    -o gpurun_out/${R}_prof_attn_fwd python tests/gpu_first_light.py attn_perf > gpurun_out/${R}_prof_attn_fwd.log 2>&1
echo "attn fwd capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:attn_bwd64_kernel -s 3 -c 2 -f \
    -o gpurun_out/${R}_prof_attn_bwd python tests/gpu_first_light.py attn_perf > gpurun_out/${R}_prof_attn_bwd.log 2>&1
echo "attn bwd capture rc=$?"
ls -la gpurun_out | tail -20
