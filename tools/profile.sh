#!/bin/bash
# Run under gpurun (1 GPU).  (1) ncu launch list of ONE full eager fwd+bwd step of the bench workload
# (NVTX range "lyco_step"), (2) `--set full` captures of the engine kernels on the dominant shapes
# (tools/ncu_target.py: GEMMs, merge / factor-grad kernels, convolutions; tools/ncu_layout_target.py: layout kernels).
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "lyco_step" --csv \
    --log-file gpurun_out/launches_step.csv \
    python bench.py --steps 1 --warmup 1 --no-graph --skip-cpu-baseline --nvtx-step > gpurun_out/bench_under_ncu.log 2>&1
REPS=1 ncu --set full --clock-control none --import-source on \
    -k regex:"gemm_sm100_kernel|conv_sm100_kernel|merge_lokr|grad_lokr|lokr_mix|lokr_w1grad" -c 40 \
    -o gpurun_out/prof_r02 python tools/ncu_target.py > gpurun_out/ncu_full.log 2>&1
REPS=1 ncu --set full --clock-control none --import-source on \
    -k regex:"transpose_cast|filter_|conv_sm100" -c 8 \
    -o gpurun_out/prof_layout python tools/ncu_layout_target.py > gpurun_out/ncu_layout.log 2>&1
ls -la gpurun_out/
