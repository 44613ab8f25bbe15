#!/bin/bash
# Run under gpurun (1 GPU).  (1) ncu launch list of ONE full eager fwd+bwd step of the bench workload
# (NVTX range "lyco_step"), (2) one `--set full` capture of the engine kernels on the dominant shapes.
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "lyco_step" --csv \
    --log-file gpurun_out/launches_step.csv \
    python bench.py --steps 1 --warmup 1 --no-graph --skip-cpu-baseline --nvtx-step > gpurun_out/bench_under_ncu.log 2>&1
REPS=2 ncu --set full --clock-control none --import-source on \
    -k regex:"gemm_sm100_kernel|conv_sm100_kernel|merge_lokr|grad_lokr" -c 24 \
    -o gpurun_out/prof_r01 python tools/ncu_target.py > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/
