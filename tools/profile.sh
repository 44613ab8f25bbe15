#!/bin/bash
# Run under gpurun (1 GPU): launch list of a short bench run + one full ncu capture of the GEMM and
# weight-side kernels.  Outputs land in gpurun_out/; summaries are copied to profiles/ by hand.
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-graph --skip-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
REPS=2 ncu --set full --clock-control none --import-source on -k regex:"gemm_.*sm100_kernel|merge_lokr|grad_lokr" -c 20 \
    -o gpurun_out/prof_r01 python tools/ncu_target.py > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/
