#!/bin/bash
# Run under gpurun (1 GPU).  (1) ncu launch list of ONE full eager fwd+bwd step of the bench workload (NVTX range
# "lyco_step"), (2) `--set full` captures of the engine kernels on the dominant shapes (tools/ncu_target.py: GEMMs,
# merge / factor-grad / structured-gradient kernels, convolutions; tools/ncu_layout_target.py: layout kernels).
# The summaries are produced ON THE BOX: gpurun only copies gpurun_out/ back when it is under 64 MiB, so oversized
# .ncu-rep files are dropped at the end (the text summaries and the traffic record survive).
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "lyco_step" --csv \
    --log-file gpurun_out/launches_step.csv \
    python bench.py --steps 1 --warmup 1 --no-graph --skip-cpu-baseline --skip-gpu-reference --nvtx-step > gpurun_out/bench_under_ncu.log 2>&1
python tools/ncu_launch_list_summary.py gpurun_out/launches_step.csv > gpurun_out/launch_list_step_summary.txt 2>&1
gzip -f gpurun_out/launches_step.csv
REPS=1 ncu --set full --clock-control none --import-source on \
    -k regex:"gemm_sm100_kernel|conv_sm100_kernel|hada_sm100_kernel|merge_lokr|grad_lokr|lokr_mix|lokr_w1grad" -c 40 \
    -o gpurun_out/prof_r02 python tools/ncu_target.py > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/prof_r02.ncu-rep --page raw --csv > gpurun_out/prof_r02_raw.csv 2>/dev/null
python tools/ncu_summarize.py < gpurun_out/prof_r02_raw.csv > gpurun_out/ncu_full_summary.txt 2>&1
python tools/ncu_traffic_json.py gpurun_out/ncu_traffic_cfg4.json cfg4 gemm_sm100_kernel 0 < gpurun_out/prof_r02_raw.csv > gpurun_out/ncu_traffic.log 2>&1
REPS=1 ncu --set full --clock-control none \
    -k regex:"transpose_cast|filter_|dora_|delta_weight" -c 10 \
    -o gpurun_out/prof_layout python tools/ncu_layout_target.py > gpurun_out/ncu_layout.log 2>&1
ncu -i gpurun_out/prof_layout.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summarize.py > gpurun_out/ncu_layout_summary.txt 2>&1
gzip -f gpurun_out/prof_r02_raw.csv
find gpurun_out -name '*.ncu-rep' -size +20M -print -delete
du -sh gpurun_out
ls -la gpurun_out/
