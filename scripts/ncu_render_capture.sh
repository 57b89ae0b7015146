#!/bin/bash
# GPU box: launch list + one `ncu --set full` capture of a rendered frame's kernels (B200_PROFILING.md recipe).
CMD="python bench.py --steps 1 --warmup 3 --no-train --no-cpu-baseline --no-torch-gpu"
$CMD > gpurun_out/r2_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches.csv $CMD > gpurun_out/r2_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"field_fwd|composite_fwd|sample_pdf" -s 15 -c 5 -o gpurun_out/r2_render $CMD > gpurun_out/r2_ncu_full.log 2>&1
tail -2 gpurun_out/r2_ncu_full.log
