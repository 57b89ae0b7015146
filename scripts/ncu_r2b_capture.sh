#!/bin/bash
# GPU box, end of round 2: launch lists of the render frame and of a Stage-II frame, `ncu --set full` of the new
# sample_pdf kernel and of the Stage-II conv launches of RefineNetwork.layer7 + layer8 (B200_PROFILING.md recipe; every
# command has exited 0 without ncu before).
CMD="python bench.py --steps 1 --warmup 3 --no-train --no-cpu-baseline --no-torch-gpu --no-stage2"
S2="python scripts/gpu_spade_profile.py --no-baselines"
$CMD > gpurun_out/r2b_plain.log 2>&1 && $S2 > gpurun_out/r2b_plain_s2.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r2b_bench_launches.csv $CMD > gpurun_out/r2b_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"sample_pdf" -c 1 -o gpurun_out/r2b_sample_pdf $CMD > gpurun_out/r2b_ncu2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 115 --csv --log-file gpurun_out/r2b_stage2_launches.csv $S2 > gpurun_out/r2b_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"spade_conv" -s 59 -c 11 -o gpurun_out/r2b_spade $S2 > gpurun_out/r2b_ncu4.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
