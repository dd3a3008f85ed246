#!/bin/bash
# ncu --set full captures of the level-1 analysis, the level-1 histogram and the final synthesis launch of one
# 128-plane step (run under gpurun; each capture after a plain run of the same command)
CMD="python bench.py --steps 2 --warmup 1 --no-extra --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:analysis_tma_kernel -s 3 -c 1 -o gpurun_out/r2b_analysis_L1 -f $CMD > gpurun_out/ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hist_kernel -s 24 -c 1 -o gpurun_out/r2b_hist_L1 -f $CMD > gpurun_out/ncu_h.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:synth_kernel -s 31 -c 1 -o gpurun_out/r2b_final_synth -f $CMD > gpurun_out/ncu_s.log 2>&1
ls -la gpurun_out/*.ncu-rep
