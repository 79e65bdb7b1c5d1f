#!/bin/bash
# ncu capture of the two-lane layout microbenchmark kernel (sqr+2nline launch)
M="sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed_pipe_alu.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed.sum,sm__cycles_active.avg"
ncu --set full --metrics $M --import-source on --clock-control none -k regex:k_lz2 -s 4 -c 1 -f -o gpurun_out/lz2_prof tools/build/lzbench 16 > gpurun_out/ncu_lz2.log 2>&1
tail -3 gpurun_out/ncu_lz2.log
