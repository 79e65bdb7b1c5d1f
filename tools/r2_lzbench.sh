#!/bin/bash
# GPU-box recipe for the layout microbenchmark (tools/lzbench.cu): plain run, then one ncu capture of each layout's sqr+2nline launch.
set -x
mkdir -p gpurun_out
M="sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed_pipe_fmalite.sum,sm__inst_executed_pipe_alu.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed.sum,sm__cycles_active.avg"
tools/build/lzbench 64 > gpurun_out/lzbench.json 2> gpurun_out/lzbench.err || { echo "lzbench failed"; tail -5 gpurun_out/lzbench.err; cat gpurun_out/lzbench.json; exit 1; }
cat gpurun_out/lzbench.json
ncu --set full --metrics $M --import-source on --clock-control none -k regex:k_lz -s 6 -c 1 -f -o gpurun_out/lz_smem tools/build/lzbench 16 > gpurun_out/ncu_lz.log 2>&1
ncu --set full --metrics $M --import-source on --clock-control none -k regex:k_old -s 3 -c 1 -f -o gpurun_out/lz_old tools/build/lzbench 16 > gpurun_out/ncu_old.log 2>&1
tail -3 gpurun_out/ncu_lz.log gpurun_out/ncu_old.log
ls -la gpurun_out
