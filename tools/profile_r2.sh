#!/bin/bash
# GPU-box recipe (round 2): plain bench run, the ncu launch list of the same command, one --set full capture of the two heavy kernels
# in their one-launch forms (serial chain: --chunks 1) with the fmaheavy / alu pipe counters VERDICT round 1 asked for.
mkdir -p gpurun_out
M="sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed_pipe_fmalite.sum,sm__inst_executed_pipe_alu.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed.sum,sm__cycles_active.avg"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/r2_profile_plain.json 2> gpurun_out/r2_profile_plain.err || { echo "plain run failed"; tail -5 gpurun_out/r2_profile_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_launches.log 2>&1
CMD1="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --chunks 1"
$CMD1 > gpurun_out/r2_profile_serial.json 2> gpurun_out/r2_profile_serial.err || { echo "serial run failed"; exit 1; }
ncu --set full --metrics $M --import-source on --clock-control none -k regex:"k_miller_lz|k_final_exp_lz" -s 8 -c 2 -f -o gpurun_out/r2_prof_serial $CMD1 > gpurun_out/r2_ncu_serial.log 2>&1
tail -n 3 gpurun_out/r2_ncu_serial.log
ls -la gpurun_out | head -30
