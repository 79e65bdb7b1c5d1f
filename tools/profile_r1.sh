#!/bin/bash
# Profiling recipe of profiles/r1_*: run on the GPU box (gpurun -- 'bash tools/profile_r1.sh').  Every ncu pass follows a plain run of the
# same command that exited 0; numbers printed under ncu are never bench values.
cd /root/repo; mkdir -p gpurun_out; T=/tmp/zkvprof; mkdir -p $T
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err || exit 1
# 1. launch list of the bench command (durations only)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1_launches.csv $B > gpurun_out/ncu_launches.log 2>&1
# 2. full sections of the one-kernel forms (serial chain) and of one chunk's segment / stage kernels (default path)
S="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$S --chunks 1 > /dev/null 2>&1 || exit 1
ncu --set full --import-source on --clock-control none -k 'regex:^k_(vkx|g2_check|miller_norm|final_exp)$' -c 6 -o $T/prof_r1_serial -f $S --chunks 1 > gpurun_out/ncu_serial.log 2>&1
ncu --set full --import-source on --clock-control none -k 'regex:^k_(miller_norm_seg|final_exp_stage)$' -c 12 -o $T/prof_r1_chunked -f $S > gpurun_out/ncu_chunked.log 2>&1
python tools/ncu_summary.py gpurun_out/r1_ncu_summary.json $T/prof_r1_serial.ncu-rep $T/prof_r1_chunked.ncu-rep > gpurun_out/ncu_summary.log 2>&1
tail -20 gpurun_out/ncu_summary.log
