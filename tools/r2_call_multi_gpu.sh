#!/bin/bash
# 8-GPU box recipe: one handle over all devices (GPU test + the literal configs[2] / configs[4] strong-scaling lines, one process)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | sort | uniq -c
( time python -m pytest tests -m gpu -x -q -k "one_handle" ) > gpurun_out/r2_multi_test.log 2>&1; tail -4 gpurun_out/r2_multi_test.log
( time python tools/single_handle_bench.py --config sp1 --sweep 1,2,4,8 --steps 3 --warmup 2 ) > gpurun_out/r2_single_handle_sp1.jsonl 2> gpurun_out/r2_single_handle_sp1.err; cat gpurun_out/r2_single_handle_sp1.jsonl | cut -c1-220; tail -4 gpurun_out/r2_single_handle_sp1.err
( time python tools/single_handle_bench.py --config pairing --sweep 1,2,4,8 --steps 2 --warmup 1 ) > gpurun_out/r2_single_handle_pairing.jsonl 2> gpurun_out/r2_single_handle_pairing.err; cat gpurun_out/r2_single_handle_pairing.jsonl | cut -c1-220; tail -4 gpurun_out/r2_single_handle_pairing.err
