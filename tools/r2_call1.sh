#!/bin/bash
# GPU-box recipe: validate the tree (GPU parity tests), bench line, instruction-rate microbenchmarks, first slice of the exactness campaign.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
nproc
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_gpu_tests.log 2>&1; tail -3 gpurun_out/r2_gpu_tests.log
python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -c 600 gpurun_out/r2_bench.json
tools/build/microbench > gpurun_out/r2_microbench.json 2>&1
tools/build/latbench > gpurun_out/r2_latbench.txt 2>&1
tools/build/fp2bench > gpurun_out/r2_fp2bench.txt 2>&1
( time python tools/campaign_gpu.py --first 0 --count 40 --procs 8 --out gpurun_out/campaign ) > gpurun_out/campaign_gpu.log 2>&1; tail -3 gpurun_out/campaign_gpu.log
