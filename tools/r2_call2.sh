#!/bin/bash
# GPU-box recipe: rest of the exactness campaign's GPU half (batches 40..153), cleaned microbenchmarks, ncu capture of the heavy kernels.
mkdir -p gpurun_out
( time python tools/campaign_gpu.py --first 40 --count 114 --procs 8 --out gpurun_out/campaign ) > gpurun_out/campaign_gpu2.log 2>&1; tail -3 gpurun_out/campaign_gpu2.log
tools/build/microbench > gpurun_out/r2_microbench.json 2>&1; cat gpurun_out/r2_microbench.json
bash tools/profile_r2.sh
