#!/bin/bash
# GPU-box recipe for the numbers committed under profiles/ (round 2): exactness campaign (GPU half, all batches, current build), GPU parity
# tests, bench lines of both arms, the literal-size configs[2] / configs[4] lines, ncu launch list + full capture of the heavy kernels.
mkdir -p gpurun_out; rm -rf gpurun_out/campaign
( time python tools/campaign_gpu.py --first 0 --count 154 --procs 8 --out gpurun_out/campaign ) > gpurun_out/campaign_gpu_final.log 2>&1; tail -2 gpurun_out/campaign_gpu_final.log
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_gpu_tests_final.log 2>&1; tail -3 gpurun_out/r2_gpu_tests_final.log
python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; cut -c1-300 gpurun_out/r2_bench.json
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_reference.json 2>> gpurun_out/r2_bench.err; cut -c1-200 gpurun_out/r2_bench_reference.json
python bench.py --shape sp1 --n 1048576 --steps 3 --no-cpu-baseline > gpurun_out/r2_bench_sp1_2p20.json 2>> gpurun_out/r2_bench.err; cut -c1-300 gpurun_out/r2_bench_sp1_2p20.json
python tools/pairing_bench.py --n 4194304 --steps 2 > gpurun_out/r2_pairing_bench_2p22.json 2>> gpurun_out/r2_bench.err; cut -c1-300 gpurun_out/r2_pairing_bench_2p22.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python - > gpurun_out/r2_vk_load_ms.txt 2>&1 <<'PY'
import time, stylus_zkvm_verifiers_b200 as Z
Z.VerificationKey.risc0().close()
ts = []
for _ in range(5):
    t0 = time.perf_counter(); k = Z.VerificationKey.risc0(); ts.append(1e3 * (time.perf_counter() - t0)); k.close()
print("zkv_vk_load_risc0 (tables, line tables, Miller(alpha, beta)) on one device, ms per call:", [round(t, 2) for t in ts])
PY
cat gpurun_out/r2_vk_load_ms.txt
python tools/campaign_oracle.py check --gpu gpurun_out/campaign --json gpurun_out/r2_campaign_check.json | cut -c1-300
bash tools/profile_r2.sh
