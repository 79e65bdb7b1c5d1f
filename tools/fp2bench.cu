// Isolated throughput of the generated Fp2 primitives at low occupancy (the occupancy the pairing kernels run at).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/fp2bench.bin tools/fp2bench.cu
// Prints cycles per call per warp for 1, 2, 3, 4 warps per scheduler: the IMAD.WIDE pipe bound is 4 cycles x IMAD.WIDE count x warps/scheduler.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../stylus_zkvm_verifiers_b200/csrc/fp_ptx.cuh"

template <int MODE, int MINB>
__global__ void __launch_bounds__(128, MINB) k_chain(uint32_t* out, int iters, long long* cyc, const uint32_t* __restrict__ seed) {
    uint32_t a0[8], a1[8], b0[8], b1[8];
    for (int i = 0; i < 8; i++) { a0[i] = threadIdx.x * 7 + i; a1[i] = blockIdx.x + 3 * i; b0[i] = seed[i] * (i + 1); b1[i] = seed[8 + i] ^ (i << 8); }   // run-time operands: nothing folds into immediates
    a0[7] &= 0x0fffffff; a1[7] &= 0x0fffffff; b0[7] &= 0x0fffffff; b1[7] &= 0x0fffffff;
    long long t0 = clock64();
#pragma unroll 1
    for (int k = 0; k < iters; k++) {
        if (MODE == 0) { uint32_t r0[8], r1[8]; fp2_mul_ptx(r0, r1, a0, a1, b0, b1); for (int i = 0; i < 8; i++) { a0[i] = r0[i]; a1[i] = r1[i]; } }
        if (MODE == 1) { uint32_t r0[8], r1[8]; fp2_sqr_ptx(r0, r1, a0, a1); for (int i = 0; i < 8; i++) { a0[i] = r0[i]; a1[i] = r1[i]; } }
        if (MODE == 2) { uint32_t r0[8]; fp_mul_ptx(r0, a0, b0); for (int i = 0; i < 8; i++) a0[i] = r0[i]; }
    }
    long long t1 = clock64();
    uint32_t x = 0; for (int i = 0; i < 8; i++) x ^= a0[i] ^ a1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    uint32_t* d; cudaMalloc(&d, 1 << 24); cudaMemset(d, 0x5a, 1 << 24); long long* dc; cudaMalloc(&dc, 8);
    const int iters = 2000;
    const char* names[3] = {"fp2_mul", "fp2_sqr", "fp_mul"};
    for (int mode = 0; mode < 3; mode++)
        for (int wps = 1; wps <= 2; wps++) {      // warps per scheduler = resident blocks of 128 threads per SM (register budget 255/255/168/128)
            int blocks = p.multiProcessorCount * wps;
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            for (int rep = 0; rep < 2; rep++) {
                cudaEventRecord(e0);
#define L(M, B) k_chain<M, B><<<blocks, 128>>>(d, iters, dc, d + (1 << 20))
                if (mode == 0) { if (wps == 1) L(0, 1); else if (wps == 2) L(0, 2); else if (wps == 3) L(0, 3); else L(0, 4); }
                if (mode == 1) { if (wps == 1) L(1, 1); else if (wps == 2) L(1, 2); else if (wps == 3) L(1, 3); else L(1, 4); }
                if (mode == 2) { if (wps == 1) L(2, 1); else if (wps == 2) L(2, 2); else if (wps == 3) L(2, 3); else L(2, 4); }
                cudaEventRecord(e1); cudaDeviceSynchronize();
            }
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
            double ev_cyc = ms * 1e-3 * p.clockRate * 1e3 / iters;      // kernel duration in SM cycles (at the nominal max clock) per call
            printf("%s warps/scheduler=%d clock64: cycles/call/warp=%.0f per-scheduler=%.0f | events: per-scheduler cycles/call=%.0f (%.3f ms, clockRate %d kHz)\n", names[mode], wps,
                   (double)c / iters, (double)c / iters / wps, ev_cyc / wps, ms, p.clockRate);
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
