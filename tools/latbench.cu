// Dependent-issue latencies of the carry-chain instructions (one warp per scheduler, clock64 around an unrolled chain).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/latbench.bin tools/latbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define REP8(x) x x x x x x x x
__global__ void k_lat(uint32_t* out, long long* cyc, const uint32_t* seed, int iters) {
    uint32_t a = seed[threadIdx.x], b = seed[32 + threadIdx.x], c0 = seed[64], c1 = seed[65], c2 = seed[66], c3 = seed[67], c4 = seed[68], c5 = seed[69], c6 = seed[70], c7 = seed[71];
    long long t[8];
    // (0) one carry chain of IADD3.X: addc.cc r_k, r_k, b  (carry AND data dependent on nothing but the flag: distinct registers)
    t[0] = clock64();
    for (int k = 0; k < iters; k++)
        asm volatile("add.cc.u32 %0, %0, %8;\n\t" REP8("addc.cc.u32 %1, %1, %8;\n\taddc.cc.u32 %2, %2, %8;\n\taddc.cc.u32 %3, %3, %8;\n\taddc.cc.u32 %4, %4, %8;\n\t")
                     "addc.u32 %5, %5, %8;" : "+r"(c0), "+r"(c1), "+r"(c2), "+r"(c3), "+r"(c4), "+r"(c5), "+r"(c6), "+r"(c7) : "r"(b));
    t[1] = clock64();
    // (1) one carry chain of IMAD.WIDE.X: (mad.lo.cc, madc.hi.cc) pairs, flag dependent only (distinct accumulators)
    for (int k = 0; k < iters; k++)
        asm volatile("mad.lo.cc.u32 %0, %8, %9, %0;\n\tmadc.hi.cc.u32 %1, %8, %9, %1;\n\t"
                     REP8("madc.lo.cc.u32 %2, %8, %9, %2;\n\tmadc.hi.cc.u32 %3, %8, %9, %3;\n\tmadc.lo.cc.u32 %4, %8, %9, %4;\n\tmadc.hi.cc.u32 %5, %8, %9, %5;\n\t")
                     "madc.lo.cc.u32 %6, %8, %9, %6;\n\tmadc.hi.u32 %7, %8, %9, %7;" : "+r"(c0), "+r"(c1), "+r"(c2), "+r"(c3), "+r"(c4), "+r"(c5), "+r"(c6), "+r"(c7) : "r"(a), "r"(b));
    t[2] = clock64();
    // (2) data-dependent IMAD.WIDE (no carry): acc = a * acc_lo + acc
    uint64_t acc = ((uint64_t)c1 << 32) | c0;
    for (int k = 0; k < iters; k++) {
#pragma unroll
        for (int u = 0; u < 18; u++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"((uint32_t)acc), "r"(a));
    }
    t[3] = clock64();
    // (3) data-dependent IADD3 (no carry)
    for (int k = 0; k < iters; k++) {
#pragma unroll
        for (int u = 0; u < 34; u++) asm volatile("add.u32 %0, %0, %1;" : "+r"(c2) : "r"(b));
    }
    t[4] = clock64();
    // (4) independent IMAD.WIDE (no carry, 8 accumulators): pipe issue interval
    uint64_t q0 = c0, q1 = c1, q2 = c2, q3 = c3, q4 = c4, q5 = c5, q6 = c6, q7 = c7;
    for (int k = 0; k < iters; k++) {
#pragma unroll
        for (int u = 0; u < 4; u++)
            asm volatile("mad.wide.u32 %0, %8, %9, %0;\n\tmad.wide.u32 %1, %10, %9, %1;\n\tmad.wide.u32 %2, %11, %9, %2;\n\tmad.wide.u32 %3, %12, %9, %3;\n\t"
                         "mad.wide.u32 %4, %8, %13, %4;\n\tmad.wide.u32 %5, %10, %13, %5;\n\tmad.wide.u32 %6, %11, %13, %6;\n\tmad.wide.u32 %7, %12, %13, %7;"
                         : "+l"(q0), "+l"(q1), "+l"(q2), "+l"(q3), "+l"(q4), "+l"(q5), "+l"(q6), "+l"(q7) : "r"((uint32_t)q7), "r"(b), "r"((uint32_t)q6), "r"((uint32_t)q5), "r"((uint32_t)q4), "r"(a));
    }
    t[5] = clock64();
    out[threadIdx.x] = c0 ^ c1 ^ c2 ^ c3 ^ c4 ^ c5 ^ c6 ^ c7 ^ (uint32_t)acc ^ (uint32_t)(acc >> 32) ^ (uint32_t)(q0 ^ q1 ^ q2 ^ q3 ^ q4 ^ q5 ^ q6 ^ q7) ^ (uint32_t)((q0 ^ q3) >> 32);
    if (threadIdx.x == 0) for (int i = 0; i < 5; i++) cyc[i] = t[i + 1] - t[i];
}
int main() {
    uint32_t* d; cudaMalloc(&d, 4096); cudaMemset(d, 0x3c, 4096);
    long long* dc; cudaMalloc(&dc, 64);
    const int iters = 1000;
    for (int r = 0; r < 2; r++) { k_lat<<<1, 32>>>(d + 512, dc, d, iters); cudaDeviceSynchronize(); }
    long long c[5]; cudaMemcpy(c, dc, 40, cudaMemcpyDeviceToHost);
    printf("IADD3.X carry chain: %.2f cycles/instr\n", (double)c[0] / iters / 34);
    printf("IMAD.WIDE.X carry chain: %.2f cycles/instr\n", (double)c[1] / iters / 18);
    printf("IMAD.WIDE data-dependent: %.2f cycles/instr\n", (double)c[2] / iters / 18);
    printf("IADD3 data-dependent: %.2f cycles/instr\n", (double)c[3] / iters / 34);
    printf("IMAD.WIDE independent x8: %.2f cycles/instr\n", (double)c[4] / iters / 32);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
