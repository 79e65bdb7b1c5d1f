// K0 microbenchmarks: issue rates of the integer instructions the Fp multiplier is built from,
// measured on the box the kernels run on.  Prints one JSON object.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench.bin tools/microbench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define REP8(x) x x x x x x x x
// 8 independent mad.wide.u32 accumulators, no carries
__global__ void k_wide(uint64_t* out, uint32_t a0, uint32_t b0, int iters) {
    uint32_t a = a0 + threadIdx.x, b = b0 ^ blockIdx.x;
    uint64_t c0 = 1, c1 = 2, c2 = 3, c3 = 4, c4 = 5, c5 = 6, c6 = 7, c7 = 8;
    for (int k = 0; k < iters; k++) {
        REP8(asm volatile("mad.wide.u32 %0, %8, %9, %0;\n\tmad.wide.u32 %1, %8, %9, %1;\n\tmad.wide.u32 %2, %8, %9, %2;\n\tmad.wide.u32 %3, %8, %9, %3;\n\t"
                          "mad.wide.u32 %4, %8, %9, %4;\n\tmad.wide.u32 %5, %8, %9, %5;\n\tmad.wide.u32 %6, %8, %9, %6;\n\tmad.wide.u32 %7, %8, %9, %7;"
                          : "+l"(c0), "+l"(c1), "+l"(c2), "+l"(c3), "+l"(c4), "+l"(c5), "+l"(c6), "+l"(c7) : "r"(a), "r"(b));)
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0 ^ c1 ^ c2 ^ c3 ^ c4 ^ c5 ^ c6 ^ c7;
}
// carry chains: two independent 8-limb accumulators, each updated by a 4-product chain (mad.lo.cc / madc.hi.cc), as one row of the multiplier
__global__ void k_chain(uint32_t* out, uint32_t a0, uint32_t b0, int iters) {
    uint32_t a1 = a0 + threadIdx.x, a2 = a1 * 3, a3 = a1 * 5, a4 = a1 * 7, b = b0 ^ blockIdx.x;
    uint32_t e0 = 1, e1 = 2, e2 = 3, e3 = 4, e4 = 5, e5 = 6, e6 = 7, e7 = 8, o0 = 9, o1 = 10, o2 = 11, o3 = 12, o4 = 13, o5 = 14, o6 = 15, o7 = 16;
    for (int k = 0; k < iters; k++) {
        REP8(asm volatile("mad.lo.cc.u32 %0, %16, %20, %0;\n\tmadc.hi.cc.u32 %1, %16, %20, %1;\n\tmadc.lo.cc.u32 %2, %17, %20, %2;\n\tmadc.hi.cc.u32 %3, %17, %20, %3;\n\t"
                          "madc.lo.cc.u32 %4, %18, %20, %4;\n\tmadc.hi.cc.u32 %5, %18, %20, %5;\n\tmadc.lo.cc.u32 %6, %19, %20, %6;\n\tmadc.hi.u32 %7, %19, %20, %7;\n\t"
                          "mad.lo.cc.u32 %8, %17, %20, %8;\n\tmadc.hi.cc.u32 %9, %17, %20, %9;\n\tmadc.lo.cc.u32 %10, %18, %20, %10;\n\tmadc.hi.cc.u32 %11, %18, %20, %11;\n\t"
                          "madc.lo.cc.u32 %12, %19, %20, %12;\n\tmadc.hi.cc.u32 %13, %19, %20, %13;\n\tmadc.lo.cc.u32 %14, %16, %20, %14;\n\tmadc.hi.u32 %15, %16, %20, %15;"
                          : "+r"(e0), "+r"(e1), "+r"(e2), "+r"(e3), "+r"(e4), "+r"(e5), "+r"(e6), "+r"(e7), "+r"(o0), "+r"(o1), "+r"(o2), "+r"(o3), "+r"(o4), "+r"(o5), "+r"(o6), "+r"(o7)
                          : "r"(a1), "r"(a2), "r"(a3), "r"(a4), "r"(b));)
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = e0 ^ e1 ^ e2 ^ e3 ^ e4 ^ e5 ^ e6 ^ e7 ^ o0 ^ o1 ^ o2 ^ o3 ^ o4 ^ o5 ^ o6 ^ o7;
}
// same 8 products per group but without carries: distinct multiplicands, 8 independent 64-bit accumulators
__global__ void k_wide_distinct(uint64_t* out, uint32_t a0, uint32_t b0, int iters) {
    uint32_t a1 = a0 + threadIdx.x, a2 = a1 * 3, a3 = a1 * 5, a4 = a1 * 7, b = b0 ^ blockIdx.x;
    uint64_t c0 = 1, c1 = 2, c2 = 3, c3 = 4, c4 = 5, c5 = 6, c6 = 7, c7 = 8;
    for (int k = 0; k < iters; k++) {
        REP8(asm volatile("mad.wide.u32 %0, %8, %12, %0;\n\tmad.wide.u32 %1, %9, %12, %1;\n\tmad.wide.u32 %2, %10, %12, %2;\n\tmad.wide.u32 %3, %11, %12, %3;\n\t"
                          "mad.wide.u32 %4, %9, %12, %4;\n\tmad.wide.u32 %5, %10, %12, %5;\n\tmad.wide.u32 %6, %11, %12, %6;\n\tmad.wide.u32 %7, %8, %12, %7;"
                          : "+l"(c0), "+l"(c1), "+l"(c2), "+l"(c3), "+l"(c4), "+l"(c5), "+l"(c6), "+l"(c7) : "r"(a1), "r"(a2), "r"(a3), "r"(a4), "r"(b));)
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0 ^ c1 ^ c2 ^ c3 ^ c4 ^ c5 ^ c6 ^ c7;
}
// plain 32-bit IMAD (lo)
__global__ void k_imad(uint32_t* out, uint32_t a0, uint32_t b0, int iters) {
    uint32_t a = a0 + threadIdx.x, b = b0 ^ blockIdx.x;
    uint32_t c0 = 1, c1 = 2, c2 = 3, c3 = 4, c4 = 5, c5 = 6, c6 = 7, c7 = 8;
    for (int k = 0; k < iters; k++) {
        REP8(asm volatile("mad.lo.u32 %0, %8, %9, %0;\n\tmad.lo.u32 %1, %8, %9, %1;\n\tmad.lo.u32 %2, %8, %9, %2;\n\tmad.lo.u32 %3, %8, %9, %3;\n\t"
                          "mad.lo.u32 %4, %8, %9, %4;\n\tmad.lo.u32 %5, %8, %9, %5;\n\tmad.lo.u32 %6, %8, %9, %6;\n\tmad.lo.u32 %7, %8, %9, %7;"
                          : "+r"(c0), "+r"(c1), "+r"(c2), "+r"(c3), "+r"(c4), "+r"(c5), "+r"(c6), "+r"(c7) : "r"(a), "r"(b));)
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0 ^ c1 ^ c2 ^ c3 ^ c4 ^ c5 ^ c6 ^ c7;
}
// add-with-carry chains (IADD3.X): two independent 8-limb additions per group
__global__ void k_addc(uint32_t* out, uint32_t a0, int iters) {
    uint32_t a = a0 + threadIdx.x;
    uint32_t e0 = 1, e1 = 2, e2 = 3, e3 = 4, e4 = 5, e5 = 6, e6 = 7, e7 = 8, o0 = 9, o1 = 10, o2 = 11, o3 = 12, o4 = 13, o5 = 14, o6 = 15, o7 = 16;
    for (int k = 0; k < iters; k++) {
        REP8(asm volatile("add.cc.u32 %0, %0, %16;\n\taddc.cc.u32 %1, %1, %16;\n\taddc.cc.u32 %2, %2, %16;\n\taddc.cc.u32 %3, %3, %16;\n\t"
                          "addc.cc.u32 %4, %4, %16;\n\taddc.cc.u32 %5, %5, %16;\n\taddc.cc.u32 %6, %6, %16;\n\taddc.u32 %7, %7, %16;\n\t"
                          "add.cc.u32 %8, %8, %16;\n\taddc.cc.u32 %9, %9, %16;\n\taddc.cc.u32 %10, %10, %16;\n\taddc.cc.u32 %11, %11, %16;\n\t"
                          "addc.cc.u32 %12, %12, %16;\n\taddc.cc.u32 %13, %13, %16;\n\taddc.cc.u32 %14, %14, %16;\n\taddc.u32 %15, %15, %16;"
                          : "+r"(e0), "+r"(e1), "+r"(e2), "+r"(e3), "+r"(e4), "+r"(e5), "+r"(e6), "+r"(e7), "+r"(o0), "+r"(o1), "+r"(o2), "+r"(o3), "+r"(o4), "+r"(o5), "+r"(o6), "+r"(o7)
                          : "r"(a));)
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = e0 ^ e1 ^ e2 ^ e3 ^ e4 ^ e5 ^ e6 ^ e7 ^ o0 ^ o1 ^ o2 ^ o3 ^ o4 ^ o5 ^ o6 ^ o7;
}
// double-precision FMA: 8 independent accumulators
__global__ void k_dfma(double* out, double a0, double b0, int iters) {
    double a = a0 + threadIdx.x, b = b0 + blockIdx.x;
    double c0 = 1, c1 = 2, c2 = 3, c3 = 4, c4 = 5, c5 = 6, c6 = 7, c7 = 8;
    for (int k = 0; k < iters; k++) {
        REP8(asm volatile("fma.rn.f64 %0, %8, %9, %0;\n\tfma.rn.f64 %1, %8, %9, %1;\n\tfma.rn.f64 %2, %8, %9, %2;\n\tfma.rn.f64 %3, %8, %9, %3;\n\t"
                          "fma.rn.f64 %4, %8, %9, %4;\n\tfma.rn.f64 %5, %8, %9, %5;\n\tfma.rn.f64 %6, %8, %9, %6;\n\tfma.rn.f64 %7, %8, %9, %7;"
                          : "+d"(c0), "+d"(c1), "+d"(c2), "+d"(c3), "+d"(c4), "+d"(c5), "+d"(c6), "+d"(c7) : "d"(a), "d"(b));)
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0 + c1 + c2 + c3 + c4 + c5 + c6 + c7;
}
// interleaved DFMA + IMAD.WIDE: do the two pipes overlap?
__global__ void k_dfma_wide(double* out, double a0, double b0, uint32_t ia, uint32_t ib, int iters) {
    double a = a0 + threadIdx.x, b = b0 + blockIdx.x;
    uint32_t x = ia + threadIdx.x, y = ib ^ blockIdx.x;
    double c0 = 1, c1 = 2, c2 = 3, c3 = 4;
    uint64_t d0 = 1, d1 = 2, d2 = 3, d3 = 4;
    for (int k = 0; k < iters; k++) {
        REP8(asm volatile("fma.rn.f64 %0, %8, %9, %0;\n\tmad.wide.u32 %4, %10, %11, %4;\n\tfma.rn.f64 %1, %8, %9, %1;\n\tmad.wide.u32 %5, %10, %11, %5;\n\t"
                          "fma.rn.f64 %2, %8, %9, %2;\n\tmad.wide.u32 %6, %10, %11, %6;\n\tfma.rn.f64 %3, %8, %9, %3;\n\tmad.wide.u32 %7, %10, %11, %7;"
                          : "+d"(c0), "+d"(c1), "+d"(c2), "+d"(c3), "+l"(d0), "+l"(d1), "+l"(d2), "+l"(d3) : "d"(a), "d"(b), "r"(x), "r"(y));)
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0 + c1 + c2 + c3 + (double)(d0 ^ d1 ^ d2 ^ d3);
}
// IMAD.WIDE interleaved with independent ALU work (IADD3): do ALU instructions ride along for free?
__global__ void k_wide_alu(uint64_t* out, uint32_t a0, uint32_t b0, int iters) {
    uint32_t a = a0 + threadIdx.x, b = b0 ^ blockIdx.x;
    uint64_t c0 = 1, c1 = 2, c2 = 3, c3 = 4;
    uint32_t s0 = 5, s1 = 6, s2 = 7, s3 = 8;
    for (int k = 0; k < iters; k++) {
        REP8(asm volatile("mad.wide.u32 %0, %8, %9, %0;\n\tadd.u32 %4, %4, %8;\n\tmad.wide.u32 %1, %8, %9, %1;\n\txor.b32 %5, %5, %9;\n\t"
                          "mad.wide.u32 %2, %8, %9, %2;\n\tadd.u32 %6, %6, %9;\n\tmad.wide.u32 %3, %8, %9, %3;\n\tshf.l.wrap.b32 %7, %7, %7, 3;"
                          : "+l"(c0), "+l"(c1), "+l"(c2), "+l"(c3), "+r"(s0), "+r"(s1), "+r"(s2), "+r"(s3) : "r"(a), "r"(b));)
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0 ^ c1 ^ c2 ^ c3 ^ s0 ^ s1 ^ s2 ^ s3;
}

template <class F>
static double time_ms(F launch) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 4; r++) { cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms; }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return best;
}
int main() {
    cudaDeviceProp prop; if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess) { printf("{\"error\": \"no CUDA device\"}\n"); return 1; }
    int sms = prop.multiProcessorCount, blocks = sms * 8, threads = 256, iters = 2048;
    void* buf; cudaMalloc(&buf, (size_t)blocks * threads * 8);
    double nthr = (double)blocks * threads;
    double t;
    printf("{\"device\": \"%s\", \"sms\": %d", prop.name, sms);
    t = time_ms([&] { k_wide<<<blocks, threads>>>((uint64_t*)buf, 0x9e3779b9u, 0x7f4a7c15u, iters); });          printf(", \"imad_wide_per_s\": %.4e", nthr * iters * 64 / (t * 1e-3));
    t = time_ms([&] { k_wide_distinct<<<blocks, threads>>>((uint64_t*)buf, 0x9e3779b9u, 0x7f4a7c15u, iters); }); printf(", \"imad_wide_distinct_per_s\": %.4e", nthr * iters * 64 / (t * 1e-3));
    t = time_ms([&] { k_chain<<<blocks, threads>>>((uint32_t*)buf, 0x9e3779b9u, 0x7f4a7c15u, iters); });         printf(", \"imad_wide_carry_chain_mac_per_s\": %.4e", nthr * iters * 64 / (t * 1e-3));
    t = time_ms([&] { k_imad<<<blocks, threads>>>((uint32_t*)buf, 0x9e3779b9u, 0x7f4a7c15u, iters); });          printf(", \"imad_lo_per_s\": %.4e", nthr * iters * 64 / (t * 1e-3));
    t = time_ms([&] { k_addc<<<blocks, threads>>>((uint32_t*)buf, 0x9e3779b9u, iters); });                       printf(", \"iadd3_x_per_s\": %.4e", nthr * iters * 128 / (t * 1e-3));
    t = time_ms([&] { k_dfma<<<blocks, threads>>>((double*)buf, 1.000001, 0.999999, iters); });                  printf(", \"dfma_per_s\": %.4e", nthr * iters * 64 / (t * 1e-3));
    t = time_ms([&] { k_dfma_wide<<<blocks, threads>>>((double*)buf, 1.000001, 0.999999, 3, 5, iters); });       printf(", \"dfma_plus_wide_pairs_per_s\": %.4e", nthr * iters * 32 / (t * 1e-3));
    t = time_ms([&] { k_wide_alu<<<blocks, threads>>>((uint64_t*)buf, 0x9e3779b9u, 0x7f4a7c15u, iters); });      printf(", \"wide_plus_alu_pairs_per_s\": %.4e", nthr * iters * 32 / (t * 1e-3));
    printf("}\n");
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { fprintf(stderr, "CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
