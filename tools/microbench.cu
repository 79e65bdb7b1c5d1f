// (round 2: the loop-invariant mad.wide kernels of round 1 were removed - ptxas strength-reduces them to 64-bit adds, so their
// figures were IADD3 rates; the 9x9 / immediate / constant-bank / register-operand mad.wide kernels went the same way in round 2: all but
// one multiplicand were loop invariant, so nvcc hoisted the products out of the loop.  Every kernel left here has its multiplies INSIDE
// the timed loop in the SASS (cuobjdump -sass: k_chain 320 IMAD.WIDE.U32[.X] per iteration, k_asm_*: data-dependent multiplicands).)
// K0 microbenchmarks: issue rates of the integer instructions the Fp multiplier is built from,
// measured on the box the kernels run on.  Prints one JSON object.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench.bin tools/microbench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define REP8(x) x x x x x x x x
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



// every multiply takes the low word of its own accumulator as multiplicand, so nothing is loop invariant (reg)
__global__ void k_asm_u32_rr(uint64_t* out, uint32_t a0, uint32_t b0, int iters) {
    uint32_t b[8]; uint64_t c[16];
    for (int i = 0; i < 8; i++) b[i] = (b0 * (i + 3) ^ blockIdx.x) | 1u;
    for (int i = 0; i < 16; i++) c[i] = a0 * (i + 1) + threadIdx.x;
    for (int k = 0; k < iters; k++) {
        asm volatile("{\n\t.reg .b32 tl<16>, th<16>;\n\t"
            "mov.b64 {tl0, th0}, %0;\n\t"
            "mad.wide.u32 %0, tl0, %16, %0;\n\t"
            "mov.b64 {tl1, th1}, %1;\n\t"
            "mad.wide.u32 %1, tl1, %21, %1;\n\t"
            "mov.b64 {tl2, th2}, %2;\n\t"
            "mad.wide.u32 %2, tl2, %18, %2;\n\t"
            "mov.b64 {tl3, th3}, %3;\n\t"
            "mad.wide.u32 %3, tl3, %23, %3;\n\t"
            "mov.b64 {tl4, th4}, %4;\n\t"
            "mad.wide.u32 %4, tl4, %20, %4;\n\t"
            "mov.b64 {tl5, th5}, %5;\n\t"
            "mad.wide.u32 %5, tl5, %17, %5;\n\t"
            "mov.b64 {tl6, th6}, %6;\n\t"
            "mad.wide.u32 %6, tl6, %22, %6;\n\t"
            "mov.b64 {tl7, th7}, %7;\n\t"
            "mad.wide.u32 %7, tl7, %19, %7;\n\t"
            "mov.b64 {tl8, th8}, %8;\n\t"
            "mad.wide.u32 %8, tl8, %16, %8;\n\t"
            "mov.b64 {tl9, th9}, %9;\n\t"
            "mad.wide.u32 %9, tl9, %21, %9;\n\t"
            "mov.b64 {tl10, th10}, %10;\n\t"
            "mad.wide.u32 %10, tl10, %18, %10;\n\t"
            "mov.b64 {tl11, th11}, %11;\n\t"
            "mad.wide.u32 %11, tl11, %23, %11;\n\t"
            "mov.b64 {tl12, th12}, %12;\n\t"
            "mad.wide.u32 %12, tl12, %20, %12;\n\t"
            "mov.b64 {tl13, th13}, %13;\n\t"
            "mad.wide.u32 %13, tl13, %17, %13;\n\t"
            "mov.b64 {tl14, th14}, %14;\n\t"
            "mad.wide.u32 %14, tl14, %22, %14;\n\t"
            "mov.b64 {tl15, th15}, %15;\n\t"
            "mad.wide.u32 %15, tl15, %19, %15;\n\t"
            "mov.b64 {tl0, th0}, %0;\n\t"
            "mad.wide.u32 %0, tl0, %17, %0;\n\t"
            "mov.b64 {tl1, th1}, %1;\n\t"
            "mad.wide.u32 %1, tl1, %22, %1;\n\t"
            "mov.b64 {tl2, th2}, %2;\n\t"
            "mad.wide.u32 %2, tl2, %19, %2;\n\t"
            "mov.b64 {tl3, th3}, %3;\n\t"
            "mad.wide.u32 %3, tl3, %16, %3;\n\t"
            "mov.b64 {tl4, th4}, %4;\n\t"
            "mad.wide.u32 %4, tl4, %21, %4;\n\t"
            "mov.b64 {tl5, th5}, %5;\n\t"
            "mad.wide.u32 %5, tl5, %18, %5;\n\t"
            "mov.b64 {tl6, th6}, %6;\n\t"
            "mad.wide.u32 %6, tl6, %23, %6;\n\t"
            "mov.b64 {tl7, th7}, %7;\n\t"
            "mad.wide.u32 %7, tl7, %20, %7;\n\t"
            "mov.b64 {tl8, th8}, %8;\n\t"
            "mad.wide.u32 %8, tl8, %17, %8;\n\t"
            "mov.b64 {tl9, th9}, %9;\n\t"
            "mad.wide.u32 %9, tl9, %22, %9;\n\t"
            "mov.b64 {tl10, th10}, %10;\n\t"
            "mad.wide.u32 %10, tl10, %19, %10;\n\t"
            "mov.b64 {tl11, th11}, %11;\n\t"
            "mad.wide.u32 %11, tl11, %16, %11;\n\t"
            "mov.b64 {tl12, th12}, %12;\n\t"
            "mad.wide.u32 %12, tl12, %21, %12;\n\t"
            "mov.b64 {tl13, th13}, %13;\n\t"
            "mad.wide.u32 %13, tl13, %18, %13;\n\t"
            "mov.b64 {tl14, th14}, %14;\n\t"
            "mad.wide.u32 %14, tl14, %23, %14;\n\t"
            "mov.b64 {tl15, th15}, %15;\n\t"
            "mad.wide.u32 %15, tl15, %20, %15;\n\t"
            "}"
            : "+l"(c[0]), "+l"(c[1]), "+l"(c[2]), "+l"(c[3]), "+l"(c[4]), "+l"(c[5]), "+l"(c[6]), "+l"(c[7]), "+l"(c[8]), "+l"(c[9]), "+l"(c[10]), "+l"(c[11]), "+l"(c[12]), "+l"(c[13]), "+l"(c[14]), "+l"(c[15]) : "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
    }
    uint64_t x = 0;
    for (int i = 0; i < 16; i++) x ^= c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

// every multiply takes the low word of its own accumulator as multiplicand, so nothing is loop invariant (reg)
__global__ void k_asm_s32_rr(uint64_t* out, uint32_t a0, uint32_t b0, int iters) {
    uint32_t b[8]; uint64_t c[16];
    for (int i = 0; i < 8; i++) b[i] = (b0 * (i + 3) ^ blockIdx.x) | 1u;
    for (int i = 0; i < 16; i++) c[i] = a0 * (i + 1) + threadIdx.x;
    for (int k = 0; k < iters; k++) {
        asm volatile("{\n\t.reg .b32 tl<16>, th<16>;\n\t"
            "mov.b64 {tl0, th0}, %0;\n\t"
            "mad.wide.s32 %0, tl0, %16, %0;\n\t"
            "mov.b64 {tl1, th1}, %1;\n\t"
            "mad.wide.s32 %1, tl1, %21, %1;\n\t"
            "mov.b64 {tl2, th2}, %2;\n\t"
            "mad.wide.s32 %2, tl2, %18, %2;\n\t"
            "mov.b64 {tl3, th3}, %3;\n\t"
            "mad.wide.s32 %3, tl3, %23, %3;\n\t"
            "mov.b64 {tl4, th4}, %4;\n\t"
            "mad.wide.s32 %4, tl4, %20, %4;\n\t"
            "mov.b64 {tl5, th5}, %5;\n\t"
            "mad.wide.s32 %5, tl5, %17, %5;\n\t"
            "mov.b64 {tl6, th6}, %6;\n\t"
            "mad.wide.s32 %6, tl6, %22, %6;\n\t"
            "mov.b64 {tl7, th7}, %7;\n\t"
            "mad.wide.s32 %7, tl7, %19, %7;\n\t"
            "mov.b64 {tl8, th8}, %8;\n\t"
            "mad.wide.s32 %8, tl8, %16, %8;\n\t"
            "mov.b64 {tl9, th9}, %9;\n\t"
            "mad.wide.s32 %9, tl9, %21, %9;\n\t"
            "mov.b64 {tl10, th10}, %10;\n\t"
            "mad.wide.s32 %10, tl10, %18, %10;\n\t"
            "mov.b64 {tl11, th11}, %11;\n\t"
            "mad.wide.s32 %11, tl11, %23, %11;\n\t"
            "mov.b64 {tl12, th12}, %12;\n\t"
            "mad.wide.s32 %12, tl12, %20, %12;\n\t"
            "mov.b64 {tl13, th13}, %13;\n\t"
            "mad.wide.s32 %13, tl13, %17, %13;\n\t"
            "mov.b64 {tl14, th14}, %14;\n\t"
            "mad.wide.s32 %14, tl14, %22, %14;\n\t"
            "mov.b64 {tl15, th15}, %15;\n\t"
            "mad.wide.s32 %15, tl15, %19, %15;\n\t"
            "mov.b64 {tl0, th0}, %0;\n\t"
            "mad.wide.s32 %0, tl0, %17, %0;\n\t"
            "mov.b64 {tl1, th1}, %1;\n\t"
            "mad.wide.s32 %1, tl1, %22, %1;\n\t"
            "mov.b64 {tl2, th2}, %2;\n\t"
            "mad.wide.s32 %2, tl2, %19, %2;\n\t"
            "mov.b64 {tl3, th3}, %3;\n\t"
            "mad.wide.s32 %3, tl3, %16, %3;\n\t"
            "mov.b64 {tl4, th4}, %4;\n\t"
            "mad.wide.s32 %4, tl4, %21, %4;\n\t"
            "mov.b64 {tl5, th5}, %5;\n\t"
            "mad.wide.s32 %5, tl5, %18, %5;\n\t"
            "mov.b64 {tl6, th6}, %6;\n\t"
            "mad.wide.s32 %6, tl6, %23, %6;\n\t"
            "mov.b64 {tl7, th7}, %7;\n\t"
            "mad.wide.s32 %7, tl7, %20, %7;\n\t"
            "mov.b64 {tl8, th8}, %8;\n\t"
            "mad.wide.s32 %8, tl8, %17, %8;\n\t"
            "mov.b64 {tl9, th9}, %9;\n\t"
            "mad.wide.s32 %9, tl9, %22, %9;\n\t"
            "mov.b64 {tl10, th10}, %10;\n\t"
            "mad.wide.s32 %10, tl10, %19, %10;\n\t"
            "mov.b64 {tl11, th11}, %11;\n\t"
            "mad.wide.s32 %11, tl11, %16, %11;\n\t"
            "mov.b64 {tl12, th12}, %12;\n\t"
            "mad.wide.s32 %12, tl12, %21, %12;\n\t"
            "mov.b64 {tl13, th13}, %13;\n\t"
            "mad.wide.s32 %13, tl13, %18, %13;\n\t"
            "mov.b64 {tl14, th14}, %14;\n\t"
            "mad.wide.s32 %14, tl14, %23, %14;\n\t"
            "mov.b64 {tl15, th15}, %15;\n\t"
            "mad.wide.s32 %15, tl15, %20, %15;\n\t"
            "}"
            : "+l"(c[0]), "+l"(c[1]), "+l"(c[2]), "+l"(c[3]), "+l"(c[4]), "+l"(c[5]), "+l"(c[6]), "+l"(c[7]), "+l"(c[8]), "+l"(c[9]), "+l"(c[10]), "+l"(c[11]), "+l"(c[12]), "+l"(c[13]), "+l"(c[14]), "+l"(c[15]) : "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
    }
    uint64_t x = 0;
    for (int i = 0; i < 16; i++) x ^= c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

// every multiply takes the low word of its own accumulator as multiplicand, so nothing is loop invariant (imm)
__global__ void k_asm_u32_ri(uint64_t* out, uint32_t a0, uint32_t b0, int iters) {
    uint32_t b[8]; uint64_t c[16];
    for (int i = 0; i < 8; i++) b[i] = (b0 * (i + 3) ^ blockIdx.x) | 1u;
    for (int i = 0; i < 16; i++) c[i] = a0 * (i + 1) + threadIdx.x;
    for (int k = 0; k < iters; k++) {
        asm volatile("{\n\t.reg .b32 tl<16>, th<16>;\n\t"
            "mov.b64 {tl0, th0}, %0;\n\t"
            "mad.wide.u32 %0, tl0, 0x187cfd47, %0;\n\t"
            "mov.b64 {tl1, th1}, %1;\n\t"
            "mad.wide.u32 %1, tl1, 0x02db40c0, %1;\n\t"
            "mov.b64 {tl2, th2}, %2;\n\t"
            "mad.wide.u32 %2, tl2, 0x1c72a34f, %2;\n\t"
            "mov.b64 {tl3, th3}, %3;\n\t"
            "mad.wide.u32 %3, tl3, 0x0e5c2634, %3;\n\t"
            "mov.b64 {tl4, th4}, %4;\n\t"
            "mad.wide.u32 %4, tl4, 0x1585d978, %4;\n\t"
            "mov.b64 {tl5, th5}, %5;\n\t"
            "mad.wide.u32 %5, tl5, 0x010460b6, %5;\n\t"
            "mov.b64 {tl6, th6}, %6;\n\t"
            "mad.wide.u32 %6, tl6, 0x00a6e141, %6;\n\t"
            "mov.b64 {tl7, th7}, %7;\n\t"
            "mad.wide.u32 %7, tl7, 0x02d522d0, %7;\n\t"
            "mov.b64 {tl8, th8}, %8;\n\t"
            "mad.wide.u32 %8, tl8, 0x187cfd47, %8;\n\t"
            "mov.b64 {tl9, th9}, %9;\n\t"
            "mad.wide.u32 %9, tl9, 0x02db40c0, %9;\n\t"
            "mov.b64 {tl10, th10}, %10;\n\t"
            "mad.wide.u32 %10, tl10, 0x1c72a34f, %10;\n\t"
            "mov.b64 {tl11, th11}, %11;\n\t"
            "mad.wide.u32 %11, tl11, 0x0e5c2634, %11;\n\t"
            "mov.b64 {tl12, th12}, %12;\n\t"
            "mad.wide.u32 %12, tl12, 0x1585d978, %12;\n\t"
            "mov.b64 {tl13, th13}, %13;\n\t"
            "mad.wide.u32 %13, tl13, 0x010460b6, %13;\n\t"
            "mov.b64 {tl14, th14}, %14;\n\t"
            "mad.wide.u32 %14, tl14, 0x00a6e141, %14;\n\t"
            "mov.b64 {tl15, th15}, %15;\n\t"
            "mad.wide.u32 %15, tl15, 0x02d522d0, %15;\n\t"
            "mov.b64 {tl0, th0}, %0;\n\t"
            "mad.wide.u32 %0, tl0, 0x010460b6, %0;\n\t"
            "mov.b64 {tl1, th1}, %1;\n\t"
            "mad.wide.u32 %1, tl1, 0x00a6e141, %1;\n\t"
            "mov.b64 {tl2, th2}, %2;\n\t"
            "mad.wide.u32 %2, tl2, 0x02d522d0, %2;\n\t"
            "mov.b64 {tl3, th3}, %3;\n\t"
            "mad.wide.u32 %3, tl3, 0x187cfd47, %3;\n\t"
            "mov.b64 {tl4, th4}, %4;\n\t"
            "mad.wide.u32 %4, tl4, 0x02db40c0, %4;\n\t"
            "mov.b64 {tl5, th5}, %5;\n\t"
            "mad.wide.u32 %5, tl5, 0x1c72a34f, %5;\n\t"
            "mov.b64 {tl6, th6}, %6;\n\t"
            "mad.wide.u32 %6, tl6, 0x0e5c2634, %6;\n\t"
            "mov.b64 {tl7, th7}, %7;\n\t"
            "mad.wide.u32 %7, tl7, 0x1585d978, %7;\n\t"
            "mov.b64 {tl8, th8}, %8;\n\t"
            "mad.wide.u32 %8, tl8, 0x010460b6, %8;\n\t"
            "mov.b64 {tl9, th9}, %9;\n\t"
            "mad.wide.u32 %9, tl9, 0x00a6e141, %9;\n\t"
            "mov.b64 {tl10, th10}, %10;\n\t"
            "mad.wide.u32 %10, tl10, 0x02d522d0, %10;\n\t"
            "mov.b64 {tl11, th11}, %11;\n\t"
            "mad.wide.u32 %11, tl11, 0x187cfd47, %11;\n\t"
            "mov.b64 {tl12, th12}, %12;\n\t"
            "mad.wide.u32 %12, tl12, 0x02db40c0, %12;\n\t"
            "mov.b64 {tl13, th13}, %13;\n\t"
            "mad.wide.u32 %13, tl13, 0x1c72a34f, %13;\n\t"
            "mov.b64 {tl14, th14}, %14;\n\t"
            "mad.wide.u32 %14, tl14, 0x0e5c2634, %14;\n\t"
            "mov.b64 {tl15, th15}, %15;\n\t"
            "mad.wide.u32 %15, tl15, 0x1585d978, %15;\n\t"
            "}"
            : "+l"(c[0]), "+l"(c[1]), "+l"(c[2]), "+l"(c[3]), "+l"(c[4]), "+l"(c[5]), "+l"(c[6]), "+l"(c[7]), "+l"(c[8]), "+l"(c[9]), "+l"(c[10]), "+l"(c[11]), "+l"(c[12]), "+l"(c[13]), "+l"(c[14]), "+l"(c[15]) : "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
    }
    uint64_t x = 0;
    for (int i = 0; i < 16; i++) x ^= c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
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
    t = time_ms([&] { k_chain<<<blocks, threads>>>((uint32_t*)buf, 0x9e3779b9u, 0x7f4a7c15u, iters); });         printf(", \"imad_wide_carry_chain_mac_per_s\": %.4e", nthr * iters * 64 / (t * 1e-3));
    t = time_ms([&] { k_imad<<<blocks, threads>>>((uint32_t*)buf, 0x9e3779b9u, 0x7f4a7c15u, iters); });          printf(", \"imad_lo_per_s\": %.4e", nthr * iters * 64 / (t * 1e-3));
    t = time_ms([&] { k_addc<<<blocks, threads>>>((uint32_t*)buf, 0x9e3779b9u, iters); });                       printf(", \"iadd3_x_per_s\": %.4e", nthr * iters * 128 / (t * 1e-3));
    t = time_ms([&] { k_dfma<<<blocks, threads>>>((double*)buf, 1.000001, 0.999999, iters); });                  printf(", \"dfma_per_s\": %.4e", nthr * iters * 64 / (t * 1e-3));
    int blocks2 = sms * 4;   // 64+ registers per thread: 4 x 256 threads per SM
    double nthr2 = (double)blocks2 * threads; int it3 = 4096;
    t = time_ms([&] { k_asm_u32_rr<<<blocks2, threads>>>((uint64_t*)buf, 0x1234567, 0x7654321, it3); }); printf(", \"asm_mad_wide_u32_distinct_regs_per_s\": %.4e", nthr2 * it3 * 32 / (t * 1e-3));
    t = time_ms([&] { k_asm_s32_rr<<<blocks2, threads>>>((uint64_t*)buf, 0x1234567, 0x7654321, it3); }); printf(", \"asm_mad_wide_s32_distinct_regs_per_s\": %.4e", nthr2 * it3 * 32 / (t * 1e-3));
    t = time_ms([&] { k_asm_u32_ri<<<blocks2, threads>>>((uint64_t*)buf, 0x1234567, 0x7654321, it3); }); printf(", \"asm_mad_wide_u32_imm_per_s\": %.4e", nthr2 * it3 * 32 / (t * 1e-3));
    printf("}\n");
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { fprintf(stderr, "CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
