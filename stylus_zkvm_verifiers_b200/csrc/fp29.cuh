// BN254 base field for sm_100a: 9 limbs of 29 bits in signed 32-bit registers, Montgomery R = 2^261,
// products accumulated in signed 64-bit columns with IMAD.WIDE and NO carry flags.
//
// Why not 8 x 32-bit limbs with mad.lo.cc / madc.hi.cc chains (the layout BASELINE.json's north star names)?  It was
// built and measured first (git history: csrc/fp_ptx.cuh): on B200 an IMAD.WIDE that consumes or produces a carry
// predicate issues at HALF the rate of a plain IMAD.WIDE (profiles/r1_microbench.json: 9.2e12 vs 18.4e12 MAC32/s),
// and every carry-chained instruction depends on its predecessor, so that multiplier tops out at 50 % of the integer
// pipe.  With 29-bit limbs a 64-bit column holds 9 products plus the reduction terms without overflow, all 171
// multiply-adds are plain IMAD.WIDE and independent per column, and additions / subtractions are 9 carry-less adds
// (values are kept lazily reduced: see the bounds below).  This replaces the field arithmetic inside the EVM
// precompiles 0x06-0x08 the reference static-calls (/root/reference/contracts/src/common/groth16.rs:54-55,121-125).
//
// Lazy representation.  An element is any limb vector whose value is congruent to x * 2^261 mod p.  Limbs 0..7 are
// "normalised" (N) when they lie in [-2^28, 2^28]; the top limb is signed and small.  Two numbers are tracked per
// value by the bounds build (ZKV_BOUNDS, host only): L = max |limb 0..7| and V = max |value| / p.
//   fp_add/sub : L = La + Lb (< 2^31), V = Va + Vb
//   fp_norm    : one parallel carry pass, L -> 2^28 + small
//   fp_mul     : needs 9 La Lb + 9 2^58 < 2^63; result is N with V = Va Vb p/2^261 + 1  (p/2^261 ~ 1/170)
// All control flow that decides these bounds is input independent, so one bounds-build run of the whole pairing
// (tests/test_host_emu.py) proves the absence of overflow for every input.
#pragma once
#include <stdint.h>

#if defined(ZKV_BOUNDS)
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <execinfo.h>
#endif

namespace zkv {

#define ZKV_LB 29
#define ZKV_LMASK 0x1fffffff
#define ZKV_LHALF 0x10000000

#if defined(ZKV_BOUNDS)
struct fp { int32_t v[9]; double L, V; };
#define ZKV_B(x) x
static inline void zkv_bfail(const char* what, double a, double b) {
    fprintf(stderr, "ZKV_BOUNDS violation: %s (%.3g, %.3g)\n", what, a, b);
    void* bt[24]; int n = backtrace(bt, 24); backtrace_symbols_fd(bt, n, 2);
    abort();
}
#else
struct alignas(8) fp { int32_t v[9]; };      // sizeof == 40: the implicit tail padding keeps 8-byte vector loads/stores
#define ZKV_B(x)
#endif
static const double ZKV_2P28 = 268435456.0, ZKV_2P63 = 9223372036854775808.0;

ZKV_HD ZKV_INLINE fp fp_const(const int32_t* c) { fp r; for (int i = 0; i < 9; i++) r.v[i] = c[i]; ZKV_B(r.L = 2 * ZKV_2P28; r.V = 1;) return r; }
ZKV_HD ZKV_INLINE fp fp_zero() { fp r; for (int i = 0; i < 9; i++) r.v[i] = 0; ZKV_B(r.L = 0; r.V = 0;) return r; }
ZKV_HD ZKV_INLINE fp fp_one() { return fp_const(C_ONE); }

ZKV_HD ZKV_INLINE void fp_add(fp& r, const fp& a, const fp& b) {
    for (int i = 0; i < 9; i++) r.v[i] = a.v[i] + b.v[i];
    ZKV_B(r.L = a.L + b.L; r.V = a.V + b.V; if (r.L >= 8 * ZKV_2P28 - 64 || r.V > 500) zkv_bfail("fp_add", r.L, r.V);)
}
ZKV_HD ZKV_INLINE void fp_sub(fp& r, const fp& a, const fp& b) {
    for (int i = 0; i < 9; i++) r.v[i] = a.v[i] - b.v[i];
    ZKV_B(r.L = a.L + b.L; r.V = a.V + b.V; if (r.L >= 8 * ZKV_2P28 - 64 || r.V > 500) zkv_bfail("fp_sub", r.L, r.V);)
}
ZKV_HD ZKV_INLINE void fp_neg(fp& r, const fp& a) { for (int i = 0; i < 9; i++) r.v[i] = -a.v[i]; ZKV_B(r.L = a.L; r.V = a.V;) }
ZKV_HD ZKV_INLINE void fp_dbl(fp& r, const fp& a) { fp_add(r, a, a); }
// one parallel carry pass to balanced limbs: value unchanged, limbs 0..7 in [-2^28 - c, 2^28 + c]
ZKV_HD ZKV_INLINE void fp_norm(fp& r, const fp& a) {
    int32_t c[8], u[8];
    for (int i = 0; i < 8; i++) { int32_t t = a.v[i] + ZKV_LHALF; c[i] = t >> ZKV_LB; u[i] = t & ZKV_LMASK; }
    int32_t top = a.v[8] + c[7];
    r.v[0] = u[0] - ZKV_LHALF;
    for (int i = 1; i < 8; i++) r.v[i] = u[i] + (c[i - 1] - ZKV_LHALF);
    r.v[8] = top;
    ZKV_B(r.L = ZKV_2P28 + std::ceil(a.L / (2 * ZKV_2P28)) + 1; r.V = a.V;)
}

// 32 x 32 -> 64 products.  The signed form is plain C++ (mul.wide.s32 + add.s64, which ptxas fuses into one IMAD.WIDE with the
// column as addend); the unsigned product by a constant limb of p is spelled in PTX because the C++ front end otherwise widens it
// to a 64-bit multiply with dead high-word corrections.
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ int64_t zkv_umul(uint32_t a, uint32_t b) { int64_t r; asm("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b)); return r; }
#else
static inline int64_t zkv_umul(uint32_t a, uint32_t b) { return (int64_t)((uint64_t)a * b); }
#endif
ZKV_HD ZKV_INLINE int64_t zkv_smul(int32_t a, int32_t b) { return (int64_t)a * b; }
ZKV_HD ZKV_INLINE int64_t zkv_smad(int32_t a, int32_t b, int64_t c) { return c + (int64_t)a * b; }
ZKV_HD ZKV_INLINE int64_t zkv_umad(uint32_t a, uint32_t b, int64_t c) { return c + zkv_umul(a, b); }

// ---- 64-bit column accumulators ------------------------------------------------------------------------------
// cols = sum of up to a few 9x9 limb products, before Montgomery reduction.  Columns 9..16 start at the bias 2^28
// that turns the final carry pass into a balanced one (removed again in fp_redc).
struct cols {
    int64_t t[18];
#if defined(ZKV_BOUNDS)
    double B, V;      // B: bound on |column| without the reduction rows; V: bound on |value| / p^2
#endif
};
#if defined(ZKV_BOUNDS)
static inline double zkv_lim(const fp& a) { return std::fmax(a.L, a.V * 4194304.0 + 16); }
#endif
ZKV_HD ZKV_INLINE void fp_prod_set(cols& c, const fp& a, const fp& b) {
    for (int k = 0; k < 9; k++) c.t[k] = zkv_smul(a.v[0], b.v[k]);
    for (int k = 9; k < 17; k++) c.t[k] = ZKV_LHALF;
    c.t[17] = 0;
#pragma unroll
    for (int i = 1; i < 9; i++)
#pragma unroll
        for (int j = 0; j < 9; j++) c.t[i + j] = zkv_smad(a.v[i], b.v[j], c.t[i + j]);
    ZKV_B(c.B = 9.0 * zkv_lim(a) * zkv_lim(b); c.V = a.V * b.V;)
}
ZKV_HD ZKV_INLINE void fp_prod_add(cols& c, const fp& a, const fp& b) {
#pragma unroll
    for (int i = 0; i < 9; i++)
#pragma unroll
        for (int j = 0; j < 9; j++) c.t[i + j] = zkv_smad(a.v[i], b.v[j], c.t[i + j]);
    ZKV_B(c.B += 9.0 * zkv_lim(a) * zkv_lim(b); c.V += a.V * b.V;)
}
ZKV_HD ZKV_INLINE void fp_prod_sub(cols& c, const fp& a, const fp& b) {
    int32_t n[9];
    for (int i = 0; i < 9; i++) n[i] = -a.v[i];
#pragma unroll
    for (int i = 0; i < 9; i++)
#pragma unroll
        for (int j = 0; j < 9; j++) c.t[i + j] = zkv_smad(n[i], b.v[j], c.t[i + j]);
    ZKV_B(c.B += 9.0 * zkv_lim(a) * zkv_lim(b); c.V += a.V * b.V;)
}
// squares: 45 products
ZKV_HD ZKV_INLINE void fp_prod_sqr_set(cols& c, const fp& a) {
    int32_t d[9];
    for (int i = 0; i < 9; i++) d[i] = a.v[i] + a.v[i];
    c.t[0] = zkv_smul(a.v[0], a.v[0]);
    for (int k = 1; k < 9; k++) c.t[k] = zkv_smul(a.v[0], d[k]);
    for (int k = 9; k < 17; k++) c.t[k] = ZKV_LHALF;
    c.t[17] = 0;
#pragma unroll
    for (int i = 1; i < 9; i++) {
        c.t[2 * i] = zkv_smad(a.v[i], a.v[i], c.t[2 * i]);
#pragma unroll
        for (int j = i + 1; j < 9; j++) c.t[i + j] = zkv_smad(a.v[i], d[j], c.t[i + j]);
    }
    ZKV_B(c.B = 9.0 * zkv_lim(a) * zkv_lim(a); c.V = a.V * a.V; if (zkv_lim(a) >= 4 * ZKV_2P28) zkv_bfail("fp_prod_sqr_set", a.L, a.V);)
}
// Montgomery reduction of the columns: r = cols / 2^261 (mod p), limbs 0..7 balanced in [-2^28, 2^28)
ZKV_HD ZKV_INLINE void fp_redc(fp& r, cols& c) {
    ZKV_B(if (c.B + 9.0 * 288230376151711744.0 + 1e12 >= ZKV_2P63 || c.V > 12000) zkv_bfail("fp_redc columns", c.B, c.V);)
    const int32_t pl[9] = {ZKV_P0, ZKV_P1, ZKV_P2, ZKV_P3, ZKV_P4, ZKV_P5, ZKV_P6, ZKV_P7, ZKV_P8};
#pragma unroll
    for (int i = 0; i < 9; i++) {
        int32_t m = (int32_t)(((uint32_t)c.t[i] * ZKV_PINV29) & ZKV_LMASK);
#pragma unroll
        for (int j = 0; j < 9; j++) c.t[i + j] = zkv_umad((uint32_t)m, (uint32_t)pl[j], c.t[i + j]);
        c.t[i + 1] += c.t[i] >> ZKV_LB;
    }
#pragma unroll
    for (int k = 0; k < 8; k++) { r.v[k] = (int32_t)((uint32_t)c.t[9 + k] & ZKV_LMASK) - ZKV_LHALF; c.t[10 + k] += c.t[9 + k] >> ZKV_LB; }
    r.v[8] = (int32_t)c.t[17];
   
    ZKV_B(r.L = ZKV_2P28; r.V = c.V / 169.0 + 1;)
}
// Montgomery product a * b / 2^261 (mod p): 81 + 9 + 81 IMAD(.WIDE), the reduction rows use p's limbs as immediates
ZKV_HD ZKV_INLINE void fp_mul(fp& r, const fp& a, const fp& b) { cols c; fp_prod_set(c, a, b); fp_redc(r, c); }
ZKV_HD ZKV_INLINE void fp_sqr(fp& r, const fp& a) { cols c; fp_prod_sqr_set(c, a); fp_redc(r, c); }
// a / 2 (mod p): make the value even by adding p when limb 0 is odd, then shift the whole number right by one bit
ZKV_HD ZKV_INLINE void fp_half(fp& r, const fp& a) {
    const int32_t pl[9] = {ZKV_P0, ZKV_P1, ZKV_P2, ZKV_P3, ZKV_P4, ZKV_P5, ZKV_P6, ZKV_P7, ZKV_P8};
    int32_t odd = -(a.v[0] & 1), t[9];
    for (int i = 0; i < 9; i++) t[i] = a.v[i] + (pl[i] & odd);
    for (int i = 0; i < 8; i++) r.v[i] = (t[i] >> 1) + ((t[i + 1] & 1) << 28);
    r.v[8] = t[8] >> 1;
    ZKV_B(r.L = (a.L + 2 * ZKV_2P28) / 2 + ZKV_2P28 + 1; r.V = (a.V + 1) / 2; if (a.L + 2 * ZKV_2P28 >= 8 * ZKV_2P28) zkv_bfail("fp_half", a.L, a.V);)
}
ZKV_HD ZKV_INLINE void fp_to_mont(fp& r, const fp& a) { fp r2 = fp_const(C_R2); fp_mul(r, a, r2); }

// Canonical representative: limbs 0..8 in [0, 2^29), value in [0, p).  Used for comparisons and byte output only.
ZKV_HD ZKV_INLINE void fp_canon(fp& r, const fp& a) {
    ZKV_B(if (a.V > 150) zkv_bfail("fp_canon", a.L, a.V);)
    fp one = fp_const(C_ONE), y;
    ZKV_B(one.L = 2 * ZKV_2P28;)
    fp n; fp_norm(n, a);
    fp_mul(y, n, one);                                   // same residue, value in (-p, 2p)
    const int32_t pl[9] = {ZKV_P0, ZKV_P1, ZKV_P2, ZKV_P3, ZKV_P4, ZKV_P5, ZKV_P6, ZKV_P7, ZKV_P8};
    int32_t t[9], c = 0;
    for (int i = 0; i < 9; i++) { int32_t s = y.v[i] + pl[i] + c; if (i < 8) { t[i] = s & ZKV_LMASK; c = s >> ZKV_LB; } else t[i] = s; }   // y + p in (0, 3p)
    for (int rep = 0; rep < 2; rep++) {
        int32_t d[9], bw = 0;
        for (int i = 0; i < 9; i++) { int32_t s = t[i] - pl[i] + bw; if (i < 8) { d[i] = s & ZKV_LMASK; bw = s >> ZKV_LB; } else d[i] = s; }
        int32_t keep = d[8] >> 31;                       // all ones if t < p
        for (int i = 0; i < 9; i++) t[i] = (t[i] & keep) | (d[i] & ~keep);
    }
    for (int i = 0; i < 9; i++) r.v[i] = t[i];
   
    ZKV_B(r.L = 2 * ZKV_2P28; r.V = 1;)
}
ZKV_HD ZKV_INLINE bool fp_is_zero(const fp& a) { fp c; fp_canon(c, a); int32_t t = 0; for (int i = 0; i < 9; i++) t |= c.v[i]; return t == 0; }
ZKV_HD ZKV_INLINE bool fp_eq(const fp& a, const fp& b) { fp d; fp_sub(d, a, b); return fp_is_zero(d); }

// raw (non-Montgomery) 256-bit compare a >= m on 8 x 32-bit words
ZKV_HD ZKV_INLINE bool u256_geq(const uint32_t* a, const uint32_t* m) {
    uint32_t borrow = 0;
    for (int i = 0; i < 8; i++) { uint64_t t = (uint64_t)a[i] - m[i] - borrow; borrow = (uint32_t)(t >> 63); }
    return borrow == 0;
}
// 8 x 32-bit words (value < 2^256) -> limbs (no reduction, not Montgomery)
ZKV_HD ZKV_INLINE void fp_from_words(fp& r, const uint32_t* w) {
    for (int i = 0; i < 9; i++) {
        int bit = ZKV_LB * i, wi = bit >> 5, sh = bit & 31;
        uint32_t lo = w[wi] >> sh;
        if (sh > 3 && wi + 1 < 8) lo |= w[wi + 1] << (32 - sh);
        r.v[i] = (int32_t)(lo & ZKV_LMASK);
    }
   
    ZKV_B(r.L = 2 * ZKV_2P28; r.V = 5.3;)    // 2^256 / p
}
// canonical limbs -> 8 x 32-bit words
ZKV_HD ZKV_INLINE void fp_to_words(uint32_t* w, const fp& c) {
    for (int j = 0; j < 8; j++) {
        int bit = 32 * j, li = bit / ZKV_LB, sh = bit % ZKV_LB;
        uint64_t v = (uint64_t)(uint32_t)c.v[li] >> sh;
        v |= (uint64_t)(uint32_t)c.v[li + 1] << (ZKV_LB - sh);
        if (li + 2 < 9) v |= (uint64_t)(uint32_t)c.v[li + 2] << (2 * ZKV_LB - sh);
        w[j] = (uint32_t)v;
    }
}
// Montgomery -> canonical integer words
ZKV_HD ZKV_INLINE void fp_from_mont_words(uint32_t* w, const fp& a) {
    fp n, one = fp_zero(), y, c; one.v[0] = 1; ZKV_B(one.L = 1; one.V = 1;)
    fp_norm(n, a); fp_mul(y, n, one);        // a / R: the plain residue, lazily reduced
    // canonicalise y without another Montgomery multiplication
    const int32_t pl[9] = {ZKV_P0, ZKV_P1, ZKV_P2, ZKV_P3, ZKV_P4, ZKV_P5, ZKV_P6, ZKV_P7, ZKV_P8};
    int32_t t[9], cy = 0;
    for (int i = 0; i < 9; i++) { int32_t s = y.v[i] + pl[i] + cy; if (i < 8) { t[i] = s & ZKV_LMASK; cy = s >> ZKV_LB; } else t[i] = s; }
    for (int rep = 0; rep < 2; rep++) {
        int32_t d[9], bw = 0;
        for (int i = 0; i < 9; i++) { int32_t s = t[i] - pl[i] + bw; if (i < 8) { d[i] = s & ZKV_LMASK; bw = s >> ZKV_LB; } else d[i] = s; }
        int32_t keep = d[8] >> 31;
        for (int i = 0; i < 9; i++) t[i] = (t[i] & keep) | (d[i] & ~keep);
    }
    for (int i = 0; i < 9; i++) c.v[i] = t[i];
    fp_to_words(w, c);
}
// a^(p-2); inv(0) = 0
ZKV_HD ZKV_NOINLINE void fp_inv(fp& r, const fp& a) {
    fp base; fp_norm(base, a);
    fp acc = fp_one();
    for (int i = 253; i >= 0; i--) {
        fp_sqr(acc, acc);
        if ((C_PM2W[i >> 5] >> (i & 31)) & 1) fp_mul(acc, acc, base);
    }
    r = acc;
}

}  // namespace zkv
