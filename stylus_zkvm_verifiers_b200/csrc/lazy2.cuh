// TWO LANES PER PROOF: the lazily reduced Fp12 arithmetic of lazy.cuh with every Fp2 value split over two adjacent lanes of a warp.
//
// lazy.cuh (one proof per thread) is bounded by what one thread can hold: 255 registers and 28 shared-memory slots per proof allow two
// warps per scheduler, the Karatsuba recombinations of the Fp2 products cost as many additions as there are multiplications, and a warp
// that is adding leaves the multiplier pipe idle unless the one other warp of its scheduler happens to be multiplying (executed IMAD.WIDE
// issue share 0.56, profiles/r2_ncu_summary.json).  Here lane 2i owns the REAL and lane 2i+1 the IMAGINARY component of every Fp2 value
// of proof i, and both run one instruction stream (operands are picked by lane parity, nothing diverges):
//   * an Fp2 product is ONE 128-IMAD.WIDE accumulation chain per lane (lz_mulw2: x0 t1 + x1 t2 with (t1, t2) = (y0, p - y1) on the real
//     lane and (y1, y0) on the imaginary one): no Karatsuba recombination, no wide subtraction, a quarter of the additions;
//   * half the live state per lane: 128 registers, FOUR warps per scheduler with the same 28 slots x 128 proofs of shared memory per block;
//   * the only cross-lane traffic is the multiplication by xi = 9 + u of an unreduced value (16 shuffles) and `__syncwarp` around
//     shared-memory slots that the partner lane reads.
// Costs: schoolbook instead of Karatsuba over Fp2, i.e. +23 % IMAD.WIDE per sparse / dense Fp6 product (squarings cost the same).
// Slot layout as in lazy.cuh (LZ_NT proofs per block, slot s of proof i at lz_sm[(2 s + h) * LZ_NT + i]); a block has 2 * LZ_NT threads.
// The routines of csrc/lazy2_gen.cuh are generated and bound-checked for both lanes by tools/gen_lazy2.py.  All values at rest are canonical
// Montgomery residues, so results are bit-identical to lazy.cuh, bn254.cuh and the oracle.
// (arithmetic behind the reference's ecPairing precompile call, /root/reference/contracts/src/common/groth16.rs:121-125)
#pragma once
#include "lazy.cuh"

namespace zkv {

struct lz2_r3 { fp c[3]; };      // own components of three Fp2 values
struct lz2_r2 { fp c[2]; };

#if defined(__CUDACC__)
LZ_INL void lz_sel8(uint32_t* r, const uint32_t* x, const uint32_t* y, uint32_t im) { for (int i = 0; i < 8; i++) r[i] = im ? y[i] : x[i]; }
LZ_INL void lz_xchgw(uint32_t* r, const uint32_t* w) { for (int i = 0; i < 16; i++) r[i] = __shfl_xor_sync(0xffffffffu, w[i], 1); }
// r = im ? x : c - x   (c >= x: a multiple of p 2^224, so the real lane gets a non-negative representative of -x)
LZ_INL void lz_negsel(uint32_t* r, const uint32_t* x, const uint32_t* c, uint32_t im) {
    uint32_t t[16];
    lz_subw(t, c, x);
    for (int i = 0; i < 16; i++) r[i] = im ? x[i] : t[i];
}
LZ_INL uint32_t lz2_pid() { return threadIdx.x >> 1; }
LZ_INL uint32_t lz2_im() { return threadIdx.x & 1; }
LZ_INL void lz2_sync() { __syncwarp(); }

#include "lazy2_gen.cuh"

// own-component helpers on canonical values
LZ_INL fp fpv_add(const fp& a, const fp& b) { fp r; fp_add(r, a, b); return r; }
LZ_INL fp fpv_sub(const fp& a, const fp& b) { fp r; fp_sub(r, a, b); return r; }
LZ_INL fp fpv_neg(const fp& a) { fp r; fp_neg(r, a); return r; }
LZ_INL fp fpv_sel(const fp& x, const fp& y, uint32_t im) { fp r; for (int i = 0; i < 8; i++) r.v[i] = im ? y.v[i] : x.v[i]; return r; }
// own component of (9 + u) a given a's own and other component: real lane 9 re - im, imaginary lane 9 im + re
LZ_INL fp lz2_xi(const fp& own, const fp& oth, uint32_t im) {
    fp t; fp_dbl(t, own); fp_dbl(t, t); fp_dbl(t, t); fp_add(t, t, own);
    return fpv_add(t, fpv_sel(fpv_neg(oth), oth, im));
}
struct Lz2 { uint32_t own, oth, im; };       // slot offsets of the lane's own / other component, lane parity
LZ_INL Lz2 lz2_ctx() { Lz2 c; c.im = lz2_im(); c.own = c.im * LZ_SLOT; c.oth = LZ_SLOT - c.own; return c; }
LZ_INL fp lz2_ldo(uint32_t base, int k, const Lz2& c) { return lz_ldfp(base + 2 * k * LZ_SLOT + c.own); }       // own component of Fp2 number k
LZ_INL fp lz2_ldx(uint32_t base, int k, const Lz2& c) { return lz_ldfp(base + 2 * k * LZ_SLOT + c.oth); }       // the other component
LZ_INL void lz2_sto(uint32_t base, int k, const Lz2& c, const fp& v) { lz_stfp(base + 2 * k * LZ_SLOT + c.own, v); }

// f (12 slots at `f`) <- f^2, complex squaring as lz_f12sqr; `t` = 6-slot temporary
LZ_FN2 void lz2_f12sqr(uint32_t f, uint32_t t, Lz2 c) {
    const uint32_t f1 = f + 6 * LZ_SLOT;
    { lz2_r3 ab = lz2_f6mul(f, f1, c.own, c.oth, c.im); for (int k = 0; k < 3; k++) lz2_sto(t, k, c, ab.c[k]); }
    {
        fp x0 = lz2_ldo(f, 0, c), x1 = lz2_ldo(f, 1, c), x2 = lz2_ldo(f, 2, c);
        fp y0 = lz2_ldo(f1, 0, c), y1 = lz2_ldo(f1, 1, c), y2 = lz2_ldo(f1, 2, c), y2x = lz2_ldx(f1, 2, c);
        lz2_sync();                                  // the partner has read what it needs of f before anything is overwritten
        lz2_sto(f, 0, c, fpv_add(x0, y0)); lz2_sto(f, 1, c, fpv_add(x1, y1)); lz2_sto(f, 2, c, fpv_add(x2, y2));
        lz2_sto(f1, 0, c, fpv_add(x0, lz2_xi(y2, y2x, c.im))); lz2_sto(f1, 1, c, fpv_add(x1, y0)); lz2_sto(f1, 2, c, fpv_add(x2, y1));
        lz2_sync();
    }
    lz2_r3 p = lz2_f6mul(f, f1, c.own, c.oth, c.im);
    fp t0 = lz2_ldo(t, 0, c), t1 = lz2_ldo(t, 1, c), t2 = lz2_ldo(t, 2, c), t2x = lz2_ldx(t, 2, c);
    lz2_sync();
    lz2_sto(f, 0, c, fpv_sub(fpv_sub(p.c[0], t0), lz2_xi(t2, t2x, c.im)));
    lz2_sto(f, 1, c, fpv_sub(fpv_sub(p.c[1], t1), t0));
    lz2_sto(f, 2, c, fpv_sub(fpv_sub(p.c[2], t2), t1));
    lz2_sto(f1, 0, c, fpv_add(t0, t0)); lz2_sto(f1, 1, c, fpv_add(t1, t1)); lz2_sto(f1, 2, c, fpv_add(t2, t2));
    lz2_sync();
}
// f *= 1 + (c3 + c4 v) w with c3, c4 in the two Fp2 slots at `l` (as lz_mul_nline); `t` = 6-slot temporary
LZ_FN2 void lz2_mul_nline(uint32_t f, uint32_t t, uint32_t l, Lz2 c) {
    const uint32_t f1 = f + 6 * LZ_SLOT;
    { lz2_r3 a = lz2_f6mul01(f1, l, c.own, c.oth, c.im); for (int k = 0; k < 3; k++) lz2_sto(t, k, c, a.c[k]); }
    lz2_r3 b = lz2_f6mul01(f, l, c.own, c.oth, c.im);
    lz2_sync();                                      // t is complete, and the partner is done reading f
    for (int k = 0; k < 3; k++) lz2_sto(f1, k, c, fpv_add(lz2_ldo(f1, k, c), b.c[k]));
    lz2_sto(f, 0, c, fpv_add(lz2_ldo(f, 0, c), lz2_xi(lz2_ldo(t, 2, c), lz2_ldx(t, 2, c), c.im)));
    lz2_sto(f, 1, c, fpv_add(lz2_ldo(f, 1, c), lz2_ldo(t, 0, c)));
    lz2_sto(f, 2, c, fpv_add(lz2_ldo(f, 2, c), lz2_ldo(t, 1, c)));
    lz2_sync();
}
#endif

}  // namespace zkv
