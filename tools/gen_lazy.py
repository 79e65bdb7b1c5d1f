#!/usr/bin/env python3
"""Lazy-reduction compiler for the BN254 tower kernels (writes csrc/lazy_gen.cuh).

The Fp6-level routines of the Miller loop and of the final exponentiation keep the 512-bit products of an Fp6 / sparse
multiplication unreduced, combine them there (Karatsuba terms, the multiplication by xi = 9 + u) and run ONE Montgomery
reduction per output coefficient (Aranha, Karabina, Longa, Gebotys, Lopez: "Faster explicit formulas for computing pairings
over ordinary curves", section 3-5).  An Fp6 multiplication then costs 18 x 64 + 6 x 72 = 1584 IMAD.WIDE instead of
6 x 336 = 2016.

What makes that safe is bookkeeping nobody should do by hand: every wide value must stay inside [0, 2^512), every
subtraction must stay non-negative (offsets that are multiples of p are added where it could not), every reduction input
must be below 4 p 2^256.  This script is a small compiler that does the bookkeeping:

  * a routine is written once against the `Gen` API below (ld / add8 / mulw / subw / mul_xi / redc / ...);
  * every value carries (a) its exact expression as an integer linear combination of products of the routine's inputs --
    so cancellations like (x0+x1)(y0+y1) - x0y0 - x1y1 = x0y1 + x1y0 >= 0 are KNOWN, not estimated -- from which the tightest
    sound bounds follow, (b) the constant offset added so far, (c) concrete values for a set of random and extreme inputs;
  * `subw` adds the smallest sufficient multiple of p 2^224, `ensure` inserts conditional subtractions of k p 2^256, and
    every claim is asserted on the concrete values as well; results are compared with plain modular arithmetic;
  * the routine is emitted as straight-line C++ over the leaf primitives of csrc/fp_ptx.cuh (whose PTX instruction lists
    tools/gen_fp_ptx.py verifies by simulation).  tests/host_emu compiles the same C++ against portable leaves.

Usage: gen_lazy.py [--check]   (--check: run all bound / value assertions, write nothing)
"""
import os
import random
import sys

P = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
R = 1 << 256
RINV = pow(R, -1, P)
BW = P << 256            # "B": multiples of it leave a Montgomery reduction's result unchanged modulo p
LIM = 1 << 512
UNIT = P << 224          # granularity of the offsets (9 limbs starting at limb 7)
NCASE = 24
XI = (9, 1)


# ------------------------------------------------------------------------------------------------------ values
class Atom:
    def __init__(self, name, hi, vals):
        self.name, self.hi, self.vals = name, hi, vals      # value in [0, hi)


class Nv:  # (see R_ATOM below: the constant 2^256 as an atom, so that n * 2^256 has an exact linear form)
    """narrow value: 8 limbs, a non-negative integer combination of atoms"""
    def __init__(self, name, lin, vals, canon=False):
        self.name, self.lin, self.vals, self.canon = name, lin, vals, canon

    def hi(self):    # inclusive maximum
        return sum(c * (a.hi - 1) for a, c in self.lin.items())


class Wv:
    """wide value: 16 limbs; stored = sum(coeff * key) + off, keys are products of two narrow atoms or opaque wide atoms"""
    def __init__(self, name, lin, off, vals):
        self.name, self.lin, self.off, self.vals = name, lin, off, vals

    @staticmethod
    def kmax(k):
        return (k[0].hi - 1) * (k[1].hi - 1) if len(k) == 2 else k[0].hi - 1

    def lo(self):
        return self.off + sum(c * Wv.kmax(k) for k, c in self.lin.items() if c < 0)

    def hi(self):    # inclusive maximum of the stored value
        return self.off + sum(c * Wv.kmax(k) for k, c in self.lin.items() if c > 0)


R_ATOM = Atom("R", R + 1, [R] * NCASE)


def lin_add(a, b, sb=1):
    r = dict(a)
    for k, c in b.items():
        r[k] = r.get(k, 0) + sb * c
        if r[k] == 0:
            del r[k]
    return r


def limbs(x, n):
    return ", ".join("0x%08xu" % ((x >> (32 * i)) & 0xFFFFFFFF) for i in range(n))


class Gen:
    def __init__(self, name, seed=1):
        self.name = name
        self.lines = []
        self.nid = 0
        self.rnd = random.Random(0xB200 ^ seed)
        self.stats = {"mulw": 0, "redc": 0, "csubw": 0, "csub8": 0, "addhi": 0, "wide_addsub": 0, "narrow": 0}

    # ---- plumbing
    def _n(self, lin, vals, canon=False):
        self.nid += 1
        name = "n%d" % self.nid
        self.lines.append("uint32_t %s[8];" % name)
        for v in vals:
            assert 0 <= v < R, "narrow overflow in %s" % self.name
        return Nv(name, lin, vals, canon)

    def _w(self, lin, off, vals):
        self.nid += 1
        name = "w%d" % self.nid
        self.lines.append("uint32_t %s[16];" % name)
        w = Wv(name, lin, off, vals)
        assert w.lo() >= 0 and w.hi() < LIM, "wide bound violated in %s: [%d, %.3f B]" % (self.name, w.lo(), w.hi() / BW)
        for v in vals:
            assert w.lo() <= v <= w.hi(), "bound engine disagrees with a concrete value in %s" % self.name
        return w

    def atom_vals(self, hi):
        vals = [hi - 1, 0]
        while len(vals) < NCASE:
            m = self.rnd.random()
            vals.append(hi - 1 if m < 0.25 else 0 if m < 0.35 else 1 if m < 0.4 else self.rnd.randrange(hi))
        return vals

    def fresh_n(self, name, vals):
        a = Atom(name, P, vals)
        return a

    def emit(self, s):
        self.lines.append(s)

    # ---- narrow
    def input_n(self, cname):
        """a canonical narrow input that already lives in the C array `cname`"""
        a = Atom(cname, P, self.atom_vals(P))
        return Nv(cname, {a: 1}, list(a.vals), True)

    def ld(self, base, k):
        a = Atom("%s[%d]" % (base, k), P, self.atom_vals(P))
        v = self._n({a: 1}, list(a.vals), True)
        self.emit("lz_ld(%s, %s + %d * LZ_SLOT);" % (v.name, base, k))
        return v

    def st(self, base, k, v):
        assert v.canon
        self.emit("lz_st(%s + %d * LZ_SLOT, %s);" % (base, k, v.name))

    def add8(self, a, b):
        self.stats["narrow"] += 1
        v = self._n(lin_add(a.lin, b.lin), [x + y for x, y in zip(a.vals, b.vals)])
        assert v.hi() < R
        self.emit("lz_add8(%s, %s, %s);" % (v.name, a.name, b.name))
        return v

    def _modop(self, fn, pyf, *args):
        for a in args:
            assert a.canon
        self.stats["narrow"] += 3
        vals = [pyf(*[a.vals[i] for a in args]) % P for i in range(NCASE)]
        at = Atom("m", P, vals)
        v = self._n({at: 1}, vals, True)
        self.emit("%s(%s, %s);" % (fn, v.name, ", ".join(a.name for a in args)))
        return v

    def fpadd(self, a, b): return self._modop("fp_add_ptx", lambda x, y: x + y, a, b)
    def fpsub(self, a, b): return self._modop("fp_sub_ptx", lambda x, y: x - y, a, b)
    def fpdbl(self, a): return self._modop("fp_add_ptx", lambda x, y: x + y, a, a)

    def fpneg(self, a):
        assert a.canon
        vals = [(-x) % P for x in a.vals]
        at = Atom("m", P, vals)
        v = self._n({at: 1}, vals, True)
        self.emit("lz_fp_neg(%s, %s);" % (v.name, a.name))
        return v

    # ---- wide
    def mulw(self, a, b):
        self.stats["mulw"] += 1
        lin = {}
        for ka, ca in a.lin.items():
            for kb, cb in b.lin.items():
                key = (ka, kb) if id(ka) <= id(kb) else (kb, ka)
                lin[key] = lin.get(key, 0) + ca * cb
        w = self._w(lin, 0, [x * y for x, y in zip(a.vals, b.vals)])
        self.emit("lz_mulw(%s, %s, %s);" % (w.name, a.name, b.name))
        return w

    def hiw(self, n):
        """n * 2^256 as a VIRTUAL wide value (never materialised): addw / subw add or subtract it on the upper eight limbs only.
        With n a Montgomery residue x R this is x R^2, the scale of the products, so redc(T -+ hiw(n)) = redc(T) -+ n (mod p)."""
        lin = {}
        for ka, ca in n.lin.items():
            lin[(ka, R_ATOM) if id(ka) <= id(R_ATOM) else (R_ATOM, ka)] = ca
        w = Wv("hi:" + n.name, lin, 0, [v << 256 for v in n.vals])
        w.hi_src = n.name
        assert w.hi() < LIM
        return w

    def csub(self, x, k):
        assert not getattr(x, "hi_src", None), "a virtual n * 2^256 value cannot be reduced"
        self.stats["csubw"] += 1
        assert 1 <= k <= 4 and x.lo() >= 0
        hi = max(k * BW - 1, x.hi() - k * BW)
        vals = [v - k * BW if v >= k * BW else v for v in x.vals]
        self.nid += 1
        at = Atom("cs%d" % self.nid, hi + 1, vals)
        w = self._w({(at,): 1}, 0, vals)
        self.emit("{ const uint32_t k_[8] = {%s}; lz_csubw(%s, %s, k_); }" % (limbs(k * P, 8), w.name, x.name))
        return w

    def ensure(self, x, limit):
        """conditional subtractions of k p 2^256 until the stored value is known to be <= limit"""
        while x.hi() > limit:
            k = 4
            while k > 1 and k * BW > x.hi() - limit and k * BW - 1 > limit:
                k //= 2
            while k * BW > x.hi():
                k //= 2
            assert k >= 1, "cannot reduce below the limit"
            x = self.csub(x, k)
        return x

    def addoff(self, x, off):
        assert off % UNIT == 0 and off > 0 and not getattr(x, "hi_src", None)
        self.stats["addhi"] += 1
        x = self.ensure(x, LIM - 1 - off)
        w = self._w(x.lin, x.off + off, [v + off for v in x.vals])
        self.emit("{ const uint32_t c_[9] = {%s}; lz_addhi(%s, %s, c_); }" % (limbs(off >> 224, 9), w.name, x.name))
        return w

    def addw(self, a, b):
        self.stats["wide_addsub"] += 1
        if getattr(a, "hi_src", None):
            a, b = b, a
        assert not getattr(a, "hi_src", None)
        vb = getattr(b, "hi_src", None)
        while a.hi() + b.hi() >= LIM:
            if a.hi() >= b.hi() or vb:
                a = self.ensure(a, max(BW - 1, a.hi() // 2))
            else:
                b = self.ensure(b, max(BW - 1, b.hi() // 2))
        w = self._w(lin_add(a.lin, b.lin), a.off + b.off, [x + y for x, y in zip(a.vals, b.vals)])
        if vb:
            self.emit("lz_addw_hi(%s, %s, %s);" % (w.name, a.name, vb))
        else:
            self.emit("lz_addw(%s, %s, %s);" % (w.name, a.name, b.name))
        return w

    def subw(self, a, b):
        """a - b, preceded by the addition of the smallest multiple of p 2^224 that keeps the difference non-negative"""
        self.stats["wide_addsub"] += 1
        while True:
            lin = lin_add(a.lin, b.lin, -1)
            probe = Wv("", lin, a.off - b.off, [])
            if probe.lo() >= 0:
                break
            need = -probe.lo()
            a = self.addoff(a, -(-need // UNIT) * UNIT)      # (may first reduce a, which forgets its expression: hence the loop)
        w = self._w(lin, a.off - b.off, [x - y for x, y in zip(a.vals, b.vals)])
        assert not getattr(a, "hi_src", None)
        if getattr(b, "hi_src", None):
            self.emit("lz_subw_hi(%s, %s, %s);" % (w.name, a.name, b.hi_src))
        else:
            self.emit("lz_subw(%s, %s, %s);" % (w.name, a.name, b.name))
        return w

    def dblw(self, a):
        a = self.ensure(a, LIM // 2 - 1)
        self.stats["wide_addsub"] += 1
        w = self._w({k: 2 * c for k, c in a.lin.items()}, 2 * a.off, [2 * v for v in a.vals])
        self.emit("lz_addw(%s, %s, %s);" % (w.name, a.name, a.name))
        return w

    def shl3w(self, a):
        """8 a by sixteen independent funnel shifts (no carry chain)"""
        a = self.ensure(a, LIM // 8 - 1)
        self.stats["wide_addsub"] += 1
        w = self._w({k: 8 * c for k, c in a.lin.items()}, 8 * a.off, [8 * v for v in a.vals])
        self.emit("lz_shl3w(%s, %s);" % (w.name, a.name))
        return w

    def mul9(self, x):
        if x.hi() < LIM // 8:                      # 8 x fits: sixteen independent funnel shifts
            return self.addw(self.shl3w(x), x)
        return self.addw(self.dblw(self.dblw(self.dblw(x))), x)      # else three doublings, conditionally reduced on the way

    def mul_xi(self, re, im):
        """(re + im u)(9 + u) = (9 re - im) + (9 im + re) u"""
        r9, i9 = self.mul9(re), self.mul9(im)
        return self.subw(r9, im), self.addw(i9, re)

    def redc(self, x):
        """Montgomery reduction to a canonical narrow value"""
        self.stats["redc"] += 1
        x = self.ensure(x, 4 * BW - 1)
        vals = [(v * RINV) % P for v in x.vals]
        self.nid += 1
        name = "n%d" % self.nid
        self.lines.append("uint32_t %s[8];" % name)
        self.emit("lz_redc(%s, %s);" % (name, x.name))
        hi = (x.hi() >> 256) + P          # inclusive bound of the unsubtracted result
        assert hi < R
        k = 4
        while k >= 1:
            if hi >= k * P:
                self.stats["csub8"] += 1
                self.emit("{ const uint32_t k_[8] = {%s}; lz_csub8(%s, %s, k_); }" % (limbs(k * P, 8), name, name))
                hi = max(k * P - 1, hi - k * P)
            k //= 2
        assert hi < P
        at = Atom("r%d" % self.nid, P, vals)
        return Nv(name, {at: 1}, vals, True)

    def body(self):
        return "\n".join("    " + l for l in self.lines)


# ------------------------------------------------------------------------------------------------------ reference arithmetic (per test case)
def f2_mul_ref(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def f2_add_ref(a, b, s=1):
    return ((a[0] + s * b[0]) % P, (a[1] + s * b[1]) % P)


def f2_xi_ref(a):
    return f2_mul_ref(a, XI)


def f6_mul_ref(a, b):
    m = f2_mul_ref
    ad = f2_add_ref
    c0 = ad(m(a[0], b[0]), f2_xi_ref(ad(m(a[1], b[2]), m(a[2], b[1]))))
    c1 = ad(ad(m(a[0], b[1]), m(a[1], b[0])), f2_xi_ref(m(a[2], b[2])))
    c2 = ad(ad(m(a[0], b[2]), m(a[1], b[1])), m(a[2], b[0]))
    return [c0, c1, c2]


def case(vs, i):
    """the i-th concrete value of a list of (re, im) Nv pairs"""
    return [(re.vals[i], im.vals[i]) for re, im in vs]


def expect(outs, ref):
    """outs: list of (re, im) Nv pairs; ref(i) -> list of (re, im) integer pairs WITHOUT the Montgomery factor of one reduction"""
    for i in range(NCASE):
        want = ref(i)
        for (re, im), (wr, wi) in zip(outs, want):
            assert re.vals[i] == wr * RINV % P and im.vals[i] == wi * RINV % P, "value mismatch"
            assert re.canon and im.canon


# ------------------------------------------------------------------------------------------------------ building blocks
def f2_mulw(g, X, Y):
    """wide product of two Fp2 operands (components may be unreduced sums): 3 mulw"""
    sx, sy = g.add8(X[0], X[1]), g.add8(Y[0], Y[1])
    t2 = g.mulw(sx, sy)
    t0 = g.mulw(X[0], Y[0])
    im = g.subw(t2, t0)
    t1 = g.mulw(X[1], Y[1])
    im = g.subw(im, t1)
    re = g.subw(t0, t1)
    return re, im


def f2_add8(g, X, Y):
    return g.add8(X[0], Y[0]), g.add8(X[1], Y[1])


def w2_add(g, a, b):
    return g.addw(a[0], b[0]), g.addw(a[1], b[1])


def w2_sub(g, a, b):
    return g.subw(a[0], b[0]), g.subw(a[1], b[1])


def w2_redc(g, a):
    return g.redc(a[0]), g.redc(a[1])


def ld2(g, base, i):
    return g.ld(base, 2 * i), g.ld(base, 2 * i + 1)


# ------------------------------------------------------------------------------------------------------ routines
def gen_f6mul():
    """fp6 = a * b for two Fp6 operands in shared-memory slots: 18 mulw + 6 redc"""
    g = Gen("lz_f6mul")
    A = [None] * 3
    Bv = [None] * 3
    A[0], Bv[0] = ld2(g, "a", 0), ld2(g, "b", 0)
    v0 = f2_mulw(g, A[0], Bv[0])
    A[1], Bv[1] = ld2(g, "a", 1), ld2(g, "b", 1)
    v1 = f2_mulw(g, A[1], Bv[1])
    m01 = f2_mulw(g, f2_add8(g, A[0], A[1]), f2_add8(g, Bv[0], Bv[1]))
    c1 = w2_sub(g, w2_sub(g, m01, v0), v1)
    A[2], Bv[2] = ld2(g, "a", 2), ld2(g, "b", 2)
    v2 = f2_mulw(g, A[2], Bv[2])
    c1 = w2_add(g, c1, g.mul_xi(*v2))
    o1 = w2_redc(g, c1)
    m02 = f2_mulw(g, f2_add8(g, A[0], A[2]), f2_add8(g, Bv[0], Bv[2]))
    c2 = w2_add(g, w2_sub(g, w2_sub(g, m02, v0), v2), v1)
    o2 = w2_redc(g, c2)
    m12 = f2_mulw(g, f2_add8(g, A[1], A[2]), f2_add8(g, Bv[1], Bv[2]))
    s = w2_sub(g, w2_sub(g, m12, v1), v2)
    c0 = w2_add(g, g.mul_xi(*s), v0)
    o0 = w2_redc(g, c0)
    outs = [o0, o1, o2]
    expect(outs, lambda i: f6_mul_ref(case(A, i), case(Bv, i)))
    return g, outs


def gen_f6mul01():
    """fp6 = a * (b0 + b1 v): Fp6 operand a and the two Fp2 coefficients b0, b1 in shared-memory slots: 15 mulw + 6 redc"""
    g = Gen("lz_f6mul01", 2)
    A = [ld2(g, "a", 0), ld2(g, "a", 1), None]
    Bv = [ld2(g, "b", 0), ld2(g, "b", 1)]
    v0 = f2_mulw(g, A[0], Bv[0])
    v1 = f2_mulw(g, A[1], Bv[1])
    m01 = f2_mulw(g, f2_add8(g, A[0], A[1]), f2_add8(g, Bv[0], Bv[1]))
    o1 = w2_redc(g, w2_sub(g, w2_sub(g, m01, v0), v1))                  # c1 = a0 b1 + a1 b0
    A[2] = ld2(g, "a", 2)
    t = f2_mulw(g, A[2], Bv[0])
    o2 = w2_redc(g, w2_add(g, t, v1))                                   # c2 = a1 b1 + a2 b0
    t = f2_mulw(g, A[2], Bv[1])
    o0 = w2_redc(g, w2_add(g, g.mul_xi(*t), v0))                        # c0 = a0 b0 + xi a2 b1
    outs = [o0, o1, o2]
    z = (0, 0)
    expect(outs, lambda i: f6_mul_ref(case(A, i), case(Bv, i) + [z]))
    return g, outs


def gen_f4sqr():
    """(t0, t1) = (a^2 + xi b^2, 2 a b) for a + b s in Fp4 = Fp2[s]/(s^2 - xi), a and b in shared-memory slots: 6 mulw + 4 redc"""
    g = Gen("lz_f4sqr", 3)
    a, b = ld2(g, "a", 0), ld2(g, "b", 0)

    def sqrw(x):      # (x0 + x1 u)^2 = (x0 + x1)(x0 - x1) + 2 x0 x1 u, the difference taken modulo p
        re = g.mulw(g.add8(x[0], x[1]), g.fpsub(x[0], x[1]))
        im = g.mulw(g.add8(x[0], x[0]), x[1])
        return re, im
    a2, b2 = sqrw(a), sqrw(b)
    o0 = w2_redc(g, w2_add(g, g.mul_xi(*b2), a2))
    s = (g.fpadd(a[0], b[0]), g.fpadd(a[1], b[1]))
    s2 = sqrw(s)
    o1 = w2_redc(g, w2_sub(g, w2_sub(g, s2, a2), b2))
    sq = lambda x: f2_mul_ref(x, x)
    expect([o0, o1], lambda i: [f2_add_ref(sq(case([a], i)[0]), f2_xi_ref(sq(case([b], i)[0]))),
                                f2_mul_ref(f2_add_ref(case([a], i)[0], case([a], i)[0]), case([b], i)[0])])
    return g, [o0, o1]


def st2(g, base, i, v):
    g.st(base, 2 * i, v[0]); g.st(base, 2 * i + 1, v[1])


def w2_addhi(g, w, n, sign=1):
    """w +- n * 2^256 for an Fp2 pair of wide values w and an Fp2 pair of narrow values n"""
    op = g.addw if sign > 0 else g.subw
    return op(w[0], g.hiw(n[0])), op(w[1], g.hiw(n[1]))


def gen_f6mul01_acc(kind):
    g = Gen("lz_f6mul01_" + kind, 5 if kind == "add" else 6)
    A = [ld2(g, "a", 0), ld2(g, "a", 1), None]
    Bv = [ld2(g, "b", 0), ld2(g, "b", 1)]
    Cv = [None] * 3
    outs = [None] * 3
    tgt = {"add": (1, 2, 0), "vadd": (2, 0, 1)}[kind]            # where (ab).c1, (ab).c2, (ab).c0 go: o = c + ab, or o = c + v ab
    v0 = f2_mulw(g, A[0], Bv[0])
    v1 = f2_mulw(g, A[1], Bv[1])
    m01 = f2_mulw(g, f2_add8(g, A[0], A[1]), f2_add8(g, Bv[0], Bv[1]))
    x = w2_sub(g, w2_sub(g, m01, v0), v1)                               # (ab).c1 = a0 b1 + a1 b0
    k = tgt[0]; Cv[k] = ld2(g, "c", k); outs[k] = w2_redc(g, w2_addhi(g, x, Cv[k])); st2(g, "o", k, outs[k])
    A[2] = ld2(g, "a", 2)
    t = f2_mulw(g, A[2], Bv[0])
    x = w2_add(g, t, v1)                                                # (ab).c2 = a1 b1 + a2 b0
    if kind == "vadd":
        x = g.mul_xi(*x)
    k = tgt[1]; Cv[k] = ld2(g, "c", k); outs[k] = w2_redc(g, w2_addhi(g, x, Cv[k])); st2(g, "o", k, outs[k])
    t = f2_mulw(g, A[2], Bv[1])
    x = w2_add(g, g.mul_xi(*t), v0)                                     # (ab).c0 = a0 b0 + xi a2 b1
    k = tgt[2]; Cv[k] = ld2(g, "c", k); outs[k] = w2_redc(g, w2_addhi(g, x, Cv[k])); st2(g, "o", k, outs[k])
    z = (0, 0)

    def ref(i):
        ab = f6_mul_ref(case(A, i), case(Bv, i) + [z])
        if kind == "vadd":
            ab = [f2_xi_ref(ab[2]), ab[0], ab[1]]
        c = case(Cv, i)
        return [((ab[j][0] + c[j][0] * R) % P, (ab[j][1] + c[j][1] * R) % P) for j in range(3)]
    expect(outs, ref)
    return g, []


def gen_f6mul01_add():
    """o = c + a * (b0 + b1 v): Fp6 a, c and the Fp2 pair b in shared-memory slots, result stored to the slots at o (o may be c): 15 mulw + 6 redc"""
    return gen_f6mul01_acc("add")


def gen_f6mul01_vadd():
    """o = c + v * a * (b0 + b1 v), as above (o must not overlap a, b or c): 15 mulw + 6 redc"""
    return gen_f6mul01_acc("vadd")


def gen_gs(xi):
    g = Gen("lz_gs%d" % xi, 7 + xi)
    a, b = ld2(g, "a", 0), ld2(g, "b", 0)

    def sqrw(x):
        re = g.mulw(g.add8(x[0], x[1]), g.fpsub(x[0], x[1]))
        im = g.mulw(g.add8(x[0], x[0]), x[1])
        return re, im

    def tri(w):
        return g.addw(g.dblw(w), w)
    a2, b2 = sqrw(a), sqrw(b)
    t0 = w2_add(g, g.mul_xi(*b2), a2)                                   # a^2 + xi b^2
    ga = ld2(g, "ga", 0)
    oa = w2_redc(g, w2_addhi(g, (tri(t0[0]), tri(t0[1])), f2_add8(g, ga, ga), -1))
    s = (g.fpadd(a[0], b[0]), g.fpadd(a[1], b[1]))
    s2 = sqrw(s)
    t1 = w2_sub(g, w2_sub(g, s2, a2), b2)                               # 2 a b
    if xi:
        t1 = g.mul_xi(*t1)
    gb = ld2(g, "gb", 0)
    ob = w2_redc(g, w2_addhi(g, (tri(t1[0]), tri(t1[1])), f2_add8(g, gb, gb)))
    st2(g, "oa", 0, oa); st2(g, "ob", 0, ob)
    sq = lambda x: f2_mul_ref(x, x)

    def ref(i):
        av, bv, gav, gbv = case([a], i)[0], case([b], i)[0], case([ga], i)[0], case([gb], i)[0]
        u0 = f2_add_ref(sq(av), f2_xi_ref(sq(bv)))
        u1 = f2_mul_ref(f2_add_ref(av, av), bv)
        if xi:
            u1 = f2_xi_ref(u1)
        return [tuple((3 * u0[j] - 2 * gav[j] * R) % P for j in range(2)), tuple((3 * u1[j] + 2 * gbv[j] * R) % P for j in range(2))]
    expect([oa, ob], ref)
    return g, []


def gen_gs0():
    """Granger-Scott step on one Fp4 pair: with (t0, t1) = (a^2 + xi b^2, 2 a b), slots oa <- 3 t0 - 2 ga, ob <- 3 t1 + 2 gb (stored at the end: oa, ob may be a, b, ga, gb): 6 mulw + 4 redc"""
    return gen_gs(0)


def gen_gs1():
    """as lz_gs0 with ob <- 3 xi t1 + 2 gb"""
    return gen_gs(1)


VSIG_ACC = "uint32_t o, uint32_t a, uint32_t b, uint32_t c"
VSIG_GS = "uint32_t a, uint32_t b, uint32_t ga, uint32_t gb, uint32_t oa, uint32_t ob"
ROUTINES = [("lz_f6mul", gen_f6mul, "uint32_t a, uint32_t b", "fp6"), ("lz_f6mul01", gen_f6mul01, "uint32_t a, uint32_t b", "fp6"),
            ("lz_f4sqr", gen_f4sqr, "uint32_t a, uint32_t b", "fp4")]
# Fused forms that were generated, bound-checked and COUNTED in round 2 but are not emitted (gen_lazy.py --experiments prints their
# statistics): folding the accumulator into the wide domain (o = c + a b, o = c + v a b through n * 2^256 terms) and the Granger-Scott
# combinations 3 t -+ 2 g push the unreduced values past 4 p 2^256, and the conditional subtractions that brings back (csubw: 9 for the
# v-form against 2, 16 - 30 for a Granger-Scott pair against 2) cost more instructions than the narrow modular additions they replace.
EXPERIMENTS = [("lz_f6mul01_add", gen_f6mul01_add, VSIG_ACC, None), ("lz_f6mul01_vadd", gen_f6mul01_vadd, VSIG_ACC, None),
               ("lz_gs0", gen_gs0, VSIG_GS, None), ("lz_gs1", gen_gs1, VSIG_GS, None)]

HEADER = """// GENERATED by tools/gen_lazy.py (bounds proved on exact linear forms, values checked on random and extreme inputs). Do not edit.
// Lazily reduced Fp6-level routines of the BN254 tower over operands in shared-memory slots: the 512-bit products of a routine are
// combined unreduced and every output coefficient is reduced once.  Leaves: csrc/fp_ptx.cuh (device) or tests/host_emu/lazy_leaf_host.h.
// This is the arithmetic behind the reference's ecPairing precompile call (/root/reference/contracts/src/common/groth16.rs:121-125).
#pragma once
"""


def render(verbose=False):
    out = [HEADER]
    for name, fn, sig, rtype in ROUTINES:
        g, outs = fn()
        if verbose:
            print("%-12s %s" % (name, g.stats))
        if rtype is None:
            out.append("// %s\nLZ_FN void %s(%s) {\n%s\n}\n" % (fn.__doc__.strip().split("\n")[0], name, sig, g.body()))
            continue
        ret = "\n".join("    for (int i_ = 0; i_ < 8; i_++) { r_.c%d.c0.v[i_] = %s[i_]; r_.c%d.c1.v[i_] = %s[i_]; }" % (k, re.name, k, im.name)
                        for k, (re, im) in enumerate(outs))
        out.append("// %s\nLZ_FN %s %s(%s) {\n%s\n    %s r_;\n%s\n    return r_;\n}\n" % (fn.__doc__.strip().split("\n")[0], rtype, name, sig, g.body(), rtype, ret))
    return "\n".join(out)


def main():
    text = render(verbose=True)
    if "--experiments" in sys.argv:
        for name, fn, sig, rtype in EXPERIMENTS:
            g, outs = fn()
            print("%-16s %s" % (name, g.stats))
    if "--check" in sys.argv:
        return
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, "stylus_zkvm_verifiers_b200/csrc/lazy_gen.cuh"), "w") as f:
        f.write(text)
    print("wrote lazy_gen.cuh")


if __name__ == "__main__":
    main()
