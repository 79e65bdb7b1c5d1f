#!/usr/bin/env python3
"""Generate (and verify by simulation) the inline-PTX BN254 Fp primitives for sm_100a.

Why generated: the multiplier is ~190 PTX instructions whose correctness hinges on carry-flag
plumbing.  This script builds each routine as an instruction list, *simulates that exact list*
(32-bit registers + CC.CF semantics of the PTX ISA) against Python big-int arithmetic on random
and edge inputs, and only then emits it as one `asm` statement.  `--check` runs the simulation
only (used by tests/test_fp_ptx_sim.py); default also writes csrc/fp_ptx.cuh.

Multiplier layout (8 x 32-bit limbs, Montgomery R = 2^256): products of even limbs of `a`
accumulate in E (E[k] at column k), products of odd limbs in O (O[k] at column k+1), so every
(mad.lo.cc, madc.hi.cc) pair works on one 64-bit product and ptxas fuses the pair into a single
IMAD.WIDE.U32[.X] with the carry in a predicate: 8 rows x (8 a*b + 1 m + 8 m*p) = 136 IMAD.
After each row's reduction E[0]==0; the frame shifts by one limb by *renaming* (new E = old O,
new O = old E >> 2 limbs) plus one add of old E[1] into new E[0].
"""
import os
import random
import sys

P = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
N = 8
MASK = 0xFFFFFFFF
M0 = (-pow(P, -1, 1 << 32)) & MASK
assert M0 == 0xE4866389
PL = [(P >> (32 * i)) & MASK for i in range(N)]
RMONT = 1 << 256
GRAIN_MUL = int(os.environ.get("ZKV_GRAIN_MUL", "1"))      # carry chains emitted per stream per round of interleave()
GRAIN_REDC = int(os.environ.get("ZKV_GRAIN_REDC", "1"))


class Prog:
    def __init__(self, name, ins, outs):
        self.name, self.ins, self.outs = name, ins, outs
        self.code = []
        self.ntmp = 0
        self.npred = 0

    def tmp(self):
        self.ntmp += 1
        return "t%d" % (self.ntmp - 1)

    def emit(self, op, dst, *src):
        self.code.append((op, dst) + src)
        return dst

    def op(self, op, *src):
        return self.emit(op, self.tmp(), *src)

    # ---- simulator -------------------------------------------------------------------
    def run(self, inputs):
        reg = dict(inputs)
        cf = 0
        val = lambda x: x if isinstance(x, int) else reg[x]
        for ins in self.code:
            op, dst, src = ins[0], ins[1], [val(s) for s in ins[2:]]
            if op == "assert_nc":
                assert cf == 0, "dropped carry in %s" % self.name
                continue
            if op == "assert_zero":
                assert src[0] == 0, "overflow beyond the top limb in %s" % self.name
                continue
            if op in ("mul.lo", "mul.hi"):
                pr = src[0] * src[1]
                reg[dst] = (pr & MASK) if op == "mul.lo" else (pr >> 32)
                continue
            if op.startswith("mad"):
                pr = src[0] * src[1]
                part = (pr & MASK) if ".lo" in op else (pr >> 32)
                t = part + src[2] + (cf if op.startswith("madc") else 0)
                reg[dst] = t & MASK
                if op.endswith(".cc"):
                    cf = t >> 32
                continue
            if op.startswith("add"):
                t = src[0] + src[1] + (cf if op.startswith("addc") else 0)
                reg[dst] = t & MASK
                if op.endswith(".cc"):
                    cf = t >> 32
                continue
            if op.startswith("sub"):
                # PTX: sub.cc writes CC.CF = borrow-out ; subc consumes it as borrow-in
                t = src[0] - src[1] - (cf if op.startswith("subc") else 0)
                reg[dst] = t & MASK
                if op.endswith(".cc"):
                    cf = 1 if t < 0 else 0
                continue
            if op == "selp_eqz":      # dst = (src2 == 0) ? src0 : src1
                reg[dst] = src[0] if src[2] == 0 else src[1]
                continue
            if op == "and":
                reg[dst] = src[0] & src[1]
                continue
            if op == "shf.l":         # funnel shift left: the upper 32 bits of (hi:lo) << n  (src = lo, hi, n)
                reg[dst] = (((src[1] << 32) | src[0]) << src[2] >> 32) & MASK
                continue
            if op == "shl":
                reg[dst] = (src[0] << src[1]) & MASK
                continue
            if op == "mov":
                reg[dst] = src[0]
                continue
            raise ValueError(op)
        return [reg[o] for o in self.outs]

    # ---- emitter ---------------------------------------------------------------------
    def cuda(self):
        opnum = {}
        for i, o in enumerate(self.outs):
            opnum[o] = "%%%d" % i
        for i, o in enumerate(self.ins):
            opnum[o] = "%%%d" % (i + len(self.outs))

        def fmt(x):
            if isinstance(x, int):
                return "0x%08x" % x
            return opnum.get(x, x)

        lines = []
        if self.ntmp:
            lines.append(".reg .u32 t<%d>;" % self.ntmp)
        lines.append(".reg .pred q;")
        for ins in self.code:
            op, dst, src = ins[0], ins[1], ins[2:]
            if op in ("assert_nc", "assert_zero"):
                continue
            if op == "selp_eqz":
                lines.append("setp.eq.u32 q, %s, 0;" % fmt(src[2]))
                lines.append("selp.u32 %s, %s, %s, q;" % (fmt(dst), fmt(src[0]), fmt(src[1])))
                continue
            pop = {"and": "and.b32", "mov": "mov.u32", "shf.l": "shf.l.wrap.b32", "shl": "shl.b32"}.get(op, op + ".u32")
            lines.append("%s %s, %s;" % (pop, fmt(dst), ", ".join(fmt(s) for s in src)))
        body = "\n".join('        "%s\\n\\t"' % l for l in lines)
        outs = ", ".join('"=r"(%s)' % self._c(o) for o in self.outs)
        ins = ", ".join('"r"(%s)' % self._c(i) for i in self.ins)
        return '    asm("{\\n\\t"\n%s\n        "}"\n        : %s\n        : %s);' % (body, outs, ins)

    @staticmethod
    def _c(name):
        return "%s[%s]" % (name[0], name[1:])

    def count(self, prefix):
        return sum(1 for c in self.code if c[0].startswith(prefix))


def final_sub(p, r, outs, bound2p=True):
    """outs = r - P if r >= P else r  (r < 2P)."""
    s = [p.op("sub.cc" if k == 0 else "subc.cc", r[k], PL[k]) for k in range(N)]
    bor = p.op("subc", 0, 0)          # 0 or 0xffffffff
    for k in range(N):
        p.emit("selp_eqz", outs[k], s[k], r[k], bor)


def reduce_row(p, E, O, carry_in=False):
    m = p.op("mul.lo", E[0], M0)
    for idx, j in enumerate((1, 3, 5, 7)):
        O[j - 1] = p.op("mad.lo.cc" if (idx == 0 and not carry_in) else "madc.lo.cc", m, PL[j], O[j - 1])
        O[j] = p.op("madc.hi.cc", m, PL[j], O[j])
    p.emit("assert_nc", None)
    for idx, j in enumerate((0, 2, 4, 6)):
        E[j] = p.op("mad.lo.cc" if idx == 0 else "madc.lo.cc", m, PL[j], E[j])
        E[j + 1] = p.op("madc.hi.cc", m, PL[j], E[j + 1])
    O[7] = p.op("addc", O[7], 0)


def gen_mul(name="fp_mul", reduce_final=True, sqr=False):
    ins = ["a%d" % i for i in range(N)] + ([] if sqr else ["b%d" % i for i in range(N)])
    outs = ["r%d" % i for i in range(N)]
    p = Prog(name, ins, outs)
    a = ["a%d" % i for i in range(N)]
    b = a if sqr else ["b%d" % i for i in range(N)]
    E, O = [None] * N, [None] * N
    for j in (0, 2, 4, 6):
        E[j] = p.op("mul.lo", a[j], b[0]); E[j + 1] = p.op("mul.hi", a[j], b[0])
    for j in (1, 3, 5, 7):
        O[j - 1] = p.op("mul.lo", a[j], b[0]); O[j] = p.op("mul.hi", a[j], b[0])
    reduce_row(p, E, O)
    for i in range(1, N):
        X = E[1]
        Oin = E[2:] + [0, 0]
        E = O
        O = [None] * N
        E[0] = p.op("add.cc", E[0], X)
        for j in (1, 3, 5, 7):
            O[j - 1] = p.op("madc.lo.cc", a[j], b[i], Oin[j - 1])
            O[j] = p.op("madc.hi.cc", a[j], b[i], Oin[j])
        p.emit("assert_nc", None)
        for idx, j in enumerate((0, 2, 4, 6)):
            E[j] = p.op("mad.lo.cc" if idx == 0 else "madc.lo.cc", a[j], b[i], E[j])
            E[j + 1] = p.op("madc.hi.cc", a[j], b[i], E[j + 1])
        O[7] = p.op("addc", O[7], 0)
        reduce_row(p, E, O)
    r = [None] * N
    r[0] = p.op("add.cc", E[1], O[0])
    for k in range(1, 7):
        r[k] = p.op("addc.cc", E[k + 1], O[k])
    r[7] = p.op("addc", O[7], 0)
    if reduce_final:
        final_sub(p, r, outs)
    else:
        for k in range(N):
            p.emit("mov", outs[k], r[k])
    return p


def gen_add():
    p = Prog("fp_add", ["a%d" % i for i in range(N)] + ["b%d" % i for i in range(N)], ["r%d" % i for i in range(N)])
    r = [p.op("add.cc" if k == 0 else ("addc.cc" if k < 7 else "addc"), "a%d" % k, "b%d" % k) for k in range(N)]
    final_sub(p, r, p.outs)
    return p


def gen_sub():
    p = Prog("fp_sub", ["a%d" % i for i in range(N)] + ["b%d" % i for i in range(N)], ["r%d" % i for i in range(N)])
    d = [p.op("sub.cc" if k == 0 else "subc.cc", "a%d" % k, "b%d" % k) for k in range(N)]
    bor = p.op("subc", 0, 0)
    pm = [p.op("and", bor, PL[k]) for k in range(N)]
    for k in range(N):
        p.emit("add.cc" if k == 0 else ("addc.cc" if k < 7 else "addc"), p.outs[k], d[k], pm[k])
    return p


# ---------------------------------------------------------------------------------------------------------------------
# Lazily reduced Fp2 arithmetic (Aranha et al., "Faster explicit formulas for computing pairings over ordinary curves",
# section 3): the three Karatsuba products of an Fp2 multiplication stay 512 bits wide, are combined there, and only the two
# result coefficients go through a Montgomery reduction: 3 x 64 + 2 x 72 = 336 IMAD.WIDE instead of 3 x 136 = 408.
def mul_wide(p, a, b):
    """16-limb product of two 8-limb operands (no reduction), 64 IMAD.WIDE; same even/odd accumulator framing as gen_mul."""
    out = []
    E, O = [None] * N, [None] * N
    for j in (0, 2, 4, 6):
        E[j] = p.op("mul.lo", a[j], b[0]); E[j + 1] = p.op("mul.hi", a[j], b[0])
    for j in (1, 3, 5, 7):
        O[j - 1] = p.op("mul.lo", a[j], b[0]); O[j] = p.op("mul.hi", a[j], b[0])
    out.append(E[0])
    for i in range(1, N):
        X = E[1]
        Oin = E[2:] + [0, 0]
        E = O
        O = [None] * N
        E[0] = p.op("add.cc", E[0], X)
        for j in (1, 3, 5, 7):
            O[j - 1] = p.op("madc.lo.cc", a[j], b[i], Oin[j - 1])
            O[j] = p.op("madc.hi.cc", a[j], b[i], Oin[j])
        p.emit("assert_nc", None)
        for idx, j in enumerate((0, 2, 4, 6)):
            E[j] = p.op("mad.lo.cc" if idx == 0 else "madc.lo.cc", a[j], b[i], E[j])
            E[j + 1] = p.op("madc.hi.cc", a[j], b[i], E[j + 1])
        O[7] = p.op("addc.cc", O[7], 0)
        p.emit("assert_nc", None)
        out.append(E[0])
    out.append(p.op("add.cc", E[1], O[0]))
    for k in range(1, 7):
        out.append(p.op("addc.cc", E[k + 1], O[k]))
    out.append(p.op("addc.cc", O[7], 0))
    p.emit("assert_nc", None)
    return out


def mul_wide2(p, a, b, c, d):
    """16-limb a b + c d for 8-limb operands whose product sum stays below 2^512 (asserted): 128 IMAD.WIDE and no wide addition.
    Same even / odd framing as mul_wide; at every offset the row of a and the row of c land on the same accumulators, and the carry
    out of limb offset + 8 is kept in TOP, which enters the odd accumulator of limb offset + 9 at the next slide."""
    out = []
    E, O = [None] * N, [None] * N
    for j in (0, 2, 4, 6):
        E[j] = p.op("mul.lo", a[j], b[0]); E[j + 1] = p.op("mul.hi", a[j], b[0])
    for j in (1, 3, 5, 7):
        O[j - 1] = p.op("mul.lo", a[j], b[0]); O[j] = p.op("mul.hi", a[j], b[0])
    top = 0
    for i in range(N):
        if i:
            X = E[1]
            Oin = E[2:] + [0, top]
            E = O
            O = [None] * N
            E[0] = p.op("add.cc", E[0], X)
            for j in (1, 3, 5, 7):
                O[j - 1] = p.op("madc.lo.cc", a[j], b[i], Oin[j - 1])
                O[j] = p.op("madc.hi.cc", a[j], b[i], Oin[j])
            top = p.op("addc", 0, 0)
            for idx, j in enumerate((0, 2, 4, 6)):
                E[j] = p.op("mad.lo.cc" if idx == 0 else "madc.lo.cc", a[j], b[i], E[j])
                E[j + 1] = p.op("madc.hi.cc", a[j], b[i], E[j + 1])
            O[7] = p.op("addc.cc", O[7], 0)
            top = p.op("addc", top, 0)
        for idx, j in enumerate((1, 3, 5, 7)):
            O[j - 1] = p.op("mad.lo.cc" if idx == 0 else "madc.lo.cc", c[j], d[i], O[j - 1])
            O[j] = p.op("madc.hi.cc", c[j], d[i], O[j])
        top = p.op("addc", top, 0)
        for idx, j in enumerate((0, 2, 4, 6)):
            E[j] = p.op("mad.lo.cc" if idx == 0 else "madc.lo.cc", c[j], d[i], E[j])
            E[j + 1] = p.op("madc.hi.cc", c[j], d[i], E[j + 1])
        O[7] = p.op("addc.cc", O[7], 0)
        top = p.op("addc", top, 0)
        out.append(E[0])
    out.append(p.op("add.cc", E[1], O[0]))
    for k in range(1, 7):
        out.append(p.op("addc.cc", E[k + 1], O[k]))
    out.append(p.op("addc.cc", O[7], 0))
    p.emit("assert_nc", None)
    p.emit("assert_zero", None, top)
    return out


def redc_wide(p, T, outs):
    """outs = T / 2^256 mod P, fully reduced, for a 16-limb T < P * 2^256: 72 IMAD.WIDE."""
    E, O = list(T[:N]), [0] * N
    for i in range(N):
        reduce_row(p, E, O, carry_in=(i > 0))
        if i < N - 1:
            X = E[1]
            Oin = E[2:] + [0, 0]
            E = O
            O = Oin
            E[0] = p.op("add.cc", E[0], X)        # carry consumed by the first madc of the next reduce_row
    u = [p.op("add.cc", E[1], O[0])]
    for k in range(1, 7):
        u.append(p.op("addc.cc", E[k + 1], O[k]))
    u.append(p.op("addc.cc", O[7], 0))
    p.emit("assert_nc", None)
    r = [p.op("add.cc" if k == 0 else "addc.cc", u[k], T[N + k]) for k in range(N)]
    p.emit("assert_nc", None)
    final_sub(p, r, outs)


def add_raw(p, a, b, n=N):
    r = [p.op("add.cc" if k == 0 else "addc.cc", a[k], b[k]) for k in range(n)]
    p.emit("assert_nc", None)
    return r


def sub_mod(p, a, b):
    d = [p.op("sub.cc" if k == 0 else "subc.cc", a[k], b[k]) for k in range(N)]
    bor = p.op("subc", 0, 0)
    pm = [p.op("and", bor, PL[k]) for k in range(N)]
    return [p.op("add.cc" if k == 0 else ("addc.cc" if k < 7 else "addc"), d[k], pm[k]) for k in range(N)]


def _reads_cf(op):
    return op.startswith(("madc", "addc", "subc"))


def _writes_cf(op):
    return op.endswith(".cc")


def interleave(p, builders, grain=1):
    """Run each builder (it emits into p), then merge the instruction streams round-robin at carry-chain granularity (a cut is only
    made where CC.CF is dead), `grain` chains at a time.  The streams must be independent of each other.  ptxas keeps the source order
    as its tie-break, so this is what puts several independent IMAD.WIDE.X carry chains in flight per warp: one chain alone issues
    an instruction per carry-predicate latency, not per pipe slot."""
    results, streams = [], []
    for b in builders:
        start = len(p.code)
        results.append(b())
        streams.append(p.code[start:])
        del p.code[start:]
    segs = []
    for code in streams:
        live = [False] * (len(code) + 1)
        for k in range(len(code) - 1, -1, -1):
            op = code[k][0]
            if op == "assert_nc":
                live[k] = live[k + 1]
            elif _reads_cf(op):
                live[k] = True
            elif _writes_cf(op):
                live[k] = False
            else:
                live[k] = live[k + 1]
        out, cur = [], []
        for k, ins in enumerate(code):
            if cur and not live[k] and ins[0] != "assert_nc":
                out.append(cur); cur = []
            cur.append(ins)
        if cur:
            out.append(cur)
        segs.append(out)
    pos = [0] * len(segs)
    while any(pos[i] < len(segs[i]) for i in range(len(segs))):
        for i in range(len(segs)):
            for _ in range(grain):
                if pos[i] < len(segs[i]):
                    p.code.extend(segs[i][pos[i]]); pos[i] += 1
    return results


def gen_fp2_mul():
    """(r, q) = (x + y u)(z + w u) in Fp[u]/(u^2+1):  r = xz - yw,  q = (x+y)(z+w) - xz - yw."""
    nm = lambda c: ["%s%d" % (c, i) for i in range(N)]
    x, y, z, w = nm("x"), nm("y"), nm("z"), nm("w")
    p = Prog("fp2_mul", x + y + z + w, nm("r") + nm("q"))
    s = add_raw(p, x, y)                    # < 2P < 2^255
    t = add_raw(p, z, w)
    T0, T1, T2 = interleave(p, [lambda: mul_wide(p, x, z), lambda: mul_wide(p, y, w), lambda: mul_wide(p, s, t)], GRAIN_MUL)   # T2 < 4 P^2 < 2^510
    # C0 = T0 - T1 (+ P * 2^256 if negative)  in [0, P * 2^256)
    d = [p.op("sub.cc" if k == 0 else "subc.cc", T0[k], T1[k]) for k in range(2 * N)]
    bor = p.op("subc", 0, 0)
    pm = [p.op("and", bor, PL[k]) for k in range(N)]
    C0 = d[:N] + [p.op("add.cc" if k == 0 else ("addc.cc" if k < 7 else "addc"), d[N + k], pm[k]) for k in range(N)]
    S = add_raw(p, T0, T1, 2 * N)           # < 2 P^2
    C1 = [p.op("sub.cc" if k == 0 else "subc.cc", T2[k], S[k]) for k in range(2 * N)]
    p.emit("assert_nc", None)               # xw + yz >= 0
    interleave(p, [lambda: redc_wide(p, C0, nm("r")), lambda: redc_wide(p, C1, nm("q"))], GRAIN_REDC)
    return p


def gen_fp2_sqr():
    """(r, q) = (x + y u)^2:  r = (x+y)(x-y),  q = (2x) y.  Operands of the two products are < 2P, so the unreduced sum / double is enough."""
    nm = lambda c: ["%s%d" % (c, i) for i in range(N)]
    x, y = nm("x"), nm("y")
    p = Prog("fp2_sqr", x + y, nm("r") + nm("q"))
    s = add_raw(p, x, y)
    d = sub_mod(p, x, y)
    m = add_raw(p, x, x)
    A, B = interleave(p, [lambda: mul_wide(p, s, d), lambda: mul_wide(p, m, y)], GRAIN_MUL)
    interleave(p, [lambda: redc_wide(p, A, nm("r")), lambda: redc_wide(p, B, nm("q"))], GRAIN_REDC)
    return p



# ---------------------------------------------------------------------------------------------------------------------
# Leaf primitives of the lazily reduced tower (csrc/lazy_gen.cuh, generated by tools/gen_lazy.py, composes them):
# wide (16-limb) products, raw wide / narrow carry chains, additions of constants that are multiples of p (offsets that
# keep a difference non-negative), conditional subtractions of multiples of p * 2^256, and a Montgomery reduction
# without the final subtraction.  Each is an instruction list simulated against Python integers like the rest.
def nm(c, n=N):
    return ["%s%d" % (c, i) for i in range(n)]


def gen_mulw():
    p = Prog("lz_mulw", nm("a") + nm("b"), nm("w", 16))
    out = mul_wide(p, nm("a"), nm("b"))
    for k in range(16):
        p.emit("mov", "w%d" % k, out[k])
    return p


def gen_mulw2():
    p = Prog("lz_mulw2", nm("a") + nm("b") + nm("c") + nm("d"), nm("w", 16))
    out = mul_wide2(p, nm("a"), nm("b"), nm("c"), nm("d"))
    for k in range(16):
        p.emit("mov", "w%d" % k, out[k])
    return p


def gen_redc_nf():
    """r = (T + m p) / 2^256 for a 16-limb T < 4 p 2^256, WITHOUT the final subtraction: r < T / 2^256 + p < 5 p < 2^256."""
    p = Prog("lz_redc", nm("w", 16), nm("r"))
    T = nm("w", 16)
    E, O = list(T[:N]), [0] * N
    for i in range(N):
        reduce_row(p, E, O, carry_in=(i > 0))
        if i < N - 1:
            X = E[1]
            Oin = E[2:] + [0, 0]
            E = O
            O = Oin
            E[0] = p.op("add.cc", E[0], X)
    u = [p.op("add.cc", E[1], O[0])]
    for k in range(1, 7):
        u.append(p.op("addc.cc", E[k + 1], O[k]))
    u.append(p.op("addc.cc", O[7], 0))
    p.emit("assert_nc", None)
    for k in range(N):
        p.emit("add.cc" if k == 0 else "addc.cc", "r%d" % k, u[k], T[N + k])
    p.emit("assert_nc", None)
    return p


def gen_chain(name, op, n):
    """raw n-limb r = a op b (op in add, sub); the simulator asserts there is no carry / borrow out."""
    p = Prog(name, nm("a", n) + nm("b", n), nm("r", n))
    for k in range(n):
        p.emit(("%s.cc" % op) if k == 0 else ("%sc.cc" % op), "r%d" % k, "a%d" % k, "b%d" % k)
    p.emit("assert_nc", None)
    return p


def gen_addhi():
    """r = x + c * 2^224 for a 9-limb constant c (a multiple of p: the value modulo p is unchanged)."""
    p = Prog("lz_addhi", nm("x", 16) + nm("c", 9), nm("r", 16))
    for k in range(7):
        p.emit("mov", "r%d" % k, "x%d" % k)
    for k in range(9):
        p.emit("add.cc" if k == 0 else "addc.cc", "r%d" % (7 + k), "x%d" % (7 + k), "c%d" % k)
    p.emit("assert_nc", None)
    return p


def gen_hi(name, op):
    """r = x op (n * 2^256) for a 16-limb x and an 8-limb n (op in add, sub): only the upper half moves; no carry / borrow out."""
    p = Prog(name, nm("x", 16) + nm("n", 8), nm("r", 16))
    for k in range(8):
        p.emit("mov", "r%d" % k, "x%d" % k)
    for k in range(8):
        p.emit(("%s.cc" % op) if k == 0 else ("%sc.cc" % op), "r%d" % (8 + k), "x%d" % (8 + k), "n%d" % k)
    p.emit("assert_nc", None)
    return p


def gen_shl3w():
    """r = 8 x for a 16-limb x < 2^509: sixteen independent funnel shifts instead of three dependent 16-limb doublings (the 9 x of xi = 9 + u)."""
    p = Prog("lz_shl3w", nm("x", 16), nm("r", 16))
    p.emit("shl", "r0", "x0", 3)
    for k in range(1, 16):
        p.emit("shf.l", "r%d" % k, "x%d" % (k - 1), "x%d" % k, 3)
    return p


def gen_csub(name, n):
    """the top 8 limbs of an n-limb x (n = 8 or 16) minus the 8-limb constant k if that does not borrow, else unchanged."""
    p = Prog(name, nm("x", n) + nm("k", 8), nm("r", n))
    lo = n - 8
    for k in range(lo):
        p.emit("mov", "r%d" % k, "x%d" % k)
    s = [p.op("sub.cc" if k == 0 else "subc.cc", "x%d" % (lo + k), "k%d" % k) for k in range(8)]
    bor = p.op("subc", 0, 0)
    for k in range(8):
        p.emit("selp_eqz", "r%d" % (lo + k), s[k], "x%d" % (lo + k), bor)
    return p


def leaf_progs():
    return [gen_mulw(), gen_redc_nf(), gen_chain("lz_add8", "add", 8), gen_chain("lz_sub8", "sub", 8), gen_chain("lz_addw", "add", 16),
            gen_chain("lz_subw", "sub", 16), gen_addhi(), gen_csub("lz_csubw", 16), gen_csub("lz_csub8", 8), gen_hi("lz_addw_hi", "add"), gen_hi("lz_subw_hi", "sub"), gen_mulw2(), gen_shl3w()]


def check_leaves(iters=600):
    rnd = random.Random(0x1A27)
    val = lambda v: sum(x << (32 * i) for i, x in enumerate(v))
    lim = lambda x, n: {i: (x >> (32 * i)) & MASK for i in range(n)}
    def run(prog, **kw):
        inp = {}
        for c, (x, n) in kw.items():
            inp.update({"%s%d" % (c, i): v for i, v in lim(x, n).items()})
        return val(prog.run(inp))
    mulw, redc, add8, sub8, addw, subw, addhi, csubw, csub8, addw_hi, subw_hi, mulw2, shl3w = leaf_progs()
    Bw = P << 256
    for it in range(iters):
        a, b = rnd.getrandbits(256), rnd.getrandbits(256)
        if it < 4:
            a, b = [(0, 0), ((1 << 256) - 1, (1 << 256) - 1), (P - 1, P - 1), (1, (1 << 256) - 1)][it]
        assert run(mulw, a=(a, 8), b=(b, 8)) == a * b
        c2, d2 = rnd.getrandbits(256), rnd.getrandbits(256)
        a2, b2 = a, b
        if it < 4:
            c2, d2 = [((1 << 256) - 1, 1 << 255), (0, 0), ((1 << 256) - 1, (1 << 256) - 1), (P - 1, P - 1)][it]
        while a2 * b2 + c2 * d2 >= 1 << 512:
            a2 >>= 1; c2 >>= 1
        assert run(mulw2, a=(a2, 8), b=(b2, 8), c=(c2, 8), d=(d2, 8)) == a2 * b2 + c2 * d2, "mulw2"
        t = rnd.randrange(4 * Bw) if it > 3 else [0, 4 * Bw - 1, Bw, Bw - 1][it]
        r = run(redc, w=(t, 16))
        assert r < (t >> 256) + P + 1 and (r << 256) % P == t % P, "redc"
        x, y = sorted((rnd.getrandbits(255), rnd.getrandbits(255)))
        assert run(add8, a=(x, 8), b=(y, 8)) == x + y and run(sub8, a=(y, 8), b=(x, 8)) == y - x
        x, y = sorted((rnd.getrandbits(511), rnd.getrandbits(511)))
        assert run(addw, a=(x, 16), b=(y, 16)) == x + y and run(subw, a=(y, 16), b=(x, 16)) == y - x
        xs = rnd.getrandbits(509) if it > 3 else [0, (1 << 509) - 1, 1, (1 << 508) + 5][it]
        assert run(shl3w, x=(xs, 16)) == xs << 3
        nn = rnd.getrandbits(255)
        assert run(addw_hi, x=(x, 16), n=(nn, 8)) == x + (nn << 256)
        big = x | (1 << 511)
        assert run(subw_hi, x=(big, 16), n=(nn, 8)) == big - (nn << 256)
        c = rnd.randrange(1 << 32) * P
        assert run(addhi, x=(x, 16), c=(c, 9)) == x + (c << 224)
        k = rnd.randrange(1, 5)
        w = rnd.randrange(1 << 512) if it > 3 else [k * Bw, k * Bw - 1, 0, (1 << 512) - 1][it]
        assert run(csubw, x=(w, 16), k=(k * P, 8)) == (w - k * Bw if w >= k * Bw else w)
        w = rnd.getrandbits(256) if it > 3 else [k * P, k * P - 1, 0, (1 << 256) - 1][it]
        assert run(csub8, x=(w, 8), k=(k * P, 8)) == (w - k * P if w >= k * P else w)
    return True


LEAF_SIGS = {
    "lz_mulw": "uint32_t* w, const uint32_t* a, const uint32_t* b", "lz_redc": "uint32_t* r, const uint32_t* w",
    "lz_add8": "uint32_t* r, const uint32_t* a, const uint32_t* b", "lz_sub8": "uint32_t* r, const uint32_t* a, const uint32_t* b",
    "lz_addw": "uint32_t* r, const uint32_t* a, const uint32_t* b", "lz_subw": "uint32_t* r, const uint32_t* a, const uint32_t* b",
    "lz_addhi": "uint32_t* r, const uint32_t* x, const uint32_t* c", "lz_csubw": "uint32_t* r, const uint32_t* x, const uint32_t* k",
    "lz_csub8": "uint32_t* r, const uint32_t* x, const uint32_t* k",
    "lz_addw_hi": "uint32_t* r, const uint32_t* x, const uint32_t* n", "lz_subw_hi": "uint32_t* r, const uint32_t* x, const uint32_t* n",
    "lz_mulw2": "uint32_t* w, const uint32_t* a, const uint32_t* b, const uint32_t* c, const uint32_t* d",
    "lz_shl3w": "uint32_t* r, const uint32_t* x",
}


def limbs(x):
    return [(x >> (32 * i)) & MASK for i in range(N)]


def unl(v):
    return sum(x << (32 * i) for i, x in enumerate(v))


def check(verbose=True, iters=3000):
    rnd = random.Random(0xB200)
    rinv = pow(RMONT, -1, P)
    mul, add, sub = gen_mul(), gen_add(), gen_sub()
    edge = [0, 1, 2, P - 1, P - 2, (P - 1) // 2, (P + 1) // 2, RMONT % P, (1 << 253), (1 << 254) - 1 if (1 << 254) - 1 < P else P - 3]
    cases = [(x, y) for x in edge for y in edge] + [(rnd.randrange(P), rnd.randrange(P)) for _ in range(iters)]
    for x, y in cases:
        inp = {"a%d" % i: v for i, v in enumerate(limbs(x))}
        inp.update({"b%d" % i: v for i, v in enumerate(limbs(y))})
        assert unl(mul.run(inp)) == x * y * rinv % P, ("mul", hex(x), hex(y))
        assert unl(add.run(inp)) == (x + y) % P, ("add", hex(x), hex(y))
        assert unl(sub.run(inp)) == (x - y) % P, ("sub", hex(x), hex(y))
    f2m, f2s = gen_fp2_mul(), gen_fp2_sqr()
    top = (1 << 256) - 1
    cases2 = [(a, b, c, d) for a in (0, 1, P - 1) for b in (0, P - 1) for c in (0, 2, P - 1) for d in (0, 1, P - 1)]
    cases2 += [tuple(rnd.randrange(P) for _ in range(4)) for _ in range(iters // 3)]
    for a0, a1, b0, b1 in cases2:
        inp = {}
        for c, v in (("x", a0), ("y", a1), ("z", b0), ("w", b1)):
            inp.update({"%s%d" % (c, i): l for i, l in enumerate(limbs(v))})
        o = f2m.run(inp)
        assert unl(o[:N]) == (a0 * b0 - a1 * b1) * rinv % P and unl(o[N:]) == (a0 * b1 + a1 * b0) * rinv % P, ("fp2_mul", a0, a1, b0, b1)
        o = f2s.run(inp)
        assert unl(o[:N]) == (a0 * a0 - a1 * a1) * rinv % P and unl(o[N:]) == 2 * a0 * a1 * rinv % P, ("fp2_sqr", a0, a1)
    # the wide multiplier alone on full-range operands (no dropped carry for any 256-bit input)
    mw = Prog("mul_wide", ["x%d" % i for i in range(N)] + ["y%d" % i for i in range(N)], [])
    mw.outs = mul_wide(mw, ["x%d" % i for i in range(N)], ["y%d" % i for i in range(N)])
    for a0, b0 in [(top, top), (top, 1), (1 << 255, top), (P, P)] + [(rnd.getrandbits(256), rnd.getrandbits(256)) for _ in range(iters // 3)]:
        inp = {"x%d" % i: l for i, l in enumerate(limbs(a0))}
        inp.update({"y%d" % i: l for i, l in enumerate(limbs(b0))})
        assert sum(v << (32 * i) for i, v in enumerate(mw.run(inp))) == a0 * b0
    check_leaves()
    if verbose:
        print("fp2_mul: %d IMAD.WIDE-equivalent pairs, %d add/sub/logic ops; fp2_sqr: %d pairs; all %d cases ok"
              % ((f2m.count("mad") + f2m.count("mul") - 16) // 2, f2m.count("add") + f2m.count("sub") + f2m.count("and") + f2m.count("selp"),
                 (f2s.count("mad") + f2s.count("mul") - 16) // 2, len(cases2)))
        print("fp_mul: %d mad/mul ops (%d IMAD.WIDE-equivalent pairs + %d single), %d add/sub ops; all %d cases ok"
              % (mul.count("mad") + mul.count("mul"), (mul.count("mad") + mul.count("mul") - 8) // 2, 8,
                 mul.count("add") + mul.count("sub"), len(cases)))
    return True


HEADER = """// GENERATED by tools/gen_fp_ptx.py (instruction lists verified by simulation before emission). Do not edit.
// BN254 Fp primitives, 8 x 32-bit limbs, Montgomery R = 2^256, operands and results in [0, p).
// Replaces the field arithmetic inside EVM precompiles 0x06-0x08 that the reference static-calls
// (/root/reference/contracts/src/common/groth16.rs:54-55,121-125).
#pragma once
#include <stdint.h>
"""


def render():
    out = [HEADER]
    for prog, sig in ((gen_mul(), "const uint32_t* a, const uint32_t* b"), (gen_add(), "const uint32_t* a, const uint32_t* b"),
                      (gen_sub(), "const uint32_t* a, const uint32_t* b")):
        out.append("__device__ __forceinline__ void %s_ptx(uint32_t* r, %s) {\n%s\n}\n" % (prog.name, sig, prog.cuda()))
    out.append("// (r + q u) = (x + y u)(z + w u): lazily reduced, 3 x 64 + 2 x 72 IMAD.WIDE\n"
               "__device__ __forceinline__ void fp2_mul_ptx(uint32_t* r, uint32_t* q, const uint32_t* x, const uint32_t* y, const uint32_t* z, const uint32_t* w) {\n%s\n}\n"
               % gen_fp2_mul().cuda())
    out.append("// (r + q u) = (x + y u)^2\n"
               "__device__ __forceinline__ void fp2_sqr_ptx(uint32_t* r, uint32_t* q, const uint32_t* x, const uint32_t* y) {\n%s\n}\n"
               % gen_fp2_sqr().cuda())
    out.append("// ---- leaf primitives of the lazily reduced tower (composed by csrc/lazy_gen.cuh)\n")
    for prog in leaf_progs():
        out.append("__device__ __forceinline__ void %s(%s) {\n%s\n}\n" % (prog.name, LEAF_SIGS[prog.name], prog.cuda()))
    return "\n".join(out)


def main():
    check()
    if "--check" in sys.argv:
        return
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, "stylus_zkvm_verifiers_b200/csrc/fp_ptx.cuh"), "w") as f:
        f.write(render())
    print("wrote fp_ptx.cuh")


if __name__ == "__main__":
    main()
