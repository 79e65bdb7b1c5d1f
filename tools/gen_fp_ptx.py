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
            if op == "assert_nc":
                continue
            if op == "selp_eqz":
                lines.append("setp.eq.u32 q, %s, 0;" % fmt(src[2]))
                lines.append("selp.u32 %s, %s, %s, q;" % (fmt(dst), fmt(src[0]), fmt(src[1])))
                continue
            pop = {"and": "and.b32", "mov": "mov.u32"}.get(op, op + ".u32")
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


def reduce_row(p, E, O):
    m = p.op("mul.lo", E[0], M0)
    for idx, j in enumerate((1, 3, 5, 7)):
        O[j - 1] = p.op("mad.lo.cc" if idx == 0 else "madc.lo.cc", m, PL[j], O[j - 1])
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
    if verbose:
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
