"""Synthetic Groth16/BN254 workloads for the RISC Zero and SP1 proof shapes.

Proofs are *simulated with known trapdoors*: a random verification key is built from scalars
(alpha, beta, gamma, delta, ic_0..ic_k) as multiples of the group generators, and a valid proof for
public inputs s is (A, B, C) = (a G1, b G2, c G1) with c solved from the verification equation the
reference checks (/root/reference/contracts/src/common/groth16.rs:86-107):

    RISC Zero (A is negated by the verifier, groth16.rs:96-99):  c = (a b - alpha beta - x gamma) / delta
    SP1 (vk stores beta' = -beta etc., A used as is, :100-103):   c = -(a b + alpha beta' + x gamma') / delta'

with x = ic_0 + sum s_i ic_{i+1} (mod r).  Scalar arithmetic is plain Python integers; the group
multiplications are delegated to a *backend* object with `g1_mul(scalars)` / `g2_mul(scalars)`
(multiples of the generators, returned as EVM words).  bench.py passes the GPU backend (the
library's own batched ecMul service); tests may pass the CPU oracle instead, and check each against
the other.  Nothing here imports the oracle.
"""
import hashlib
import struct

P = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47   # Q, groth16.rs:10
R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001   # R, groth16.rs:9

G1_GEN = (1).to_bytes(32, "big") + (2).to_bytes(32, "big")
# standard G2 generator in wire order x_im, x_re, y_im, y_re (== RISC Zero gamma2, risc0/crypto.rs:32-41)
G2_GEN = b"".join(int(h, 16).to_bytes(32, "big") for h in (
    "198e9393920d483a7260bfb731fb5d25f1aa493335a9e71297e485b7aef312c2",
    "1800deef121f1e76426a00665e5c4479674322d4f75edadd46debd5cd992f6ed",
    "090689d0585ff075ec9e99ad690c3395bc4b313370b38ef355acdadcd122975b",
    "12c85ea5db8c6deb4aab71808dcb408fe3d1e7690c43d37b4ce6cc0166fa7daa"))

ST_OK, ST_INVALID_INITIALIZATION, ST_INVALID_PROOF_DATA, ST_SELECTOR_MISMATCH, ST_VERIFICATION_FAILED = range(5)
SP1_SELECTOR = bytes.fromhex("a4594c59")   # sp1/config.rs:4-9,18-20


def w32(v):
    return int(v).to_bytes(32, "big")


class SplitMix64:
    def __init__(self, seed):
        self.s = seed & 0xFFFFFFFFFFFFFFFF

    def next(self):
        self.s = (self.s + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return z ^ (z >> 31)

    def below(self, n):
        return self.next() % n

    def u256(self):
        return (self.next() << 192) | (self.next() << 128) | (self.next() << 64) | self.next()

    def fr(self):
        while True:
            v = self.u256() >> 2
            if 0 < v < R:
                return v

    def bytes(self, n):
        out = b""
        while len(out) < n:
            out += struct.pack(">Q", self.next())
        return out[:n]


# ------------------------------------------------------------------ public-input computation (independent of the library: hashlib)
def risc0_claim_digest(image_id, journal_digest, sys0):
    """ReceiptClaim::ok(image_id, journal).digest(), risc0/types.rs:44-95."""
    sha = lambda b: hashlib.sha256(b).digest()
    out = sha(sha(b"risc0.Output") + journal_digest + bytes(32) + b"\x02\x00")
    return sha(sha(b"risc0.ReceiptClaim") + bytes(32) + image_id + sys0 + out + bytes(4) + bytes(4) + b"\x04\x00")


def split_digest(d):
    """risc0/crypto.rs:103-110 -> (low, high) as integers."""
    rev = d[::-1]
    return int.from_bytes(rev[16:], "big"), int.from_bytes(rev[:16], "big")


def risc0_signals(control_root, bn254_control_id, claim):
    c0, c1 = split_digest(control_root)
    lo, hi = split_digest(claim)
    return [c0, c1, lo, hi, int.from_bytes(bn254_control_id, "big")]


def sp1_signals(vkey, public_values):
    h = int.from_bytes(hashlib.sha256(public_values).digest(), "big") & ((1 << 253) - 1)
    return [int.from_bytes(vkey, "big"), h % R]


# ------------------------------------------------------------------ keys
class SynthVk:
    def __init__(self, vm, trap, alpha, beta, gamma, delta, ic):
        self.vm, self.trap = vm, trap
        self.alpha, self.beta, self.gamma, self.delta, self.ic = alpha, beta, gamma, delta, ic

    @property
    def k(self):
        return len(self.ic) - 1

    def x_of(self, signals):
        t = self.trap
        x = t["ic"][0]
        for s, c in zip(signals, t["ic"][1:]):
            x = (x + s * c) % R
        return x


def make_vk(backend, vm, n_ic, seed):
    """Random key of the given shape.  For vm == 1 (SP1) the stored beta/gamma/delta play the role of the
    reference's pre-negated constants (sp1/crypto.rs:14,30,46); being random they need no explicit negation."""
    rng = SplitMix64(seed)
    trap = {"alpha": rng.fr(), "beta": rng.fr(), "gamma": rng.fr(), "delta": rng.fr(), "ic": [rng.fr() for _ in range(n_ic)]}
    g1 = backend.g1_mul([trap["alpha"]] + trap["ic"])
    g2 = backend.g2_mul([trap["beta"], trap["gamma"], trap["delta"]])
    return SynthVk(vm, trap, g1[0], g2[0], g2[1], g2[2], g1[1:])


class Pools:
    """(a, a G1) and (b, b G2) pools so that only C needs a fresh group multiplication per proof."""

    def __init__(self, backend, rng, size):
        self.a = [rng.fr() for _ in range(size)]
        self.b = [rng.fr() for _ in range(size)]
        self.A = backend.g1_mul(self.a)
        self.B = backend.g2_mul(self.b)


def solve_c(vk, a, b, x):
    t = vk.trap
    if vk.vm == 0:
        num = (a * b - t["alpha"] * t["beta"] - x * t["gamma"]) % R
    else:
        num = (-(a * b + t["alpha"] * t["beta"] + x * t["gamma"])) % R
    if "delta_inv" not in t:
        t["delta_inv"] = pow(t["delta"], -1, R)
    return num * t["delta_inv"] % R


def make_proofs(backend, vk, signals_list, rng, pools):
    """One valid 256-byte proof (a, b, c words; common/groth16.rs:23-31 argument order) per signal vector."""
    n = len(signals_list)
    ja = [rng.below(len(pools.a)) for _ in range(n)]
    jb = [rng.below(len(pools.b)) for _ in range(n)]
    cs = [solve_c(vk, pools.a[ja[i]], pools.b[jb[i]], vk.x_of(signals_list[i])) for i in range(n)]
    C = backend.g1_mul(cs)
    return [pools.A[ja[i]] + pools.B[jb[i]] + C[i] for i in range(n)]


# ------------------------------------------------------------------ batches of the two shapes
class Risc0Batch:
    def __init__(self, seals, image_ids, journals, expect=None, classes=None):
        self.seals, self.image_ids, self.journals, self.expect, self.classes = seals, image_ids, journals, expect, classes


def make_risc0_batch(backend, vk, selector, control_root, bn254_control_id, sys0, n, seed, pool=1024):
    rng = SplitMix64(seed)
    pools = Pools(backend, rng, min(pool, max(n, 1)))
    image_ids = [rng.bytes(32) for _ in range(n)]
    journals = [rng.bytes(32) for _ in range(n)]
    sigs = [risc0_signals(control_root, bn254_control_id, risc0_claim_digest(image_ids[i], journals[i], sys0)) for i in range(n)]
    proofs = make_proofs(backend, vk, sigs, rng, pools)
    return Risc0Batch([selector + p for p in proofs], image_ids, journals, [ST_OK] * n, ["valid"] * n)


class Sp1Batch:
    def __init__(self, proofs, vkeys, public_values, expect=None, classes=None):
        self.proofs, self.vkeys, self.public_values, self.expect, self.classes = proofs, vkeys, public_values, expect, classes


def make_sp1_batch(backend, vk, n, seed, pv_len=96, pool=1024, selector=SP1_SELECTOR):
    rng = SplitMix64(seed)
    pools = Pools(backend, rng, min(pool, max(n, 1)))
    vkeys = [bytes([0]) + rng.bytes(31) for _ in range(n)]        # real vkeys have the top byte clear (sp1 fixture)
    pvs = [rng.bytes(pv_len) for _ in range(n)]
    sigs = [sp1_signals(vkeys[i], pvs[i]) for i in range(n)]
    proofs = make_proofs(backend, vk, sigs, rng, pools)
    return Sp1Batch([selector + p for p in proofs], vkeys, pvs, [ST_OK] * n, ["valid"] * n)


# ------------------------------------------------------------------ Fp2 helpers for off-subgroup twist points
def _f2mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def _f2pow(a, e):
    r = (1, 0)
    while e:
        if e & 1:
            r = _f2mul(r, a)
        a = _f2mul(a, a)
        e >>= 1
    return r


def _f2sqrt(a):
    """Square root in Fp2 = Fp[u]/(u^2+1), p = 3 mod 4 (complex method); None if a is a non-residue."""
    if a == (0, 0):
        return (0, 0)
    n = (a[0] * a[0] + a[1] * a[1]) % P
    s = pow(n, (P + 1) // 4, P)
    if s * s % P != n:
        return None
    inv2 = pow(2, -1, P)
    for sg in (s, (-s) % P):
        t = (a[0] + sg) * inv2 % P
        x = pow(t, (P + 1) // 4, P)
        if x * x % P == t and x:
            y = a[1] * pow(2 * x, -1, P) % P
            if _f2mul((x, y), (x, y)) == (a[0] % P, a[1] % P):
                return (x, y)
    return None


_XI_INV = None


def twist_b():
    global _XI_INV
    if _XI_INV is None:
        d = pow(82, -1, P)                        # 1/(9+u) = (9-u)/82
        _XI_INV = (9 * d % P, (-d) % P)
    return _f2mul((3, 0), _XI_INV)


def random_twist_point(rng):
    """A point on y^2 = x^3 + 3/(9+u) that is (with overwhelming probability) NOT in the order-r subgroup:
    the twist has order r * (2p - r) (SURVEY.md section 7 numeric anchors)."""
    b = twist_b()
    while True:
        x = (rng.u256() % P, rng.u256() % P)
        x3 = _f2mul(_f2mul(x, x), x)
        y = _f2sqrt(((x3[0] + b[0]) % P, (x3[1] + b[1]) % P))
        if y is not None:
            return w32(x[1]) + w32(x[0]) + w32(y[1]) + w32(y[0])


# ------------------------------------------------------------------ config 4: mixed valid / invalid / malformed batch
MIXED_CLASSES = (("valid", 50), ("tampered", 15), ("off_curve", 10), ("coord_ge_p", 5), ("wrong_subgroup", 10), ("infinity", 5), ("malformed", 5))


def _pick_class(rng):
    v = rng.below(100)
    for name, w in MIXED_CLASSES:
        if v < w:
            return name
        v -= w
    return "valid"


def _word(seal, i):
    return int.from_bytes(seal[4 + 32 * i:36 + 32 * i], "big")


def _set_word(seal, i, v):
    return seal[:4 + 32 * i] + w32(v % (1 << 256)) + seal[36 + 32 * i:]


def mutate_risc0(batch, backend, rng, pools=None, wrong_pts=None):
    """Turn a valid RISC Zero-shape batch into the mixed batch of SURVEY.md section 8d config 4, in place.
    `expect` is left None where the outcome is not known by construction (decided by the oracle in tests)."""
    n = len(batch.seals)
    if wrong_pts is None:
        wrong_pts = [random_twist_point(rng) for _ in range(32)]
    for i in range(n):
        cls = _pick_class(rng)
        batch.classes[i] = cls
        seal = batch.seals[i]
        if cls == "valid":
            continue
        batch.expect[i] = ST_VERIFICATION_FAILED
        sub = rng.below(6)
        if cls == "tampered":
            if sub == 0:
                b = bytearray(batch.image_ids[i]); b[rng.below(32)] ^= 1 << rng.below(8); batch.image_ids[i] = bytes(b)
            elif sub == 1:
                b = bytearray(batch.journals[i]); b[rng.below(32)] ^= 1 << rng.below(8); batch.journals[i] = bytes(b)
            elif sub in (2, 3) and pools is not None:
                other = pools.A[rng.below(len(pools.A))]
                at = 4 if sub == 2 else 4 + 192
                if seal[at:at + 64] == other:
                    other = pools.A[(pools.A.index(other) + 1) % len(pools.A)]
                seal = seal[:at] + other + seal[at + 64:]
            else:
                other = pools.B[rng.below(len(pools.B))] if pools is not None else G2_GEN
                if seal[68:196] == other:
                    other = G2_GEN
                seal = seal[:68] + other + seal[196:]
        elif cls == "off_curve":
            wi = (1, 7, 5)[sub % 3]                      # A.y, C.y, B.y_re
            seal = _set_word(seal, wi, (_word(seal, wi) + 1) % P)
        elif cls == "coord_ge_p":
            if sub == 0:
                seal = _set_word(_set_word(seal, 0, 0), 1, P)        # A = (0, Q): negate_g1 maps it to (0,0) = infinity
                batch.expect[i] = None
            elif sub == 1:
                seal = _set_word(seal, 1, 0)                          # A = (x, 0) -> y' = Q -> rejected
            else:
                wi = (0, 1, 6, 7, 2, 3, 4, 5)[rng.below(8)]
                v = _word(seal, wi) + P
                if v >= (1 << 256):
                    v = P
                seal = _set_word(seal, wi, v)
        elif cls == "wrong_subgroup":
            seal = seal[:68] + wrong_pts[rng.below(len(wrong_pts))] + seal[196:]
        elif cls == "infinity":
            batch.expect[i] = None                       # A or B = infinity leaves a 3-pair product: decided by the oracle
            if sub % 3 == 0:
                seal = seal[:4] + bytes(64) + seal[68:]
            elif sub % 3 == 1:
                seal = seal[:68] + bytes(128) + seal[196:]
            else:
                seal = seal[:196] + bytes(64)
        elif cls == "malformed":
            if sub == 0:
                seal = seal[:rng.below(4)]; batch.expect[i] = ST_INVALID_PROOF_DATA
            elif sub == 1:
                seal = seal[:4 + rng.below(256)]; batch.expect[i] = ST_INVALID_PROOF_DATA
            elif sub == 2:
                seal = seal + rng.bytes(1 + rng.below(64)); batch.expect[i] = ST_INVALID_PROOF_DATA
            else:
                b = bytearray(seal); b[rng.below(4)] ^= 1 << rng.below(8); seal = bytes(b); batch.expect[i] = ST_SELECTOR_MISMATCH
        batch.seals[i] = seal
    return batch


def mutate_sp1(batch, backend, rng, pools=None, wrong_pts=None):
    """SP1-shape counterpart of mutate_risc0 (proof words a, b, c at the same offsets; no A negation)."""
    n = len(batch.proofs)
    if wrong_pts is None:
        wrong_pts = [random_twist_point(rng) for _ in range(32)]
    for i in range(n):
        cls = _pick_class(rng)
        batch.classes[i] = cls
        pr = batch.proofs[i]
        if cls == "valid":
            continue
        batch.expect[i] = ST_VERIFICATION_FAILED
        sub = rng.below(6)
        if cls == "tampered":
            if sub == 0:
                b = bytearray(batch.public_values[i]); b[rng.below(len(b))] ^= 1 << rng.below(8); batch.public_values[i] = bytes(b)
            elif sub == 1:
                b = bytearray(batch.vkeys[i]); b[1 + rng.below(31)] ^= 1 << rng.below(8); batch.vkeys[i] = bytes(b)
            elif sub == 2:
                batch.vkeys[i] = w32(R + rng.below(1 << 64))          # signal >= R, groth16.rs:32-34
            else:
                pr = pr[:68] + G2_GEN + pr[196:]
        elif cls == "off_curve":
            wi = (1, 7, 5)[sub % 3]
            pr = _set_word(pr, wi, (_word(pr, wi) + 1) % P)
        elif cls == "coord_ge_p":
            if sub == 0:
                pr = _set_word(_set_word(pr, 0, 0), 1, P)             # no negation on the SP1 path: y = Q is simply out of range
            else:
                wi = (0, 1, 6, 7, 2, 3, 4, 5)[rng.below(8)]
                v = _word(pr, wi) + P
                if v >= (1 << 256):
                    v = P
                pr = _set_word(pr, wi, v)
        elif cls == "wrong_subgroup":
            pr = pr[:68] + wrong_pts[rng.below(len(wrong_pts))] + pr[196:]
        elif cls == "infinity":
            batch.expect[i] = None
            if sub % 3 == 0:
                pr = pr[:4] + bytes(64) + pr[68:]
            elif sub % 3 == 1:
                pr = pr[:68] + bytes(128) + pr[196:]
            else:
                pr = pr[:196] + bytes(64)
        elif cls == "malformed":
            if sub == 0:
                pr = pr[:rng.below(4)]; batch.expect[i] = ST_INVALID_PROOF_DATA
            elif sub == 1:
                pr = pr[:4 + rng.below(256)]; batch.expect[i] = ST_INVALID_PROOF_DATA
            elif sub == 2:
                pr = pr + rng.bytes(1 + rng.below(64)); batch.expect[i] = ST_INVALID_PROOF_DATA
            else:
                b = bytearray(pr); b[rng.below(4)] ^= 1 << rng.below(8); pr = bytes(b); batch.expect[i] = ST_SELECTOR_MISMATCH
        batch.proofs[i] = pr
    return batch


def make_pairing4_batch(backend, vk, n, seed, pool=256, valid_fraction=50):
    """Config 5: n instances (4 G1 points + 1 G2 point) for zkv_pairing4_batch with the fixed G2 points of `vk`
    (treated as plain beta, gamma, delta).  About valid_fraction % satisfy the product equation
    e(P0,Q) e(P1,beta) e(P2,gamma) e(P3,delta) = 1; returns (g1s, g2s, expect_ok)."""
    rng = SplitMix64(seed)
    t = vk.trap
    pools = Pools(backend, rng, min(pool, max(n, 1)))
    p1 = [rng.fr() for _ in range(n)]
    p2 = [rng.fr() for _ in range(n)]
    ja = [rng.below(len(pools.a)) for _ in range(n)]
    jb = [rng.below(len(pools.b)) for _ in range(n)]
    dinv = pow(t["delta"], -1, R)
    p3, expect = [], []
    for i in range(n):
        good = rng.below(100) < valid_fraction
        c = (-(pools.a[ja[i]] * pools.b[jb[i]] + p1[i] * t["beta"] + p2[i] * t["gamma"])) * dinv % R
        if not good:
            c = (c + 1 + rng.below(1 << 32)) % R
        p3.append(c); expect.append(1 if good else 0)
    pts = backend.g1_mul(p1 + p2 + p3)
    g1s = [pools.A[ja[i]] + pts[i] + pts[n + i] + pts[2 * n + i] for i in range(n)]
    g2s = [pools.B[jb[i]] for i in range(n)]
    return g1s, g2s, expect
