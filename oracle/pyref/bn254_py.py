"""Slow, independent BN254 referee (TEST INFRASTRUCTURE ONLY -- never on the product path).

Purpose: pin the C oracle (oracle/zkv_oracle.c) with an implementation that shares
nothing with it: Fp12 is the flat polynomial ring Fp[w]/(w^12 - 18 w^6 + 82), curve
arithmetic is affine with modular inverses, the final exponentiation is a plain
pow(f, (p^12-1)/r).  It follows EIP-196/197 semantics for the three precompiles the
reference static-calls (/root/reference/contracts/src/common/groth16.rs:12-14,60-73,
109-128) and restates the reference verify logic on top of them
(groth16.rs:23-49, risc0/verifier.rs:128-197, sp1/verifier.rs:58-111).
"""
import hashlib

P = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47  # groth16.rs:10 (Q)
R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001  # groth16.rs:9
U = 4965661367192848881
ATE = 6 * U + 2
assert P == 36 * U**4 + 36 * U**3 + 24 * U**2 + 6 * U + 1
assert R == 36 * U**4 + 36 * U**3 + 18 * U**2 + 6 * U + 1

# hard-part convention shared by the C oracle and the CUDA path:
# GT = m^((p^6-1)(p^2+1)(L0 + L1 p + L2 p^2 + L3 p^3)),  sum = LAMBDA * (p^4-p^2+1)/r
L0 = 1 + 6 * U + 12 * U**2 + 12 * U**3
L1 = 4 * U + 6 * U**2 + 12 * U**3
L2 = 6 * U + 6 * U**2 + 12 * U**3
L3 = -1 + 4 * U + 6 * U**2 + 12 * U**3
LAMBDA = 2 * U * (6 * U**2 + 3 * U + 1)
assert (L0 + L1 * P + L2 * P**2 + L3 * P**3) * R == LAMBDA * (P**4 - P**2 + 1)


def inv(a, m=P):
    return pow(a, -1, m) if a % m else 0


# ---------------------------------------------------------------- Fp2 (tuples)
def f2add(a, b): return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)
def f2sub(a, b): return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)
def f2neg(a): return (-a[0] % P, -a[1] % P)
def f2mul(a, b): return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)
def f2scal(a, k): return (a[0] * k % P, a[1] * k % P)
def f2inv(a):
    d = inv(a[0] * a[0] + a[1] * a[1])
    return (a[0] * d % P, -a[1] * d % P)
def f2conj(a): return (a[0], -a[1] % P)


XI = (9, 1)
B2 = f2mul((3, 0), f2inv(XI))  # twist b' = 3/(9+u)


# ---------------------------------------------------------------- Fp12 = Fp[w]/(w^12-18w^6+82)
def f12mul(a, b):
    t = [0] * 23
    for i, x in enumerate(a):
        if x:
            for j, y in enumerate(b):
                t[i + j] += x * y
    for k in range(22, 11, -1):
        c = t[k]
        if c:
            t[k - 6] += 18 * c
            t[k - 12] -= 82 * c
    return [x % P for x in t[:12]]


F12_ONE = [1] + [0] * 11


def f12pow(a, e):
    r = F12_ONE
    for bit in bin(e)[2:]:
        r = f12mul(r, r)
        if bit == "1":
            r = f12mul(r, a)
    return r


def f12_from_f2(c, k):
    """(c0 + c1*u) * w^k  with u = w^6 - 9 ; k < 6."""
    out = [0] * 12
    out[k] = (c[0] - 9 * c[1]) % P
    out[k + 6] = c[1] % P
    return out


def f12add(a, b): return [(x + y) % P for x, y in zip(a, b)]


def tower_to_poly(t):
    """12 Fp coefficients in tower order c{0,1}(w) . c{0,1,2}(v) . c{0,1}(u) -> flat poly."""
    out = [0] * 12
    idx = 0
    for i in range(2):
        for j in range(3):
            c = (t[idx], t[idx + 1]); idx += 2
            out = f12add(out, f12_from_f2(c, i + 2 * j))
    return out


# ---------------------------------------------------------------- curves (affine, None = infinity)
def g1_on_curve(pt):
    x, y = pt
    return (y * y - x * x * x - 3) % P == 0


def g1_add(a, b):
    if a is None: return b
    if b is None: return a
    if a[0] == b[0]:
        if (a[1] + b[1]) % P == 0: return None
        lam = 3 * a[0] * a[0] * inv(2 * a[1]) % P
    else:
        lam = (b[1] - a[1]) * inv(b[0] - a[0]) % P
    x = (lam * lam - a[0] - b[0]) % P
    return (x, (lam * (a[0] - x) - a[1]) % P)


def g1_mul(pt, k):
    acc = None
    while k:
        if k & 1: acc = g1_add(acc, pt)
        pt = g1_add(pt, pt); k >>= 1
    return acc


def g1_neg(pt): return None if pt is None else (pt[0], -pt[1] % P)


def g2_on_curve(pt):
    x, y = pt
    return f2sub(f2mul(y, y), f2add(f2mul(f2mul(x, x), x), B2)) == (0, 0)


def g2_add(a, b):
    if a is None: return b
    if b is None: return a
    if a[0] == b[0]:
        if f2add(a[1], b[1]) == (0, 0): return None
        lam = f2mul(f2scal(f2mul(a[0], a[0]), 3), f2inv(f2scal(a[1], 2)))
    else:
        lam = f2mul(f2sub(b[1], a[1]), f2inv(f2sub(b[0], a[0])))
    x = f2sub(f2sub(f2mul(lam, lam), a[0]), b[0])
    return (x, f2sub(f2mul(lam, f2sub(a[0], x)), a[1]))


def g2_mul(pt, k):
    acc = None
    while k:
        if k & 1: acc = g2_add(acc, pt)
        pt = g2_add(pt, pt); k >>= 1
    return acc


def g2_neg(pt): return None if pt is None else (pt[0], f2neg(pt[1]))


G1 = (1, 2)
G2 = ((0x1800DEEF121F1E76426A00665E5C4479674322D4F75EDADD46DEBD5CD992F6ED,
       0x198E9393920D483A7260BFB731FB5D25F1AA493335A9E71297E485B7AEF312C2),
      (0x12C85EA5DB8C6DEB4AAB71808DCB408FE3D1E7690C43D37B4CE6CC0166FA7DAA,
       0x090689D0585FF075EC9E99AD690C3395BC4B313370B38EF355ACDADCD122975B))
assert g1_on_curve(G1) and g2_on_curve(G2)

# Frobenius on the twist: (x,y) -> (conj(x)*xi^((p-1)/3), conj(y)*xi^((p-1)/2))
def _f2pow(a, e):
    r = (1, 0)
    while e:
        if e & 1: r = f2mul(r, a)
        a = f2mul(a, a); e >>= 1
    return r


TW_X = _f2pow(XI, (P - 1) // 3)
TW_Y = _f2pow(XI, (P - 1) // 2)


def g2_frob(pt):
    return (f2mul(f2conj(pt[0]), TW_X), f2mul(f2conj(pt[1]), TW_Y))


# ---------------------------------------------------------------- pairing
def _line(T, Q2, Pt):
    """Exact line through untwisted T,Q2 (twist points, affine) at G1 point Pt; returns (l, T+Q2)."""
    if T[0] == Q2[0] and T[1] == Q2[1]:
        lam = f2mul(f2scal(f2mul(T[0], T[0]), 3), f2inv(f2scal(T[1], 2)))
    elif T[0] == Q2[0]:
        # vertical line: x_P - x_T*w^2
        l = [Pt[0] % P] + [0] * 11
        l = f12add(l, f12_from_f2(f2neg(T[0]), 2))
        return l, None
    else:
        lam = f2mul(f2sub(Q2[1], T[1]), f2inv(f2sub(Q2[0], T[0])))
    # l = yP - lam*xP*w + (lam*xT - yT)*w^3
    l = [Pt[1] % P] + [0] * 11
    l = f12add(l, f12_from_f2(f2scal(f2neg(lam), Pt[0]), 1))
    l = f12add(l, f12_from_f2(f2sub(f2mul(lam, T[0]), T[1]), 3))
    x = f2sub(f2sub(f2mul(lam, lam), T[0]), Q2[0])
    y = f2sub(f2mul(lam, f2sub(T[0], x)), T[1])
    return l, (x, y)


def miller(Pt, Q):
    """Optimal-ate Miller value f_{6u+2,Q}(P) * two Frobenius lines. Pt in G1, Q on the twist (affine)."""
    if Pt is None or Q is None:
        return F12_ONE
    f = F12_ONE
    T = Q
    for bit in bin(ATE)[3:]:
        l, T2 = _line(T, T, Pt)
        f = f12mul(f12mul(f, f), l); T = T2
        if bit == "1":
            l, T2 = _line(T, Q, Pt)
            f = f12mul(f, l); T = T2
    Q1 = g2_frob(Q)
    nQ2 = g2_neg(g2_frob(Q1))
    l, T2 = _line(T, Q1, Pt); f = f12mul(f, l); T = T2
    l, T2 = _line(T, nQ2, Pt); f = f12mul(f, l)
    return f


FINAL_EXP = (P**12 - 1) // R


def final_exp(f):
    return f12pow(f, FINAL_EXP)


def pairing_product(pairs):
    f = F12_ONE
    for Pt, Q in pairs:
        f = f12mul(f, miller(Pt, Q))
    return final_exp(f)


# ---------------------------------------------------------------- EIP-196/197 precompile byte semantics
def _be(b): return int.from_bytes(b, "big")


def _dec_g1(b):
    x, y = _be(b[:32]), _be(b[32:64])
    if x >= P or y >= P: raise ValueError("coord >= p")
    if x == 0 and y == 0: return None
    if not g1_on_curve((x, y)): raise ValueError("off curve")
    return (x, y)


def _enc_g1(pt):
    if pt is None: return b"\0" * 64
    return pt[0].to_bytes(32, "big") + pt[1].to_bytes(32, "big")


def _dec_g2(b):
    xi, xr, yi, yr = (_be(b[i:i + 32]) for i in range(0, 128, 32))
    if max(xi, xr, yi, yr) >= P: raise ValueError("coord >= p")
    if xi == xr == yi == yr == 0: return None
    Q = ((xr, xi), (yr, yi))
    if not g2_on_curve(Q): raise ValueError("off twist")
    if g2_mul(Q, R) is not None: raise ValueError("not in subgroup")
    return Q


def ec_add(data):
    data = (data + b"\0" * 128)[:128]
    return _enc_g1(g1_add(_dec_g1(data[:64]), _dec_g1(data[64:])))


def ec_mul(data):
    data = (data + b"\0" * 96)[:96]
    return _enc_g1(g1_mul(_dec_g1(data[:64]), _be(data[64:96])))


def ec_pairing(data):
    if len(data) % 192: raise ValueError("bad length")
    pairs = [(_dec_g1(data[i:i + 64]), _dec_g2(data[i + 64:i + 192])) for i in range(0, len(data), 192)]
    ok = pairing_product(pairs) == F12_ONE
    return (1 if ok else 0).to_bytes(32, "big")


# ---------------------------------------------------------------- reference verify logic on top of the precompiles
RISC0, SP1 = 0, 1


def groth16_verify(vm, vk, a, b, c, signals):
    """groth16.rs:23-49. vk = dict(alpha=(x,y), beta/gamma/delta=((x0,x1),(y0,y1)) wire order, ic=[(x,y)...])."""
    w = lambda v: v.to_bytes(32, "big")
    if len(signals) + 1 != len(vk["ic"]) or any(s >= R for s in signals):
        return False
    try:
        vkx = w(vk["ic"][0][0]) + w(vk["ic"][0][1])
        for s, ic in zip(signals, vk["ic"][1:]):
            m = ec_mul(w(ic[0]) + w(ic[1]) + w(s))
            vkx = ec_add(vkx + m)
        if vm == RISC0:  # negate_g1, groth16.rs:75-84
            a = a if (a[0] == 0 and a[1] == 0) else (a[0], (P - a[1]) % (1 << 256))
        g1s = [w(a[0]) + w(a[1]), w(vk["alpha"][0]) + w(vk["alpha"][1]), vkx, w(c[0]) + w(c[1])]
        g2 = lambda q: w(q[0][0]) + w(q[0][1]) + w(q[1][0]) + w(q[1][1])
        g2s = [g2(b), g2(vk["beta"]), g2(vk["gamma"]), g2(vk["delta"])]
        ret = ec_pairing(b"".join(x + y for x, y in zip(g1s, g2s)))
        return _be(ret) != 0
    except ValueError:
        return False


def sha256(b): return hashlib.sha256(b).digest()


def split_digest(d):
    rev = d[::-1]
    return rev[16:], rev[:16]  # (low, high)  risc0/crypto.rs:103-110


SYSTEM_STATE_ZERO_DIGEST = bytes.fromhex("a3acc27117418996340b84e5a90f3ef4c49d22c79e44aad822ec9c313e1eb8e2")


def risc0_claim_digest(image_id, journal_digest):
    out = sha256(sha256(b"risc0.Output") + journal_digest + b"\0" * 32 + b"\x02\x00")
    return sha256(sha256(b"risc0.ReceiptClaim") + b"\0" * 32 + image_id + SYSTEM_STATE_ZERO_DIGEST + out
                  + b"\0\0\0\0" + b"\0\0\0\0" + b"\x04\x00")


def risc0_vk_digest(vk):
    w = lambda v: v.to_bytes(32, "big")
    ic_tag = sha256(b"risc0_groth16.VerifyingKey.IC")
    cur = b"\0" * 32
    for pt in reversed(vk["ic"]):
        cur = sha256(ic_tag + sha256(w(pt[0]) + w(pt[1])) + cur + b"\x02\x00")
    g2d = lambda q: sha256(w(q[0][0]) + w(q[0][1]) + w(q[1][0]) + w(q[1][1]))
    return sha256(sha256(b"risc0_groth16.VerifyingKey") + sha256(w(vk["alpha"][0]) + w(vk["alpha"][1]))
                  + g2d(vk["beta"]) + g2d(vk["gamma"]) + g2d(vk["delta"]) + cur + b"\x05\x00")


def risc0_selector(control_root, bn254_control_id, vk):
    d = sha256(sha256(b"risc0.Groth16ReceiptVerifierParameters") + control_root + bn254_control_id[::-1]
               + risc0_vk_digest(vk) + b"\x03\x00")
    return d[:4]


# ---------------------------------------------------------------- the two verifiers' control flow, as status codes (0 ok, 2 InvalidProofData,
# 3 SelectorMismatch / WrongVerifierSelector, 4 VerificationFailed): restated from the reference independently of oracle/zkv_oracle.c
ST_OK, ST_INVALID_PROOF_DATA, ST_SELECTOR_MISMATCH, ST_VERIFICATION_FAILED = 0, 2, 3, 4


def _front(seal, selector):
    """risc0/verifier.rs:151-170 and sp1/verifier.rs:64-83: length < 4, selector, then strict abi_decode of 8 x uint256"""
    if len(seal) < 4: return ST_INVALID_PROOF_DATA, None
    if seal[:4] != selector: return ST_SELECTOR_MISMATCH, None
    if len(seal) - 4 != 256: return ST_INVALID_PROOF_DATA, None
    return None, [_be(seal[4 + 32 * i:36 + 32 * i]) for i in range(8)]


def risc0_verify_status(vk, selector, control_root, bn254_control_id, seal, image_id, journal_digest):
    """RiscZeroVerifier::verify of an initialised verifier (risc0/verifier.rs:78-92, 146-197)"""
    st, w = _front(seal, selector)
    if st is not None: return st
    c0, c1 = split_digest(control_root)
    lo, hi = split_digest(risc0_claim_digest(image_id, journal_digest))
    sig = [_be(c0), _be(c1), _be(lo), _be(hi), _be(bn254_control_id)]
    return ST_OK if groth16_verify(RISC0, vk, w[0:2], [w[2:4], w[4:6]], w[6:8], sig) else ST_VERIFICATION_FAILED


def sp1_verify_status(vk, selector, vkey, public_values, proof):
    """Sp1Verifier::verify_proof (sp1/verifier.rs:58-111, sp1/types.rs:21-38)"""
    st, w = _front(proof, selector)
    if st is not None: return st
    sig = [_be(vkey), _be(sha256(public_values)) & ((1 << 253) - 1)]
    return ST_OK if groth16_verify(SP1, vk, w[0:2], [w[2:4], w[4:6]], w[6:8], sig) else ST_VERIFICATION_FAILED
