"""The synthetic-proof generator (stylus_zkvm_verifiers_b200/synth.py) against the oracle: trapdoor
proofs of both shapes must be accepted, and every class of the mixed batch
must get the status the generator predicts."""
import numpy as np

import oracle_lib as O
from conftest import oracle_vk
from stylus_zkvm_verifiers_b200 import synth as S


def test_risc0_shape_valid_and_mixed(oracle_backend, fx):
    vk = S.make_vk(oracle_backend, 0, 6, 0xB2000001)
    ro = O.Risc0Oracle(oracle_vk(vk)); ro.initialize(fx["control_root"], fx["bn254_control_id"])
    sel = ro.selector()
    n = 96
    batch = S.make_risc0_batch(oracle_backend, vk, sel, fx["control_root"], fx["bn254_control_id"], fx["sys0"], n, 0xB2000001, pool=16)
    st = ro.verify_batch(batch.seals, batch.image_ids, batch.journals)
    assert (st == O.ST_OK).all()
    rng = S.SplitMix64(0xB2000004)
    pools = S.Pools(oracle_backend, rng, 8)
    S.mutate_risc0(batch, oracle_backend, rng, pools)
    st = ro.verify_batch(batch.seals, batch.image_ids, batch.journals)
    seen = set(batch.classes)
    assert {"valid", "tampered", "off_curve", "wrong_subgroup"} <= seen
    for i in range(n):
        if batch.expect[i] is not None:
            assert st[i] == batch.expect[i], (i, batch.classes[i], st[i], batch.expect[i])
        else:
            assert st[i] in (O.ST_OK, O.ST_VERIFICATION_FAILED)


def test_sp1_shape_valid_and_mixed(oracle_backend):
    vk = S.make_vk(oracle_backend, 1, 3, 0xB2000003)
    ovk = oracle_vk(vk)
    n = 64
    batch = S.make_sp1_batch(oracle_backend, vk, n, 0xB2000003, pool=16)
    st = O.sp1_verify_batch(ovk, S.SP1_SELECTOR, batch.vkeys, batch.public_values, batch.proofs)
    assert (st == O.ST_OK).all()
    rng = S.SplitMix64(0xB2000004)
    S.mutate_sp1(batch, oracle_backend, rng)
    st = O.sp1_verify_batch(ovk, S.SP1_SELECTOR, batch.vkeys, batch.public_values, batch.proofs)
    for i in range(n):
        if batch.expect[i] is not None:
            assert st[i] == batch.expect[i], (i, batch.classes[i], st[i], batch.expect[i])


def test_crafted_infinity_accepts(oracle_backend, fx):
    """A = infinity (both as (0,0) and through the (0,Q) quirk) with c = -(alpha beta + x gamma)/delta is ACCEPTED
    for any B (SURVEY.md 8d config 2 note): pairs with an infinity member contribute 1."""
    vk = S.make_vk(oracle_backend, 0, 6, 77)
    ro = O.Risc0Oracle(oracle_vk(vk)); ro.initialize(fx["control_root"], fx["bn254_control_id"])
    im, jd = bytes(range(32)), bytes(range(32, 64))
    sig = S.risc0_signals(fx["control_root"], fx["bn254_control_id"], S.risc0_claim_digest(im, jd, fx["sys0"]))
    t = vk.trap
    c = (-(t["alpha"] * t["beta"] + vk.x_of(sig) * t["gamma"])) * pow(t["delta"], -1, S.R) % S.R
    C = oracle_backend.g1_mul([c])[0]
    B = oracle_backend.g2_mul([12345])[0]
    for A in (bytes(64), S.w32(0) + S.w32(S.P)):
        assert ro.verify(ro.selector() + A + B + C, im, jd) == O.ST_OK
    assert ro.verify(ro.selector() + bytes(64) + bytes(128) + C, im, jd) == O.ST_OK          # B = infinity too
    assert ro.verify(ro.selector() + S.G1_GEN + B + C, im, jd) == O.ST_VERIFICATION_FAILED


def test_pairing4_generator(oracle_backend):
    vk = S.make_vk(oracle_backend, 0, 2, 5)
    g1s, g2s, expect = S.make_pairing4_batch(oracle_backend, vk, 12, 0xB2000005, pool=4)
    blob = b"".join(g1s[i][0:64] + g2s[i] + g1s[i][64:128] + vk.beta + g1s[i][128:192] + vk.gamma + g1s[i][192:256] + vk.delta for i in range(12))
    ok, _, _ = O.pairing4_batch(blob, 12)
    assert list(ok) == expect and 0 in expect and 1 in expect
