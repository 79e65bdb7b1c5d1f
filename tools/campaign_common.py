"""Shared by the two halves of the round-2 exactness campaign (BASELINE config 4, north star: >= 10^7 mixed proofs compared 1:1 with
the reference path's CPU restatement):
  tools/campaign_gpu.py     on the GPU box: builds batch k from seed base+k with the library's own GPU ecMul services, mutates it into
                            the mixed classes, verifies it through the C ABI and writes the status bytes (1 B per proof);
  tools/campaign_oracle.py  on a CPU box: rebuilds the SAME batch k from the same seed with the oracle's ecMul (test infrastructure),
                            runs every proof through the oracle and compares 1:1 with the GPU's bytes.
Both sides call batch() below, so the inputs are identical by construction (exact arithmetic, seeded SplitMix64); the input bytes are
additionally fingerprinted (SHA-256 over seals / proofs and public inputs) on both sides and the fingerprints must agree."""
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
BASE_SEED = 0xB2020000
BATCH = 1 << 16
KEY_SEEDS = (0xB2000001, 0xB2000003)


def keys(backend):
    from stylus_zkvm_verifiers_b200 import synth as S
    return S.make_vk(backend, 0, 6, KEY_SEEDS[0]), S.make_vk(backend, 1, 3, KEY_SEEDS[1])


def batch(backend, k, vk0, vk1, selector0, consts, n=BATCH):
    """batch k: RISC Zero shape for even k, SP1 shape for odd k; returns (shape, batch, fingerprint)"""
    from stylus_zkvm_verifiers_b200 import synth as S
    h = bytes.fromhex
    r = consts["risc0_fixture"]
    seed = BASE_SEED + k
    rng = S.SplitMix64(seed ^ 0x5EED)
    fp = hashlib.sha256()
    if k % 2 == 0:
        b = S.make_risc0_batch(backend, vk0, selector0, h(r["control_root"]), h(r["bn254_control_id"]), h(consts["risc0_system_state_zero_digest"]), n, seed, pool=1024)
        S.mutate_risc0(b, backend, rng)
        for x in (b.seals, b.image_ids, b.journals):
            for y in x:
                fp.update(len(y).to_bytes(4, "little")); fp.update(y)
        return "risc0", b, fp.hexdigest()
    b = S.make_sp1_batch(backend, vk1, n, seed, pool=1024)
    S.mutate_sp1(b, backend, rng)
    for x in (b.proofs, b.vkeys, b.public_values):
        for y in x:
            fp.update(len(y).to_bytes(4, "little")); fp.update(y)
    return "sp1", b, fp.hexdigest()
