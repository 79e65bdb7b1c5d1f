#!/usr/bin/env python3
"""Accept/reject exactness campaign (BASELINE config 4): fresh mixed batches (valid / tampered / off-curve / coordinate >= p /
wrong-subgroup G2 / infinity / malformed) of both proof shapes, verified on the GPU and compared
  * with the status the mutation guarantees by construction, for every proof that has one, and
  * 1:1 with the oracle on a random-offset window of every batch (the classes whose outcome is not known by construction are only
    checked there).
usage: python tools/campaign.py --minutes 8 --batch 65536 --window 1024      prints one JSON line with the totals."""
import argparse, json, os, sys, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--minutes", type=float, default=5.0)
    ap.add_argument("--batch", type=int, default=1 << 16)
    ap.add_argument("--window", type=int, default=1024)
    ap.add_argument("--seed", type=lambda s: int(s, 0), default=0xB2000004)
    a = ap.parse_args()
    import oracle_lib as O
    import stylus_zkvm_verifiers_b200 as Z
    from stylus_zkvm_verifiers_b200 import synth as S
    from conftest import oracle_vk
    c = O.constants(); h = bytes.fromhex; r = c["risc0_fixture"]
    gpu = Z.GpuBackend(0)
    vk0 = S.make_vk(gpu, 0, 6, 0xB2000001); kv0 = Z.VerificationKey(0, vk0.alpha, vk0.beta, vk0.gamma, vk0.delta, vk0.ic)
    v0 = Z.RiscZeroVerifier(kv0); v0.initialize(h(r["control_root"]), h(r["bn254_control_id"]))
    ro = O.Risc0Oracle(oracle_vk(vk0)); ro.initialize(h(r["control_root"]), h(r["bn254_control_id"]))
    vk1 = S.make_vk(gpu, 1, 3, 0xB2000003); kv1 = Z.VerificationKey(1, vk1.alpha, vk1.beta, vk1.gamma, vk1.delta, vk1.ic)
    v1 = Z.Sp1Verifier(kv1); ovk1 = oracle_vk(vk1)
    sys0 = h(c["risc0_system_state_zero_digest"])
    tot = {"proofs": 0, "checked_by_construction": 0, "checked_against_oracle": 0, "mismatches": 0, "accepted": 0, "batches": 0}
    t_end = time.time() + 60 * a.minutes
    seed = a.seed
    n, w = a.batch, a.window
    while time.time() < t_end:
        seed += 1
        rng = S.SplitMix64(seed)
        if seed & 1:
            b = S.make_risc0_batch(gpu, vk0, v0.get_selector(), h(r["control_root"]), h(r["bn254_control_id"]), sys0, n, seed, pool=1024)
            S.mutate_risc0(b, gpu, rng)
            st = v0.verify_batch(b.seals, b.image_ids, b.journals)
            o = rng.below(n - w)
            want = ro.verify_batch(b.seals[o:o + w], b.image_ids[o:o + w], b.journals[o:o + w])
        else:
            b = S.make_sp1_batch(gpu, vk1, n, seed, pool=1024)
            S.mutate_sp1(b, gpu, rng)
            st = v1.verify_batch(b.vkeys, b.public_values, b.proofs)
            o = rng.below(n - w)
            want = O.sp1_verify_batch(ovk1, S.SP1_SELECTOR, b.vkeys[o:o + w], b.public_values[o:o + w], b.proofs[o:o + w])
        bad = sum(1 for i in range(n) if b.expect[i] is not None and st[i] != b.expect[i])
        bad += int((st[o:o + w] != want).sum())
        tot["proofs"] += n; tot["batches"] += 1
        tot["checked_by_construction"] += sum(1 for e in b.expect if e is not None)
        tot["checked_against_oracle"] += w
        tot["mismatches"] += bad
        tot["accepted"] += int((st == 0).sum())
    print(json.dumps(tot), flush=True)


if __name__ == "__main__":
    main()
