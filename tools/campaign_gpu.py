#!/usr/bin/env python3
"""GPU half of the exactness campaign (see tools/campaign_common.py): python tools/campaign_gpu.py --first 0 --count 160 --out gpurun_out/campaign
Writes <out>/status_<k>.bin (one status byte per proof) and <out>/manifest.json (per batch: shape, input fingerprint, class histogram,
statuses checked against what the mutation guarantees by construction)."""
import argparse, collections, json, os, sys, time
import campaign_common as CC
sys.path.insert(0, os.path.join(CC.ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--first", type=int, default=0); ap.add_argument("--count", type=int, default=160)
    ap.add_argument("--out", default="gpurun_out/campaign"); ap.add_argument("--n", type=int, default=CC.BATCH)
    a = ap.parse_args()
    import numpy as np
    import stylus_zkvm_verifiers_b200 as Z
    consts = json.load(open(os.path.join(CC.ROOT, "tests", "golden", "reference_constants.json")))
    h = bytes.fromhex; r = consts["risc0_fixture"]
    os.makedirs(a.out, exist_ok=True)
    gpu = Z.GpuBackend(0)
    vk0, vk1 = CC.keys(gpu)
    kv0 = Z.VerificationKey(0, vk0.alpha, vk0.beta, vk0.gamma, vk0.delta, vk0.ic); v0 = Z.RiscZeroVerifier(kv0); v0.initialize(h(r["control_root"]), h(r["bn254_control_id"]))
    kv1 = Z.VerificationKey(1, vk1.alpha, vk1.beta, vk1.gamma, vk1.delta, vk1.ic); v1 = Z.Sp1Verifier(kv1)
    man = {"base_seed": CC.BASE_SEED, "batch": a.n, "batches": [], "library": os.environ.get("ZKV_LIB", "libzkv_b200.so")}
    t0 = time.time()
    for k in range(a.first, a.first + a.count):
        shape, b, fpr = CC.batch(gpu, k, vk0, vk1, v0.get_selector(), consts, a.n)
        st = v0.verify_batch(b.seals, b.image_ids, b.journals) if shape == "risc0" else v1.verify_batch(b.vkeys, b.public_values, b.proofs)
        st = np.asarray(st, dtype=np.uint8)
        st.tofile(os.path.join(a.out, "status_%04d.bin" % k))
        bad = sum(1 for i in range(a.n) if b.expect[i] is not None and st[i] != b.expect[i])
        man["batches"].append({"k": k, "shape": shape, "fingerprint": fpr, "classes": dict(collections.Counter(b.classes)), "accepted": int((st == 0).sum()),
                               "known_by_construction": sum(1 for e in b.expect if e is not None), "mismatch_vs_construction": bad})
        json.dump(man, open(os.path.join(a.out, "manifest.json"), "w"))
        print("batch %d %s accepted %d bad %d  (%.0f s)" % (k, shape, int((st == 0).sum()), bad, time.time() - t0), flush=True)


if __name__ == "__main__":
    main()
