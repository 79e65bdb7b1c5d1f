#!/usr/bin/env python3
"""GPU half of the exactness campaign (see tools/campaign_common.py):
  python tools/campaign_gpu.py --first 0 --count 154 --procs 8 --out gpurun_out/campaign
Writes <out>/status_<k>.bin (one status byte per proof) and <out>/manifest.json (per batch: shape, input fingerprint, class histogram,
statuses checked against what the mutation guarantees by construction).  Batch generation is host-side Python, so --procs worker
processes (each with its own CUDA context on device 0) build and verify batches k = w, w + procs, ... side by side."""
import argparse, collections, json, os, sys, time
import campaign_common as CC
sys.path.insert(0, os.path.join(CC.ROOT, "tests"))


def worker(args):
    w, procs, first, count, out, n = args
    import numpy as np
    import stylus_zkvm_verifiers_b200 as Z
    consts = json.load(open(os.path.join(CC.ROOT, "tests", "golden", "reference_constants.json")))
    h = bytes.fromhex; r = consts["risc0_fixture"]
    gpu = Z.GpuBackend(0)
    vk0, vk1 = CC.keys(gpu)
    kv0 = Z.VerificationKey(0, vk0.alpha, vk0.beta, vk0.gamma, vk0.delta, vk0.ic); v0 = Z.RiscZeroVerifier(kv0); v0.initialize(h(r["control_root"]), h(r["bn254_control_id"]))
    kv1 = Z.VerificationKey(1, vk1.alpha, vk1.beta, vk1.gamma, vk1.delta, vk1.ic); v1 = Z.Sp1Verifier(kv1)
    rows = []
    t0 = time.time()
    for k in range(first + w, first + count, procs):
        shape, b, fpr = CC.batch(gpu, k, vk0, vk1, v0.get_selector(), consts, n)
        st = v0.verify_batch(b.seals, b.image_ids, b.journals) if shape == "risc0" else v1.verify_batch(b.vkeys, b.public_values, b.proofs)
        st = np.asarray(st, dtype=np.uint8)
        st.tofile(os.path.join(out, "status_%04d.bin" % k))
        bad = sum(1 for i in range(n) if b.expect[i] is not None and st[i] != b.expect[i])
        rows.append({"k": k, "shape": shape, "fingerprint": fpr, "classes": dict(collections.Counter(b.classes)), "accepted": int((st == 0).sum()),
                     "known_by_construction": sum(1 for e in b.expect if e is not None), "mismatch_vs_construction": bad})
        json.dump(rows, open(os.path.join(out, "manifest_part_%d.json" % w), "w"))
        print("worker %d batch %d %s accepted %d bad %d  (%.0f s)" % (w, k, shape, int((st == 0).sum()), bad, time.time() - t0), flush=True)
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--first", type=int, default=0); ap.add_argument("--count", type=int, default=154)
    ap.add_argument("--out", default="gpurun_out/campaign"); ap.add_argument("--n", type=int, default=CC.BATCH)
    ap.add_argument("--procs", type=int, default=1)
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    jobs = [(w, a.procs, a.first, a.count, a.out, a.n) for w in range(a.procs)]
    t0 = time.time()
    if a.procs == 1:
        parts = [worker(jobs[0])]
    else:
        import multiprocessing as mp
        with mp.get_context("spawn").Pool(a.procs) as pool:
            parts = pool.map(worker, jobs)
    rows = sorted((r for p in parts for r in p), key=lambda r: r["k"])
    man = {"base_seed": CC.BASE_SEED, "batch": a.n, "batches": rows, "library": os.environ.get("ZKV_LIB", "libzkv_b200.so"),
           "seconds": time.time() - t0, "procs": a.procs}
    prev = os.path.join(a.out, "manifest.json")
    if os.path.exists(prev):   # extend an earlier run of other batch indices
        old = json.load(open(prev))
        have = {r["k"] for r in rows}
        man["batches"] = sorted([r for r in old["batches"] if r["k"] not in have] + rows, key=lambda r: r["k"])
    json.dump(man, open(prev, "w"))
    print("campaign: %d batches, %d proofs, %d mismatches vs construction, %.0f s" % (len(rows), len(rows) * a.n, sum(r["mismatch_vs_construction"] for r in rows), time.time() - t0))


if __name__ == "__main__":
    main()
