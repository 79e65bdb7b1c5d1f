#!/usr/bin/env python3
"""CPU half of the exactness campaign (see tools/campaign_common.py).  TEST INFRASTRUCTURE: runs the oracle, never shipped or timed.
  python tools/campaign_oracle.py compute --first 0 --count 160 --out gpurun_out/campaign_oracle [--procs 4]
      rebuilds batch k with the oracle's ecMul and stores the oracle's status bytes + input fingerprint + class histogram
  python tools/campaign_oracle.py compare --gpu gpurun_out/campaign --oracle gpurun_out/campaign_oracle --json profiles/r2_campaign_oracle.json
      1:1 comparison of every status byte, per mutation class"""
import argparse, collections, json, os, sys, time
import campaign_common as CC
sys.path.insert(0, os.path.join(CC.ROOT, "tests"))


class OracleBackend:
    def g1_mul(self, sc):
        import oracle_lib as O
        from stylus_zkvm_verifiers_b200 import synth as S
        return [O.g1_mul(S.G1_GEN, s) for s in sc]

    def g2_mul(self, sc):
        import oracle_lib as O
        from stylus_zkvm_verifiers_b200 import synth as S
        return [O.g2_mul(S.G2_GEN, s) for s in sc]


def compute_one(args):
    k, out, n = args
    path = os.path.join(out, "oracle_%04d.json" % k)
    if os.path.exists(path):
        return k, 0.0
    import numpy as np
    import oracle_lib as O
    from stylus_zkvm_verifiers_b200 import synth as S
    consts = O.constants(); h = bytes.fromhex; r = consts["risc0_fixture"]
    t0 = time.time()
    ob = OracleBackend()
    vk0, vk1 = CC.keys(ob)
    ro = O.Risc0Oracle(O.Vk(0, vk0.alpha, vk0.beta, vk0.gamma, vk0.delta, vk0.ic)); ro.initialize(h(r["control_root"]), h(r["bn254_control_id"]))
    shape, b, fpr = CC.batch(ob, k, vk0, vk1, ro.selector(), consts, n)
    if shape == "risc0":
        st = ro.verify_batch(b.seals, b.image_ids, b.journals)
    else:
        st = O.sp1_verify_batch(O.Vk(1, vk1.alpha, vk1.beta, vk1.gamma, vk1.delta, vk1.ic), S.SP1_SELECTOR, b.vkeys, b.public_values, b.proofs)
    st = np.asarray(st, dtype=np.uint8)
    st.tofile(os.path.join(out, "oracle_%04d.bin" % k))
    cls = np.asarray([CLASS_ID[c] for c in b.classes], dtype=np.uint8); cls.tofile(os.path.join(out, "classes_%04d.bin" % k))
    json.dump({"k": k, "shape": shape, "fingerprint": fpr, "accepted": int((st == 0).sum()), "seconds": time.time() - t0}, open(path, "w"))
    return k, time.time() - t0


CLASSES = ("valid", "tampered", "off_curve", "coord_ge_p", "wrong_subgroup", "infinity", "malformed")
CLASS_ID = {c: i for i, c in enumerate(CLASSES)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", choices=["compute", "compare", "digest", "check"])
    ap.add_argument("--digests", default="profiles/r2_campaign_oracle_digests.json")
    ap.add_argument("--first", type=int, default=0); ap.add_argument("--count", type=int, default=160); ap.add_argument("--n", type=int, default=CC.BATCH)
    ap.add_argument("--out", default="gpurun_out/campaign_oracle"); ap.add_argument("--procs", type=int, default=4)
    ap.add_argument("--gpu", default="gpurun_out/campaign"); ap.add_argument("--oracle", default="gpurun_out/campaign_oracle"); ap.add_argument("--json", default="profiles/r2_campaign_oracle.json")
    a = ap.parse_args()
    if a.mode == "compute":
        os.makedirs(a.out, exist_ok=True)
        os.environ.setdefault("OMP_NUM_THREADS", str(max(1, (os.cpu_count() or 1) // a.procs)))
        import multiprocessing as mp
        with mp.Pool(a.procs) as pool:
            for k, dt in pool.imap_unordered(compute_one, [(k, a.out, a.n) for k in range(a.first, a.first + a.count)]):
                print("oracle batch %d done in %.0f s" % (k, dt), flush=True)
        return
    import hashlib
    import numpy as np
    if a.mode == "digest":
        # the oracle's status bytes are 64 KiB per batch (10 MB for the campaign): what is committed is their SHA-256 per batch, with the input
        # fingerprint, the accept count and the per-class accept / reject counts, so that the GPU half can be re-checked after any rebuild
        # (mode `check`) without repeating the 2.5 CPU-hours of the oracle half
        out = {"base_seed": CC.BASE_SEED, "batch": a.n, "classes": list(CLASSES), "batches": []}
        k = 0
        while os.path.exists(os.path.join(a.oracle, "oracle_%04d.json" % k)):
            oj = json.load(open(os.path.join(a.oracle, "oracle_%04d.json" % k)))
            o = np.fromfile(os.path.join(a.oracle, "oracle_%04d.bin" % k), dtype=np.uint8)
            c = np.fromfile(os.path.join(a.oracle, "classes_%04d.bin" % k), dtype=np.uint8)
            out["batches"].append({"k": k, "shape": oj["shape"], "fingerprint": oj["fingerprint"], "status_sha256": hashlib.sha256(o.tobytes()).hexdigest(),
                                   "accepted": int((o == 0).sum()), "per_class": [[int((c == i).sum()), int(((c == i) & (o == 0)).sum())] for i in range(len(CLASSES))],
                                   "status_histogram": {str(int(x)): int(n) for x, n in zip(*np.unique(o, return_counts=True))}})
            k += 1
        json.dump(out, open(a.digests, "w"))
        print("wrote %s: %d batches, %d proofs" % (a.digests, k, k * a.n))
        return
    if a.mode == "check":
        dg = json.load(open(a.digests)); man = json.load(open(os.path.join(a.gpu, "manifest.json")))
        byk = {b["k"]: b for b in dg["batches"]}
        tot = {"proofs": 0, "batches": 0, "batches_equal_to_oracle": 0, "fingerprint_mismatches": 0, "accepted": 0, "library": man.get("library")}
        for e in man["batches"]:
            b = byk.get(e["k"])
            if b is None:
                continue
            g = np.fromfile(os.path.join(a.gpu, "status_%04d.bin" % e["k"]), dtype=np.uint8)
            tot["batches"] += 1; tot["proofs"] += len(g); tot["accepted"] += int((g == 0).sum())
            tot["fingerprint_mismatches"] += int(b["fingerprint"] != e["fingerprint"])
            tot["batches_equal_to_oracle"] += int(hashlib.sha256(g.tobytes()).hexdigest() == b["status_sha256"] and b["fingerprint"] == e["fingerprint"])
        tot["mismatching_batches"] = tot["batches"] - tot["batches_equal_to_oracle"]
        tot["how"] = "SHA-256 of each batch's GPU status bytes against the committed digest of the oracle's status bytes for the same seeded batch (%s)" % a.digests
        json.dump(tot, open(a.json, "w"), indent=1)
        print(json.dumps(tot))
        return
    man = json.load(open(os.path.join(a.gpu, "manifest.json")))
    tot = {"proofs": 0, "compared_1_to_1_with_oracle": 0, "mismatches": 0, "accepted": 0, "batches": 0, "fingerprint_mismatches": 0,
           "per_class": {c: {"proofs": 0, "accepted": 0, "mismatches": 0} for c in CLASSES}, "per_shape": collections.Counter(), "status_histogram": collections.Counter()}
    for e in man["batches"]:
        k = e["k"]
        op = os.path.join(a.oracle, "oracle_%04d.json" % k)
        if not os.path.exists(op):
            continue
        oj = json.load(open(op))
        g = np.fromfile(os.path.join(a.gpu, "status_%04d.bin" % k), dtype=np.uint8)
        o = np.fromfile(os.path.join(a.oracle, "oracle_%04d.bin" % k), dtype=np.uint8)
        c = np.fromfile(os.path.join(a.oracle, "classes_%04d.bin" % k), dtype=np.uint8)
        assert len(g) == len(o) == len(c)
        if oj["fingerprint"] != e["fingerprint"]:
            tot["fingerprint_mismatches"] += 1
            continue
        tot["batches"] += 1; tot["proofs"] += len(g); tot["compared_1_to_1_with_oracle"] += len(g); tot["per_shape"][e["shape"]] += len(g)
        tot["mismatches"] += int((g != o).sum()); tot["accepted"] += int((g == 0).sum())
        for i, name in enumerate(CLASSES):
            m = c == i
            pc = tot["per_class"][name]; pc["proofs"] += int(m.sum()); pc["accepted"] += int((g[m] == 0).sum()); pc["mismatches"] += int((g[m] != o[m]).sum())
        for s, cnt in zip(*np.unique(g, return_counts=True)):
            tot["status_histogram"][str(int(s))] += int(cnt)
    tot["per_shape"] = dict(tot["per_shape"]); tot["status_histogram"] = dict(tot["status_histogram"])
    tot["library"] = man.get("library"); tot["base_seed"] = man["base_seed"]
    tot["how"] = "tools/campaign_gpu.py on a B200 (status bytes through the C ABI), tools/campaign_oracle.py on CPU (oracle/zkv_oracle.c on inputs rebuilt from the same seeds; input fingerprints agree)"
    json.dump(tot, open(a.json, "w"), indent=1)
    print(json.dumps(tot))


if __name__ == "__main__":
    main()
