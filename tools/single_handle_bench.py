#!/usr/bin/env python3
"""ONE process, ONE handle over N GPUs (north star: "partitioned across the GPUs ... results gathered on the host"): the literal
BASELINE.json configs[2] and configs[4] through the host-buffer C ABI, strong scaling (the batch is fixed, the devices grow).

    python tools/single_handle_bench.py --config sp1 --n 1048576 --gpus 8       # 2^20 SP1-shape proofs, one zkv_sp1_verify_batch call per step
    python tools/single_handle_bench.py --config pairing --n 4194304 --gpus 8   # 2^22 4-pair instances, one zkv_pairing4_batch call per step
    python tools/single_handle_bench.py --config sp1 --sweep 1,2,4,8            # one line per device count, same batch

A step is one host call: the library cuts the batch into contiguous per-device ranges (for_each_device, csrc/zkv.cu), one host thread per
device stages, uploads, runs the kernel chains and downloads; the status / ok bytes are gathered into the caller's array.  Timed with the
host clock around the call (everything is inside: host staging, H2D, kernels, D2H, gather).  Every line checks all result bytes against the
generator's expectation (valid proofs accepted, the tampered ones rejected).  Prints one JSON line per device count.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="sp1", choices=["sp1", "risc0", "pairing"])
    ap.add_argument("--n", type=int, default=0)
    ap.add_argument("--gpus", type=int, default=0, help="devices of the handle (0 = all)")
    ap.add_argument("--sweep", default="", help="comma-separated device counts, one line each")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--pool", type=int, default=4096)
    ap.add_argument("--pageable", action="store_true", help="keep the input arrays in pageable memory (default: page-locked, uploaded in place)")
    args = ap.parse_args()
    import numpy as np
    import torch
    assert torch.cuda.is_available(), "no CPU path"
    import stylus_zkvm_verifiers_b200 as Z
    from stylus_zkvm_verifiers_b200 import _native as N
    from stylus_zkvm_verifiers_b200 import synth as S
    consts = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_constants.json")))
    h = bytes.fromhex
    have = Z.device_count()
    counts = [int(x) for x in args.sweep.split(",")] if args.sweep else [args.gpus or have]
    counts = [c for c in counts if c <= have]
    n = args.n or ((1 << 22) if args.config == "pairing" else (1 << 20))
    gpu = Z.GpuBackend(0)
    pool = min(args.pool, n)
    assert n % pool == 0
    reps = n // pool
    rng = S.SplitMix64(0xB2000031)

    if args.config == "pairing":
        vk = S.make_vk(gpu, 0, 6, 0xB2000001)
        g1s, g2s, expect = S.make_pairing4_batch(gpu, vk, pool, 0xB2000005, pool=min(pool, 1024))
        h_g1 = np.tile(np.frombuffer(b"".join(g1s), dtype=np.uint8), reps); h_g2 = np.tile(np.frombuffer(b"".join(g2s), dtype=np.uint8), reps)
        want = np.tile(np.asarray(expect, dtype=np.uint8), reps)
        in_bytes = h_g1.nbytes + h_g2.nbytes

        def make(devs):
            kv = Z.VerificationKey(0, vk.alpha, vk.beta, vk.gamma, vk.delta, vk.ic, devices=devs)
            out = np.zeros(n, dtype=np.uint8)
            return kv, lambda: N.check(N.lib().zkv_pairing4_batch(kv._h, N.buf(h_g1), N.buf(h_g2), n, out.ctypes.data, None, None)), out
        metric, unit, label = "pairing4_instances_per_sec", "instances/s", "configs[4]: 2^%d 4-pair product checks (1 variable + 3 fixed G2), pool of %d distinct instances tiled" % (n.bit_length() - 1, pool)
    else:
        r = consts["risc0_fixture"]
        if args.config == "sp1":
            vk = S.make_vk(gpu, 1, 3, 0xB2000003)
            b = S.make_sp1_batch(gpu, vk, pool, 0xB2000003, pool=min(pool, 4096))
            proofs, a32, pvs = list(b.proofs), list(b.vkeys), list(b.public_values)
        else:
            vk = S.make_vk(gpu, 0, 6, 0xB2000001)
            v0 = Z.RiscZeroVerifier(Z.VerificationKey(0, vk.alpha, vk.beta, vk.gamma, vk.delta, vk.ic, devices=[0]), devices=[0]); v0.initialize(h(r["control_root"]), h(r["bn254_control_id"]))
            b = S.make_risc0_batch(gpu, vk, v0.get_selector(), h(r["control_root"]), h(r["bn254_control_id"]), h(consts["risc0_system_state_zero_digest"]), pool, 0xB2000001, pool=min(pool, 4096))
            proofs, a32, pvs = list(b.seals), list(b.image_ids), list(b.journals)
        exp = np.zeros(pool, dtype=np.uint8)
        for i in range(0, pool, 61):                      # tamper a byte of C.x in every 61st proof of the pool: must come back rejected
            p = bytearray(proofs[i]); p[4 + 200] ^= 1; proofs[i] = bytes(p); exp[i] = 1
        h_pr = np.tile(np.frombuffer(b"".join(proofs), dtype=np.uint8), reps); h_po = np.arange(n + 1, dtype=np.uint64) * 260
        h_a = np.tile(np.frombuffer(b"".join(a32), dtype=np.uint8), reps)
        pvl = len(pvs[0]); assert all(len(x) == pvl for x in pvs)
        h_pv = np.tile(np.frombuffer(b"".join(pvs), dtype=np.uint8), reps); h_vo = np.arange(n + 1, dtype=np.uint64) * pvl
        want = np.tile(exp, reps)
        if not args.pageable:
            h_pr, h_po, h_a, h_pv, h_vo = [Z.pinned_copy(x) for x in (h_pr, h_po, h_a, h_pv, h_vo)]
        in_bytes = h_pr.nbytes + h_a.nbytes + h_pv.nbytes + h_po.nbytes + (h_vo.nbytes if args.config == "sp1" else 0)

        def make(devs):
            out = np.zeros(n, dtype=np.uint8)
            if args.config == "sp1":
                v = Z.Sp1Verifier(Z.VerificationKey(1, vk.alpha, vk.beta, vk.gamma, vk.delta, vk.ic, devices=devs), devices=devs)
                return v, lambda: v.verify_batch_packed(h_a, h_pv, h_vo, h_pr, h_po, n, out), out
            v = Z.RiscZeroVerifier(Z.VerificationKey(0, vk.alpha, vk.beta, vk.gamma, vk.delta, vk.ic, devices=devs), devices=devs)
            v.initialize(h(r["control_root"]), h(r["bn254_control_id"]))
            return v, lambda: v.verify_batch_packed(h_pr, h_po, h_a, h_pv, n, out), out
        metric, unit = "groth16_verifies_per_sec", "verifies/s"
        label = "%s: 2^%d %s-shape proofs in ONE host-buffer batch call, pool of %d distinct proofs tiled, 1 in 61 tampered" % (
            "configs[2]" if args.config == "sp1" else "configs[1] shape", n.bit_length() - 1, "SP1 v5" if args.config == "sp1" else "RISC Zero", pool)

    base = None
    for nd in counts:
        devs = list(range(nd))
        keep, call, out = make(devs)
        for _ in range(max(args.warmup, 1)):
            call()
        good = (out == want).all() if args.config == "pairing" else ((out != 0).astype(np.uint8) == want).all()
        assert good, "result bytes differ from the generator's expectation"
        out[:] = 255
        t0 = time.perf_counter()
        for _ in range(args.steps):
            call()
        dt = time.perf_counter() - t0
        if args.config == "pairing":
            assert (out == want).all()
        else:
            assert ((out != 0).astype(np.uint8) == want).all()
        rate = n * args.steps / dt
        base = base or (rate, nd)
        print(json.dumps({"metric": metric, "value": rate, "unit": unit, "n_gpus": nd, "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": 1e3 * dt / args.steps,
                          "higher_is_better": True, "scaling": "strong", "single_process": True,
                          "config": {"workload": label, "items_per_step": n, "sharding": "one handle over devices %s: contiguous ranges, one host thread per device, status bytes gathered on the host, no collective" % devs},
                          "e2e": {"value": rate, "unit": unit, "h2d_bytes_per_step": int(in_bytes), "d2h_bytes_per_step": int(n)},
                          "timing": "host clock around the C-ABI call (H2D, kernels, D2H and gather inside; inputs in %s host memory)" % ("pageable" if args.pageable else "page-locked"),
                          "speedup_vs_first_line": rate / base[0] * 1.0, "first_line_gpus": base[1]}), flush=True)
        del keep, call


if __name__ == "__main__":
    main()
