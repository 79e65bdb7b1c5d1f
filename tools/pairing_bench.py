"""Pairing microbench (BASELINE.json configs[4]): n instances of the 4-pair product check
e(P0,Q) e(P1,beta) e(P2,gamma) e(P3,delta) == 1 through zkv_pairing4_batch_device (unscaled lines, 4-pair Miller loop, final
exponentiation, ok byte per instance), inputs resident in HBM, CUDA events on the launching stream.  The instances are a seeded pool of
`--pool` distinct ones (about half of them satisfy the equation) tiled to n on the device; the ok bytes of every tile must equal the
generator's expectation.  Prints one JSON line: instances/s and the fraction of the IMAD.WIDE issue roofline (W_pair4 of bench.py).

    python tools/pairing_bench.py --n 1048576 --steps 3
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1 << 20)
    ap.add_argument("--pool", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    import numpy as np
    import torch
    assert torch.cuda.is_available(), "the pairing service has no CPU path"
    import bench as B
    import stylus_zkvm_verifiers_b200 as Z
    from stylus_zkvm_verifiers_b200 import _native as N
    from stylus_zkvm_verifiers_b200 import synth as S
    dev, n, pool = 0, args.n, args.pool
    assert n % pool == 0
    torch.cuda.set_device(dev)
    gpu = Z.GpuBackend(dev)
    vk = S.make_vk(gpu, 0, 6, 0xB2000001)
    kv = Z.VerificationKey(0, vk.alpha, vk.beta, vk.gamma, vk.delta, vk.ic)
    g1s, g2s, expect = S.make_pairing4_batch(gpu, vk, pool, 0xB2000005, pool=min(pool, 1024))
    d_g1 = torch.frombuffer(bytearray(b"".join(g1s)), dtype=torch.uint8).cuda().repeat(n // pool)
    d_g2 = torch.frombuffer(bytearray(b"".join(g2s)), dtype=torch.uint8).cuda().repeat(n // pool)
    d_ok = torch.full((n,), 255, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
    sp = stream.cuda_stream

    def launch():
        N.check(N.lib().zkv_pairing4_batch_device(kv._h, dev, d_g1.data_ptr(), d_g2.data_ptr(), n, d_ok.data_ptr(), None, sp))

    for _ in range(max(args.warmup, 3)):
        launch()
    torch.cuda.synchronize()
    want = torch.tensor(expect, dtype=torch.uint8, device="cuda").repeat(n // pool)
    assert torch.equal(d_ok, want), "ok bytes differ from the generator's expectation"
    imad_peak, _ = Z.imad_peak(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ms = 0.0
    for k in range(args.steps):
        flush.fill_(k)                                         # > 126 MB L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream); launch(); e1.record(stream)
        torch.cuda.synchronize()
        ms += e0.elapsed_time(e1)
    assert torch.equal(d_ok, want)
    rate = n * args.steps / (ms * 1e-3)
    w = (B.W_MILLER4_M + B.W_FINALEXP_M) * B.M_MAC32
    print(json.dumps({"metric": "pairing4_instances_per_sec", "value": rate, "unit": "instances/s", "n_gpus": 1, "steps": args.steps,
                      "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "dtype": "u256 (8x32-bit Montgomery limbs, IMAD.WIDE.U32)",
                      "config": {"workload": "configs[4]: 2^%d 4-pair multi-Miller loop + final exponentiation instances (1 variable + 3 fixed G2), pool of %d distinct instances tiled on the device, %d accepted per tile"
                                 % (n.bit_length() - 1, pool, sum(1 for e in expect if e == 1)), "l2": "flushed between timed steps (256 MiB fill)"},
                      "roofline": {"bound": "imad", "achieved": rate * w / 1e12, "peak": imad_peak / 1e12, "unit": "TMAC32/s", "frac": rate * w / imad_peak,
                                   "mac32_per_instance": w}}))


if __name__ == "__main__":
    main()
