"""Multi-GPU form of the path: contiguous proof ranges per rank, no data-path collective, one status
byte per proof gathered on rank 0.  Covered here on the CPU with the gloo backend at world_size 2 (and 3, ragged): each
rank verifies its own range (with the oracle standing in for the GPU, this is a test of the host logic only) and the
gathered bytes must equal the unsharded result."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from stylus_zkvm_verifiers_b200 import sharding as SH


def test_shard_ranges_cover_and_match_for_each_device():
    for n in (0, 1, 2, 7, 4095, 4096, 65536, (1 << 20) + 3):
        for world in (1, 2, 3, 4, 8):
            r = SH.shard_ranges(n, world)
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[g][1] == r[g + 1][0] for g in range(world - 1))
            sizes = [e - b for b, e in r]
            assert max(sizes) - min(sizes) <= 1
            # csrc/zkv.cu for_each_device: b = n*d/nd, e = n*(d+1)/nd
            assert r == [(n * d // world, n * (d + 1) // world) for d in range(world)]
    with pytest.raises(ValueError):
        SH.shard_range(10, 2, 2)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n, seed, q):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    import oracle_lib as O
    from conftest import OracleBackend, oracle_vk
    from stylus_zkvm_verifiers_b200 import synth as S
    dist.init_process_group("gloo", rank=rank, world_size=world)
    c = O.constants(); h = bytes.fromhex; r = c["risc0_fixture"]
    be = OracleBackend()
    vk = S.make_vk(be, 0, 6, seed)                      # same seed on every rank -> same key and batch
    ro = O.Risc0Oracle(oracle_vk(vk)); ro.initialize(h(r["control_root"]), h(r["bn254_control_id"]))
    batch = S.make_risc0_batch(be, vk, ro.selector(), h(r["control_root"]), h(r["bn254_control_id"]), h(c["risc0_system_state_zero_digest"]), n, seed, pool=8)
    S.mutate_risc0(batch, be, S.SplitMix64(seed + 1))
    verify = lambda b, e: ro.verify_batch(batch.seals[b:e], batch.image_ids[b:e], batch.journals[b:e])
    got = SH.verify_sharded(verify, n, dist, dst=0)
    if rank == 0:
        want = ro.verify_batch(batch.seals, batch.image_ids, batch.journals)
        q.put((got.tolist(), np.asarray(want).tolist()))
    else:
        assert got is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 24), (3, 19)])
def test_sharded_verify_matches_unsharded_gloo(world, n):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(g, world, port, n, 0xB2000004, q)) for g in range(world)]
    for p in procs:
        p.start()
    got, want = q.get(timeout=240)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert got == want
    assert len(set(want)) > 1, "the mixed batch should contain accepted and rejected proofs"
