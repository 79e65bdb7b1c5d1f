import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle", "pyref")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")


@pytest.fixture(scope="session")
def consts():
    import oracle_lib
    return oracle_lib.constants()


@pytest.fixture(scope="session")
def fx(consts):
    """The two golden fixtures of the reference (examples/*/examples/interact.rs) as bytes."""
    r, s = consts["risc0_fixture"], consts["sp1_fixture"]
    h = bytes.fromhex
    return {
        "control_root": h(r["control_root"]), "bn254_control_id": h(r["bn254_control_id"]), "image_id": h(r["image_id"]),
        "seal": h(r["seal"]), "journal_digest": h(r["journal_digest"]),
        "sp1_vkey": h(s["vkey"]), "sp1_public_values": h(s["public_values"]), "sp1_proof": h(s["proof"]),
        "sp1_selector": h(consts["sp1_verifier_hash"])[:4], "sys0": h(consts["risc0_system_state_zero_digest"]),
    }


class OracleBackend:
    """synth.py backend served by the CPU oracle (tests only)."""

    def g1_mul(self, scalars):
        import oracle_lib as O
        from stylus_zkvm_verifiers_b200.synth import G1_GEN
        return [O.g1_mul(G1_GEN, s) for s in scalars]

    def g2_mul(self, scalars):
        import oracle_lib as O
        from stylus_zkvm_verifiers_b200.synth import G2_GEN
        return [O.g2_mul(G2_GEN, s) for s in scalars]


@pytest.fixture(scope="session")
def oracle_backend():
    return OracleBackend()


def oracle_vk(synth_vk):
    import oracle_lib as O
    return O.Vk(synth_vk.vm, synth_vk.alpha, synth_vk.beta, synth_vk.gamma, synth_vk.delta, synth_vk.ic)
