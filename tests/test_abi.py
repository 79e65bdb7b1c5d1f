"""The C-ABI boundary without a GPU: the library loads, exports every symbol include/zkv.h declares, and
fails loudly (no CPU fallback) when asked to verify without a CUDA device."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "zkv.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(zkv_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    from stylus_zkvm_verifiers_b200 import _native as N
    L = N.lib()
    names = header_functions()
    assert len(names) >= 35
    for n in names:
        assert hasattr(L, n), "libzkv_b200.so does not export %s" % n
    assert set(names) == set(N.EXPORTS), set(names) ^ set(N.EXPORTS)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    import stylus_zkvm_verifiers_b200 as Z
    assert Z._native.lib().zkv_device_count() == 0
    with pytest.raises(Z.ZkvError) as e:
        Z.RiscZeroVerifier()
    assert e.value.code == Z._native.ZKV_ERR_CUDA
    with pytest.raises(Z.ZkvError):
        Z.Sp1Verifier()
    with pytest.raises(Z.ZkvError):
        Z.fp_mul_batch(bytes(32), bytes(32), 1)


def test_product_never_imports_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may touch oracle/."""
    pkg = os.path.join(ROOT, "stylus_zkvm_verifiers_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle_lib" not in txt and "libzkv_oracle" not in txt and "zkvo_" not in txt, f


def test_error_payloads():
    from stylus_zkvm_verifiers_b200 import errors as E
    assert E.VerificationFailed().payload == E.keccak256(b"VerificationFailed()")[:4]
    p = E.SelectorMismatch(b"\x01\x02\x03\x04", b"\x9f\x39\x69\x6c").payload
    assert len(p) == 68 and p[4:8] == b"\x01\x02\x03\x04" and p[8:36] == bytes(28) and p[36:40] == b"\x9f\x39\x69\x6c"
    assert E.WrongVerifierSelector(b"abcd", b"efgh").payload[:4] == E.keccak256(b"WrongVerifierSelector(bytes4,bytes4)")[:4]


def test_tools_and_bench_parse():
    """bench.py and the measurement tools under tools/ are run on the GPU box only: keep them at least syntactically valid here."""
    import ast, glob
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for path in [os.path.join(root, "bench.py"), os.path.join(root, "__graft_entry__.py")] + sorted(glob.glob(os.path.join(root, "tools", "*.py"))):
        with open(path) as f:
            ast.parse(f.read(), filename=path)
