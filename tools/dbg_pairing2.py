import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import oracle_lib as O
import stylus_zkvm_verifiers_b200 as Z
from stylus_zkvm_verifiers_b200 import synth as S
gpu = Z.GpuBackend(0)
vk = S.make_vk(gpu, 0, 2, 5)
kv = Z.VerificationKey(0, vk.alpha, vk.beta, vk.gamma, vk.delta, vk.ic)
g1s, g2s, expect = S.make_pairing4_batch(gpu, vk, 1, 0xB2000005, pool=4)
inf = bytes(64)
cases = {"all4": g1s[0], "only_var": g1s[0][:64] + inf * 3, "only_beta": inf + g1s[0][64:128] + inf * 2, "only_gamma": inf * 2 + g1s[0][128:192] + inf, "only_delta": inf * 3 + g1s[0][192:256]}
for name, g1 in cases.items():
    ok, gt, ml = Z.pairing4_batch(kv, g1, g2s[0], 1, want_gt=True, want_miller=True)
    blob = g1[0:64] + g2s[0] + g1[64:128] + vk.beta + g1[128:192] + vk.gamma + g1[192:256] + vk.delta
    ook, ogt, oml = O.pairing4_batch(blob, 1, want_gt=True, want_miller=True)
    print(name, "miller", ml.tobytes() == oml.tobytes(), "gt", gt.tobytes() == ogt.tobytes(), int(ok[0]), int(ook[0]))
