import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import oracle_lib as O
import stylus_zkvm_verifiers_b200 as Z
from stylus_zkvm_verifiers_b200 import synth as S
gpu = Z.GpuBackend(0)
vk = S.make_vk(gpu, 0, 2, 5)
kv = Z.VerificationKey(0, vk.alpha, vk.beta, vk.gamma, vk.delta, vk.ic)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
g1s, g2s, expect = S.make_pairing4_batch(gpu, vk, n, 0xB2000005, pool=4)
ok, gt, ml = Z.pairing4_batch(kv, b"".join(g1s), b"".join(g2s), n, want_gt=True, want_miller=True)
blob = b"".join(g1s[i][0:64] + g2s[i] + g1s[i][64:128] + vk.beta + g1s[i][128:192] + vk.gamma + g1s[i][192:256] + vk.delta for i in range(n))
ook, ogt, oml = O.pairing4_batch(blob, n, want_gt=True, want_miller=True)
print("ok", list(ok), list(ook), expect)
for i in range(n):
    m_ok = ml[384*i:384*i+384].tobytes() == oml[384*i:384*i+384].tobytes(); g_ok = gt[384*i:384*i+384].tobytes() == ogt[384*i:384*i+384].tobytes()
    # does the oracle's final exp of OUR miller value give OUR gt?
    fe = O.final_exp(ml[384*i:384*i+384].tobytes()) == gt[384*i:384*i+384].tobytes() if max(ml[384*i:384*i+384]) else None
    print(i, "miller", m_ok, "gt", g_ok, "fe(our miller)==our gt", fe)
