import sys; sys.path.insert(0,'/root/repo')
import stylus_zkvm_verifiers_b200 as Z
print(Z.imad_peak(0))
