"""Profiling driver: one short C3 (PsiCNN 10x10) sampling + E_loc call (run under ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import annongpu_b200 as A
from annongpu_b200 import factories as F
A.setDevice(0)
spec, H = F.config_C3()
psi, op = spec.build(True), H.build(True)
chains = 1184
mc = A.MonteCarloSpins(chains, 1, 1, chains, True, seed=1)
ev = A.ExpectationValue(True)
print(ev(op, psi, mc))
