"""Profiling driver (run under ncu): a few C1 gradient calls (ExactSummation, PsiRBM 16x32, TFIM ring)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import annongpu_b200 as A
from annongpu_b200 import factories as F

A.setDevice(0)
spec, H = F.config_C1()
psi, op, es = spec.build(True), H.build(True), A.ExactSummationSpins(16, True)
psi.normalize(es)
ev = A.ExpectationValue(True)
for _ in range(3):
    g, E = ev.gradient(op, psi, es)
print(E)
