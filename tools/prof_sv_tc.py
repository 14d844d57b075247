"""Profiling driver (run under ncu): one exact and one tensor-core factorised S.v at C2 or a C5 shard."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import annongpu_b200 as A
from annongpu_b200 import factories as F
A.setDevice(0)
name = sys.argv[1] if len(sys.argv) > 1 else "C2"
(spec, H), chains = (F.config_C2(), 8192) if name == "C2" else (F.config_C5(), 16384)
psi, op = spec.build(True), H.build(True)
mc = A.MonteCarloSpins(chains, 1, 1, chains, True, seed=11)
t = A.TDVP(psi.num_params, True)
t.eval_F(op, psi, mc)
v = np.ones(psi.num_params, dtype=complex)
for rep in range(2):
    a = t.S_dot_vector(v)
    b = t.set_tensorcore_products(True).S_dot_vector(v)
    t.set_tensorcore_products(False)
print(np.linalg.norm(a - b) / np.linalg.norm(a))
