"""Profiling driver: one C2 eval_F followed by a few S.v products (run under ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import annongpu_b200 as A
from annongpu_b200 import factories as F
A.setDevice(0)
cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
spec, H = F.config_C2() if cfg == "C2" else F.config_C5()
chains = 8192 if cfg == "C2" else 4096
psi, op = spec.build(True), H.build(True)
mc = A.MonteCarloSpins(chains, 1, 2, chains, True, seed=1)
t = A.TDVP(psi.num_params, True)
t.eval_F(op, psi, mc)
v = np.random.default_rng(0).normal(size=psi.num_params) + 0j
for _ in range(4):
    t.S_dot_vector(v)
x, it, rr = t.solve_cg(tol=1e-6, max_iter=16, shift_rel=1e-3)
print("done", it, rr)
