"""C4 dense SR step timing: eval (exact / tensor-core S) + dense solve; prints phase times."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import annongpu_b200 as A
from annongpu_b200 import factories as F
A.setDevice(0)
spec, H = F.config_C4()
psi, op = spec.build(True), H.build(True)
mc = A.MonteCarloSpins(16384, 1, 10, 16384, True, seed=4)
t = A.TDVP(psi.num_params, True); t.set_profile(True)
for tol in (1e-5, 0.0):
    for rep in range(3):
        t.eval(op, psi, mc, s_tolerance=tol)
        ph_eval = t.phase_ms
        x = t.solve(shift_rel=1e-3)
        ph = t.phase_ms
    S = t.S_matrix; Fv = t.F_vector
    d = np.real(np.diag(S))
    r = (S + np.diag(1e-3 * d)) @ x - Fv
    print(json.dumps({"s_tolerance": tol, "eval_total_ms": ph_eval["total"], "s_build_ms": ph_eval["s_build"], "solve_ms": ph["solve"],
                      "rel_residual": float(np.linalg.norm(r) / np.linalg.norm(Fv)), "P": psi.num_params}))
