"""Profiling driver (run under ncu): one reduced pass of C2 (eval_F + 2 CG iterations), C3 and C4 (exact S build)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import annongpu_b200 as A
from annongpu_b200 import factories as F

A.setDevice(0)
todo = sys.argv[1:] or ["C2", "C3", "C4"]
if "C2" in todo:
    spec, H = F.config_C2()
    psi, op = spec.build(True), H.build(True)
    mc = A.MonteCarloSpins(8192, 1, 10, 8192, True, seed=2)
    t = A.TDVP(psi.num_params, True)
    t.eval_F(op, psi, mc)
    print("C2", t.E_local, t.solve_cg(tol=1e-30, max_iter=2, shift_abs=0.0, shift_rel=1e-3)[1:])
if "C3" in todo:
    spec, H = F.config_C3()
    psi, op = spec.build(True), H.build(True)
    mc = A.MonteCarloSpins(2368, 1, 1, 2368, True, seed=3)
    t = A.TDVP(psi.num_params, True)
    t.eval_F(op, psi, mc)
    print("C3", t.E_local)
if "C4" in todo:
    spec, H = F.config_C4()
    psi, op = spec.build(True), H.build(True)
    mc = A.MonteCarloSpins(4096, 1, 2, 4096, True, seed=4)
    t = A.TDVP(psi.num_params, True)
    t.eval(op, psi, mc)
    print("C4", t.E_local)
