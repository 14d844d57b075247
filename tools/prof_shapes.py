"""Profiling driver (run under ncu): one short pass of the C3 / C4 / C5 shapes through the public API, sized so that an
`ncu --set full` capture of the samplers and the HBM-bound mat-vecs stays within a few minutes.
usage: python tools/prof_shapes.py [C3] [C4] [C5]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import annongpu_b200 as A
from annongpu_b200 import factories as F

A.setDevice(0)
todo = sys.argv[1:] or ["C3", "C4", "C5"]
if "C5" in todo:
    spec, H = F.config_C5()
    psi, op = spec.build(True), H.build(True)
    mc = A.MonteCarloSpins(2368, 1, 2, 2368, True, seed=5)        # 16 chains per SM, 2+1 sweeps
    t = A.TDVP(psi.num_params, True)
    t.eval_F(op, psi, mc)
    print("C5", t.E_local, t.solve_cg(tol=1e-30, max_iter=2, shift_abs=0.0, shift_rel=1e-3)[1:])
if "C4" in todo:
    spec, H = F.config_C4()
    psi, op = spec.build(True), H.build(True)
    mc = A.MonteCarloSpins(4736, 1, 2, 4736, True, seed=4)
    t = A.TDVP(psi.num_params, True)
    t.eval(op, psi, mc)
    os.environ["ANGPU_CG_MATRIX_FREE"] = "0"
    print("C4 S", t.E_local, t.solve_cg(tol=1e-30, max_iter=2, shift_abs=0.0, shift_rel=1e-3)[1:])
    os.environ["ANGPU_CG_MATRIX_FREE"] = "1"
    print("C4 O", t.solve_cg(tol=1e-30, max_iter=2, shift_abs=0.0, shift_rel=1e-3)[1:])
    os.environ["ANGPU_CG_MATRIX_FREE"] = "0"
if "C3" in todo:
    spec, H = F.config_C3()
    psi, op = spec.build(True), H.build(True)
    mc = A.MonteCarloSpins(2368, 1, 1, 2368, True, seed=3)
    t = A.TDVP(psi.num_params, True)
    t.eval_F(op, psi, mc)
    print("C3", t.E_local)
