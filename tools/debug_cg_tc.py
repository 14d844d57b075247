"""Python-level PCG at C2 with the library's S.v: per-iteration error of the tensor-core product on the actual CG directions."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import annongpu_b200 as A
from annongpu_b200 import factories as F
A.setDevice(0)
spec, H = F.config_C2(); psi, op = spec.build(True), H.build(True)
mc = A.MonteCarloSpins(8192, 1, 10, 8192, True, seed=11)
t = A.TDVP(psi.num_params, True); t.eval_F(op, psi, mc)
P = psi.num_params
b = t.F_vector.copy()
# diagonal of S via the library's Jacobi data is not exposed: estimate with |O|^2 means from O_k samples is heavy; use shift = 1e-3*mean diag
diag = np.real(np.array([0.0]))
def Sv(v, tc):
    t.set_tensorcore_products(tc); out = t.S_dot_vector(v); t.set_tensorcore_products(False); return out
e = np.zeros(P, complex)
# cheap diagonal: S_kk for a few k, then use a constant shift (conditioning comparable)
ks = np.arange(0, P, 257); d = []
for k in ks:
    e[k] = 1; d.append(Sv(e, False)[k].real); e[k] = 0
shift = 1e-3 * float(np.mean(d))
for mode in ("tc_all", "tc_refresh8"):
    x = np.zeros(P, complex); r = b.copy(); p = r.copy(); rr = np.vdot(r, r).real; b2 = rr
    for it in range(1, 161):
        Ap_tc = Sv(p, True) + shift * p
        Ap_ex = Sv(p, False) + shift * p
        err = np.linalg.norm(Ap_tc - Ap_ex) / np.linalg.norm(Ap_ex)
        pAp_tc, pAp_ex = np.vdot(p, Ap_tc).real, np.vdot(p, Ap_ex).real
        alpha = rr / pAp_tc
        x += alpha * p; r -= alpha * Ap_tc
        if mode == "tc_refresh8" and it % 8 == 0: r = b - (Sv(x, False) + shift * x)
        rr_new = np.vdot(r, r).real; p = r + (rr_new / rr) * p; rr = rr_new
        true = np.linalg.norm(b - (Sv(x, False) + shift * x)) / np.sqrt(b2)
        if it % 8 == 0 or err > 1e-4: print(mode, it, "prod_err %.2e" % err, "pAp rel diff %.2e" % (abs(pAp_tc - pAp_ex) / abs(pAp_ex)), "rec %.2e" % np.sqrt(rr / b2), "true %.2e" % true, flush=True)
        if true < 1e-7 or not np.isfinite(true): break
