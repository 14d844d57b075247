"""A/B of the PsiRBM E_loc kernels at C2: ANGPU_ELOC_TILE=0 (team-per-sample k_eloc_rbm) vs 1 (k_eloc_rbm_tile); prints the E_loc phase time."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    import numpy as np, annongpu_b200 as A
    from annongpu_b200 import factories as F
    A.setDevice(0)
    spec, H = F.config_C2(); psi, op = spec.build(True), H.build(True)
    mc = A.MonteCarloSpins(8192, 1, 10, 8192, True, seed=4242)
    t = A.TDVP(psi.num_params, True); t.set_profile(True)
    ts = []
    for it in range(12):
        t.eval_F(op, psi, mc); ts.append(dict(t.phase_ms))
    print(json.dumps({"tile": os.environ.get("ANGPU_ELOC_TILE"), "eloc_ms": float(np.median([x["eloc"] for x in ts[2:]])),
                      "total_ms": float(np.median([x["total"] for x in ts[2:]])), "E": [t.E_local.real, t.E_local.imag]}))
else:
    for tile in ("0", "1"):
        p = subprocess.run([sys.executable, __file__, "--child"], env={**os.environ, "ANGPU_ELOC_TILE": tile}, capture_output=True, text=True)
        print(p.stdout.strip().splitlines()[-1] if p.returncode == 0 else p.stderr[-1500:], flush=True)
