"""Dense-solve driver (also run under ncu): angpu_hpd_solve on random HPD matrices; prints the relative residuals.
    python tools/prof_chol.py [n ...]      default: edge sizes + the C4 size (P = 8384)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import annongpu_b200 as A
A.setDevice(0)
sizes = [int(a) for a in sys.argv[1:]] or [1, 5, 127, 128, 129, 300, 1000, 8384]
rng = np.random.default_rng(0)
for n in sizes:
    k = max(8, n // 16)
    B = rng.standard_normal((n, k)) + 1j * rng.standard_normal((n, k))
    M = B @ B.conj().T / k + np.eye(n)
    b = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    t0 = time.time()
    x = A.hpd_solve(M, b)
    print(n, "rel residual", np.linalg.norm(M @ x - b) / np.linalg.norm(b), "wall s (incl. upload)", round(time.time() - t0, 3), flush=True)
