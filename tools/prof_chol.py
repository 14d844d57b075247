"""Profiling driver (run under ncu): one dense solve at the C4 size on a random HPD matrix (P = 8384)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import annongpu_b200 as A
A.setDevice(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8384
rng = np.random.default_rng(0)
B = rng.standard_normal((n, 64)) + 1j * rng.standard_normal((n, 64))
M = B @ B.conj().T / 64 + np.eye(n)
b = rng.standard_normal(n) + 0j
x = A.hpd_solve(M, b)
print(np.linalg.norm(M @ x - b) / np.linalg.norm(b))
