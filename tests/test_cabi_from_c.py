"""The drop-in boundary exercised from plain C: tests/cabi/vmc_from_c.c is compiled with gcc against include/angpu.h and
linked to annongpu_b200/libangpu.so (no Python, no torch in that process), run on cuda:0, and its Monte-Carlo energy is
compared with the same computation through the Python mirror."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build(tmp_path):
    exe = str(tmp_path / "vmc_from_c")
    libdir = os.path.join(ROOT, "annongpu_b200")
    cmd = ["gcc", "-std=c99", "-Wall", "-O1", os.path.join(ROOT, "tests", "cabi", "vmc_from_c.c"), "-I", os.path.join(ROOT, "include"),
           "-L", libdir, "-langpu", "-lm", "-Wl,-rpath," + libdir, "-o", exe]
    subprocess.run(cmd, check=True)
    return exe


def test_c_program_compiles_and_links(tmp_path):
    """CPU: the header is valid C99 and every symbol the program uses resolves against the shared library."""
    assert os.path.exists(build(tmp_path))


@pytest.mark.gpu
def test_c_program_matches_python_mirror(tmp_path, gpu):
    exe = build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    tok = r.stdout.split()
    E_c, acc_c, it_c = complex(float(tok[1]), float(tok[2])), float(tok[4]), int(tok[6])
    # the same run through the Python mirror (same seed -> same Philox chains -> same sums up to reduction order)
    N, M, chains = 12, 24, 2048
    k = np.arange(N * M)
    W = (0.05 * (np.sin(0.37 * k + 0.1) + 1j * np.cos(0.11 * k))).reshape(N, M)
    sys.path.insert(0, ROOT)
    from annongpu_b200 import factories as F
    psi = gpu.PsiRBM(W, 2.0, 0.0, True)
    H = gpu.Operator(F.heisenberg(N, F.ring_bonds(N)), True)
    mc = gpu.MonteCarloSpins(chains, 1, 10, chains, True, seed=42)
    t = gpu.TDVP(psi.num_params, True)
    t.eval_F(H, psi, mc)
    assert abs(t.E_local - E_c) <= 1e-9 * max(1.0, abs(E_c))
    assert abs(mc.acceptance_rate - acc_c) <= 1e-4 and it_c > 0
