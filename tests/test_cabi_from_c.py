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


def build_two_ranks(tmp_path):
    exe = str(tmp_path / "vmc_two_ranks")
    libdir = os.path.join(ROOT, "annongpu_b200")
    cmd = ["gcc", "-std=gnu99", "-Wall", "-O1", os.path.join(ROOT, "tests", "cabi", "vmc_two_ranks.c"), "-I", os.path.join(ROOT, "include"),
           "-L", libdir, "-langpu", "-lm", "-Wl,-rpath," + libdir, "-o", exe]
    subprocess.run(cmd, check=True)
    return exe


def test_two_rank_c_program_compiles_and_links(tmp_path):
    assert os.path.exists(build_two_ranks(tmp_path))


@pytest.mark.gpu
def test_two_ranks_from_plain_c(tmp_path):
    """libangpu's own NCCL communicator driven from C: two forked ranks reproduce the one-rank E and F to 1e-10."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([build_two_ranks(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert any(line.startswith("E2 ") for line in r.stdout.splitlines()), r.stdout      # NCCL may print its version banner first


def build_cxx(tmp_path):
    exe = str(tmp_path / "vmc_from_cxx")
    libdir = os.path.join(ROOT, "annongpu_b200")
    cmd = ["g++", "-std=c++14", "-Wall", "-O1", os.path.join(ROOT, "tests", "cabi", "vmc_from_cxx.cpp"), "-I", os.path.join(ROOT, "include"),
           "-L", libdir, "-langpu", "-Wl,-rpath," + libdir, "-o", exe]
    subprocess.run(cmd, check=True)
    return exe


def test_cxx_program_compiles_and_links(tmp_path):
    """CPU: include/angpu.hpp (the RAII classes with the reference's names) is valid C++14 over the C ABI."""
    assert os.path.exists(build_cxx(tmp_path))


@pytest.mark.gpu
def test_cxx_program_runs_and_matches_the_c_program(tmp_path, gpu):
    rc = subprocess.run([build(tmp_path)], capture_output=True, text=True, timeout=300)
    rx = subprocess.run([build_cxx(tmp_path)], capture_output=True, text=True, timeout=300)
    assert rc.returncode == 0 and rx.returncode == 0, rc.stderr + rx.stderr
    tc, tx = rc.stdout.split(), rx.stdout.split()
    assert abs(float(tc[1]) - float(tx[1])) <= 1e-12 and int(tc[6]) == int(tx[4])      # same E, same CG iteration count


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
