"""On-hardware multi-rank correctness: the product path under a real process group (NCCL inside libangpu, and the callback
transport) must reproduce the one-rank results for the same global chains.  Needs >= 2 GPUs (skipped otherwise; run with
`gpurun --gpus 2`).  Single-GPU parts: a failing callback must fail the call, an unsharded ensemble must not be reduced."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from annongpu_b200 import factories as F

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("transport", ["nccl", "hook"])
def test_two_ranks_reproduce_one_rank(transport):
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    env = {**os.environ, "ANGPU_COMM": transport}
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29731" if transport == "nccl" else "29732", os.path.join(ROOT, "tests", "multirank_worker.py")],
                       env=env, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, (p.stdout[-2000:], p.stderr[-4000:])
    line = [l for l in p.stdout.splitlines() if l.startswith("MULTIRANK ")][-1]
    r = json.loads(line[len("MULTIRANK "):])
    for key in ("rbm_E", "rbm_F", "rbm_Ok", "rbm_Sv", "fluct", "es_E", "es_grad", "deep_S", "deep_F"):
        assert r[key] <= 1e-10, (key, r)
    assert r["rbm_cg_x"] <= 1e-7 and r["deep_solve"] <= 1e-7, r          # solves: conditioning times the reduction-order noise
    assert r["deep_S_tensorcore_vs_fp64"] <= 1e-4, r


def test_failing_allreduce_callback_fails_the_call(gpu):
    spec = F.rbm_spec(8, 8, noise=0.05, final_weight=1.0, seed=1)
    psi, op = spec.build(True), F.heisenberg(8, F.ring_bonds(8)).build(True)
    calls = []

    def bad(ptr, count):
        calls.append(count)
        raise RuntimeError("transport down")

    gpu.set_allreduce(bad)
    try:
        t = gpu.TDVP(psi.num_params, True)
        mc = gpu.MonteCarloSpins(64, 1, 2, 64, True, seed=3).set_shard(0, 2)
        with pytest.raises(gpu.AngpuError) as ei:
            t.eval_F(op, psi, mc)
        assert isinstance(ei.value.__cause__, RuntimeError) and calls
        # an ensemble that is NOT sharded is never summed over ranks, whatever transport is installed
        calls.clear()
        t.eval_F(op, psi, gpu.MonteCarloSpins(64, 1, 2, 64, True, seed=3))
        assert not calls
    finally:
        gpu.set_allreduce(None)


def test_apply_update_on_device_matches_host_update(gpu):
    spec, H = F.rbm_spec(10, 20, noise=0.05, final_weight=1.0, seed=2), F.heisenberg(10, F.ring_bonds(10))
    psi, op = spec.build(True), H.build(True)
    mc = gpu.MonteCarloSpins(1024, 1, 3, 1024, True, seed=8)
    t = gpu.TDVP(psi.num_params, True)
    t.eval_F(op, psi, mc)
    x, it, rr = t.solve_cg(tol=1e-9, max_iter=500)
    p0 = psi.params
    _, it2, _ = t.solve_cg(tol=1e-9, max_iter=500, keep_on_device=True)
    t.apply_update(psi, -0.01 + 0.002j)
    assert it == it2
    assert np.abs(psi.params - (p0 + (-0.01 + 0.002j) * x)).max() <= 1e-14
    # the device copies used by the sampler and by E_loc see the new weights: same results as a freshly built psi
    fresh = gpu.PsiRBM(psi.W, spec.final_weight, psi.log_prefactor)
    c0, l0 = gpu.MonteCarloSpins(256, 1, 2, 256, True, seed=5).sample(psi)
    c1, l1 = gpu.MonteCarloSpins(256, 1, 2, 256, True, seed=5).sample(fresh)
    assert np.array_equal(c0, c1) and np.abs(l0 - l1).max() <= 1e-13


@pytest.mark.parametrize("n", [1, 7, 100, 128, 129, 300, 1000])
def test_blocked_cholesky_solve_against_numpy(gpu, n):
    """The hand-written dense solve (csrc/cholesky.cu) on random Hermitian positive definite systems whose sizes straddle the
    128-row blocks and the 64-column tiles of the tensor-core trailing update: residual and solution against numpy."""
    rng = np.random.default_rng(n)
    B = rng.standard_normal((n, n + 5)) + 1j * rng.standard_normal((n, n + 5))
    A = B @ B.conj().T / n + 0.05 * np.eye(n)
    b = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    x = gpu.hpd_solve(A, b)
    x_ref = np.linalg.solve(A, b)
    assert np.linalg.norm(A @ x - b) <= 1e-11 * np.linalg.norm(b) * np.linalg.cond(A)
    assert np.linalg.norm(x - x_ref) <= 1e-10 * np.linalg.norm(x_ref) * max(1.0, np.linalg.cond(A) / 100)
    # only the upper triangle is read
    A2 = A.copy()
    A2[np.tril_indices(n, -1)] = 123.0
    assert np.array_equal(gpu.hpd_solve(A2, b), x)
    if n > 1:
        with pytest.raises(gpu.AngpuError):
            gpu.hpd_solve(A - 2.0 * np.trace(A).real / n * np.eye(n), b)
