"""Factorised S.v of a PsiRBM on the tcgen05 tensor cores (csrc/sv_tc.cu, opt-in: TDVP.set_tensorcore_products): one product against the
exact FP64-tensor-core product, and the CG solve -- tensor-core search directions, exact residual refresh -- against the exact solve.
Tolerances: the product is TF32 hi+lo planes with fp32 accumulation (1e-5 relative to ||S v||); the solve keeps its fp64 tolerance."""
import numpy as np
import pytest

from annongpu_b200 import factories as F
from helpers import make_op, make_psi

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,M,ns", [(24, 72, 1000), (64, 256, 4096), (100, 200, 777), (200, 416, 1536)])
def test_tensorcore_product_matches_exact(gpu, N, M, ns):
    """Shapes cover: K padding of the site dimension (24, 100, 200 are not multiples of 16), more than one row tile of sites (200),
    column tiles with a ragged edge (72, 200, 416), a sample count that is not a multiple of 128 or 16 (1000, 777)."""
    spec = F.rbm_spec(N, M, noise=0.05, final_weight=1.0, seed=3)
    psi, op = make_psi(gpu, spec), make_op(gpu, F.heisenberg(N, F.ring_bonds(N)))
    mc = gpu.MonteCarloSpins(ns, 1, 3, ns, True, seed=11)
    t = gpu.TDVP(psi.num_params, True)
    t.eval_F(op, psi, mc)
    rng = np.random.default_rng(0)
    for scale in (1.0, 1e-6):
        v = scale * (rng.standard_normal(psi.num_params) + 1j * rng.standard_normal(psi.num_params))
        exact = t.S_dot_vector(v)
        tc = t.set_tensorcore_products(True).S_dot_vector(v)
        t.set_tensorcore_products(False)
        assert np.linalg.norm(tc - exact) <= 1e-5 * np.linalg.norm(exact), (np.linalg.norm(tc - exact) / np.linalg.norm(exact))
        assert np.array_equal(t.S_dot_vector(v), exact)                  # switching back restores the exact, deterministic product


def test_cg_with_tensorcore_products_meets_the_fp64_tolerance(gpu):
    spec = F.rbm_spec(32, 128, noise=0.02, final_weight=1.0, seed=5)
    psi, op = make_psi(gpu, spec), make_op(gpu, F.heisenberg(32, F.ring_bonds(32)))
    mc = gpu.MonteCarloSpins(4096, 1, 10, 4096, True, seed=2)
    t = gpu.TDVP(psi.num_params, True)
    t.eval_F(op, psi, mc)
    x_ex, it_ex, rr_ex = t.solve_cg(tol=1e-6, max_iter=2000, shift_rel=1e-3)
    x_tc, it_tc, rr_tc = t.set_tensorcore_products(True).solve_cg(tol=1e-6, max_iter=2000, shift_rel=1e-3)
    t.set_tensorcore_products(False)
    assert rr_ex <= 1e-6 and rr_tc <= 1e-6 and it_tc <= 2 * it_ex + 16
    # the residual of the tensor-core run is recomputed with exact products at every check: both solutions satisfy the same fp64
    # criterion, so they agree to (tolerance x conditioning), and S x - F (the unshifted part of the residual) has the same size
    assert np.linalg.norm(x_tc - x_ex) <= 1e-3 * np.linalg.norm(x_ex)
    res, res_ex = t.S_dot_vector(x_tc) - t.F_vector, t.S_dot_vector(x_ex) - t.F_vector
    assert abs(np.linalg.norm(res) - np.linalg.norm(res_ex)) <= 1e-3 * np.linalg.norm(t.F_vector)


def test_auto_mode_keeps_small_solves_exact(gpu):
    """Default (auto): S_dot_vector is exact and solve_cg uses the tensor cores only for ns N M >= 1e9 -- a small solve is bit-identical
    to one with the tensor-core products switched off."""
    spec = F.rbm_spec(16, 32, noise=0.05, final_weight=1.0, seed=9)
    psi, op = make_psi(gpu, spec), make_op(gpu, F.heisenberg(16, F.ring_bonds(16)))
    t = gpu.TDVP(psi.num_params, True)
    t.eval_F(op, psi, gpu.MonteCarloSpins(2048, 1, 5, 2048, True, seed=4))
    x_auto, it_auto, _ = t.solve_cg(tol=1e-8, max_iter=500, shift_rel=1e-3)
    x_off, it_off, _ = t.set_tensorcore_products(False).solve_cg(tol=1e-8, max_iter=500, shift_rel=1e-3)
    assert it_auto == it_off and np.array_equal(x_auto, x_off)
