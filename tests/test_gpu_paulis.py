"""The CUDA path on the Pauli-string (density-matrix) basis -- MonteCarloPaulis / ExactSummationPaulis with a PsiDeep of
N = 3 num_sites input units -- against the compiled reference's golden vectors and the oracle port."""
import os

import numpy as np
import pytest

from helpers import make_op, make_psi, pauli_zoo
from oracle import port_oracle as P

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_golden_paulis.npz"))
TOL = dict(rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("name", list(pauli_zoo()))
def test_exact_summation_paulis_matches_reference_golden(gpu, name):
    spec, H, ns = pauli_zoo()[name]
    psi, op, es = make_psi(gpu, spec), make_op(gpu, H), gpu.ExactSummationPaulis(ns)
    assert es.num_steps == 4 ** ns and psi.pauli_sites == ns
    np.testing.assert_allclose(gpu.log_psi_vector(psi, es), G[f"{name}/log_psi_vector"], rtol=1e-12, atol=1e-13)
    ev = gpu.ExpectationValue(True)
    np.testing.assert_allclose(ev(op, psi, es), G[f"{name}/E"], **TOL)
    g, e = ev.gradient(op, psi, es)
    np.testing.assert_allclose(g, G[f"{name}/gradient"], **TOL)
    t = gpu.TDVP(psi.num_params, True)
    t.eval(op, psi, es)
    np.testing.assert_allclose(t.S_matrix, G[f"{name}/tdvp_S"], **TOL)
    np.testing.assert_allclose(t.F_vector, G[f"{name}/tdvp_F"], **TOL)
    np.testing.assert_allclose(t.O_k_vector, G[f"{name}/tdvp_O_k"], **TOL)
    np.testing.assert_allclose(t.E_local, G[f"{name}/tdvp_E"], **TOL)
    # configurations cross the boundary as units masks in PauliString::enumerate order
    confs, lp = es.sample(psi)
    assert np.array_equal(confs, P.enumerate_pauli_units(ns))
    for a, b, lps, ok in zip(G[f"{name}/probe_a"], G[f"{name}/probe_b"], G[f"{name}/probe_log_psi"], G[f"{name}/probe_O_k"]):
        u = gpu.paulis_to_units(a, b, ns)
        assert gpu.units_to_paulis(u, ns) == (int(a), int(b))
        np.testing.assert_allclose(gpu.log_psi_s(psi, u), lps, rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(gpu.psi_O_k(psi, u), ok, rtol=1e-12, atol=1e-13)


def test_pauli_local_energies_match_the_port_string_by_string(gpu):
    """Every branch of PauliString o PauliString: random operator strings on random Pauli configurations, 40 sites (two words of
    units), E_loc per configuration against the port."""
    ns = 40
    spec = __import__("annongpu_b200").factories.deep_spec(ns, 3 * ns, [60, 30], [12, 6], noise=0.05, final_weights=0.5, seed=5)
    rng = np.random.default_rng(3)
    from annongpu_b200 import factories as F
    H = 0.7 * F.sigma_x(0, ns)
    for k in range(24):
        term = complex(rng.normal(), rng.normal())
        for s in rng.choice(ns, size=rng.integers(1, 4), replace=False):
            term = term * (F.sigma_x, F.sigma_y, F.sigma_z)[rng.integers(3)](int(s), ns)
        H = H + term
    H = H + 0.3
    psi_g, psi_p, op_g, op_p = make_psi(gpu, spec), make_psi(P, spec), make_op(gpu, H), make_op(P, H, words=1)
    confs = []
    for _ in range(64):
        a, b = int(rng.integers(0, 1 << ns)), int(rng.integers(0, 1 << ns))
        confs.append(gpu.paulis_to_units(a, b, ns))
    confs = np.stack(confs)
    lp_g, el_g = gpu.local_energies(psi_g, op_g, confs)
    lp_p, el_p, _ = P.eval_samples(psi_p, op_p, confs)
    np.testing.assert_allclose(lp_g, lp_p, rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(el_g, el_p, rtol=1e-10, atol=1e-12)


def test_monte_carlo_paulis_chains_identical_to_the_port(gpu):
    """Same Philox streams, same Init / Update policies: the recorded Pauli strings coincide chain by chain, across two calls."""
    spec, H, ns = pauli_zoo()["pdeep4"]
    psi_g, psi_p = make_psi(gpu, spec), make_psi(P, spec)
    mg, mp = gpu.MonteCarloPaulis(512, 2, 3, 128, True, seed=77), P.MonteCarloPaulis(512, 2, 3, 128, seed=77)
    for call in range(2):
        cg, lg = mg.sample(psi_g)
        cp, lpp = mp.sample(psi_p)
        assert np.array_equal(cg, cp)
        np.testing.assert_allclose(lg, lpp, rtol=1e-11, atol=1e-12)
        assert mg.acceptances == (mp.acceptances, mp.rejections)


def test_monte_carlo_paulis_estimates_within_3_sigma(gpu):
    spec, H, ns = pauli_zoo()["pdeep3"]
    psi, op = make_psi(gpu, spec), make_op(gpu, H)
    ev = gpu.ExpectationValue(True)
    lp = gpu.log_psi_vector(psi, gpu.ExactSummationPaulis(ns))
    exact = ev(op, psi, gpu.ExactSummationPaulis(ns)) / np.sum(np.exp(2 * lp.real))
    chains = 2048
    mc = gpu.MonteCarloPaulis(chains * 8, 2, 20, chains, True, seed=5)
    t = gpu.TDVP(psi.num_params, True)
    t.eval_F(op, psi, mc)
    sigma = np.sqrt(max(t.var_H, 0.0) / chains)                 # chains are independent; samples within a chain are not
    assert abs(t.E_local.real - exact.real) < 3 * sigma + 1e-9
    assert 0.3 < mc.acceptance_rate <= 1.0


def test_basis_mismatch_is_an_error(gpu):
    spec, H, ns = pauli_zoo()["pdeep3"]
    psi, op = make_psi(gpu, spec), make_op(gpu, H)
    with pytest.raises(RuntimeError, match="Pauli"):
        gpu.ExpectationValue(True)(op, psi, gpu.MonteCarloSpins(64, 1, 1, 64, True))
    with pytest.raises(RuntimeError, match="Pauli"):
        gpu.ExpectationValue(True)(op, psi, gpu.ExactSummationSpins(3 * ns))
    from annongpu_b200 import factories as F
    spin_psi = make_psi(gpu, F.deep_spec(6, 6, [6], [3], seed=1))
    with pytest.raises(RuntimeError, match="Pauli"):
        gpu.ExpectationValue(True)(make_op(gpu, F.heisenberg(6, F.ring_bonds(6))), spin_psi, gpu.ExactSummationPaulis(6))
