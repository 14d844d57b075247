"""Pins the CPU oracle (oracle/port) to the reference: against the golden vectors generated from the compiled,
unmodified reference (tests/golden/ref_golden.npz, made by tests/golden/make_golden.py), and — where
oracle/_ref/liboracle_ref.so exists — against the live reference on further seeded inputs."""
import os

import numpy as np
import pytest

from annongpu_b200 import factories as F
from helpers import classical_zoo, hsd_cases, kl_cases, kl_sequence, make_classical, make_op, make_psi, rel_err, zoo

GOLDEN = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_golden.npz"))
PROBES = [0x2A5, 0x13, 0x3FF, 0x0]
TOL = 1e-10


def check_against_golden(mod, name, psi, H, N, ensemble_cls):
    g = lambda k: GOLDEN[f"{name}/{k}"]   # noqa: E731
    op, es = make_op(mod, H), ensemble_cls(N)
    assert abs(mod.psi_norm(psi, es) - g("norm")) <= TOL * g("norm")
    psi.log_prefactor = complex(g("log_prefactor"))
    assert abs(mod.expectation(op, psi, es) - g("E")) <= TOL * max(1.0, abs(g("E")))
    assert abs(mod.fluctuation(op, psi, es)[0] - g("fluctuation")) <= 1e-9 * max(1.0, g("fluctuation"))
    assert rel_err(mod.gradient(op, psi, es)[0], g("gradient")) <= TOL
    t = mod.TDVP(psi.num_params)
    t.eval(op, psi, es)
    assert rel_err(t.F_vector, g("F")) <= TOL and rel_err(t.O_k_vector, g("Ok")) <= TOL
    assert abs(t.var_H - g("var_H")) <= 1e-9 * max(1.0, abs(g("var_H")))
    assert rel_err(np.diag(t.S_matrix), g("S_diag")) <= TOL
    assert rel_err(t.S_matrix @ g("v"), g("Sv_dense")) <= 1e-9
    assert rel_err(t.S_dot_vector(g("v"), es), g("Sv")) <= 1e-9
    mask = (1 << N) - 1
    for i, c in enumerate(PROBES):
        assert abs(mod.log_psi_s(psi, [c & mask]) - g("log_psi_s")[i]) <= TOL * max(1.0, abs(g("log_psi_s")[i]))
        assert rel_err(mod.psi_O_k(psi, [c & mask]), g("O_k")[i]) <= TOL
    assert rel_err(mod.psi_vector(psi, es)[:64], g("psi_vector_head")) <= TOL
    assert rel_err(mod.apply_operator(psi, op, es)[:64], g("apply_operator_head")) <= TOL
    assert abs(mod.log_psi(psi, es) - g("log_psi_mean")) <= 1e-9
    if hasattr(mod, "exp_sigma_z"):
        assert abs(mod.exp_sigma_z(op, psi, es) - g("exp_sigma_z")) <= TOL * abs(g("exp_sigma_z"))


def check_wref_against_golden(mod, name, psi, H, N, ensemble_cls):
    """TDVP::eval(..., true_t) = eval_with_psi_ref of a PsiClassical, after check_against_golden fixed its log_prefactor."""
    g = lambda k: GOLDEN[f"{name}/wref/{k}"]   # noqa: E731
    t = mod.TDVP(psi.num_params)
    t.eval_with_psi_ref(make_op(mod, H), psi, ensemble_cls(N))
    assert abs(t.total_weight - g("total_weight")) <= TOL * g("total_weight")
    assert abs(t.E_local - g("E")) <= TOL * max(1.0, abs(g("E")))
    assert rel_err(t.F_vector, g("F")) <= TOL and rel_err(t.O_k_vector, g("Ok")) <= TOL
    assert rel_err(t.S_matrix, g("S")) <= TOL


@pytest.mark.parametrize("name", sorted(zoo()))
def test_port_matches_reference_golden(port, name):
    spec, H, N = zoo()[name]
    check_against_golden(port, name, make_psi(port, spec), H, N, port.ExactSummation)


@pytest.mark.parametrize("name", sorted(classical_zoo()))
def test_port_matches_reference_golden_classical(port, name):
    N, order, Hl, pr, ref_spec, lp, H = classical_zoo()[name]
    psi = make_classical(port, N, order, Hl, pr, ref_spec, lp)
    check_against_golden(port, name, psi, H, N, port.ExactSummation)
    check_wref_against_golden(port, name, psi, H, N, port.ExactSummation)


def check_hsd_against_golden(mod, name, ensemble_cls):
    spec, spec_prime, OP, is_unitary, N = hsd_cases()[name]
    psi, psi_prime, op, es = make_psi(mod, spec), make_psi(mod, spec_prime), make_op(mod, OP), ensemble_cls(N)
    d = mod.hilbert_space_distance(psi, psi_prime, op, is_unitary, es)
    assert abs(d - GOLDEN[f"hsd/{name}/distance"]) <= 1e-9
    g, d2 = mod.hilbert_space_distance_gradient(psi, psi_prime, op, is_unitary, es, 1.0)
    assert abs(d2 - GOLDEN[f"hsd/{name}/distance_g"]) <= 1e-9
    assert rel_err(g, GOLDEN[f"hsd/{name}/gradient"]) <= 1e-8


@pytest.mark.parametrize("name", sorted(hsd_cases()))
def test_port_hilbert_space_distance_matches_reference_golden(port, name):
    check_hsd_against_golden(port, name, port.ExactSummation)


def check_kl_against_golden(mod, name, ensemble_cls):
    got = kl_sequence(mod, name, ensemble_cls)
    g = lambda k: GOLDEN[f"kl/{name}/{k}"]   # noqa: E731
    for k in ("v1", "v2", "v3"):
        assert abs(got[k] - g(k)) <= 1e-9 * max(1.0, abs(g(k))), k
    assert rel_err(got["g"], g("g")) <= 1e-8 and rel_err(got["gn"], g("gn")) <= 1e-8
    assert abs(got["total_weight"] - g("total_weight")) <= TOL * g("total_weight")
    assert abs(got["mean_deviation"] - g("mean_deviation")) <= 1e-9 * max(1.0, abs(g("mean_deviation")))
    if name.endswith("_cnn"):
        # PsiCNN::foreach_O_k emits a parameter several times (one partial per lattice site, PsiCNN.hpp:185-266);
        # upstream squares each partial in the noise sums, which is not a function of O_k: the noise is not compared
        return
    ref_noise = g("noise")
    ok = np.isfinite(ref_noise)                       # the upstream variance estimate can be negative (sqrt -> nan)
    assert np.array_equal(np.isfinite(got["noise"]), ok)
    assert np.allclose(got["noise"][ok], ref_noise[ok], rtol=1e-6, atol=1e-12)


@pytest.mark.parametrize("name", sorted(kl_cases()))
def test_port_kullback_leibler_matches_reference_golden(port, name):
    check_kl_against_golden(port, name, port.ExactSummation)


def test_port_primitives_match_golden(port):
    a, b, c = GOLDEN["pauli/a"], GOLDEN["pauli/b"], GOLDEN["pauli/conf"]
    for i in range(len(a)):
        coeff, out = port.pauli_apply([a[i]], [b[i]], [c[i]], 1)
        assert coeff == complex(GOLDEN["pauli/coeff"][i])          # bit-exact: phases are +-1, +-i
        assert int(out[0]) == int(GOLDEN["pauli/conf_out"][i])
    for layer in (0, 1, 2):
        for z, lc, th in zip(GOLDEN["act/z"], GOLDEN[f"act/lc{layer}"], GOLDEN[f"act/th{layer}"]):
            plc, pth = port.activation(z, layer)
            assert abs(plc - lc) <= 1e-15 * max(1, abs(lc)) and abs(pth - th) <= 1e-15 * max(1, abs(th))


def test_philox_known_answers(port):
    """Random123 known-answer vectors for Philox4x32-10."""
    kat = [([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
           ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
           ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0], [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    for ctr, key, out in kat:
        assert [int(x) for x in port.philox(ctr, key)] == out


def test_port_matches_live_reference_on_fresh_inputs(port, ref):
    """Beyond the committed fixtures: new seeds, the reference's largest RBM (N=16, M=128) on sampled configurations."""
    rng = np.random.default_rng(2024)
    spec = F.rbm_spec(16, 128, noise=3e-2, final_weight=2, seed=77)
    H = F.heisenberg(16, F.ring_bonds(16))
    pr, pp = make_psi(ref, spec), make_psi(port, spec)
    opr, opp = make_op(ref, H), make_op(port, H)
    for c in rng.integers(0, 1 << 16, size=16):
        c = int(c)
        assert abs(ref.log_psi_s(pr, c) - port.log_psi_s(pp, [c])) <= 1e-12
        assert rel_err(port.psi_O_k(pp, [c]), ref.psi_O_k(pr, c)) <= 1e-12
    esr, esp = ref.ExactSummation(12), port.ExactSummation(12)
    spec2 = F.deep_spec(12, 12, [24, 12], [6, 12], noise=2e-2, a=0.05, final_weights=2, seed=78)
    H2 = F.tfim(12, F.ring_bonds(12), h=1.3)
    p2r, p2p = make_psi(ref, spec2), make_psi(port, spec2)
    o2r, o2p = make_op(ref, H2), make_op(port, H2)
    gr, er = ref.gradient(o2r, p2r, esr)
    gp, ep = port.gradient(o2p, p2p, esp)
    assert rel_err(gp, gr) <= TOL and abs(ep - er) <= TOL * abs(er)
    # the reference's CPU Monte-Carlo (chain 0, mt19937) and the port's (Philox) agree statistically
    n = ref.psi_norm(p2r, esr)
    p2r.log_prefactor = -np.log(n)
    p2p.log_prefactor = -np.log(n)
    E_exact = ref.expectation(o2r, p2r, esr)
    mcr = ref.MonteCarlo(4000, 2, 50, 1)
    f_r, E_r = ref.fluctuation(o2r, p2r, mcr)
    mcp = port.MonteCarlo(4000, 2, 50, 8, seed=3)
    f_p, E_p = port.fluctuation(o2p, p2p, mcp)
    assert abs(E_r - E_exact) <= 5 * f_r / np.sqrt(4000 / 8) and abs(E_p - E_exact) <= 5 * f_p / np.sqrt(4000 / 8)
