"""Parity of the CUDA path (through the C ABI) against the CPU oracle, on the same seeded inputs.

Tolerances (BASELINE.json north_star): fp64 quantities within 1e-10 relative; Monte-Carlo estimates within 3 sigma;
basis enumeration and Pauli-string action bit-exact.
"""
import numpy as np
import pytest

from annongpu_b200 import factories as F
from helpers import classical_zoo, make_classical, make_op, make_psi, rel_err, wide_zoo, zoo

pytestmark = pytest.mark.gpu

TOL = 1e-10   # north_star: ExactSummation energies / gradients / log psi / E_loc within 1e-10 relative (fp64)


# ------------------------------------------------------------------------------------------------ bit-exact integer work

def test_spins_enumerate_bit_exact(gpu):
    for idx in [0, 1, 0x2A5, 65535, (1 << 31) - 1]:
        s = gpu.Spins.enumerate(idx, 64)
        assert s.configuration == idx
    s = gpu.Spins.enumerate(12345, 200)
    assert s.configuration == 12345 and len(s.words()) == 4


@pytest.mark.parametrize("num_sites", [10, 64, 100, 200])
def test_pauli_apply_bit_exact(gpu, port, num_sites):
    rng = np.random.default_rng(num_sites)
    words = F.words_for(num_sites)
    top = (1 << num_sites) - 1
    for _ in range(40):
        a = int.from_bytes(rng.bytes(32), "little") & top
        b = int.from_bytes(rng.bytes(32), "little") & top
        c = int.from_bytes(rng.bytes(32), "little") & top
        cg, sg = gpu.pauli_apply(a, b, c, num_sites)
        cp_, sp_ = port.pauli_apply(F.masks_to_words([a], words)[0], F.masks_to_words([b], words)[0],
                                    F.masks_to_words([c], words)[0], words)
        assert cg == cp_                                   # phases are exactly +-1, +-i
        assert sg.configuration == port.conf_int(sp_)
        assert sg.configuration == c ^ (a ^ b)


def test_activation_polynomials(gpu, port):
    rng = np.random.default_rng(0)
    for layer in (0, 1, 2):
        for _ in range(10):
            z = complex(rng.normal(), rng.normal()) * 0.7
            lc, th = port.activation(z, layer)
            assert abs(gpu.activation_function(z, layer) - lc) <= 1e-14 * max(1.0, abs(lc))
            assert abs(gpu.activation_derivative(z, layer) - th) <= 1e-14 * max(1.0, abs(th))


# ------------------------------------------------------------------------------------------------ ExactSummation parity

def _es_compare(gpu, port, pg, pp, H, N):
    og, op_ = make_op(gpu, H), make_op(port, H)
    eg, ep = gpu.ExactSummationSpins(N), port.ExactSummation(N)
    n_p = port.psi_norm(pp, ep)
    assert abs(pg.norm(eg) - n_p) <= TOL * n_p
    pg.log_prefactor = pg.log_prefactor - np.log(n_p)
    pp.log_prefactor = pp.log_prefactor - np.log(n_p)
    assert pg.num_params == pp.num_params
    assert rel_err(pg.params, pp.params) == 0.0

    rng = np.random.default_rng(N)
    for c in rng.integers(0, 1 << N, size=6):
        c = int(c)
        lp_p = port.log_psi_s(pp, [c])
        assert abs(gpu.log_psi_s(pg, gpu.Spins(c, N)) - lp_p) <= TOL * max(1.0, abs(lp_p))
        assert rel_err(gpu.psi_O_k(pg, c), port.psi_O_k(pp, [c])) <= TOL
        assert abs(gpu.local_energies(pg, og, [[c]])[1][0] - port.local_energy(pp, op_, [c])) <= TOL * 10

    assert rel_err(gpu.psi_vector(pg, eg), port.psi_vector(pp, ep)) <= TOL
    assert rel_err(gpu.log_psi_vector(pg, eg), port.log_psi_vector(pp, ep)) <= TOL
    assert abs(gpu.log_psi(pg, eg) - port.log_psi(pp, ep)) <= TOL * 10
    assert rel_err(gpu.apply_operator(pg, og, eg), port.apply_operator(pp, op_, ep)) <= TOL
    assert rel_err(gpu.psi_O_k_vector(pg, eg), port.psi_O_k_vector(pp, ep)) <= 1e-9

    ev = gpu.ExpectationValue(True)
    E_p = port.expectation(op_, pp, ep)
    assert abs(ev(og, pg, eg) - E_p) <= TOL * max(1.0, abs(E_p))
    f_g, m_g = ev.fluctuation(og, pg, eg)
    f_p, m_p = port.fluctuation(op_, pp, ep)
    assert abs(f_g - f_p) <= 1e-9 * max(1.0, f_p) and abs(m_g - m_p) <= TOL * max(1.0, abs(m_p))
    g_g, e_g = ev.gradient(og, pg, eg)
    g_p, e_p = port.gradient(op_, pp, ep)
    assert rel_err(g_g, g_p) <= TOL and abs(e_g - e_p) <= TOL * max(1.0, abs(e_p))

    tg, tp = gpu.TDVP(pg.num_params, True), port.TDVP(pp.num_params)
    tg.eval(og, pg, eg)
    tp.eval(op_, pp, ep)
    assert rel_err(tg.S_matrix, tp.S_matrix) <= TOL
    assert rel_err(tg.F_vector, tp.F_vector) <= TOL
    assert rel_err(tg.O_k_vector, tp.O_k_vector) <= TOL
    assert abs(tg.var_H - tp.var_H) <= 1e-9 * max(1.0, abs(tp.var_H))
    assert abs(tg.E_local - tp.E_local) <= TOL * max(1.0, abs(tp.E_local))
    assert rel_err(tg.O_k_samples.reshape(1 << N, -1), tp.O_k_samples) <= TOL
    assert rel_err(tg.weight_samples, tp.weight_samples) <= TOL
    v = rng.normal(size=pg.num_params) + 1j * rng.normal(size=pg.num_params)
    sv_p = tp.S_dot_vector(v)
    assert rel_err(tg.S_dot_vector(v, eg), sv_p) <= 1e-9
    # eval_F keeps PsiRBM samples factorised: the matrix-free product must agree with the dense one
    tg.eval_F(og, pg, eg)
    assert rel_err(tg.F_vector, tp.F_vector) <= TOL
    assert rel_err(tg.S_dot_vector(v, eg), sv_p) <= 1e-9
    # the (new) solvers: residual of the shifted system against the oracle's S and F
    shift = 1e-3
    A = tp.S_matrix + shift * np.eye(pg.num_params)
    x_cg, it, rr = tg.solve_cg(tol=1e-10, max_iter=4000, shift_abs=shift, shift_rel=0.0)
    assert np.linalg.norm(A @ x_cg - tp.F_vector) <= 1e-7 * np.linalg.norm(tp.F_vector)
    tg.eval(og, pg, eg)
    x_d = tg.solve(shift_abs=shift, shift_rel=0.0)
    assert np.linalg.norm(A @ x_d - tp.F_vector) <= 1e-8 * np.linalg.norm(tp.F_vector)


@pytest.mark.parametrize("name", sorted(zoo()))
def test_exact_summation_parity(gpu, port, name):
    spec, H, N = zoo()[name]
    _es_compare(gpu, port, make_psi(gpu, spec), make_psi(port, spec), H, N)


@pytest.mark.parametrize("name", sorted(classical_zoo()))
def test_exact_summation_parity_classical(gpu, port, name):
    N, order, Hl, pr, ref_spec, lp, H = classical_zoo()[name]
    pg = make_classical(gpu, N, order, Hl, pr, ref_spec, lp)
    pp = make_classical(port, N, order, Hl, pr, ref_spec, lp)
    _es_compare(gpu, port, pg, pp, H, N)


def test_params_roundtrip_and_copy(gpu, port):
    spec, H, N = zoo()["deep2"]
    pg, pp = make_psi(gpu, spec), make_psi(port, spec)
    rng = np.random.default_rng(3)
    new = pg.params + 1e-2 * (rng.normal(size=pg.num_params) + 1j * rng.normal(size=pg.num_params))
    pg2 = +pg
    pg2.params = new
    pp.params = new
    assert rel_err(pg2.params, new) == 0.0
    c = 0x5A
    assert abs(gpu.log_psi_s(pg2, c) - port.log_psi_s(pp, [c])) <= TOL * 10
    assert abs(gpu.log_psi_s(pg, c) - gpu.log_psi_s(pg2, c)) > 1e-6      # the copy is independent
    # O_k against central finite differences of log_psi_s (the reference's own test, test/test_Psi.py:259-324)
    eps, O = 1e-5, gpu.psi_O_k(pg2, c)
    for k in rng.integers(N, pg.num_params, size=8):     # k < N: input_weights, not used by log psi
        p1, p2 = new.copy(), new.copy()
        p1[k] += eps
        p2[k] -= eps
        pg.params = p1
        f1 = gpu.log_psi_s(pg, c)
        pg.params = p2
        f2 = gpu.log_psi_s(pg, c)
        assert abs((f1 - f2) / (2 * eps) - O[k]) <= 1e-6 * max(1.0, abs(O[k]))


def test_expectation_list_and_hermitian_identity(gpu, port):
    """<psi|H|psi> = v^dagger H v with the dense matrix built from PauliString.apply (test/test_ExpectationValue.py:6-30)."""
    spec, H, N = zoo()["rbm10"]
    pg, pp = make_psi(gpu, spec), make_psi(port, spec)
    eg = gpu.ExactSummationSpins(N)
    pg.normalize(eg)
    v = pg.vector(eg)
    Hm = make_op(port, H).dense_matrix(N)
    ev = gpu.ExpectationValue(True)
    og = make_op(gpu, H)
    assert abs(ev(og, pg, eg) - np.vdot(v, Hm @ v)) <= 1e-9
    sx = F.PauliSum(N).add(1.0, {N // 2: "X"})
    sy = F.PauliSum(N).add(1.0, {N // 2: "Y"})
    sz = F.PauliSum(N).add(1.0, {N // 2: "Z"})
    many = ev([make_op(gpu, o) for o in (sx, sy, sz, H)], pg, eg)
    for val, o in zip(many, (sx, sy, sz, H)):
        assert abs(val - np.vdot(v, make_op(port, o).dense_matrix(N) @ v)) <= 1e-9
    # H @ psi (test/test_operator.py:6-26)
    assert rel_err(gpu.apply_operator(pg, og, eg), Hm @ v) <= 1e-9


# ------------------------------------------------------------------------------------------------ Monte Carlo

@pytest.mark.parametrize("name", ["rbm10", "rbm8_cfw", "rbm12_m40", "rbm_wide", "deep1", "deep2", "deep3", "cnn"])
def test_mc_chains_identical_to_oracle(gpu, port, name):
    """Same Philox stream, same proposals: the sampled configurations must coincide chain by chain."""
    spec, H, N = {**zoo(), **wide_zoo()}[name]
    pg, pp = make_psi(gpu, spec), make_psi(port, spec)
    chains, per_chain = 24, 3
    mg = gpu.MonteCarloSpins(chains * per_chain, 2, 3, chains, True, seed=1234)
    mp = port.MonteCarlo(chains * per_chain, 2, 3, chains, seed=1234)
    for call in range(2):                           # the stream continues across calls
        cg, lg = mg.sample(pg)
        cp_, lp_ = mp.sample(pp)
        assert np.array_equal(cg, cp_)
        assert rel_err(lg, lp_) <= 1e-9
        a, r = mg.acceptances
        assert (a, r) == (mp.acceptances, mp.rejections)
        assert a + r == chains * (3 + 2 * per_chain) * N
    # sharded run: two "ranks" reproduce the halves of the single-process run
    mfull = gpu.MonteCarloSpins(chains * per_chain, 2, 3, chains, True, seed=99)
    full, _ = mfull.sample(pg)
    parts = []
    for rank in range(2):
        m = gpu.MonteCarloSpins(chains * per_chain, 2, 3, chains, True, seed=99).set_shard(rank, 2)
        parts.append(m.sample(pg)[0].reshape(per_chain, chains // 2, -1))
    assert np.array_equal(np.concatenate(parts, axis=1).reshape(full.shape), full)


def test_mc_energy_within_3_sigma_of_exact(gpu, port):
    spec, H, N = zoo()["rbm10"]
    pg = make_psi(gpu, spec)
    og = make_op(gpu, H)
    eg = gpu.ExactSummationSpins(N)
    pg.normalize(eg)
    ev = gpu.ExpectationValue(True)
    E_exact = ev(og, pg, eg)
    chains, per_chain = 2048, 8
    # |psi|^2 of this model is sharply peaked (~5 effective states): single-spin-flip chains need ~200 sweeps to equilibrate
    mc = gpu.MonteCarloSpins(chains * per_chain, 2, 200, chains, True, seed=7)
    fluct, E_mc = ev.fluctuation(og, pg, mc)
    sigma = fluct / np.sqrt(chains)          # conservative: chains are independent, samples within a chain are not
    assert abs(E_mc - E_exact) <= 3 * sigma + 1e-12
    assert 0.05 < mc.acceptance_rate < 1.0
    g_mc, _ = ev.gradient(og, pg, mc)
    g_ex, _ = ev.gradient(og, pg, eg)
    assert np.linalg.norm(g_mc - g_ex) <= 0.25 * np.linalg.norm(g_ex) + 10 * fluct / np.sqrt(chains * per_chain)


# ------------------------------------------------------------------------------------------------ BASELINE shapes

def _sample_confs(rng, N, ns):
    words = F.words_for(N)
    out = np.zeros((ns, words), dtype=np.uint64)
    for s in range(ns):
        v = int.from_bytes(rng.bytes(32), "little") & ((1 << N) - 1)
        for w in range(words):
            out[s, w] = (v >> (64 * w)) & 0xFFFFFFFFFFFFFFFF
    return out


@pytest.mark.parametrize("cfg", ["C2", "C5_small", "C3", "C4"])
def test_per_configuration_parity_at_baseline_shapes(gpu, port, cfg):
    """log psi, E_loc and O_k rows on identical configurations at the BASELINE.json shapes (C5 with M reduced to 400
    so the oracle finishes in seconds; the full-size C5 run is covered by test_c5_properties)."""
    if cfg == "C2":
        spec, H = F.config_C2()
    elif cfg == "C5_small":
        spec, H = F.config_C5(N=200, alpha=2)
    elif cfg == "C3":
        spec, H = F.config_C3()
    else:
        spec, H = F.config_C4()
    N = spec.num_sites
    pg, pp = make_psi(gpu, spec), make_psi(port, spec)
    og, op_ = make_op(gpu, H), make_op(port, H)
    confs = _sample_confs(np.random.default_rng(5), N, 48)
    lp_g, el_g = gpu.local_energies(pg, og, confs)
    lp_p, el_p, O_p = port.eval_samples(pp, op_, confs, want_O=True)
    assert rel_err(lp_g, lp_p) <= TOL
    assert rel_err(el_g, el_p) <= TOL
    for s in (0, 17, 47):
        assert rel_err(gpu.psi_O_k(pg, confs[s]), O_p[s]) <= TOL


def _diag_S(t, ns, P):
    O = t.O_k_samples.reshape(ns, P)
    w = t.weight_samples
    return (w[:, None] * np.abs(O) ** 2).sum(axis=0) - np.abs(t.O_k_vector) ** 2


def test_c2_tdvp_mc_step_properties(gpu, port):
    """The SR step at C2 size (RBM 64x256, Heisenberg ring, 8192 chains): oracle re-evaluation of the GPU's own
    samples, factorised vs dense S.v, linearity and Hermiticity of S.v, CG residual."""
    spec, H = F.config_C2()
    pg, pp = make_psi(gpu, spec), make_psi(port, spec)
    og, op_ = make_op(gpu, H), make_op(port, H)
    chains = 2048          # (the full 8192-chain step is what bench.py times)
    mc = gpu.MonteCarloSpins(chains, 1, 10, chains, True, seed=0xA11CE)
    t = gpu.TDVP(pg.num_params, True)
    t.eval_F(og, pg, mc)
    assert 0.05 < mc.acceptance_rate < 0.999
    # re-evaluate a slice of the GPU's samples with the oracle
    mc2 = gpu.MonteCarloSpins(chains, 1, 10, chains, True, seed=0xA11CE)
    confs, lp = mc2.sample(pg)
    el = t.E_local_samples
    idx = np.arange(0, chains, 67)
    lp_p, el_p, O_p = port.eval_samples(pp, op_, confs[idx], want_O=True)
    assert rel_err(lp[idx], lp_p) <= TOL and rel_err(el[idx], el_p) <= TOL
    w = t.weight_samples
    assert np.allclose(w, 1.0 / chains)
    assert abs(t.E_local - np.sum(w * el)) <= 1e-10 * abs(t.E_local)
    rng = np.random.default_rng(1)
    P = pg.num_params
    v1 = rng.normal(size=P) + 1j * rng.normal(size=P)
    v2 = rng.normal(size=P) + 1j * rng.normal(size=P)
    s1, s2 = t.S_dot_vector(v1), t.S_dot_vector(v2)
    assert rel_err(t.S_dot_vector(2.0 * v1 - 0.5j * v2), 2.0 * s1 - 0.5j * s2) <= 1e-10       # linearity
    assert abs(np.vdot(v2, s1) - np.conj(np.vdot(v1, s2))) <= 1e-9 * abs(np.vdot(v2, s1))     # Hermitian
    assert np.vdot(v1, s1).real > 0 and abs(np.vdot(v1, s1).imag) <= 1e-9 * abs(np.vdot(v1, s1))
    # rows of the oracle reproduce <O_k>, F on the slice-independent identities: compare S.v against numpy on a subsample
    O_slice = t.O_k_samples.reshape(chains, P)[idx]
    assert rel_err(O_slice, O_p) <= TOL
    # after O_k_samples the dense path is used: same answer as the factorised one
    assert rel_err(t.S_dot_vector(v1), s1) <= 1e-10
    # 2048 samples << P = 16384: S + 1e-3 diag(S) is ill-conditioned, CG needs O(10^3) iterations here
    x, it, rr = t.solve_cg(tol=1e-6, max_iter=6000, shift_abs=0.0, shift_rel=1e-3)
    assert rr <= 1e-6
    b = t.F_vector
    r = t.S_dot_vector(x) + 1e-3 * _diag_S(t, chains, P) * x - b
    assert np.linalg.norm(r) <= 1e-5 * np.linalg.norm(b)


def test_c5_properties(gpu, port):
    """RBM alpha=8, N=200 (4-word masks, M=1600 > the register-resident sampler): a reduced chain count of the C5 shard."""
    spec, H = F.config_C5()
    pg, pp = make_psi(gpu, spec), make_psi(port, spec)
    og, op_ = make_op(gpu, H), make_op(port, H)
    chains = 256
    mc = gpu.MonteCarloSpins(chains, 1, 2, chains, True, seed=5)
    t = gpu.TDVP(pg.num_params, True)
    t.eval_F(og, pg, mc)
    mc2 = gpu.MonteCarloSpins(chains, 1, 2, chains, True, seed=5)
    confs, lp = mc2.sample(pg)
    idx = np.arange(0, chains, 37)
    lp_p, el_p, _ = port.eval_samples(pp, op_, confs[idx])
    assert rel_err(lp[idx], lp_p) <= TOL and rel_err(t.E_local_samples[idx], el_p) <= TOL
    x, it, rr = t.solve_cg(tol=1e-6, max_iter=300, shift_abs=0.0, shift_rel=1e-3)
    assert rr <= 1e-6


# ------------------------------------------------------------------------------------------------ tensor-core S build

@pytest.mark.parametrize("name,ns_mc", [("deep2", 0), ("rbm12_m40", 0), ("deep2", 3000)])
def test_tensorcore_S_build_matches_fp64(gpu, name, ns_mc):
    """The opt-in tcgen05 path (3xTF32 split, fp32 accumulation in TMEM) against the exact fp64 S of the same samples:
    tolerance 1e-5 relative to ||S|| (BASELINE.json's fp32 tolerance); Hermitian to rounding."""
    spec, H, N = zoo()[name]
    psi, op = make_psi(gpu, spec), make_op(gpu, H)
    if ns_mc:
        ens = gpu.MonteCarloSpins(ns_mc, 1, 5, ns_mc, True, seed=11)
    else:
        ens = gpu.ExactSummationSpins(N)
        psi.normalize(ens)
    t = gpu.TDVP(psi.num_params, True)
    t.eval(op, psi, ens)
    S64 = t.S_matrix
    t.build_S_tensorcore()
    S32 = t.S_matrix
    scale = np.abs(S64).max()
    assert np.abs(S32 - S64).max() <= 1e-5 * scale
    assert np.abs(S32 - S32.conj().T).max() <= 1e-12 * scale
    # the same path selected from eval by its accuracy budget (same samples: a fresh ensemble with the same seed)
    ens2 = gpu.MonteCarloSpins(ns_mc, 1, 5, ns_mc, True, seed=11) if ns_mc else ens
    t2 = gpu.TDVP(psi.num_params, True)
    t2.eval(op, psi, ens2, s_tolerance=1e-5)
    assert np.array_equal(t2.S_matrix, S32) and np.abs(t2.F_vector - t.F_vector).max() <= 1e-14 * np.abs(t.F_vector).max()
    with pytest.raises(gpu.AngpuError):
        t2.eval(op, psi, ens2, s_tolerance=1e-8)
    # un-normalised ExactSummation weights (sum w != 1): the reference's convention S = sum w O*O - <O>*<O> is kept
    if not ns_mc:
        psi.log_prefactor = psi.log_prefactor + 0.3
        t.eval(op, psi, ens)
        S64 = t.S_matrix
        t.build_S_tensorcore()
        assert np.abs(t.S_matrix - S64).max() <= 1e-5 * np.abs(S64).max()


def test_cg_on_dense_S_matches_matrix_free_and_cholesky(gpu):
    """After eval() the CG products stream the dense S (k_smat_vec); ANGPU_CG_MATRIX_FREE=1 forces the O-based
    products.  Both, and the Cholesky solve, must give the same solution of (S + shift) x = F."""
    import os
    spec, H, N = zoo()["deep2"]
    psi, op = make_psi(gpu, spec), make_op(gpu, H)
    ens = gpu.MonteCarloSpins(4096, 1, 5, 4096, True, seed=5)
    t = gpu.TDVP(psi.num_params, True)
    t.eval(op, psi, ens)
    S, F = t.S_matrix, t.F_vector
    A = S + 1e-3 * np.diag(np.diag(S).real) + 1e-4 * np.eye(psi.num_params)
    x_d = t.solve(shift_abs=1e-4, shift_rel=1e-3)
    assert np.linalg.norm(A @ x_d - F) <= 1e-8 * np.linalg.norm(F)
    try:
        xs = {}
        for env in ("0", "1"):
            os.environ["ANGPU_CG_MATRIX_FREE"] = env
            x, it, rr = t.solve_cg(tol=1e-10, max_iter=20000, shift_abs=1e-4, shift_rel=1e-3)
            assert np.linalg.norm(A @ x - F) <= 1e-7 * np.linalg.norm(F), (env, it, rr)
            xs[env] = x
    finally:
        os.environ["ANGPU_CG_MATRIX_FREE"] = "0"
    assert rel_err(xs["0"], xs["1"]) <= 1e-5


def test_deep_block_sampler_matches_generic_at_c4_shape(gpu):
    """PsiDeep 64-64-64 (C4): the block-per-chain sampler (deep_kernels.cuh) and the generic warp-per-chain kernel run the
    same Philox stream; configurations must coincide and log psi agree to rounding."""
    import os
    spec, H = F.config_C4()
    psi = make_psi(gpu, spec)
    out = {}
    try:
        for mode in ("block", "generic"):
            os.environ["ANGPU_DEEP_SAMPLER"] = mode
            mc = gpu.MonteCarloSpins(98 * 2, 1, 4, 98, True, seed=21)      # 98: a ragged last block of the 4-chain kernel
            out[mode] = mc.sample(psi) + (mc.acceptances,)
    finally:
        os.environ["ANGPU_DEEP_SAMPLER"] = "block"
    assert np.array_equal(out["block"][0], out["generic"][0])
    assert rel_err(out["block"][1], out["generic"][1]) <= 1e-10
    assert out["block"][2] == out["generic"][2]


# ------------------------------------------------------------------------------------------------ SURVEY 8f rank 2

def test_reweighted_expectation_and_exp_sigma_z(gpu, port):
    """ExpectationValue(op, psi, psi_sampling, ens) and exp_sigma_z against the oracle: exact summation within 1e-10,
    Monte Carlo on identical chains (same Philox stream) within 1e-9."""
    specA, H, N = zoo()["rbm10"]
    specB = F.rbm_spec(10, 20, noise=4e-2, final_weight=8, seed=31)
    ev = gpu.ExpectationValue(True)
    pg, sg, og = make_psi(gpu, specA), make_psi(gpu, specB), make_op(gpu, H)
    pp, sp, op_ = make_psi(port, specA), make_psi(port, specB), make_op(port, H)
    eg, ep = gpu.ExactSummationSpins(N), port.ExactSummation(N)
    r_p = port.expectation_reweighted(op_, pp, sp, ep)
    assert abs(ev(og, pg, sg, eg) - r_p) <= TOL * max(1.0, abs(r_p))
    # reweighting with exact summation must reproduce the normalised direct expectation value
    direct = port.expectation(op_, pp, ep) / port.psi_norm(pp, ep) ** 2
    assert abs(ev(og, pg, sg, eg) - direct) <= 1e-9 * max(1.0, abs(direct))
    z_p = port.exp_sigma_z(op_, pp, ep)
    assert abs(ev.exp_sigma_z(og, pg, eg) - z_p) <= TOL * abs(z_p)
    mg, mp = gpu.MonteCarloSpins(512, 2, 5, 64, True, seed=77), port.MonteCarlo(512, 2, 5, 64, seed=77)
    r_p = port.expectation_reweighted(op_, pp, sp, mp)
    assert abs(ev(og, pg, sg, mg) - r_p) <= 1e-9 * max(1.0, abs(r_p))
    z_p = port.exp_sigma_z(op_, pp, mp)
    assert abs(ev.exp_sigma_z(og, pg, mg) - z_p) <= 1e-9 * abs(z_p)


@pytest.mark.parametrize("name", sorted(classical_zoo()))
def test_eval_with_psi_ref_matches_oracle(gpu, port, name):
    """TDVP.eval_with_psi_ref (samples from the classical state's reference, un-normalised reweighted sums)."""
    N, order, Hl, pr, ref_spec, lp, H = classical_zoo()[name]
    pg, pp = make_classical(gpu, N, order, Hl, pr, ref_spec, lp), make_classical(port, N, order, Hl, pr, ref_spec, lp)
    og, op_ = make_op(gpu, H), make_op(port, H)
    for eg, ep, tol in ((gpu.ExactSummationSpins(N), port.ExactSummation(N), TOL),
                        (gpu.MonteCarloSpins(256, 1, 4, 32, True, seed=5), port.MonteCarlo(256, 1, 4, 32, seed=5), 1e-9)):
        tg, tp = gpu.TDVP(pg.num_params, True), port.TDVP(pp.num_params)
        tg.eval_with_psi_ref(og, pg, eg)
        tp.eval_with_psi_ref(op_, pp, ep)
        assert abs(tg.total_weight - tp.total_weight) <= tol * tp.total_weight
        assert abs(tg.E_local - tp.E_local) <= tol * max(1.0, abs(tp.E_local))
        assert rel_err(tg.F_vector, tp.F_vector) <= tol and rel_err(tg.O_k_vector, tp.O_k_vector) <= tol
        assert rel_err(tg.S_matrix, tp.S_matrix) <= tol


def test_eval_with_explicit_sampling_state(gpu, port):
    """Additive: any psi can be importance-sampled from another (here PsiDeep from a PsiRBM)."""
    spec, H, N = zoo()["deep2"]
    samp = F.rbm_spec(8, 16, noise=3e-2, final_weight=2, seed=41)
    pg, sg, og = make_psi(gpu, spec), make_psi(gpu, samp), make_op(gpu, H)
    pp, sp, op_ = make_psi(port, spec), make_psi(port, samp), make_op(port, H)
    tg, tp = gpu.TDVP(pg.num_params, True), port.TDVP(pp.num_params)
    tg.eval_with_psi_ref(og, pg, gpu.ExactSummationSpins(N), psi_sampling=sg)
    tp.eval_with_psi_ref(op_, pp, port.ExactSummation(N), psi_sampling=sp)
    assert rel_err(tg.S_matrix, tp.S_matrix) <= TOL and rel_err(tg.F_vector, tp.F_vector) <= TOL
    assert abs(tg.total_weight - tp.total_weight) <= TOL * tp.total_weight


# ------------------------------------------------------------------------------------------------ SURVEY 8f rank 1

@pytest.mark.parametrize("pair", ["rbm_rbm", "deep_rbm", "rbm_wide"])
def test_hilbert_space_distance_monte_carlo_and_mixed_models(gpu, port, pair):
    """HilbertSpaceDistance on Monte-Carlo samples (identical chains) and on model pairs the reference does not
    instantiate, incl. the factorised PsiRBM rows (M = 600 > the reference's limit) as psi_prime."""
    if pair == "rbm_rbm":
        spec, H, N = zoo()["rbm10"]
        spec_p = F.rbm_spec(10, 20, noise=4e-2, final_weight=9, seed=32)
    elif pair == "deep_rbm":
        spec, H, N = zoo()["deep2"]
        spec_p = F.rbm_spec(8, 16, noise=3e-2, final_weight=2, seed=41)
    else:
        spec, H, N = wide_zoo()["rbm_wide"]
        spec_p = F.rbm_spec(12, 600, noise=2.5e-2, final_weight=0.5, seed=14)
    U = F.propagator(H, 0.03)
    pg, qg, og = make_psi(gpu, spec), make_psi(gpu, spec_p), make_op(gpu, U)
    pp, qp, op_ = make_psi(port, spec), make_psi(port, spec_p), make_op(port, U)
    # normalise both states (un-normalised exact-summation weights of the wide RBM under/overflow v^2 in the gradient)
    for a, b in ((pg, pp), (qg, qp)):
        lp = b.log_prefactor - np.log(port.psi_norm(b, port.ExactSummation(N)))
        a.log_prefactor = b.log_prefactor = lp
    hsd = gpu.HilbertSpaceDistance(qg.num_params, True)
    for eg, ep, tol in ((gpu.ExactSummationSpins(N), port.ExactSummation(N), 1e-9),
                        (gpu.MonteCarloSpins(384, 1, 4, 48, True, seed=9), port.MonteCarlo(384, 1, 4, 48, seed=9), 1e-8)):
        d_p = port.hilbert_space_distance(pp, qp, op_, True, ep)
        assert abs(hsd(pg, qg, og, True, eg) - d_p) <= tol
        if isinstance(ep, port.MonteCarlo):       # fresh ensembles: the Philox call counter advanced above
            eg, ep = gpu.MonteCarloSpins(384, 1, 4, 48, True, seed=9), port.MonteCarlo(384, 1, 4, 48, seed=9)
        g_p, d_p = port.hilbert_space_distance_gradient(pp, qp, op_, True, ep, 0.5)
        g_g, d_g = hsd.gradient(pg, qg, og, True, eg, 0.5)
        assert abs(d_g - d_p) <= tol and rel_err(g_g, g_p) <= 1e-7


@pytest.mark.parametrize("name", ["cnn", "cnn_sym", "C3"])
def test_cnn_incremental_sampler_bit_identical_to_full_forward(gpu, name):
    """The incremental PsiCNN sampler (cnn_kernels.cuh: only the receptive cone of the flipped site is recomputed) must
    reproduce the full-forward kernel exactly: same configurations, same acceptances, log psi equal bit for bit."""
    import os
    spec = F.config_C3()[0] if name == "C3" else zoo()[name][0]
    psi = make_psi(gpu, spec)
    out = {}
    try:
        for mode in ("incremental", "generic"):
            os.environ["ANGPU_CNN_SAMPLER"] = mode
            mc = gpu.MonteCarloSpins(37 * 3, 2, 3, 37, True, seed=17)
            out[mode] = mc.sample(psi) + (mc.acceptances,)
    finally:
        os.environ["ANGPU_CNN_SAMPLER"] = "incremental"
    assert np.array_equal(out["incremental"][0], out["generic"][0])
    assert np.array_equal(out["incremental"][1], out["generic"][1])
    assert out["incremental"][2] == out["generic"][2]


# ------------------------------------------------------------------------------------------------ SURVEY 8f rank 4 (wire format)

@pytest.mark.parametrize("name", ["rbm8_cfw", "deep3", "cnn_sym"])
def test_json_round_trip(gpu, name):
    """psi.to_json() (the reference's key names and ndarray encoding, pyANNonGPU/Psi*.py) survives json.dumps/loads and
    rebuilds a state with identical parameters and amplitudes."""
    import json
    spec, H, N = zoo()[name]
    psi = make_psi(gpu, spec)
    psi.log_prefactor = 0.3 - 0.2j
    obj = json.loads(json.dumps(psi.to_json()))
    assert obj["type"] == type(psi).__name__
    psi2 = type(psi).from_json(obj, True)
    assert np.array_equal(psi2.params, psi.params) and psi2.log_prefactor == psi.log_prefactor
    es = gpu.ExactSummationSpins(N)
    assert np.array_equal(gpu.log_psi_vector(psi2, es), gpu.log_psi_vector(psi, es))


@pytest.mark.parametrize("name", ["cnn", "cnn_sym", "C3"])
def test_cnn_cone_local_energy_bit_identical_to_full_forward(gpu, name):
    """The cone-based PsiCNN E_loc (only the union of the receptive cones of a flip group is recomputed) against the
    generic kernel (one full forward pass per flip group): identical local energies, for two different operators in a row
    (the per-operator cone lists are cached by content)."""
    import os
    if name == "C3":
        spec, H = F.config_C3()
        N = 100
    else:
        spec, H, N = zoo()[name]
    H2 = F.tfim(N, F.ring_bonds(N), h=0.7)
    psi = make_psi(gpu, spec)
    rng = np.random.default_rng(5)
    confs = _sample_confs(rng, N, 40)
    out = {}
    try:
        for mode in ("cone", "generic"):
            os.environ["ANGPU_CNN_ELOC"] = mode
            out[mode] = [gpu.local_energies(psi, make_op(gpu, h), confs)[1] for h in (H, H2, H)]
    finally:
        os.environ["ANGPU_CNN_ELOC"] = "cone"
    for a, b in zip(out["cone"], out["generic"]):
        assert rel_err(a, b) <= 1e-13
    assert np.array_equal(out["cone"][0], out["cone"][2])


@pytest.mark.parametrize("name", ["deep2", "deep3", "C4"])
def test_deep_block_local_energy_matches_generic(gpu, name):
    """PsiDeep E_loc with register-resident deep layers (block per sample, 4 flip groups per round) against the generic
    warp-per-sample kernel."""
    import os
    if name == "C4":
        spec, H = F.config_C4()
        N = 64
    else:
        spec, H, N = zoo()[name]
    psi, op = make_psi(gpu, spec), make_op(gpu, H)
    confs = _sample_confs(np.random.default_rng(9), N, 37)
    out = {}
    try:
        for mode in ("block", "generic"):
            os.environ["ANGPU_DEEP_ELOC"] = mode
            out[mode] = gpu.local_energies(psi, op, confs)[1]
    finally:
        os.environ["ANGPU_DEEP_ELOC"] = "block"
    assert rel_err(out["block"], out["generic"]) <= 1e-11


# ------------------------------------------------------------------------------------------------ SURVEY 8f rank 3 (KullbackLeibler)

def test_kullback_leibler_monte_carlo_and_rbm_prime(gpu, port):
    """KullbackLeibler on Monte-Carlo samples (identical chains) and with a PsiRBM as psi_prime (factorised rows), which
    the reference does not instantiate; the state carried between calls (mean deviation) must agree as well."""
    N, order, Hl, pr, ref_spec, lp, H = classical_zoo()["clfp2"]
    spec_p = F.rbm_spec(6, 12, noise=5e-2, final_weight=2, seed=51)
    pg, pp = make_classical(gpu, N, order, Hl, pr, ref_spec, lp), make_classical(port, N, order, Hl, pr, ref_spec, lp)
    qg, qp = make_psi(gpu, spec_p), make_psi(port, spec_p)
    kg, kp = gpu.KullbackLeibler(qg.num_params, True), port.KullbackLeibler(qp.num_params)
    for call in range(2):
        mg, mp = gpu.MonteCarloSpins(512, 1, 4, 64, True, seed=3 + call), port.MonteCarlo(512, 1, 4, 64, seed=3 + call)
        g_g, v_g = kg.gradient(pg, qg, mg, 1.0, 0.0)
        g_p, v_p = kp.gradient(pp, qp, mp, 1.0, 0.0)
        assert abs(v_g - v_p) <= 1e-9 * max(1.0, v_p) and rel_err(g_g, g_p) <= 1e-8
        assert abs(kg.mean_deviation - kp.mean_deviation) <= 1e-9 and abs(kg.total_weight - kp.total_weight) <= 1e-9 * kp.total_weight
    eg, ep = gpu.ExactSummationSpins(N), port.ExactSummation(N)
    gn_g, n_g, v_g = kg.gradient_with_noise(pg, qg, eg, 0.0, 1e-2)
    gn_p, n_p, v_p = kp.gradient_with_noise(pp, qp, ep, 0.0, 1e-2)
    assert abs(v_g - v_p) <= 1e-9 * max(1.0, v_p) and rel_err(gn_g, gn_p) <= 1e-8
    ok = np.isfinite(n_p)
    assert np.array_equal(np.isfinite(n_g), ok) and np.allclose(n_g[ok], n_p[ok], rtol=1e-6, atol=1e-12)
