"""Edge cases of the hot path through the C ABI, against the CPU oracle: ragged sample counts, single chains, word
boundaries of the configuration masks, degenerate operators, empty shards, and the error behaviour of the boundary."""
import numpy as np
import pytest

from annongpu_b200 import factories as F
from helpers import make_op, make_psi, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _pair(gpu, port, spec, H):
    return make_psi(gpu, spec), make_psi(port, spec), make_op(gpu, H), make_op(port, H)


@pytest.mark.parametrize("num_samples,chains", [(10, 4), (7, 1), (33, 32), (5, 5)])
def test_ragged_sample_counts_follow_the_reference_integer_division(gpu, port, num_samples, chains):
    """num_mc_steps_per_chain = num_samples / num_chains (integer division), weight = 1 / num_samples
    (source/ensembles/MonteCarlo.cu:33-35): samples beyond chains * steps are never produced."""
    spec = F.rbm_spec(9, 18, noise=5e-2, final_weight=4, seed=3)
    H = F.heisenberg(9, F.ring_bonds(9))
    pg, pp, og, op_ = _pair(gpu, port, spec, H)
    mg, mp = gpu.MonteCarloSpins(num_samples, 1, 2, chains, True, seed=8), port.MonteCarlo(num_samples, 1, 2, chains, seed=8)
    cg, lg = mg.sample(pg)
    cp_, lp_ = mp.sample(pp)
    assert len(cg) == (num_samples // chains) * chains == len(cp_)
    assert np.array_equal(cg, cp_) and rel_err(lg, lp_) <= 1e-9
    tg, tp = gpu.TDVP(pg.num_params, True), port.TDVP(pp.num_params)
    mg, mp = gpu.MonteCarloSpins(num_samples, 1, 2, chains, True, seed=8), port.MonteCarlo(num_samples, 1, 2, chains, seed=8)
    tg.eval(og, pg, mg)
    tp.eval(op_, pp, mp)
    assert abs(tg.total_weight - (num_samples // chains) * chains / num_samples) <= 1e-14
    assert abs(tg.E_local - tp.E_local) <= 1e-9 * max(1.0, abs(tp.E_local))
    assert rel_err(tg.F_vector, tp.F_vector) <= 1e-8 and rel_err(tg.S_matrix, tp.S_matrix) <= 1e-8


@pytest.mark.parametrize("N", [1, 2, 63, 64, 65, 128, 129, 256])
def test_word_boundaries_of_the_configuration_masks(gpu, port, N):
    """Sites 63/64/65, 128/129 and the 256-site maximum: bit-exact Pauli action and log psi / O_k / E_loc parity."""
    M = max(N, 4)
    spec = F.rbm_spec(N, M, noise=2e-2, final_weight=1.5, seed=N)
    H = F.PauliSum(N)
    for i in {0, N // 2, N - 1}:
        H.add(0.7, {i: "Z", (i + 1) % N: "Z"} if N > 1 else {i: "Z"}).add(-0.4 + 0.1j, {i: "X"}).add(0.3, {i: "Y", (i + N // 2) % N: "Z"} if N > 2 else {i: "Y"})
    pg, pp, og, op_ = _pair(gpu, port, spec, H)
    rng = np.random.default_rng(N)
    words = F.words_for(N)
    for _ in range(4):
        v = int.from_bytes(rng.bytes(32), "little") & ((1 << N) - 1)
        conf = np.array([(v >> (64 * w)) & 0xFFFFFFFFFFFFFFFF for w in range(words)], dtype=np.uint64)
        lp_p = port.log_psi_s(pp, conf)
        assert abs(gpu.log_psi_s(pg, conf) - lp_p) <= TOL * max(1.0, abs(lp_p))
        assert rel_err(gpu.psi_O_k(pg, conf), port.psi_O_k(pp, conf)) <= TOL
        assert abs(gpu.local_energies(pg, og, conf[None, :])[1][0] - port.local_energy(pp, op_, conf)) <= 1e-9
    chains = 6
    mg, mp = gpu.MonteCarloSpins(chains, 1, 1, chains, True, seed=N), port.MonteCarlo(chains, 1, 1, chains, seed=N)
    assert np.array_equal(mg.sample(pg)[0], mp.sample(pp)[0])


def test_degenerate_operators(gpu, port):
    """Identity only, diagonal only, and flip groups of 3, 4 (RBM fast path) and 5 sites (generic fallback)."""
    N = 8
    spec = F.rbm_spec(N, 16, noise=5e-2, final_weight=3, seed=12)
    es_g, es_p = gpu.ExactSummationSpins(N), port.ExactSummation(N)
    ev = gpu.ExpectationValue(True)
    ops = {
        "identity": F.PauliSum(N).add(2.5 - 1j, {}),
        "diagonal": F.PauliSum(N).add(1.0, {0: "Z", 3: "Z"}).add(-0.5, {5: "Z"}),
        "flip3": F.PauliSum(N).add(0.3, {0: "X", 1: "Y", 2: "X"}).add(1.0, {4: "Z"}),
        "flip4": F.PauliSum(N).add(0.2j, {0: "X", 2: "X", 4: "Y", 6: "X"}),
        "flip5": F.PauliSum(N).add(0.1, {0: "X", 1: "X", 2: "Y", 3: "X", 4: "X"}).add(0.4, {7: "X"}),
    }
    for name, H in ops.items():
        pg, pp, og, op_ = _pair(gpu, port, spec, H)
        E_p = port.expectation(op_, pp, es_p)
        assert abs(ev(og, pg, es_g) - E_p) <= TOL * max(1.0, abs(E_p)), name
        g_g, _ = ev.gradient(og, pg, es_g)
        g_p, _ = port.gradient(op_, pp, es_p)
        assert np.abs(g_g - g_p).max() <= 1e-9 * max(1.0, np.abs(g_p).max()), name
    pg = make_psi(gpu, spec)
    pg.normalize(es_g)
    assert abs(ev(make_op(gpu, ops["identity"]), pg, es_g) - (2.5 - 1j)) <= 1e-12


def test_empty_shards_and_sharded_sums(gpu, port):
    """More ranks than chains: the ranks without chains contribute exact zeros; the shards add up to the full result."""
    spec = F.rbm_spec(8, 16, noise=5e-2, final_weight=3, seed=2)
    H = F.heisenberg(8, F.ring_bonds(8))
    pg, og = make_psi(gpu, spec), make_op(gpu, H)
    full = gpu.MonteCarloSpins(6, 1, 2, 3, True, seed=4)
    t = gpu.TDVP(pg.num_params, True)
    t.eval_F(og, pg, full)
    E_full, F_full, O_full = t.E_local, t.F_vector, t.O_k_vector
    world, E_sum, O_sum, n_local = 5, 0.0, 0.0, []
    for rank in range(world):
        m = gpu.MonteCarloSpins(6, 1, 2, 3, True, seed=4).set_shard(rank, world)
        n_local.append(m.local_steps)
        ts = gpu.TDVP(pg.num_params, True)
        ts.eval_F(og, pg, m)          # without an all-reduce hook every shard reports its own partial sums
        E_sum, O_sum = E_sum + ts.E_local, O_sum + ts.O_k_vector
        if m.local_steps == 0:
            assert ts.E_local == 0 and not np.any(ts.O_k_vector) and ts.total_weight == 0.0
    assert sum(n_local) == 6 and 0 in n_local
    assert abs(E_sum - E_full) <= 1e-12 * max(1.0, abs(E_full)) and rel_err(O_sum, O_full) <= 1e-12
    # exact summation: shards of the basis
    es_full = gpu.ExactSummationSpins(8)
    ev = gpu.ExpectationValue(True)
    parts = [ev(og, pg, gpu.ExactSummationSpins(8).set_shard(r, 3)) for r in range(3)]
    assert abs(sum(parts) - ev(og, pg, es_full)) <= 1e-12 * abs(ev(og, pg, es_full))


def test_boundary_errors(gpu):
    """Errors surface as exceptions with a message (CUDA_CHECK -> std::runtime_error -> RuntimeError in the reference)."""
    spec = F.rbm_spec(8, 16, noise=5e-2, final_weight=3, seed=2)
    psi = make_psi(gpu, spec)
    H = make_op(gpu, F.heisenberg(8, F.ring_bonds(8)))
    with pytest.raises(Exception):
        gpu.PsiRBM(spec.W, spec.final_weight, 0.0, False)                       # no CPU fallback
    with pytest.raises(RuntimeError, match="num_params"):
        gpu.TDVP(psi.num_params + 1, True).eval_F(H, psi, gpu.MonteCarloSpins(8, 1, 1, 8, True))
    with pytest.raises(RuntimeError, match="num_sites"):
        gpu.ExpectationValue(True)(H, psi, gpu.ExactSummationSpins(7))
    with pytest.raises(RuntimeError, match="eval"):
        gpu.TDVP(psi.num_params, True).solve_cg()
    H100 = make_op(gpu, F.heisenberg(100, F.ring_bonds(100)))
    with pytest.raises(RuntimeError, match="sites"):                            # a 100-site operator on an 8-site state
        gpu.ExpectationValue(True)(H100, psi, gpu.MonteCarloSpins(8, 1, 1, 8, True))
    H12 = make_op(gpu, F.heisenberg(12, F.ring_bonds(12)))
    with pytest.raises(RuntimeError, match="acts on site"):                     # same word count, still too wide
        gpu.ExpectationValue(True)(H12, psi, gpu.MonteCarloSpins(8, 1, 1, 8, True))
    t = gpu.TDVP(psi.num_params, True)
    t.eval(H, psi, gpu.ExactSummationSpins(8))
    with pytest.raises(RuntimeError, match="positive definite"):
        t.solve(shift_abs=-10.0, shift_rel=0.0)


def test_destroyed_psi_is_an_error_not_a_use_after_free(gpu):
    """TDVP materialises the dense O_k rows lazily from the psi of its last eval; a C-ABI caller that destroyed that psi
    in between gets an error message (the Python wrapper normally keeps the psi alive through TDVP._keep)."""
    spec = F.rbm_spec(8, 16, noise=5e-2, final_weight=3, seed=2)
    psi = make_psi(gpu, spec)
    H = make_op(gpu, F.heisenberg(8, F.ring_bonds(8)))
    t = gpu.TDVP(psi.num_params, True)
    t.eval_F(H, psi, gpu.ExactSummationSpins(8))
    t._keep = None
    psi.__del__()                                   # angpu_psi_destroy
    with pytest.raises(RuntimeError, match="destroyed"):
        t.O_k_samples
