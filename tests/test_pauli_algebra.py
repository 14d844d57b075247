"""The native Pauli-operator algebra (annongpu_b200.factories.PauliSum: +, *, scalar, dagger, roll, matrix) against dense
matrices built from the oracle's PauliString::apply semantics (include/basis/PauliString.hpp:193-255)."""
import numpy as np
import pytest

from annongpu_b200 import factories as F


def random_sum(rng, n, terms):
    H = F.PauliSum(n)
    for _ in range(terms):
        paulis = {int(i): "XYZ"[int(rng.integers(0, 3))] for i in rng.choice(n, size=int(rng.integers(0, n + 1)), replace=False)}
        H.add(complex(rng.normal(), rng.normal()), paulis)
    return H


def test_matrix_matches_oracle_apply_semantics(port):
    rng = np.random.default_rng(1)
    H = random_sum(rng, 5, 7)
    c, a, b = H.arrays(1)
    assert np.allclose(H.matrix(), port.Operator(c, a, b).dense_matrix(5), atol=1e-14)


def test_single_site_products():
    X, Y, Z = (F.PauliSum(1).add(1.0, {0: k}) for k in "XYZ")
    I = F.PauliSum(1).add(1.0, {})
    for A, B, C in ((X, Y, Z), (Y, Z, X), (Z, X, Y)):
        assert np.allclose((A * B).matrix(), 1j * C.matrix())
        assert np.allclose((B * A).matrix(), -1j * C.matrix())
        assert np.allclose((A * A).matrix(), I.matrix())


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_algebra_matches_dense_linear_algebra(seed):
    rng = np.random.default_rng(seed)
    n = 4
    A, B = random_sum(rng, n, 5), random_sum(rng, n, 6)
    MA, MB = A.matrix(), B.matrix()
    assert np.allclose((A * B).matrix(), MA @ MB, atol=1e-12)
    assert np.allclose((A + B).matrix(), MA + MB, atol=1e-12)
    assert np.allclose((A - 2.5j * B).matrix(), MA - 2.5j * MB, atol=1e-12)
    assert np.allclose((0.5 + A).matrix(), 0.5 * np.eye(1 << n) + MA, atol=1e-12)
    assert np.allclose(A.dagger().matrix(), MA.conj().T, atol=1e-12)
    assert np.allclose(A.commutator(B).matrix(), MA @ MB - MB @ MA, atol=1e-12)


def test_roll_translates_sites_on_the_ring():
    n = 5
    H = F.PauliSum(n).add(1.0, {0: "X", 1: "Z"}).add(0.5j, {4: "Y"})
    R = H.roll(2)
    expect = F.PauliSum(n).add(1.0, {2: "X", 3: "Z"}).add(0.5j, {1: "Y"})
    assert np.allclose(R.matrix(), expect.matrix())
    assert np.allclose(H.roll(n).matrix(), H.matrix())
    # the Heisenberg ring is translation invariant
    Hh = F.heisenberg(n, F.ring_bonds(n))
    assert np.allclose(Hh.roll(1).matrix(), Hh.matrix())


def test_heisenberg_matrix_is_hermitian_with_known_ground_state_energy():
    # 4-site Heisenberg ring in Pauli-matrix normalisation: E0 = -8
    M = F.heisenberg(4, F.ring_bonds(4)).matrix()
    assert np.allclose(M, M.conj().T)
    assert abs(np.linalg.eigvalsh(M).min() + 8.0) < 1e-12
