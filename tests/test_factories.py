"""The BASELINE.json configurations as built by annongpu_b200.factories: shapes, parameter counts and operator sizes fixed
in SURVEY.md §8 (C1..C5), and the seeded network factories' invariants (pyANNonGPU/new_*.py restatements)."""
import numpy as np

from annongpu_b200 import factories as F


def num_params_deep(spec):
    return len(spec.input_weights) + sum(b.size + w.size for b, w in zip(spec.biases, spec.weights))


def test_baseline_shapes():
    s1, H1 = F.config_C1()
    assert s1.W.shape == (16, 32) and s1.W.size == 512 and H1.num_strings == 32
    s2, H2 = F.config_C2()
    assert s2.W.shape == (64, 256) and s2.W.size == 16384 and H2.num_strings == 192
    s3, H3 = F.config_C3()
    assert list(s3.extent) == [1, 10, 10] and len(s3.params) == 189 and H3.num_strings == 600 and H3.words == 2
    s4, H4 = F.config_C4()
    assert num_params_deep(s4) == 8384 and H4.num_strings == 192
    s5, H5 = F.config_C5()
    assert s5.W.shape == (200, 1600) and s5.W.size == 320000 and H5.num_strings == 600 and H5.words == 4


def test_factories_are_seeded_and_regular():
    a, b = F.rbm_spec(8, 16, seed=3), F.rbm_spec(8, 16, seed=3)
    assert np.array_equal(a.W, b.W) and not np.array_equal(a.W, F.rbm_spec(8, 16, seed=4).W)
    d = F.deep_spec(8, 8, [16, 8], [4, 8], seed=1)
    for c, w, prev in zip(d.connections, d.weights, [8, 16]):
        assert c.shape == w.shape and c.max() < prev
        # every unit of the previous layer feeds the same number of units (needed for the rhs tables, PsiDeep.cu:214-242)
        counts = np.bincount(c.ravel(), minlength=prev)
        assert counts.min() == counts.max()
    c = F.cnn_spec([4, 4], [(2, [2, 2]), (3, [3, 3])], seed=2)
    assert len(c.params) == (1 * 2 * 4 + 2 * 3 * 9) * 1 and list(c.extent) == [1, 4, 4]


def test_lattice_bonds():
    assert len(F.ring_bonds(7)) == 7
    bonds = F.square_lattice_bonds(8, 8)
    assert len(bonds) == 128 and all(0 <= i < 64 and 0 <= j < 64 and i != j for i, j in bonds)
    H = F.tfim(64, bonds)
    assert H.num_strings == 128 + 64
