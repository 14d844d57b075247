"""The ndarray wire format of annongpu_b200.json_numpy against the reference's layout (pyANNonGPU/json_numpy.py:5-38)."""
import json

import numpy as np

from annongpu_b200.json_numpy import NumpyDecoder, NumpyEncoder, plain, restore


def test_ndarray_encoding_matches_reference_layout():
    z = np.array([[1 + 2j, 3 - 4j]], dtype=np.complex128)
    enc = json.loads(json.dumps(z, cls=NumpyEncoder))
    assert enc == {"type": "ndarray", "dtype": "complex128", "complex": True, "real": [[1.0, 3.0]], "imag": [[2.0, -4.0]]}
    u = np.arange(4, dtype=np.uint32).reshape(2, 2)
    enc = json.loads(json.dumps(u, cls=NumpyEncoder))
    assert enc == {"type": "ndarray", "dtype": "uint32", "complex": False, "data": [[0, 1], [2, 3]]}


def test_round_trip_nested():
    obj = {"type": "PsiDeep", "W": [np.ones((2, 3)) * (1 + 1j), np.zeros((3, 1), dtype=complex)], "c": [np.arange(6, dtype=np.uint32).reshape(2, 3)], "x": 1.5}
    back = restore(plain(obj))
    assert back["type"] == "PsiDeep" and back["x"] == 1.5
    for a, b in zip(obj["W"], back["W"]):
        assert a.dtype == b.dtype and np.array_equal(a, b)
    assert back["c"][0].dtype == np.uint32 and np.array_equal(back["c"][0], obj["c"][0])
    assert json.loads(json.dumps(plain(obj)), cls=NumpyDecoder)["W"][0].shape == (2, 3)


def test_documents_written_by_the_reference_encoder_decode(tmp_path):
    """tests/golden/ref_json_blobs.json was produced by the reference's own NumpyEncoder (tests/golden/make_json_fixture.py)
    with the keys of the reference's to_json methods: the decoder here must give back the arrays bit for bit."""
    import os
    from annongpu_b200 import factories as F
    with open(os.path.join(os.path.dirname(__file__), "golden", "ref_json_blobs.json")) as f:
        blobs = json.load(f)
    rbm = restore(blobs["PsiRBM"])
    spec = F.rbm_spec(6, 12, noise=0.05, final_weight=3.0, seed=101)
    assert rbm["type"] == "PsiRBM" and rbm["W"].dtype == np.complex128 and np.array_equal(rbm["W"], spec.W)
    assert rbm["final_weight"] == 3.0 and (rbm["log_prefactor_re"], rbm["log_prefactor_im"]) == (0.25, -0.5)
    deep = restore(blobs["PsiDeep"])
    dspec = F.deep_spec(6, 6, [6, 3], [3, 2], noise=0.05, final_weights=2.0, seed=102)
    assert np.array_equal(deep["a"], dspec.input_weights)
    for got, want in zip(deep["W"], dspec.weights):
        assert np.array_equal(got, want)
    for got, want in zip(deep["connections"], dspec.connections):
        assert got.dtype == want.dtype and np.array_equal(got, want)
    cnn = restore(blobs["PsiCNN"])
    cspec = F.cnn_spec([1, 2, 3], [(2, [1, 2, 2]), (1, [1, 2, 3])], noise=0.05, final_factor=1.5, seed=103)
    assert cnn["extent"].dtype == np.uint32 and np.array_equal(cnn["params"], cspec.params)
    assert np.array_equal(cnn["connectivity_list"], np.asarray(cspec.connectivity_list))
    # and the documents written here are byte-identical to the reference encoder's for the same content
    mine = plain(dict(type="PsiRBM", W=spec.W, final_weight=3.0, log_prefactor_re=0.25, log_prefactor_im=-0.5))
    assert mine == blobs["PsiRBM"]


def test_pauli_sum_json_and_exp():
    import scipy.linalg
    from annongpu_b200.factories import PauliSum, sigma_x, sigma_y, sigma_z
    h = sigma_z(0) * sigma_z(1) + 1.1 * sigma_x(0) + (0.3 - 0.2j) * sigma_y(2) * sigma_x(0)
    back = PauliSum.from_json(json.loads(json.dumps(h.to_json())))
    assert np.array_equal(back.matrix(), h.matrix())
    for term in (0.4j * sigma_x(1) * sigma_y(2), (-0.2 + 0.1j) * sigma_z(0), 0.7 * sigma_x(0, 3) * sigma_x(0, 3)):
        assert np.abs(term.exp(0).matrix() - scipy.linalg.expm(term.matrix())).max() <= 1e-14
    import pytest
    with pytest.raises(ValueError):
        h.exp()
