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
