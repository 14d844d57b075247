"""JSON (de)serialisation of numpy arrays in the reference's wire format (pyANNonGPU/json_numpy.py:5-38):
``{"type": "ndarray", "dtype": "...", "complex": bool, "real": [...], "imag": [...]}`` or ``"data": [...]`` for real
arrays, so that files written by pyANNonGPU's ``psi.to_json()`` load here and vice versa."""
import json

import numpy as np


class NumpyEncoder(json.JSONEncoder):
    def default(self, o):
        if isinstance(o, np.ndarray):
            if np.iscomplexobj(o):
                return {"type": "ndarray", "dtype": str(o.dtype), "complex": True, "real": o.real.tolist(), "imag": o.imag.tolist()}
            return {"type": "ndarray", "dtype": str(o.dtype), "complex": False, "data": o.tolist()}
        if isinstance(o, (np.integer,)):
            return int(o)
        if isinstance(o, (np.floating,)):
            return float(o)
        if isinstance(o, complex):
            return {"type": "complex", "real": o.real, "imag": o.imag}
        return super().default(o)


def _hook(d):
    if d.get("type") == "ndarray":
        dt = np.dtype(d["dtype"])
        if d["complex"]:
            return np.array(d["real"], dtype=dt) + 1j * np.array(d["imag"], dtype=dt)
        return np.array(d["data"], dtype=dt)
    if d.get("type") == "complex":
        return complex(d["real"], d["imag"])
    return d


class NumpyDecoder(json.JSONDecoder):
    def __init__(self, *args, **kwargs):
        kwargs.setdefault("object_hook", _hook)
        super().__init__(*args, **kwargs)


def plain(obj):
    """obj -> JSON-compatible python structure (what the reference's to_json returns)."""
    return json.loads(json.dumps(obj, cls=NumpyEncoder))


def restore(obj):
    """Inverse of `plain` (accepts an already decoded structure as well)."""
    return json.loads(json.dumps(obj, cls=NumpyEncoder), cls=NumpyDecoder)
