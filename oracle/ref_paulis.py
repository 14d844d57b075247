"""TEST INFRASTRUCTURE — ctypes loader for ``oracle/_ref/liboracle_ref_paulis.so``: the UNMODIFIED reference compiled with
``-DENABLE_PAULIS`` (Pauli-string / density-matrix basis) and PsiDeep only, by ``oracle/Makefile`` (target ``ref_paulis``) plus
``oracle/ref_shim/shim_paulis.cu``.  Used by ``tests/golden/make_golden_paulis.py`` and ``tests/test_oracle_pinned.py`` only.
Limits of the true reference: <= 64 sites, ExactSummationPaulis <= 15 sites (4^N configurations in an unsigned int)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "liboracle_ref_paulis.so")
ES, MC = 0, 1
_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        vp, u32, u64, dbl, i32 = C.c_void_p, C.c_uint, C.c_uint64, C.c_double, C.c_int
        for name in ("refp_op_create", "refp_deep_create", "refp_es_create", "refp_mc_create", "refp_tdvp_create"):
            getattr(L, name).restype = vp
        L.refp_psi_num_params.restype = u32
        L.refp_ens_num_steps.restype = u32
        L.refp_network_unit_at.restype = i32
        L.refp_enumerate.argtypes = [u32, vp, vp]
        L.refp_pauli_mul.argtypes = [u64, u64, u64, u64, vp, vp, vp]
        L.refp_network_unit_at.argtypes = [u64, u64, u32]
        L.refp_op_create.argtypes = [u32, vp, vp, vp]
        L.refp_op_destroy.argtypes = [vp]
        L.refp_deep_create.argtypes = [u32, u32, vp, u32, vp, vp, vp, vp, vp, vp, dbl, dbl]
        L.refp_psi_destroy.argtypes = [vp]
        L.refp_psi_num_params.argtypes = [vp]
        L.refp_es_create.argtypes = [u32]
        L.refp_mc_create.argtypes = [u32, u32, u32, u32]
        L.refp_ens_destroy.argtypes = [i32, vp]
        L.refp_ens_num_steps.argtypes = [i32, vp]
        L.refp_log_psi_s.argtypes = [vp, u64, u64, vp]
        L.refp_psi_O_k.argtypes = [vp, u64, u64, vp]
        L.refp_log_psi_vector.argtypes = [vp, i32, vp, vp]
        L.refp_expectation.argtypes = [vp, vp, i32, vp, vp]
        L.refp_fluctuation.argtypes = [vp, vp, i32, vp, vp]
        L.refp_gradient.argtypes = [vp, vp, i32, vp, vp, vp]
        L.refp_tdvp_create.argtypes = [u32]
        L.refp_tdvp_destroy.argtypes = [vp]
        L.refp_tdvp_eval.argtypes = [vp, vp, vp, i32, vp]
        L.refp_tdvp_get.argtypes = [vp, vp, vp, vp, vp]
        _lib = L
    return _lib


def _c128(x):
    return np.ascontiguousarray(x, dtype=np.complex128)


def _u32(x):
    return np.ascontiguousarray(x, dtype=np.uint32)


def _u64(x):
    return np.ascontiguousarray(x, dtype=np.uint64)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def enumerate(index):
    a, b = np.zeros(1, np.uint64), np.zeros(1, np.uint64)
    lib().refp_enumerate(index, _p(a), _p(b))
    return int(a[0]), int(b[0])


def pauli_mul(Pa, Pb, xa, xb):
    c, a, b = np.empty(1, np.complex128), np.zeros(1, np.uint64), np.zeros(1, np.uint64)
    lib().refp_pauli_mul(int(Pa), int(Pb), int(xa), int(xb), _p(c), _p(a), _p(b))
    return complex(c[0]), int(a[0]), int(b[0])


def network_unit_at(a, b, idx):
    return int(lib().refp_network_unit_at(int(a), int(b), idx))


class Operator:
    def __init__(self, coeffs, a, b):
        self.coeffs, self.a, self.b = _c128(coeffs), _u64(a), _u64(b)
        self.h = lib().refp_op_create(len(self.coeffs), _p(self.coeffs), _p(self.a), _p(self.b))

    def __del__(self):
        if getattr(self, "h", None):
            lib().refp_op_destroy(self.h); self.h = None


class PsiDeep:
    def __init__(self, num_sites, input_weights, biases, connections, weights, final_weights, log_prefactor):
        a = _c128(input_weights)
        sizes = _u32([len(b) for b in biases])
        conn = _u32([np.asarray(c).shape[0] for c in connections])
        b_cat = _c128(np.concatenate([np.asarray(b).ravel() for b in biases]))
        c_cat = _u32(np.concatenate([np.asarray(c).ravel() for c in connections]))
        w_cat = _c128(np.concatenate([np.asarray(w).ravel() for w in weights]))
        fw = _c128(final_weights)
        lp = complex(log_prefactor)
        self.num_sites, self.N = num_sites, len(a)
        self.h = lib().refp_deep_create(num_sites, len(a), _p(a), len(sizes), _p(sizes), _p(conn), _p(b_cat), _p(c_cat), _p(w_cat),
                                        _p(fw), lp.real, lp.imag)
        self.num_params = int(lib().refp_psi_num_params(self.h))

    def __del__(self):
        if getattr(self, "h", None):
            lib().refp_psi_destroy(self.h); self.h = None


class ExactSummationPaulis:
    kind = ES

    def __init__(self, num_sites):
        self.num_sites = num_sites
        self.h = lib().refp_es_create(num_sites)
        self.num_steps = int(lib().refp_ens_num_steps(ES, self.h))

    def __del__(self):
        if getattr(self, "h", None):
            lib().refp_ens_destroy(ES, self.h); self.h = None


def log_psi_s(psi, a, b):
    out = np.empty(1, np.complex128)
    lib().refp_log_psi_s(psi.h, int(a), int(b), _p(out))
    return complex(out[0])


def psi_O_k(psi, a, b):
    out = np.empty(psi.num_params, np.complex128)
    lib().refp_psi_O_k(psi.h, int(a), int(b), _p(out))
    return out


def log_psi_vector(psi, ens):
    out = np.empty(ens.num_steps, np.complex128)
    lib().refp_log_psi_vector(psi.h, ens.kind, ens.h, _p(out))
    return out


def expectation(op, psi, ens):
    out = np.empty(1, np.complex128)
    lib().refp_expectation(psi.h, op.h, ens.kind, ens.h, _p(out))
    return complex(out[0])


def fluctuation(op, psi, ens):
    out = np.empty(3)
    lib().refp_fluctuation(psi.h, op.h, ens.kind, ens.h, _p(out))
    return float(out[0]), complex(out[1], out[2])


def gradient(op, psi, ens):
    g, e = np.empty(psi.num_params, np.complex128), np.empty(1, np.complex128)
    lib().refp_gradient(psi.h, op.h, ens.kind, ens.h, _p(g), _p(e))
    return g, complex(e[0])


def tdvp_eval(op, psi, ens):
    P = psi.num_params
    t = lib().refp_tdvp_create(P)
    lib().refp_tdvp_eval(t, psi.h, op.h, ens.kind, ens.h)
    S, F, Ok, sc = np.empty((P, P), np.complex128), np.empty(P, np.complex128), np.empty(P, np.complex128), np.empty(4)
    lib().refp_tdvp_get(t, _p(S), _p(F), _p(Ok), _p(sc))
    lib().refp_tdvp_destroy(t)
    return dict(S=S, F=F, O_k=Ok, E=complex(sc[0], sc[1]), E2=float(sc[2]), var_H=float(sc[3]))
