"""TEST INFRASTRUCTURE — ctypes loader for ``oracle/_ref/liboracle_ref.so``.

That library is the UNMODIFIED reference (heikoburau/ANNonGPU) host path compiled from
``/root/reference`` by ``oracle/Makefile`` plus ``oracle/ref_shim/shim.cu``.  It only exists where
it was built (this container) or where the prebuilt ``.so`` travelled to (the GPU box).  Nothing in
the product package may import this module; only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` legs do.

Limits of the true reference: N <= 64 sites, PsiRBM M <= 128, PsiDeep width <= 64,
PsiCNN channels*N <= 128, ExactSummation N <= 31.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "liboracle_ref.so")

RBM, DEEP, CNN, CLFP1, CLFP2, CLANN1, CLANN2 = range(7)
ES, MC = 0, 1

_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        vp, u32, u64, dbl, i32 = C.c_void_p, C.c_uint, C.c_uint64, C.c_double, C.c_int
        for name in ("ref_op_create", "ref_rbm_create", "ref_deep_create", "ref_cnn_create",
                     "ref_classical_create", "ref_es_create", "ref_mc_create", "ref_tdvp_create"):
            getattr(L, name).restype = vp
        L.ref_spins_enumerate.restype = u64
        L.ref_spins_at.restype = dbl
        L.ref_psi_norm.restype = dbl
        L.ref_psi_num_params.restype = u32
        L.ref_ens_num_steps.restype = u32
        L.ref_spins_enumerate.argtypes = [u32]
        L.ref_spins_at.argtypes = [u64, u32]
        L.ref_pauli_apply.argtypes = [u64, u64, u64, vp, vp]
        L.ref_activation.argtypes = [dbl, dbl, u32, vp, vp]
        L.ref_op_create.argtypes = [u32, vp, vp, vp]
        L.ref_rbm_create.argtypes = [u32, u32, vp, dbl, dbl, dbl, dbl]
        L.ref_deep_create.argtypes = [u32, u32, vp, u32, vp, vp, vp, vp, vp, vp, dbl, dbl]
        L.ref_cnn_create.argtypes = [vp, u32, vp, vp, vp, vp, u32, dbl, dbl, dbl]
        L.ref_cnn_init_gradient.argtypes = [vp, u32]
        L.ref_classical_create.argtypes = [u32, u32, u32, vp, vp, u32, vp, dbl, dbl]
        L.ref_psi_destroy.argtypes = [i32, vp]
        L.ref_psi_num_params.argtypes = [i32, vp]
        L.ref_psi_get_params.argtypes = [i32, vp, vp]
        L.ref_psi_set_params.argtypes = [i32, vp, vp]
        L.ref_psi_get_log_prefactor.argtypes = [i32, vp, vp]
        L.ref_psi_set_log_prefactor.argtypes = [i32, vp, dbl, dbl]
        L.ref_es_create.argtypes = [u32]
        L.ref_mc_create.argtypes = [u32, u32, u32, u32]
        L.ref_ens_destroy.argtypes = [i32, vp]
        L.ref_ens_num_steps.argtypes = [i32, vp]
        L.ref_mc_acceptance.argtypes = [vp, vp]
        L.ref_log_psi_s.argtypes = [i32, vp, u64, vp]
        L.ref_psi_O_k.argtypes = [i32, vp, u64, vp]
        for name in ("ref_psi_vector", "ref_log_psi_vector", "ref_log_psi_mean"):
            getattr(L, name).argtypes = [i32, vp, i32, vp, vp]
        L.ref_psi_norm.argtypes = [i32, vp, vp]
        L.ref_psi_O_k_vector.argtypes = [i32, vp, vp, vp]
        L.ref_apply_operator.argtypes = [i32, vp, vp, i32, vp, vp]
        L.ref_expectation.argtypes = [i32, vp, vp, i32, vp, vp]
        L.ref_fluctuation.argtypes = [i32, vp, vp, i32, vp, vp]
        L.ref_gradient.argtypes = [i32, vp, vp, i32, vp, vp, vp]
        L.ref_tdvp_create.argtypes = [u32]
        L.ref_tdvp_destroy.argtypes = [vp]
        L.ref_tdvp_eval.argtypes = [vp, i32, vp, vp, i32, vp]
        L.ref_tdvp_eval_F.argtypes = [vp, i32, vp, vp, i32, vp]
        L.ref_tdvp_get.argtypes = [vp, vp, vp, vp, vp]
        L.ref_tdvp_get_samples.argtypes = [vp, vp, vp]
        L.ref_tdvp_S_dot_vector.argtypes = [vp, vp, i32, vp, vp]
        L.ref_exp_sigma_z.argtypes = [i32, vp, vp, i32, vp, vp]
        L.ref_tdvp_eval_with_psi_ref.argtypes = [vp, i32, vp, vp, i32, vp]
        L.ref_tdvp_eval_with_psi_ref.restype = dbl
        L.ref_hilbert_space_distance.argtypes = [i32, vp, vp, vp, i32, i32, vp, vp, C.c_float]
        L.ref_hilbert_space_distance.restype = dbl
        L.ref_kullback_leibler.argtypes = [i32, vp, i32, vp, i32, vp, i32, dbl, dbl, dbl, dbl, dbl, vp, vp, vp]
        L.ref_kullback_leibler.restype = dbl
        L.ref_set_gpu.argtypes = [i32]
        L.ref_device_synchronize.restype = i32
        _lib = L
    return _lib


def set_gpu(on):
    """Objects created afterwards use the reference's own CUDA path (gpu=true).  Timing baseline only
    (oracle/ref_gpu_bench.py); the parity oracle is the default gpu=false host path."""
    lib().ref_set_gpu(1 if on else 0)


def _c128(x):
    return np.ascontiguousarray(x, dtype=np.complex128)


def _u32(x):
    return np.ascontiguousarray(x, dtype=np.uint32)


def _u64(x):
    return np.ascontiguousarray(x, dtype=np.uint64)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _cout(n):
    return np.empty(n, dtype=np.complex128)


# ---------------------------------------------------------------- primitives

def pauli_apply(a, b, conf):
    coeff = _cout(1)
    out = np.zeros(1, dtype=np.uint64)
    lib().ref_pauli_apply(int(a), int(b), int(conf), _p(coeff), _p(out))
    return complex(coeff[0]), int(out[0])


def spins_enumerate(index):
    return int(lib().ref_spins_enumerate(int(index)))


def spins_at(conf, i):
    return float(lib().ref_spins_at(int(conf), int(i)))


def activation(z, layer):
    lc, th = _cout(1), _cout(1)
    z = complex(z)
    lib().ref_activation(z.real, z.imag, int(layer), _p(lc), _p(th))
    return complex(lc[0]), complex(th[0])


# ---------------------------------------------------------------- objects

class Operator:
    def __init__(self, coeffs, a, b):
        self.coeffs, self.a, self.b = _c128(coeffs), _u64(a), _u64(b)
        self.num_strings = len(self.coeffs)
        self.h = lib().ref_op_create(self.num_strings, _p(self.coeffs), _p(self.a), _p(self.b))

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_op_destroy(self.h)


class Psi:
    kind = None

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_psi_destroy(self.kind, self.h)

    @property
    def num_params(self):
        return int(lib().ref_psi_num_params(self.kind, self.h))

    @property
    def params(self):
        out = _cout(self.num_params)
        lib().ref_psi_get_params(self.kind, self.h, _p(out))
        return out

    @params.setter
    def params(self, value):
        value = _c128(value)
        assert value.size == self.num_params
        lib().ref_psi_set_params(self.kind, self.h, _p(value))

    @property
    def log_prefactor(self):
        out = _cout(1)
        lib().ref_psi_get_log_prefactor(self.kind, self.h, _p(out))
        return complex(out[0])

    @log_prefactor.setter
    def log_prefactor(self, value):
        value = complex(value)
        lib().ref_psi_set_log_prefactor(self.kind, self.h, value.real, value.imag)


class PsiRBM(Psi):
    kind = RBM

    def __init__(self, W, final_weight, log_prefactor):
        W = _c128(W)
        self.N, self.M = W.shape
        self.num_sites = self.N
        fw, lp = complex(final_weight), complex(log_prefactor)
        self.h = lib().ref_rbm_create(self.N, self.M, _p(W), fw.real, fw.imag, lp.real, lp.imag)


class PsiDeep(Psi):
    kind = DEEP

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
        self.h = lib().ref_deep_create(num_sites, len(a), _p(a), len(sizes), _p(sizes), _p(conn), _p(b_cat),
                                       _p(c_cat), _p(w_cat), _p(fw), lp.real, lp.imag)


class PsiCNN(Psi):
    kind = CNN

    def __init__(self, extent, num_channels_list, connectivity_list, symmetry_classes, params, final_factor, log_prefactor):
        ext = _u32(extent)
        assert ext.size == 3
        nc, conn, sym, p = _u32(num_channels_list), _u32(connectivity_list), _u32(symmetry_classes), _c128(params)
        lp = complex(log_prefactor)
        self.num_sites = self.N = int(np.prod(ext))
        self.h = lib().ref_cnn_create(_p(ext), len(nc), _p(nc), _p(conn), _p(sym), _p(p), p.size,
                                      float(final_factor), lp.real, lp.imag)

    def init_gradient(self, num_steps):
        lib().ref_cnn_init_gradient(self.h, int(num_steps))


class PsiClassical(Psi):
    def __init__(self, num_sites, order, H_local, params, psi_ref, log_prefactor):
        self.kind = {(1, False): CLFP1, (2, False): CLFP2, (1, True): CLANN1, (2, True): CLANN2}[(order, psi_ref is not None)]
        self.num_sites = self.N = num_sites
        self._ops = list(H_local)
        handles = (C.c_void_p * len(self._ops))(*[op.h for op in self._ops])
        p = _c128(params)
        lp = complex(log_prefactor)
        self.h = lib().ref_classical_create(num_sites, order, len(self._ops), handles, _p(p), p.size,
                                            psi_ref.h if psi_ref is not None else None, lp.real, lp.imag)


class ExactSummation:
    kind = ES

    def __init__(self, num_sites):
        self.h = lib().ref_es_create(num_sites)
        self.num_steps = 1 << num_sites

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_ens_destroy(self.kind, self.h)


class MonteCarlo:
    """CPU reference Monte-Carlo runs Markov chain 0 only (MonteCarlo.hpp:67-71): use num_chains=1."""
    kind = MC

    def __init__(self, num_samples, num_sweeps, num_thermalization_sweeps, num_markov_chains=1):
        self.h = lib().ref_mc_create(num_samples, num_sweeps, num_thermalization_sweeps, num_markov_chains)
        self.num_steps = num_samples

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_ens_destroy(self.kind, self.h)

    @property
    def acceptance(self):
        out = np.zeros(2, dtype=np.uint32)
        lib().ref_mc_acceptance(self.h, _p(out))
        return int(out[0]), int(out[1])


# ---------------------------------------------------------------- functions

def log_psi_s(psi, conf):
    out = _cout(1)
    lib().ref_log_psi_s(psi.kind, psi.h, int(conf), _p(out))
    return complex(out[0])


def psi_O_k(psi, conf):
    out = _cout(psi.num_params)
    lib().ref_psi_O_k(psi.kind, psi.h, int(conf), _p(out))
    return out


def psi_vector(psi, ens):
    out = _cout(ens.num_steps)
    lib().ref_psi_vector(psi.kind, psi.h, ens.kind, ens.h, _p(out))
    return out


def log_psi_vector(psi, ens):
    out = _cout(ens.num_steps)
    lib().ref_log_psi_vector(psi.kind, psi.h, ens.kind, ens.h, _p(out))
    return out


def log_psi(psi, ens):
    out = _cout(1)
    lib().ref_log_psi_mean(psi.kind, psi.h, ens.kind, ens.h, _p(out))
    return complex(out[0])


def psi_norm(psi, es):
    return float(lib().ref_psi_norm(psi.kind, psi.h, es.h))


def psi_O_k_vector(psi, es):
    out = _cout(psi.num_params)
    lib().ref_psi_O_k_vector(psi.kind, psi.h, es.h, _p(out))
    return out


def apply_operator(psi, op, ens):
    out = _cout(ens.num_steps)
    lib().ref_apply_operator(psi.kind, psi.h, op.h, ens.kind, ens.h, _p(out))
    return out


def expectation(op, psi, ens):
    out = _cout(1)
    lib().ref_expectation(psi.kind, psi.h, op.h, ens.kind, ens.h, _p(out))
    return complex(out[0])


def fluctuation(op, psi, ens):
    out = np.empty(3)
    lib().ref_fluctuation(psi.kind, psi.h, op.h, ens.kind, ens.h, _p(out))
    return float(out[0]), complex(out[1], out[2])


def exp_sigma_z(op, psi, ens):
    out = np.empty(2)
    lib().ref_exp_sigma_z(psi.kind, psi.h, op.h, ens.kind, ens.h, _p(out))
    return complex(out[0], out[1])


def gradient(op, psi, ens):
    g, e = _cout(psi.num_params), _cout(1)
    lib().ref_gradient(psi.kind, psi.h, op.h, ens.kind, ens.h, _p(g), _p(e))
    return g, complex(e[0])


def hilbert_space_distance(psi, psi_prime, op, is_unitary, ens):
    """HilbertSpaceDistance::distance; (PsiDeep, PsiDeep) or (PsiCNN, PsiCNN) only."""
    assert psi.kind == psi_prime.kind and psi.kind in (DEEP, CNN)
    return float(lib().ref_hilbert_space_distance(psi.kind, psi.h, psi_prime.h, op.h, int(bool(is_unitary)), ens.kind, ens.h, None, 0.0))


def hilbert_space_distance_gradient(psi, psi_prime, op, is_unitary, ens, nu):
    """HilbertSpaceDistance::gradient -> (gradient[P'], distance)."""
    assert psi.kind == psi_prime.kind and psi.kind in (DEEP, CNN)
    g = _cout(psi_prime.num_params)
    d = lib().ref_hilbert_space_distance(psi.kind, psi.h, psi_prime.h, op.h, int(bool(is_unitary)), ens.kind, ens.h, _p(g), float(nu))
    return g, float(d)


class KullbackLeibler:
    """KullbackLeibler(num_params, gpu) (pyANNonGPU/main.cpp.template:445-461): psi a PsiClassical kind, psi_prime a PsiDeep
    or PsiCNN; samples are drawn from psi_prime.  last_mean_deviation is carried from call to call like upstream."""

    def __init__(self, num_params):
        self.num_params, self.log_psi_scale = int(num_params), 1.0
        self.last_mean_deviation, self.mean_deviation, self.total_weight = 0j, 0j, 0.0

    def _run(self, mode, psi, psi_prime, ens, nu, threshold):
        assert psi_prime.kind in (DEEP, CNN) and psi.kind in (CLFP1, CLFP2, CLANN1, CLANN2)
        g, noise, extra = _cout(self.num_params), np.zeros(self.num_params), np.zeros(3)
        v = lib().ref_kullback_leibler(psi.kind, psi.h, psi_prime.kind, psi_prime.h, ens.kind, ens.h, mode, float(nu), float(threshold),
                                       float(self.log_psi_scale), self.last_mean_deviation.real, self.last_mean_deviation.imag,
                                       _p(g), _p(noise), _p(extra))
        self.total_weight, self.mean_deviation = float(extra[0]), complex(extra[1], extra[2])
        self.last_mean_deviation = self.mean_deviation
        return float(v), g, noise

    def __call__(self, psi, psi_prime, ens, threshold):
        return self._run(0, psi, psi_prime, ens, 0.0, threshold)[0]

    def gradient(self, psi, psi_prime, ens, nu, threshold):
        v, g, _ = self._run(1, psi, psi_prime, ens, nu, threshold)
        return g, v

    def gradient_with_noise(self, psi, psi_prime, ens, nu, threshold):
        v, g, noise = self._run(2, psi, psi_prime, ens, nu, threshold)
        return g, noise, v


class TDVP:
    def __init__(self, num_params):
        self.P = int(num_params)
        self.h = lib().ref_tdvp_create(self.P)

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_tdvp_destroy(self.h)

    def eval(self, op, psi, ens):
        self._ns = ens.num_steps
        lib().ref_tdvp_eval(self.h, psi.kind, psi.h, op.h, ens.kind, ens.h)

    def eval_with_psi_ref(self, op, psi, ens):
        """TDVP::eval(..., true_t) (PsiClassical kinds only); returns total_weight of this call."""
        self._ns = ens.num_steps
        self.total_weight = float(lib().ref_tdvp_eval_with_psi_ref(self.h, psi.kind, psi.h, op.h, ens.kind, ens.h))
        return self.total_weight

    def eval_F(self, op, psi, ens):
        self._ns = ens.num_steps
        lib().ref_tdvp_eval_F(self.h, psi.kind, psi.h, op.h, ens.kind, ens.h)

    def _get(self, want_S):
        S = _cout(self.P * self.P) if want_S else None
        F, Ok, scal = _cout(self.P), _cout(self.P), np.empty(4)
        lib().ref_tdvp_get(self.h, _p(S) if want_S else None, _p(F), _p(Ok), _p(scal))
        return S, F, Ok, scal

    @property
    def S_matrix(self):
        return self._get(True)[0].reshape(self.P, self.P)

    @property
    def F_vector(self):
        return self._get(False)[1]

    @property
    def O_k_vector(self):
        return self._get(False)[2]

    @property
    def E_local(self):
        s = self._get(False)[3]
        return complex(s[0], s[1])

    @property
    def var_H(self):
        return float(self._get(False)[3][3])

    @property
    def O_k_samples(self):
        out = _cout(self._ns * self.P)
        lib().ref_tdvp_get_samples(self.h, _p(out), None)
        return out.reshape(self._ns, self.P)

    @property
    def weight_samples(self):
        out = np.empty(self._ns)
        lib().ref_tdvp_get_samples(self.h, None, _p(out))
        return out

    def S_dot_vector(self, vec, ens):
        vec, out = _c128(vec), _cout(self.P)
        lib().ref_tdvp_S_dot_vector(self.h, _p(vec), ens.kind, ens.h, _p(out))
        return out
