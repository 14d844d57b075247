"""Host-side mirror of the reference's binding surface for the VMC hot path, over the C ABI.

Same names, argument order and return types as the pybind11 module ``_pyANNonGPU``
(``pyANNonGPU/main.cpp.template:65-541``) plus the monkey-patched helpers of ``pyANNonGPU/Psi*.py``:

    PsiRBM, PsiDeep, PsiCNN, PsiClassicalFP_1/_2, PsiClassicalANN_1/_2, PsiFullyPolarized,
    Operator, Spins, MonteCarloSpins, ExactSummationSpins, ExpectationValue, TDVP,
    log_psi_s, psi_O_k, psi_O_k_vector, log_psi, psi_vector, log_psi_vector, apply_operator,
    activation_function, setDevice, start_profiling, stop_profiling

Differences, all additive or forced by the B200-only scope:
  * ``gpu`` arguments are accepted for signature compatibility but must be truthy — there is no CPU path;
  * ``Operator`` also accepts raw ``(coefficients, a_masks, b_masks)`` / a ``factories.PauliSum`` (the reference
    needs a ``QuantumExpression.PauliExpression``; any object with ``.coeffs/.a/.b/.num_sites`` or iterable of
    ``(pauli_string_with_.a_.b, coefficient)`` terms is duck-typed);
  * ``MonteCarloSpins`` takes an optional ``seed``; ``Spins`` may hold more than 64 sites;
  * ``TDVP.solve`` / ``TDVP.solve_cg`` are new (the reference has no solver).
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import call, lib, ALLREDUCE_FN
from .factories import PauliSum, masks_to_words, words_for

__all__ = [
    "PsiRBM", "PsiDeep", "PsiCNN", "PsiClassicalFP_1", "PsiClassicalFP_2", "PsiClassicalANN_1", "PsiClassicalANN_2",
    "PsiFullyPolarized", "Operator", "Spins", "MonteCarloSpins", "ExactSummationSpins", "MonteCarloPaulis", "ExactSummationPaulis",
    "paulis_to_units", "units_to_paulis", "ExpectationValue", "TDVP",
    "HilbertSpaceDistance", "KullbackLeibler",
    "log_psi_s", "psi_O_k", "psi_O_k_vector", "log_psi", "psi_vector", "log_psi_vector", "apply_operator",
    "local_energies", "activation_function", "pauli_apply", "setDevice", "start_profiling", "stop_profiling",
    "synchronize", "launch_count", "set_stream", "measure_fp64_tflops", "AngpuError",
    "comm_unique_id", "comm_init", "comm_destroy", "comm_rank", "hpd_solve",
]
AngpuError = _lib.AngpuError


def _c128(x):
    return np.ascontiguousarray(x, dtype=np.complex128)


def _u32(x):
    return np.ascontiguousarray(x, dtype=np.uint32)


def _u64(x):
    return np.ascontiguousarray(x, dtype=np.uint64)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _pair(z):
    z = complex(z)
    return np.array([z.real, z.imag], dtype=np.float64)


def _require_gpu(gpu):
    if not gpu:
        raise ValueError("annongpu_b200 is GPU-only (sm_100a); the reference's gpu=False host path is not provided")


def setDevice(device):
    """source/ANNonGPU.cu:7-9"""
    call("angpu_init", int(device))


def set_stream(cuda_stream):
    """Run the library on the given CUDA stream handle; None restores the library's own stream.  Handle 0 (the legacy
    default stream, what torch.cuda.default_stream().cuda_stream returns) is passed as cudaStreamLegacy."""
    if cuda_stream is None:
        call("angpu_set_stream", None)
    else:
        call("angpu_set_stream", C.c_void_p(int(cuda_stream) or 1))


def synchronize():
    call("angpu_synchronize")


def start_profiling():
    call("angpu_profiler_start")


def stop_profiling():
    call("angpu_profiler_stop")


def measure_fp64_tflops():
    out = C.c_double()
    call("angpu_measure_fp64_tflops", C.byref(out))
    return float(out.value)


def launch_count(reset=False):
    return int(lib.angpu_launch_count(1 if reset else 0))


def activation_function(z, layer=0):
    """my_logcosh(z, layer) (pyANNonGPU/main.cpp.template:534-536), evaluated on the device."""
    lc, th = np.empty(2), np.empty(2)
    call("angpu_activation", _p(_pair(z)), int(layer), _p(lc), _p(th))
    return complex(lc[0], lc[1])


def activation_derivative(z, layer=0):
    """my_tanh(z, layer), evaluated on the device."""
    lc, th = np.empty(2), np.empty(2)
    call("angpu_activation", _p(_pair(z)), int(layer), _p(lc), _p(th))
    return complex(th[0], th[1])


# -------------------------------------------------------------------------------------------- Spins

class Spins:
    """Spins(configuration, num_spins) (pyANNonGPU/main.cpp.template:340-344; include/basis/Spins.h)."""

    def __init__(self, configuration, num_spins=64):
        self.num_spins = int(num_spins)
        self.configuration = int(configuration) & ((1 << self.num_spins) - 1)

    @staticmethod
    def enumerate(index, num_spins=64):
        words = words_for(num_spins)
        out = np.zeros(words, dtype=np.uint64)
        call("angpu_spins_enumerate", int(index), words, _p(out))
        return Spins(sum(int(x) << (64 * w) for w, x in enumerate(out)), num_spins)

    def words(self, words=None):
        words = words or words_for(self.num_spins)
        return np.array([(self.configuration >> (64 * w)) & 0xFFFFFFFFFFFFFFFF for w in range(words)], dtype=np.uint64)

    def array(self, num_spins=None):
        n = num_spins or self.num_spins
        return np.array([1.0 if (self.configuration >> i) & 1 else -1.0 for i in range(n)])

    def roll(self, shift, N):
        """Spins::roll (include/basis/Spins.h:349-354)."""
        c = self.configuration & ((1 << N) - 1)
        return Spins(((c << shift) | (c >> (N - shift))) & ((1 << N) - 1), N)

    def flip(self, position):
        return Spins(self.configuration ^ (1 << position), self.num_spins)

    def __eq__(self, other):
        return isinstance(other, Spins) and self.configuration == other.configuration

    def __repr__(self):
        return f"Spins({self.configuration:#x}, {self.num_spins})"


def _conf_words(conf, words):
    if isinstance(conf, Spins):
        return conf.words(words)
    if isinstance(conf, (int, np.integer)):
        return Spins(int(conf), 64 * words).words(words)
    arr = _u64(conf).ravel()
    assert arr.size == words, f"configuration needs {words} uint64 words"
    return arr


def pauli_apply(a, b, conf, num_sites=64):
    """PauliString(a, b).apply(Spins) evaluated on the device: returns (coefficient, Spins)."""
    words = words_for(num_sites)
    am, bm = masks_to_words([a], words)[0], masks_to_words([b], words)[0]
    c = _conf_words(conf, words)
    coeff, out = np.empty(2), np.zeros(words, dtype=np.uint64)
    call("angpu_pauli_apply", _p(am), _p(bm), _p(c), words, _p(coeff), _p(out))
    return complex(coeff[0], coeff[1]), Spins(sum(int(x) << (64 * w) for w, x in enumerate(out)), num_sites)


# -------------------------------------------------------------------------------------------- Operator

class Operator:
    """Operator(expr, gpu) (pyANNonGPU/main.cpp.template:321-333, source/operator/Operator.cpp:18-39)."""

    def __init__(self, expr, gpu=True, num_sites=None, _raw=None):
        _require_gpu(gpu)
        if _raw is not None:
            coeffs, a, b, num_sites = _raw
        elif isinstance(expr, PauliSum) or all(hasattr(expr, k) for k in ("coeffs", "a", "b")):
            coeffs, a, b = list(expr.coeffs), list(expr.a), list(expr.b)
            num_sites = num_sites or getattr(expr, "num_sites", None)
        else:  # duck-typed PauliExpression: iterable of (pauli_string, coefficient) with .a / .b masks
            coeffs, a, b = [], [], []
            for string, coeff in expr:
                coeffs.append(complex(coeff)); a.append(int(string.a)); b.append(int(string.b))
        if num_sites is None:
            top = max([int(x | y).bit_length() for x, y in zip(a, b)] + [1])
            num_sites = max(top, 1)
        self.num_sites = int(num_sites)
        self.words = words_for(self.num_sites)
        self.coefficients = _c128(coeffs)
        self.a_masks, self.b_masks = [int(x) for x in a], [int(x) for x in b]
        self._a, self._b = masks_to_words(self.a_masks, self.words), masks_to_words(self.b_masks, self.words)
        self._h = C.c_void_p()
        call("angpu_operator_create", len(self.coefficients), _p(self.coefficients), _p(self._a), _p(self._b),
             self.words, C.byref(self._h))
        self.gpu = True

    @classmethod
    def from_arrays(cls, coefficients, a_masks, b_masks, num_sites, gpu=True):
        return cls(None, gpu, _raw=(coefficients, a_masks, b_masks, num_sites))

    def with_words(self, words):
        """Same operator with masks padded to `words` 64-bit words (to match a wider wavefunction)."""
        if words == self.words:
            return self
        if words < self.words and any(int(m) >> (64 * words) for m in list(self.a_masks) + list(self.b_masks)):
            raise RuntimeError(f"operator acts on sites beyond {64 * words}: it does not fit a wavefunction of {words} mask word(s)")
        return Operator(None, True, _raw=(self.coefficients, self.a_masks, self.b_masks, 64 * words))

    @property
    def num_strings(self):
        return len(self.coefficients)

    def __del__(self):
        if getattr(self, "_h", None) and self._h:
            lib.angpu_operator_destroy(self._h)
            self._h = None


# -------------------------------------------------------------------------------------------- wavefunctions

class _Psi:
    _h = None
    gpu = True

    def __del__(self):
        if getattr(self, "_h", None):
            lib.angpu_psi_destroy(self._h)
            self._h = None

    @classmethod
    def _wrap(cls, handle, template):
        obj = object.__new__(type(template))
        obj.__dict__.update({k: v for k, v in template.__dict__.items() if k != "_h"})
        obj._h = handle
        return obj

    def copy(self):
        h = C.c_void_p()
        call("angpu_psi_copy", self._h, C.byref(h))
        return self._wrap(h, self)

    def __pos__(self):
        return self.copy()

    @property
    def pauli_sites(self):
        """!= 0: a network on the Pauli-string basis (PsiDeep with N = 3 num_sites input units), to be used with MonteCarloPaulis /
        ExactSummationPaulis; its configurations are units masks (paulis_to_units)."""
        n, N = getattr(self, "num_sites", 0), getattr(self, "N", 0)
        return n if (isinstance(self, PsiDeep) and n and N == 3 * n) else 0

    @property
    def words(self):
        """uint64 words of one configuration as it crosses the boundary."""
        return words_for(3 * self.num_sites) if self.pauli_sites else words_for(self.num_sites)

    @property
    def num_params(self):
        n = C.c_uint()
        call("angpu_psi_num_params", self._h, C.byref(n))
        return int(n.value)

    @property
    def params(self):
        out = np.empty(self.num_params, dtype=np.complex128)
        call("angpu_psi_get_params", self._h, _p(out))
        return out

    @params.setter
    def params(self, value):
        value = _c128(value).ravel()
        if value.size != self.num_params:
            raise ValueError(f"expected {self.num_params} parameters, got {value.size}")
        call("angpu_psi_set_params", self._h, _p(value))

    @property
    def log_prefactor(self):
        out = np.empty(2)
        call("angpu_psi_get_log_prefactor", self._h, _p(out))
        return complex(out[0], out[1])

    @log_prefactor.setter
    def log_prefactor(self, value):
        call("angpu_psi_set_log_prefactor", self._h, _p(_pair(value)))

    # pyANNonGPU/PsiRBM.py:39-66 (same helpers are patched onto every Psi class)
    def _vector(self, exact_summation):
        return psi_vector(self, exact_summation)

    @property
    def vector(self):
        return self._vector

    def norm(self, exact_summation):
        out = C.c_double()
        call("angpu_psi_norm", self._h, exact_summation._h, C.byref(out))
        return float(out.value)

    def normalize(self, exact_summation):
        self.log_prefactor -= np.log(self.norm(exact_summation))

    def calibrate(self, ensemble):
        if type(ensemble).__name__.startswith("ExactSummation"):
            self.normalize(ensemble)
            self.log_prefactor -= 1j * log_psi(self, ensemble).imag
        else:
            self.log_prefactor = 0
            self.log_prefactor -= log_psi(self, ensemble)


class PsiRBM(_Psi):
    """PsiRBM(W, final_weight, log_prefactor, gpu) (pyANNonGPU/main.cpp.template:120-154)."""

    def __init__(self, W, final_weight, log_prefactor=0.0, gpu=True):
        _require_gpu(gpu)
        W = _c128(W)
        assert W.ndim == 2
        self.N, self.M = (int(x) for x in W.shape)
        self.num_sites = self.N
        self.final_weight = complex(final_weight)
        self.symmetric = False
        self._h = C.c_void_p()
        call("angpu_rbm_create", self.N, self.M, _p(W), _p(_pair(final_weight)), _p(_pair(log_prefactor)), C.byref(self._h))

    @property
    def W(self):
        return self.params.reshape(self.N, self.M)

    def to_json(self):
        """pyANNonGPU/PsiRBM.py:7-19 (same keys and array encoding)."""
        from .json_numpy import plain
        fw = self.final_weight
        return plain(dict(type="PsiRBM", W=self.W, final_weight=fw.real if fw.imag == 0.0 else fw,
                          log_prefactor_re=self.log_prefactor.real, log_prefactor_im=self.log_prefactor.imag))

    @staticmethod
    def from_json(json_obj, gpu=True):
        from .json_numpy import restore
        obj = restore(json_obj)
        return PsiRBM(np.asarray(obj["W"]), obj["final_weight"], obj["log_prefactor_re"] + 1j * obj["log_prefactor_im"], gpu)


class PsiDeep(_Psi):
    """PsiDeep(num_sites, input_weights, biases, lhs_connections, lhs_weights, final_weights, log_prefactor, gpu)
    (pyANNonGPU/main.cpp.template:71-114)."""

    def __init__(self, num_sites, input_weights, biases, connections, weights, final_weights, log_prefactor=0.0, gpu=True):
        _require_gpu(gpu)
        a = _c128(input_weights).ravel()
        self.num_sites, self.N = int(num_sites), int(a.size)
        self._sizes = _u32([len(b) for b in biases])
        self._conn = _u32([np.asarray(c).shape[0] for c in connections])
        b_cat = _c128(np.concatenate([np.asarray(b).ravel() for b in biases]))
        c_cat = _u32(np.concatenate([np.asarray(c).ravel() for c in connections]))
        w_cat = _c128(np.concatenate([np.asarray(w).ravel() for w in weights]))
        self.connections = [np.array(c, dtype=np.uint32) for c in connections]
        self._final_weights = _c128(final_weights).ravel()
        self.symmetric = False
        self.N_i = self.N_j = 0
        self._h = C.c_void_p()
        call("angpu_deep_create", self.num_sites, self.N, _p(a), len(self._sizes), _p(self._sizes), _p(self._conn),
             _p(b_cat), _p(c_cat), _p(w_cat), _p(self._final_weights), _p(_pair(log_prefactor)), C.byref(self._h))

    def _split(self):
        p = self.params
        a, off = p[:self.N], self.N
        b, W = [], []
        for size, conn in zip(self._sizes, self._conn):
            b.append(p[off:off + size]); off += size
            W.append(p[off:off + size * conn].reshape(conn, size)); off += size * conn
        return a, b, W

    @property
    def a(self):
        return self._split()[0]

    input_weights = a

    @property
    def b(self):
        return self._split()[1]

    @property
    def W(self):
        return self._split()[2]

    @property
    def final_weights(self):
        return self._final_weights.copy()


def _deep_to_json(self):
    """pyANNonGPU/PsiDeep.py:7-22 (same keys and array encoding)."""
    from .json_numpy import plain
    return plain(dict(type="PsiDeep", num_sites=self.num_sites, a=self.a, b=list(self.b), connections=list(self.connections),
                      W=list(self.W), final_weights=self.final_weights,
                      log_prefactor_re=self.log_prefactor.real, log_prefactor_im=self.log_prefactor.imag))


def _deep_from_json(json_obj, gpu=True):
    from .json_numpy import restore
    obj = restore(json_obj)
    return PsiDeep(obj["num_sites"], obj["a"], obj["b"], obj["connections"], obj["W"], obj["final_weights"],
                   obj["log_prefactor_re"] + 1j * obj["log_prefactor_im"], gpu)


PsiDeep.to_json = _deep_to_json
PsiDeep.from_json = staticmethod(_deep_from_json)


class PsiCNN(_Psi):
    """PsiCNN(extent, num_channels_list, connectivity_list, symmetry_classes, params, final_factor, log_prefactor, gpu)
    (pyANNonGPU/main.cpp.template:160-204)."""

    def __init__(self, extent, num_channels_list, connectivity_list, symmetry_classes, params, final_factor, log_prefactor=0.0, gpu=True):
        _require_gpu(gpu)
        extent = list(int(x) for x in extent)
        conn = np.atleast_2d(np.asarray(connectivity_list, dtype=np.uint32))
        while len(extent) < 3:   # the bound type is PsiCNN_t<3>; lower-dimensional lattices get leading 1s
            extent = [1] + extent
            conn = np.concatenate([np.ones((conn.shape[0], 1), dtype=np.uint32), conn], axis=1)
        self.extent = extent
        self.dim = 3
        self.num_channels_list = _u32(num_channels_list)
        self.connectivity_list = _u32(conn)
        self.symmetry_classes = _u32(symmetry_classes)
        self.N = self.num_sites = int(np.prod(extent))
        self.final_factor = float(final_factor)
        self.num_symmetry_classes = len(set(int(s) for s in self.symmetry_classes))
        p = _c128(params).ravel()
        self._h = C.c_void_p()
        ext = _u32(extent)
        call("angpu_cnn_create", _p(ext), len(self.num_channels_list), _p(self.num_channels_list), _p(self.connectivity_list),
             _p(self.symmetry_classes), _p(p), p.size, self.final_factor, _p(_pair(log_prefactor)), C.byref(self._h))

    def init_gradient(self, num_steps):
        """Kept for API parity (source/quantum_state/PsiCNN.cpp:75-78): the per-sample angle scratch lives in shared
        memory here, so there is nothing to size."""

    def channel_link(self, layer, prev_channel, channel):
        """Weights of one channel link, shaped (num_symmetry_classes, volume) (cf. test/test_Psi.py:65-69)."""
        off = 0
        for l in range(layer + 1):
            nch = int(self.num_channels_list[l]); prev = int(self.num_channels_list[l - 1]) if l > 0 else 1
            vol = int(np.prod(self.connectivity_list[l]))
            if l == layer:
                off += (prev_channel * nch + channel) * self.num_symmetry_classes * vol
                return self.params[off:off + self.num_symmetry_classes * vol].reshape(self.num_symmetry_classes, vol)
            off += nch * prev * self.num_symmetry_classes * vol


def _cnn_to_json(self):
    """pyANNonGPU/PsiCNN.py:7-23 (same keys and array encoding)."""
    from .json_numpy import plain
    return plain(dict(type="PsiCNN", extent=np.asarray(self.extent), num_channels_list=np.asarray(self.num_channels_list),
                      connectivity_list=np.asarray(self.connectivity_list), symmetry_classes=np.asarray(self.symmetry_classes),
                      params=self.params, final_factor=self.final_factor,
                      log_prefactor_re=self.log_prefactor.real, log_prefactor_im=self.log_prefactor.imag))


def _cnn_from_json(json_obj, gpu=True):
    from .json_numpy import restore
    obj = restore(json_obj)
    return PsiCNN(obj["extent"], obj["num_channels_list"], obj["connectivity_list"], obj["symmetry_classes"], obj["params"],
                  obj["final_factor"], obj["log_prefactor_re"] + 1j * obj["log_prefactor_im"], gpu)


PsiCNN.to_json = _cnn_to_json
PsiCNN.from_json = staticmethod(_cnn_from_json)


class PsiFullyPolarized:
    """PsiFullyPolarized(num_sites, log_prefactor) (pyANNonGPU/main.cpp.template:284-291): log psi(s) = 0 for every s,
    no parameters (include/quantum_state/PsiFullyPolarized.hpp:41-49 ignores log_prefactor).  As a sampling
    distribution (TDVP.eval_with_psi_ref of a PsiClassicalFP) it is the uniform one; on the device it is a PsiClassical
    with no local operators."""

    def __init__(self, num_sites, log_prefactor=0.0):
        self.num_sites = self.N = int(num_sites)
        self.log_prefactor, self.num_params, self.gpu = complex(log_prefactor), 0, True
        self._handle = None

    @property
    def _h(self):
        if self._handle is None:
            self._handle = C.c_void_p()
            call("angpu_classical_create", self.num_sites, 1, 0, None, None, 0, None, _p(_pair(0.0)), C.byref(self._handle))
        return self._handle

    def __del__(self):
        if getattr(self, "_handle", None):
            lib.angpu_psi_destroy(self._handle)
            self._handle = None

    def to_json(self):
        """pyANNonGPU/PsiFullyPolarized.py:4-10."""
        return dict(type="PsiFullyPolarized", num_sites=self.num_sites, log_prefactor_re=self.log_prefactor.real,
                    log_prefactor_im=self.log_prefactor.imag)

    @staticmethod
    def from_json(obj):
        return PsiFullyPolarized(obj["num_sites"], obj["log_prefactor_re"] + 1j * obj["log_prefactor_im"])


class _PsiClassical(_Psi):
    _order = 1
    _ann = False

    def __init__(self, num_sites, H_local, params, psi_ref, log_prefactor=0.0, gpu=True):
        _require_gpu(gpu)
        self.num_sites = self.N = int(num_sites)
        self.H_local = list(H_local)
        self.psi_ref = psi_ref
        p = _c128(params).ravel()
        handles = (C.c_void_p * max(1, len(self.H_local)))(*[op._h for op in self.H_local])
        ref = psi_ref._h if isinstance(psi_ref, PsiCNN) else None
        if self._ann and ref is None:
            raise ValueError("PsiClassicalANN needs a PsiCNN reference state")
        self._h = C.c_void_p()
        call("angpu_classical_create", self.num_sites, self._order, len(self.H_local), handles, _p(p), p.size, ref,
             _p(_pair(log_prefactor)), C.byref(self._h))

    @property
    def order(self):
        return self._order

    def update_psi_ref_kernel(self):
        pass

    def to_json(self, ansatz=None):
        """pyANNonGPU/PsiClassical.py:7-25 (same keys).  `ansatz`: the local operators as PauliSum expressions; defaults
        to the expressions the H_local operators were built from."""
        from .json_numpy import plain
        if ansatz is None:
            ansatz = [PauliSum(op.num_sites, list(op.coefficients), list(op.a_masks), list(op.b_masks)) for op in self.H_local]
        ref = self.psi_ref.to_json() if hasattr(self.psi_ref, "to_json") else self.psi_ref
        return plain(dict(type="PsiClassical", num_sites=self.num_sites, order=self.order, ansatz=[a.to_json() for a in ansatz],
                          params=self.params[:len(self.H_local)],
                          psi_ref=ref, log_prefactor_re=self.log_prefactor.real, log_prefactor_im=self.log_prefactor.imag))

    @staticmethod
    def from_json(json_obj, gpu=True):
        """pyANNonGPU/PsiClassical.py:28-67: PsiClassical{FP,ANN}_{1,2} by `order` and the type of psi_ref."""
        from .json_numpy import restore
        obj = restore(json_obj)
        kind = obj["psi_ref"]["type"]
        if kind == "PsiFullyPolarized":
            psi_ref = PsiFullyPolarized.from_json(obj["psi_ref"])
        elif kind == "PsiCNN":
            psi_ref = PsiCNN.from_json(obj["psi_ref"], gpu)
        else:
            raise ValueError(f"PsiClassical.from_json: reference state of type {kind} is not supported (PsiFullyPolarized, PsiCNN)")
        H_local = [Operator(PauliSum.from_json(h), gpu, num_sites=obj["num_sites"]) for h in obj["ansatz"]]
        lp = obj["log_prefactor_re"] + 1j * obj["log_prefactor_im"]
        cls = {(1, True): PsiClassicalFP_1, (2, True): PsiClassicalFP_2, (1, False): PsiClassicalANN_1,
               (2, False): PsiClassicalANN_2}[(int(obj["order"]), kind == "PsiFullyPolarized")]
        return cls(obj["num_sites"], H_local, np.asarray(obj["params"]), psi_ref, lp, gpu)


class PsiClassicalFP_1(_PsiClassical):
    _order, _ann = 1, False


class PsiClassicalFP_2(_PsiClassical):
    _order, _ann = 2, False


class PsiClassicalANN_1(_PsiClassical):
    _order, _ann = 1, True


class PsiClassicalANN_2(_PsiClassical):
    _order, _ann = 2, True


# -------------------------------------------------------------------------------------------- ensembles

class _Ensemble:
    _h = None
    gpu = True

    def __del__(self):
        if getattr(self, "_h", None):
            lib.angpu_ensemble_destroy(self._h)
            self._h = None

    @property
    def num_steps(self):
        n = C.c_ulonglong()
        call("angpu_ensemble_num_steps", self._h, C.byref(n))
        return int(n.value)

    @property
    def local_steps(self):
        n = C.c_ulonglong()
        call("angpu_ensemble_local_steps", self._h, C.byref(n))
        return int(n.value)

    def set_shard(self, rank, world):
        """Multi-GPU: own chains / basis indices [rank*n/world, (rank+1)*n/world)."""
        call("angpu_ensemble_set_shard", self._h, int(rank), int(world))
        return self

    def sample(self, psi):
        """(configurations [local_steps, words] uint64, log_psi [local_steps]) of one run of the sampler."""
        n = self.local_steps
        confs = np.zeros((max(n, 1), psi.words), dtype=np.uint64)
        lp = np.zeros(max(n, 1), dtype=np.complex128)
        call("angpu_ensemble_sample", self._h, psi._h, _p(confs), _p(lp))
        return confs[:n], lp[:n]


class ExactSummationSpins(_Ensemble):
    """ExactSummationSpins(num_sites, gpu) (pyANNonGPU/main.cpp.template:395-399)."""

    def __init__(self, num_sites, gpu=True):
        _require_gpu(gpu)
        self.num_sites = int(num_sites)
        self._h = C.c_void_p()
        call("angpu_es_create", self.num_sites, C.byref(self._h))


class ExactSummationPaulis(_Ensemble):
    """ExactSummationPaulis(num_sites, gpu) (pyANNonGPU/main.cpp.template:400-405): all 4^num_sites Pauli strings, for a PsiDeep with
    N = 3 num_sites input units (the density-matrix basis)."""

    def __init__(self, num_sites, gpu=True):
        _require_gpu(gpu)
        self.num_sites = int(num_sites)
        self._h = C.c_void_p()
        call("angpu_es_paulis_create", self.num_sites, C.byref(self._h))


class MonteCarloSpins(_Ensemble):
    """MonteCarloSpins(num_samples, num_sweeps, num_thermalization_sweeps, num_markov_chains, gpu)
    (pyANNonGPU/main.cpp.template:358-366) + seed; or MonteCarloSpins(other) (copy)."""

    def __init__(self, num_samples, num_sweeps=None, num_thermalization_sweeps=None, num_markov_chains=None, gpu=True, seed=0xA11CE):
        self._h = C.c_void_p()
        if isinstance(num_samples, MonteCarloSpins):
            other = num_samples
            assert type(other) is type(self), "copy between MonteCarloSpins and MonteCarloPaulis"
            self.__dict__.update({k: v for k, v in other.__dict__.items() if k != "_h"})
            call("angpu_ensemble_copy", other._h, C.byref(self._h))
            return
        _require_gpu(gpu)
        self.num_samples, self.num_sweeps = int(num_samples), int(num_sweeps)
        self.num_thermalization_sweeps, self.num_markov_chains = int(num_thermalization_sweeps), int(num_markov_chains)
        self.seed = int(seed)
        call(self._create, self.num_samples, self.num_sweeps, self.num_thermalization_sweeps, self.num_markov_chains,
             self.seed, C.byref(self._h))

    _create = "angpu_mc_create"

    @property
    def acceptances(self):
        out = np.zeros(2, dtype=np.uint64)
        call("angpu_mc_acceptance", self._h, _p(out))
        return int(out[0]), int(out[1])

    @property
    def acceptance_rate(self):
        a, r = self.acceptances
        return float(a) / float(a + r) if a + r else float("nan")

    @property
    def call_index(self):
        """Number of sampling calls made so far = position of the random streams (they continue across calls)."""
        n = C.c_uint()
        call("angpu_mc_get_call_index", self._h, C.byref(n))
        return int(n.value)

    @call_index.setter
    def call_index(self, value):
        call("angpu_mc_set_call_index", self._h, int(value))

    @property
    def exact_decisions(self):
        """Proposals of the last call whose accept/reject decision needed the fp64 evaluation (fp32-screened PsiRBM
        sampler; 0 for the all-fp64 samplers).  No reference counterpart."""
        out = np.zeros(4, dtype=np.uint64)
        call("angpu_mc_counters", self._h, _p(out))
        return int(out[2])


# -------------------------------------------------------------------------------------------- functionals

def _match(op, psi):
    w = words_for(psi.num_sites) if psi.pauli_sites else psi.words          # operator masks are SITE masks on either basis
    return op.with_words(w) if op.words != w else op


class MonteCarloPaulis(MonteCarloSpins):
    """MonteCarloPaulis(num_samples, num_sweeps, num_thermalization_sweeps, num_markov_chains, gpu) (pyANNonGPU/main.cpp.template:369-377):
    Markov chains over Pauli strings (Init_Policy / Update_Policy<PauliString>), for a PsiDeep with N = 3 num_sites input units.
    A sweep is 3 num_sites proposals (MonteCarlo.hpp:90-101 counts psi.get_num_input_units())."""
    _create = "angpu_mc_paulis_create"


def paulis_to_units(a, b, num_sites):
    """The boundary form of a Pauli string (a, b): bit 3 s + t set iff site s carries type t + 1 (PauliString::network_unit_at)."""
    a, b = int(a), int(b)
    words = (3 * num_sites + 63) // 64
    v = 0
    for s in range(num_sites):
        t = ((a >> s) & 1) | (((b >> s) & 1) << 1)
        if t:
            v |= 1 << (3 * s + t - 1)
    return np.array([(v >> (64 * w)) & 0xFFFFFFFFFFFFFFFF for w in range(words)], dtype=np.uint64)


def units_to_paulis(units, num_sites):
    v = sum(int(x) << (64 * w) for w, x in enumerate(np.asarray(units, dtype=np.uint64).ravel()))
    a = b = 0
    for s in range(num_sites):
        t = 0
        for k in range(3):
            if (v >> (3 * s + k)) & 1:
                t = k + 1
        a |= (t & 1) << s
        b |= (t >> 1) << s
    return a, b


class ExpectationValue:
    """ExpectationValue(gpu) (pyANNonGPU/main.cpp.template:420-433)."""

    def __init__(self, gpu=True):
        _require_gpu(gpu)
        self._h = C.c_void_p()
        call("angpu_expval_create", C.byref(self._h))

    def __del__(self):
        if getattr(self, "_h", None):
            lib.angpu_expval_destroy(self._h)
            self._h = None

    def __call__(self, operator, psi, *args):
        """(operator | [operators], psi, ensemble) or the importance-reweighted (operator, psi, psi_sampling, ensemble)."""
        if len(args) == 2:
            psi_sampling, ensemble = args
            op = _match(operator, psi)
            out = np.empty(2)
            call("angpu_expectation_reweighted", self._h, op._h, psi._h, psi_sampling._h, ensemble._h, _p(out))
            return complex(out[0], out[1])
        (ensemble,) = args
        if isinstance(operator, (list, tuple)):
            ops = [_match(o, psi) for o in operator]
            handles = (C.c_void_p * max(1, len(ops)))(*[o._h for o in ops])
            out = np.empty(len(ops), dtype=np.complex128)
            call("angpu_expectation_many", self._h, len(ops), handles, psi._h, ensemble._h, _p(out))
            return out
        op = _match(operator, psi)
        out = np.empty(2)
        call("angpu_expectation", self._h, op._h, psi._h, ensemble._h, _p(out))
        return complex(out[0], out[1])

    def fluctuation(self, operator, psi, ensemble):
        op = _match(operator, psi)
        f, m = C.c_double(), np.empty(2)
        call("angpu_fluctuation", self._h, op._h, psi._h, ensemble._h, C.byref(f), _p(m))
        return float(f.value), complex(m[0], m[1])

    def gradient(self, operator, psi, ensemble):
        op = _match(operator, psi)
        g, m = np.empty(psi.num_params, dtype=np.complex128), np.empty(2)
        call("angpu_gradient", self._h, op._h, psi._h, ensemble._h, _p(g), _p(m))
        return g, complex(m[0], m[1])

    def exp_sigma_z(self, operator, psi, ensemble):
        """sum_s w_s exp(sum_n c_n coefficient_n(s)) (ExpectationValue.cu.template:52-82)."""
        op = _match(operator, psi)
        out = np.empty(2)
        call("angpu_exp_sigma_z", self._h, op._h, psi._h, ensemble._h, _p(out))
        return complex(out[0], out[1])

    def gradient_with_noise(self, operator, psi, ensemble):
        raise NotImplementedError("gradient_with_noise: the reference's body is commented out and returns uninitialised "
                                  "arrays (ExpectationValue.cu.template:276-333); there is nothing to reproduce")


class HilbertSpaceDistance:
    """HilbertSpaceDistance(num_params, gpu) (pyANNonGPU/main.cpp.template:436-440): ``hsd(psi, psi_prime, operator,
    is_unitary, ensemble) -> distance`` and ``hsd.gradient(psi, psi_prime, operator, is_unitary, ensemble, nu) ->
    (gradient[num_params of psi_prime], distance)``.  The reference binds it for (PsiDeep, PsiDeep) and (PsiCNN, PsiCNN);
    here any pair of models on the same lattice works."""

    def __init__(self, num_params, gpu=True):
        _require_gpu(gpu)
        self.num_params = int(num_params)
        self._h = C.c_void_p()
        call("angpu_hsd_create", self.num_params, C.byref(self._h))

    def __del__(self):
        if getattr(self, "_h", None):
            lib.angpu_hsd_destroy(self._h)
            self._h = None

    def __call__(self, psi, psi_prime, operator_, is_unitary, spin_ensemble):
        op = _match(operator_, psi)
        d = C.c_double()
        call("angpu_hsd_distance", self._h, psi._h, psi_prime._h, op._h, int(bool(is_unitary)), spin_ensemble._h, C.byref(d))
        return float(d.value)

    def gradient(self, psi, psi_prime, operator_, is_unitary, spin_ensemble, nu):
        op = _match(operator_, psi)
        g, d = np.empty(self.num_params, dtype=np.complex128), C.c_double()
        call("angpu_hsd_gradient", self._h, psi._h, psi_prime._h, op._h, int(bool(is_unitary)), spin_ensemble._h, float(nu),
             _p(g), C.byref(d))
        return g, float(d.value)


class KullbackLeibler:
    """KullbackLeibler(num_params, gpu) (pyANNonGPU/main.cpp.template:445-461): ``kl(psi, psi_prime, ensemble, threshold)``,
    ``kl.gradient(psi, psi_prime, ensemble, nu, threshold) -> (gradient, value)``, ``kl.gradient_with_noise(...) ->
    (gradient, noise, value)``; properties ``total_weight``, ``mean_deviation``, ``log_psi_scale`` (rw).  Samples are drawn
    from psi_prime; the mean deviation of one call is subtracted in the next (``last_mean_deviation`` upstream)."""

    def __init__(self, num_params, gpu=True):
        _require_gpu(gpu)
        self.num_params = int(num_params)
        self._h = C.c_void_p()
        call("angpu_kl_create", self.num_params, C.byref(self._h))

    def __del__(self):
        if getattr(self, "_h", None):
            lib.angpu_kl_destroy(self._h)
            self._h = None

    def _state(self):
        out = np.empty(4)
        call("angpu_kl_get_state", self._h, _p(out))
        return out

    total_weight = property(lambda self: float(self._state()[0]))
    mean_deviation = property(lambda self: complex(self._state()[1], self._state()[2]))

    @property
    def log_psi_scale(self):
        return float(self._state()[3])

    @log_psi_scale.setter
    def log_psi_scale(self, value):
        call("angpu_kl_set_log_psi_scale", self._h, float(value))

    def __call__(self, psi, psi_prime, ensemble, threshold):
        v = C.c_double()
        call("angpu_kl_value", self._h, psi._h, psi_prime._h, ensemble._h, float(threshold), C.byref(v))
        return float(v.value)

    def gradient(self, psi, psi_prime, ensemble, nu, threshold):
        g, v = np.empty(self.num_params, dtype=np.complex128), C.c_double()
        call("angpu_kl_gradient", self._h, psi._h, psi_prime._h, ensemble._h, float(nu), float(threshold), _p(g), C.byref(v))
        return g, float(v.value)

    def gradient_with_noise(self, psi, psi_prime, ensemble, nu, threshold):
        g, n, v = np.empty(self.num_params, dtype=np.complex128), np.empty(self.num_params), C.c_double()
        call("angpu_kl_gradient_with_noise", self._h, psi._h, psi_prime._h, ensemble._h, float(nu), float(threshold), _p(g), _p(n),
             C.byref(v))
        return g, n, float(v.value)


class TDVP:
    """TDVP(num_params, gpu) (pyANNonGPU/main.cpp.template:465-489)."""

    def __init__(self, num_params, gpu=True):
        _require_gpu(gpu)
        self.num_params = int(num_params)
        self.threshold = -1e6
        self._h = C.c_void_p()
        self._keep = None
        call("angpu_tdvp_create", self.num_params, C.byref(self._h))

    def __del__(self):
        if getattr(self, "_h", None):
            lib.angpu_tdvp_destroy(self._h)
            self._h = None

    def eval(self, operator, psi, ensemble, s_tolerance=0.0):
        """TDVP::eval.  `s_tolerance` (additive): 0 = S in exact fp64; >= 1e-5 (relative to ||S||) = S on the tcgen05
        tensor cores (3xTF32, ~2e-6 measured), several times faster for P ~ 1e4."""
        op = _match(operator, psi)
        self._keep = (op, psi, ensemble)
        if s_tolerance:
            call("angpu_tdvp_eval_tol", self._h, op._h, psi._h, ensemble._h, float(s_tolerance))
        else:
            call("angpu_tdvp_eval", self._h, op._h, psi._h, ensemble._h)

    def eval_F(self, operator, psi, ensemble):
        op = _match(operator, psi)
        self._keep = (op, psi, ensemble)
        call("angpu_tdvp_eval_F", self._h, op._h, psi._h, ensemble._h)

    def eval_with_psi_ref(self, operator, psi, ensemble, psi_sampling=None):
        """TDVP::eval(..., true_t) (include/network_functions/TDVP.hpp:90-93): samples drawn from psi.psi_ref (a
        PsiClassical's reference state), weights w |psi/psi_ref|^2, un-normalised sums; see `total_weight`.
        `psi_sampling` (additive) overrides the sampling state, so any psi can be importance-sampled."""
        if psi_sampling is None and not isinstance(psi, _PsiClassical):
            raise TypeError("eval_with_psi_ref needs a PsiClassical (its reference state is sampled) or an explicit psi_sampling")
        op = _match(operator, psi)
        self._keep = (op, psi, ensemble, psi_sampling)
        call("angpu_tdvp_eval_reweighted", self._h, op._h, psi._h, psi_sampling._h if psi_sampling is not None else None, ensemble._h)

    def _vec(self, name, n):
        out = np.empty(n, dtype=np.complex128)
        call(name, self._h, _p(out))
        return out

    @property
    def S_matrix(self):
        return self._vec("angpu_tdvp_get_S", self.num_params * self.num_params).reshape(self.num_params, self.num_params)

    @property
    def F_vector(self):
        return self._vec("angpu_tdvp_get_F", self.num_params)

    @property
    def O_k_vector(self):
        return self._vec("angpu_tdvp_get_O_k", self.num_params)

    def _scalars(self):
        out = np.empty(5)
        call("angpu_tdvp_get_scalars", self._h, _p(out))
        return out

    @property
    def E_local(self):
        s = self._scalars()
        return complex(s[0], s[1])

    @property
    def var_H(self):
        return float(self._scalars()[3])

    @property
    def total_weight(self):
        return float(self._scalars()[4])

    @property
    def num_local_samples(self):
        n = C.c_ulonglong()
        call("angpu_tdvp_num_local_samples", self._h, C.byref(n))
        return int(n.value)

    @property
    def O_k_samples(self):
        """Flat, as the reference binding returns it (main.cpp.template:473)."""
        return self._vec("angpu_tdvp_get_O_k_samples", self.num_local_samples * self.num_params)

    @property
    def weight_samples(self):
        out = np.empty(self.num_local_samples)
        call("angpu_tdvp_get_weights", self._h, _p(out))
        return out

    @property
    def E_local_samples(self):
        return self._vec("angpu_tdvp_get_E_local_samples", self.num_local_samples)

    def S_dot_vector(self, vec, ensemble=None):
        vec = _c128(vec).ravel()
        assert vec.size == self.num_params
        out = np.empty(self.num_params, dtype=np.complex128)
        call("angpu_tdvp_S_dot_vector", self._h, _p(vec), _p(out))
        return out

    def build_S_tensorcore(self):
        """NEW, opt-in: rebuild S on the tcgen05 tensor cores (3xTF32, ~1e-5 relative to ||S||); then S_matrix / solve use it."""
        call("angpu_tdvp_build_S_tensorcore", self._h)

    def set_tensorcore_products(self, enable=True):
        """PsiRBM: S_dot_vector and the CG search directions on the tcgen05 tensor cores (~1e-6 relative); solve_cg still recomputes the
        true residual with the exact product every 32 iterations and before accepting it.  True / False / None (None = auto, the
        default: S_dot_vector exact, solve_cg by problem size).  No reference counterpart."""
        call("angpu_tdvp_set_tensorcore_products", self._h, -1 if enable is None else (1 if enable else 0))
        return self

    def set_profile(self, enable=True):
        call("angpu_tdvp_set_profile", self._h, 1 if enable else 0)

    @property
    def phase_ms(self):
        """{sample, eloc, ok_reduce, total} device milliseconds of the last eval / eval_F (needs set_profile(True))."""
        out = np.empty(6)
        call("angpu_tdvp_phase_ms", self._h, _p(out))
        return dict(sample=float(out[0]), eloc=float(out[1]), ok_reduce=float(out[2]), total=float(out[3]),
                    s_build=float(out[4]), solve=float(out[5]))

    def solve_cg(self, tol=1e-6, max_iter=1000, shift_abs=0.0, shift_rel=1e-3, rhs_phase=1.0, keep_on_device=False):
        """NEW: matrix-free CG for (S + shift_abs I + shift_rel diag S) x = rhs_phase F. Returns (x, iterations, rel_residual);
        with keep_on_device the solution is not copied out (x is None) -- follow up with apply_update."""
        x = None if keep_on_device else np.empty(self.num_params, dtype=np.complex128)
        it, rr = C.c_uint(), C.c_double()
        call("angpu_tdvp_solve_cg", self._h, float(tol), int(max_iter), float(shift_abs), float(shift_rel),
             _p(_pair(rhs_phase)), None if x is None else _p(x), C.byref(it), C.byref(rr))
        return x, int(it.value), float(rr.value)

    def apply_update(self, psi, alpha):
        """NEW: psi.params += alpha * x for the x of the last solve, on the device (the SR / TDVP parameter step)."""
        call("angpu_tdvp_apply_update", self._h, psi._h, _p(_pair(alpha)))

    def solve(self, shift_abs=0.0, shift_rel=1e-3, rhs_phase=1.0):
        """NEW: dense Cholesky solve of the same system (needs eval, i.e. a dense S)."""
        x = np.empty(self.num_params, dtype=np.complex128)
        call("angpu_tdvp_solve_dense", self._h, float(shift_abs), float(shift_rel), _p(_pair(rhs_phase)), _p(x))
        return x


# -------------------------------------------------------------------------------------------- free functions

def log_psi_s(psi, conf):
    out = np.empty(2)
    call("angpu_log_psi_s", psi._h, _p(_conf_words(conf, psi.words)), _p(out))
    return complex(out[0], out[1])


def psi_O_k(psi, conf):
    out = np.empty(psi.num_params, dtype=np.complex128)
    call("angpu_psi_O_k", psi._h, _p(_conf_words(conf, psi.words)), _p(out))
    return out


def psi_O_k_vector(psi, ensemble):
    out = np.empty(psi.num_params, dtype=np.complex128)
    call("angpu_psi_O_k_vector", psi._h, ensemble._h, _p(out))
    return out


def log_psi(psi, ensemble):
    out = np.empty(2)
    call("angpu_log_psi_mean", psi._h, ensemble._h, _p(out))
    return complex(out[0], out[1])


def psi_vector(psi, ensemble):
    out = np.empty(ensemble.local_steps, dtype=np.complex128)
    call("angpu_psi_vector", psi._h, ensemble._h, _p(out))
    return out


def log_psi_vector(psi, ensemble):
    out = np.empty(ensemble.local_steps, dtype=np.complex128)
    call("angpu_log_psi_vector", psi._h, ensemble._h, _p(out))
    return out


def apply_operator(psi, op, ensemble):
    op = _match(op, psi)
    out = np.empty(ensemble.local_steps, dtype=np.complex128)
    call("angpu_apply_operator", psi._h, op._h, ensemble._h, _p(out))
    return out


def local_energies(psi, op, confs):
    """(log_psi, E_loc) on caller-given configurations [ns, words] (additive testing/analysis aid)."""
    op = _match(op, psi)
    confs = _u64(confs).reshape(-1, psi.words)
    lp, el = np.empty(len(confs), dtype=np.complex128), np.empty(len(confs), dtype=np.complex128)
    call("angpu_local_energies", psi._h, op._h, _p(confs), len(confs), _p(lp), _p(el))
    return lp, el


# keep a reference so the ctypes callback is not collected
_allreduce_cb = None


def set_allreduce(fn):
    """fn(dev_ptr:int, count:int) must sum `count` float64 at dev_ptr in place over all ranks; None disables."""
    global _allreduce_cb
    if fn is None:
        _allreduce_cb = None
        lib.angpu_set_allreduce(C.cast(None, ALLREDUCE_FN), None)
        return
    def trampoline(ptr, count, user):
        # ctypes would print and swallow an exception raised here, and the library would go on with un-reduced partial
        # sums: park it, report failure to C (the entry point then fails) and let _lib.check() re-raise it
        try:
            fn(int(ptr), int(count))
            return 0
        except BaseException as exc:                 # noqa: BLE001  (must not propagate into the C frames)
            _lib.pending_callback_error.append(exc)
            return 1

    _allreduce_cb = ALLREDUCE_FN(trampoline)
    lib.angpu_set_allreduce(_allreduce_cb, None)


def hpd_solve(A, b):
    """x = A^{-1} b for a Hermitian positive definite A on the device: the hand-written blocked Cholesky behind TDVP.solve."""
    A = _c128(A)
    b = _c128(b).ravel()
    assert A.ndim == 2 and A.shape[0] == A.shape[1] == b.size
    x = np.empty(b.size, dtype=np.complex128)
    call("angpu_hpd_solve", b.size, _p(A), _p(b), _p(x))
    return x


def comm_unique_id():
    """128-byte NCCL unique id (rank 0 creates it, every rank passes it to comm_init)."""
    buf = (C.c_ubyte * 128)()
    call("angpu_comm_unique_id", buf)
    return bytes(buf)


def comm_init(unique_id, rank, world):
    """Creates the library's own NCCL communicator on the current device; ensembles created afterwards are sharded."""
    buf = (C.c_ubyte * 128).from_buffer_copy(bytes(unique_id))
    call("angpu_comm_init", buf, int(rank), int(world))


def comm_destroy():
    call("angpu_comm_destroy")


def comm_rank():
    r, w = C.c_int(), C.c_int()
    call("angpu_comm_rank", C.byref(r), C.byref(w))
    return r.value, w.value
