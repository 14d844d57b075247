"""TEST INFRASTRUCTURE — the CPU oracle of the VMC hot path.

ctypes wrapper over ``oracle/_build/liboracle_port.so`` (``oracle/port/vmc_port.c``, the plain-C
restatement of the reference's per-sample arithmetic, multi-word capable) plus the numpy
restatement of the reference's *reductions over samples*:

  ExactSummation weights      include/ensembles/ExactSummation.hpp:54-72      (w = exp(2 Re log psi), un-normalised)
  MonteCarlo weights          source/ensembles/MonteCarlo.cu:33               (w = 1/num_samples)
  ExpectationValue::operator()  source/network_functions/ExpectationValue.cu.template:20-50
  ExpectationValue::fluctuation ...:176-216
  ExpectationValue::gradient    ...:220-275   grad_k = <O_k* E> - <E><O_k*>
  TDVP::eval / eval_F_vector    source/network_functions/TDVP.cu.template:78-126, 182-334
                                S = <O_k* O_k'> - <O_k>*<O_k'>, F = <E O_k*> - <E><O_k>*, var_H (TDVP.hpp:69-71)
  TDVP::S_dot_vector            ...:337-443
  psi_vector / log_psi / psi_norm / apply_operator / psi_O_k_vector   source/network_functions/{PsiVector,PsiNorm,
                                ApplyOperator,PsiOkVector}.cu.template

Pinned against the compiled reference by tests/test_oracle_pinned.py and the fixtures in tests/golden/.
Nothing in the product package may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "liboracle_port.so")
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "port"])


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "port", "vmc_port.c")
        if not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
            build()
        L = C.CDLL(LIB_PATH)
        vp, u32, u64, ul, dbl, i32 = C.c_void_p, C.c_uint, C.c_uint64, C.c_ulong, C.c_double, C.c_int
        for n in ("port_op_create", "port_rbm_create", "port_deep_create", "port_cnn_create", "port_classical_create"):
            getattr(L, n).restype = vp
        L.port_psi_num_params.restype = u32
        L.port_max_threads.restype = i32
        L.port_activation.argtypes = [dbl, dbl, u32, vp, vp]
        L.port_op_create.argtypes = [u32, vp, vp, vp, u32]
        L.port_op_destroy.argtypes = [vp]
        L.port_pauli_apply.argtypes = [vp, vp, vp, u32, vp, vp]
        L.port_rbm_create.argtypes = [u32, u32, vp, vp, vp]
        L.port_deep_create.argtypes = [u32, u32, vp, u32, vp, vp, vp, vp, vp, vp, vp]
        L.port_cnn_create.argtypes = [vp, u32, vp, vp, vp, vp, u32, dbl, vp]
        L.port_classical_create.argtypes = [u32, u32, u32, vp, vp, u32, vp, vp]
        L.port_psi_destroy.argtypes = [vp]
        L.port_psi_num_params.argtypes = [vp]
        L.port_psi_set_log_prefactor.argtypes = [vp, dbl, dbl]
        L.port_psi_get_params.argtypes = [vp, vp]
        L.port_psi_set_params.argtypes = [vp, vp]
        L.port_log_psi_s.argtypes = [vp, vp, vp]
        L.port_psi_O_k.argtypes = [vp, vp, vp]
        L.port_local_energy.argtypes = [vp, vp, vp, vp]
        L.port_eval_samples.argtypes = [vp, vp, vp, ul, vp, vp, vp, i32]
        L.port_philox.argtypes = [vp, vp, vp]
        L.port_mc_sample.argtypes = [vp, ul, u32, u32, ul, u64, C.c_uint32, ul, vp, vp, vp, i32]
        L.port_mc_gradient.argtypes = [vp, vp, ul, u32, u32, ul, u64, C.c_uint32, vp, vp, vp, i32]
        L.port_paulis_to_units.argtypes = [vp, vp, u32, vp]
        L.port_units_to_paulis.argtypes = [vp, u32, vp, vp]
        L.port_pauli_mul.argtypes = [vp, vp, vp, vp, u32, vp, vp, vp]
        L.port_paulis_enumerate.argtypes = [u64, u32, vp, vp]
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


def _cpair(z):
    z = complex(z)
    return np.array([z.real, z.imag])


def words_for(num_sites):
    return (int(num_sites) + 63) // 64


def conf_words(value, words):
    """Python int (arbitrary precision bitmask, bit i <-> site i) -> uint64[words]."""
    value = int(value)
    return np.array([(value >> (64 * w)) & 0xFFFFFFFFFFFFFFFF for w in range(words)], dtype=np.uint64)


def conf_int(arr):
    return sum(int(x) << (64 * w) for w, x in enumerate(np.asarray(arr, dtype=np.uint64).ravel()))


def enumerate_confs(num_sites):
    """Spins::enumerate for every basis index (include/basis/Spins.h:291-296): index == bitmask."""
    return np.arange(1 << num_sites, dtype=np.uint64).reshape(-1, 1)


def activation(z, layer):
    lc, th = np.empty(1, np.complex128), np.empty(1, np.complex128)
    z = complex(z)
    lib().port_activation(z.real, z.imag, int(layer), _p(lc), _p(th))
    return complex(lc[0]), complex(th[0])


def philox(ctr, key):
    ctr, key, out = _u32(ctr), _u32(key), np.zeros(4, np.uint32)
    lib().port_philox(_p(ctr), _p(key), _p(out))
    return out


# ---------------------------------------------------------------- objects

class Operator:
    """coeffs[n]; a, b: (n, words) uint64 masks (or (n,) when words == 1)."""

    def __init__(self, coeffs, a, b, words=None):
        self.coeffs = _c128(coeffs)
        self.num_strings = n = len(self.coeffs)
        a, b = _u64(a), _u64(b)
        if words is None:
            words = 1 if a.ndim == 1 else a.shape[1]
        self.words = words
        self.a, self.b = a.reshape(n, words), b.reshape(n, words)
        self.h = lib().port_op_create(n, _p(self.coeffs), _p(self.a), _p(self.b), words)

    def __del__(self):
        if getattr(self, "h", None):
            lib().port_op_destroy(self.h)

    def pauli_apply(self, n, conf):
        conf = _u64(conf)
        coeff, out = np.empty(1, np.complex128), np.zeros(self.words, np.uint64)
        lib().port_pauli_apply(_p(self.a[n]), _p(self.b[n]), _p(conf), self.words, _p(coeff), _p(out))
        return complex(coeff[0]), out

    def dense_matrix(self, num_sites):
        """<s'|H|s> as a dense 2^N x 2^N matrix from PauliString::apply semantics (independent check,
        mirrors the reference tests' use of ``H.matrix(N, 'spins')``, test/test_ExpectationValue.py:22-30)."""
        dim = 1 << num_sites
        H = np.zeros((dim, dim), dtype=np.complex128)
        for n in range(self.num_strings):
            for s in range(dim):
                f, sp = self.pauli_apply(n, conf_words(s, self.words))
                # apply() yields the matrix element entering E_loc(s) = sum_n c_n f_n psi(s')/psi(s)  => H[s, s']
                H[s, conf_int(sp)] += self.coeffs[n] * f
        return H


def pauli_apply(a, b, conf, words=1):
    a, b, conf = _u64(a).reshape(words), _u64(b).reshape(words), _u64(conf).reshape(words)
    coeff, out = np.empty(1, np.complex128), np.zeros(words, np.uint64)
    lib().port_pauli_apply(_p(a), _p(b), _p(conf), words, _p(coeff), _p(out))
    return complex(coeff[0]), out


class Psi:
    def __del__(self):
        if getattr(self, "h", None):
            lib().port_psi_destroy(self.h)

    @property
    def words(self):
        return words_for(getattr(self, "N", self.num_sites))      # N = 3 num_sites for a PsiDeep on the Pauli-string basis

    @property
    def num_params(self):
        return int(lib().port_psi_num_params(self.h))

    @property
    def params(self):
        out = np.empty(self.num_params, np.complex128)
        lib().port_psi_get_params(self.h, _p(out))
        return out

    @params.setter
    def params(self, value):
        value = _c128(value)
        assert value.size == self.num_params
        lib().port_psi_set_params(self.h, _p(value))

    @property
    def log_prefactor(self):
        return self._lp

    @log_prefactor.setter
    def log_prefactor(self, value):
        self._lp = complex(value)
        lib().port_psi_set_log_prefactor(self.h, self._lp.real, self._lp.imag)


class PsiRBM(Psi):
    def __init__(self, W, final_weight, log_prefactor):
        W = _c128(W)
        self.N, self.M = W.shape
        self.num_sites = self.N
        self._lp = complex(log_prefactor)
        self.h = lib().port_rbm_create(self.N, self.M, _p(W), _p(_cpair(final_weight)), _p(_cpair(log_prefactor)))


class PsiDeep(Psi):
    def __init__(self, num_sites, input_weights, biases, connections, weights, final_weights, log_prefactor):
        a = _c128(input_weights)
        sizes = _u32([len(b) for b in biases])
        conn = _u32([np.asarray(c).shape[0] for c in connections])
        b_cat = _c128(np.concatenate([np.asarray(b).ravel() for b in biases]))
        c_cat = _u32(np.concatenate([np.asarray(c).ravel() for c in connections]))
        w_cat = _c128(np.concatenate([np.asarray(w).ravel() for w in weights]))
        fw = _c128(final_weights)
        self.num_sites, self.N = num_sites, len(a)
        self._lp = complex(log_prefactor)
        self.h = lib().port_deep_create(num_sites, len(a), _p(a), len(sizes), _p(sizes), _p(conn), _p(b_cat),
                                        _p(c_cat), _p(w_cat), _p(fw), _p(_cpair(log_prefactor)))


class PsiCNN(Psi):
    def __init__(self, extent, num_channels_list, connectivity_list, symmetry_classes, params, final_factor, log_prefactor):
        ext = _u32(extent)
        nc, conn, sym, p = _u32(num_channels_list), _u32(connectivity_list), _u32(symmetry_classes), _c128(params)
        self.num_sites = self.N = int(np.prod(ext))
        self._lp = complex(log_prefactor)
        self.h = lib().port_cnn_create(_p(ext), len(nc), _p(nc), _p(conn), _p(sym), _p(p), p.size,
                                       float(final_factor), _p(_cpair(log_prefactor)))

    def init_gradient(self, num_steps):
        pass


class PsiClassical(Psi):
    def __init__(self, num_sites, order, H_local, params, psi_ref, log_prefactor):
        self.num_sites = self.N = num_sites
        self._ops, self._ref = list(H_local), psi_ref
        handles = (C.c_void_p * max(1, len(self._ops)))(*[op.h for op in self._ops])
        p = _c128(params)
        self._lp = complex(log_prefactor)
        self.h = lib().port_classical_create(num_sites, order, len(self._ops), handles, _p(p), p.size,
                                             psi_ref.h if psi_ref is not None else None, _p(_cpair(log_prefactor)))


# ---------------------------------------------------------------- per-configuration probes

def log_psi_s(psi, conf):
    conf, out = _u64(conf).reshape(psi.words), np.empty(1, np.complex128)
    lib().port_log_psi_s(psi.h, _p(conf), _p(out))
    return complex(out[0])


def psi_O_k(psi, conf):
    conf, out = _u64(conf).reshape(psi.words), np.empty(psi.num_params, np.complex128)
    lib().port_psi_O_k(psi.h, _p(conf), _p(out))
    return out


def local_energy(psi, op, conf):
    conf, out = _u64(conf).reshape(psi.words), np.empty(1, np.complex128)
    lib().port_local_energy(psi.h, op.h, _p(conf), _p(out))
    return complex(out[0])


def eval_samples(psi, op, confs, want_O=False, nthreads=0):
    """Per-sample (log_psi, E_loc, O rows) on the given configurations (ns, words)."""
    confs = _u64(confs).reshape(-1, psi.words)
    ns = confs.shape[0]
    lp = np.empty(ns, np.complex128)
    el = np.empty(ns, np.complex128) if op is not None else None
    O = np.empty((ns, psi.num_params), np.complex128) if want_O else None
    lib().port_eval_samples(psi.h, op.h if op is not None else None, _p(confs), ns, _p(lp), _p(el), _p(O), nthreads)
    return lp, el, O


# ---------------------------------------------------------------- ensembles

class ExactSummation:
    def __init__(self, num_sites):
        self.num_sites = num_sites
        self.num_steps = 1 << num_sites

    def samples(self, psi):
        confs = enumerate_confs(self.num_sites)
        return confs, None  # weights derive from log psi


# ---- Pauli-string basis (density-matrix ensembles).  A PsiDeep with N == 3 num_sites input units lives on
# this basis; configurations are handed around as the network-side "units" mask (vmc_port.c: port_paulis_to_units).

def paulis_to_units(a, b, num_sites):
    a, b = _u64(a).reshape(-1), _u64(b).reshape(-1)
    out = np.zeros(words_for(3 * num_sites), np.uint64)
    lib().port_paulis_to_units(_p(a), _p(b), num_sites, _p(out))
    return out


def units_to_paulis(units, num_sites):
    units = _u64(units).reshape(-1)
    a, b = np.zeros(words_for(num_sites), np.uint64), np.zeros(words_for(num_sites), np.uint64)
    lib().port_units_to_paulis(_p(units), num_sites, _p(a), _p(b))
    return a, b


def pauli_mul(Pa, Pb, xa, xb, words=1):
    """PauliString::apply(PauliString) (include/basis/PauliString.hpp:257-277): (factor, a', b')."""
    Pa, Pb, xa, xb = (conf_words(v, words) if np.isscalar(v) else _u64(v) for v in (Pa, Pb, xa, xb))
    c, oa, ob = np.empty(1, np.complex128), np.zeros(words, np.uint64), np.zeros(words, np.uint64)
    lib().port_pauli_mul(_p(Pa), _p(Pb), _p(xa), _p(xb), words, _p(c), _p(oa), _p(ob))
    return complex(c[0]), oa, ob


def paulis_enumerate(index, num_sites):
    a, b = np.zeros(words_for(num_sites), np.uint64), np.zeros(words_for(num_sites), np.uint64)
    lib().port_paulis_enumerate(int(index), num_sites, _p(a), _p(b))
    return a, b


def enumerate_pauli_units(num_sites):
    """All 4^num_sites Pauli strings in PauliString::enumerate order, as units masks (ns, words_for(3 num_sites))."""
    return np.stack([paulis_to_units(*paulis_enumerate(i, num_sites), num_sites) for i in range(4 ** num_sites)])


class ExactSummationPaulis:
    """ExactSummation_t<PauliString> (include/ensembles/ExactSummation.hpp:124-126)."""

    def __init__(self, num_sites):
        self.num_sites = num_sites
        self.num_steps = 4 ** num_sites


class MonteCarlo:
    """Philox-driven restatement of MonteCarlo_t (all chains run; see vmc_port.c header).  With a PsiDeep on the Pauli-string
    basis (N == 3 num_sites) the chain uses Init_Policy / Update_Policy<PauliString> -- MonteCarloPaulis is the same class."""

    def __init__(self, num_samples, num_sweeps, num_thermalization_sweeps, num_markov_chains, seed=0xA11CE, chain0=0):
        self.num_samples, self.num_sweeps = num_samples, num_sweeps
        self.num_thermalization_sweeps, self.num_markov_chains = num_thermalization_sweeps, num_markov_chains
        self.seed, self.call, self.chain0 = seed, 0, chain0
        self.num_steps = num_samples
        self.acceptances = self.rejections = 0

    def sample(self, psi, nthreads=0):
        ns = (self.num_samples // self.num_markov_chains) * self.num_markov_chains
        confs = np.zeros((max(ns, 1), psi.words), np.uint64)
        lp = np.zeros(max(ns, 1), np.complex128)
        ar = np.zeros(2, np.uint64)
        lib().port_mc_sample(psi.h, self.num_samples, self.num_sweeps, self.num_thermalization_sweeps,
                             self.num_markov_chains, self.seed, self.call, self.chain0, _p(confs), _p(lp), _p(ar), nthreads)
        self.call += 1
        self.acceptances, self.rejections = int(ar[0]), int(ar[1])
        return confs[:ns], lp[:ns]

    @property
    def acceptance_rate(self):
        return self.acceptances / max(1, self.acceptances + self.rejections)


MonteCarloPaulis = MonteCarlo


def _samples_and_weights(psi, ens):
    if isinstance(ens, ExactSummationPaulis):
        assert psi.N == 3 * ens.num_sites
        confs = enumerate_pauli_units(ens.num_sites)
        lp, _, _ = eval_samples(psi, None, confs)
        return confs, lp, np.exp(2.0 * lp.real)
    if isinstance(ens, ExactSummation):
        confs = enumerate_confs(ens.num_sites)
        lp, _, _ = eval_samples(psi, None, confs)
        return confs, lp, np.exp(2.0 * lp.real)
    confs, lp = ens.sample(psi)
    return confs, lp, np.full(len(confs), 1.0 / ens.num_samples)


def _reweighted(psi, psi_sampling, ens):
    """Configurations from |psi_sampling|^2, weights w |psi/psi_sampling|^2 (ExpectationValue.cu.template:127-172,
    TDVP.cu.template:28-59)."""
    confs, lp_s, w = _samples_and_weights(psi_sampling, ens)
    lp, _, _ = eval_samples(psi, None, confs)
    return confs, lp, w * np.exp(2.0 * (lp.real - lp_s.real))


def fully_polarized(num_sites):
    """PsiFullyPolarized (include/quantum_state/PsiFullyPolarized.hpp:41-49): log psi = 0, no parameters."""
    return PsiClassical(num_sites, 1, [], np.zeros(0, dtype=complex), None, 0.0)


# ---------------------------------------------------------------- network functions (reductions in numpy)

def psi_vector(psi, ens):
    return np.exp(log_psi_vector(psi, ens))


def log_psi_vector(psi, ens):
    return _samples_and_weights(psi, ens)[1]


def log_psi(psi, ens):
    _, lp, w = _samples_and_weights(psi, ens)
    return complex(np.sum(w * lp))


def psi_norm(psi, es):
    _, _, w = _samples_and_weights(psi, es)
    return float(np.sqrt(np.sum(w)))


def psi_O_k_vector(psi, es):
    confs, _, _ = _samples_and_weights(psi, es)
    return eval_samples(psi, None, confs, want_O=True)[2].sum(axis=0)


def apply_operator(psi, op, ens):
    confs, lp, _ = _samples_and_weights(psi, ens)
    _, el, _ = eval_samples(psi, op, confs)
    return np.exp(lp) * el


def expectation(op, psi, ens):
    confs, _, w = _samples_and_weights(psi, ens)
    _, el, _ = eval_samples(psi, op, confs)
    return complex(np.sum(w * el))


def expectation_reweighted(op, psi, psi_sampling, ens):
    """The evident intent of ExpectationValue::operator()(op, psi, psi_sampling, ens) (:127-172); the reference itself
    never accumulates its denominator."""
    confs, _, w = _reweighted(psi, psi_sampling, ens)
    _, el, _ = eval_samples(psi, op, confs)
    return complex(np.sum(w * el) / np.sum(w))


def exp_sigma_z(op, psi, ens):
    """sum_s w_s exp(fast_local_energy(s)) (ExpectationValue.cu.template:52-82, Operator.hpp:123-136)."""
    confs, _, w = _samples_and_weights(psi, ens)
    e = np.zeros(len(confs), dtype=complex)
    for s, conf in enumerate(confs):
        for n in range(op.num_strings):
            e[s] += op.coeffs[n] * op.pauli_apply(n, conf)[0]
    return complex(np.sum(w * np.exp(e)))


def fluctuation(op, psi, ens):
    confs, _, w = _samples_and_weights(psi, ens)
    _, el, _ = eval_samples(psi, op, confs)
    A, A2 = np.sum(w * el), np.sum(w * np.abs(el) ** 2)
    return float(np.sqrt(A2 - abs(A) ** 2)), complex(A)


def gradient(op, psi, ens):
    confs, _, w = _samples_and_weights(psi, ens)
    _, el, O = eval_samples(psi, op, confs, want_O=True)
    E = np.sum(w * el)
    Oc = np.conj(O)
    return (w * el) @ Oc - E * (w @ Oc), complex(E)


def _hsd_averages(psi, psi_prime, op, is_unitary, ens, want_gradient):
    """kernel::HilbertSpaceDistance::compute_averages (source/network_functions/HilbertSpaceDistance.cu.template:16-89)."""
    confs, lp, w = _samples_and_weights(psi, ens)
    _, el, _ = eval_samples(psi, op, confs)
    lpp, _, Op = eval_samples(psi_prime, None, confs, want_O=want_gradient)
    if is_unitary:
        omega = np.exp(np.conj(lpp - lp)) * el
        next_state_norm = np.sum(w * np.abs(el) ** 2)
    else:
        omega = np.exp(el + np.conj(lpp - lp))
        next_state_norm = np.sum(w * np.exp(2.0 * el.real))
    pr = np.exp(2.0 * (lpp.real - lp.real))
    out = {"omega": np.sum(w * omega), "pr": np.sum(w * pr), "nsn": next_state_norm}
    if want_gradient:
        out["omega_Ok"] = (w * omega) @ np.conj(Op)
        out["pr_Ok"] = (w * pr) @ np.conj(Op)
    return out


def hilbert_space_distance(psi, psi_prime, op, is_unitary, ens):
    """HilbertSpaceDistance::distance (:121-137)."""
    a = _hsd_averages(psi, psi_prime, op, is_unitary, ens, False)
    u, v = abs(a["omega"]) ** 2, a["nsn"] * a["pr"]
    return float(np.sqrt(max(1.0 - u / v, 1e-8)))


def hilbert_space_distance_gradient(psi, psi_prime, op, is_unitary, ens, nu):
    """HilbertSpaceDistance::gradient (:140-171) -> (gradient[P'], distance); nu is a float in the reference."""
    a = _hsd_averages(psi, psi_prime, op, is_unitary, ens, True)
    u, v = abs(a["omega"]) ** 2, a["nsn"] * a["pr"]
    distance = float(np.sqrt(max(1.0 - u / v, 1e-8)))
    prefactor = distance ** float(np.float32(nu))
    u_k = np.conj(a["omega"]) * a["omega_Ok"]
    v_k = a["nsn"] * a["pr_Ok"]
    return -(u_k * v - u * v_k) / (v * v) / prefactor, distance


class KullbackLeibler:
    """Restatement of KullbackLeibler (source/network_functions/KullbackLeibler.cu.template:16-321): samples from psi_prime,
    weights w' |psi^scale / psi'|^2, deviation = log psi' - scale log psi - last_mean_deviation (the latter carried from
    the previous call), terms with |deviation| <= threshold dropped from the deviation sums."""

    def __init__(self, num_params):
        self.num_params, self.log_psi_scale = int(num_params), 1.0
        self.last_mean_deviation, self.mean_deviation, self.total_weight = 0j, 0j, 0.0

    def _averages(self, psi, psi_prime, ens, threshold, want_O):
        confs, lpp, wp = _samples_and_weights(psi_prime, ens)
        lp = eval_samples(psi, None, confs)[0] * self.log_psi_scale
        w = wp * np.exp(2.0 * (lp.real - lpp.real))
        tw = np.sum(w)
        dev = lpp - lp - self.last_mean_deviation
        dev2 = np.abs(dev) ** 2
        keep = dev2 > threshold * threshold
        out = {"tw": tw, "mean_dev": np.sum(w * (lpp - lp)) / tw, "d": np.sum((w * dev)[keep]) / tw, "d2": np.sum((w * dev2)[keep]) / tw}
        if want_O:
            O = eval_samples(psi_prime, None, confs, want_O=True)[2]
            wk, O2 = np.where(keep, w, 0.0), np.abs(O) ** 2
            out.update(O=(w @ O) / tw, dOc=((wk * dev) @ np.conj(O)) / tw, d2O2=((wk * dev2) @ O2) / tw, dO=((wk * dev) @ O) / tw,
                       dO2=((wk * dev) @ O2) / tw, d2O=((wk * dev2) @ O) / tw, O2=(w @ O2) / tw)
        self.total_weight, self.mean_deviation = float(tw), complex(out["mean_dev"])
        self.last_mean_deviation = self.mean_deviation
        out["value"] = float(np.sqrt(max(1e-8, out["d2"] - abs(out["d"]) ** 2)))
        return out

    def __call__(self, psi, psi_prime, ens, threshold):
        return self._averages(psi, psi_prime, ens, threshold, False)["value"]

    def gradient(self, psi, psi_prime, ens, nu, threshold):
        a = self._averages(psi, psi_prime, ens, threshold, True)
        return (a["dOc"] - a["d"] * np.conj(a["O"])) / a["value"] ** nu, a["value"]

    def gradient_with_noise(self, psi, psi_prime, ens, nu, threshold):
        a = self._averages(psi, psi_prime, ens, threshold, True)
        f = a["value"] ** nu
        d, d2, O = a["d"], a["d2"], a["O"]
        var = (a["d2O2"] - np.abs(a["dOc"]) ** 2 + 2.0 * (a["dO"] * np.conj(d) * np.conj(O) + 2.0 * np.conj(a["dOc"]) * d * np.conj(O)
                                                              - a["d2O"] * np.conj(O) - a["dO2"] * np.conj(d)).real
               + d2 * np.abs(O) ** 2 + abs(d) ** 2 * a["O2"] - 4.0 * abs(d) ** 2 * np.abs(O) ** 2)
        with np.errstate(invalid="ignore"):
            noise = np.sqrt(var / ens.num_steps) / f
        return (a["dOc"] - d * np.conj(O)) / f, noise, a["value"]


class TDVP:
    def __init__(self, num_params):
        self.P = int(num_params)

    def eval(self, op, psi, ens, want_S=True, psi_sampling=None):
        confs, _, w = _samples_and_weights(psi, ens) if psi_sampling is None else _reweighted(psi, psi_sampling, ens)
        self.total_weight = float(np.sum(w))
        _, el, O = eval_samples(psi, op, confs, want_O=True)
        self.confs, self.E_loc_samples = confs, el
        self.O_k_samples, self.weight_samples = O, w
        self.E_local = complex(np.sum(w * el))
        self.E2_local = float(np.sum(w * np.abs(el) ** 2))
        self.O_k_vector = w @ O
        self.F_vector = (w * el) @ np.conj(O) - self.E_local * np.conj(self.O_k_vector)
        if want_S:
            self.S_matrix = (np.conj(O).T * w) @ O - np.outer(np.conj(self.O_k_vector), self.O_k_vector)
        self.var_H = self.E2_local - abs(self.E_local) ** 2

    def eval_F(self, op, psi, ens):
        self.eval(op, psi, ens, want_S=False)

    def eval_with_psi_ref(self, op, psi, ens, psi_sampling=None):
        """TDVP::eval(..., true_t) (TDVP.cu.template:15-74): samples from psi's reference state, un-normalised sums."""
        if psi_sampling is None:
            psi_sampling = psi._ref if psi._ref is not None else fully_polarized(psi.num_sites)
        self.eval(op, psi, ens, want_S=True, psi_sampling=psi_sampling)

    def S_dot_vector(self, vec, ens=None):
        vec = _c128(vec)
        O, w = self.O_k_samples, self.weight_samples
        return np.conj(O).T @ (w * (O @ vec)) - np.conj(self.O_k_vector) * (self.O_k_vector @ vec)


def mc_gradient_timed(psi, op, num_samples, num_sweeps, num_therm, num_chains, seed=0xA11CE, call=0, nthreads=0):
    """One ExpectationValue::gradient call over a Monte-Carlo ensemble, entirely in C (the CPU baseline)."""
    g, e, ar = np.empty(psi.num_params, np.complex128), np.empty(1, np.complex128), np.zeros(2, np.uint64)
    lib().port_mc_gradient(psi.h, op.h, num_samples, num_sweeps, num_therm, num_chains, seed, call, _p(g), _p(e), _p(ar), nthreads)
    return g, complex(e[0]), (int(ar[0]), int(ar[1]))


def max_threads():
    return int(lib().port_max_threads())
