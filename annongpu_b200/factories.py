"""Synthetic-input factories for the VMC hot path (host side, numpy only).

Seeded restatements of the reference's random-init factories — ``pyANNonGPU/new_RBM.py:14-29``,
``pyANNonGPU/new_neural_network.py:31-131``, ``pyANNonGPU/new_convolutional_network.py:30-82`` — that
return plain *specs* (dataclasses of numpy arrays).  A spec can be turned into a device object of this
package (``spec.build()``), or handed to any other implementation of the same interface.

The reference builds operators from ``QuantumExpression.PauliExpression`` (absent here); the raw
form it reduces them to is (coefficient, a-mask, b-mask) per Pauli string
(``source/operator/Operator.cpp:18-39``) with X=(1,0), Y=(0,1), Z=(1,1)
(``include/basis/PauliString.hpp:37-56``).  ``PauliSum`` below is that raw form with arbitrary-width
masks (Python ints), so N > 64 sites is representable.
"""
import math
from dataclasses import dataclass, field

import numpy as np


def _real_noise(rng, shape):
    return 2.0 * rng.random(shape) - 1.0


def _complex_noise(rng, shape):
    return _real_noise(rng, shape) + 1j * _real_noise(rng, shape)


def _noise_vector(rng, shape, real):
    return _real_noise(rng, shape) if real else _complex_noise(rng, shape)


def words_for(num_sites):
    return (int(num_sites) + 63) // 64


def masks_to_words(masks, words):
    """list of Python-int bitmasks -> uint64 array (len, words), little-endian words."""
    out = np.zeros((len(masks), words), dtype=np.uint64)
    for n, m in enumerate(masks):
        m = int(m)
        for w in range(words):
            out[n, w] = (m >> (64 * w)) & 0xFFFFFFFFFFFFFFFF
    return out


# ---------------------------------------------------------------------------- operators

@dataclass
class PauliSum:
    """Sum of Pauli strings in the reference's raw (coefficient, a, b) form."""
    num_sites: int
    coeffs: list = field(default_factory=list)
    a: list = field(default_factory=list)
    b: list = field(default_factory=list)

    def add(self, coeff, paulis):
        """paulis: dict site -> 'X' | 'Y' | 'Z'."""
        a = b = 0
        for site, kind in paulis.items():
            assert 0 <= site < self.num_sites
            if kind in ("X", "Z"):
                a |= 1 << site
            if kind in ("Y", "Z"):
                b |= 1 << site
        self.coeffs.append(complex(coeff))
        self.a.append(a)
        self.b.append(b)
        return self

    @property
    def num_strings(self):
        return len(self.coeffs)

    @property
    def words(self):
        return words_for(self.num_sites)

    def arrays(self, words=None):
        words = words or self.words
        return (np.asarray(self.coeffs, dtype=np.complex128), masks_to_words(self.a, words), masks_to_words(self.b, words))

    def build(self, gpu=True):
        from .api import Operator
        return Operator(self, gpu)

    # ---- operator algebra: a native stand-in for the QuantumExpression `PauliExpression` the reference's Python layer
    # and tests build operators with (source/operator/Operator.cpp:18-39 consumes only term.first.a/.b and term.second).
    # Per site the masks encode I = (0,0), X = (1,0), Y = (0,1), Z = (1,1) (include/basis/PauliString.hpp:18-21).
    def copy(self):
        return PauliSum(self.num_sites, list(self.coeffs), list(self.a), list(self.b))

    def simplified(self, tol=0.0):
        """Merge equal strings, drop those with |coefficient| <= tol."""
        acc = {}
        for c, a, b in zip(self.coeffs, self.a, self.b):
            acc[(a, b)] = acc.get((a, b), 0.0) + c
        out = PauliSum(self.num_sites)
        for (a, b), c in acc.items():
            if abs(c) > tol:
                out.coeffs.append(complex(c)); out.a.append(a); out.b.append(b)
        return out

    def __add__(self, other):
        if isinstance(other, (int, float, complex)):
            other = PauliSum(self.num_sites).add(other, {})
        # (expressions built site by site, e.g. sigma_z(0) * sigma_z(1), grow to the larger support)
        return PauliSum(max(self.num_sites, other.num_sites), self.coeffs + other.coeffs, self.a + other.a, self.b + other.b).simplified()

    __radd__ = __add__

    def __neg__(self):
        return PauliSum(self.num_sites, [-c for c in self.coeffs], list(self.a), list(self.b))

    def __sub__(self, other):
        return self + (-other if isinstance(other, PauliSum) else -complex(other))

    def __rsub__(self, other):
        return (-self) + other

    @staticmethod
    def _string_product(a1, b1, a2, b2):
        """(phase, a, b) of the product P1 P2 of two Pauli strings (P1 acts after P2)."""
        # per-site type code t = a + 2 b: 0 I, 1 X, 2 Y, 3 Z;  X Y = i Z, Y Z = i X, Z X = i Y (and -i reversed)
        phase_pow = 0                    # accumulated power of i
        x1, y1, z1 = a1 & ~b1, ~a1 & b1, a1 & b1
        x2, y2, z2 = a2 & ~b2, ~a2 & b2, a2 & b2
        plus = (x1 & y2) | (y1 & z2) | (z1 & x2)
        minus = (y1 & x2) | (z1 & y2) | (x1 & z2)
        phase_pow = (bin(plus).count("1") - bin(minus).count("1")) % 4
        # result type per site: bitwise in the (x, z) symplectic picture: X-part = a ^ b ... with this encoding the
        # flip mask is f = a ^ b and the sign mask is b; products XOR both
        f = (a1 ^ b1) ^ (a2 ^ b2)
        bb = b1 ^ b2
        return 1j ** phase_pow, f ^ bb, bb

    def __mul__(self, other):
        if isinstance(other, (int, float, complex)):
            return PauliSum(self.num_sites, [c * other for c in self.coeffs], list(self.a), list(self.b))
        out = PauliSum(max(self.num_sites, other.num_sites))
        for c1, a1, b1 in zip(self.coeffs, self.a, self.b):
            for c2, a2, b2 in zip(other.coeffs, other.a, other.b):
                ph, a, b = self._string_product(a1, b1, a2, b2)
                out.coeffs.append(c1 * c2 * ph); out.a.append(a); out.b.append(b)
        return out.simplified()

    def __rmul__(self, other):
        return self * other

    def dagger(self):
        """Hermitian conjugate (Pauli strings are Hermitian: only the coefficients are conjugated)."""
        return PauliSum(self.num_sites, [c.conjugate() for c in self.coeffs], list(self.a), list(self.b))

    def roll(self, shift):
        """Translate every string by `shift` sites on the ring of num_sites sites (QuantumExpression's .roll)."""
        n, mask = self.num_sites, (1 << self.num_sites) - 1
        shift %= n
        rot = lambda m: ((m << shift) | (m >> (n - shift))) & mask if shift else m   # noqa: E731
        return PauliSum(n, list(self.coeffs), [rot(a) for a in self.a], [rot(b) for b in self.b])

    def commutator(self, other):
        return self * other - other * self

    def exp(self, threshold=0.0):
        """exp(self) as a PauliSum for a SINGLE Pauli string c P (P^2 = 1): cosh(c) 1 + sinh(c) P -- the only use the
        reference makes of QuantumExpression's .exp (pyANNonGPU/LearningByGradientDescent.py:356, `term.exp(0)`, the
        factors of a Trotterised propagator).  Terms with |coefficient| <= threshold are dropped."""
        t = self.simplified()
        if t.num_strings == 0:
            return PauliSum(self.num_sites).add(1.0, {})
        if t.num_strings != 1:
            raise ValueError("PauliSum.exp is defined for a single Pauli string (factor a sum term by term)")
        c, a, b = t.coeffs[0], t.a[0], t.b[0]
        if a == 0 and b == 0:
            return PauliSum(self.num_sites).add(np.exp(c), {})
        import cmath
        out = PauliSum(self.num_sites, [cmath.cosh(c), cmath.sinh(c)], [0, a], [0, b])
        return out.simplified(threshold)

    def to_json(self):
        """{"type": "PauliSum", "num_sites": N, "terms": [{"re": .., "im": .., "paulis": {"3": "X", ...}}, ...]}
        (QuantumExpression's own PauliExpression.to_json encoding is not available -- the dependency is absent and
        unpinned, SURVEY.md 8c -- so operators inside JSON documents use this self-describing form)."""
        terms = []
        for c, a, b in zip(self.coeffs, self.a, self.b):
            paulis = {}
            for i in range(self.num_sites):
                code = ((a >> i) & 1) | (((b >> i) & 1) << 1)
                if code:
                    paulis[str(i)] = " XYZ"[code]
            terms.append({"re": c.real, "im": c.imag, "paulis": paulis})
        return {"type": "PauliSum", "num_sites": self.num_sites, "terms": terms}

    @staticmethod
    def from_json(obj):
        out = PauliSum(int(obj["num_sites"]))
        for t in obj["terms"]:
            out.add(complex(t["re"], t["im"]), {int(k): v for k, v in t["paulis"].items()})
        return out

    def matrix(self):
        """Dense 2^N x 2^N matrix with the reference's conventions: basis index = configuration bitmask (bit i <-> site i,
        1 <-> spin up), M[s, s'] = sum_n c_n <s| P_n |s'> as used by E_loc(s) = sum_s' M[s, s'] psi(s') / psi(s)
        (include/basis/PauliString.hpp:193-255, include/operator/Operator.hpp:38-121).  Small N only."""
        n, dim = self.num_sites, 1 << self.num_sites
        assert n <= 14
        M = np.zeros((dim, dim), dtype=np.complex128)
        s = np.arange(dim, dtype=np.int64)
        for c, a, b in zip(self.coeffs, self.a, self.b):
            ny = bin(~a & b & (dim - 1)).count("1")
            neg = np.array([bin(int(v)).count("1") for v in (~s & b & (dim - 1))]) & 1
            coeff = c * ((-1j) ** ny) * np.where(neg, -1.0, 1.0)
            M[s, s ^ (a ^ b)] += coeff
        return M


def sigma_x(site, num_sites=None):
    """QuantumExpression-style constructors: sigma_x(i), sigma_y(i), sigma_z(i) (the support grows under + and *)."""
    return PauliSum(num_sites or site + 1).add(1.0, {site: "X"})


def sigma_y(site, num_sites=None):
    return PauliSum(num_sites or site + 1).add(1.0, {site: "Y"})


def sigma_z(site, num_sites=None):
    return PauliSum(num_sites or site + 1).add(1.0, {site: "Z"})


def propagator(H, dt):
    """First-order time-step operator 1 - i dt H as a PauliSum (identity string a = b = 0): the kind of operator the
    reference's HilbertSpaceDistance is used with (pyANNonGPU/LearningByGradientDescent.py)."""
    U = PauliSum(H.num_sites).add(1.0, {})
    for c, a, b in zip(H.coeffs, H.a, H.b):
        U.coeffs.append(-1j * dt * c)
        U.a.append(a)
        U.b.append(b)
    return U


def scaled(H, factor):
    """factor * H as a new PauliSum."""
    out = PauliSum(H.num_sites)
    out.coeffs, out.a, out.b = [factor * c for c in H.coeffs], list(H.a), list(H.b)
    return out


def ring_bonds(n):
    return [(i, (i + 1) % n) for i in range(n)]


def square_lattice_bonds(rows, cols):
    """Periodic rows x cols lattice, site = row * cols + col; 2 * rows * cols bonds."""
    bonds = []
    for r in range(rows):
        for c in range(cols):
            s = r * cols + c
            bonds.append((s, r * cols + (c + 1) % cols))
            bonds.append((s, ((r + 1) % rows) * cols + c))
    return bonds


def tfim(num_sites, bonds, J=1.0, h=1.0):
    """-J sum_<ij> Z_i Z_j - h sum_i X_i   (SURVEY.md §8d)."""
    H = PauliSum(num_sites)
    for i, j in bonds:
        H.add(-J, {i: "Z", j: "Z"})
    for i in range(num_sites):
        H.add(-h, {i: "X"})
    return H


def heisenberg(num_sites, bonds, J=1.0):
    """J sum_<ij> (X_i X_j + Y_i Y_j + Z_i Z_j)   (SURVEY.md §8d)."""
    H = PauliSum(num_sites)
    for i, j in bonds:
        for k in "XYZ":
            H.add(J, {i: k, j: k})
    return H


# ---------------------------------------------------------------------------- wavefunction specs

@dataclass
class RBMSpec:
    W: np.ndarray
    final_weight: complex
    log_prefactor: complex = 0.0

    @property
    def num_sites(self):
        return self.W.shape[0]

    def build(self, gpu=True):
        from .api import PsiRBM
        return PsiRBM(self.W, self.final_weight, self.log_prefactor, gpu)


@dataclass
class DeepSpec:
    num_sites: int
    input_weights: np.ndarray
    biases: list
    connections: list
    weights: list
    final_weights: np.ndarray
    log_prefactor: complex = 0.0

    def build(self, gpu=True):
        from .api import PsiDeep
        return PsiDeep(self.num_sites, self.input_weights, self.biases, self.connections, self.weights,
                       self.final_weights, self.log_prefactor, gpu)


@dataclass
class CNNSpec:
    extent: tuple
    num_channels_list: np.ndarray
    connectivity_list: np.ndarray
    symmetry_classes: np.ndarray
    params: np.ndarray
    final_factor: float
    log_prefactor: complex = 0.0

    @property
    def num_sites(self):
        return int(np.prod(self.extent))

    def build(self, gpu=True):
        from .api import PsiCNN
        return PsiCNN(self.extent, self.num_channels_list, self.connectivity_list, self.symmetry_classes,
                      self.params, self.final_factor, self.log_prefactor, gpu)


def rbm_spec(N, M, initial_value=(0.01 + 1j * math.pi / 4), noise=1e-4, final_weight=10, seed=0):
    """pyANNonGPU/new_RBM.py:14-29 with an explicit seed."""
    assert M >= N
    rng = np.random.default_rng(seed)
    W = noise * _complex_noise(rng, (N, M))
    for alpha in range(M // N):
        W[:, alpha * N:(alpha + 1) * N] += initial_value * np.eye(N)
    return RBMSpec(W, complex(final_weight), 0.0)


def _prod(xs):
    r = 1
    for x in xs:
        r *= x
    return r


def deep_spec(num_sites, N, M, C, initial_value=(0.01 + 1j * math.pi / 4), a=0, noise=1e-4,
              noise_modulation="auto", final_weights=10, seed=0):
    """pyANNonGPU/new_neural_network.py:31-131 (1-D connectivity form) with an explicit seed."""
    assert not isinstance(N, (list, tuple)), "only the 1-D form is restated here"
    rng = np.random.default_rng(seed)
    for n, m, c in zip([N] + M[:-1], M, C):
        assert (m * c) % n == 0 and c <= n
    a = a * np.ones(N, dtype=complex) if isinstance(a, (float, int, complex)) else np.array(a, dtype=complex)
    is_real = (complex(initial_value).imag == 0)

    b = [noise * _noise_vector(rng, m, is_real).astype(complex) for m in M]
    w = (noise * _noise_vector(rng, (C[0], M[0]), is_real)).astype(complex)
    w[C[0] // 2, :] += initial_value
    W = [w]
    if noise_modulation == "auto":
        noise_modulation = [math.sqrt(6 / (c + next_c)) for c, m, next_c in zip(C[1:], M[1:], C[2:] + [1])]
    for c, m, next_c, nm in zip(C[1:], M[1:], C[2:] + [1], noise_modulation):
        W.append((math.sqrt(6 / (c + next_c)) * _real_noise(rng, (c, m)) + noise * _noise_vector(rng, (c, m), is_real)).astype(complex))

    def delta_func(n, m, c):
        if m > n:
            return 1
        if n % m == 0:
            return n // m
        return c

    connections = []
    for n, m, c in zip([N] + M[:-1], M, C):
        dj = delta_func(n, m, c)
        connections.append(np.array([[(j * dj + i) % n for j in range(m)] for i in range(c)], dtype=np.uint32))

    if isinstance(final_weights, (float, int)):
        final_weights = final_weights * np.ones(M[-1])
    assert len(final_weights) == M[-1]
    return DeepSpec(num_sites, a, b, connections, W, np.asarray(final_weights, dtype=complex), 0.0)


def cnn_spec(L, layers, initial_value=(0.01 + 1j * math.pi / 4), noise=1e-4, final_factor=10,
             symmetry_classes=None, real=False, seed=0):
    """pyANNonGPU/new_convolutional_network.py:30-82 with an explicit seed. L is padded to 3 dims (leading 1s),
    matching the bound type PsiCNN_t<3> (include/quantum_state/PsiCNN.hpp:381)."""
    rng = np.random.default_rng(seed)
    L = list(L)
    layers = [(nc, list(conn)) for nc, conn in layers]
    while len(L) < 3:
        L = [1] + L
        layers = [(nc, [1] + conn) for nc, conn in layers]
    num_channels_list = np.array([nc for nc, _ in layers], dtype=np.uint32)
    connectivity_list = np.array([conn for _, conn in layers], dtype=np.uint32)
    if symmetry_classes is None:
        symmetry_classes = np.zeros(_prod(L), dtype=np.uint32)
    num_symmetry_classes = len(set(int(s) for s in symmetry_classes))

    params = []
    for layer, (num_channels, nd_conn) in enumerate(layers):
        for c, l in zip(nd_conn, L):
            assert c <= l
        connectivity = _prod(nd_conn)
        num_prev = layers[layer - 1][0] if layer > 0 else 1
        for _ in range(num_channels * num_prev):
            for _ in range(num_symmetry_classes):
                link = (noise * _noise_vector(rng, connectivity, real)).astype(complex)
                if layer == 0:
                    link[connectivity // 2] = complex(initial_value).real if real else initial_value
                else:
                    link += math.sqrt(6 / (connectivity * num_prev + connectivity * num_channels)) * _real_noise(rng, connectivity)
                params += list(link)
    return CNNSpec(tuple(L), num_channels_list, connectivity_list, np.asarray(symmetry_classes, dtype=np.uint32),
                   np.array(params, dtype=complex), float(final_factor), 0.0)


# ---------------------------------------------------------------------------- BASELINE.json configurations

def config_C1():
    """PsiRBM alpha=2, N=16, TFIM ring, ExactSummation."""
    return rbm_spec(16, 32, noise=1e-2, final_weight=10, seed=1234), tfim(16, ring_bonds(16))


def config_C2(N=64, alpha=4):
    """PsiRBM alpha=4, N=64, Heisenberg ring, 8192 chains."""
    return rbm_spec(N, alpha * N, noise=0.02 / math.sqrt(alpha), final_weight=1, seed=1234), heisenberg(N, ring_bonds(N))


def config_C3():
    """PsiCNN 10x10, 3 layers x 3 channels, 3x3 kernels, J1 Heisenberg, 32768 chains."""
    return cnn_spec([10, 10], [(3, [3, 3])] * 3, noise=1e-2, final_factor=1, seed=1236), heisenberg(100, square_lattice_bonds(10, 10))


def config_C4():
    """PsiDeep 64 -> 64 -> 64 on the 8x8 TFIM."""
    return (deep_spec(64, 64, [64, 64], [64, 64], noise=1e-3, a=0, final_weights=1, seed=1235),
            tfim(64, square_lattice_bonds(8, 8)))


def config_C5(N=200, alpha=8):
    """PsiRBM alpha=8, N=200, Heisenberg ring, 131072 chains over 8 GPUs."""
    return rbm_spec(N, alpha * N, noise=0.02 / math.sqrt(alpha), final_weight=1, seed=1234), heisenberg(N, ring_bonds(N))
