"""Shared test helpers: build the same model in the product (annongpu_b200), the CPU oracle (oracle.port_oracle)
and the compiled reference (oracle.ref_oracle) from one spec of annongpu_b200.factories."""
import numpy as np

from annongpu_b200 import factories as F


def make_psi(mod, spec, log_prefactor=None):
    """mod: annongpu_b200 | oracle.port_oracle | oracle.ref_oracle."""
    lp = spec.log_prefactor if log_prefactor is None else log_prefactor
    if isinstance(spec, F.RBMSpec):
        return mod.PsiRBM(spec.W, spec.final_weight, lp)
    if isinstance(spec, F.DeepSpec):
        return mod.PsiDeep(spec.num_sites, spec.input_weights, spec.biases, spec.connections, spec.weights, spec.final_weights, lp)
    if isinstance(spec, F.CNNSpec):
        return mod.PsiCNN(spec.extent, spec.num_channels_list, spec.connectivity_list, spec.symmetry_classes, spec.params,
                          spec.final_factor, lp)
    raise TypeError(spec)


def make_op(mod, H, words=None):
    name = mod.__name__
    words = words or H.words
    c, a, b = H.arrays(words)
    if name.endswith("ref_oracle"):
        assert words == 1
        return mod.Operator(c, a[:, 0], b[:, 0])
    if name.endswith("port_oracle"):
        return mod.Operator(c, a, b, words)
    return mod.Operator(H)


def make_classical(mod, num_sites, order, local_ops, params, ref_spec, log_prefactor):
    """local_ops: list of PauliSum; ref_spec: CNNSpec or None."""
    name = mod.__name__
    ops = [make_op(mod, h) for h in local_ops]
    psi_ref = make_psi(mod, ref_spec) if ref_spec is not None else None
    if psi_ref is not None and hasattr(psi_ref, "init_gradient"):
        # the reference copies psi_ref (with its per-sample angle scratch) into PsiClassical: size it first
        psi_ref.init_gradient(1 << num_sites)
    if name.endswith("oracle"):
        psi = mod.PsiClassical(num_sites, order, ops, params, psi_ref, log_prefactor)
    else:
        cls = {(1, False): mod.PsiClassicalFP_1, (2, False): mod.PsiClassicalFP_2,
               (1, True): mod.PsiClassicalANN_1, (2, True): mod.PsiClassicalANN_2}[(order, ref_spec is not None)]
        psi = cls(num_sites, ops, params, psi_ref if psi_ref is not None else mod.PsiFullyPolarized(num_sites, log_prefactor),
                  log_prefactor)
    psi._keepalive = (ops, psi_ref)
    return psi


def rel_err(a, b):
    a, b = np.asarray(a), np.asarray(b)
    scale = max(np.abs(b).max() if b.size else 0.0, 1e-300)
    return float(np.abs(a - b).max() / scale) if a.size else 0.0


# ---------------------------------------------------------------- small model zoo used by the parity tests

def zoo():
    """name -> (spec or classical-builder args, Hamiltonian PauliSum, num_sites). Sizes the oracle finishes in seconds."""
    z = {}
    z["rbm10"] = (F.rbm_spec(10, 20, noise=5e-2, final_weight=10, seed=1), F.heisenberg(10, F.ring_bonds(10)), 10)
    z["rbm8_cfw"] = (F.rbm_spec(8, 24, noise=5e-2, final_weight=2.0 - 0.5j, seed=11), F.tfim(8, F.ring_bonds(8), h=0.8), 8)
    z["rbm12_m40"] = (F.rbm_spec(12, 40, noise=3e-2, final_weight=3, seed=12), F.heisenberg(12, F.ring_bonds(12)), 12)
    z["deep1"] = (F.deep_spec(6, 6, [12], [6], noise=1e-2, a=0, final_weights=2, seed=4), F.heisenberg(6, F.ring_bonds(6)), 6)
    z["deep2"] = (F.deep_spec(8, 8, [16, 8], [4, 8], noise=1e-2, a=0.1, final_weights=3, seed=2), F.tfim(8, F.ring_bonds(8), h=0.7), 8)
    z["deep3"] = (F.deep_spec(6, 6, [18, 9, 3], [3, 2, 3], noise=1e-2, a=0, final_weights=2, seed=3), F.heisenberg(6, F.ring_bonds(6)), 6)
    z["cnn"] = (F.cnn_spec([2, 2, 3], [(2, [2, 2, 2]), (3, [1, 2, 2])], noise=1e-1, final_factor=2, seed=5),
                F.heisenberg(12, F.ring_bonds(12)), 12)
    z["cnn_sym"] = (F.cnn_spec([3, 4], [(2, [2, 3]), (2, [3, 2]), (1, [2, 2])], noise=1e-1, final_factor=2, seed=6,
                               symmetry_classes=np.array([0, 1] * 6)), F.tfim(12, F.square_lattice_bonds(3, 4)), 12)
    return z


def wide_zoo():
    """Shapes beyond the reference's compiled-in limits (M > 128): checked against the port only."""
    return {"rbm_wide": (F.rbm_spec(12, 600, noise=2e-2, final_weight=0.5, seed=13), F.heisenberg(12, F.ring_bonds(12)), 12)}


def classical_zoo():
    N = 6
    Hl = [F.PauliSum(N).add(1.0, {i: "Z", (i + 1) % N: "Z"}).add(0.3, {i: "X"}).add(0.2j, {i: "Y", (i + 2) % N: "Z"}) for i in range(N)]
    rng = np.random.default_rng(7)
    pr = 0.1 * (rng.normal(size=N) + 1j * rng.normal(size=N))
    ref_spec = F.cnn_spec([1, 2, 3], [(2, [1, 2, 2])], noise=1e-1, final_factor=2, seed=8)
    H = F.heisenberg(N, F.ring_bonds(N))
    return {
        "clfp1": (N, 1, Hl, pr, None, -1.0, H),
        "clfp2": (N, 2, Hl, pr, None, -1.0, H),
        "clann1": (N, 1, Hl, pr, ref_spec, 0.0, H),
        "clann2": (N, 2, Hl, pr, ref_spec, 0.0, H),
    }


def hsd_cases():
    """name -> (spec of psi, spec of psi_prime, operator PauliSum, is_unitary, num_sites) for HilbertSpaceDistance."""
    z = zoo()
    deep, Hd, Nd = z["deep2"]
    deep_p = F.deep_spec(8, 8, [16, 8], [4, 8], noise=2e-2, a=0.05, final_weights=3, seed=22)
    cnn, Hc, Nc = z["cnn"]
    cnn_p = F.cnn_spec([2, 2, 3], [(2, [2, 2, 2]), (3, [1, 2, 2])], noise=1.2e-1, final_factor=2, seed=55)
    return {
        "deep_unitary": (deep, deep_p, F.propagator(Hd, 0.05), True, Nd),
        "deep_exp": (deep, deep_p, F.scaled(Hd, -0.05j), False, Nd),
        "cnn_unitary": (cnn, cnn_p, F.propagator(Hc, 0.02), True, Nc),
        "cnn_exp": (cnn, cnn_p, F.scaled(Hc, -0.001j), False, Nc),
    }


def kl_cases():
    """name -> (classical zoo entry name, spec of psi_prime) for KullbackLeibler (psi: PsiClassical, psi_prime: Deep | CNN)."""
    deep_p = F.deep_spec(6, 6, [12, 6], [6, 12], noise=5e-2, a=0.05, final_weights=2, seed=33)
    cnn_p = F.cnn_spec([1, 2, 3], [(2, [1, 2, 2]), (2, [1, 2, 2])], noise=2e-1, final_factor=2, seed=34)
    return {"clfp1_deep": ("clfp1", deep_p), "clfp2_cnn": ("clfp2", cnn_p), "clann1_cnn": ("clann1", cnn_p), "clann2_deep": ("clann2", deep_p)}


def kl_sequence(mod, name, ensemble_cls):
    """Three consecutive calls on one KullbackLeibler object (the mean deviation is carried from call to call):
    value(threshold 0), gradient(nu 0.5, threshold 1e-3), gradient_with_noise(nu 1, threshold 0)."""
    cname, pspec = kl_cases()[name]
    N, order, Hl, pr, ref_spec, lp, H = classical_zoo()[cname]
    psi, psi_prime = make_classical(mod, N, order, Hl, pr, ref_spec, lp), make_psi(mod, pspec)
    if hasattr(psi_prime, "init_gradient"):
        psi_prime.init_gradient(1 << N)
    es = ensemble_cls(N)
    kl = mod.KullbackLeibler(psi_prime.num_params) if mod.__name__.endswith("oracle") else mod.KullbackLeibler(psi_prime.num_params, True)
    kl.log_psi_scale = 0.9
    v1 = kl(psi, psi_prime, es, 0.0)
    g, v2 = kl.gradient(psi, psi_prime, es, 0.5, 1e-3)
    gn, noise, v3 = kl.gradient_with_noise(psi, psi_prime, es, 1.0, 0.0)
    return dict(v1=v1, v2=v2, v3=v3, g=g, gn=gn, noise=noise, total_weight=kl.total_weight, mean_deviation=kl.mean_deviation)


class GpuAdapter:
    """Presents annongpu_b200 with the oracle modules' function names, so one checker serves both."""
    __name__ = "annongpu_b200"

    def __init__(self, A):
        self.A = A
        self.ev = A.ExpectationValue(True)
        for k in ("PsiRBM", "PsiDeep", "PsiCNN", "PsiClassicalFP_1", "PsiClassicalFP_2", "PsiClassicalANN_1", "PsiClassicalANN_2",
                  "PsiFullyPolarized", "Operator", "log_psi", "psi_vector", "apply_operator", "KullbackLeibler"):
            setattr(self, k, getattr(A, k))

    def ExactSummation(self, N):
        return self.A.ExactSummationSpins(N, True)

    def psi_norm(self, psi, es):
        return psi.norm(es)

    def expectation(self, op, psi, ens):
        return self.ev(op, psi, ens)

    def fluctuation(self, op, psi, ens):
        return self.ev.fluctuation(op, psi, ens)

    def gradient(self, op, psi, ens):
        return self.ev.gradient(op, psi, ens)

    def exp_sigma_z(self, op, psi, ens):
        return self.ev.exp_sigma_z(op, psi, ens)

    def hilbert_space_distance(self, psi, psi_prime, op, is_unitary, ens):
        return self.A.HilbertSpaceDistance(psi_prime.num_params, True)(psi, psi_prime, op, is_unitary, ens)

    def hilbert_space_distance_gradient(self, psi, psi_prime, op, is_unitary, ens, nu):
        return self.A.HilbertSpaceDistance(psi_prime.num_params, True).gradient(psi, psi_prime, op, is_unitary, ens, nu)

    def TDVP(self, P):
        return self.A.TDVP(P, True)

    def log_psi_s(self, psi, conf):
        return self.A.log_psi_s(psi, np.asarray(conf, dtype=np.uint64))

    def psi_O_k(self, psi, conf):
        return self.A.psi_O_k(psi, np.asarray(conf, dtype=np.uint64))


def pauli_zoo():
    """Pauli-string-basis (density-matrix) cases, SURVEY.md §8f rank 3: name -> (DeepSpec with N = 3 num_sites, PauliSum, num_sites).
    The operator has complex coefficients and every Pauli type, so that all branches of PauliString::apply(PauliString) are hit."""
    z = {}
    for name, ns, M, C, seed in (("pdeep3", 3, [9, 3], [3, 3], 31), ("pdeep4", 4, [12], [6], 32), ("pdeep2x2h", 2, [6, 6], [6, 3], 33)):
        spec = F.deep_spec(ns, 3 * ns, M, C, noise=0.3, final_weights=0.7, seed=seed)
        H = F.heisenberg(ns, F.ring_bonds(ns)) if ns > 2 else F.heisenberg(ns, [(0, 1)])
        H = H + (0.3 - 0.2j) * F.sigma_x(0, ns) + (0.1 + 0.4j) * F.sigma_y(ns - 1, ns) * F.sigma_z(0, ns) + 0.25 * F.sigma_z(1, ns) + 0.5
        z[name] = (spec, H, ns)
    return z
