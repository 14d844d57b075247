"""annongpu_b200 — B200-native (sm_100a) variational-Monte-Carlo hot path of heikoburau/ANNonGPU.

Hand-written CUDA kernels behind a C ABI (``include/angpu.h``, ``libangpu.so``), with this package as the
host-side mirror of the reference's Python binding surface (``pyANNonGPU``).  GPU-only: no CPU fallback.
"""
from . import factories
from .api import *  # noqa: F401,F403
from .api import set_allreduce, activation_derivative
from .factories import PauliSum, heisenberg, tfim, ring_bonds, square_lattice_bonds, sigma_x, sigma_y, sigma_z


def new_RBM(N, M, initial_value=(0.01 + 1j * 3.141592653589793 / 4), noise=1e-4, gpu=True, final_weight=10, seed=None):
    """pyANNonGPU/new_RBM.py:14-29 (+ optional seed)."""
    return factories.rbm_spec(N, M, initial_value, noise, final_weight, seed).build(gpu)


def new_deep_neural_network(num_sites, N, M, C, initial_value=(0.01 + 1j * 3.141592653589793 / 4), a=0, noise=1e-4,
                            gpu=True, noise_modulation="auto", final_weights=10, seed=None):
    """pyANNonGPU/new_neural_network.py:31-131 (+ optional seed)."""
    return factories.deep_spec(num_sites, N, M, C, initial_value, a, noise, noise_modulation, final_weights, seed).build(gpu)


def new_convolutional_network(L, layers, initial_value=(0.01 + 1j * 3.141592653589793 / 4), noise=1e-4, final_factor=10,
                              symmetry_classes=None, real=False, gpu=True, seed=None):
    """pyANNonGPU/new_convolutional_network.py:30-82 (+ optional seed)."""
    return factories.cnn_spec(L, layers, initial_value, noise, final_factor, symmetry_classes, real, seed).build(gpu)


def new_classical_network(num_sites, order, H_local, symmetric=False, distance="max", params=0, psi_ref="fully polarized",
                          use_super_operator=False, gpu=True):
    """pyANNonGPU/new_classical_network.py:6-81 for the spin basis: H_local is a list of Pauli expressions (PauliSum), one
    parameter each; psi_ref "fully polarized" gives PsiClassicalFP_<order>, a PsiCNN gives PsiClassicalANN_<order>.
    (symmetric / SuperOperator variants belong to the Pauli-basis feature set and are rejected.)"""
    import numpy as np
    if order not in (1, 2):
        raise ValueError("order must be 1 or 2")
    if symmetric or use_super_operator:
        raise NotImplementedError("new_classical_network: symmetric / super-operator ansaetze are outside the spin-basis hot path")
    if isinstance(H_local, PauliSum):
        H_local = [H_local]
    ops = [Operator(h, gpu, num_sites=num_sites) for h in H_local]
    if isinstance(params, (int, float)) and params == 0:
        params = np.zeros(len(ops), dtype=complex)
    log_prefactor = float(np.log(1.0 / 2.0 ** (num_sites / 2.0)))
    if isinstance(psi_ref, str):
        ref = PsiFullyPolarized(num_sites, log_prefactor)
        cls = PsiClassicalFP_1 if order == 1 else PsiClassicalFP_2
        return cls(num_sites, ops, params, ref, log_prefactor, gpu)
    cls = PsiClassicalANN_1 if order == 1 else PsiClassicalANN_2
    return cls(num_sites, ops, params, psi_ref, 0.0, gpu)
