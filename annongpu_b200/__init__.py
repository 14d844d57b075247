"""annongpu_b200 — B200-native (sm_100a) variational-Monte-Carlo hot path of heikoburau/ANNonGPU.

Hand-written CUDA kernels behind a C ABI (``include/angpu.h``, ``libangpu.so``), with this package as the
host-side mirror of the reference's Python binding surface (``pyANNonGPU``).  GPU-only: no CPU fallback.
"""
from . import factories
from .api import *  # noqa: F401,F403
from .api import set_allreduce, activation_derivative
from .factories import PauliSum, heisenberg, tfim, ring_bonds, square_lattice_bonds


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
