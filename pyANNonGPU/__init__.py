"""Drop-in import name of the reference's Python package (``/root/reference/pyANNonGPU/__init__.py:1`` is
``from ._pyANNonGPU import *``): ``import pyANNonGPU`` resolves to the B200-native implementation in ``annongpu_b200``
(the C ABI of ``include/angpu.h`` behind the same class / function names), so scripts and tests written against the
reference run unchanged on the hot path.  GPU-only: every ``gpu`` argument must be True."""
from annongpu_b200 import *                                   # noqa: F401,F403
from annongpu_b200 import (                                   # noqa: F401
    new_RBM, new_deep_neural_network, new_convolutional_network, new_classical_network, PauliSum,
    sigma_x, sigma_y, sigma_z, set_allreduce, factories, distributed, json_numpy,
)
from annongpu_b200.api import (                               # noqa: F401
    PsiRBM, PsiDeep, PsiCNN, PsiClassicalFP_1, PsiClassicalFP_2, PsiClassicalANN_1, PsiClassicalANN_2, PsiFullyPolarized,
    Operator, Spins, MonteCarloSpins, ExactSummationSpins, MonteCarloPaulis, ExactSummationPaulis, ExpectationValue, TDVP, HilbertSpaceDistance, KullbackLeibler,
    log_psi_s, psi_O_k, psi_O_k_vector, log_psi, psi_vector, log_psi_vector, apply_operator, activation_function,
    setDevice, start_profiling, stop_profiling,
)
