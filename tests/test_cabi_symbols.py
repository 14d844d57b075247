"""The C-ABI library loads on a CPU-only box and exports every symbol include/angpu.h declares (no compute calls)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "angpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(angpu_[a-z0-9_]+)\s*\(", text, flags=re.I)) - {"angpu_allreduce_fn"})


def test_header_symbols_are_exported():
    lib = ctypes.CDLL(os.path.join(ROOT, "annongpu_b200", "libangpu.so"))
    names = declared_symbols()
    assert len(names) >= 60
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_python_binding_covers_the_header():
    from annongpu_b200 import _lib
    assert sorted(_lib.EXPORTED) == declared_symbols()


def test_fails_loudly_without_a_gpu_or_on_bad_arguments():
    """No CPU fallback: with no device the first call raises; with a device, bad arguments raise (never a silent result)."""
    import numpy as np
    import pytest
    import annongpu_b200 as A
    with pytest.raises((A._lib.AngpuError, ValueError)):
        A.PsiRBM(np.zeros((4, 8), dtype=complex), 1.0, 0.0, gpu=False)      # the reference's host path is not provided
    with pytest.raises(A._lib.AngpuError):
        # N = 300 exceeds the 256-site mask width -> rejected before any CUDA call; on a CPU-only box the CUDA init fails first
        A.PsiRBM(np.zeros((300, 300), dtype=complex), 1.0, 0.0, True)


def test_reference_binding_surface_names():
    """Names of pyANNonGPU/main.cpp.template:65-541 that belong to the hot path."""
    import annongpu_b200 as A
    for name in ["PsiRBM", "PsiDeep", "PsiCNN", "PsiClassicalFP_1", "PsiClassicalFP_2", "PsiClassicalANN_1", "PsiClassicalANN_2",
                 "PsiFullyPolarized", "Operator", "Spins", "MonteCarloSpins", "ExactSummationSpins", "ExpectationValue", "TDVP",
                 "log_psi_s", "psi_O_k", "psi_O_k_vector", "log_psi", "psi_vector", "log_psi_vector", "apply_operator",
                 "activation_function", "setDevice", "start_profiling", "stop_profiling",
                 "new_RBM", "new_deep_neural_network", "new_convolutional_network"]:
        assert hasattr(A, name), name
    for cls, members in [(A.TDVP, ["eval", "eval_F", "S_dot_vector", "S_matrix", "F_vector", "O_k_vector", "var_H", "total_weight",
                                   "O_k_samples", "E_local_samples", "solve", "solve_cg"]),
                         (A.ExpectationValue, ["__call__", "fluctuation", "gradient"]),
                         (A.MonteCarloSpins, ["num_steps", "acceptance_rate"]),
                         (A.PsiRBM, ["copy", "num_params", "params", "W", "norm", "normalize", "calibrate", "vector", "log_prefactor"]),
                         (A.PsiDeep, ["a", "b", "W", "final_weights", "input_weights", "params"]),
                         (A.PsiCNN, ["init_gradient", "params", "channel_link"])]:
        for m in members:
            assert hasattr(cls, m), (cls.__name__, m)
