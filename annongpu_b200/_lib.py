"""ctypes loader for ``libangpu.so`` (the C ABI declared in ``include/angpu.h``).

There is no CPU fallback: if the CUDA library is missing, importing the package fails loudly.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libangpu.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "angpu.h")


class AngpuError(RuntimeError):
    """Raised when a libangpu call fails (the reference raises RuntimeError from CUDA_CHECK, include/types.h:139-145)."""


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(or `make -C annongpu_b200/csrc`). annongpu_b200 has no CPU fallback."
    )

lib = C.CDLL(LIB_PATH)

vp, u32, u64, ull, dbl, i32 = C.c_void_p, C.c_uint, C.c_uint64, C.c_ulonglong, C.c_double, C.c_int
ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_ulonglong, C.c_void_p)

_SIGNATURES = {
    "angpu_init": [i32],
    "angpu_set_stream": [vp],
    "angpu_synchronize": [],
    "angpu_profiler_start": [],
    "angpu_profiler_stop": [],
    "angpu_set_allreduce": [ALLREDUCE_FN, vp],
    "angpu_comm_unique_id": [vp],
    "angpu_comm_init": [vp, i32, i32],
    "angpu_comm_destroy": [],
    "angpu_comm_rank": [vp, vp],
    "angpu_spins_enumerate": [u64, u32, vp],
    "angpu_pauli_apply": [vp, vp, vp, u32, vp, vp],
    "angpu_activation": [vp, u32, vp, vp],
    "angpu_operator_create": [u32, vp, vp, vp, u32, vp],
    "angpu_operator_destroy": [vp],
    "angpu_operator_num_strings": [vp, vp],
    "angpu_rbm_create": [u32, u32, vp, vp, vp, vp],
    "angpu_deep_create": [u32, u32, vp, u32, vp, vp, vp, vp, vp, vp, vp, vp],
    "angpu_cnn_create": [vp, u32, vp, vp, vp, vp, u32, dbl, vp, vp],
    "angpu_classical_create": [u32, u32, u32, vp, vp, u32, vp, vp, vp],
    "angpu_psi_copy": [vp, vp],
    "angpu_psi_destroy": [vp],
    "angpu_psi_kind": [vp, vp],
    "angpu_psi_num_sites": [vp, vp],
    "angpu_psi_num_params": [vp, vp],
    "angpu_psi_get_params": [vp, vp],
    "angpu_psi_set_params": [vp, vp],
    "angpu_psi_get_log_prefactor": [vp, vp],
    "angpu_psi_set_log_prefactor": [vp, vp],
    "angpu_es_create": [u32, vp],
    "angpu_mc_create": [ull, u32, u32, u32, u64, vp],
    "angpu_es_paulis_create": [u32, vp],
    "angpu_mc_paulis_create": [ull, u32, u32, u32, u64, vp],
    "angpu_ensemble_copy": [vp, vp],
    "angpu_ensemble_destroy": [vp],
    "angpu_ensemble_num_steps": [vp, vp],
    "angpu_ensemble_local_steps": [vp, vp],
    "angpu_ensemble_set_shard": [vp, u32, u32],
    "angpu_mc_acceptance": [vp, vp],
    "angpu_mc_counters": [vp, vp],
    "angpu_mc_get_call_index": [vp, vp],
    "angpu_mc_set_call_index": [vp, u32],
    "angpu_ensemble_sample": [vp, vp, vp, vp],
    "angpu_log_psi_s": [vp, vp, vp],
    "angpu_psi_O_k": [vp, vp, vp],
    "angpu_log_psi_vector": [vp, vp, vp],
    "angpu_psi_vector": [vp, vp, vp],
    "angpu_log_psi_mean": [vp, vp, vp],
    "angpu_psi_norm": [vp, vp, vp],
    "angpu_psi_O_k_vector": [vp, vp, vp],
    "angpu_apply_operator": [vp, vp, vp, vp],
    "angpu_local_energies": [vp, vp, vp, ull, vp, vp],
    "angpu_expval_create": [vp],
    "angpu_expval_destroy": [vp],
    "angpu_expectation": [vp, vp, vp, vp, vp],
    "angpu_expectation_many": [vp, u32, vp, vp, vp, vp],
    "angpu_expectation_reweighted": [vp, vp, vp, vp, vp, vp],
    "angpu_exp_sigma_z": [vp, vp, vp, vp, vp],
    "angpu_fluctuation": [vp, vp, vp, vp, vp, vp],
    "angpu_gradient": [vp, vp, vp, vp, vp, vp],
    "angpu_tdvp_create": [u32, vp],
    "angpu_tdvp_destroy": [vp],
    "angpu_tdvp_eval": [vp, vp, vp, vp],
    "angpu_tdvp_eval_F": [vp, vp, vp, vp],
    "angpu_tdvp_eval_tol": [vp, vp, vp, vp, dbl],
    "angpu_kl_create": [u32, vp],
    "angpu_kl_destroy": [vp],
    "angpu_kl_set_log_psi_scale": [vp, dbl],
    "angpu_kl_get_state": [vp, vp],
    "angpu_kl_value": [vp, vp, vp, vp, dbl, vp],
    "angpu_kl_gradient": [vp, vp, vp, vp, dbl, dbl, vp, vp],
    "angpu_kl_gradient_with_noise": [vp, vp, vp, vp, dbl, dbl, vp, vp, vp],
    "angpu_hsd_create": [u32, vp],
    "angpu_hsd_destroy": [vp],
    "angpu_hsd_distance": [vp, vp, vp, vp, C.c_int, vp, vp],
    "angpu_hsd_gradient": [vp, vp, vp, vp, C.c_int, vp, C.c_float, vp, vp],
    "angpu_tdvp_eval_reweighted": [vp, vp, vp, vp, vp],
    "angpu_tdvp_get_S": [vp, vp],
    "angpu_tdvp_get_F": [vp, vp],
    "angpu_tdvp_get_O_k": [vp, vp],
    "angpu_tdvp_get_scalars": [vp, vp],
    "angpu_tdvp_num_local_samples": [vp, vp],
    "angpu_tdvp_get_O_k_samples": [vp, vp],
    "angpu_tdvp_get_weights": [vp, vp],
    "angpu_tdvp_get_E_local_samples": [vp, vp],
    "angpu_tdvp_S_dot_vector": [vp, vp, vp],
    "angpu_tdvp_solve_cg": [vp, dbl, u32, dbl, dbl, vp, vp, vp, vp],
    "angpu_tdvp_solve_dense": [vp, dbl, dbl, vp, vp],
    "angpu_tdvp_apply_update": [vp, vp, vp],
    "angpu_hpd_solve": [u32, vp, vp, vp],
    "angpu_tdvp_build_S_tensorcore": [vp],
    "angpu_tdvp_set_profile": [vp, i32],
    "angpu_tdvp_set_tensorcore_products": [vp, i32],
    "angpu_tdvp_phase_ms": [vp, vp],
    "angpu_measure_fp64_tflops": [vp],
}

for _name, _args in _SIGNATURES.items():
    _f = getattr(lib, _name)
    _f.argtypes = _args
    _f.restype = i32
lib.angpu_last_error.restype = C.c_char_p
lib.angpu_last_error.argtypes = []
lib.angpu_launch_count.restype = ull
lib.angpu_launch_count.argtypes = [i32]

EXPORTED = sorted(list(_SIGNATURES) + ["angpu_last_error", "angpu_launch_count"])


# an exception raised inside the Python all-reduce callback cannot cross the C frames: the trampoline (api.set_allreduce)
# parks it here and returns non-zero, the failing entry point returns an error, and check() re-raises the original
pending_callback_error = []


def check(status):
    if status != 0:
        msg = lib.angpu_last_error().decode("utf-8", "replace")
        if pending_callback_error:
            exc = pending_callback_error.pop()
            pending_callback_error.clear()
            raise AngpuError(msg) from exc
        raise AngpuError(msg)


def call(name, *args):
    check(getattr(lib, name)(*args))
