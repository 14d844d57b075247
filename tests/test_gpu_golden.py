"""The CUDA path against the golden vectors generated from the compiled, unmodified reference."""
import numpy as np
import pytest

from helpers import GpuAdapter, classical_zoo, hsd_cases, kl_cases, make_classical, make_psi, zoo
from test_oracle_pinned import (GOLDEN, check_against_golden, check_hsd_against_golden, check_kl_against_golden,
                                check_wref_against_golden)

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(zoo()))
def test_gpu_matches_reference_golden(gpu, name):
    spec, H, N = zoo()[name]
    ad = GpuAdapter(gpu)
    check_against_golden(ad, name, make_psi(gpu, spec), H, N, ad.ExactSummation)


@pytest.mark.parametrize("name", sorted(classical_zoo()))
def test_gpu_matches_reference_golden_classical(gpu, name):
    N, order, Hl, pr, ref_spec, lp, H = classical_zoo()[name]
    ad = GpuAdapter(gpu)
    psi = make_classical(gpu, N, order, Hl, pr, ref_spec, lp)
    check_against_golden(ad, name, psi, H, N, ad.ExactSummation)
    check_wref_against_golden(ad, name, psi, H, N, ad.ExactSummation)


@pytest.mark.parametrize("name", sorted(hsd_cases()))
def test_gpu_hilbert_space_distance_matches_reference_golden(gpu, name):
    ad = GpuAdapter(gpu)
    check_hsd_against_golden(ad, name, ad.ExactSummation)


@pytest.mark.parametrize("name", sorted(kl_cases()))
def test_gpu_kullback_leibler_matches_reference_golden(gpu, name):
    ad = GpuAdapter(gpu)
    check_kl_against_golden(ad, name, ad.ExactSummation)


def test_gpu_primitives_match_golden(gpu):
    a, b, c = GOLDEN["pauli/a"], GOLDEN["pauli/b"], GOLDEN["pauli/conf"]
    for i in range(len(a)):
        coeff, out = gpu.pauli_apply(int(a[i]), int(b[i]), int(c[i]), 64)
        assert coeff == complex(GOLDEN["pauli/coeff"][i])
        assert out.configuration == int(GOLDEN["pauli/conf_out"][i])
    for layer in (0, 1, 2):
        for z, lc, th in zip(GOLDEN["act/z"], GOLDEN[f"act/lc{layer}"], GOLDEN[f"act/th{layer}"]):
            assert abs(gpu.activation_function(z, layer) - lc) <= 1e-14 * max(1, abs(lc))
            assert abs(gpu.activation_derivative(z, layer) - th) <= 1e-14 * max(1, abs(th))
