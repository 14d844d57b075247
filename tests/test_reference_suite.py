"""The reference's two live tests of the wavefunction layer, restated against ``import pyANNonGPU`` (the drop-in name):
  test/test_Psi.py:30-79    test_psi_CNN_s  log psi(s) of a PsiCNN equals an explicit periodic cross-correlation in numpy
  test/test_Psi.py:259-324  test_O_k        psi_O_k equals central finite differences of log psi(s) in every parameter
with the model the reference's conftest parametrises them with (test/conftest.py:38-47, 63-67):
new_convolutional_network([2, 4, 4], [(2, [2, 4, 4]), (4, [2, 2, 2])], noise=1e-2).  The test files themselves live in
/root/reference, which does not exist on the GPU box -- hence the restatement; tolerances are the reference's."""
import random

import numpy as np
import pytest
from pytest import approx

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def P(gpu):
    import pyANNonGPU
    return pyANNonGPU


def _psi_cnn(P):
    return P.new_convolutional_network([2, 4, 4], [(2, [2, 4, 4]), (4, [2, 2, 2])], noise=1e-2, gpu=True, seed=4321)


def _nd_cross_correlate(inputs, weights):
    """out[x] = sum_k w[k] in[(x + k) mod L]: every shift of the periodic input against the window (test_Psi.py:30-42)."""
    window = tuple(slice(0, w) for w in weights.shape)
    return np.array([np.sum(np.roll(inputs, -np.array(shift), axis=tuple(range(inputs.ndim)))[window] * weights)
                     for shift in np.ndindex(*inputs.shape)]).reshape(inputs.shape)


def test_psi_CNN_s(P):
    random.seed(7)
    psi = _psi_cnn(P)
    N = psi.num_sites
    act = np.vectorize(P.activation_function, excluded=[1])
    for _ in range(10):
        spins = P.Spins(random.randint(0, 2 ** N - 1), 64)
        activations = [spins.array(N)]
        prev_counts = [1] + list(psi.num_channels_list)[:-1]
        for layer, (n_prev, n_ch) in enumerate(zip(prev_counts, psi.num_channels_list)):
            conn = tuple(int(c) for c in psi.connectivity_list[layer])
            activations = np.array([
                act(sum(_nd_cross_correlate(np.asarray(activations[pc]).reshape(tuple(int(e) for e in psi.extent)),
                                            psi.channel_link(layer, pc, ch).reshape(conn)).flatten() for pc in range(n_prev)), layer)
                for ch in range(n_ch)])
        log_psi_ref = psi.log_prefactor + np.sum(activations) * psi.final_factor
        assert P.log_psi_s(psi, spins) == approx(log_psi_ref, 1e-4)


def test_O_k(P):
    random.seed(11)
    psi = _psi_cnn(P)
    eps = 1e-4
    psi_plus = +psi

    def shifted(k, delta):
        params = psi.params
        params[k] += delta
        psi_plus.params = params
        return psi_plus

    for _ in range(10):
        conf = P.Spins.enumerate(random.randint(0, 2 ** psi.num_sites - 1))
        ref = np.array([(P.log_psi_s(shifted(k, eps), conf) - P.log_psi_s(shifted(k, -eps), conf)) / (2 * eps)
                        for k in range(psi.num_params)])
        assert np.allclose(ref, P.psi_O_k(psi, conf), rtol=1e-3, atol=1e-4)


def test_classical_network_json_round_trip(P):
    """pyANNonGPU/PsiClassical.py:7-67 with the native operator encoding: same log psi after to_json -> from_json."""
    h = [P.sigma_z(0) * P.sigma_z(1) + P.sigma_x(0), 0.7 * P.sigma_z(1) * P.sigma_z(2)]
    for order, ref in ((1, "fully polarized"), (2, "fully polarized"), (2, "cnn")):
        psi_ref = ref if ref != "cnn" else P.new_convolutional_network([1, 1, 4], [(2, [1, 1, 2])], noise=5e-2, final_factor=1, seed=5)
        psi = P.new_classical_network(4, order, h, params=np.array([0.1 + 0.2j, -0.3j]), psi_ref=psi_ref, gpu=True)
        doc = psi.to_json()
        assert doc["type"] == "PsiClassical" and doc["order"] == order and len(doc["ansatz"]) == 2
        back = type(psi).from_json(doc, True)
        assert type(back) is type(psi) and back.num_params == psi.num_params
        for idx in (0, 5, 10, 15):
            s = P.Spins.enumerate(idx)
            assert abs(P.log_psi_s(back, s) - P.log_psi_s(psi, s)) <= 1e-13
