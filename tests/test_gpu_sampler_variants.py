"""PsiRBM sampler instantiations: the C2-shaped warp kernel <K=8, WORDS=1> against the oracle's trajectory, and the
opt-in fp32-screened sampler (csrc/rbm_sampler.cuh, ANGPU_MC_SCREEN=1) against the all-fp64 one, chain by chain."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from annongpu_b200 import factories as F
from helpers import make_op, make_psi, rel_err

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c2_shape_chains_identical_to_oracle(gpu, port):
    """N = 64, M = 256 (BASELINE C2): k_mc_rbm<8, 1, true> draws the oracle's configurations, call after call."""
    spec, _ = F.config_C2()
    pg, pp = make_psi(gpu, spec), make_psi(port, spec)
    chains, per_chain = 12, 2
    mg = gpu.MonteCarloSpins(chains * per_chain, 1, 2, chains, True, seed=31337)
    mp = port.MonteCarlo(chains * per_chain, 1, 2, chains, seed=31337)
    for _ in range(2):
        cg, lg = mg.sample(pg)
        cp_, lp_ = mp.sample(pp)
        assert np.array_equal(cg, cp_)
        assert rel_err(lg, lp_) <= 1e-9
        assert mg.acceptances == (mp.acceptances, mp.rejections)


CHILD = r"""
import json, sys
import numpy as np
sys.path.insert(0, {root!r})
import annongpu_b200 as A
from annongpu_b200 import factories as F
A.setDevice(0)
out = {{}}
for name, (N, M, noise) in {{"n64m256": (64, 256, None), "n40m100": (40, 100, 0.05), "n100m500": (100, 500, 0.02), "n20m40": (20, 40, 0.1)}}.items():
    spec = F.config_C2()[0] if noise is None else F.rbm_spec(N, M, noise=noise, final_weight=1.0, seed=77)
    psi = spec.build(True)
    mc = A.MonteCarloSpins(512, 2, 3, 256, True, seed=99)
    rows = []
    for call in range(2):
        c, lp = mc.sample(psi)
        rows.append([c.tolist(), lp.real.tolist(), lp.imag.tolist(), list(mc.acceptances), mc.exact_decisions])
    out[name] = rows
print(json.dumps(out))
"""


def _run(screen):
    env = {**os.environ, "ANGPU_MC_SCREEN": "1" if screen else "0"}
    p = subprocess.run([sys.executable, "-c", CHILD.format(root=ROOT)], env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    return json.loads(p.stdout.strip().splitlines()[-1])


def test_fp32_screened_sampler_reproduces_the_fp64_chains(gpu):
    """Every accept/reject decision of the screened sampler is the fp64 one: identical configurations and acceptance
    counts, log psi equal to rounding, and only a small fraction of the proposals needs the fp64 evaluation."""
    ref, scr = _run(False), _run(True)
    for name in ref:
        for (c0, re0, im0, acc0, ex0), (c1, re1, im1, acc1, ex1) in zip(ref[name], scr[name]):
            assert c0 == c1, name
            assert acc0 == acc1, name
            assert ex0 == 0 and 0 <= ex1 <= 0.02 * sum(acc1), (name, ex1)
            lp0, lp1 = np.array(re0) + 1j * np.array(im0), np.array(re1) + 1j * np.array(im1)
            assert np.abs(lp0 - lp1).max() <= 1e-11, name


@pytest.mark.parametrize("N,M,model", [(64, 256, "heisenberg"), (40, 100, "heisenberg"), (24, 72, "tfim"), (64, 256, "tfim")])
def test_tile_local_energy_kernel_matches_the_oracle(gpu, port, N, M, model):
    """k_eloc_rbm_tile (W rows of a flip group in registers, a tile of samples per block) is taken for 64 < M <= 256, <= 2 flips per group
    and >= 1184 samples: E_loc per configuration against the port, on random configurations (ragged last tile included)."""
    spec = F.rbm_spec(N, M, noise=0.03, final_weight=1.3, seed=21)
    H = F.heisenberg(N, F.ring_bonds(N)) if model == "heisenberg" else F.tfim(N, F.ring_bonds(N), J=1.0, h=0.7)
    psi_g, psi_p, op_g, op_p = make_psi(gpu, spec), make_psi(port, spec), make_op(gpu, H), make_op(port, H)
    rng = np.random.default_rng(8)
    ns = 1184 + 777
    confs = rng.integers(0, 1 << min(N, 62), size=ns, dtype=np.uint64).reshape(ns, 1)
    if N > 62:
        confs ^= rng.integers(0, 2, size=(ns, 1), dtype=np.uint64) << np.uint64(63)
    lp_g, el_g = gpu.local_energies(psi_g, op_g, confs)
    lp_p, el_p, _ = port.eval_samples(psi_p, op_p, confs)
    np.testing.assert_allclose(lp_g, lp_p, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(el_g, el_p, rtol=1e-10, atol=1e-10)
