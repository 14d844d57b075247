"""Generates tests/golden/ref_golden_paulis.npz from the COMPILED, UNMODIFIED reference built with -DENABLE_PAULIS
(oracle/_ref/liboracle_ref_paulis.so, `make -C oracle ref_paulis`): the Pauli-string (density-matrix) basis of SURVEY.md §8f
rank 3 -- PsiDeep with N = 3 num_sites input units over ExactSummationPaulis.  Run in the build container:
    python tests/golden/make_golden_paulis.py
Inputs are the seeded specs of tests/helpers.pauli_zoo()."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from helpers import pauli_zoo                      # noqa: E402
from oracle import ref_paulis as RP                # noqa: E402


def main():
    out = {}
    for name, (spec, H, ns) in pauli_zoo().items():
        psi = RP.PsiDeep(spec.num_sites, spec.input_weights, spec.biases, spec.connections, spec.weights, spec.final_weights,
                         spec.log_prefactor)
        c, a, b = H.arrays(1)
        op = RP.Operator(c, a[:, 0], b[:, 0])
        es = RP.ExactSummationPaulis(ns)
        assert es.num_steps == 4 ** ns
        lpv = RP.log_psi_vector(psi, es)
        out[f"{name}/log_psi_vector"] = lpv
        out[f"{name}/E"] = RP.expectation(op, psi, es)
        f, m = RP.fluctuation(op, psi, es)
        out[f"{name}/fluctuation"] = f
        g, e = RP.gradient(op, psi, es)
        out[f"{name}/gradient"] = g
        t = RP.tdvp_eval(op, psi, es)
        for k in ("S", "F", "O_k", "E", "E2", "var_H"):
            out[f"{name}/tdvp_{k}"] = t[k]
        probes = [RP.enumerate(i) for i in (0, 1, 7, 4 ** ns - 1, (4 ** ns) // 3)]
        out[f"{name}/probe_a"] = np.array([p[0] for p in probes], dtype=np.uint64)
        out[f"{name}/probe_b"] = np.array([p[1] for p in probes], dtype=np.uint64)
        out[f"{name}/probe_log_psi"] = np.array([RP.log_psi_s(psi, *p) for p in probes])
        out[f"{name}/probe_O_k"] = np.array([RP.psi_O_k(psi, *p) for p in probes])
    # primitives (bit-exact): enumerate, Pauli o Pauli, network units
    idx = np.arange(0, 4 ** 6, 37, dtype=np.uint32)
    en = [RP.enumerate(int(i)) for i in idx]
    out["enum/index"], out["enum/a"], out["enum/b"] = idx, np.array([e[0] for e in en], dtype=np.uint64), np.array([e[1] for e in en], dtype=np.uint64)
    rng = np.random.default_rng(11)
    Pa, Pb, xa, xb = (rng.integers(0, 1 << 63, size=96, dtype=np.uint64) for _ in range(4))
    res = [RP.pauli_mul(*v) for v in zip(Pa, Pb, xa, xb)]
    out["mul/Pa"], out["mul/Pb"], out["mul/xa"], out["mul/xb"] = Pa, Pb, xa, xb
    out["mul/coeff"] = np.array([r[0] for r in res])
    out["mul/a"], out["mul/b"] = np.array([r[1] for r in res], dtype=np.uint64), np.array([r[2] for r in res], dtype=np.uint64)
    ua, ub = int(Pa[0]) & 0xFFFFF, int(Pb[0]) & 0xFFFFF
    out["units/a"], out["units/b"] = np.uint64(ua), np.uint64(ub)
    out["units/values"] = np.array([RP.network_unit_at(ua, ub, i) for i in range(60)], dtype=np.int32)
    np.savez_compressed(os.path.join(HERE, "ref_golden_paulis.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
