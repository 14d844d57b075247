"""Generates tests/golden/ref_golden.npz from the COMPILED, UNMODIFIED reference (oracle/_ref/liboracle_ref.so, built by
oracle/Makefile from /root/reference).  Run in the build container:  python tests/golden/make_golden.py

The reference ships no golden vectors of its own; these fixtures pin the oracle port — and through it
the CUDA path — on boxes where the reference cannot be built.  Inputs are the seeded specs of tests/helpers.py."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from helpers import classical_zoo, hsd_cases, kl_cases, kl_sequence, make_classical, make_op, make_psi, zoo   # noqa: E402
from oracle import ref_oracle as R                                           # noqa: E402

PROBES = [0x2A5, 0x13, 0x3FF, 0x0]


def record(out, name, psi, H, N):
    op = make_op(R, H)
    es = R.ExactSummation(N)
    n = R.psi_norm(psi, es)
    psi.log_prefactor = psi.log_prefactor - np.log(n)
    if hasattr(psi, "init_gradient"):
        psi.init_gradient(1 << N)
    P = psi.num_params
    out[f"{name}/norm"] = n
    out[f"{name}/log_prefactor"] = psi.log_prefactor
    out[f"{name}/E"] = R.expectation(op, psi, es)
    f, m = R.fluctuation(op, psi, es)
    out[f"{name}/fluctuation"] = f
    g, e = R.gradient(op, psi, es)
    out[f"{name}/gradient"] = g
    t = R.TDVP(P)
    t.eval(op, psi, es)
    out[f"{name}/F"], out[f"{name}/Ok"], out[f"{name}/var_H"] = t.F_vector, t.O_k_vector, t.var_H
    S = t.S_matrix
    out[f"{name}/S_diag"] = np.diag(S).copy()
    rng = np.random.default_rng(42)
    v = rng.normal(size=P) + 1j * rng.normal(size=P)
    out[f"{name}/v"], out[f"{name}/Sv_dense"], out[f"{name}/Sv"] = v, S @ v, t.S_dot_vector(v, es)
    mask = (1 << N) - 1
    out[f"{name}/log_psi_s"] = np.array([R.log_psi_s(psi, c & mask) for c in PROBES])
    out[f"{name}/O_k"] = np.array([R.psi_O_k(psi, c & mask) for c in PROBES])
    out[f"{name}/psi_vector_head"] = R.psi_vector(psi, es)[:64]
    out[f"{name}/apply_operator_head"] = R.apply_operator(psi, op, es)[:64]
    out[f"{name}/log_psi_mean"] = R.log_psi(psi, es)
    out[f"{name}/exp_sigma_z"] = R.exp_sigma_z(op, psi, es)


def main():
    out = {}
    for name, (spec, H, N) in zoo().items():
        record(out, name, make_psi(R, spec), H, N)
    for name, (N, order, Hl, pr, ref_spec, lp, H) in classical_zoo().items():
        psi = make_classical(R, N, order, Hl, pr, ref_spec, lp)
        record(out, name, psi, H, N)
        # TDVP::eval(..., true_t) = eval_with_psi_ref: samples from the classical state's reference state
        t = R.TDVP(psi.num_params)
        out[f"{name}/wref/total_weight"] = t.eval_with_psi_ref(make_op(R, H), psi, R.ExactSummation(N))
        out[f"{name}/wref/E"], out[f"{name}/wref/F"], out[f"{name}/wref/Ok"] = t.E_local, t.F_vector, t.O_k_vector
        out[f"{name}/wref/S"] = t.S_matrix
    # HilbertSpaceDistance (the reference instantiates (PsiDeep, PsiDeep) and (PsiCNN, PsiCNN)); inputs: helpers.hsd_cases()
    for name, (spec, spec_prime, OP, is_unitary, N) in hsd_cases().items():
        psi, psi_prime, op, es = make_psi(R, spec), make_psi(R, spec_prime), make_op(R, OP), R.ExactSummation(N)
        for p in (psi, psi_prime):
            if hasattr(p, "init_gradient"):
                p.init_gradient(1 << N)
        out[f"hsd/{name}/distance"] = R.hilbert_space_distance(psi, psi_prime, op, is_unitary, es)
        g, d = R.hilbert_space_distance_gradient(psi, psi_prime, op, is_unitary, es, 1.0)
        out[f"hsd/{name}/gradient"], out[f"hsd/{name}/distance_g"] = g, d
    # KullbackLeibler (psi: PsiClassical kinds, psi_prime: PsiDeep | PsiCNN): three consecutive calls, helpers.kl_sequence()
    for name in kl_cases():
        for k, v in kl_sequence(R, name, R.ExactSummation).items():
            out[f"kl/{name}/{k}"] = v
    # primitives: Pauli action (bit-exact) and activation polynomials
    rng = np.random.default_rng(7)
    a, b, c = (rng.integers(0, 1 << 63, size=64, dtype=np.uint64) for _ in range(3))
    res = [R.pauli_apply(int(x), int(y), int(z)) for x, y, z in zip(a, b, c)]
    out["pauli/a"], out["pauli/b"], out["pauli/conf"] = a, b, c
    out["pauli/coeff"] = np.array([r[0] for r in res])
    out["pauli/conf_out"] = np.array([r[1] for r in res], dtype=np.uint64)
    z = 0.7 * (rng.normal(size=16) + 1j * rng.normal(size=16))
    out["act/z"] = z
    for layer in (0, 1, 2):
        vals = [R.activation(x, layer) for x in z]
        out[f"act/lc{layer}"] = np.array([v[0] for v in vals])
        out[f"act/th{layer}"] = np.array([v[1] for v in vals])
    np.savez_compressed(os.path.join(HERE, "ref_golden.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
