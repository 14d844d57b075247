"""Worker of tests/test_gpu_multirank.py (launched under torchrun, one rank per GPU): drives the PRODUCT multi-rank path --
distributed.init_from_env, sharded ensembles, TDVP.eval_F / eval / S_dot_vector / solve_cg / build_S_tensorcore,
ExpectationValue -- and compares every result with a one-rank run of the same global chains on rank 0."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import annongpu_b200 as A                                   # noqa: E402
from annongpu_b200 import distributed as D                  # noqa: E402
from annongpu_b200 import factories as F                    # noqa: E402


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def main():
    rank, world = D.init_from_env()
    assert world > 1
    report = {"world": world, "transport": os.environ.get("ANGPU_COMM", "nccl")}
    inherit = report["transport"] == "nccl"                 # ensembles inherit (rank, world) from the library communicator

    def sharded(ens):
        return ens if inherit else ens.set_shard(rank, world)

    # ---- PsiRBM (factorised rows): eval_F, S.v, matrix-free CG
    spec = F.rbm_spec(12, 24, noise=0.05, final_weight=1.0, seed=5)
    H = F.heisenberg(12, F.ring_bonds(12))
    psi, op = spec.build(True), H.build(True)
    chains = 1024
    mc = sharded(A.MonteCarloSpins(chains, 2, 5, chains, True, seed=11))
    t = A.TDVP(psi.num_params, True)
    t.eval_F(op, psi, mc)
    v = np.exp(1j * np.arange(psi.num_params))
    sv = t.S_dot_vector(v, mc)
    x, it, rr = t.solve_cg(tol=1e-10, max_iter=500, shift_rel=1e-3)
    E, Fv, Ok = t.E_local, t.F_vector, t.O_k_vector
    ev = A.ExpectationValue(True)
    fl, Em = ev.fluctuation(op, psi, sharded(A.MonteCarloSpins(chains, 2, 5, chains, True, seed=12)))
    # exact summation sharded over basis ranges
    es = sharded(A.ExactSummationSpins(12, True))
    g_es, E_es = ev.gradient(op, psi, es)
    if rank == 0:
        one = A.MonteCarloSpins(chains, 2, 5, chains, True, seed=11).set_shard(0, 1)
        t1 = A.TDVP(psi.num_params, True)
        t1.eval_F(op, psi, one)
        report["rbm_E"] = abs(t1.E_local - E) / abs(E)
        report["rbm_F"] = rel(Fv, t1.F_vector)
        report["rbm_Ok"] = rel(Ok, t1.O_k_vector)
        report["rbm_Sv"] = rel(sv, t1.S_dot_vector(v, one))
        x1, it1, rr1 = t1.solve_cg(tol=1e-10, max_iter=500, shift_rel=1e-3)
        report["rbm_cg_x"] = rel(x, x1)
        report["rbm_cg_it"] = [it, it1]
        fl1, Em1 = ev.fluctuation(op, psi, A.MonteCarloSpins(chains, 2, 5, chains, True, seed=12).set_shard(0, 1))
        report["fluct"] = abs(fl - fl1) / abs(fl1) + abs(Em - Em1) / abs(Em1)
        g1, E1 = ev.gradient(op, psi, A.ExactSummationSpins(12, True).set_shard(0, 1))
        report["es_E"] = abs(E_es - E1) / abs(E1)
        report["es_grad"] = rel(g_es, g1)

    # ---- PsiDeep (dense rows): eval with the exact S, the tcgen05 S build, CG on S and the dense solve
    dspec = F.deep_spec(8, 8, [8, 8], [8, 8], noise=0.05, final_weights=1.0, seed=6)
    Hd = F.tfim(8, F.ring_bonds(8))
    dpsi, dop = dspec.build(True), Hd.build(True)
    dmc = sharded(A.MonteCarloSpins(2048, 2, 5, 512, True, seed=21))
    td = A.TDVP(dpsi.num_params, True)
    td.eval(dop, dpsi, dmc)
    S = td.S_matrix
    xd = td.solve(shift_rel=1e-3)
    td.build_S_tensorcore()
    S_tc = td.S_matrix
    if rank == 0:
        one = A.MonteCarloSpins(2048, 2, 5, 512, True, seed=21).set_shard(0, 1)
        t1 = A.TDVP(dpsi.num_params, True)
        t1.eval(dop, dpsi, one)
        report["deep_S"] = rel(S, t1.S_matrix)
        report["deep_F"] = rel(td.F_vector, t1.F_vector)
        report["deep_solve"] = rel(xd, t1.solve(shift_rel=1e-3))
        report["deep_S_tensorcore_vs_fp64"] = rel(S_tc, t1.S_matrix)
        print("MULTIRANK " + json.dumps(report), flush=True)
    D.shutdown()


if __name__ == "__main__":
    main()
