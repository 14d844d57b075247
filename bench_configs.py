#!/usr/bin/env python
"""Secondary measurements: one JSON line per BASELINE.json configuration (C1..C5) on ONE B200, for BASELINE.md §4.

bench.py is the contract benchmark (C2); this script times the other shapes through the same public API:
  C1  PsiRBM a=2 N=16, TFIM ring, ExactSummation: energy + gradient (states/s)
  C3  PsiCNN 10x10, 3x3 channels, J1 Heisenberg, 32768 chains: E_loc + O_k (samples/s)
  C4  PsiDeep 64-64-64 on the 8x8 TFIM, 16384 samples: TDVP.eval with dense S (P = 8384) + dense solve
  C5  PsiRBM a=8 N=200 (one GPU's shard: 16384 chains), Heisenberg ring: eval_F + matrix-free CG
CUDA events via the library's phase timers; 3 warm-up + `--steps` timed repetitions; chains can be scaled down with
--scale for a quick run.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def timed(fn, steps, warmup=3):
    import torch
    for _ in range(warmup):
        fn()
    evs = []
    for _ in range(steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in evs) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="C1,C3,C4,C5")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--scale", type=float, default=1.0, help="scale the number of chains / samples")
    args = ap.parse_args()
    import torch
    import annongpu_b200 as A
    from annongpu_b200 import factories as F
    from annongpu_b200 import distributed as D
    rank, world = D.init_from_env()
    fp64_peak = A.measure_fp64_tflops()
    todo = args.configs.split(",")

    if "C1" in todo:
        spec, H = F.config_C1()
        psi, op, es = spec.build(True), H.build(True), A.ExactSummationSpins(16, True)
        psi.normalize(es)
        ev = A.ExpectationValue(True)
        ms = timed(lambda: ev.gradient(op, psi, es), args.steps)
        g, E = ev.gradient(op, psi, es)
        print(json.dumps({"config": "C1", "what": "ExactSummation energy + gradient, PsiRBM 16x32 (P=512), TFIM ring", "states": 65536,
                          "ms": ms, "states_per_s": 65536 / (ms * 1e-3), "E": [E.real, E.imag]}))

    if "C3" in todo:
        spec, H = F.config_C3()
        psi, op = spec.build(True), H.build(True)
        chains = max(148, int(32768 * args.scale))
        mc = A.MonteCarloSpins(chains, 1, 10, chains, True, seed=3)
        t = A.TDVP(psi.num_params, True)
        t.set_profile(True)
        ms = timed(lambda: t.eval_F(op, psi, mc), args.steps, warmup=1)
        ph = t.phase_ms
        print(json.dumps({"config": "C3", "what": "PsiCNN 10x10, 3 layers x 3 channels, 3x3 kernels (P=189), J1 Heisenberg (600 strings), "
                          "10+1 sweeps, E_loc + O_k + F", "chains": chains, "ms": ms, "samples_per_s": chains / (ms * 1e-3),
                          "phase_ms": ph, "acceptance": mc.acceptance_rate, "E": t.E_local.real}))

    if "C4" in todo:
        spec, H = F.config_C4()
        psi, op = spec.build(True), H.build(True)
        ns = max(148, int(16384 * args.scale))
        P = psi.num_params
        mc = A.MonteCarloSpins(ns, 1, 10, ns, True, seed=4)
        t = A.TDVP(P, True)
        t.set_profile(True)
        ms_eval = timed(lambda: t.eval(op, psi, mc), args.steps, warmup=1)
        ph = t.phase_ms
        import numpy as np
        solve_ms = []
        for _ in range(3):                        # the first call creates the cuSOLVER handle and workspace
            x = t.solve(shift_abs=0.0, shift_rel=1e-3)
            solve_ms.append(t.phase_ms["solve"])
        ms_solve = min(solve_ms)
        Sm, Fv = t.S_matrix, t.F_vector
        Sm[np.diag_indices(P)] *= 1.0 + 1e-3
        solve_residual = float(np.linalg.norm(Sm @ x - Fv) / np.linalg.norm(Fv))
        del Sm
        return_cg = {}
        for label, env in (("cg_on_dense_S", "0"), ("cg_matrix_free_on_O", "1")):
            os.environ["ANGPU_CG_MATRIX_FREE"] = env
            # a bandwidth probe of the two S.v products, NOT a solve: at this shift cond(S) is far too large for CG
            # (the Cholesky solve above is the C4 path), so the iteration count is capped
            x_cg, it_cg, rr_cg = t.solve_cg(tol=1e-6, max_iter=300, shift_abs=0.0, shift_rel=1e-3)
            ms_cg = t.phase_ms["solve"]
            bytes_per_it = (P * P * 16.0) if env == "0" else (2.0 * ns * P * 16)
            return_cg[label] = {"iterations": it_cg, "rel_residual": rr_cg, "ms": ms_cg, "ms_per_iteration": ms_cg / max(1, it_cg),
                                "algorithmic_GB_per_s": bytes_per_it * it_cg / (ms_cg * 1e-3) / 1e9,
                                "bytes_per_iteration": bytes_per_it,
                                "note": "capped at 300 iterations: mat-vec bandwidth probe, not converged"}
        os.environ["ANGPU_CG_MATRIX_FREE"] = "0"
        cg = return_cg
        flops_S = 4.0 * ns * P * P
        # opt-in tensor-core rebuild of S from the same samples (3xTF32 on tcgen05): time and deviation from the fp64 S
        import numpy as np
        S64 = t.S_matrix
        tc_ms = []
        for _ in range(3):
            t.build_S_tensorcore()
            tc_ms.append(t.phase_ms["s_build"])
        S32 = t.S_matrix
        tc_err = float(np.abs(S32 - S64).max() / np.abs(S64).max())
        del S64, S32
        mma_flops = 12.0 * ns * P * P            # 12 real 128x128x8 products per k-step on the upper-triangular tile pairs
        tc = {"ms": min(tc_ms), "equivalent fp64 TFLOP/s (4 Ns P^2)": flops_S / (min(tc_ms) * 1e-3) / 1e12,
              "issued tf32 TFLOP/s (12 Ns P^2)": mma_flops / (min(tc_ms) * 1e-3) / 1e12, "max rel deviation from fp64 S": tc_err,
              "includes": "fp64 centring + TF32 hi/lo packing (k_pack_planes) + k_sbuild_tf32"}
        print(json.dumps({"config": "C4", "what": "PsiDeep 64->64->64 (P=8384), 8x8 TFIM (192 strings), TDVP.eval: sampling + E_loc + O_k + dense S, "
                          "then Cholesky solve", "samples": ns, "ms_eval": ms_eval, "sr_steps_per_s": 1e3 / (ms_eval + ms_solve),
                          "phase_ms": ph, "ms_dense_solve": ms_solve, "ms_dense_solve_calls": solve_ms, "dense_solve_rel_residual": solve_residual,
                          "S_build": {"ms": ph["s_build"], "TFLOP/s (4 Ns P^2)": flops_S / (ph["s_build"] * 1e-3) / 1e12,
                                      "fp64_peak_measured": fp64_peak, "frac_of_fp64_peak": flops_S / (ph["s_build"] * 1e-3) / 1e12 / fp64_peak},
                          "S_build_tensorcore": tc, "cg": cg,
                          "acceptance": mc.acceptance_rate, "E": t.E_local.real, "x_norm": float(abs(x).max())}))

    if "C5" in todo:
        spec, H = F.config_C5()
        psi, op = spec.build(True), H.build(True)
        chains = max(148, int(16384 * args.scale)) * world       # under torchrun: 16384 chains per GPU (8 GPUs = BASELINE config 5)
        mc = A.MonteCarloSpins(chains, 1, 10, chains, True, seed=5)
        if world > 1:
            mc.set_shard(rank, world)
        t = A.TDVP(psi.num_params, True)
        t.set_profile(True)
        ms = timed(lambda: t.eval_F(op, psi, mc), args.steps, warmup=1)
        ph = t.phase_ms
        cg_runs = []
        for _ in range(3):                       # the first solve of a process carries ~150 ms of one-time costs (kernel loading, workspaces)
            x, it, rr = t.solve_cg(tol=1e-6, max_iter=2000, shift_abs=0.0, shift_rel=1e-3)
            cg_runs.append(t.phase_ms["solve"])
        ms_cg = min(cg_runs)
        if rank == 0:
          print(json.dumps({"config": "C5 (one GPU's shard)" if world == 1 else f"C5 on {world} GPUs", "n_gpus": world, "what": "PsiRBM 200x1600 (P=320000), Heisenberg ring (600 strings), 10+1 sweeps, "
                          "eval_F + matrix-free CG (tol 1e-6, shift 1e-3 diag)", "chains": chains, "ms_eval_F": ms,
                          "samples_per_s": chains / (ms * 1e-3), "phase_ms": ph, "cg_iterations": it, "cg_rel_residual": rr, "ms_cg": ms_cg,
                          "ms_per_cg_iteration": ms_cg / max(1, it), "ms_cg_runs": cg_runs, "sr_steps_per_s": 1e3 / (ms + ms_cg),
                          "acceptance": mc.acceptance_rate, "E": t.E_local.real}))
    D.shutdown()


if __name__ == "__main__":
    main()
