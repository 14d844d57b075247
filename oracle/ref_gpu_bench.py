#!/usr/bin/env python
"""Second baseline: the UNMODIFIED reference's own CUDA kernels (gpu=true) on the same B200, next to this library,
on the shapes the reference's fixed limits allow: C1 as is, C2 at M = 128 (the reference's PsiRBM
max), C4 (PsiDeep 64-64-64, dense S through its atomics kernel).  oracle/_ref/liboracle_ref.so was compiled from
/root/reference with -arch=sm_100 (oracle/Makefile), so it carries the reference's device code.

Each case runs in a child process under `timeout` (a hung reference kernel must not take the box down); one JSON line
per case: wall-clock ms per call of the reference (it synchronises internally by copying results to the host) and
CUDA-event ms of ours for the same call.  Not a parity test; energies are printed only as a sanity check.
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))   # repo root (this file lives in oracle/: test infrastructure)
sys.path.insert(0, ROOT)


def wall(fn, reps):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps * 1e3


def case(name, scale):
    import numpy as np
    import annongpu_b200 as A
    from annongpu_b200 import factories as F
    from oracle import ref_oracle as R
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    out = {"case": name}
    # the reference's serial host path (gpu=false) on a bounded sample of the same workload, one core
    R.set_gpu(False)
    if name == "C1":
        spec, H = F.config_C1()
        rp, ro, re_ = helpers.make_psi(R, spec), helpers.make_op(R, H), R.ExactSummation(16)
        out["ref_cpu_ms"] = wall(lambda: R.gradient(ro, rp, re_), 1)
        out["ref_cpu_unit"] = "ms per call (65536 states), 1 core"
    elif name == "C2_M128":
        N, M = 64, 128
        spec, H = F.rbm_spec(N, M, noise=0.02 / np.sqrt(2.0), final_weight=1.0, seed=1234), F.heisenberg(N, F.ring_bonds(N))
        # the C2 workload per sample = 10 thermalisation sweeps + 1 sweep + E_loc + O_k for EVERY chain: the reference's
        # host path runs one chain per ensemble, so one call of MonteCarloSpins(1, 1, 10, 1) is one C2 sample
        rp, ro, rm = helpers.make_psi(R, spec), helpers.make_op(R, H), R.MonteCarlo(1, 1, 10, 1)
        rt = R.TDVP(rp.num_params)
        ms = wall(lambda: rt.eval_F(ro, rp, rm), 64)
        out["ref_cpu_samples_per_s"] = 1.0 / (ms * 1e-3)
        out["ref_cpu_unit"] = "64 calls of TDVP.eval_F with MonteCarloSpins(1, 1, 10, 1) on 1 core (same per-sample work as C2)"
    elif name == "C4":
        spec, H = F.config_C4()
        rp, ro, rm = helpers.make_psi(R, spec), helpers.make_op(R, H), R.MonteCarlo(32, 1, 10, 1)
        rt = R.TDVP(rp.num_params)
        ms = wall(lambda: rt.eval(ro, rp, rm), 1)
        out["ref_cpu_samples_per_s"] = 32 / (ms * 1e-3)
        out["ref_cpu_unit"] = "TDVP.eval with MonteCarloSpins(32, 1, 10, 1) on 1 core (S fill is O(Ns P^2) on the host)"
    R.set_gpu(True)
    if name == "C1":
        spec, H = F.config_C1()
        rp, ro, re_ = helpers.make_psi(R, spec), helpers.make_op(R, H), R.ExactSummation(16)
        gp, go, ge = spec.build(True), H.build(True), A.ExactSummationSpins(16, True)
        ev = A.ExpectationValue(True)
        out["ref_gpu_ms"] = wall(lambda: R.gradient(ro, rp, re_), 5)
        out["ours_ms"] = wall(lambda: ev.gradient(go, gp, ge), 20)
        out["E_ref"], out["E_ours"] = R.gradient(ro, rp, re_)[1].real, ev.gradient(go, gp, ge)[1].real
        out["what"] = "ExpectationValue.gradient, PsiRBM 16x32, TFIM ring, ExactSummation 65536 states"
    elif name == "C2_M128":
        N, M = 64, 128
        spec, H = F.rbm_spec(N, M, noise=0.02 / np.sqrt(2.0), final_weight=1.0, seed=1234), F.heisenberg(N, F.ring_bonds(N))
        chains = max(64, int(8192 * scale))
        rp, ro, rm = helpers.make_psi(R, spec), helpers.make_op(R, H), R.MonteCarlo(chains, 1, 10, chains)
        gp, go, gm = spec.build(True), H.build(True), A.MonteCarloSpins(chains, 1, 10, chains, True, seed=2)
        rt, gt = R.TDVP(rp.num_params), A.TDVP(gp.num_params, True)
        out["ref_gpu_ms"] = wall(lambda: rt.eval_F(ro, rp, rm), 3)
        out["ours_ms"] = wall(lambda: gt.eval_F(go, gp, gm), 10)
        out["E_ref"], out["E_ours"] = rt.E_local.real, gt.E_local.real
        out["chains"] = chains
        out["what"] = "TDVP.eval_F (sampling 10+1 sweeps + E_loc + O_k + F), PsiRBM 64x128 (reference's M limit), Heisenberg ring"
    elif name == "C4":
        spec, H = F.config_C4()
        ns = max(64, int(16384 * scale))
        rp, ro, rm = helpers.make_psi(R, spec), helpers.make_op(R, H), R.MonteCarlo(ns, 1, 10, ns)
        gp, go, gm = spec.build(True), H.build(True), A.MonteCarloSpins(ns, 1, 10, ns, True, seed=4)
        rt, gt = R.TDVP(rp.num_params), A.TDVP(gp.num_params, True)
        out["ref_gpu_ms"] = wall(lambda: rt.eval(ro, rp, rm), 1)
        out["ours_ms"] = wall(lambda: gt.eval(go, gp, gm), 3)
        out["E_ref"], out["E_ours"] = rt.E_local.real, gt.E_local.real
        out["samples"] = ns
        out["what"] = "TDVP.eval (sampling + E_loc + O_k + dense S, P = 8384), PsiDeep 64-64-64, 8x8 TFIM"
    out["speedup"] = out["ref_gpu_ms"] / out["ours_ms"]
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="C1,C2_M128,C4")
    ap.add_argument("--case", default=None, help="(internal) run one case in this process")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--timeout", type=int, default=240)
    args = ap.parse_args()
    if args.case:
        case(args.case, args.scale)
        return
    for c in args.cases.split(","):
        cmd = ["timeout", str(args.timeout), sys.executable, os.path.abspath(__file__), "--case", c, "--scale", str(args.scale)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        line = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if line:
            print(line[-1], flush=True)
        else:
            print(json.dumps({"case": c, "failed": r.returncode, "stderr_tail": r.stderr[-400:]}), flush=True)


if __name__ == "__main__":
    main()
