#!/usr/bin/env python
"""Benchmark of the VMC hot path on B200 (contract: see the task's bench.py section).

  python bench.py --gpus N --steps K --warmup W            our arm (one rank per GPU under torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  the CPU arm: the reference's algorithm on the host cores

Workload (BASELINE.json configs[1], "C2"): PsiRBM alpha=4 on the 1-D Heisenberg ring, N=64, M=256, P=16384,
MonteCarlo with 8192 chains per GPU, num_samples = num_chains, 10 thermalisation sweeps + 1 sweep per sample.
One "step" = one call of the path  sampling -> E_loc -> O_k -> <E>, <O_k>, F  (TDVP.eval_F, which is also
ExpectationValue.gradient): 8192 MC samples per GPU.  `value` = MC samples/s with everything resident in HBM;
`e2e` = the same through the public API with host buffers (parameters uploaded from pinned memory, F and E read
back every step).  `sr` reports full SR steps/s (eval_F + matrix-free CG to 1e-6 + parameter update).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "MC samples/s (sampling + E_loc + O_k + F; PsiRBM C2)"
UNIT = "samples/s"
CHAINS_PER_GPU = 8192
THERM, SWEEPS = 10, 1
SEED = 0xA11CE


def workload_name(chains):
    return (f"C2: PsiRBM alpha=4 N=64 M=256 (P=16384), Heisenberg ring (192 Pauli strings), MonteCarloSpins("
            f"num_samples={chains}, num_sweeps={SWEEPS}, num_thermalization_sweeps={THERM}, num_markov_chains={chains}), "
            f"TDVP.eval_F = sampling + E_loc + O_k + <E>,<O_k>,F")


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------ clocks

class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.device_index, self.rows, self.proc, self.thread = device_index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-i", str(self.device_index), "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); smax = float(r[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU arm

def cpu_gradient_rate(seconds_target, threads=0, calls=1):
    """Times the oracle port's ExpectationValue::gradient over MonteCarlo (sampling + E_loc + O_k) on the host cores.
    The unmodified reference cannot run C2 (PsiRBM::max_N = 128 < M = 256), so the C port
    (oracle/port/vmc_port.c) of the same algorithm is timed: kind = "port"."""
    from oracle import port_oracle as P            # bench.py's cpu_baseline leg: the one place bench may run oracle/
    from annongpu_b200 import factories as F
    spec, H = F.config_C2()
    c, a, b = H.arrays(1)
    psi, op = P.PsiRBM(spec.W, spec.final_weight, 0.0), P.Operator(c, a, b)
    cores = P.max_threads() if threads <= 0 else threads
    # calibrate on 2 chains per thread, then size the sample for ~seconds_target of wall time
    n0, dt = 2 * cores, 0.0
    for _ in range(4):                               # grow the calibration batch until it runs for >= 0.5 s
        t0 = time.perf_counter()
        P.mc_gradient_timed(psi, op, n0, SWEEPS, THERM, n0, seed=SEED, call=0, nthreads=cores)
        dt = time.perf_counter() - t0
        if dt >= 0.5:
            break
        n0 = max(2 * n0, int(n0 * 0.7 / max(dt, 1e-3)) // cores * cores)
    n = max(2 * cores, int(n0 * seconds_target / max(dt, 1e-3)) // cores * cores)
    times = []
    for k in range(calls):
        t0 = time.perf_counter()
        P.mc_gradient_timed(psi, op, n, SWEEPS, THERM, n, seed=SEED, call=1 + k, nthreads=cores)
        times.append(time.perf_counter() - t0)
    return n, times, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = args.steps, args.warmup
    per_step_s = max(1.0, min(20.0, 120.0 / max(1, steps + warmup)))
    n, times, cores = cpu_gradient_rate(per_step_s, calls=steps + warmup)
    timed = times[warmup:]
    total = sum(timed)
    value = n * len(timed) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": 1e3 * total / len(timed), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64 (complex)", "data": "synthetic",
        "config": {"workload": workload_name(CHAINS_PER_GPU), "sample": f"{n} chains per step (same per-chain work)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{n} of the {CHAINS_PER_GPU} chains per step, {len(timed)} steps, OpenMP over chains; "
                                   "unmodified reference cannot run M=256 (PsiRBM::max_N=128)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ------------------------------------------------------------------------------------------------ our arm

def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import annongpu_b200 as A
    from annongpu_b200 import distributed as D
    from annongpu_b200 import factories as F

    rank, world = D.init_from_env()
    dev = torch.device("cuda", torch.cuda.current_device())
    chains_local = args.chains
    chains = chains_local * world                     # weak scaling: per-GPU work fixed
    spec, H = F.config_C2()
    psi, op = spec.build(True), H.build(True)
    P = psi.num_params
    mc = A.MonteCarloSpins(chains, SWEEPS, THERM, chains, True, seed=SEED).set_shard(rank, world)
    tdvp = A.TDVP(P, True)
    tdvp.set_profile(True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    params_pinned = torch.from_numpy(psi.params.copy()).pin_memory()
    params_host = params_pinned.numpy()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_loop(step, n):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        phases = []
        barrier()
        for e0, e1 in evs:
            flush.zero_()
            e0.record()
            step()
            e1.record()
            phases.append(tdvp.phase_ms)
        barrier()
        ms = sum(e0.elapsed_time(e1) for e0, e1 in evs)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), phases

    def step_device():
        tdvp.eval_F(op, psi, mc)

    out = {}

    def step_e2e():
        psi.params = params_host                      # H2D from pinned memory (P complex128)
        tdvp.eval_F(op, psi, mc)
        out["F"] = tdvp.F_vector                      # D2H (P complex128)
        out["E"] = tdvp.E_local

    def step_sr():
        tdvp.eval_F(op, psi, mc)
        x, it, rr = tdvp.solve_cg(tol=1e-6, max_iter=2000, shift_abs=0.0, shift_rel=1e-3)
        out["cg_it"], out["cg_rr"] = it, rr
        psi.params = psi.params - 1e-3 * x            # SR / imaginary-time update

    for _ in range(max(3, args.warmup)):
        step_device()
    clocks = ClockSampler(torch.cuda.current_device())
    clocks.start()
    A.launch_count(reset=True)
    ms_total, phases = timed_loop(step_device, args.steps)
    launches = A.launch_count()
    clock_info = clocks.stop()
    value = chains * args.steps / (ms_total * 1e-3)

    for _ in range(2):
        step_e2e()
    ms_e2e, _ = timed_loop(step_e2e, args.steps)
    e2e_value = chains * args.steps / (ms_e2e * 1e-3)

    sr_steps = max(2, min(args.steps, 5))
    params0 = psi.params.copy()
    step_sr()
    ms_sr, _ = timed_loop(step_sr, sr_steps)
    psi.params = params0
    acceptance = mc.acceptance_rate

    # ---- roofline of the dominant kernel (k_mc_rbm, the sampler), timed live by CUDA events inside the timed region
    N, M, words = 64, 256, 1
    t_sample = 1e-3 * sum(p["sample"] for p in phases) / len(phases)
    t_eloc = 1e-3 * sum(p["eloc"] for p in phases) / len(phases)
    t_ok = 1e-3 * sum(p["ok_reduce"] for p in phases) / len(phases)
    t_tot = 1e-3 * sum(p["total"] for p in phases) / len(phases)
    proposals = (THERM + SWEEPS) * N
    # algorithmic HBM bytes of one launch: W read once + per chain (conf + log_psi + M cached angles) written
    hbm_bytes = N * M * 16 + chains_local * (8 * words + 16 + 16 * M)
    # algorithmic flops: per proposal and hidden unit 18 (angle update 2 FMA = 4, Re lc0 in (p, q) form: 14; real final
    # weight); the undo on rejection and the fp32-screened acceptance are not counted
    flops = chains_local * proposals * M * 18.0
    hbm_peak, peak_src = measured_peaks()
    fp64_peak = A.measure_fp64_tflops()
    roofline = {"kernel": "k_mc_rbm<8,true>", "bound": "hbm", "achieved": hbm_bytes / t_sample / 1e9, "peak": hbm_peak,
                "unit": "GB/s", "frac": hbm_bytes / t_sample / 1e9 / hbm_peak,
                # dram__bytes_read.sum + dram__bytes_write.sum of one launch at 8192 chains, from the ncu --set full capture
                # profiles/r01_ncu_full_sampler_eloc_after_tuning.csv (0.281 MB + 0.047 MB): BELOW the algorithmic bytes
                # because the 34 MB of outputs stay in the 126 MB L2 until the consumers read them
                "traffic": 0.328e6 if chains_local == 8192 else None, "algorithmic_bytes": hbm_bytes, "peak_source": peak_src,
                "note": "the sampler is FP64-pipe bound by design (W and the angle cache stay on chip, SURVEY.md §8d): "
                        "see roofline_fp64 for the binding roof"}
    roofline_fp64 = {"kernel": "k_mc_rbm<8,true>", "bound": "fp64", "achieved": flops / t_sample / 1e12, "peak": fp64_peak,
                     "unit": "TFLOP/s", "frac": flops / t_sample / 1e12 / fp64_peak,
                     "peak_source": "measured in this run (angpu_measure_fp64_tflops: independent DFMA streams)"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64 (complex)", "data": "synthetic",
        "config": {"workload": workload_name(chains_local), "chains_per_gpu": chains_local, "total_chains": chains,
                   "l2": "flushed between timed iterations (256 MiB device write)", "parallelism": f"chains sharded x{world}",
                   "acceptance_rate": acceptance},
        "clocks": clock_info,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": P * 16, "d2h_bytes_per_step": P * 16 + 16,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "roofline": roofline, "roofline_fp64": roofline_fp64,
        "phase_ms": {"sample": 1e3 * t_sample, "eloc": 1e3 * t_eloc, "ok_reduce": 1e3 * t_ok, "total": 1e3 * t_tot},
        "sr": {"steps_per_sec": sr_steps / (ms_sr * 1e-3), "ms_per_step": ms_sr / sr_steps, "cg_iterations": out.get("cg_it"),
               "cg_rel_residual": out.get("cg_rr"), "what": "eval_F + matrix-free CG (tol 1e-6, shift 1e-3*diag S) + parameter update"},
    }
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            n, times, cores = cpu_gradient_rate(args.cpu_seconds)
            line["cpu_baseline"] = {"value": n / times[0], "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{n} of the {chains_local} chains (same per-chain work: {THERM}+{SWEEPS} sweeps, E_loc, O_k), "
                                              f"one call, {times[0]:.1f} s, OpenMP over chains"}
        emit(line)
    D.shutdown()


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line of the contract, written to the process' original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    # libraries (NCCL prints its version banner to stdout) must not pollute the one-line contract: everything written to
    # fd 1 while the benchmark runs goes to stderr; the JSON line is written to the saved descriptor at the end
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chains", type=int, default=CHAINS_PER_GPU, help="chains per GPU")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
