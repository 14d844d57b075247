#!/usr/bin/env python
"""Benchmark of the VMC hot path on B200 (contract: see the task's bench.py section).

  python bench.py --gpus N --steps K --warmup W            our arm (one rank per GPU under torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  the CPU arm: the reference's algorithm on the host cores
  python bench.py --config C5 ...                          BASELINE.json configs[4] instead of the headline C2

Workload C2 (BASELINE.json configs[1], the default): PsiRBM alpha=4 on the 1-D Heisenberg ring, N=64, M=256, P=16384,
MonteCarlo with 8192 chains per GPU, num_samples = num_chains, 10 thermalisation sweeps + 1 sweep per sample.
Workload C5 (configs[4]): PsiRBM alpha=8, N=200, M=1600, P=320000, 16384 chains per GPU (131072 over 8 GPUs).
One "step" = one call of the path  sampling -> E_loc -> O_k -> <E>, <O_k>, F  (TDVP.eval_F, which is also
ExpectationValue.gradient).  `value` = MC samples/s with everything resident in HBM; `e2e` = the same through the
public API with host buffers (parameters uploaded from pinned memory, F and E read back every step).  `sr` reports full
SR steps/s (eval_F + matrix-free CG to 1e-6 + parameter update on the device).
"""
import argparse
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNIT = "samples/s"
THERM, SWEEPS = 10, 1
SEED = 0xA11CE
CONFIGS = {
    # name: (factory, N, M, chains per GPU, Pauli strings)
    "C2": ("config_C2", 64, 256, 8192, 192),
    "C5": ("config_C5", 200, 1600, 16384, 600),
}


def metric_name(cfg):
    return f"MC samples/s (sampling + E_loc + O_k + F; PsiRBM {cfg})"


def config_dict(cfg, chains_local, world):
    """The SAME dictionary in both arms (the reference arm times a bounded sample of this workload: cpu_baseline.sample)."""
    _, N, M, _, strings = CONFIGS[cfg]
    chains = chains_local * world
    return {"workload": (f"{cfg}: PsiRBM alpha={M // N} N={N} M={M} (P={N * M}), Heisenberg ring ({strings} Pauli strings), "
                         f"MonteCarloSpins(num_samples={chains}, num_sweeps={SWEEPS}, num_thermalization_sweeps={THERM}, "
                         f"num_markov_chains={chains}), TDVP.eval_F = sampling + E_loc + O_k + <E>,<O_k>,F"),
            "chains_per_gpu": chains_local, "total_chains": chains,
            "l2": "flushed between timed iterations (256 MiB device write)", "parallelism": f"chains sharded x{world}"}


def load_factories():
    """annongpu_b200/factories.py WITHOUT importing the package (its __init__ loads libangpu.so): the reference arm must not
    touch the CUDA library."""
    spec = importlib.util.spec_from_file_location("angpu_factories_standalone", os.path.join(ROOT, "annongpu_b200", "factories.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod           # dataclasses resolve their module through sys.modules
    spec.loader.exec_module(mod)
    return mod


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def host_threads():
    """Threads for the CPU arm.  torchrun exports OMP_NUM_THREADS=1 to its workers, which is not a statement about the box:
    ANGPU_CPU_THREADS overrides, else every core this process may run on."""
    env = os.environ.get("ANGPU_CPU_THREADS")
    if env:
        return max(1, int(env))
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


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

def cpu_gradient_rate(cfg, seconds_target, calls=1):
    """Times the oracle port's ExpectationValue::gradient over MonteCarlo (sampling + E_loc + O_k) on the host cores.
    The unmodified reference cannot run C2 / C5 (PsiRBM::max_N = 128 < M, 64-bit masks), so the C port
    (oracle/port/vmc_port.c) of the same algorithm is timed: kind = "port"."""
    from oracle import port_oracle as P            # bench.py's cpu_baseline leg: the one place bench may run oracle/
    F = load_factories()
    spec, H = getattr(F, CONFIGS[cfg][0])()
    c, a, b = H.arrays(F.words_for(CONFIGS[cfg][1]))
    psi, op = P.PsiRBM(spec.W, spec.final_weight, 0.0), P.Operator(c, a, b, F.words_for(CONFIGS[cfg][1]))
    cores = host_threads()
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
    n, times, cores = cpu_gradient_rate(args.config, per_step_s, calls=steps + warmup)
    timed = times[warmup:]
    total = sum(timed)
    value = n * len(timed) / total
    chains_local = args.chains or CONFIGS[args.config][3]
    line = {
        "impl": "reference", "metric": metric_name(args.config), "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": 1e3 * total / len(timed), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64 (complex)", "data": "synthetic",
        "config": config_dict(args.config, chains_local, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{n} chains per step with the workload's per-chain work ({THERM}+{SWEEPS} sweeps, E_loc, O_k), "
                                   f"{len(timed)} timed steps, OpenMP over chains on {cores} threads; the unmodified reference "
                                   "cannot run this shape (PsiRBM::max_N=128)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ------------------------------------------------------------------------------------------------ our arm

def ncu_pipe_fractions(cfg):
    """Pipe utilisation of the sampler from the committed ncu --set full capture (profiles/), quoted next to the live
    roofline; None when no capture of this build is committed."""
    path = os.path.join(ROOT, "profiles", f"r02_ncu_sampler_{cfg}.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f)
    return None


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import annongpu_b200 as A
    from annongpu_b200 import distributed as D
    from annongpu_b200 import factories as F

    cfg = args.config
    factory, N, M, chains_default, _ = CONFIGS[cfg]
    rank, world = D.init_from_env()
    dev = torch.device("cuda", torch.cuda.current_device())
    chains_local = args.chains or chains_default
    chains = chains_local * world                     # weak scaling: per-GPU work fixed
    spec, H = getattr(F, factory)()
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
            D.barrier()
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
        return D.max_over_ranks(ms), phases

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
        _, it, rr = tdvp.solve_cg(tol=1e-6, max_iter=2000, shift_abs=0.0, shift_rel=1e-3, keep_on_device=True)
        out["cg_it"], out["cg_rr"] = it, rr
        tdvp.apply_update(psi, -1e-3)                 # SR / imaginary-time step, parameters updated on the device

    for _ in range(max(3, args.warmup)):
        step_device()
    clocks = ClockSampler(torch.cuda.current_device())
    clocks.start()
    A.launch_count(reset=True)
    ms_total, phases = timed_loop(step_device, args.steps)
    launches = A.launch_count()
    clock_info = clocks.stop()
    value = chains * args.steps / (ms_total * 1e-3)
    acceptance = mc.acceptance_rate

    # ---- N-rank check: the sharded E / F equal a one-rank run of the same global chains (Philox is keyed by the global
    # chain id, so the results agree to reduction order).  Rank 0 re-runs ALL chains unsharded; the others wait.
    rank_check = None
    if world > 1 and not args.no_rank_check:
        call_before = mc.call_index
        tdvp.eval_F(op, psi, mc)
        E_sh, F_sh = tdvp.E_local, tdvp.F_vector
        if rank == 0:
            full = A.MonteCarloSpins(chains, SWEEPS, THERM, chains, True, seed=SEED).set_shard(0, 1)
            full.call_index = call_before
            t1 = A.TDVP(P, True)
            t1.eval_F(op, psi, full)
            dE = abs(t1.E_local - E_sh) / abs(t1.E_local)
            dF = float(np.abs(t1.F_vector - F_sh).max() / np.abs(t1.F_vector).max())
            rank_check = {"ranks": world, "rel_dev_E": float(dE), "rel_dev_F": dF, "tolerance": 1e-10,
                          "what": "eval_F over the sharded chains vs the same global chains on one rank"}
            assert dE <= 1e-10 and dF <= 1e-10, rank_check
            del t1, full
        barrier()

    for _ in range(2):
        step_e2e()
    ms_e2e, _ = timed_loop(step_e2e, args.steps)
    e2e_value = chains * args.steps / (ms_e2e * 1e-3)

    sr_steps = max(2, min(args.steps, 5))
    params0 = psi.params.copy()
    step_sr()
    ms_sr, _ = timed_loop(step_sr, sr_steps)
    psi.params = params0

    # ---- roofline of the dominant kernel (the RBM sampler), timed live by CUDA events inside the timed region
    words = F.words_for(N)
    t_sample = 1e-3 * sum(p["sample"] for p in phases) / len(phases)
    t_eloc = 1e-3 * sum(p["eloc"] for p in phases) / len(phases)
    t_ok = 1e-3 * sum(p["ok_reduce"] for p in phases) / len(phases)
    t_tot = 1e-3 * sum(p["total"] for p in phases) / len(phases)
    proposals = (THERM + SWEEPS) * N
    # algorithmic flops: per proposal and hidden unit 18 (angle update 2 FMA = 4, Re lc0 in (p, q) form: 14; real final
    # weight); the undo on rejection and the fp32-screened exp of the acceptance test are not counted
    flops = chains_local * proposals * M * 18.0
    # algorithmic HBM bytes of one launch: W read once + per chain (conf + log_psi + M cached angles) written
    hbm_bytes = N * M * 16 + chains_local * (8 * words + 16 + 16 * M)
    hbm_peak, peak_src = measured_peaks()
    fp64_peak = A.measure_fp64_tflops()
    kernel = "k_mc_rbm<8,1,true>" if M <= 512 else f"k_mc_rbm_block<{(M + 255) // 256},true>"
    roofline = {"kernel": kernel, "bound": "fp64", "achieved": flops / t_sample / 1e12, "peak": fp64_peak,
                "unit": "TFLOP/s", "frac": flops / t_sample / 1e12 / fp64_peak,
                # dram__bytes_read + dram__bytes_write of the 8192-chain launch in the ncu --set full capture of this build
                # (profiles/r02_ncu_sampler_C2.json): 33.9 MB read -- theta_0 left in HBM by the angle GEMM -- and 0.05 MB written
                # (the 34 MB of outputs stay in the 126 MB L2); W and the angle cache stay on chip
                "traffic": 33.95e6 if (cfg == "C2" and chains_local == 8192) else None,
                "algorithmic_flops": flops,
                "peak_source": "measured in this run (angpu_measure_fp64_tflops: independent DFMA streams; MEASURED_PEAKS.json holds no "
                               "fp64 figure)",
                "ncu": ncu_pipe_fractions(cfg),
                "note": "FP64-pipe bound by design: 18 flops per proposal and hidden unit, no HBM traffic in the loop (SURVEY.md §8d)"}
    roofline_hbm = {"kernel": kernel, "bound": "hbm", "achieved": hbm_bytes / t_sample / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": hbm_bytes / t_sample / 1e9 / hbm_peak, "algorithmic_bytes": hbm_bytes, "peak_source": peak_src}

    line = {
        "metric": metric_name(cfg), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64 (complex)", "data": "synthetic",
        "config": config_dict(cfg, chains_local, world),
        "acceptance_rate": acceptance,
        "clocks": clock_info,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": P * 16, "d2h_bytes_per_step": P * 16 + 16,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "roofline": roofline, "roofline_hbm": roofline_hbm,
        "phase_ms": {"sample": 1e3 * t_sample, "eloc": 1e3 * t_eloc, "ok_reduce": 1e3 * t_ok, "total": 1e3 * t_tot},
        "sr": {"steps_per_sec": sr_steps / (ms_sr * 1e-3), "ms_per_step": ms_sr / sr_steps, "cg_iterations": out.get("cg_it"),
               "cg_rel_residual": out.get("cg_rr"),
               "what": "eval_F + matrix-free CG (tol 1e-6, shift 1e-3*diag S) + parameter update on the device"},
        "comm": os.environ.get("ANGPU_COMM", "nccl (in-library)") if world > 1 else None,
        "rank_check": rank_check,
    }
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            n, times, cores = cpu_gradient_rate(cfg, args.cpu_seconds)
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
    ap.add_argument("--config", default="C2", choices=sorted(CONFIGS))
    ap.add_argument("--chains", type=int, default=0, help="chains per GPU (default: the configuration's)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-rank-check", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
