"""A/B driver of the PsiRBM samplers: runs the same chains with the all-fp64 sampler (ANGPU_MC_SCREEN=0) and the
fp32-screened one and checks that configurations, acceptance counts and log psi coincide; prints the sampler phase time.
    python tools/ab_sampler.py            (spawns one subprocess per mode; the switch is read once per process)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SHAPES = {          # name: (N, M, chains, sweeps, therm, noise)
    "C2": (64, 256, 8192, 1, 10, None),
    "n40m100": (40, 100, 4096, 2, 5, 0.05),
    "n100m500": (100, 500, 2048, 1, 3, 0.02),
    "n20m40": (20, 40, 4096, 2, 5, 0.1),
}


def child(shape, out):
    import numpy as np
    import annongpu_b200 as A
    from annongpu_b200 import factories as F
    A.setDevice(0)
    N, M, chains, sweeps, therm, noise = SHAPES[shape]
    if shape == "C2":
        spec, H = F.config_C2()
    else:
        spec = F.rbm_spec(N, M, noise=noise, final_weight=1.0, seed=77)
        H = F.heisenberg(N, F.ring_bonds(N))
    psi, op = spec.build(True), H.build(True)
    mc = A.MonteCarloSpins(chains, sweeps, therm, chains, True, seed=4242)
    res = {}
    confs, lps, accs, exact = [], [], [], []
    for call in range(2):
        c, lp = mc.sample(psi)
        confs.append(c); lps.append(lp); accs.append(mc.acceptances); exact.append(mc.exact_decisions)
    tdvp = A.TDVP(psi.num_params, True)
    tdvp.set_profile(True)
    ts = []
    for it in range(12):
        tdvp.eval_F(op, psi, mc)
        ts.append(tdvp.phase_ms["sample"])
    res["sample_ms"] = float(np.median(ts[2:]))
    res["phase_ms"] = tdvp.phase_ms
    res["acc"] = accs; res["exact"] = exact
    res["E"] = [tdvp.E_local.real, tdvp.E_local.imag]
    np.savez(out, confs=np.stack(confs), lps=np.stack(lps))
    print(json.dumps(res))


def main():
    import numpy as np
    shapes = sys.argv[1:] or list(SHAPES)
    modes = {"fp64": {"ANGPU_MC_SCREEN": "0"}, "scr_r32": {"ANGPU_MC_SCREEN": "1", "ANGPU_MC_REFRESH": "32"},
             "scr_r16": {"ANGPU_MC_SCREEN": "1", "ANGPU_MC_REFRESH": "16"},
             "scr_r32_m10": {"ANGPU_MC_SCREEN": "1", "ANGPU_MC_REFRESH": "32", "ANGPU_MC_MINB": "10"}, "scr_r16_m10": {"ANGPU_MC_SCREEN": "1", "ANGPU_MC_REFRESH": "16", "ANGPU_MC_MINB": "10"}}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    for shape in shapes:
        ref = None
        for mode, env in modes.items():
            out = f"/tmp/ab_{shape}_{mode}.npz"
            p = subprocess.run([sys.executable, __file__, "--child", shape, out], env={**os.environ, **env}, capture_output=True, text=True)
            if p.returncode != 0:
                print(shape, mode, "FAILED", p.stderr[-2000:]); continue
            res = json.loads(p.stdout.strip().splitlines()[-1])
            d = np.load(out)
            line = {"shape": shape, "mode": mode, "sample_ms": res["sample_ms"], "acc": res["acc"], "exact": res["exact"], "E": res["E"],
                    "phase_ms": res["phase_ms"]}
            if ref is None:
                ref = d
            else:
                line["conf_identical"] = bool(np.array_equal(ref["confs"], d["confs"]))
                line["log_psi_max_abs_diff"] = float(np.abs(ref["lps"] - d["lps"]).max())
            print(json.dumps(line), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child(sys.argv[2], sys.argv[3])
    else:
        main()
