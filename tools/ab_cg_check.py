"""A/B of the convergence-check period of the exact CG (ANGPU_CG_CHECK) on the C2 SR step of bench.py."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for c in (sys.argv[1:] or ["8", "4", "2", "8", "4"]):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "20", "--warmup", "3", "--no-cpu-baseline"],
                       env={**os.environ, "ANGPU_CG_CHECK": c}, capture_output=True, text=True)
    for l in p.stdout.splitlines():
        if l.startswith("{"):
            d = json.loads(l)
            print("check_every", c, "sr_ms", round(d["sr"]["ms_per_step"], 3), "cg_it", d["sr"]["cg_iterations"], "steps/s", round(d["sr"]["steps_per_sec"], 2), flush=True)
