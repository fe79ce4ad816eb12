"""Config 5 (BASELINE.json configs[4]) probe: SE3 sphere pose graphs of growing size on one GPU.
  python tests/config5_probe.py 200 300 500 1000      (nodes per level = laps = N, N*N poses)
One JSON line per size: structure time, per-iteration wall time, chi2 trajectory, factor statistics, or the error."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import openslam_g2o_b200 as g  # noqa: E402
from openslam_g2o_b200 import synth  # noqa: E402


def run(n, iters=3, nd=0):
    rec = {"nodes_per_level": n, "poses": n * n, "nd_levels": nd}
    t0 = time.perf_counter()
    prob = synth.sphere(n, n, seed=n * n)
    rec["edges"] = int(len(prob["edge_v0"]))
    rec["generate_s"] = time.perf_counter() - t0
    opt = g.SparseOptimizer(device=0)
    opt.set_algorithm("lm_fix6_3")
    synth.feed(prob, opt)
    opt.setup_cli()
    opt.initialize_optimization()
    try:
        opt._ensure_uploaded()
        ctx = opt.context
        if nd:
            ctx.set_ordering(nd)
        t0 = time.perf_counter()
        ok = ctx.build_structure()
        rec["build_structure_s"] = time.perf_counter() - t0
        rec["structure_ok"] = bool(ok)
        rec["factor"] = ctx.factor_info()
        times, chi = [], []
        for it in range(iters):
            t0 = time.perf_counter()
            rc, st = ctx.algorithm_solve(g.LEVENBERG, it)
            ctx.synchronize()
            times.append(time.perf_counter() - t0)
            chi.append(st.chi2)
            rec.setdefault("rc", []).append(int(rc))
            rec.setdefault("trials", []).append(int(st.levenberg_iterations))
        rec["iteration_s"] = times
        rec["chi2"] = chi
        ctx.set_profiling(True)
        ctx.algorithm_solve(g.LEVENBERG, iters)
        ph = ctx.phase_times()
        ctx.set_profiling(False)
        rec["phases_ms"] = {k: 1e3 * v[0] for k, v in ph.items() if v[1] > 0}
        f = rec["factor"]["factor_flops"]
        if "chol_factor_flow" in rec["phases_ms"]:
            rec["factor_tflops"] = f / (rec["phases_ms"]["chol_factor_flow"] * 1e-3) / 1e12
    except Exception as e:  # report the limit that stopped it
        rec["error"] = repr(e)
    opt.close()
    print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    nd = int(os.environ.get("ND_LEVELS", "0"))
    for a in sys.argv[1:]:
        run(int(a), nd=nd)
