"""profiling helper: a few LM iterations of one bench workload (for `ncu ... python tests/prof_run.py venice`)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import openslam_g2o_b200 as g
from openslam_g2o_b200 import synth
wl = sys.argv[1] if len(sys.argv) > 1 else "venice"
if wl.startswith("sphere") and len(wl) > 6 and wl != "sphere2500":   # sphereN = N x N poses; sphere2500 = the default 50 x 50
    n = int(wl[6:]); p = synth.sphere(n, n, seed=n * n)   # config-5 family: n x n poses
else:
  p = synth.venice_like() if wl == "venice" else synth.venice_like(10000, 2000000, seed=10000, fixed_obs=10) if wl == "ba10k" else synth.sphere()
opt = g.SparseOptimizer(device=0); opt.set_algorithm("lm_fix6_3"); synth.feed(p, opt); opt.setup_cli(); opt.initialize_optimization()
opt.context.set_profiling(True)  # plain launches (no CUDA-graph replay) so that every kernel shows up by name
print("iterations", opt.optimize(int(sys.argv[2]) if len(sys.argv) > 2 else 3))
