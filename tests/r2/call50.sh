#!/bin/bash
# round 2, session 3: compute-sanitizer memcheck over smoke() (pose graph, BA with the packets linearisation, expmap BA,
# 2D landmark SLAM + sparse-inverse marginals) and over the forced wide-landmark Schur path
out=gpurun_out/r2b
mkdir -p $out
timeout 240 compute-sanitizer --tool memcheck --print-limit 5 python -c "import __graft_entry__ as g; g.smoke()" > $out/c50_memcheck_smoke.txt 2>&1
tail -4 $out/c50_memcheck_smoke.txt
G2O_B200_SR_WIDE=3 timeout 120 compute-sanitizer --tool memcheck --print-limit 5 python - > $out/c50_memcheck_wide.txt 2>&1 <<'PY'
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import openslam_g2o_b200 as g
from openslam_g2o_b200 import synth
from oracle_binding import LM, Oracle
p = synth.venice_like(20, 300, seed=4)
opt = g.SparseOptimizer(device=0); opt.set_algorithm("lm_fix6_3"); o = Oracle()
synth.feed(p, opt); synth.feed(p, o)
opt.set_edge_robust_kernel(range(0, len(p["edge_v0"]), 3), "Huber", 2.0); o.set_edge_robust_kernel(range(0, len(p["edge_v0"]), 3), "Huber", 2.0)
opt.setup_cli(); o.setup_cli(True); opt.initialize_optimization(); o.initialize_optimization()
n = opt.optimize(3); no, st = o.optimize(LM, 3)
cg = np.array([s.chi2 for s in opt.batch_statistics]); co = np.array([s.chi2 for s in st[:no]])
assert np.abs(cg - co).max() <= 1e-6 * co.max(), (cg, co)
print("wide path + per-edge kernels under memcheck: chi2", cg[-1], "segments", opt.context.factor_info()["schur_segments"])
PY
tail -4 $out/c50_memcheck_wide.txt
