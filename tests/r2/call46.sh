#!/bin/bash
# round 2, session 3, call 7: sparse-inverse sweep with the column-parallel diagonal kernel: parity + timing
out=gpurun_out/r2b
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_widening_landmark_slam.py -q -m gpu -k "marginals or landmark_slam_matches" 2>&1 | tail -15 > $out/c46_pytest.txt
tail -5 $out/c46_pytest.txt
timeout 300 python - <<'PY' 2>&1 | tail -5
import sys, time
sys.path.insert(0, ".")
import numpy as np
import openslam_g2o_b200 as g
from openslam_g2o_b200 import synth
opt = g.SparseOptimizer(device=0); opt.set_algorithm("lm_fix6_3"); synth.feed(synth.sphere(), opt)
opt.setup_cli(); opt.initialize_optimization(); opt._ensure_uploaded()
ctx = opt.context; assert ctx.build_structure(); ctx.compute_active_errors(); ctx.build_system()
nb = ctx.dims()["numPoses"]; diag = [(i, i) for i in range(nb)]
ctx.compute_marginals(diag[:4])
for _ in range(3):
    t = time.perf_counter(); m = ctx.compute_marginals(diag); print("all diagonal blocks: %.2f ms" % ((time.perf_counter() - t) * 1e3))
PY
