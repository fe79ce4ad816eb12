"""debug helper: cycle counters inside the tail-chain kernels (instrumented build:
   make -C openslam_g2o_b200/csrc timing ; G2O_B200_LIB=openslam_g2o_b200/libg2o_b200_timing.so python tests/chain_timing.py venice)"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import openslam_g2o_b200 as g
from openslam_g2o_b200 import synth
wl = sys.argv[1] if len(sys.argv) > 1 else "venice"
p = synth.venice_like() if wl == "venice" else synth.venice_like(10000, 2000000, seed=10000, fixed_obs=10)
opt = g.SparseOptimizer(device=0); opt.set_algorithm("lm_fix6_3"); synth.feed(p, opt); opt.setup_cli(); opt.initialize_optimization()
opt.optimize(2)
out = (C.c_ulonglong * 32)()
ctx = opt.context
ctx.build_system(); ctx.set_lambda(1e-3)
ctx.solve(); ctx.synchronize(); g.lib.b200_debug_chain_timing(out, 1)
ctx.solve(); ctx.synchronize(); g.lib.b200_debug_chain_timing(out, 1)
names = ["link prologue (wait panel + add)", "step: top -> barrier A (tid 0)", "step: barrier A -> B (column solve)", "step: own update (tid 0)",
         "pivot block (pivot thread)", "step: own update (tid 448)", "re-index", "backward: wait for data", "backward: x_J + rest (after GEMV1 barrier)", "backward: request + re-index of x", "backward: GEMV1 (tid 0)", "backward: GEMV1 barrier wait"]
info = ctx.factor_info()
print(wl, info)
print("one factorisation + solve: total cycles, events, cycles/event (us at 1.965 GHz)")
for i, n in enumerate(names):
    c, k = out[i], out[16 + i]
    print("  %-38s %12d %7d %10.0f  (%.3f us)" % (n, c, k, c / max(k, 1), c / max(k, 1) / 1965.0))
