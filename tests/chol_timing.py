"""debug helper: cycle counters inside the dataflow Cholesky kernel (needs the instrumented build:
   make -C openslam_g2o_b200/csrc timing ; G2O_B200_LIB=openslam_g2o_b200/libg2o_b200_timing.so python tests/chol_timing.py venice)"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import openslam_g2o_b200 as g
from openslam_g2o_b200 import synth
wl = sys.argv[1] if len(sys.argv) > 1 else "venice"
p = synth.venice_like() if wl == "venice" else synth.sphere()
opt = g.SparseOptimizer(device=0); opt.set_algorithm("lm_fix6_3"); synth.feed(p, opt); opt.setup_cli(); opt.initialize_optimization()
opt.optimize(2)
out = (C.c_ulonglong * 32)()
ctx = opt.context
ctx.build_system(); ctx.set_lambda(1e-3)
ctx.solve(); ctx.synchronize(); g.lib.b200_debug_chol_timing(out, 1)
ctx.solve(); ctx.synchronize(); g.lib.b200_debug_chol_timing(out, 1)
names = ["item wait", "item compute", "chunk panel load", "chunk factor", "chunk signal", "chunk inverse", "chunk wait", "rtile wait", "chunk rhs gather", "chunk stores", "chunk contrib", "item staging", "item product"]
print(wl, "one factorisation, thread 0 of every CTA: total cycles, events, cycles/event (us at 1.965 GHz)")
for i, n in enumerate(names):
    c, k = out[i], out[16 + i]
    print("  %-14s %12d %7d %10.0f  (%.2f us)" % (n, c, k, c / max(k, 1), c / max(k, 1) / 1965.0))
print(ctx.factor_info())
