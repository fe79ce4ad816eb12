"""debug helper: phase cycle counts inside chol factor_chunk (needs the -DCHOL_TIMING build)"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import openslam_g2o_b200 as g
from openslam_g2o_b200 import synth
wl = sys.argv[1] if len(sys.argv) > 1 else "venice"
p = synth.venice_like() if wl == "venice" else synth.sphere()
opt = g.SparseOptimizer(device=0); opt.set_algorithm("lm_fix6_3"); synth.feed(p, opt); opt.setup_cli(); opt.initialize_optimization()
opt.optimize(2)
out = (C.c_ulonglong * 8)()
g.lib.b200_debug_chol_timing(out, 1)
ctx = opt.context
ctx.build_system(); ctx.set_lambda(1e-3); 
import time
ctx.solve(); g.lib.b200_debug_chol_timing(out, 1)
ctx.solve(); ctx.synchronize(); g.lib.b200_debug_chol_timing(out, 1)
names = ["load", "pivot", "trsm", "barrierA", "update", "barrierB", "store"]
tot = sum(out[i] for i in range(7))
print(wl, "one factorisation, thread 0 of every chunk CTA, cycles:")
for i, n in enumerate(names): print("  %-9s %10d  %5.1f%%" % (n, out[i], 100.0 * out[i] / max(tot, 1)))
print("  total %d cycles = %.1f us summed over chunk CTAs" % (tot, tot / 1.965e3))
