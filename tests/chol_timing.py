"""debug helper: cycle counters inside the dataflow Cholesky kernel (needs the instrumented build:
   make -C openslam_g2o_b200/csrc timing ; G2O_B200_LIB=openslam_g2o_b200/libg2o_b200_timing.so python tests/chol_timing.py venice)"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import openslam_g2o_b200 as g
from openslam_g2o_b200 import synth
wl = sys.argv[1] if len(sys.argv) > 1 else "venice"
if wl.startswith("sphere") and len(wl) > 6:
    n = int(wl[6:]); p = synth.sphere(n, n, seed=n * n)   # config-5 family: n x n poses
else:
    p = synth.venice_like() if wl == "venice" else synth.sphere()
opt = g.SparseOptimizer(device=0); opt.set_algorithm("lm_fix6_3"); synth.feed(p, opt); opt.setup_cli(); opt.initialize_optimization()
opt.optimize(2)
out = (C.c_ulonglong * 32)()
ctx = opt.context
ctx.build_system(); ctx.set_lambda(1e-3)
stamps = (C.c_ulonglong * (6 * 4096))()
ctx.solve(); ctx.synchronize(); g.lib.b200_debug_chol_timing(out, 1); g.lib.b200_debug_chol_stamps(stamps, 1)
ctx.solve(); ctx.synchronize(); g.lib.b200_debug_chol_timing(out, 1); g.lib.b200_debug_chol_stamps(stamps, 1)
names = ["item wait", "item compute", "chunk panel load", "chunk factor", "chunk signal", "chunk inverse", "chunk wait", "wide: poll", "chunk rhs gather", "chunk stores", "chunk contrib", "item staging", "item product", "wide: own copies landed", "wide: stage barrier", "-"]
print(wl, "one factorisation, thread 0 of every CTA: total cycles, events, cycles/event (us at 1.965 GHz)")
for i, n in enumerate(names):
    c, k = out[i], out[16 + i]
    print("  %-14s %12d %7d %10.0f  (%.2f us)" % (n, c, k, c / max(k, 1), c / max(k, 1) / 1965.0))
tot = sum(out[i] for i in (0, 11, 12, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10))
print(ctx.factor_info())
ctx.set_profiling(True); ctx.solve(); ctx.synchronize(); ph = ctx.phase_times(); ctx.set_profiling(False)
print({k: round(1e3 * v[0], 3) for k, v in ph.items() if v[1] > 0})

import numpy as np
st = np.array(list(stamps), dtype=np.uint64).reshape(6, 4096).astype(np.float64)
n = ctx.factor_info()["supernodes"]
saw, sig, lastupd, firstuse = st[0, :n], st[1, :n], st[2, :n], st[3, :n]
t0 = sig[sig > 0].min()
ok = (firstuse < 1e19) & (sig > 0)
print("chunk signal -> first consumer sees it: median %.2f us" % np.median((firstuse - sig)[ok] / 1e3))
ok2 = (lastupd > 0) & (saw > 0)
print("last update signal -> chunk sees it:   median %.2f us" % np.median((saw - lastupd)[ok2] / 1e3))
print("chunk busy (saw -> signalled):         median %.2f us" % np.median((sig - saw)[saw > 0] / 1e3))
lastsaw, proddone = st[5, :n], st[4, :n]
ok3 = ok & (lastsaw > 0)
print("chunk signal -> LAST consumer sees it: median %.2f us" % np.median((lastsaw - sig)[ok3] / 1e3))
ok4 = (proddone > 0) & (lastupd > 0)
print("last product done -> last update signal: median %.2f us" % np.median((lastupd - proddone)[ok4] / 1e3))
# chain view: supernode J is updated by J-1 (same chain: J-2 on the two-chain Venice graph); report both
for dJ in (1, 2):
    a = proddone[dJ:] - lastsaw[:-dJ]
    m = (proddone[dJ:] > 0) & (lastsaw[:-dJ] > 0) & (a > 0) & (a < 5e4)
    if m.any(): print("last consumer saw K -> products of K+%d done: median %.2f us (n=%d)" % (dJ, np.median(a[m]) / 1e3, m.sum()))
print("whole factorisation (first chunk signal -> last): %.1f us" % ((sig.max() - t0) / 1e3))
