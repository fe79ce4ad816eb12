"""host-side probe (no GPU work): does the two-phase .g2o loader scale with threads on this box?
usage: python tests/loader_scaling.py [points]   -> prints load seconds per thread count"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openslam_g2o_b200 import synth  # noqa: E402
from openslam_g2o_b200.optimizer import SparseOptimizer  # noqa: E402

points = int(sys.argv[1]) if len(sys.argv) > 1 else 120000
path = "/tmp/loader_scaling.g2o"
synth.write_g2o(synth.venice_like(871, points, seed=871), path)
size = os.path.getsize(path) / 1e6
print("file %.1f MB, host cores %d" % (size, os.cpu_count()))
for threads in (1, 2, 4, 8, 16):
    os.environ["G2O_B200_LOADER_THREADS"] = str(threads)
    best = 1e9
    for _ in range(3):
        o = SparseOptimizer(device=-1)
        t = time.perf_counter()
        o.load(path)
        best = min(best, time.perf_counter() - t)
        o.close()
    print("threads %2d: %.3f s  %.0f MB/s" % (threads, best, size / best))
