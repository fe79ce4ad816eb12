"""Regenerates tests/golden/slam2d.npz / slam3d.npz: the oracle's LM trajectory on the seeded synthetic landmark-SLAM graphs
(openslam_g2o_b200/synth.py: landmark_slam_2d(), landmark_slam_3d() with their default arguments; nothing marginalized, block
ordering).  The reference ships no landmark datasets, so these are oracle outputs pinned against drift: a CPU test checks that
the oracle still reproduces them, the GPU parity test compares the device trajectory with them as well.
usage: python tests/golden/make_golden_synth.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle_binding import LM, Oracle, fnv1a64  # noqa: E402
from openslam_g2o_b200 import synth  # noqa: E402

ITERATIONS = 8


def trajectory(prob):
    """what the fixtures hold, computed by the oracle"""
    o = Oracle()
    synth.feed(prob, o)
    gauge = o.setup_cli(False)
    o.set_block_ordering(True)
    assert o.initialize_optimization()
    n, st = o.optimize(LM, ITERATIONS)
    ids = np.concatenate([prob["pose_ids"], prob["lm_ids"]]).astype(np.int32)
    est = np.stack([np.pad(o.vertex_estimate(int(i)), (0, 12))[:12] for i in ids])
    return dict(gauge=np.int32(gauge), iterations=np.int32(n), chi2=np.array([s.chi2 for s in st[:n]]),
                lam=np.array([s.lambda_ for s in st[:n]]), lev=np.array([s.levenberg_iterations for s in st[:n]], np.int32),
                perm_hash=np.array(fnv1a64(o.block_perm())), lnz=np.int64(o.lnz()), final_ids=ids, final_est=est)


def problems():
    return {"slam2d": synth.landmark_slam_2d(), "slam3d": synth.landmark_slam_3d()}


def main():
    for name, prob in problems().items():
        out = trajectory(prob)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, int(out["iterations"]), out["chi2"][-1], str(out["perm_hash"]), int(out["lnz"]))


if __name__ == "__main__":
    main()
