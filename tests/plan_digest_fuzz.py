"""Regression aid for host-side changes of the structure phase (no GPU): prints b200_debug_upload_digest - a digest of
the complete device-side plan - for random bundle-adjustment graphs (duplicates, shuffled edges, fixed points and
cameras, landmark shards, wide landmarks, both camera models) and random SE3 pose graphs (AMD and nested dissection).

    G2O_B200_LIB=/path/to/old/libg2o_b200.so python tests/plan_digest_fuzz.py > old.txt
    G2O_B200_HOST_GRAIN=4 G2O_B200_HOST_THREADS=3 python tests/plan_digest_fuzz.py > new.txt ; diff old.txt new.txt
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openslam_g2o_b200 import lib, synth  # noqa: E402
from openslam_g2o_b200.optimizer import SparseOptimizer  # noqa: E402


def digest(p, shard=0, shards=1, fixed=(), nd=0):
    o = SparseOptimizer(device=-1, shard=shard, num_shards=shards)
    synth.feed(p, o)
    for f in fixed:
        o.set_fixed(int(f))
    o.setup_cli()
    o.initialize_optimization()
    if nd:
        o.context.set_ordering(nd)
    o._ensure_uploaded()
    lib.b200_debug_upload_digest(1)
    try:
        o.context.build_structure()
    except Exception as e:  # e.g. nothing left to optimise: must fail identically in both builds
        return "ERR " + str(e)[:60]
    return "%016x" % lib.b200_debug_upload_digest(1)


def main():
    rng = np.random.default_rng(123)
    for case in range(80):
        cams = int(rng.integers(3, 60))
        pts = int(rng.integers(5, 1500))
        seed = int(rng.integers(1, 10 ** 6))
        fo = int(rng.integers(2, min(cams, 40) + 1)) if case % 4 == 0 else None
        p = synth.venice_like(cams, pts, seed=seed, fixed_obs=fo) if case % 3 else synth.expmap_ba(cams, pts, seed=seed)
        n = len(p["edge_v0"])
        idx = np.concatenate([np.arange(n), rng.integers(0, n, int(rng.integers(0, n // 5 + 1)))])
        rng.shuffle(idx)
        p["edge_v0"], p["edge_v1"], p["edge_payload"] = p["edge_v0"][idx], p["edge_v1"][idx], p["edge_payload"][idx]
        fixed = list(rng.choice(p["point_ids"], int(rng.integers(0, pts // 3 + 1)), replace=False)) + \
            list(rng.choice(p["cam_ids"], int(rng.integers(0, 3)), replace=False))
        shards = int(rng.choice([1, 1, 2, 3, 8]))
        shard = int(rng.integers(0, shards))
        print("ba", case, cams, pts, shards, digest(p, shard, shards, fixed))
    rng = np.random.default_rng(7)
    for case in range(40):
        npl, laps, nd = int(rng.integers(4, 60)), int(rng.integers(3, 40)), int(rng.choice([0, 0, 2, 4]))
        print("se3", case, npl, laps, nd, digest(synth.sphere(npl, laps, seed=case + 1), nd=nd))


if __name__ == "__main__":
    main()
