"""The synthetic workloads of bench.py are pinned (the reference's sphere generator uses a tr1 RNG that
cannot be reproduced, so the inputs come from numpy's PCG64 with fixed seeds and their SHA-256 is recorded here): a bench
number always refers to exactly these graphs."""
import hashlib

import numpy as np


def workload_sha256(p):
    m = hashlib.sha256()
    for k in sorted(p):
        v = p[k]
        if isinstance(v, np.ndarray):
            m.update(k.encode())
            m.update(str(v.dtype).encode())
            m.update(str(v.shape).encode())
            m.update(np.ascontiguousarray(v).tobytes())
    return m.hexdigest()


def test_bench_workloads_are_pinned():
    from openslam_g2o_b200 import synth
    # configs[1]: sphere2500 (create_sphere.cpp defaults: 50 nodes per level, 50 laps, radius 100), seed 2500
    p = synth.sphere()
    assert len(p["vertex_ids"]) == 2500 and len(p["edge_v0"]) == 9799
    assert workload_sha256(p) == "0b5c00b7d0ffd5bcb6624ff1e3df4b5c108d3a4748142158730c2305c4e595c8"
    # configs[2]: Venice-shaped BA, seed 871
    p = synth.venice_like()
    assert len(p["cam_ids"]) == 871 and len(p["point_ids"]) == 530304 and len(p["edge_v0"]) == 2014827
    assert workload_sha256(p) == "0251b8b1653b3e062b67117231836a5819229fb1cebf06b3307614da963e1684"


def test_generators_follow_the_reference_layout():
    """sphere: 2499 odometry edges + 3 * 49 * 50 - 50 = 7300 loop closures (create_sphere.cpp:131-147); the BA
    generator gives every point at least two observations inside a window of consecutive cameras"""
    from openslam_g2o_b200 import synth
    p = synth.sphere()
    d = p["edge_v1"] - p["edge_v0"]
    assert int((d == 1).sum()) >= 2499 and len(d) - 2499 == 7300
    q = synth.venice_like(50, 2000, seed=1)
    counts = np.bincount(q["edge_v0"] - 50, minlength=2000)
    assert counts.min() >= 2
    # expmap form of the same scene: same structure, pixel coordinates shifted by the principal point
    e = synth.expmap_ba(50, 2000, seed=1)
    assert np.array_equal(e["edge_v0"], q["edge_v0"]) and np.array_equal(e["edge_v1"], q["edge_v1"])
    assert np.allclose(e["edge_payload"][:, 1:3], q["edge_payload"] + np.array([320.0, 240.0]))
