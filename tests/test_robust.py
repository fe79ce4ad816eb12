"""Robust kernels: the oracle's restatement pinned against the closed forms of
core/robust_kernel_impl.cpp, the product against the oracle on graphs with outliers."""
import ctypes as C

import numpy as np
import pytest
from conftest import needs_oracle
from helpers import load_fixture, feed_fixture, rel_err

KERNELS = ("Huber", "PseudoHuber", "Cauchy", "Saturated", "DCS")


def _closed_form(name, d, e):
    """rho(e), rho'(e), rho''(e) written independently from robust_kernel_impl.cpp:65-126 (e = squared error)"""
    if name == "Huber":
        return (e, 1.0, 0.0) if e <= d * d else (2 * np.sqrt(e) * d - d * d, d / np.sqrt(e), -0.5 * d / e ** 1.5)
    if name == "PseudoHuber":
        a = 1 + e / (d * d)
        return 2 * d * d * (np.sqrt(a) - 1), 1 / np.sqrt(a), -0.5 / (d * d) / a ** 1.5
    if name == "Cauchy":
        a = 1 + e / (d * d)
        return d * d * np.log(a), 1 / a, -1 / (d * d) / a ** 2
    if name == "Saturated":
        return (e, 1.0, 0.0) if e <= d * d else (d * d, 0.0, 0.0)
    s = min(1.0, 2 * d / (d + e))
    return s * s * e, s * s, 0.0


@needs_oracle
def test_oracle_kernels_match_closed_forms():
    from oracle_binding import Oracle
    L = Oracle().L
    L.oracle_robustify.argtypes = [C.c_int, C.c_double, C.c_double, C.c_void_p]
    rng = np.random.default_rng(3)
    for k, name in enumerate(KERNELS, start=1):
        for _ in range(200):
            d = float(rng.uniform(0.1, 5.0))
            e = float(rng.choice([rng.uniform(0, d * d), rng.uniform(d * d, 100 * d * d), d * d]))
            rho = np.zeros(3)
            L.oracle_robustify(k, d, e, rho.ctypes.data)
            ref = np.array(_closed_form(name, d, e))
            assert np.allclose(rho, ref, rtol=1e-13, atol=1e-300), (name, d, e, rho, ref)
    # derivative consistency: rho' is the derivative of rho (finite differences), away from the kinks
    for k, name in enumerate(KERNELS, start=1):
        if name in ("Saturated", "DCS"):
            continue  # piecewise / not a true derivative pair in the reference either
        d, e, h = 1.3, 4.0, 1e-6
        r0, r1, r2 = np.zeros(3), np.zeros(3), np.zeros(3)
        L.oracle_robustify(k, d, e - h, r0.ctypes.data); L.oracle_robustify(k, d, e, r1.ctypes.data); L.oracle_robustify(k, d, e + h, r2.ctypes.data)
        assert abs((r2[0] - r0[0]) / (2 * h) - r1[1]) < 1e-8
        assert abs((r2[1] - r0[1]) / (2 * h) - r1[2]) < 1e-8


@needs_oracle
def test_oracle_wide_kernel_is_no_kernel():
    """a Huber kernel wider than every residual must reproduce the golden (kernel-free) chi2 sequence"""
    from oracle_binding import Oracle, GN
    fx = load_fixture("intel")
    o = Oracle()
    feed_fixture(o, fx)
    o.set_robust_kernel("Huber", 1e12)
    o.setup_cli(True)
    o.initialize_optimization()
    n, st = o.optimize(GN, int(fx["iterations"]))
    assert n == int(fx["done"])
    assert rel_err([s.chi2 for s in st[:n]], fx["chi2"]) < 1e-12


def _outlier_graph(kind, seed):
    """small synthetic problem with a few gross outliers in the measurements"""
    from openslam_g2o_b200 import synth
    rng = np.random.default_rng(seed)
    if kind == "se3":
        p = synth.sphere(12, 8, seed=seed)
        pay = p["edge_payload"].copy()
        bad = rng.choice(len(pay), 12, replace=False)
        pay[bad, :3] += rng.normal(0, 20.0, (12, 3))
    else:
        p = synth.venice_like(12, 400, seed=seed)
        pay = p["edge_payload"].copy()
        bad = rng.choice(len(pay), 40, replace=False)
        pay[bad] += rng.normal(0, 80.0, (40, 2))
    p = dict(p)
    p["edge_payload"] = pay
    return p


@pytest.mark.gpu
@needs_oracle
@pytest.mark.parametrize("kind", ["se3", "ba"])
@pytest.mark.parametrize("kernel,width", [("Huber", 1.0), ("PseudoHuber", 2.0), ("Cauchy", 1.5), ("Saturated", 3.0), ("DCS", 2.0)])
def test_gpu_robust_lm_matches_oracle(kind, kernel, width):
    import openslam_g2o_b200 as g
    from openslam_g2o_b200 import synth
    from oracle_binding import LM, Oracle
    prob = _outlier_graph(kind, 11)
    opt = g.SparseOptimizer(device=0)
    opt.set_algorithm("lm_fix6_3")
    o = Oracle()
    synth.feed(prob, opt)
    synth.feed(prob, o)
    opt.set_robust_kernel(kernel, width)
    o.set_robust_kernel(kernel, width)
    assert opt.setup_cli() == o.setup_cli(True)
    opt.initialize_optimization()
    o.initialize_optimization()
    # robust chi2 of the initial state, then 6 LM iterations
    chi0_g = opt.compute_active_errors()
    o.algorithm_init(); o.build_structure()
    chi0_o = o.compute_active_errors()
    assert abs(chi0_g - chi0_o) <= 1e-10 * chi0_o
    n = opt.optimize(6)
    o2 = Oracle()
    synth.feed(prob, o2)
    o2.set_robust_kernel(kernel, width)
    o2.setup_cli(True); o2.initialize_optimization()
    no, st = o2.optimize(LM, 6)
    assert n == no
    chi_g = np.array([s.chi2 for s in opt.batch_statistics])
    chi_o = np.array([s.chi2 for s in st[:no]])
    assert rel_err(chi_g, chi_o) < 1e-6, (kernel, chi_g, chi_o)
    # the kernel is doing something: the robust cost is well below the plain one
    plain = g.SparseOptimizer(device=0)
    plain.set_algorithm("lm_fix6_3")
    synth.feed(prob, plain)
    plain.setup_cli(); plain.initialize_optimization()
    assert chi0_g < 0.9 * plain.compute_active_errors()


def _per_edge_problem(kind):
    from openslam_g2o_b200 import synth
    if kind == "slam2d":
        p = dict(synth.landmark_slam_2d(60, 30, seed=13))
        rng = np.random.default_rng(13)
        pay = p["obs_payload"].copy()
        bad = rng.choice(len(pay), 25, replace=False)
        pay[bad, :2] += rng.normal(0, 3.0, (25, 2))
        p["obs_payload"] = pay
        n_odo, n_obs = len(p["odo_v0"]), len(p["obs_v0"])
        # edges are added odometry first, then sightings (synth.feed): kernels on the sightings and the loop closures only
        sel = {"Huber": list(range(59, n_odo)) + list(range(n_odo, n_odo + n_obs, 2)), "Cauchy": list(range(n_odo + 1, n_odo + n_obs, 4))}
        return p, "lm_var", False, sel
    p = _outlier_graph(kind, 17)
    E = len(p["edge_v0"])
    sel = {"Huber": list(range(0, E, 3)), "DCS": list(range(1, E, 5))}
    return p, "lm_fix6_3", True, sel


@pytest.mark.gpu
@needs_oracle
@pytest.mark.parametrize("kind", ["se3", "ba", "slam2d"])
def test_gpu_per_edge_robust_kernels_match_oracle(kind):
    """Edge::setRobustKernel on individual edges (a kernel on loop closures / sightings only, different kernels and widths on
    different edges; the others carry none): chi2 of the initial state and the LM trajectory against the oracle, whose edges
    carry their own rkKind / rkDelta like the reference's (core/base_binary_edge.hpp:91-113)"""
    import openslam_g2o_b200 as g
    from openslam_g2o_b200 import synth
    from oracle_binding import LM, Oracle
    prob, name, marg, sel = _per_edge_problem(kind)
    widths = {"Huber": 1.5, "Cauchy": 2.0, "DCS": 2.5}

    def make(cls):
        t = cls() if cls is Oracle else cls(device=0)
        if cls is not Oracle:
            t.set_algorithm(name)
        synth.feed(prob, t)
        for kname, idx in sel.items():
            t.set_edge_robust_kernel(idx, kname, widths[kname])
        return t
    opt, o = make(g.SparseOptimizer), make(Oracle)
    assert opt.setup_cli() == o.setup_cli(marg)
    o.set_block_ordering(True)
    opt.initialize_optimization(); o.initialize_optimization()
    chi0_g = opt.compute_active_errors()
    o.algorithm_init(); o.build_structure()
    chi0_o = o.compute_active_errors()
    assert abs(chi0_g - chi0_o) <= 1e-10 * chi0_o
    n = opt.optimize(6)
    o2 = make(Oracle)
    o2.setup_cli(marg); o2.set_block_ordering(True); o2.initialize_optimization()
    no, st = o2.optimize(LM, 6)
    assert n == no
    chi_g = np.array([s.chi2 for s in opt.batch_statistics]); chi_o = np.array([s.chi2 for s in st[:no]])
    assert rel_err(chi_g, chi_o) < 1e-6, (chi_g, chi_o)
    # not the uniform-kernel answer and not the plain one
    plain = g.SparseOptimizer(device=0)
    plain.set_algorithm(name)
    synth.feed(prob, plain)
    plain.setup_cli(); plain.initialize_optimization()
    assert chi0_g < 0.98 * plain.compute_active_errors()


def test_per_edge_robust_kernels_reach_the_context_in_edge_order():
    """host side of the same: the graph hands (kind, width) per edge of each edge set to the context in hand-over order;
    wrong lengths and unknown kernels are refused"""
    import ctypes as C
    import openslam_g2o_b200 as g
    from openslam_g2o_b200 import synth
    from openslam_g2o_b200._lib import lib, ptr
    p = synth.landmark_slam_2d(20, 10, seed=3)
    opt = g.SparseOptimizer(device=-1)
    opt.set_algorithm("lm_var")
    synth.feed(p, opt)
    opt.set_edge_robust_kernel([0, 5, len(p["odo_v0"]) + 2], "Huber", 2.0)
    opt.setup_cli(); opt.initialize_optimization(); opt._ensure_uploaded()
    assert opt.context.build_structure()
    h = opt.context.handle
    n_obs = len(p["obs_v0"])
    kinds = np.ones(n_obs, np.uint8); deltas = np.ones(n_obs)
    assert lib.b200_set_edge_robust_kernels(h, g.EDGE_SE2_XY, n_obs, ptr(kinds), ptr(deltas)) == 0
    assert lib.b200_set_edge_robust_kernels(h, g.EDGE_SE2_XY, n_obs - 1, ptr(kinds), ptr(deltas)) < 0      # one entry per edge
    kinds[3] = 9
    assert lib.b200_set_edge_robust_kernels(h, g.EDGE_SE2_XY, n_obs, ptr(kinds), ptr(deltas)) < 0          # unknown kernel
    assert lib.b200_set_edge_robust_kernels(h, g.EDGE_SE2_XY, 0, None, None) == 0                          # back to the uniform one
    with pytest.raises(g.B200Error):
        opt.set_edge_robust_kernel([10 ** 6], "Huber", 1.0)
