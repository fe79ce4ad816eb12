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
