"""The oracle against every known answer the reference side offers for this path and against
the committed golden fixtures.  CPU only."""
import numpy as np
import pytest
from conftest import REFERENCE_DATA, needs_oracle, needs_reference
from helpers import FIXTURES, feed_fixture, load_fixture, rel_err

# produced by the reference's vendored CSparse (cs_amd / cs_schol) in the survey session
KNOWN = {
    "manhattan3500": dict(file="2d/manhattan3500/manhattanOlson3500.g2o", n=10497, lnz=187431, lnz_scalar=187440,
                          p8=[1077, 1079, 1078, 1080, 1076, 1082, 1081, 1144], fnv="75f08a96a583d76a", blocks=8949),
    "sphere_bignoise": dict(file="3d/sphere/sphere_bignoise_vertex3.g2o", n=13194, lnz=2514663, lnz_scalar=2514663,
                            p8=[698, 700, 699, 797, 799, 798, 748, 749], fnv="896d455aadd297a8", blocks=10843),
    "garage": dict(file="3d/garage/parking-garage.g2o", n=9960, lnz=414804, lnz_scalar=415344,
                   p8=[513, 656, 512, 657, 511, 658, 510, 659], fnv="f2171840e4864b87", blocks=7934),
    "intel": dict(file="2d/intel/intel.g2o", n=2826, lnz=47790, lnz_scalar=47862,
                  p8=[77, 693, 886, 888, 889, 890, 892, 891], fnv="d38cc4fe651b8958", blocks=2772),
}


@needs_oracle
@needs_reference
@pytest.mark.parametrize("name", sorted(KNOWN))
def test_known_answers_from_dataset_files(name):
    from oracle_binding import Oracle, fnv1a64
    k = KNOWN[name]
    o = Oracle()
    assert o.load(REFERENCE_DATA + "/" + k["file"])
    assert o.setup_cli(True) == 0  # gauge = vertex id 0
    assert o.initialize_optimization()
    o.algorithm_init()
    o.build_structure()
    o.compute_active_errors()
    o.build_system()
    o.set_lambda(0.0)
    assert o.solve()
    d = o.dims()
    assert d["sizePoses"] == k["n"]
    assert o.blocks(0)[0].shape[0] == k["blocks"]
    perm = o.block_perm()
    assert list(perm[:8]) == k["p8"]
    assert fnv1a64(perm) == k["fnv"]
    assert o.lnz() == k["lnz"]
    assert o.scalar_amd_lnz() == k["lnz_scalar"]


@needs_oracle
@pytest.mark.parametrize("name", FIXTURES)
def test_oracle_reproduces_golden_fixture(name):
    """the committed fixtures are what the oracle computes from the same inputs (guards against drift)"""
    from oracle_binding import Oracle, fnv1a64
    fx = load_fixture(name)
    o = Oracle()
    feed_fixture(o, fx)
    assert o.setup_cli(True) == int(fx["gauge"])
    assert o.initialize_optimization()
    n, st = o.optimize(int(fx["algorithm"]), int(fx["iterations"]))
    assert n == int(fx["done"])
    assert rel_err([s.chi2 for s in st], fx["chi2"]) < 1e-12
    assert fnv1a64(o.block_perm()) == str(fx["perm_hash"])
    assert o.lnz() == int(fx["lnz"])
    est = np.stack([np.pad(o.vertex_estimate(i), (0, 12))[:12] for i in fx["final_ids"]])
    assert rel_err(est, fx["final_est"]) < 1e-12


@needs_oracle
def test_oracle_jacobians_match_numeric_differences():
    """the reference's own self-consistency test (types/slam3d/test_slam3d_jacobian.cpp:100-149): analytic
    linearizeOplus vs central differences, allowed difference 1e-6 - applied to the oracle's quadratic form:
    b = -J^T Omega e must equal the numeric gradient of chi2/2 w.r.t. the oplus increment"""
    from oracle_binding import LM, Oracle
    from openslam_g2o_b200 import synth
    p = synth.sphere(6, 4, seed=3)
    o = Oracle()
    synth.feed(p, o)
    o.setup_cli(True)
    o.initialize_optimization()
    o.algorithm_init()
    o.build_structure()
    chi0 = o.compute_active_errors()
    o.build_system()
    b = o.b()
    # numeric gradient of chi2 for a handful of coordinates, via oplus on x = +-delta * e_k
    n = len(b)
    rng = np.random.default_rng(0)
    delta = 1e-6
    for k in rng.choice(n, 12, replace=False):
        g = 0.0
        for sgn in (+1, -1):
            o.push()
            x = np.zeros(n)
            x[k] = sgn * delta
            o.set_x(x)
            o.update()
            g += sgn * o.compute_active_errors()
            o.pop()
        g /= 2 * delta
        assert abs(-0.5 * g - b[k]) <= 1e-5 * max(1.0, abs(b[k])), (k, g, b[k])
    assert chi0 > 0


def test_marginal_covariance_recursion_matches_dense_inverse():
    """oracle restatement of LinearSolverCSparse::solvePattern + MarginalCovarianceCholesky (the reference's sparse
    inverse recursion on the real CSparse factor) against the dense inverse of the same Hpp"""
    from helpers import load_fixture, feed_fixture
    from oracle_binding import Oracle
    for name in ("intel",):  # the recursion memoises O(nnz(L)) entries: seconds on intel, minutes on the sphere
        fx = load_fixture(name)
        o = Oracle()
        feed_fixture(o, fx)
        o.setup_cli(True); o.initialize_optimization(); o.algorithm_init()
        assert o.build_structure()
        o.compute_active_errors(); o.build_system()
        rows, cols, vals = o.blocks(0)
        d = vals.shape[1]
        nb = int(max(cols)) + 1
        A = np.zeros((nb * d, nb * d))
        for r, c, v in zip(rows, cols, vals):
            A[r * d:(r + 1) * d, c * d:(c + 1) * d] = v
            A[c * d:(c + 1) * d, r * d:(r + 1) * d] = v.T
        inv = np.linalg.inv(A)
        pairs = [(0, 0), (3, 3), (2, 5), (nb - 1, nb - 1), (0, nb - 1), (7, 2)]
        got = o.compute_marginals(pairs)
        assert got is not None
        for (r, c), blk in zip(pairs, got):
            assert np.abs(blk - inv[r * d:(r + 1) * d, c * d:(c + 1) * d]).max() <= 1e-8 * np.abs(inv).max()
