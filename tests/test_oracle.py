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


@needs_oracle
def test_oracle_optima_match_an_independent_solver():
    """SE3 pose graph and SBACam bundle adjustment: the oracle's converged chi2 equals the minimum that
    scipy.optimize.least_squares finds for the same cost written independently in numpy/scipy (rotation-vector
    parametrisation, numeric Jacobians) - pins error definitions, information weighting and gauge handling"""
    from scipy.optimize import least_squares
    from scipy.spatial.transform import Rotation
    from oracle_binding import LM, Oracle
    from openslam_g2o_b200 import synth

    def oracle_chi(p, iters):
        o = Oracle()
        synth.feed(p, o)
        assert o.setup_cli(True) == 0
        o.initialize_optimization()
        n, st = o.optimize(LM, iters)
        return [s.chi2 for s in st[:n]][-1]

    # ---- SE3 pose graph: e = [t ; q_xyz (w >= 0)] of Z^-1 Xi^-1 Xj, Omega = diag blocks (create_sphere.cpp)
    p = synth.sphere(6, 4, seed=3)
    nv = len(p["vertex_ids"])
    t0 = p["vertex_payload"][:, :3]
    R0 = Rotation.from_quat(p["vertex_payload"][:, 3:7]).as_matrix()
    zt = p["edge_payload"][:, :3]
    zq = p["edge_payload"][:, 3:7]
    zR = Rotation.from_quat(zq / np.linalg.norm(zq, axis=1, keepdims=True)).as_matrix()
    iu = p["edge_payload"][:, 7:]
    Om = np.zeros((len(iu), 6, 6))
    k = 0
    for i in range(6):
        for j in range(i, 6):
            Om[:, i, j] = Om[:, j, i] = iu[:, k]
            k += 1
    Lc = np.linalg.cholesky(Om)
    a, b = p["edge_v0"], p["edge_v1"]

    def res_se3(x):
        rv = np.concatenate([np.zeros((1, 3)), x[:3 * (nv - 1)].reshape(-1, 3)])
        dt = np.concatenate([np.zeros((1, 3)), x[3 * (nv - 1):].reshape(-1, 3)])
        R = np.einsum("nij,njk->nik", R0, Rotation.from_rotvec(rv).as_matrix())
        t = t0 + dt
        Rij = np.einsum("eji,ejk->eik", R[a], R[b])                   # Xi^-1 Xj
        tij = np.einsum("eji,ej->ei", R[a], t[b] - t[a])
        Rd = np.einsum("eji,ejk->eik", zR, Rij)                       # Z^-1 (Xi^-1 Xj)
        td = np.einsum("eji,ej->ei", zR, tij - zt)
        q = Rotation.from_matrix(Rd).as_quat()
        q = q * np.where(q[:, 3:4] < 0, -1.0, 1.0)
        e = np.concatenate([td, q[:, :3]], axis=1)
        return np.einsum("eji,ej->ei", Lc, e).reshape(-1)
    sol = least_squares(res_se3, np.zeros(6 * (nv - 1)), method="lm", xtol=1e-14, ftol=1e-14, gtol=1e-14, max_nfev=50000)
    chi = oracle_chi(p, 15)
    assert abs(float(np.sum(sol.fun ** 2)) - chi) <= 1e-6 * chi, (float(np.sum(sol.fun ** 2)), chi)

    # ---- SBACam bundle adjustment: e = K [R^T | -R^T c] X projected - z, Omega = I (types_sba.cpp:204-213)
    p = synth.venice_like(6, 40, seed=3)
    cam = p["cam_payload"]
    ncam = len(cam)
    c0 = cam[:, :3]
    Rw0 = Rotation.from_quat(cam[:, 3:7]).as_matrix()                   # camera to world
    fx, fy, cx, cy = cam[0, 7:11]
    cam_of, pt_of, uv = p["edge_v1"], p["edge_v0"] - ncam, p["edge_payload"]

    def res_ba(x):
        rv = np.concatenate([np.zeros((1, 3)), x[:3 * (ncam - 1)].reshape(-1, 3)])
        dc = np.concatenate([np.zeros((1, 3)), x[3 * (ncam - 1):6 * (ncam - 1)].reshape(-1, 3)])
        X = x[6 * (ncam - 1):].reshape(-1, 3)
        Rw = np.einsum("nij,njk->nik", Rw0, Rotation.from_rotvec(rv).as_matrix())
        pc = np.einsum("eji,ej->ei", Rw[cam_of], X[pt_of] - (c0 + dc)[cam_of])
        return (np.stack([fx * pc[:, 0] / pc[:, 2] + cx, fy * pc[:, 1] / pc[:, 2] + cy], axis=1) - uv).reshape(-1)
    x0 = np.concatenate([np.zeros(6 * (ncam - 1)), p["point_payload"].reshape(-1)])
    sol = least_squares(res_ba, x0, method="lm", xtol=1e-14, ftol=1e-14, gtol=1e-14, max_nfev=50000)
    chi = oracle_chi(p, 15)
    assert abs(float(np.sum(sol.fun ** 2)) - chi) <= 1e-6 * chi, (float(np.sum(sol.fun ** 2)), chi)


@needs_oracle
def test_dq_dR_is_the_reference_object_code_and_equals_the_restatement():
    """The SE3 Jacobians of the oracle call the reference's OWN compute_dq_dR (g2o/types/slam3d/dquat2mat.cpp and its
    maxima-generated tables compiled unmodified into oracle/_ref/libg2o_slam3d_ref.so behind the Eigen shim).  The
    restatement that the device code follows (csrc/geometry.cuh) must agree with it bit for bit on all four branches of
    the rotation -> quaternion case split, and with central differences of the quaternion of R."""
    import ctypes as C
    from oracle_binding import oracle_lib
    L = oracle_lib()
    rng = np.random.default_rng(5)
    seen = set()
    for i in range(3000):
        q = rng.standard_normal(4)
        q /= np.linalg.norm(q)
        w, x, y, z = q
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        Rc = R.ravel(order="F").copy()
        a, b = np.zeros(27), np.zeros(27)
        L.oracle_dq_dR(Rc.ctypes.data_as(C.c_void_p), 0, a.ctypes.data_as(C.c_void_p))
        L.oracle_dq_dR(Rc.ctypes.data_as(C.c_void_p), 1, b.ctypes.data_as(C.c_void_p))
        assert np.array_equal(a, b)
        tr = np.trace(R)
        seen.add(0 if tr > 0 else 1 if (R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]) else 2 if R[1, 1] > R[2, 2] else 3)
    assert seen == {0, 1, 2, 3}


@needs_oracle
@pytest.mark.parametrize("d", [3, 6])
def test_oracle_pcg_matches_dense_solve(d):
    """the restated LinearSolverPCG (solvers/pcg/linear_solver_pcg.hpp:79-197): converges to the dense solution, stops by
    the relative rule dn <= tolerance * dn_0, carries the absolute residual to the next solve, honours maxIter"""
    import ctypes as C
    from helpers import random_spd_blocks
    from oracle_binding import oracle_lib
    L = oracle_lib()
    L.oracle_pcg_solve.restype = C.c_int
    rng = np.random.default_rng(40 + d)
    nb = 60
    edges = [(i, j) for i in range(nb) for j in range(i + 1, min(nb, i + 5))] + [(i, (7 * i + 3) % nb) for i in range(nb)]
    cp, ri, vals, A = random_spd_blocks(rng, nb, d, edges)
    vcm = np.ascontiguousarray(np.transpose(vals, (0, 2, 1)))
    b = rng.standard_normal(nb * d)
    x = np.zeros(nb * d)
    res = C.c_double(-1.0)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    it = L.oracle_pcg_solve(nb, d, p(cp), p(ri), p(vcm), p(x), p(b), C.c_double(1e-12), 1, -1, C.byref(res))
    xr = np.linalg.solve(A, b)
    assert 2 < it < nb * d
    assert np.abs(x - xr).max() <= 1e-6 * np.abs(xr).max()
    assert res.value > 0
    # second solve with the carried-over absolute residual: stops no later than the first
    it2 = L.oracle_pcg_solve(nb, d, p(cp), p(ri), p(vcm), p(x), p(b), C.c_double(1e-12), 1, -1, C.byref(res))
    assert it2 <= it
    res = C.c_double(-1.0)
    assert L.oracle_pcg_solve(nb, d, p(cp), p(ri), p(vcm), p(x), p(b), C.c_double(1e-12), 1, 3, C.byref(res)) == 3


@needs_oracle
def test_robust_kernels_and_se2_are_the_reference_object_code_and_equal_the_restatement():
    """The oracle's robustify calls the reference's OWN kernels (g2o/core/robust_kernel_impl.cpp with robust_kernel.cpp and
    robust_kernel_factory.cpp compiled unmodified into oracle/_ref/libg2o_ref_wrap.so), and its 2D edge errors are evaluated
    by the reference's own SE2 class (g2o/types/slam2d/se2.h instantiated by oracle/ref_wrap.cpp) - both behind the Eigen
    shim of oracle/stub.  The restatements (the formula sheets of the device code) must agree with them: the kernels bit
    for bit on both sides of every threshold, the SE2 algebra to the last ulp or two (the two translation units contract
    products differently), and the device math (host build of csrc/geometry.cuh) with the reference's SE2 errors."""
    import ctypes as C
    import os
    from oracle_binding import oracle_lib
    L = oracle_lib()
    from helpers import ROOT
    R = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libg2o_ref_wrap.so"))
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    for f in (L.oracle_robustify, L.oracle_robustify_restated, R.ref_robustify):
        f.argtypes = [C.c_int, C.c_double, C.c_double, C.c_void_p]
    rng = np.random.default_rng(12)
    for kind in range(1, 6):
        for _ in range(400):
            delta = float(rng.uniform(0.1, 5.0))
            e2 = float(rng.choice([rng.uniform(0, delta * delta), rng.uniform(delta * delta, 40 * delta * delta), delta * delta]))
            a, b, c = np.zeros(3), np.zeros(3), np.zeros(3)
            L.oracle_robustify(kind, delta, e2, p(a)); L.oracle_robustify_restated(kind, delta, e2, p(b)); R.ref_robustify(kind, delta, e2, p(c))
            assert np.array_equal(a, c)            # the oracle IS the reference here
            assert np.array_equal(a[:2], b[:2]), (kind, delta, e2, a, b)   # rho and rho' (what the path uses) bit for bit
            assert np.allclose(a[2], b[2], rtol=1e-14, atol=0)
    # SE2 algebra and the two 2D edge errors
    G = C.CDLL(os.path.join(ROOT, "tests", "csrc", "libgeometry_host.so"))
    worst = 0.0
    for _ in range(2000):
        a = np.array([rng.normal(0, 30), rng.normal(0, 30), rng.uniform(-np.pi, np.pi)])
        b = np.array([rng.normal(0, 30), rng.normal(0, 30), rng.uniform(-np.pi, np.pi)])
        z = np.array([rng.normal(0, 3), rng.normal(0, 3), rng.uniform(-np.pi, np.pi)])
        l = rng.normal(0, 30, 2)
        r_ref, r_res = np.zeros(3), np.zeros(3)
        R.ref_se2_mul(p(a), p(b), p(r_ref)); L.oracle_se2_restated(0, p(a), p(b), p(z), p(r_res))
        worst = max(worst, np.abs(r_ref - r_res).max() / 30)
        R.ref_se2_inverse(p(a), p(r_ref)); L.oracle_se2_restated(1, p(a), p(b), p(z), p(r_res))
        worst = max(worst, np.abs(r_ref - r_res).max() / 30)
        R.ref_edge_se2_error(p(a), p(b), p(z), p(r_ref)); L.oracle_se2_restated(2, p(a), p(b), p(z), p(r_res))
        worst = max(worst, np.abs(r_ref - r_res).max() / 60)
        e_dev, A, B = np.zeros(3), np.zeros(9), np.zeros(9)
        G.gh_se2(p(a), p(b), p(z), p(e_dev), p(A), p(B))                 # device math: EdgeSE2 error
        worst = max(worst, np.abs(r_ref - e_dev).max() / 60)
        e_ref, e_res, e_dev2 = np.zeros(2), np.zeros(2), np.zeros(2)
        R.ref_edge_se2_xy_error(p(a), p(l), p(z), p(e_ref)); L.oracle_se2_restated(3, p(a), p(l), p(z), p(e_res))
        G.gh_se2_xy(p(a), p(l), p(z), p(e_dev2), p(np.zeros(6)), p(np.zeros(4)))   # device math: EdgeSE2PointXY error
        worst = max(worst, np.abs(e_ref - e_res).max() / 60, np.abs(e_ref - e_dev2).max() / 60)
    assert worst <= 4 * np.finfo(float).eps, worst
