"""The device math (openslam_g2o_b200/csrc/geometry.cuh), compiled for the host by tests/csrc (test-only), against
(1) the oracle's quadratic forms and (2) central differences - the reference's own self-consistency checks
(types/slam3d/test_slam3d_jacobian.cpp:100-149, allowedDifference 1e-6)."""
import ctypes as C
import os

import numpy as np
import pytest
from conftest import needs_oracle
from helpers import ROOT, rel_err

LIB = os.path.join(ROOT, "tests", "csrc", "libgeometry_host.so")


@pytest.fixture(scope="module")
def gh():
    if not os.path.exists(LIB):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "csrc")])
    return C.CDLL(LIB)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _rand_iso(rng, scale=1.0):
    q = rng.standard_normal(4)
    q /= np.linalg.norm(q)
    x, y, z, w = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    return np.concatenate([R.T.reshape(-1), scale * rng.standard_normal(3)]), q  # R col-major, t


def _numeric(f, est_list, kind, which, dim, D, gh, delta=1e-6):
    """central differences of the error w.r.t. the oplus increment on vertex `which`"""
    J = np.zeros((D, dim))
    for k in range(dim):
        e = []
        for sgn in (+1, -1):
            ests = [x.copy() for x in est_list]
            u = np.zeros(6)
            u[k] = sgn * delta
            gh.gh_oplus(kind[which], _p(ests[which]), _p(u))
            e.append(f(ests))
        J[:, k] = (e[0] - e[1]) / (2 * delta)
    return J


def test_se3_analytic_jacobian_vs_numeric(gh):
    rng = np.random.default_rng(0)
    worst = 0.0
    for _ in range(2000):
        Xi, _ = _rand_iso(rng)
        Xj, _ = _rand_iso(rng)
        Z, _ = _rand_iso(rng)
        e, Ji, Jj = np.zeros(6), np.zeros(36), np.zeros(36)
        gh.gh_se3(_p(Xi), _p(Xj), _p(Z), _p(e), _p(Ji), _p(Jj))
        assert np.all(np.abs(e) < 1e100)  # error of se3_error and se3_jacobians agree to rounding

        def f(ests):
            ee, a, b = np.zeros(6), np.zeros(36), np.zeros(36)
            gh.gh_se3(_p(ests[0]), _p(ests[1]), _p(Z), _p(ee), _p(a), _p(b))
            return ee
        Ni = _numeric(f, [Xi, Xj], [1, 1], 0, 6, 6, gh)
        Nj = _numeric(f, [Xi, Xj], [1, 1], 1, 6, 6, gh)
        worst = max(worst, np.abs(Ji.reshape(6, 6).T - Ni).max(), np.abs(Jj.reshape(6, 6).T - Nj).max())
    assert worst < 1e-6, worst


def test_se2_and_p2mc_analytic_jacobians_vs_numeric(gh):
    rng = np.random.default_rng(1)
    for _ in range(500):
        xi, xj, z = rng.uniform(-3, 3, 3), rng.uniform(-3, 3, 3), rng.uniform(-3, 3, 3)
        e, A, B = np.zeros(3), np.zeros(9), np.zeros(9)
        gh.gh_se2(_p(xi), _p(xj), _p(z), _p(e), _p(A), _p(B))

        def f(ests):
            ee, a, b = np.zeros(3), np.zeros(9), np.zeros(9)
            gh.gh_se2(_p(ests[0]), _p(ests[1]), _p(z), _p(ee), _p(a), _p(b))
            return ee
        if abs(abs(e[2]) - np.pi) < 1e-3:
            continue  # angle wrap inside the difference stencil
        assert np.abs(A.reshape(3, 3).T - _numeric(f, [xi, xj], [0, 0], 0, 3, 3, gh)).max() < 1e-6
        assert np.abs(B.reshape(3, 3).T - _numeric(f, [xi, xj], [0, 0], 1, 3, 3, gh)).max() < 1e-6
    for _ in range(500):
        iso, q = _rand_iso(rng)
        if q[3] < 0:
            q = -q
        cam = np.concatenate([iso[9:], q, [800.0, 820.0, 10.0, -5.0, 0.0]])
        R = iso[:9].reshape(3, 3).T
        X = cam[:3] + R @ np.array([rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(2, 6)])
        zz = rng.uniform(-100, 100, 2)
        e, Jp, Jc = np.zeros(2), np.zeros(6), np.zeros(12)
        gh.gh_p2mc(_p(cam), _p(X), _p(zz), _p(e), _p(Jp), _p(Jc))

        def f(ests):
            ee, a, b = np.zeros(2), np.zeros(6), np.zeros(12)
            gh.gh_p2mc(_p(ests[1]), _p(ests[0]), _p(zz), _p(ee), _p(a), _p(b))
            return ee
        Np = _numeric(f, [X, cam], [3, 2], 0, 3, 2, gh)
        Nc = _numeric(f, [X, cam], [3, 2], 1, 6, 2, gh)
        scale = max(1.0, np.abs(Nc).max())
        assert np.abs(Jp.reshape(3, 2).T - Np).max() < 1e-5 * scale
        assert np.abs(Jc.reshape(6, 2).T - Nc).max() < 1e-5 * scale


@needs_oracle
def test_device_math_matches_oracle_quadratic_forms(gh):
    """one edge, Omega = I, nothing fixed: the oracle's Hpp/Hpl/Hll blocks and b are A^T A, A^T B, B^T B, -J^T e"""
    from oracle_binding import Oracle
    rng = np.random.default_rng(2)
    for trial in range(200):
        kind = trial % 3
        o = Oracle()
        if kind == 0:
            xi, xj, z = rng.uniform(-3, 3, 3), rng.uniform(-3, 3, 3), rng.uniform(-3, 3, 3)
            o.add_vertices(0, [0, 1], np.stack([xi, xj]))
            o.add_edges(0, [0], [1], np.concatenate([z, [1, 0, 0, 1, 0, 1]])[None])
            e, A, B = np.zeros(3), np.zeros(9), np.zeros(9)
            gh.gh_se2(_p(xi), _p(xj), _p(z), _p(e), _p(A), _p(B))
            A, B = A.reshape(3, 3).T, B.reshape(3, 3).T
        elif kind == 1:
            pays = []
            for _ in range(3):
                iso, q = _rand_iso(rng)
                pays.append(np.concatenate([iso[9:], q]))
            o.add_vertices(1, [0, 1], np.stack(pays[:2]))
            iu = np.array([1.0 if i == j else 0.0 for i in range(6) for j in range(i, 6)])
            o.add_edges(1, [0], [1], np.concatenate([pays[2], iu])[None])
            ests = [o.vertex_estimate(0), o.vertex_estimate(1)]
            Z = o.edges()[0][3]
            e, A, B = np.zeros(6), np.zeros(36), np.zeros(36)
            gh.gh_se3(_p(ests[0]), _p(ests[1]), _p(np.ascontiguousarray(Z)), _p(e), _p(A), _p(B))
            A, B = A.reshape(6, 6).T, B.reshape(6, 6).T
        else:
            iso, q = _rand_iso(rng)
            camp = np.concatenate([iso[9:], q, [800.0, 820.0, 10.0, -5.0, 0.0]])
            R = iso[:9].reshape(3, 3).T
            X = camp[:3] + R @ np.array([rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(2, 6)])
            zz = rng.uniform(-100, 100, 2)
            o.add_vertices(2, [0], camp[None])
            o.add_vertices(3, [1], X[None])
            o.add_edges(2, [1], [0], zz[None])
            cam = o.vertex_estimate(0)
            e, A, B = np.zeros(2), np.zeros(6), np.zeros(12)
            gh.gh_p2mc(_p(cam), _p(X), _p(zz), _p(e), _p(A), _p(B))
            A, B = A.reshape(3, 2).T, B.reshape(6, 2).T  # A: point (vertex 0 of the edge), B: camera
        o.initialize_optimization()
        o.algorithm_init()
        o.build_structure()
        o.compute_active_errors()
        o.build_system()
        b = o.b()
        rows, cols, H = o.blocks(0)
        if kind < 2:
            d = A.shape[1]
            Hd = {(int(r), int(c)): v for r, c, v in zip(rows, cols, H)}
            assert rel_err(Hd[(0, 0)], A.T @ A) < 1e-11
            assert rel_err(Hd[(1, 1)], B.T @ B) < 1e-11
            assert rel_err(Hd[(0, 1)], A.T @ B) < 1e-11
            assert rel_err(b, np.concatenate([-A.T @ e, -B.T @ e])) < 1e-10
        else:
            # no marginalisation requested here: camera (id 0) index 0, point (id 1) index 1, both in Hpp? no -
            # dims differ, the oracle keeps them both as "poses"; only check b and the diagonal blocks
            assert rel_err(b[:6], -B.T @ e) < 1e-10
            assert rel_err(b[6:9], -A.T @ e) < 1e-10


def test_inverse3_and_oplus(gh):
    rng = np.random.default_rng(3)
    for _ in range(200):
        m = rng.standard_normal((3, 3))
        m = m @ m.T + 0.1 * np.eye(3)
        r = np.zeros(9)
        gh.gh_inverse3(_p(np.ascontiguousarray(m.T.reshape(-1))), _p(r))
        assert rel_err(r.reshape(3, 3).T, np.linalg.inv(m)) < 1e-10
    # SE3 oplus: identity rotation when |v|^2 > 1 (isometry3d_mappings.cpp:84-91)
    est = np.concatenate([np.eye(3).reshape(-1), [1.0, 2.0, 3.0]])
    u = np.array([0.5, 0.0, 0.0, 0.9, 0.9, 0.9])
    gh.gh_oplus(1, _p(est), _p(u))
    assert np.allclose(est[:9], np.eye(3).reshape(-1)) and np.allclose(est[9:], [1.5, 2.0, 3.0])
    # SE2 oplus wraps the angle into [-pi, pi)
    e2 = np.array([0.0, 0.0, 3.0])
    gh.gh_oplus(0, _p(e2), _p(np.array([0.0, 0.0, 1.0, 0, 0, 0])))
    assert -np.pi <= e2[2] < np.pi and abs(e2[2] - (4.0 - 2 * np.pi)) < 1e-12


def _expmap_est(rng):
    """[t3 q4 f f cx cy b] of a world->camera transform that looks at a point cloud around the origin"""
    iso, q = _rand_iso(rng)
    t = np.array([rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(4, 8)])  # the scene sits in front of the camera
    if q[3] < 0:
        q = -q
    return np.concatenate([t, q, [700.0, 700.0, 320.0, 240.0, 0.0]])


def test_expmap_analytic_jacobians_vs_numeric(gh):
    """EdgeProjectXYZ2UV::linearizeOplus against central differences through VertexSE3Expmap::oplusImpl
    (exp(update) * estimate) - the reference's self-consistency criterion applied to the expmap family"""
    rng = np.random.default_rng(5)
    for _ in range(50):
        est = _expmap_est(rng)
        X = rng.uniform(-1, 1, 3)
        z = rng.uniform(0, 600, 2)

        def err(ests):
            e, a, b = np.zeros(2), np.zeros(6), np.zeros(12)
            gh.gh_xyz2uv(_p(ests[1]), _p(ests[0]), _p(z), _p(e), _p(a), _p(b))
            return e.copy()
        e, Jp, Jc = np.zeros(2), np.zeros(6), np.zeros(12)
        gh.gh_xyz2uv(_p(est), _p(X), _p(z), _p(e), _p(Jp), _p(Jc))
        Np = _numeric(err, [X, est], [3, 4], 0, 3, 2, gh)
        Nc = _numeric(err, [X, est], [3, 4], 1, 6, 2, gh)
        scale = max(1.0, np.abs(Nc).max())
        assert np.abs(Jp.reshape(3, 2).T - Np).max() < 1e-5 * scale
        assert np.abs(Jc.reshape(6, 2).T - Nc).max() < 1e-5 * scale


def test_expmap_oplus_is_the_se3_exponential(gh):
    """SE3Quat::exp (se3quat.h:216-252) is the matrix exponential of the twist [omega ; upsilon]: compare the device
    oplus with scipy.linalg.expm on the homogeneous matrices, incl. the small-angle branch (theta < 1e-5)"""
    from scipy.linalg import expm
    rng = np.random.default_rng(6)

    def hom(est):
        x, y, z, w = est[3:7]
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        T = np.eye(4)
        T[:3, :3], T[:3, 3] = R, est[:3]
        return T
    for trial in range(60):
        est = _expmap_est(rng)
        u = rng.standard_normal(6) * (1e-7 if trial % 3 == 0 else 0.5 if trial % 3 == 1 else 3.0)
        om, up = u[:3], u[3:]
        A = np.zeros((4, 4))
        A[:3, :3] = [[0, -om[2], om[1]], [om[2], 0, -om[0]], [-om[1], om[0], 0]]
        A[:3, 3] = up
        want = expm(A) @ hom(est)
        got = est.copy()
        gh.gh_oplus(4, _p(got), _p(u))
        assert abs(np.linalg.norm(got[3:7]) - 1) < 1e-14 and got[6] >= 0
        assert np.abs(hom(got) - want).max() < 1e-11 * max(1.0, np.abs(want).max())
        assert np.array_equal(got[7:], est[7:])


@needs_oracle
def test_expmap_device_math_matches_oracle(gh):
    """one XYZ2UV edge with a full information matrix: the oracle's b equals -J^T Omega e of the device math, and the
    oracle's oplus (quaternion form, as the reference codes it) equals the device oplus"""
    from oracle_binding import Oracle
    rng = np.random.default_rng(7)
    for trial in range(100):
        c2w_iso, q = _rand_iso(rng)
        if q[3] < 0:
            q = -q
        c2w = np.concatenate([c2w_iso[9:], q])
        R = c2w_iso[:9].reshape(3, 3).T
        X = c2w[:3] + R @ np.array([rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(2, 6)])
        zz = rng.uniform(0, 600, 2)
        w = np.array([rng.uniform(0.5, 2), rng.uniform(-0.3, 0.3), rng.uniform(0.5, 2)])
        o = Oracle()
        o.add_camera_parameters(3, 650.0, 300.0, 250.0, 0.1)
        o.add_vertices(4, [0], c2w[None])
        o.add_vertices(3, [1], X[None])
        o.add_edges(3, [1], [0], np.concatenate([[3.0], zz, w])[None])
        est = np.concatenate([o.vertex_estimate(0), [650.0, 650.0, 300.0, 250.0, 0.1]])
        # the estimate is the inverse of the file's cam2world: maps the camera centre to the origin
        assert np.abs(est[:3] + np.array(_rotate(est[3:7], c2w[:3]))).max() < 1e-12
        e, A, B = np.zeros(2), np.zeros(6), np.zeros(12)
        gh.gh_xyz2uv(_p(est), _p(X), _p(zz), _p(e), _p(A), _p(B))
        A, B = A.reshape(3, 2).T, B.reshape(6, 2).T
        o.initialize_optimization()
        o.algorithm_init()
        o.build_structure()
        W = np.array([[w[0], w[1]], [w[1], w[2]]])
        assert abs(o.compute_active_errors() - e @ W @ e) <= 1e-11 * max(1.0, e @ W @ e)
        o.build_system()
        b = o.b()
        assert rel_err(b[:6], -B.T @ W @ e) < 1e-10
        assert rel_err(b[6:9], -A.T @ W @ e) < 1e-10
        u = rng.standard_normal(9) * 0.05
        o.set_x(u)
        o.update()
        got = est.copy()
        gh.gh_oplus(4, _p(got), _p(u[:6]))
        assert rel_err(got[:7], o.vertex_estimate(0)) < 1e-13
        assert rel_err(X + u[6:], o.vertex_estimate(1)) < 1e-15


def _rotate(q, v):
    x, y, z, w = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    return R @ np.asarray(v)
