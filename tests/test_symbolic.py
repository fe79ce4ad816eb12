"""Supernodal symbolic plan (supernodes, row structures, update lists, relative indices, scatter plan, task/level
schedule) validated by executing it sequentially on the CPU (tests/csrc/host_exec.cpp, test-only) and comparing the
solution with numpy's dense solve."""
import ctypes as C
import os

import numpy as np
import pytest
from helpers import ROOT, random_spd_blocks

LIB = os.path.join(ROOT, "tests", "csrc", "libhost_exec.so")


@pytest.fixture(scope="module")
def hx():
    if not os.path.exists(LIB):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "csrc")])
    return C.CDLL(LIB)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("d", [3, 6])
def test_schedule_execution_solves_the_system(hx, d):
    rng = np.random.default_rng(d)
    for trial in range(24):
        nb = int(rng.integers(1, 90))
        kind = trial % 4
        if kind == 0:
            edges = [(int(rng.integers(nb)), int(rng.integers(nb))) for _ in range(3 * nb)]
        elif kind == 1:
            s = max(2, int(np.sqrt(nb)))
            edges = [(i, i + 1) for i in range(nb - 1) if (i + 1) % s] + [(i, i + s) for i in range(nb - s)]
        elif kind == 2:
            edges = [(i, j) for i in range(nb) for j in range(i + 1, min(nb, i + 6))]
        else:
            edges = [(i, j) for i in range(nb) for j in range(i + 1, nb)] if nb < 30 else []  # dense / diagonal
        cp, ri, vals, A = random_spd_blocks(rng, nb, d, edges)
        v = np.ascontiguousarray(np.transpose(vals, (0, 2, 1)))
        b = rng.standard_normal(nb * d)
        x = np.zeros(nb * d)
        for maxc, relax in [(96, 1), (12, 1), (96, 0), (6, 0)]:
            rc = hx.hx_solve(nb, d, _p(cp), _p(ri), _p(v), C.c_double(0.25), _p(b), _p(x), maxc, relax)
            assert rc == 0, (rc, trial, nb, maxc, relax)
            xr = np.linalg.solve(A + 0.25 * np.eye(nb * d), b)
            assert np.abs(x - xr).max() <= 1e-9 * np.abs(xr).max()
            # the dataflow task list of the persistent kernels: dependencies only point backwards, counters add up
            for gi in (1, 4):
                assert hx.hx_check_flow(nb, d, _p(cp), _p(ri), maxc, relax, gi) == 0, (trial, nb, maxc, relax, gi)


def test_dataflow_schedule_on_benchmark_shaped_patterns(hx):
    """ring band (the reduced camera system of the Venice-shaped BA) and a 50 x 50 torus-like grid (sphere2500)"""
    def pattern(nb, edges):
        cols = [set([c]) for c in range(nb)]
        for i, j in edges:
            cols[max(i, j)].add(min(i, j))
        cp = np.zeros(nb + 1, dtype=np.int32)
        ri = []
        for c in range(nb):
            r = sorted(cols[c])
            ri += r
            cp[c + 1] = len(ri)
        return cp, np.asarray(ri, dtype=np.int32)
    ring = [(i, (i + k) % 871) for i in range(871) for k in range(1, 9)]
    grid = [(r * 50 + c, r * 50 + (c + 1) % 50) for r in range(50) for c in range(50)] + \
           [(r * 50 + c, (r + 1) * 50 + c) for r in range(49) for c in range(50)]
    for nb, edges, d in ((871, ring, 6), (2500, grid, 6), (2500, grid, 3)):
        cp, ri = pattern(nb, edges)
        for maxc, gi in ((72, 4), (72, 1), (24, 3)):
            assert hx.hx_check_flow(nb, d, _p(cp), _p(ri), maxc, 1, gi) == 0, (nb, d, maxc, gi)


def test_not_positive_definite_is_reported(hx):
    rng = np.random.default_rng(5)
    cp, ri, vals, A = random_spd_blocks(rng, 8, 3, [(i, i + 1) for i in range(7)], shift=0.0)
    vals[0] -= 50 * np.eye(3)
    v = np.ascontiguousarray(np.transpose(vals, (0, 2, 1)))
    b = np.ones(24)
    x = np.zeros(24)
    assert hx.hx_solve(8, 3, _p(cp), _p(ri), _p(v), C.c_double(0.0), _p(b), _p(x), 96, 1) == 1


def test_nested_dissection_option(hx):
    """optional ordering for parallelism: still a permutation, the schedule still solves the system, the dataflow list
    is still valid - and on a ring band (the reduced camera system of a BA) the elimination tree gets much shallower"""
    rng = np.random.default_rng(9)
    try:
        for nd in (1, 3):
            hx.hx_set_nd_levels(nd)
            for d, nb, edges in ((3, 60, [(i, j) for i in range(60) for j in range(i + 1, min(60, i + 4))]),
                                 (6, 45, [(i, (i + k) % 45) for i in range(45) for k in range(1, 4)]),
                                 (6, 30, [(int(rng.integers(30)), int(rng.integers(30))) for _ in range(80)])):
                cp, ri, vals, A = random_spd_blocks(rng, nb, d, edges)
                v = np.ascontiguousarray(np.transpose(vals, (0, 2, 1)))
                b = rng.standard_normal(nb * d)
                x = np.zeros(nb * d)
                assert hx.hx_solve(nb, d, _p(cp), _p(ri), _p(v), C.c_double(0.5), _p(b), _p(x), 24, 1) == 0
                xr = np.linalg.solve(A + 0.5 * np.eye(nb * d), b)
                assert np.abs(x - xr).max() <= 1e-9 * np.abs(xr).max()
                assert hx.hx_check_flow(nb, d, _p(cp), _p(ri), 24, 1, 4) == 0
                perm = np.zeros(nb, np.int32)
                info = np.zeros(16, np.int64)
                hx.hx_analyze(nb, d, _p(cp), _p(ri), 24, 1, _p(info), _p(perm))
                assert sorted(perm.tolist()) == list(range(nb))
        # ring of 871 cameras, every camera coupled to the next 8: levels of the task schedule, AMD vs dissection
        from helpers import upper_pattern_from_edges
        ring = [(i, (i + k) % 871) for i in range(871) for k in range(1, 9)]
        cp, ri = upper_pattern_from_edges(871, ring)
        res = {}
        for nd in (0, 4):
            hx.hx_set_nd_levels(nd)
            info = np.zeros(16, np.int64)
            hx.hx_analyze(871, 6, _p(cp), _p(ri), 72, 1, _p(info), None)
            res[nd] = (int(info[2]), int(info[4]))  # levels, factor doubles
            assert hx.hx_check_flow(871, 6, _p(cp), _p(ri), 72, 1, 8) == 0
        assert res[4][0] * 2 < res[0][0], res          # much shallower
        assert res[4][1] < 3 * res[0][1], res          # at a bounded price in fill
    finally:
        hx.hx_set_nd_levels(0)


def test_tail_chain_plan(hx):
    """the tail chain (one CTA keeps the frontal matrix in registers along the last path of the elimination tree):
    the host executor runs the chain the way chol_chain_kernel does (re-index maps, new-row masks, filtered work items)
    and must still solve the system; on band-like patterns the chain must actually exist"""
    from helpers import upper_pattern_from_edges
    rng = np.random.default_rng(3)
    try:
        for min_links in (1, 3):
            hx.hx_set_chain(1, min_links, 31)
            cases = [(45, [(i, (i + k) % 45) for i in range(45) for k in range(1, 4)], 12),
                     (120, [(i, (i + k) % 120) for i in range(120) for k in range(1, 6)], 24),
                     (200, [(i, (i + k) % 200) for i in range(200) for k in range(1, 9)], 72),
                     (60, [(int(rng.integers(60)), int(rng.integers(60))) for _ in range(150)], 24),
                     (90, [(i, i + 1) for i in range(89)] + [(i, i + 9) for i in range(81)], 36)]
            for nb, edges, maxc in cases:
                cp, ri, vals, A = random_spd_blocks(rng, nb, 6, edges)
                v = np.ascontiguousarray(np.transpose(vals, (0, 2, 1)))
                b = rng.standard_normal(nb * 6)
                x = np.zeros(nb * 6)
                rc = hx.hx_solve(nb, 6, _p(cp), _p(ri), _p(v), C.c_double(0.5), _p(b), _p(x), maxc, 1)
                assert rc == 0, (rc, nb, maxc, min_links)
                xr = np.linalg.solve(A + 0.5 * np.eye(nb * 6), b)
                assert np.abs(x - xr).max() <= 1e-9 * np.abs(xr).max()
                for gi in (1, 4, 16):
                    assert hx.hx_check_flow(nb, 6, _p(cp), _p(ri), maxc, 1, gi) == 0, (nb, maxc, gi)
        hx.hx_set_chain(1, 3, 31)
        # the reduced camera systems of the two BA workloads: almost everything is chain
        for nb, w in ((871, 8), (2000, 9)):
            ring = [(i, (i + k) % nb) for i in range(nb) for k in range(1, w + 1)]
            cp, ri = upper_pattern_from_edges(nb, ring)
            info = np.zeros(16, np.int64)
            hx.hx_analyze(nb, 6, _p(cp), _p(ri), 72, 1, _p(info), None)
            assert info[8] >= 0.5 * info[0] - 8, info   # chain links vs supernodes
            assert hx.hx_check_flow(nb, 6, _p(cp), _p(ri), 72, 1, 16) == 0
        # block dimension 3 (SE2) and fronts wider than the chain's register file: no chain, the old path
        grid = [(r * 50 + c, r * 50 + (c + 1) % 50) for r in range(50) for c in range(50)] + \
               [(r * 50 + c, (r + 1) * 50 + c) for r in range(49) for c in range(50)]
        cp, ri = upper_pattern_from_edges(2500, grid)
        for d in (3, 6):
            info = np.zeros(16, np.int64)
            hx.hx_analyze(2500, d, _p(cp), _p(ri), 72, 1, _p(info), None)
            # d = 3: never; d = 6: at most the last few panels of the top separator (their fronts have shrunk to <= 31 rows)
            assert info[8] == 0 if d == 3 else info[8] <= 4, info
    finally:
        hx.hx_set_chain(1, 3, 31)


def test_wide_tile_plan(hx):
    """96 x 72 destination tiles (SymbolicOptions::wide_tiles, the FP64 tensor-path plan of large pose graphs): the
    sequential executor must solve the same systems through the rectangular-tile work lists, incl. dense fronts several
    tiles tall, panels cut at 12 block columns, and the tail chain receiving its updates through wide tiles."""
    rng = np.random.default_rng(77)
    hx.hx_set_wide(1)
    try:
        for trial in range(10):
            nb = int(rng.integers(20, 120))
            kind = trial % 4
            if kind == 0:
                edges = [(int(rng.integers(nb)), int(rng.integers(nb))) for _ in range(4 * nb)]
            elif kind == 1:
                s = max(2, int(np.sqrt(nb)))
                edges = [(i, i + 1) for i in range(nb - 1) if (i + 1) % s] + [(i, i + s) for i in range(nb - s)]
            elif kind == 2:
                edges = [(i, j) for i in range(nb) for j in range(i + 1, min(nb, i + 20))]
            else:
                nb = min(nb, 60)
                edges = [(i, j) for i in range(nb) for j in range(i + 1, nb)]
            cp, ri, vals, A = random_spd_blocks(rng, nb, 6, edges)
            v = np.ascontiguousarray(np.transpose(vals, (0, 2, 1)))
            b = rng.standard_normal(nb * 6)
            x = np.zeros(nb * 6)
            for maxc, relax in [(72, 1), (72, 0), (24, 1)]:
                rc = hx.hx_solve(nb, 6, _p(cp), _p(ri), _p(v), C.c_double(0.25), _p(b), _p(x), maxc, relax)
                assert rc == 0, (rc, trial, nb, maxc, relax)
                xr = np.linalg.solve(A + 0.25 * np.eye(nb * 6), b)
                assert np.abs(x - xr).max() <= 1e-9 * np.abs(xr).max()
                for gi in (1, 4):
                    assert hx.hx_check_flow(nb, 6, _p(cp), _p(ri), maxc, relax, gi) == 0, (trial, nb, maxc, relax, gi)
    finally:
        hx.hx_set_wide(-1)


@pytest.mark.parametrize("d", [3, 6])
def test_sparse_inverse_recursion_matches_dense_inverse(hx, d):
    """Takahashi recursion in supernodal form (csrc/sparse_inverse.cuh runs the same steps with the same block lookup,
    csrc/spinv_lookup.h): every block on the pattern of the factor - all diagonal blocks and all blocks of the input
    pattern among them - equals the block of the dense inverse; blocks outside the pattern are reported as such
    (core/marginal_covariance_cholesky.cpp:71-100 is the scalar form of this recursion)"""
    rng = np.random.default_rng(40 + d)
    for trial in range(16):
        nb = int(rng.integers(2, 70))
        kind = trial % 4
        if kind == 0:
            edges = [(int(rng.integers(nb)), int(rng.integers(nb))) for _ in range(2 * nb)]
        elif kind == 1:
            s = max(2, int(np.sqrt(nb)))
            edges = [(i, i + 1) for i in range(nb - 1) if (i + 1) % s] + [(i, i + s) for i in range(nb - s)]
        elif kind == 2:
            edges = [(i, j) for i in range(nb) for j in range(i + 1, min(nb, i + 5))]
        else:
            edges = [(i, i + 1) for i in range(nb - 1)]
        cp, ri, vals, A = random_spd_blocks(rng, nb, d, edges)
        v = np.ascontiguousarray(np.transpose(vals, (0, 2, 1)))
        inv = np.linalg.inv(A + 0.25 * np.eye(nb * d))
        pairs = [(i, i) for i in range(nb)] + [(int(ri[k]), c) for c in range(nb) for k in range(cp[c], cp[c + 1])]
        pairs += [(c, int(ri[k])) for c in range(nb) for k in range(cp[c], cp[c + 1])][:20]   # transposed requests
        pairs += [(int(rng.integers(nb)), int(rng.integers(nb))) for _ in range(20)]         # may be outside
        rows = np.asarray([p[0] for p in pairs], np.int32)
        cols = np.asarray([p[1] for p in pairs], np.int32)
        for maxc, relax in [(96, 1), (12, 1), (6, 0)]:
            out = np.zeros((len(pairs), d, d))
            found = np.zeros(len(pairs), np.int32)
            rc = hx.hx_sparse_inverse(nb, d, _p(cp), _p(ri), _p(v), C.c_double(0.25), len(pairs), _p(rows), _p(cols),
                                      _p(out), _p(found), maxc, relax)
            assert rc == 0, (rc, trial, nb, maxc, relax)
            n_in = nb + int(cp[nb])
            assert found[:n_in].all()   # diagonal blocks and the input pattern are always on the pattern of L
            for q, (r, c) in enumerate(pairs):
                if not found[q]:
                    continue
                ref = inv[r * d:(r + 1) * d, c * d:(c + 1) * d]
                assert np.abs(out[q].T - ref).max() <= 1e-9 * np.abs(inv).max(), (trial, q, r, c)


def test_sparse_inverse_on_an_ill_conditioned_graph_with_a_tail_chain(hx):
    """parking-garage Hessian (1660 poses, condition number 1.3e12, the factorisation ends in a 15-link tail chain): the
    supernodal recursion in plain double arithmetic agrees with numpy's dense inverse to 5e-8 of the largest entry - the
    accuracy the conditioning leaves (numpy's own inverse has a residual of 7e-8), NOT the 1e-8 the well-conditioned
    fixtures reach.  (Measured on B200 with the same matrix: the GPU sweep fails a 1e-8 bound at exactly the blocks where
    this host run exceeds it, i.e. the chain's inverse diagonal blocks reach the sweep correctly.)"""
    from conftest import have_oracle
    if not have_oracle():
        pytest.skip("oracle not built")
    from helpers import feed_fixture, load_fixture
    from oracle_binding import Oracle
    fx = load_fixture("garage")
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
    cp = np.zeros(nb + 1, np.int32)
    for c in cols:
        cp[c + 1] += 1
    cp = np.cumsum(cp).astype(np.int32)
    ri = np.asarray(rows, np.int32)
    v = np.ascontiguousarray(np.transpose(vals, (0, 2, 1)))
    rr = np.arange(nb, dtype=np.int32)
    out = np.zeros((nb, d, d))
    found = np.zeros(nb, np.int32)
    try:
        hx.hx_set_chain(1, 3, 29)   # the GPU's chain parameters
        info = np.zeros(16, np.int64)
        hx.hx_analyze(nb, d, _p(cp), _p(ri), 72, 1, _p(info), None)
        assert info[8] >= 3         # the plan does end in a tail chain
        assert hx.hx_sparse_inverse(nb, d, _p(cp), _p(ri), _p(v), C.c_double(0.0), nb, _p(rr), _p(rr), _p(out), _p(found), 72, 1) == 0
    finally:
        hx.hx_set_chain(1, 3, 31)
    assert found.all()
    err = max(np.abs(out[i].T - inv[i * d:(i + 1) * d, i * d:(i + 1) * d]).max() for i in range(nb))
    assert err <= 1e-6 * np.abs(inv).max()


def test_camera_pair_index_of_the_wide_landmark_schur_kernel():
    """schur_wide_kernel (csrc/kernels.cuh) turns a pair number q into the camera pair (a, b), a <= b, of a landmark seen by
    k cameras - the enumeration the host plan uses for its segments (a ascending, b = a .. k-1).  The same arithmetic
    (double sqrt guess + integer correction) here: every q for small k, the edges of every row for k up to the 65535-camera
    limit"""
    import math

    def pair_of(q, k):
        a = int(((2.0 * k + 1.0) - math.sqrt((2.0 * k + 1.0) * (2.0 * k + 1.0) - 8.0 * float(q))) * 0.5)
        a = max(0, min(a, k - 1))
        while a > 0 and a * k - a * (a - 1) // 2 > q:
            a -= 1
        while a + 1 < k and (a + 1) * k - (a + 1) * a // 2 <= q:
            a += 1
        return a, a + (q - (a * k - a * (a - 1) // 2))
    for k in (1, 2, 3, 7, 64, 257):
        q = 0
        for a in range(k):
            for b in range(a, k):
                assert pair_of(q, k) == (a, b), (k, q)
                q += 1
    for k in (1401, 4096, 20000, 65535):
        for a in list(range(0, k, max(1, k // 997))) + [k - 2, k - 1]:
            if a < 0:
                continue
            first = a * k - a * (a - 1) // 2
            assert pair_of(first, k) == (a, a)
            assert pair_of(first + (k - 1 - a), k) == (a, k - 1)      # last pair of the row
            if a > 0:
                assert pair_of(first - 1, k) == (a - 1, k - 1)
