"""shared helpers of the test-suite"""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
FIXTURES = ("manhattan3500", "intel", "sphere_bignoise", "garage")
ALGO_NAME = {("manhattan3500"): "gn_fix3_2", "intel": "gn_fix3_2", "sphere_bignoise": "lm_fix6_3", "garage": "lm_fix6_3"}


def load_fixture(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def feed_fixture(target, fx):
    """push a golden fixture's input graph into a product SparseOptimizer or an oracle wrapper"""
    vk = int(fx["v_kind"][0])
    ek = int(fx["e_kind"][0])
    target.add_vertices(vk, fx["v_ids"], fx["v_pay"])
    target.add_edges(ek, fx["e_a"], fx["e_b"], fx["e_pay"])
    for vid in fx["fixed"]:
        target.set_fixed(int(vid), True)


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    denom = max(np.abs(b).max(), 1e-300)
    return float(np.abs(a - b).max() / denom)


def blocks_to_dict(rows, cols, vals):
    return {(int(r), int(c)): v for r, c, v in zip(rows, cols, vals)}


def upper_pattern_from_edges(n, edges):
    cols = [set([j]) for j in range(n)]
    for i, j in edges:
        if i != j:
            cols[max(i, j)].add(min(i, j))
    cp, ri = [0], []
    for j in range(n):
        r = sorted(cols[j])
        ri += r
        cp.append(len(ri))
    return np.array(cp, np.int32), np.array(ri, np.int32)


def random_spd_blocks(rng, nb, d, edges, shift=None):
    """random SPD block matrix on the given pattern -> colptr,rowidx,values[nblk,d,d],dense A"""
    cp, ri = upper_pattern_from_edges(nb, edges)
    n = nb * d
    A = np.zeros((n, n))
    vals = np.zeros((len(ri), d, d))
    for j in range(nb):
        for q in range(cp[j], cp[j + 1]):
            i = ri[q]
            if i == j:
                B = rng.standard_normal((d, d))
                B = 0.1 * B @ B.T
            else:
                B = rng.standard_normal((d, d))
                A[j * d:(j + 1) * d, i * d:(i + 1) * d] = B.T
            A[i * d:(i + 1) * d, j * d:(j + 1) * d] = B
            vals[q] = B
    s = np.abs(A).sum(1).max() if shift is None else shift
    for j in range(nb):
        q = [q for q in range(cp[j], cp[j + 1]) if ri[q] == j][0]
        vals[q] += s * np.eye(d)
        A[j * d:(j + 1) * d, j * d:(j + 1) * d] += s * np.eye(d)
    return cp, ri, vals, A
