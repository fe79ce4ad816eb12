"""Fill-reducing ordering: the product's block AMD (b200_block_amd, host) must be bit-exact with the vendored
cs_amd(1, .) (EXTERNAL/csparse/cs_amd.c) that the reference's LinearSolverCSparse calls."""
import numpy as np
import pytest
from conftest import needs_oracle
from helpers import FIXTURES, feed_fixture, load_fixture, upper_pattern_from_edges


def _patterns():
    rng = np.random.default_rng(0)
    for trial in range(120):
        n = int(rng.integers(2, 300))
        kind = trial % 6
        if kind == 0:
            edges = [(int(rng.integers(n)), int(rng.integers(n))) for _ in range(int(rng.integers(0, 4 * n)))]
        elif kind == 1:
            w = int(rng.integers(1, 6))
            edges = [(i, i + k) for i in range(n) for k in range(1, w + 1) if i + k < n]
        elif kind == 2:  # arrow head (one dense row)
            edges = [(0, i) for i in range(n)] + [(i, i + 1) for i in range(n - 1)]
        elif kind == 3:  # 2-D grid
            s = max(2, int(np.sqrt(n)))
            edges = [(i, i + 1) for i in range(n - 1) if (i + 1) % s] + [(i, i + s) for i in range(n - s)]
        elif kind == 4:  # a few dense rows among random entries
            edges = [(int(rng.integers(n)), int(rng.integers(n))) for _ in range(2 * n)]
            edges += [(int(rng.integers(min(3, n))), i) for i in range(n)]
        else:  # empty / diagonal only, ragged tiny cases
            edges = []
        yield n, edges
    # big enough to trigger workspace compaction and the dense-node path (dense = max(16, 10 sqrt n))
    n = 20000
    edges = [(int(rng.integers(n)), int(rng.integers(n))) for _ in range(5 * n)] + [(7, i) for i in range(0, n, 3)]
    yield n, edges


@needs_oracle
def test_block_amd_is_bit_exact_with_cs_amd_on_synthetic_patterns():
    from oracle_binding import cs_amd
    from openslam_g2o_b200 import block_amd
    count = 0
    for n, edges in _patterns():
        cp, ri = upper_pattern_from_edges(n, edges)
        assert np.array_equal(block_amd(cp, ri), cs_amd(cp, ri)), (n, len(edges))
        count += 1
    assert count == 121


def test_block_amd_edge_cases():
    from openslam_g2o_b200 import block_amd
    assert list(block_amd([0, 1], [0])) == [0]
    p = block_amd([0, 1, 3], [0, 0, 1])
    assert sorted(p) == [0, 1]


@pytest.mark.parametrize("name", FIXTURES)
def test_product_ordering_matches_golden_hash(name):
    """through the real product path (host-only context): pattern from buildStructure -> ordering -> nnz(L)"""
    import openslam_g2o_b200 as g
    from oracle_binding import fnv1a64
    fx = load_fixture(name)
    opt = g.SparseOptimizer(device=-1)
    feed_fixture(opt, fx)
    assert opt.setup_cli() == int(fx["gauge"])
    opt.initialize_optimization()
    opt._ensure_uploaded()
    ctx = opt.context
    assert ctx.build_structure()
    perm = ctx.block_ordering()
    assert np.array_equal(perm, fx["perm"])
    assert fnv1a64(perm) == str(fx["perm_hash"])
    assert ctx.factor_nnz() == int(fx["lnz"])
    d = ctx.dims()
    assert [d["numPoses"], d["sizePoses"], d["numEdges"]] == [int(fx["dims"][0]), int(fx["dims"][2]), int(fx["dims"][4])]
