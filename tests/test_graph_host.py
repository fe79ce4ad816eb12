"""Standalone host graph (openslam_g2o_b200/csrc/graph_host.cpp): loader, CLI gauge / marginalisation, index
mapping and save, against the oracle (CPU only)."""
import os

import numpy as np
import pytest
from conftest import REFERENCE_DATA, needs_oracle, needs_reference
from helpers import FIXTURES, feed_fixture, load_fixture, rel_err

FILES = {"manhattan3500": "2d/manhattan3500/manhattanOlson3500.g2o", "intel": "2d/intel/intel.g2o",
         "sphere_bignoise": "3d/sphere/sphere_bignoise_vertex3.g2o", "garage": "3d/garage/parking-garage.g2o"}


@needs_reference
@pytest.mark.parametrize("name", FIXTURES)
def test_text_loader_equals_fixture_ingest(name, tmp_path):
    """parsing the .g2o text gives the same graph as feeding the parsed numbers (and as the oracle's loader)"""
    import openslam_g2o_b200 as g
    fx = load_fixture(name)
    a = g.SparseOptimizer(device=-1)
    assert a.load(os.path.join(REFERENCE_DATA, FILES[name]))
    b = g.SparseOptimizer(device=-1)
    feed_fixture(b, fx)
    for o in (a, b):
        assert o.setup_cli() == int(fx["gauge"])
        o.initialize_optimization()
    ids = fx["final_ids"]
    for vid in ids[:: max(1, len(ids) // 200)]:
        assert np.array_equal(a.vertex_estimate(int(vid)), b.vertex_estimate(int(vid)))
        assert a.vertex_info(int(vid)) == b.vertex_info(int(vid))
    # index mapping equals the oracle's (golden)
    hidx = {int(i): int(h) for i, h in zip(fx["final_ids"], fx["final_hidx"])}
    flags = {int(i): int(f) for i, f in zip(fx["final_ids"], fx["final_flags"])}
    for vid in ids[:: max(1, len(ids) // 500)]:
        info = a.vertex_info(int(vid))
        assert info["hessian_index"] == hidx[int(vid)]
        assert info["fixed"] == bool(flags[int(vid)] & 1) and info["marginalized"] == bool(flags[int(vid)] & 2)
    # save -> load round trip keeps the graph (estimates to text precision %.17g = exact for SE2)
    out = tmp_path / "saved.g2o"
    assert a.save(out)
    c = g.SparseOptimizer(device=-1)
    assert c.load(out)
    vc_a, ec_a = a.counts()
    vc_c, ec_c = c.counts()
    assert list(vc_a) == list(vc_c) and list(ec_a) == list(ec_c)
    assert c.vertex_info(int(fx["gauge"]))["fixed"]
    # SE3: write() emits a NORMALISED quaternion (toVectorQT) while read() had used the file's un-normalised one
    # (|q| = 1 to 6 digits), so R changes at the 1e-6 level exactly as it does in the reference
    tol = 1e-12 if int(fx["v_kind"][0]) == 0 else 1e-5
    for vid in ids[:: max(1, len(ids) // 100)]:
        assert rel_err(c.vertex_estimate(int(vid)), a.vertex_estimate(int(vid))) < tol


@needs_oracle
def test_ba_setup_matches_oracle():
    """BA: gauge = first camera, points marginalized, poses indexed before landmarks (sparse_optimizer.cpp:166-190)"""
    import openslam_g2o_b200 as g
    from oracle_binding import Oracle
    from openslam_g2o_b200 import synth
    p = synth.venice_like(9, 60, seed=4)
    a = g.SparseOptimizer(device=-1)
    o = Oracle()
    synth.feed(p, a)
    synth.feed(p, o)
    assert a.setup_cli() == o.setup_cli(True) == 0
    a.initialize_optimization()
    o.initialize_optimization()
    ids, kinds, hidx, flags = o.vertices()
    for i, k, h, f in zip(ids, kinds, hidx, flags):
        info = a.vertex_info(int(i))
        assert (info["kind"], info["hessian_index"], info["fixed"], info["marginalized"]) == (int(k), int(h), bool(f & 1), bool(f & 2))
    a._ensure_uploaded()
    assert a.context.build_structure()
    o.algorithm_init()
    o.build_structure()
    d, od = a.context.dims(), o.dims()
    for key in ("numPoses", "numLandmarks", "sizePoses", "sizeLandmarks", "numEdges"):
        assert d[key] == od[key], key
    # Hschur / Hpl patterns identical to BlockSolver::buildStructure's
    for which in (2, 3):
        n = g.lib.b200_get_blocks(a.context.handle, which, None, None, None)
        rows, cols = np.zeros(n, np.int32), np.zeros(n, np.int32)
        g.lib.b200_get_blocks(a.context.handle, which, rows.ctypes.data, cols.ctypes.data, None)
        orows, ocols, _ = o.blocks(which)
        assert np.array_equal(rows, orows) and np.array_equal(cols, ocols), which


def test_loader_edge_cases(tmp_path):
    import openslam_g2o_b200 as g
    # comments, unknown tags, FIX, vertices declared after an edge that creates them (createEdges=true path)
    txt = """# a comment
VERTEX_SE2 0 0 0 0
UNKNOWN_TAG 1 2 3
EDGE_SE2 0 1 1 0 0.5 10 0 0 10 0 10
VERTEX_SE2 1 9 9 9
VERTEX_SE2 2 2 0 1.0
EDGE_SE2 2 1 -1 0 -0.5 10 0 0 10 0 10
FIX 2

"""
    f = tmp_path / "t.g2o"
    f.write_text(txt)
    o = g.SparseOptimizer(device=-1)
    assert o.load(f)
    vc, ec = o.counts()
    assert list(vc) == [3, 0, 0, 0, 0, 0] and list(ec) == [2, 0, 0, 0, 0, 0]
    # vertex 1 was created by the first edge: estimate = x0 * z; the later VERTEX line is ignored (duplicate id)
    assert np.allclose(o.vertex_estimate(1), [1.0, 0.0, 0.5])
    assert o.setup_cli() == -1  # vertex 2 is already fixed -> no gauge needed
    o.initialize_optimization()
    assert o.vertex_info(2)["hessian_index"] == -1 and o.vertex_info(0)["hessian_index"] == 0
    # empty graph
    e = g.SparseOptimizer(device=-1)
    with pytest.raises(g.B200Error):
        e.initialize_optimization()


@needs_oracle
def test_expmap_ba_setup_text_round_trip_and_structure(tmp_path):
    """VERTEX_SE3:EXPMAP / PARAMS_CAMERAPARAMETERS / EDGE_PROJECT_XYZ2UV:EXPMAP (types/sba/types_six_dof_expmap.cpp):
    programmatic ingest == text loader == oracle loader; save() writes cam2world again; the structure phase gives
    the oracle's index mapping and Hpl / Hschur patterns"""
    import openslam_g2o_b200 as g
    from oracle_binding import Oracle
    from openslam_g2o_b200 import synth
    p = synth.expmap_ba(9, 60, seed=4)
    path = tmp_path / "expmap.g2o"
    synth.write_g2o(p, path)
    a, b = g.SparseOptimizer(device=-1), g.SparseOptimizer(device=-1)
    o, o2 = Oracle(), Oracle()
    synth.feed(p, a)
    synth.feed(p, o)
    assert b.load(path) and o2.load(path)
    vc, ec = b.counts()
    assert list(vc) == [0, 0, 0, 60, 9, 0] and list(ec) == [0, 0, 0, len(p["edge_v0"]), 0, 0]
    for vid in list(p["cam_ids"]) + list(p["point_ids"][::7]):
        ea, eb = a.vertex_estimate(int(vid)), b.vertex_estimate(int(vid))
        assert np.array_equal(ea, eb)
        n = len(o.vertex_estimate(int(vid)))
        assert rel_err(ea[:n], o.vertex_estimate(int(vid))) < 1e-15
        assert np.array_equal(o.vertex_estimate(int(vid)), o2.vertex_estimate(int(vid)))
    assert np.array_equal(a.vertex_estimate(0)[7:], [1000.0, 1000.0, 320.0, 240.0, 0.0])
    # save -> load reproduces the estimates (cam2world is written, inverted again on load) to rounding
    out = tmp_path / "saved.g2o"
    assert b.save(out)
    assert "PARAMS_CAMERAPARAMETERS 0 1000" in out.read_text().splitlines()[0]
    c = g.SparseOptimizer(device=-1)
    assert c.load(out)
    for vid in p["cam_ids"]:
        assert rel_err(c.vertex_estimate(int(vid)), b.vertex_estimate(int(vid))) < 1e-14
    # gauge, marginalisation, index mapping, patterns
    assert a.setup_cli() == o.setup_cli(True) == 0
    a.initialize_optimization()
    o.initialize_optimization()
    ids, kinds, hidx, flags = o.vertices()
    for i, k, h, f in zip(ids, kinds, hidx, flags):
        info = a.vertex_info(int(i))
        assert (info["kind"], info["hessian_index"], info["fixed"], info["marginalized"]) == (int(k), int(h), bool(f & 1), bool(f & 2))
    a._ensure_uploaded()
    assert a.context.build_structure()
    o.algorithm_init()
    o.build_structure()
    d, od = a.context.dims(), o.dims()
    for key in ("numPoses", "numLandmarks", "sizePoses", "sizeLandmarks", "numEdges"):
        assert d[key] == od[key], key
    for which in (2, 3):
        n = g.lib.b200_get_blocks(a.context.handle, which, None, None, None)
        rows, cols = np.zeros(n, np.int32), np.zeros(n, np.int32)
        g.lib.b200_get_blocks(a.context.handle, which, rows.ctypes.data, cols.ctypes.data, None)
        orows, ocols, _ = o.blocks(which)
        assert np.array_equal(rows, orows) and np.array_equal(cols, ocols), which
    # an edge that names an unknown parameter is rejected (resolveParameters), as is mixing the two camera models
    with pytest.raises(g.B200Error):
        a.add_edges(g.EDGE_XYZ2UV, [int(p["point_ids"][0])], [0], np.array([[5.0, 1, 2, 1, 0, 1]]))
    with pytest.raises(g.B200Error):
        a.add_edges(g.EDGE_P2MC, [int(p["point_ids"][0])], [0], np.array([[1.0, 2.0]]))
    # two parameters with different values on one pose: unsupported, reported
    a.add_camera_parameters(1, 500.0, 0.0, 0.0, 0.0)
    with pytest.raises(g.B200Error) as ei:
        a.add_edges(g.EDGE_XYZ2UV, [int(p["point_ids"][0])], [0], np.array([[1.0, 1, 2, 1, 0, 1]]))
    assert ei.value.code == g._lib.ERR_UNSUPPORTED


def test_parallel_loader_equals_sequential(tmp_path, monkeypatch):
    """the loader tokenises the text in per-thread chunks and applies the records in file order: any thread count gives
    the graph a line-by-line reader builds - including a vertex that an EDGE line creates before its VERTEX line and FIX
    lines that refer to vertices of an earlier chunk"""
    import openslam_g2o_b200 as g
    from openslam_g2o_b200 import synth
    p = synth.sphere(40, 30, seed=5)          # 1200 poses / ~4700 edges: > 1 MiB of text, i.e. several chunks
    path = tmp_path / "s.g2o"
    synth.write_g2o(p, path)
    txt = path.read_text().splitlines()
    nvert = len(p["vertex_ids"])
    # move one VERTEX line behind the edges (the first edge that touches it creates it), fix two vertices at the end
    moved = txt.pop(700)
    txt += [moved, "FIX 3 900", "# trailing comment", ""]
    path.write_text("\n".join(txt))
    assert path.stat().st_size > (1 << 20)
    graphs = []
    for threads in ("1", "2", "7"):
        monkeypatch.setenv("G2O_B200_LOADER_THREADS", threads)
        o = g.SparseOptimizer(device=-1)
        assert o.load(path)
        graphs.append(o)
    vc, ec = graphs[0].counts()
    assert vc[g.VERTEX_SE3] == nvert and ec[g.EDGE_SE3] == len(p["edge_v0"])
    for o in graphs[1:]:
        assert [list(c) for c in o.counts()] == [list(c) for c in graphs[0].counts()]
        for vid in list(range(0, nvert, 37)) + [3, 700, 900]:
            assert np.array_equal(o.vertex_estimate(vid), graphs[0].vertex_estimate(vid))
            assert o.vertex_info(vid) == graphs[0].vertex_info(vid)
    assert graphs[0].vertex_info(3)["fixed"] and graphs[0].vertex_info(900)["fixed"]
    # vertex 700 was created by an edge (initialEstimate from its neighbour), the later VERTEX line is a duplicate
    assert not np.array_equal(graphs[0].vertex_estimate(700)[9:], p["vertex_payload"][700][:3])
    for o in graphs:
        assert o.setup_cli() == -1
        o.initialize_optimization()
    assert graphs[2].vertex_info(901) == graphs[0].vertex_info(901)


def test_structure_plan_digest_is_deterministic_and_sensitive(monkeypatch):
    """b200_debug_upload_digest: the digest over everything the structure phase prepares for the device is identical
    for identical inputs (the plan does not depend on allocation addresses, thread timing or hash-map order) and changes
    when one observation moves to another camera"""
    import openslam_g2o_b200 as g
    from openslam_g2o_b200 import synth

    def digest(p):
        o = g.SparseOptimizer(device=-1)
        synth.feed(p, o)
        o.setup_cli()
        o.initialize_optimization()
        o._ensure_uploaded()
        g.lib.b200_debug_upload_digest(1)
        assert o.context.build_structure()
        return g.lib.b200_debug_upload_digest(1)
    p = synth.venice_like(40, 3000, seed=9)
    d0 = digest(p)
    assert d0 == digest(p) and d0 != 1469598103934665603
    # the host-side fork/join (csrc/host_parallel.h) never changes the plan: many tiny ranges on several threads
    monkeypatch.setenv("G2O_B200_HOST_GRAIN", "7")
    for threads in ("2", "5"):
        monkeypatch.setenv("G2O_B200_HOST_THREADS", threads)
        assert digest(p) == d0
    wide = synth.venice_like(30, 200, seed=2, fixed_obs=30)  # ranges grown beyond the default slot budget
    monkeypatch.setenv("G2O_B200_HOST_THREADS", "1")
    dw = digest(wide)
    monkeypatch.setenv("G2O_B200_HOST_THREADS", "4")
    assert digest(wide) == dw
    monkeypatch.delenv("G2O_B200_HOST_GRAIN")
    monkeypatch.delenv("G2O_B200_HOST_THREADS")
    # pose graph: the supernodal Cholesky plan (work items per tile, their order, descriptors) is part of the digest
    sp = synth.sphere(20, 20, seed=1)
    ds = digest(sp)
    monkeypatch.setenv("G2O_B200_HOST_GRAIN", "3")
    monkeypatch.setenv("G2O_B200_HOST_THREADS", "5")
    assert digest(sp) == ds
    monkeypatch.delenv("G2O_B200_HOST_GRAIN")
    monkeypatch.delenv("G2O_B200_HOST_THREADS")
    monkeypatch.setenv("G2O_B200_PANEL_COLS", "36")
    assert digest(sp) != ds
    q = dict(p)
    q["edge_v1"] = p["edge_v1"].copy()
    q["edge_v1"][17] = (q["edge_v1"][17] + 20) % 40
    assert digest(q) != d0


def test_expmap_poses_may_use_different_camera_parameters(tmp_path):
    """CameraParameters belong to the edges in g2o; the B200 path carries them in the pose rows, so different poses may
    name different parameters (a multi-camera rig) as long as each pose is consistent"""
    import openslam_g2o_b200 as g
    txt = """PARAMS_CAMERAPARAMETERS 0 1000 320 240 0
PARAMS_CAMERAPARAMETERS 7 650.5 300 250 0.1
VERTEX_SE3:EXPMAP 0 0 0 0 0 0 0 1
VERTEX_SE3:EXPMAP 1 0.5 0 0 0 0 0 1
VERTEX_XYZ 2 0.1 0.2 4
VERTEX_XYZ 3 -0.3 0.1 5
EDGE_PROJECT_XYZ2UV:EXPMAP 2 0 0 345 290 1 0 1
EDGE_PROJECT_XYZ2UV:EXPMAP 3 0 0 260 260 1 0 1
EDGE_PROJECT_XYZ2UV:EXPMAP 2 1 7 235 282 2 0.5 3
EDGE_PROJECT_XYZ2UV:EXPMAP 3 1 7 196 263 2 0.5 3
EDGE_PROJECT_XYZ2UV:EXPMAP 3 1 9 196 263 2 0.5 3
"""
    f = tmp_path / "rig.g2o"
    f.write_text(txt)
    o = g.SparseOptimizer(device=-1)
    assert o.load(f)
    vc, ec = o.counts()
    assert ec[g.EDGE_XYZ2UV] == 4          # the edge naming the unknown parameter 9 was dropped (resolveParameters)
    assert np.array_equal(o.vertex_estimate(0)[7:], [1000, 1000, 320, 240, 0])
    assert np.array_equal(o.vertex_estimate(1)[7:], [650.5, 650.5, 300, 250, 0.1])
    # estimate = inverse of the file's cam2world: identity rotation, translation negated
    assert np.array_equal(o.vertex_estimate(1)[:7], [-0.5, 0, 0, 0, 0, 0, 1])
    assert o.setup_cli() == 0
    o.initialize_optimization()
    o._ensure_uploaded()
    assert o.context.build_structure()
    out = tmp_path / "rig_out.g2o"
    assert o.save(out)
    lines = out.read_text().splitlines()
    assert lines[0].startswith("PARAMS_CAMERAPARAMETERS 0 ") and lines[1].startswith("PARAMS_CAMERAPARAMETERS 7 650.5")
    assert sum(ln.startswith("EDGE_PROJECT_XYZ2UV:EXPMAP") for ln in lines) == 4
    assert any(ln.startswith("EDGE_PROJECT_XYZ2UV:EXPMAP 3 1 7 ") for ln in lines)


def test_landmark_seen_by_more_cameras_than_a_range_holds_is_planned_not_refused(monkeypatch):
    """round 1 refused a landmark observed by more than 1400 cameras (shared-memory capacity of one Schur range); now such
    landmarks get one segment per camera pair (kernels.cuh: schur_wide_kernel) and the ranges around them stay contiguous
    runs of Hpl slots.  Host-only structure phase: segments = those of the ranges + k (k + 1) / 2 per wide landmark"""
    import openslam_g2o_b200 as g
    from openslam_g2o_b200 import synth
    cams = 1500
    p = dict(synth.venice_like(cams, 200, seed=3))
    wide_pt = p["point_ids"][57]
    p["edge_v0"] = np.concatenate([p["edge_v0"], np.full(cams, wide_pt)]).astype(np.int32)
    p["edge_v1"] = np.concatenate([p["edge_v1"], p["cam_ids"]]).astype(np.int32)
    p["edge_payload"] = np.concatenate([p["edge_payload"], np.zeros((cams, 2))])

    def plan(env):
        if env is None:
            monkeypatch.delenv("G2O_B200_SR_WIDE", raising=False)
        else:
            monkeypatch.setenv("G2O_B200_SR_WIDE", str(env))
        opt = g.SparseOptimizer(device=-1)
        opt.set_algorithm("lm_fix6_3")
        synth.feed(p, opt)
        opt.setup_cli(); opt.initialize_optimization(); opt._ensure_uploaded()
        assert opt.context.build_structure()
        info = opt.context.factor_info()
        opt.close()
        return info
    info = plan(None)
    k = cams - 1   # the gauge camera is fixed: no Hpl block for it
    assert info["schur_segments"] >= k * (k + 1) // 2
    # forcing every landmark with more than 3 cameras through the wide path changes the segments, not the contributions
    a, b = plan(1400), plan(3)
    assert a["schur_contributions"] == b["schur_contributions"] and a["hpl_slots"] == b["hpl_slots"]
    assert b["schur_segments"] > a["schur_segments"]   # (and more, shorter ranges: a wide landmark closes the range before it)


def test_pcg_solver_names_plan_on_the_host():
    """`*_pcg*` solver names (solvers/pcg/solver_pcg.cpp) select LinearSolverPCG inside the solver: the structure phase
    (incl. the symmetric block-row lists of the CG kernels) runs on a host-only context; compute stays refused without a GPU"""
    import openslam_g2o_b200 as g
    from openslam_g2o_b200 import synth
    for p, name in ((synth.venice_like(10, 200, seed=2), "lm_pcg6_3"), (synth.sphere(8, 5, seed=2), "gn_pcg6_3"),
                    (synth.landmark_slam_2d(20, 10, seed=2), "lm_pcg")):
        opt = g.SparseOptimizer(device=-1)
        opt.set_algorithm(name)
        synth.feed(p, opt)
        opt.setup_cli(); opt.initialize_optimization(); opt._ensure_uploaded()
        assert opt.context.build_structure()
        with pytest.raises(g.B200Error):
            opt.context.solve()
        opt.close()
    with pytest.raises(g.B200Error):
        g.SparseOptimizer(device=-1).set_algorithm("lm_pcg7_3")   # 7-dimensional poses: not provided
