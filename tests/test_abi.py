"""The C-ABI library loads on a CPU-only box, exports every symbol include/g2o_b200.h declares, and refuses to
compute without a device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest
from helpers import ROOT


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "g2o_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", txt)) - {"b200_allreduce_fn"})


def test_library_exports_every_declared_symbol():
    import openslam_g2o_b200._lib as L
    lib = C.CDLL(L.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 60
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    # and the Python binding declares a signature for each of them
    undeclared = [s for s in syms if s not in L.EXPORTED_SYMBOLS]
    assert not undeclared, undeclared


def test_no_device_means_loud_failure_not_fallback():
    import openslam_g2o_b200 as g
    import openslam_g2o_b200._lib as L
    if L.lib.b200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(g.B200Error) as ei:
        g.SolverContext(0)
    assert ei.value.code == L.ERR_NO_DEVICE
    with pytest.raises(g.B200Error):
        g.LinearSolverB200(0)


def test_host_only_context_refuses_every_compute_call():
    import numpy as np
    import openslam_g2o_b200 as g
    import openslam_g2o_b200._lib as L
    from openslam_g2o_b200 import synth
    opt = g.SparseOptimizer(device=-1)
    synth.feed(synth.sphere(5, 3, seed=1), opt)
    opt.setup_cli()
    opt.initialize_optimization()
    opt._ensure_uploaded()
    ctx = opt.context
    assert ctx.build_structure()
    for call in (ctx.compute_active_errors, ctx.build_system, ctx.solve, ctx.update, ctx.push, ctx.x, ctx.b,
                 lambda: ctx.optimize(L.LEVENBERG, 1)):
        with pytest.raises(g.B200Error) as ei:
            call()
        assert ei.value.code == L.ERR_NO_DEVICE
    assert ctx.launch_count() == 0
    assert np.array_equal(np.sort(ctx.block_ordering()), np.arange(14))
    # reading back what was ingested is not compute: a host-only context returns the estimates as handed over
    est = ctx.estimates(L.VERTEX_SE3, 15)
    assert est.shape == (15, 12) and np.isfinite(est).all() and np.abs(est).max() > 0


def test_product_never_imports_the_oracle():
    """the oracle is test infrastructure: nothing under openslam_g2o_b200/ may mention it"""
    pkg = os.path.join(ROOT, "openslam_g2o_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".h", "Makefile")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "liboracle" not in txt and "oracle_binding" not in txt and "oracle/" not in txt.replace("the oracle/", ""), (dirpath, f)


def test_committed_bench_line_follows_the_contract():
    """the bench line committed under profiles/ carries every key the measurement contract names"""
    import glob
    import json
    import re
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r01_bench_venice_stage*.json")),
                   key=lambda f: int(re.search(r"stage(\d+)", f).group(1)))
    assert files
    d = json.loads(open(files[-1]).read().strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["dtype"] == "f64" and d["vs_baseline"] is None and "workload" in d["config"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["gpu_launches"] > 0
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    c = d["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(c) and c["kind"] in ("port", "reference")
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_header_is_plain_c_and_binds_from_a_c_host(tmp_path):
    """include/g2o_b200.h compiles as C99 (-pedantic) and a C program drives the standalone host + structure phase
    through it; the ordering it prints is the one the Python binding sees"""
    import subprocess
    import numpy as np
    import openslam_g2o_b200 as g
    import openslam_g2o_b200._lib as L
    inc = os.path.join(ROOT, "include")
    subprocess.check_call(["gcc", "-x", "c", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only",
                           os.path.join(inc, "g2o_b200.h")])
    exe = str(tmp_path / "c_host")
    libdir = os.path.dirname(L.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", inc,
                           os.path.join(ROOT, "tests", "csrc", "c_host.c"), "-o", exe, L.LIB_PATH,
                           "-Wl,-rpath," + libdir])
    out = subprocess.check_output([exe], text=True).strip().splitlines()
    assert out[1].startswith("g2o_b200")
    fields = dict(kv.split("=", 1) for kv in out[0].split(" "))
    assert fields["poses"] == "5" and fields["edges"] == "6" and fields["nodevice"] == "1"
    # same graph through the Python mirror
    opt = g.SparseOptimizer(device=-1)
    est = np.array([[i, 0.1 * i, 0.05 * i] for i in range(6)], dtype=float)
    opt.add_vertices(g.VERTEX_SE2, np.arange(6), est)
    pay = np.tile(np.array([1.0, 0.1, 0.05, 10, 0, 0, 10, 0, 10]), (6, 1))
    opt.add_edges(g.EDGE_SE2, np.arange(6), (np.arange(6) + 1) % 6, pay)
    opt.setup_cli()
    opt.initialize_optimization()
    opt._ensure_uploaded()
    assert opt.context.build_structure()
    assert fields["perm"] == ",".join(str(int(x)) for x in opt.context.block_ordering())
    assert int(fields["lnz"]) == opt.context.factor_nnz()
