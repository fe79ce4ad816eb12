"""Host-side mirror of the reference's interface for the hot path, over the C-ABI.

  SparseOptimizer   ~ g2o::SparseOptimizer + the g2o CLI setup  (core/sparse_optimizer.h:70-278,
                      apps/g2o_cli/g2o.cpp:211-320,542-553): load / setup / initializeOptimization / optimize
  SolverContext     ~ g2o::Solver (core/solver.h:44-149): buildStructure / buildSystem / setLambda / solve / ...
  LinearSolverB200  ~ g2o::LinearSolver<MatrixType> (core/linear_solver.h:40-81)

Same method names and argument meaning as the reference (snake_case), same error behaviour (bool / int
results, B200Error only for misuse or CUDA failures).  No numerical work happens in Python.
"""
import ctypes as C

import numpy as np

from . import _lib as L
from ._lib import B200Error, IterStats, lib


def _check(rc, handle=None, ls=False, graph=False):
    if rc >= 0:
        return rc
    if graph:
        msg = lib.b200_graph_last_error(handle)
    elif ls:
        msg = lib.b200_ls_last_error(handle)
    else:
        msg = lib.b200_last_error(handle)
    raise B200Error(rc, (msg or b"").decode())


def comm_unique_id():
    """rank 0: the NCCL unique id (bytes) to hand to every rank's SolverContext.comm_init"""
    buf = C.create_string_buffer(L.COMM_ID_BYTES)
    rc = lib.b200_comm_unique_id(buf, L.COMM_ID_BYTES)
    if rc != 0:
        raise B200Error(rc, (lib.b200_last_error(None) or b"").decode())
    return buf.raw


def block_amd(colptr, rowidx):
    """Block fill-reducing ordering (bit-exact twin of cs_amd(1, .), EXTERNAL/csparse/cs_amd.c). Host only."""
    colptr = L.as_i32(colptr)
    rowidx = L.as_i32(rowidx)
    n = len(colptr) - 1
    perm = np.zeros(n, np.int32)
    _check(lib.b200_block_amd(n, L.ptr(colptr), L.ptr(rowidx), L.ptr(perm)))
    return perm


# solver names of the reference factory (solvers/csparse/solver_csparse.cpp:34-124,
# solvers/cholmod/solver_cholmod.cpp:41-132) that this path covers
_SOLVERS = {}
for _alg, _aid in (("gn", L.GAUSS_NEWTON), ("lm", L.LEVENBERG)):
    for _fix, _pd, _ld in (("fix3_2", 3, 2), ("fix6_3", 6, 3)):
        for _suffix in ("", "_cholmod", "_b200"):
            _SOLVERS["%s_%s%s" % (_alg, _fix, _suffix)] = (_aid, _pd, _ld)
    # LinearSolverPCG as the linear solver (solvers/pcg/solver_pcg.cpp: gn_pcg, gn_pcg3_2, gn_pcg6_3, lm_pcg, ...)
    for _name, _pd in (("pcg", -1), ("pcg3_2", 3), ("pcg6_3", 6)):
        for _suffix in ("", "_b200"):
            _SOLVERS["%s_%s%s" % (_alg, _name, _suffix)] = (_aid, _pd, 3 if _pd == 6 else 2 if _pd == 3 else -1)
    # variable block sizes (BlockSolverX; solvers/csparse/solver_csparse.cpp:53-55: requiresMarginalize = false): pose graphs
    # and landmark SLAM (SE2 + XY, SE3 + XYZ) with every vertex in one system
    for _suffix in ("", "_cholmod", "_b200"):
        _SOLVERS["%s_var%s" % (_alg, _suffix)] = (_aid, -1, -1)


class SolverContext:
    """Level 2/3 boundary: one b200_ctx (device-resident system + estimates)."""

    def __init__(self, device=0):
        h = C.c_void_p()
        rc = lib.b200_create(device, C.byref(h))
        if rc != 0:
            raise B200Error(rc, (lib.b200_last_error(None) or b"").decode())
        self._h = h
        self._keep = []

    def close(self):
        if self._h:
            lib.b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    # ---- ingest
    def set_vertices(self, kind, estimates, hessian_index, marginalized=None):
        est = L.as_f64(estimates)
        hidx = L.as_i32(hessian_index)
        marg = None if marginalized is None else np.ascontiguousarray(marginalized, dtype=np.uint8)
        _check(lib.b200_set_vertices(self._h, kind, len(hidx), L.ptr(est), L.ptr(hidx), L.ptr(marg)), self._h)

    def set_edges(self, kind, vi, vj, measurement, information):
        vi, vj = L.as_i32(vi), L.as_i32(vj)
        meas, info = L.as_f64(measurement), L.as_f64(information)
        _check(lib.b200_set_edges(self._h, kind, len(vi), L.ptr(vi), L.ptr(vj), L.ptr(meas), L.ptr(info)), self._h)

    def set_allreduce(self, fn, rank, world_size):
        cb = L.ALLREDUCE_FN(fn)
        self._keep.append(cb)
        _check(lib.b200_set_allreduce(self._h, cb, None, rank, world_size), self._h)

    def comm_init(self, unique_id, rank, world_size):
        """native NCCL communicator of the landmark-sharded path (collective call: every rank, same unique_id)"""
        buf = C.create_string_buffer(bytes(unique_id), L.COMM_ID_BYTES)
        _check(lib.b200_comm_init(self._h, buf, rank, world_size), self._h)

    def comm_destroy(self):
        lib.b200_comm_destroy(self._h)

    # ---- g2o::Solver
    def build_structure(self):
        return _check(lib.b200_build_structure(self._h), self._h) == 0

    def compute_active_errors(self):
        chi2 = C.c_double()
        _check(lib.b200_compute_active_errors(self._h, C.byref(chi2)), self._h)
        return chi2.value

    def build_system(self):
        return _check(lib.b200_build_system(self._h), self._h) == 0

    def set_lambda(self, lam, backup=False):
        return _check(lib.b200_set_lambda(self._h, float(lam), int(backup)), self._h) == 0

    def restore_diagonal(self):
        _check(lib.b200_restore_diagonal(self._h), self._h)

    def solve(self):
        """True on success, False when the system is not positive definite (like Solver::solve)."""
        return _check(lib.b200_solve(self._h), self._h) == 0

    def update(self):
        _check(lib.b200_update(self._h), self._h)

    def push(self):
        _check(lib.b200_push(self._h), self._h)

    def pop(self):
        _check(lib.b200_pop(self._h), self._h)

    def discard_top(self):
        _check(lib.b200_discard_top(self._h), self._h)

    # ---- g2o::OptimizationAlgorithm / SparseOptimizer::optimize
    def optimize(self, algorithm, iterations):
        stats = (IterStats * max(iterations, 1))()
        n = _check(lib.b200_optimize(self._h, algorithm, iterations, stats), self._h)
        return n, list(stats)[:max(n, 0)] if n > 0 else list(stats)[:iterations]

    def algorithm_solve(self, algorithm, iteration):
        st = IterStats()
        rc = lib.b200_algorithm_solve(self._h, algorithm, iteration, C.byref(st))
        _check(rc, self._h)  # < 0 is always a hard error; OK / Terminate / SOLVE_FAIL are >= 1
        return rc, st

    def compute_marginals(self, pairs):
        """blocks (row, col) of Hpp^-1 -> array [len(pairs), d, d] (Solver::computeMarginals); after build_system()"""
        rows = L.as_i32([p[0] for p in pairs])
        cols = L.as_i32([p[1] for p in pairs])
        d = self.dims()["poseDim"]
        out = np.zeros((len(pairs), d, d))
        rc = _check(lib.b200_compute_marginals(self._h, len(pairs), L.ptr(rows), L.ptr(cols), L.ptr(out)), self._h)
        if rc != 0:
            return None
        return np.ascontiguousarray(np.transpose(out, (0, 2, 1)))  # column-major blocks -> [row, col]

    def set_linear_solver(self, kind="cholesky", tolerance=1e-6, absolute_tolerance=True, max_iterations=-1):
        """linear solver of the (reduced) pose system: "cholesky" (default) or "pcg" = LinearSolverPCG (solvers/pcg) with its
        setTolerance / setAbsoluteTolerance / setMaxIterations; before build_structure()"""
        _check(lib.b200_set_linear_solver(self._h, {"cholesky": 0, "pcg": 1}[kind], float(tolerance), int(bool(absolute_tolerance)),
                                          int(max_iterations)), self._h)

    def linear_solver_iterations(self):
        """G2OBatchStatistics::iterationsLinearSolver of the last solve (PCG)"""
        return int(lib.b200_get_linear_solver_iterations(self._h))

    def set_ordering(self, nd_levels=0):
        """0: block AMD (the reference's ordering, default); k > 0: nested dissection with 2^k parts on top of it"""
        _check(lib.b200_set_ordering(self._h, int(nd_levels)), self._h)

    def set_robust_kernel(self, kind, delta=1.0):
        """one robust kernel on every edge (g2o -robustKernel NAME -robustKernelWidth delta)"""
        _check(lib.b200_set_robust_kernel(self._h, ROBUST_KERNELS[kind] if isinstance(kind, str) else int(kind),
                                          float(delta)), self._h)

    def set_terminate(self, fn):
        """SparseOptimizer::setForceStopFlag / terminate(): fn() -> truthy stops after the current trial / iteration"""
        cb = L.TERMINATE_FN((lambda _u: 1 if fn() else 0) if fn else 0)
        self._keep.append(cb)
        _check(lib.b200_set_terminate(self._h, cb, None), self._h)

    def set_lm_params(self, user_lambda_init=0.0, max_trials_after_failure=10):
        _check(lib.b200_set_lm_params(self._h, user_lambda_init, max_trials_after_failure), self._h)

    # ---- read-back
    def dims(self):
        d = np.zeros(8, np.int32)
        _check(lib.b200_get_dims(self._h, L.ptr(d)), self._h)
        return dict(zip(("numPoses", "numLandmarks", "sizePoses", "sizeLandmarks", "numEdges", "numVertices",
                         "poseDim", "landmarkDim"), (int(v) for v in d)))

    def _vec(self, fn):
        d = self.dims()
        out = np.zeros(d["sizePoses"] + d["sizeLandmarks"])
        _check(fn(self._h, L.ptr(out)), self._h)
        return out

    def x(self):
        return self._vec(lib.b200_get_x)

    def b(self):
        return self._vec(lib.b200_get_b)

    def hessian_diagonal(self):
        return self._vec(lib.b200_get_hessian_diagonal)

    def bschur(self):
        out = np.zeros(self.dims()["sizePoses"])
        _check(lib.b200_get_bschur(self._h, L.ptr(out)), self._h)
        return out

    def estimates(self, kind, n):
        out = np.zeros((n, L.VERTEX_EST_LEN[kind]))
        _check(lib.b200_get_estimates(self._h, kind, L.ptr(out)), self._h)
        return out

    def set_estimates(self, kind, estimates):
        est = L.as_f64(estimates)
        _check(lib.b200_set_estimates(self._h, kind, L.ptr(est)), self._h)

    def get_estimates_into(self, kind, out):
        _check(lib.b200_get_estimates(self._h, kind, L.ptr(out)), self._h)

    def blocks(self, which):
        """(rows, cols, values[n, r, c]) of Hpp(0) / Hll(1) / Hpl(2) / Hschur(3), SparseBlockMatrix order."""
        n = _check(lib.b200_get_blocks(self._h, which, None, None, None), self._h)
        d = self.dims()
        rd, cd = {0: (d["poseDim"],) * 2, 1: (3, 3), 2: (d["poseDim"], 3), 3: (d["poseDim"],) * 2}[which]
        rows = np.zeros(n, np.int32)
        cols = np.zeros(n, np.int32)
        vals = np.zeros((n, cd, rd))
        _check(lib.b200_get_blocks(self._h, which, L.ptr(rows), L.ptr(cols), L.ptr(vals)), self._h)
        return rows, cols, np.transpose(vals, (0, 2, 1))  # column-major payload -> [n, r, c]

    def block_ordering(self):
        n = _check(lib.b200_get_block_ordering(self._h, None), self._h)
        p = np.zeros(n, np.int32)
        _check(lib.b200_get_block_ordering(self._h, L.ptr(p)), self._h)
        return p

    def factor_nnz(self):
        return int(lib.b200_get_factor_nnz(self._h))

    def factor_info(self):
        out = np.zeros(24, np.int64)
        _check(lib.b200_get_factor_info(self._h, L.ptr(out)), self._h)
        return dict(zip(("supernodes", "tasks", "levels", "max_panel_rows", "max_panel_cols", "factor_doubles",
                         "flow_tasks", "schur_ranges", "schur_segments", "schur_contributions", "hpl_slots",
                         "schur_range_smem", "factor_flops", "chain_links", "chain_flops", "reserved", "wide_tiles",
                         "update_items", "update_item_flops", "split_tile_slots", "subtree_flops"), (int(v) for v in out)))

    def launch_count(self):
        return int(lib.b200_get_launch_count(self._h))

    def set_profiling(self, on):
        _check(lib.b200_set_profiling(self._h, int(on)), self._h)

    def phase_times(self):
        names = ("errors", "linearize", "schur", "factor", "trisolve", "update", "backsub", "linearize_cams", "gather",
                 "schur_inv", "scale", "collective", "chol_scatter", "chol_factor_flow", "schur_finish", "chol_chain",
                 "chol_chain_backward", "unused17", "unused18", "chol_backward")
        out = {}
        for i, nme in enumerate(names):
            s, c = C.c_double(), C.c_int64()
            lib.b200_get_phase_time(self._h, i, C.byref(s), C.byref(c))
            out[nme] = (s.value, c.value)
        return out

    def stream(self):
        return lib.b200_get_stream(self._h)

    def synchronize(self):
        _check(lib.b200_synchronize(self._h), self._h)


# names of the robust kernel factory (core/robust_kernel_impl.cpp:129-136)
ROBUST_KERNELS = {"none": 0, "Huber": 1, "PseudoHuber": 2, "Cauchy": 3, "Saturated": 4, "DCS": 5}


class SparseOptimizer:
    """g2o::SparseOptimizer as driven by the `g2o` binary, for the configured graph families.

    opt = SparseOptimizer(); opt.set_algorithm("lm_fix6_3"); opt.load(path); opt.setup_cli()
    opt.initialize_optimization(); n = opt.optimize(10)
    """

    def __init__(self, device=0, shard=0, num_shards=1):
        g = C.c_void_p()
        _check(lib.b200_graph_create(C.byref(g)))
        self._g = g
        self._device = device
        self._ctx = None
        self._algorithm = L.LEVENBERG
        self._requires_marginalize = True
        self._shard, self._num_shards = shard, num_shards
        self._uploaded = False
        self.batch_statistics = []

    def close(self):
        if self._ctx is not None:
            self._ctx.close()
            self._ctx = None
        if self._g:
            lib.b200_graph_destroy(self._g)
            self._g = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # OptimizationAlgorithmFactory::construct (apps/g2o_cli/g2o.cpp:211-212)
    def set_algorithm(self, name):
        if name not in _SOLVERS:
            raise B200Error(L.ERR_UNSUPPORTED, "solver '%s' is not provided by the B200 path (have: %s)"
                            % (name, ", ".join(sorted(_SOLVERS))))
        self._algorithm = _SOLVERS[name][0]
        self._requires_marginalize = _SOLVERS[name][1] > 0  # all fix* solvers (solver_cholmod.cpp:115-121); var: false
        self._linear_solver = "pcg" if "_pcg" in name else "cholesky"
        if self._ctx is not None:
            self._ctx.set_linear_solver(self._linear_solver)

    def load(self, path):
        return _check(lib.b200_graph_load(self._g, str(path).encode()), self._g, graph=True) == 0

    def add_vertices(self, kind, ids, payload):
        ids = L.as_i32(ids)
        payload = L.as_f64(payload)
        _check(lib.b200_graph_add_vertices(self._g, kind, len(ids), L.ptr(ids), L.ptr(payload), payload.shape[1]),
               self._g, graph=True)

    def add_edges(self, kind, id1, id2, payload):
        id1, id2 = L.as_i32(id1), L.as_i32(id2)
        payload = L.as_f64(payload)
        _check(lib.b200_graph_add_edges(self._g, kind, len(id1), L.ptr(id1), L.ptr(id2), L.ptr(payload),
                                        payload.shape[1]), self._g, graph=True)

    def set_fixed(self, vid, fixed=True):
        _check(lib.b200_graph_set_fixed(self._g, vid, int(fixed)), self._g, graph=True)

    def add_camera_parameters(self, pid, focal_length, cx, cy, baseline):
        """OptimizableGraph::addParameter(CameraParameters) (types/sba/types_six_dof_expmap.h:45-80)"""
        _check(lib.b200_graph_add_camera_parameters(self._g, int(pid), float(focal_length), float(cx), float(cy),
                                                     float(baseline)), self._g, graph=True)

    def add_se3_offset(self, pid, xyz_qxyzw):
        """OptimizableGraph::addParameter(ParameterSE3Offset) (types/slam3d/parameter_se3_offset.h): x y z qx qy qz qw"""
        o = L.as_f64(xyz_qxyzw)
        _check(lib.b200_graph_add_se3_offset(self._g, int(pid), L.ptr(o)), self._g, graph=True)

    def set_robust_kernel(self, name, width=1.0):
        """`g2o -robustKernel name -robustKernelWidth width` (apps/g2o_cli/g2o.cpp:322-336): every edge gets the kernel"""
        self.context.set_robust_kernel(name, width)

    def set_edge_robust_kernel(self, edge_indices, name, width=1.0):
        """Edge::setRobustKernel on individual edges (indices in the order the edges were added / read): e.g. a kernel on the
        loop closures only.  Edges without one of their own then carry no kernel."""
        for k in edge_indices:
            _check(lib.b200_graph_set_edge_robust_kernel(self._g, int(k), ROBUST_KERNELS[name], float(width)), self._g, graph=True)
        self._uploaded = False

    def setup_cli(self):
        """gauge + marginalisation exactly as the g2o binary does (apps/g2o_cli/g2o.cpp:272-320)."""
        return lib.b200_graph_setup_cli(self._g, int(self._requires_marginalize))

    def initialize_optimization(self):
        rc = _check(lib.b200_graph_initialize(self._g), self._g, graph=True)
        self._uploaded = False
        return rc == 0

    @property
    def context(self):
        if self._ctx is None:
            self._ctx = SolverContext(self._device)
            if getattr(self, "_linear_solver", "cholesky") != "cholesky":
                self._ctx.set_linear_solver(self._linear_solver)
        return self._ctx

    def _ensure_uploaded(self):
        if not self._uploaded:
            _check(lib.b200_graph_upload(self._g, self.context.handle, self._shard, self._num_shards), self._g,
                   graph=True)
            self._uploaded = True

    def optimize(self, iterations):
        self._ensure_uploaded()
        n, stats = self.context.optimize(self._algorithm, iterations)
        self.batch_statistics = stats
        return n

    def compute_active_errors(self):
        self._ensure_uploaded()
        if not self.context.build_structure():
            return float("nan")
        return self.context.compute_active_errors()

    def chi2(self):
        return self.compute_active_errors()

    def sync_estimates(self):
        """device -> host graph (what the Level-3 adapter does before any host read of the estimates)"""
        _check(lib.b200_graph_download(self._g, self.context.handle), self._g, graph=True)

    def vertex_estimate(self, vid):
        out = np.zeros(12)
        n = _check(lib.b200_graph_get_estimate(self._g, vid, L.ptr(out)), self._g, graph=True)
        return out[:n].copy()

    def vertex_info(self, vid):
        out = np.zeros(4, np.int32)
        _check(lib.b200_graph_get_vertex_info(self._g, vid, L.ptr(out)), self._g, graph=True)
        return dict(kind=int(out[0]), hessian_index=int(out[1]), fixed=bool(out[2]), marginalized=bool(out[3]))

    def counts(self):
        vc = np.zeros(L.NUM_VERTEX_KINDS, np.int32)
        ec = np.zeros(L.NUM_EDGE_KINDS, np.int32)
        _check(lib.b200_graph_counts(self._g, L.ptr(vc), L.ptr(ec)), self._g, graph=True)
        return vc, ec

    def save(self, path):
        return _check(lib.b200_graph_save(self._g, str(path).encode()), self._g, graph=True) == 0


class LinearSolverB200:
    """g2o::LinearSolver<MatrixType> (core/linear_solver.h:40-81) on an upper block-CCS matrix."""

    def __init__(self, device=0):
        h = C.c_void_p()
        rc = lib.b200_ls_create(device, C.byref(h))
        if rc != 0:
            raise B200Error(rc, (lib.b200_ls_last_error(None) or b"").decode())
        self._h = h

    def close(self):
        if self._h:
            lib.b200_ls_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def init(self):
        return lib.b200_ls_init(self._h) == 0

    def solve(self, colptr, rowidx, values, b):
        """values[nblk, d, d] (row, col) blocks of the upper triangle incl. the diagonal.
        Returns x, or None if the matrix is not positive definite (solve() == false in the reference)."""
        colptr, rowidx = L.as_i32(colptr), L.as_i32(rowidx)
        values = np.asarray(values, dtype=np.float64)
        d = values.shape[1]
        vals_cm = np.ascontiguousarray(np.transpose(values, (0, 2, 1)))
        b = L.as_f64(b)
        x = np.zeros_like(b)
        rc = _check(lib.b200_ls_solve(self._h, len(colptr) - 1, d, L.ptr(colptr), L.ptr(rowidx), L.ptr(vals_cm),
                                      L.ptr(x), L.ptr(b)), self._h, ls=True)
        return None if rc == L.NOT_POSITIVE_DEFINITE else x

    def solve_pcg(self, colptr, rowidx, values, b, tolerance=1e-6, absolute_tolerance=True, max_iterations=-1):
        """LinearSolverPCG::solve (solvers/pcg/linear_solver_pcg.hpp:79-160): returns (x, iterations, residual); x is None
        when a diagonal block is not positive definite."""
        colptr, rowidx = L.as_i32(colptr), L.as_i32(rowidx)
        values = np.asarray(values, dtype=np.float64)
        d = values.shape[1]
        vals_cm = np.ascontiguousarray(np.transpose(values, (0, 2, 1)))
        b = L.as_f64(b)
        x = np.zeros_like(b)
        it, res = C.c_int32(0), C.c_double(0.0)
        rc = _check(lib.b200_ls_solve_pcg(self._h, len(colptr) - 1, d, L.ptr(colptr), L.ptr(rowidx), L.ptr(vals_cm), L.ptr(x),
                                          L.ptr(b), float(tolerance), int(bool(absolute_tolerance)), int(max_iterations),
                                          C.byref(it), C.byref(res)), self._h, ls=True)
        return (None if rc == L.NOT_POSITIVE_DEFINITE else x), int(it.value), float(res.value)

    def block_ordering(self):
        n = _check(lib.b200_ls_get_block_ordering(self._h, None), self._h, ls=True)
        p = np.zeros(n, np.int32)
        lib.b200_ls_get_block_ordering(self._h, L.ptr(p))
        return p

    def factor_nnz(self):
        return int(lib.b200_ls_get_factor_nnz(self._h))
