"""ctypes binding of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY: imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs - never by the product."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_PATH = os.path.join(ROOT, "oracle", "liboracle.so")
GN, LM = 0, 1


class OracleStats(C.Structure):
    _fields_ = [("iteration", C.c_int), ("levenberg_iterations", C.c_int), ("result", C.c_int), ("reserved", C.c_int),
                ("chi2", C.c_double), ("lambda_", C.c_double), ("time_residuals", C.c_double),
                ("time_quadratic_form", C.c_double), ("time_schur", C.c_double), ("time_symbolic", C.c_double),
                ("time_numeric", C.c_double), ("time_linear_solver", C.c_double), ("time_linear_solution", C.c_double),
                ("time_update", C.c_double), ("time_iteration", C.c_double)]


_lib = None


def oracle_lib():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_PATH):
            raise FileNotFoundError("oracle not built: run `make -C oracle` (needs /root/reference) - " + ORACLE_PATH)
        L = C.CDLL(ORACLE_PATH)
        L.oracle_new.restype = C.c_void_p
        L.oracle_get_lnz.restype = C.c_int64
        L.oracle_scalar_amd_lnz.restype = C.c_int64
        for f in ("oracle_compute_active_errors", "oracle_lambda_init"):
            getattr(L, f).restype = C.c_double
        L.oracle_set_lambda.argtypes = [C.c_void_p, C.c_double, C.c_int]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def fnv1a64(perm):
    """FNV-1a-64 over 32-bit entries"""
    h = 1469598103934665603
    for p in perm:
        h ^= int(p) & 0xFFFFFFFF
        h = (h * 1099511628211) % (1 << 64)
    return "%016x" % h


EST_LEN = {0: 3, 1: 12, 2: 12, 3: 3, 4: 7, 5: 2}
MEAS_LEN = {0: 3, 1: 12, 2: 2, 3: 2, 4: 2, 5: 3}
EDGE_DIM = {0: 3, 1: 6, 2: 2, 3: 2, 4: 2, 5: 3}


class Oracle:
    """thin object wrapper with the same method names as openslam_g2o_b200.SparseOptimizer where they overlap"""

    def __init__(self):
        self.L = oracle_lib()
        self.g = C.c_void_p(self.L.oracle_new())

    def __del__(self):
        try:
            self.L.oracle_free(self.g)
        except Exception:
            pass

    def load(self, path):
        return self.L.oracle_load(self.g, str(path).encode()) == 0

    def add_vertices(self, kind, ids, payload):
        ids = np.ascontiguousarray(ids, np.int32)
        payload = np.ascontiguousarray(payload, np.float64)
        assert self.L.oracle_add_vertices(self.g, kind, len(ids), _p(ids), _p(payload), payload.shape[1]) == 0

    def add_edges(self, kind, id1, id2, payload):
        id1 = np.ascontiguousarray(id1, np.int32)
        id2 = np.ascontiguousarray(id2, np.int32)
        payload = np.ascontiguousarray(payload, np.float64)
        assert self.L.oracle_add_edges(self.g, kind, len(id1), _p(id1), _p(id2), _p(payload), payload.shape[1]) == 0

    def set_fixed(self, vid, fixed=True):
        assert self.L.oracle_set_fixed(self.g, vid, int(fixed)) == 0

    def add_camera_parameters(self, pid, focal_length, cx, cy, baseline):
        self.L.oracle_add_camera_parameters.argtypes = [C.c_void_p, C.c_int] + [C.c_double] * 4
        assert self.L.oracle_add_camera_parameters(self.g, int(pid), focal_length, cx, cy, baseline) == 0

    def add_se3_offset(self, pid, xyz_qxyzw):
        o = np.ascontiguousarray(xyz_qxyzw, np.float64)
        assert self.L.oracle_add_se3_offset(self.g, int(pid), _p(o)) == 0

    def setup_cli(self, requires_marginalize=True):
        return self.L.oracle_setup_cli(self.g, int(requires_marginalize))

    def initialize_optimization(self):
        return self.L.oracle_initialize(self.g) == 0

    def compute_marginals(self, pairs):
        rows = np.ascontiguousarray([p[0] for p in pairs], dtype=np.int32)
        cols = np.ascontiguousarray([p[1] for p in pairs], dtype=np.int32)
        d = self.dims()["poseDim"]
        out = np.zeros((len(pairs), d, d))
        self.L.oracle_compute_marginals.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        if self.L.oracle_compute_marginals(self.g, len(pairs), _p(rows), _p(cols), _p(out)) != 0:
            return None
        return np.ascontiguousarray(np.transpose(out, (0, 2, 1)))

    def set_robust_kernel(self, name, width=1.0):
        kinds = {"none": 0, "Huber": 1, "PseudoHuber": 2, "Cauchy": 3, "Saturated": 4, "DCS": 5}
        self.L.oracle_set_robust_kernel.argtypes = [C.c_void_p, C.c_int, C.c_double]
        assert self.L.oracle_set_robust_kernel(self.g, kinds[name], float(width)) == 0

    def set_linear_solver(self, kind, tolerance=1e-6, absolute_tolerance=True, max_iterations=-1):
        self.L.oracle_set_linear_solver.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_int]
        assert self.L.oracle_set_linear_solver(self.g, {"cholesky": 0, "pcg": 1}[kind], float(tolerance), int(bool(absolute_tolerance)),
                                               int(max_iterations)) == 0

    def pcg_iterations(self):
        return int(self.L.oracle_pcg_iterations(self.g))

    def set_edge_robust_kernel(self, edge_indices, name, width=1.0):
        kinds = {"none": 0, "Huber": 1, "PseudoHuber": 2, "Cauchy": 3, "Saturated": 4, "DCS": 5}
        self.L.oracle_set_edge_robust_kernel.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double]
        for k in edge_indices:
            assert self.L.oracle_set_edge_robust_kernel(self.g, int(k), kinds[name], float(width)) == 0

    def set_block_ordering(self, on):
        self.L.oracle_set_block_ordering(self.g, int(on))

    def optimize(self, algorithm, iterations):
        st = (OracleStats * max(iterations, 1))()
        n = self.L.oracle_optimize(self.g, algorithm, iterations, st)
        return n, list(st)[:iterations]

    # step-wise
    def algorithm_init(self): return self.L.oracle_algorithm_init(self.g) == 0
    def build_structure(self): return self.L.oracle_build_structure(self.g) == 0
    def compute_active_errors(self): return self.L.oracle_compute_active_errors(self.g)
    def build_system(self): return self.L.oracle_build_system(self.g) == 0
    def lambda_init(self): return self.L.oracle_lambda_init(self.g)
    def set_lambda(self, lam, backup=True): self.L.oracle_set_lambda(self.g, float(lam), int(backup))
    def restore_diagonal(self): self.L.oracle_restore_diagonal(self.g)
    def solve(self): return self.L.oracle_solve(self.g) == 1
    def update(self): self.L.oracle_update(self.g)
    def push(self): self.L.oracle_push(self.g)
    def pop(self): self.L.oracle_pop(self.g)
    def discard_top(self): self.L.oracle_discard_top(self.g)

    def dims(self):
        d = (C.c_int * 8)()
        self.L.oracle_dims(self.g, d)
        return dict(zip(("numPoses", "numLandmarks", "sizePoses", "sizeLandmarks", "numEdges", "numVertices", "poseDim",
                         "landmarkDim"), list(d)))

    def _vec(self, fn):
        d = self.dims()
        out = np.zeros(d["sizePoses"] + d["sizeLandmarks"])
        fn(self.g, _p(out))
        return out

    def b(self): return self._vec(self.L.oracle_get_b)
    def x(self): return self._vec(self.L.oracle_get_x)

    def set_x(self, x):
        x = np.ascontiguousarray(x, np.float64)
        self.L.oracle_set_x(self.g, _p(x))

    def bschur(self):
        out = np.zeros(self.dims()["sizePoses"])
        self.L.oracle_get_bschur(self.g, _p(out))
        return out

    def vertices(self):
        n = self.L.oracle_vertex_count(self.g)
        ids, kinds, hidx, flags = (np.zeros(n, np.int32) for _ in range(4))
        self.L.oracle_get_vertices(self.g, _p(ids), _p(kinds), _p(hidx), _p(flags))
        return ids, kinds, hidx, flags

    def vertex_estimate(self, vid):
        out = np.zeros(12)
        n = self.L.oracle_get_estimate(self.g, int(vid), _p(out))
        return out[:n].copy()

    def estimates(self):
        """dict id -> canonical estimate"""
        ids, _, _, _ = self.vertices()
        return {int(i): self.vertex_estimate(i) for i in ids}

    def edges(self):
        n = self.L.oracle_edge_count(self.g)
        out = []
        kind, a, b = C.c_int(), C.c_int(), C.c_int()
        meas, info = np.zeros(12), np.zeros(36)
        for k in range(n):
            self.L.oracle_get_edge(self.g, k, C.byref(kind), C.byref(a), C.byref(b), _p(meas), _p(info))
            D = EDGE_DIM[kind.value]
            out.append((kind.value, a.value, b.value, meas[:MEAS_LEN[kind.value]].copy(), info[:D * D].copy()))
        return out

    def blocks(self, which):
        n = self.L.oracle_get_blocks(self.g, which, None, None, None)
        d = self.dims()
        rd, cd = {0: (d["poseDim"],) * 2, 1: (3, 3), 2: (d["poseDim"], 3), 3: (d["poseDim"],) * 2}[which]
        rows, cols = np.zeros(n, np.int32), np.zeros(n, np.int32)
        vals = np.zeros((n, cd, rd))
        self.L.oracle_get_blocks(self.g, which, _p(rows), _p(cols), _p(vals))
        return rows, cols, np.transpose(vals, (0, 2, 1))

    def block_perm(self):
        n = self.L.oracle_get_block_perm(self.g, None)
        p = np.zeros(n, np.int32)
        self.L.oracle_get_block_perm(self.g, _p(p))
        return p

    def lnz(self): return int(self.L.oracle_get_lnz(self.g))
    def scalar_amd_lnz(self): return int(self.L.oracle_scalar_amd_lnz(self.g))


def cs_amd(colptr, rowidx):
    colptr = np.ascontiguousarray(colptr, np.int32)
    rowidx = np.ascontiguousarray(rowidx, np.int32)
    n = len(colptr) - 1
    perm = np.zeros(n, np.int32)
    assert oracle_lib().oracle_cs_amd(n, _p(colptr), _p(rowidx), _p(perm)) == 0
    return perm
