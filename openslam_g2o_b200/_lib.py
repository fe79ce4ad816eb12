"""ctypes binding of libg2o_b200.so (C-ABI declared in include/g2o_b200.h).

The library is the product; this module only declares signatures.  There is no Python or CPU fallback:
if the shared library is missing the import fails, and every compute call fails with
B200_ERR_NO_DEVICE when no CUDA device is present.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("G2O_B200_LIB", os.path.join(_HERE, "libg2o_b200.so"))  # override: instrumented debug builds

OK = 0
NOT_POSITIVE_DEFINITE = 1
ERR_INVALID, ERR_NO_DEVICE, ERR_CUDA, ERR_UNSUPPORTED, ERR_COLLECTIVE, ERR_EXCEPTION = -1, -2, -3, -4, -5, -6
VERTEX_SE2, VERTEX_SE3, VERTEX_CAM, VERTEX_XYZ, VERTEX_SE3_EXPMAP, VERTEX_XY = 0, 1, 2, 3, 4, 5
EDGE_SE2, EDGE_SE3, EDGE_P2MC, EDGE_XYZ2UV, EDGE_SE2_XY, EDGE_SE3_XYZ = 0, 1, 2, 3, 4, 5
NUM_VERTEX_KINDS, NUM_EDGE_KINDS = 6, 6
GAUSS_NEWTON, LEVENBERG = 0, 1
COMM_ID_BYTES = 128
RESULT_TERMINATE, RESULT_OK, RESULT_FAIL = 2, 1, -1   # values of IterStats.result (g2o's SolverResult)
SOLVE_FAIL = 3                                        # return value of b200_algorithm_solve for a failed solve

VERTEX_EST_LEN = {VERTEX_SE2: 3, VERTEX_SE3: 12, VERTEX_CAM: 12, VERTEX_XYZ: 3, VERTEX_SE3_EXPMAP: 12, VERTEX_XY: 2}
VERTEX_DIM = {VERTEX_SE2: 3, VERTEX_SE3: 6, VERTEX_CAM: 6, VERTEX_XYZ: 3, VERTEX_SE3_EXPMAP: 6, VERTEX_XY: 2}
EDGE_DIM = {EDGE_SE2: 3, EDGE_SE3: 6, EDGE_P2MC: 2, EDGE_XYZ2UV: 2, EDGE_SE2_XY: 2, EDGE_SE3_XYZ: 3}
EDGE_MEAS_LEN = {EDGE_SE2: 3, EDGE_SE3: 12, EDGE_P2MC: 2, EDGE_XYZ2UV: 2, EDGE_SE2_XY: 2, EDGE_SE3_XYZ: 3}


class IterStats(C.Structure):
    """POD twin of G2OBatchStatistics (reference core/batch_stats.h:40-77)."""
    _fields_ = [("iteration", C.c_int32), ("levenberg_iterations", C.c_int32), ("result", C.c_int32),
                ("reserved", C.c_int32), ("chi2", C.c_double), ("lambda_", C.c_double),
                ("time_residuals", C.c_double), ("time_quadratic_form", C.c_double), ("time_schur", C.c_double),
                ("time_symbolic", C.c_double), ("time_numeric", C.c_double), ("time_linear_solver", C.c_double),
                ("time_linear_solution", C.c_double), ("time_update", C.c_double), ("time_iteration", C.c_double)]


ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p)
TERMINATE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p)


class B200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("g2o_b200 error %d: %s" % (code, msg))
        self.code = code


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libg2o_b200.so is missing (%s). Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C openslam_g2o_b200/csrc`. There is no fallback path." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_double
    sig = {
        "b200_device_count": (i32, []),
        "b200_create": (i32, [i32, C.POINTER(vp)]),
        "b200_destroy": (None, [vp]),
        "b200_last_error": (C.c_char_p, [vp]),
        "b200_version": (C.c_char_p, []),
        "b200_set_vertices": (i32, [vp, i32, i32, vp, vp, vp]),
        "b200_set_edges": (i32, [vp, i32, i32, vp, vp, vp, vp]),
        "b200_set_sensor_offset": (i32, [vp, vp]),
        "b200_set_allreduce": (i32, [vp, ALLREDUCE_FN, vp, i32, i32]),
        "b200_add_schur_pattern": (i32, [vp, i32, vp, vp]),
        "b200_comm_unique_id": (i32, [vp, i32]),
        "b200_comm_init": (i32, [vp, vp, i32, i32]),
        "b200_comm_destroy": (i32, [vp]),
        "b200_comm_version": (i32, []),
        "b200_get_stream": (vp, [vp]),
        "b200_synchronize": (i32, [vp]),
        "b200_build_structure": (i32, [vp]),
        "b200_compute_active_errors": (i32, [vp, C.POINTER(dbl)]),
        "b200_build_system": (i32, [vp]),
        "b200_set_lambda": (i32, [vp, dbl, i32]),
        "b200_restore_diagonal": (i32, [vp]),
        "b200_solve": (i32, [vp]),
        "b200_update": (i32, [vp]),
        "b200_push": (i32, [vp]),
        "b200_pop": (i32, [vp]),
        "b200_discard_top": (i32, [vp]),
        "b200_optimize": (i32, [vp, i32, i32, vp]),
        "b200_algorithm_solve": (i32, [vp, i32, i32, vp]),
        "b200_set_lm_params": (i32, [vp, dbl, i32]),
        "b200_set_terminate": (i32, [vp, TERMINATE_FN, vp]),
        "b200_get_dims": (i32, [vp, vp]),
        "b200_get_x": (i32, [vp, vp]),
        "b200_get_b": (i32, [vp, vp]),
        "b200_get_estimates": (i32, [vp, i32, vp]),
        "b200_set_estimates": (i32, [vp, i32, vp]),
        "b200_get_hessian_diagonal": (i32, [vp, vp]),
        "b200_get_blocks": (i32, [vp, i32, vp, vp, vp]),
        "b200_get_bschur": (i32, [vp, vp]),
        "b200_get_block_ordering": (i32, [vp, vp]),
        "b200_get_factor_nnz": (i64, [vp]),
        "b200_get_factor_info": (i32, [vp, vp]),
        "b200_set_robust_kernel": (i32, [vp, i32, C.c_double]),
        "b200_set_edge_robust_kernels": (i32, [vp, i32, i32, vp, vp]),
        "b200_set_ordering": (i32, [vp, i32]),
        "b200_set_linear_solver": (i32, [vp, i32, C.c_double, i32, i32]),
        "b200_get_linear_solver_iterations": (i32, [vp]),
        "b200_compute_marginals": (i32, [vp, i32, vp, vp, vp]),
        "b200_get_launch_count": (i64, [vp]),
        "b200_debug_upload_digest": (C.c_uint64, [i32]),
        "b200_set_profiling": (i32, [vp, i32]),
        "b200_get_phase_time": (i32, [vp, i32, C.POINTER(dbl), C.POINTER(i64)]),
        "b200_ls_create": (i32, [i32, C.POINTER(vp)]),
        "b200_ls_destroy": (None, [vp]),
        "b200_ls_init": (i32, [vp]),
        "b200_ls_solve": (i32, [vp, i32, i32, vp, vp, vp, vp, vp]),
        "b200_ls_solve_pcg": (i32, [vp, i32, i32, vp, vp, vp, vp, vp, C.c_double, i32, i32, vp, vp]),
        "b200_ls_get_block_ordering": (i32, [vp, vp]),
        "b200_ls_get_factor_nnz": (i64, [vp]),
        "b200_ls_last_error": (C.c_char_p, [vp]),
        "b200_block_amd": (i32, [i32, vp, vp, vp]),
        "b200_graph_create": (i32, [C.POINTER(vp)]),
        "b200_graph_destroy": (None, [vp]),
        "b200_graph_load": (i32, [vp, C.c_char_p]),
        "b200_graph_add_vertex": (i32, [vp, i32, i32, vp, i32]),
        "b200_graph_add_edge": (i32, [vp, i32, i32, i32, vp, i32]),
        "b200_graph_add_vertices": (i32, [vp, i32, i32, vp, vp, i32]),
        "b200_graph_add_edges": (i32, [vp, i32, i32, vp, vp, vp, i32]),
        "b200_graph_set_fixed": (i32, [vp, i32, i32]),
        "b200_graph_set_edge_robust_kernel": (i32, [vp, i32, i32, C.c_double]),
        "b200_graph_add_camera_parameters": (i32, [vp, i32, dbl, dbl, dbl, dbl]),
        "b200_graph_add_se3_offset": (i32, [vp, i32, vp]),
        "b200_graph_setup_cli": (i32, [vp, i32]),
        "b200_graph_initialize": (i32, [vp]),
        "b200_graph_counts": (i32, [vp, vp, vp]),
        "b200_graph_upload": (i32, [vp, vp, i32, i32]),
        "b200_graph_download": (i32, [vp, vp]),
        "b200_graph_get_estimate": (i32, [vp, i32, vp]),
        "b200_graph_get_vertex_info": (i32, [vp, i32, vp]),
        "b200_graph_save": (i32, [vp, C.c_char_p]),
        "b200_graph_last_error": (C.c_char_p, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    lib._signatures = sig
    return lib


lib = _load()
EXPORTED_SYMBOLS = sorted(lib._signatures.keys())


def ptr(a):
    """pointer to a C-contiguous numpy array (or None)"""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def as_i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)
