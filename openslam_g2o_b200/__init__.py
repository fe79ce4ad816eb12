"""openslam_g2o_b200 - B200-native solve path for g2o (hot path only).

The product is libg2o_b200.so (hand-written sm_100a CUDA behind the C-ABI of include/g2o_b200.h).
This package is the Python-side host mirror of the reference's SparseOptimizer / LinearSolver interface
used by the tests and the benchmark; it contains no numerical code of its own.
"""
from ._lib import (B200Error, EDGE_P2MC, EDGE_SE2, EDGE_SE2_XY, EDGE_SE3, EDGE_SE3_XYZ, EDGE_XYZ2UV, VERTEX_XY, GAUSS_NEWTON, LEVENBERG, LIB_PATH, VERTEX_CAM,
                   VERTEX_SE2, VERTEX_SE3, VERTEX_SE3_EXPMAP, VERTEX_XYZ, IterStats, lib)
from .optimizer import LinearSolverB200, SolverContext, SparseOptimizer, block_amd

__all__ = ["SparseOptimizer", "SolverContext", "LinearSolverB200", "block_amd", "B200Error", "IterStats", "lib",
           "LIB_PATH", "VERTEX_SE2", "VERTEX_SE3", "VERTEX_CAM", "VERTEX_XYZ", "VERTEX_SE3_EXPMAP", "EDGE_SE2", "EDGE_SE3", "EDGE_P2MC", "EDGE_XYZ2UV", "EDGE_SE2_XY", "EDGE_SE3_XYZ", "VERTEX_XY",
           "GAUSS_NEWTON", "LEVENBERG"]
