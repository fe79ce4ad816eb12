/*
 * oracle.h - C interface of the CPU ORACLE (TEST INFRASTRUCTURE ONLY).
 *
 * The oracle is a from-scratch, Eigen-free restatement of the reference's CPU hot path
 * (g2o SparseOptimizer -> OptimizationAlgorithm{GaussNewton,Levenberg} -> BlockSolver ->
 * LinearSolverCSparse) that calls the reference's OWN vendored CSparse (compiled in place into
 * oracle/_ref/libg2o_csparse_ref.so) for ordering, symbolic analysis and numeric Cholesky, the reference's own
 * compute_dq_dR (oracle/_ref/libg2o_slam3d_ref.so), and the reference's own robust kernels and SE2 class
 * (oracle/_ref/libg2o_ref_wrap.so, entered through oracle/ref_wrap.cpp).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library.  The product (openslam_g2o_b200/) never links, imports or executes it.
 *
 * Parity pinning: the reference ships no golden vectors for this path.  The oracle is
 * pinned by (1) the known-answer table of BASELINE.md section 2 (block-AMD permutation hash, nnz(L)), which was
 * produced by the reference's vendored CSparse and which this oracle reproduces through the same
 * library, and (2) the reference's two self-consistency tests restated in tests/ (analytic vs numeric
 * Jacobian at 1e-6), and (3) independent-solver checks in tests/ (converged chi2 of the SE3, SBACam and expmap families
 * against scipy.optimize.least_squares on the same cost; SE3Quat::exp against expm).  The CHOLMOD flavour of the path is unpinned by the reference (SuiteSparse is not
 * vendored); parity is claimed against the CSparse flavour ({gn,lm}_fix* solvers).
 */
#ifndef G2O_B200_ORACLE_H
#define G2O_B200_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct oracle_graph oracle_graph;

/* vertex / edge kinds (shared numbering with include/g2o_b200.h) */
enum { ORC_VERTEX_SE2 = 0, ORC_VERTEX_SE3 = 1, ORC_VERTEX_CAM = 2, ORC_VERTEX_XYZ = 3, ORC_VERTEX_SE3_EXPMAP = 4, ORC_VERTEX_XY = 5 };
/* SE2_XY = EdgeSE2PointXY (types/slam2d/edge_se2_pointxy.h: VertexSE2 -> VertexPointXY), SE3_XYZ = EdgeSE3PointXYZ
 * (types/slam3d/edge_se3_pointxyz.h: VertexSE3 -> VertexPointXYZ, stored as ORC_VERTEX_XYZ) */
enum { ORC_EDGE_SE2 = 0, ORC_EDGE_SE3 = 1, ORC_EDGE_P2MC = 2, ORC_EDGE_XYZ2UV = 3, ORC_EDGE_SE2_XY = 4, ORC_EDGE_SE3_XYZ = 5 };
enum { ORC_GN = 0, ORC_LM = 1 };

/* one record per outer iteration; mirrors the fields of G2OBatchStatistics (core/batch_stats.h:40-77) */
typedef struct oracle_iter_stats {
  int iteration;
  int levenberg_iterations;
  int result;          /* 1 OK, 2 Terminate, -1 Fail (core/optimization_algorithm.h SolverResult) */
  int reserved;
  double chi2;         /* activeRobustChi2 after the iteration */
  double lambda;       /* _currentLambda after the iteration (LM) */
  double time_residuals, time_quadratic_form, time_schur, time_symbolic, time_numeric,
         time_linear_solver, time_linear_solution, time_update, time_iteration;
} oracle_iter_stats;

oracle_graph* oracle_new(void);
void oracle_free(oracle_graph* g);

/* OptimizableGraph::load (core/optimizable_graph.cpp:356-569) for the configured tags */
int oracle_load(oracle_graph* g, const char* path);
/* programmatic construction; payload = the numbers that follow the ids on the .g2o line */
int oracle_add_vertex(oracle_graph* g, int kind, int id, const double* payload, int n);
int oracle_add_edge(oracle_graph* g, int kind, int id1, int id2, const double* payload, int n);
int oracle_add_vertices(oracle_graph* g, int kind, int n, const int* ids, const double* payload, int stride);
int oracle_add_edges(oracle_graph* g, int kind, int n, const int* id1, const int* id2, const double* payload, int stride);
int oracle_set_fixed(oracle_graph* g, int id, int fixed);
/* PARAMS_CAMERAPARAMETERS (types/sba/types_six_dof_expmap.h:45-80); must precede the XYZ2UV edges that name it.
 * XYZ2UV edge payload = paramId u v i00 i01 i11 (types_six_dof_expmap.cpp:241-256) */
int oracle_add_camera_parameters(oracle_graph* g, int id, double focal_length, double cx, double cy, double baseline);
/* PARAMS_SE3OFFSET id x y z qx qy qz qw (types/slam3d/parameter_se3_offset.cpp:47-56); must precede the SE3_XYZ edges that
 * name it.  SE3_XYZ edge payload = paramId x y z + upper triangle of the 3x3 information; SE2_XY: x y i00 i01 i11 */
int oracle_add_se3_offset(oracle_graph* g, int id, const double* xyz_qxyzw);

/* apps/g2o_cli/g2o.cpp:272-320: gauge fixing + marginalisation of the low-dimensional vertices.
 * returns the id of the vertex fixed as gauge, -1 if none was needed, -2 on error. */
int oracle_setup_cli(oracle_graph* g, int requires_marginalize);
/* SparseOptimizer::initializeOptimization (core/sparse_optimizer.cpp:199-267) */
int oracle_initialize(oracle_graph* g);
/* LinearSolverCSparse::setBlockOrdering (solver_csparse.cpp: fix* -> 1, var -> 0) */
void oracle_set_block_ordering(oracle_graph* g, int block_ordering);
/* the BlockSolver's linear solver: 0 LinearSolverCSparse (default), 1 LinearSolverPCG (solvers/pcg/linear_solver_pcg.hpp:79-160,
 * the `*_pcg*` solvers of solvers/pcg/solver_pcg.cpp) with setTolerance / setAbsoluteTolerance / setMaxIterations */
int oracle_set_linear_solver(oracle_graph* g, int kind, double tolerance, int absolute_tolerance, int max_iterations);
int oracle_pcg_iterations(oracle_graph* g);   /* iterations of the last PCG solve */
/* one robust kernel on every edge (apps/g2o_cli/g2o.cpp:322-336; core/robust_kernel_impl.cpp:65-126):
 * kind 0 none, 1 Huber, 2 PseudoHuber, 3 Cauchy, 4 Saturated, 5 DCS */
int oracle_set_robust_kernel(oracle_graph* g, int kind, double delta);
void oracle_robustify(int kind, double delta, double e2, double* rho3);   /* the reference's object code (oracle/_ref) */
void oracle_robustify_restated(int kind, double delta, double e2, double* rho3);   /* the restatement, for the cross-check */
/* restated SE2 algebra: what = 0 a*b, 1 a^-1, 2 EdgeSE2 error (a = v1, b = v2, z), 3 EdgeSE2PointXY error (a = pose, b = point, z) */
void oracle_se2_restated(int what, const double* a, const double* b, const double* z, double* out);
/* Edge::setRobustKernel on edge k (addEdge order) only */
int oracle_set_edge_robust_kernel(oracle_graph* g, int k, int kind, double delta);
/* Solver::computeMarginals -> LinearSolverCSparse::solvePattern -> MarginalCovarianceCholesky: blocks (rows[q], cols[q])
 * of Hpp^-1 (after oracle_build_system), column-major d x d each */
int oracle_compute_marginals(oracle_graph* g, int nblocks, const int* rows, const int* cols, double* out);  /* rho, rho', rho'' */

/* SparseOptimizer::optimize (core/sparse_optimizer.cpp:354-419); returns #iterations done (0 on Fail) */
int oracle_optimize(oracle_graph* g, int algorithm, int iterations, oracle_iter_stats* stats);

/* one OptimizationAlgorithmLevenberg::solve(iteration) of a running optimisation; returns SolverResult */
int oracle_lm_iteration(oracle_graph* g, int iteration, oracle_iter_stats* stats);
/* one GN (algorithm 0) / LM (1) iteration incl. OptimizationAlgorithm::init at iteration 0; stats->chi2 = chi2 after it */
int oracle_iteration(oracle_graph* g, int algorithm, int iteration, oracle_iter_stats* stats);

/* ---- step-wise access to the same objects (for fine-grained parity tests) ---- */
int oracle_algorithm_init(oracle_graph* g);             /* OptimizationAlgorithmWithHessian::init */
int oracle_build_structure(oracle_graph* g);            /* BlockSolver::buildStructure */
double oracle_compute_active_errors(oracle_graph* g);   /* computeActiveErrors + activeRobustChi2 */
int oracle_build_system(oracle_graph* g);               /* BlockSolver::buildSystem */
double oracle_lambda_init(oracle_graph* g);             /* Levenberg::computeLambdaInit */
int oracle_set_lambda(oracle_graph* g, double lambda, int backup);
int oracle_restore_diagonal(oracle_graph* g);
int oracle_solve(oracle_graph* g);                      /* BlockSolver::solve, 1 ok / 0 not PD */
int oracle_update(oracle_graph* g);                     /* SparseOptimizer::update(solver.x()) */
int oracle_push(oracle_graph* g);
int oracle_pop(oracle_graph* g);
int oracle_discard_top(oracle_graph* g);

/* dims[0..7] = numPoses, numLandmarks, sizePoses, sizeLandmarks, #activeEdges, #activeVertices,
 *              poseDim, landmarkDim */
int oracle_dims(oracle_graph* g, int* dims);
int oracle_get_b(oracle_graph* g, double* b);           /* length sizePoses+sizeLandmarks */
int oracle_get_x(oracle_graph* g, double* x);
int oracle_set_x(oracle_graph* g, const double* x);   /* overwrite solver.x() (numeric-Jacobian checks) */
int oracle_get_errors(oracle_graph* g, double* err);    /* active edges in order, D doubles each */
/* canonical estimate layouts: SE2 [x y th]; SE3 [R col-major 9, t 3]; CAM [t3 q(xyzw)4 fx fy cx cy b];
 * XYZ [x y z]; SE3_EXPMAP [t3 q(xyzw)4] (world -> camera).  returns #doubles written or -1 */
int oracle_get_estimate(oracle_graph* g, int id, double* out);
int oracle_vertex_count(oracle_graph* g);
/* all vertices ascending id: ids[], kind[], hessianIndex[], flags (1 fixed | 2 marginalized) */
int oracle_get_vertices(oracle_graph* g, int* ids, int* kinds, int* hidx, int* flags);
int oracle_edge_count(oracle_graph* g);                 /* all edges, file order */
/* canonical edge data: ids of both vertices + measurement + information, see oracle_get_estimate.
 * meas: SE2 [x y th]; SE3 [R9 t3] of Z; P2MC [u v].  info: full DxD col-major. */
int oracle_get_edge(oracle_graph* g, int k, int* kind, int* id1, int* id2, double* meas, double* info);

/* Hessian blocks. which: 0 Hpp, 1 Hll, 2 Hpl, 3 Hschur.  First call with rows==NULL returns #blocks.
 * blocks are listed column by column, ascending row; values column-major, rdim*cdim each. */
int oracle_get_blocks(oracle_graph* g, int which, int* rows, int* cols, double* values);
int oracle_get_bschur(oracle_graph* g, double* out);
/* symbolic results of LinearSolverCSparse::computeSymbolicDecomposition
 * (solvers/csparse/linear_solver_csparse.h:246-300): block permutation (block ordering) or scalar
 * permutation, and nnz(L). */
int oracle_get_block_perm(oracle_graph* g, int* perm);  /* returns length */
int64_t oracle_get_lnz(oracle_graph* g);

/* standalone ordering known-answer helper: runs the vendored cs_amd(1, pattern) on an upper-triangular
 * block pattern in CCS form (as fillBlockStructure emits it) */
int oracle_cs_amd(int n, const int* colptr, const int* rowidx, int* perm);
/* scalar cs_schol(1, A) lnz of the current linear system (for BASELINE.md "scalar-AMD" column) */
int64_t oracle_scalar_amd_lnz(oracle_graph* g);

#ifdef __cplusplus
}
#endif
#endif
