/*
 * g2o_b200.h - C-ABI of the B200-native g2o solve path (libg2o_b200.so).
 *
 * Plain C: opaque handles, caller-owned buffers, int status returns (0 = ok, >0 = numerical outcome,
 * <0 = error; b200_last_error() gives the text).  No C++/torch types cross this boundary.
 * Everything behind it runs as hand-written sm_100a CUDA; there is NO CPU fallback: every compute entry
 * point fails with B200_ERR_NO_DEVICE when no CUDA device is usable.
 *
 * Each entry point names the reference interface it replaces (paths relative to /root/reference/g2o).
 *
 * Data layouts (all double = IEEE f64, all indices 32-bit int, matrices column-major like Eigen):
 *   vertex estimates  SE2: [x y theta]                              (types/slam2d/vertex_se2.h:64-73)
 *                     SE3: [R(3x3 col-major) t(3)] = Isometry3d     (types/slam3d/vertex_se3.h, state is R|t)
 *                     CAM: [t(3) q(x y z w) fx fy cx cy baseline]   (types/sba/sbacam.h:60-98)
 *                     XYZ: [x y z]                                  (types/sba/types_sba.h:136-156)
 *                     SE3_EXPMAP: [t(3) q(x y z w) f f cx cy baseline]  world->camera SE3Quat (types/sba/
 *                          types_six_dof_expmap.h:87-105) + the CameraParameters its edges name (:45-80)
 *                     XY:  [x y]                                    (types/slam2d/vertex_point_xy.h)
 *                     (XYZ also stands for VertexPointXYZ, types/slam3d/vertex_pointxyz.h: same state, same update)
 *   edge measurement  SE2: [x y theta] of Z; SE3: [R t] of Z (12); P2MC, XYZ2UV: [u v]; SE2_XY: [x y]; SE3_XYZ: [x y z]
 *   edge information  full D x D column-major (D = 3, 6, 2, 2, 2, 3)
 */
#ifndef G2O_B200_H
#define G2O_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define B200_OK 0
#define B200_NOT_POSITIVE_DEFINITE 1   /* LinearSolver::solve returned false (linear_solver.h:59) */
#define B200_ERR_INVALID (-1)
#define B200_ERR_NO_DEVICE (-2)
#define B200_ERR_CUDA (-3)
#define B200_ERR_UNSUPPORTED (-4)      /* graph uses types / combinations outside the families listed below: no CPU fallback */
#define B200_ERR_COLLECTIVE (-5)
#define B200_ERR_EXCEPTION (-6)        /* a C++ exception other than a CUDA failure (std::bad_alloc, ...) was caught at the ABI */

enum { B200_VERTEX_SE2 = 0, B200_VERTEX_SE3 = 1, B200_VERTEX_CAM = 2, B200_VERTEX_XYZ = 3, B200_VERTEX_SE3_EXPMAP = 4,
       B200_VERTEX_XY = 5 };
/* XYZ2UV = EdgeProjectXYZ2UV (types/sba/types_six_dof_expmap.h:133-155), the monocular edge of ba_demo / SE3 expmap BA.
 * Landmark SLAM: SE2_XY = EdgeSE2PointXY (types/slam2d/edge_se2_pointxy.h:44-51, .cpp:67-95),
 * SE3_XYZ = EdgeSE3PointXYZ (types/slam3d/edge_se3_pointxyz.cpp:98-135) with its ParameterSE3Offset
 * (b200_set_sensor_offset).  They go with the pose-pose edges of their pose kind in ONE context, landmarks NOT
 * marginalized: the reference's variable-block-size path (BlockSolverX, `gn_var` / `lm_var`, solvers/csparse/
 * solver_csparse.cpp:53-55, the set-up of examples/tutorial_slam2d).  Here the landmark blocks are padded to the pose
 * dimension (2 -> 3, 3 -> 6; the padding rows / columns are decoupled unit-diagonal unknowns that stay 0), so the
 * block pattern - and with it the block AMD ordering - is exactly the reference's, and x / b / Hpp blocks / marginals
 * are reported in the padded layout (poseDim entries per vertex). */
enum { B200_EDGE_SE2 = 0, B200_EDGE_SE3 = 1, B200_EDGE_P2MC = 2, B200_EDGE_XYZ2UV = 3, B200_EDGE_SE2_XY = 4,
       B200_EDGE_SE3_XYZ = 5 };
#define B200_NUM_VERTEX_KINDS 6
#define B200_NUM_EDGE_KINDS 6
enum { B200_GAUSS_NEWTON = 0, B200_LEVENBERG = 1 };
/* OptimizationAlgorithm::SolverResult (core/optimization_algorithm.h:49): the value stored in b200_iter_stats.result.
 * The RETURN value of b200_algorithm_solve never uses -1 for it (that is B200_ERR_INVALID): a failed linear solve
 * (GN: not positive definite) is returned as B200_SOLVE_FAIL, hard errors are < 0. */
enum { B200_RESULT_TERMINATE = 2, B200_RESULT_OK = 1, B200_RESULT_FAIL = -1 };
#define B200_SOLVE_FAIL 3

typedef struct b200_ctx b200_ctx;

/* POD twin of G2OBatchStatistics (core/batch_stats.h:40-77); times in seconds.  time_iteration (monotonic clock) and
 * time_symbolic (iteration 0) are always filled.  The per-phase fields (time_residuals, time_quadratic_form,
 * time_schur, time_numeric, time_linear_solver, time_linear_solution, time_update: CUDA-event intervals on the solver
 * stream, summed over the trials of the iteration) are filled while b200_set_profiling(ctx, 1) is on - `g2o -stats`
 * mode: the iteration then runs as individual launches instead of CUDA-graph replays - and are 0 otherwise. */
typedef struct b200_iter_stats {
  int32_t iteration;
  int32_t levenberg_iterations;
  int32_t result;
  int32_t reserved;
  double chi2;
  double lambda;
  double time_residuals, time_quadratic_form, time_schur, time_symbolic, time_numeric,
         time_linear_solver, time_linear_solution, time_update, time_iteration;
} b200_iter_stats;

/* optional collective for landmark-sharded BA: all-reduce (op 0 = sum, 1 = max) `count` doubles living at DEVICE
 * pointer `dev_ptr`, ordered on CUDA stream `stream` (a cudaStream_t).  The host side supplies it (NCCL through
 * torch.distributed in the Python harness, ncclAllReduce(ncclDouble, ncclSum|ncclMax) in a C++ host).
 * return 0 on success. */
typedef int (*b200_allreduce_fn)(void* dev_ptr, int64_t count, int op, void* stream, void* user);

/* ------------------------------------------------------------------ lifetime */
int b200_device_count(void);
int b200_create(int device, b200_ctx** out);
void b200_destroy(b200_ctx* ctx);
const char* b200_last_error(const b200_ctx* ctx);   /* ctx may be NULL: last creation error */
const char* b200_version(void);

/* ------------------------------------------------------------------ graph ingest
 * What a Solver adapter extracts in buildStructure() from SparseOptimizer::indexMapping()/activeEdges()
 * (core/block_solver.hpp:142-295).  One pose kind and at most one landmark kind (XYZ) per context.
 * hessian_index: g2o's v->hessianIndex() (-1 = fixed);  marginalized: v->marginalized() (may be NULL). */
int b200_set_vertices(b200_ctx* ctx, int kind, int n, const double* estimates,
                      const int32_t* hessian_index, const uint8_t* marginalized);
/* vi/vj index into the vertex array of the kind the edge type expects
 * (SE2: SE2,SE2; SE3: SE3,SE3; P2MC: vi = XYZ point, vj = CAM; XYZ2UV: vi = XYZ point, vj = SE3_EXPMAP;
 * SE2_XY: vi = SE2 pose, vj = XY point; SE3_XYZ: vi = SE3 pose, vj = XYZ point).
 * Order = active edge order (internalId).  A context holds one pose-pose / projection edge set and, for landmark
 * SLAM, one pose-landmark edge set (SE2_XY with SE2, SE3_XYZ with SE3): a call replaces the set of its own class;
 * n = 0 removes it. */
int b200_set_edges(b200_ctx* ctx, int kind, int n, const int32_t* vi, const int32_t* vj,
                   const double* measurement, const double* information);
/* ParameterSE3Offset of the SE3_XYZ edges (types/slam3d/parameter_se3_offset.h: sensor pose on the robot), one for
 * the whole context: isometry [R(3x3 col-major) t(3)].  Default: identity. */
int b200_set_sensor_offset(b200_ctx* ctx, const double* isometry12);
/* landmark sharding: this context owns the landmarks / edges it was given; cameras are
 * replicated.  After the local Schur reduction [Hschur | bschur | scalars] is all-reduced through fn. */
int b200_set_allreduce(b200_ctx* ctx, b200_allreduce_fn fn, void* user, int rank, int world_size);
/* Native communicator for the sharded path (preferred over the callback): NCCL, bound at run time (dlopen of
 * libnccl.so.2), one rank per context / GPU.  Rank 0 calls b200_comm_unique_id and ships the B200_COMM_ID_BYTES bytes
 * to the other ranks over any out-of-band channel (MPI, a file, torch.distributed's store); every rank then calls
 * b200_comm_init (collective: ncclCommInitRank on the context's device).  From then on each LM trial issues exactly
 * two ncclAllReduce(ncclDouble, ncclSum) on the context's stream, captured in its CUDA graph: one over
 * [Hschur | bschur | chi2 before the trial] and one over 2 doubles after the update (chi2, LM scale). */
#define B200_COMM_ID_BYTES 128
int b200_comm_unique_id(void* out, int capacity);
int b200_comm_init(b200_ctx* ctx, const void* unique_id, int rank, int world_size);
int b200_comm_destroy(b200_ctx* ctx);
int b200_comm_version(void);   /* NCCL_VERSION_CODE of the bound library, 0 if none can be loaded */
/* sharded BA: blocks (rows[i] <= cols[i]) that OTHER shards contribute to the reduced camera matrix, so that
 * every rank builds the identical Hschur pattern (core/block_solver.hpp:262-288 over the whole graph).
 * Call AFTER b200_set_vertices (which forgets the keys of the previous graph) and before b200_build_structure;
 * indices are Hessian block indices of free poses (checked against numPoses at b200_build_structure). */
int b200_add_schur_pattern(b200_ctx* ctx, int n, const int32_t* rows, const int32_t* cols);
/* the CUDA stream (cudaStream_t) all work of this context is ordered on; b200_synchronize waits for it */
void* b200_get_stream(b200_ctx* ctx);
int b200_synchronize(b200_ctx* ctx);

/* ------------------------------------------------------------------ Level 2: g2o::Solver (core/solver.h:44-149) */
/* Solver::buildStructure (core/block_solver.hpp:142-295) + LinearSolver symbolic phase
 * (solvers/csparse/linear_solver_csparse.h:246-300): block pattern, block AMD, supernodal analysis. */
int b200_build_structure(b200_ctx* ctx);
/* SparseOptimizer::computeActiveErrors + activeRobustChi2 (core/sparse_optimizer.cpp:61-114) */
int b200_compute_active_errors(b200_ctx* ctx, double* chi2);
/* Solver::buildSystem (core/block_solver.hpp:501-560) */
int b200_build_system(b200_ctx* ctx);
/* Solver::setLambda / restoreDiagonal (core/block_solver.hpp:563-604) */
int b200_set_lambda(b200_ctx* ctx, double lambda, int backup);
int b200_restore_diagonal(b200_ctx* ctx);
/* Solver::solve (core/block_solver.hpp:354-486): [Schur] + sparse Cholesky + back-substitution.
 * returns B200_OK or B200_NOT_POSITIVE_DEFINITE. */
int b200_solve(b200_ctx* ctx);
/* SparseOptimizer::update(solver->x()) (core/sparse_optimizer.cpp:421-434) */
int b200_update(b200_ctx* ctx);
/* SparseOptimizer::push / pop / discardTop (core/sparse_optimizer.cpp:515-553, 599-612) */
int b200_push(b200_ctx* ctx);
int b200_pop(b200_ctx* ctx);
int b200_discard_top(b200_ctx* ctx);

/* ------------------------------------------------------------------ Level 3: g2o::OptimizationAlgorithm
 * SparseOptimizer::optimize loop over OptimizationAlgorithm{GaussNewton,Levenberg}::solve
 * (core/sparse_optimizer.cpp:354-419, core/optimization_algorithm_levenberg.cpp:57-147,
 * core/optimization_algorithm_gauss_newton.cpp:50-93), whole iteration device-resident.
 * stats may be NULL, else room for max_iterations records.  returns #iterations done (0 on Fail), <0 error */
int b200_optimize(b200_ctx* ctx, int algorithm, int max_iterations, b200_iter_stats* stats);
/* SparseOptimizer::terminate() / setForceStopFlag (core/sparse_optimizer.h:189, apps/g2o_cli/g2o.cpp:89-99,552): the
 * host's stop request, polled where the reference polls it - between the trials of one LM iteration
 * (core/optimization_algorithm_levenberg.cpp:142) and between the iterations of b200_optimize
 * (core/sparse_optimizer.cpp:376).  fn returns non-zero to stop; NULL removes the hook. */
typedef int (*b200_terminate_fn)(void* user);
int b200_set_terminate(b200_ctx* ctx, b200_terminate_fn fn, void* user);
/* one OptimizationAlgorithm::solve(iteration).  returns B200_RESULT_OK (1), B200_RESULT_TERMINATE (2) or
 * B200_SOLVE_FAIL (3; stats->result = B200_RESULT_FAIL = -1 as in the reference's enum); < 0: hard error (nothing
 * was solved - missing structure, CUDA failure, out of memory), never an ordinary LM/GN outcome */
int b200_algorithm_solve(b200_ctx* ctx, int algorithm, int iteration, b200_iter_stats* stats);
/* fill-reducing ordering of the (reduced) pose system.  nd_levels = 0 (default): block AMD, bit-exact with the
 * reference's cs_amd(1, .) on the block pattern (solvers/csparse/linear_solver_csparse.h:252-294).  nd_levels = k > 0:
 * nested dissection with 2^k parts on top of it (separators from breadth-first level structures, AMD inside the parts):
 * an ordering for PARALLELISM - band-like reduced camera systems, whose AMD elimination tree is one long chain,
 * become 2^k independent subtrees.  Same solution (to rounding), different elimination order; call before
 * b200_build_structure.  b200_get_block_ordering reports the ordering in use. */
int b200_set_ordering(b200_ctx* ctx, int nd_levels);
/* linear solver of the (reduced) pose system inside Solver::solve: the supernodal Cholesky (default; LinearSolverCSparse /
 * LinearSolverCholmod, `*_fix*`, `*_var`) or the block-Jacobi preconditioned conjugate gradients of LinearSolverPCG
 * (solvers/pcg/linear_solver_pcg.hpp:79-160; solver names `gn_pcg`, `lm_pcg`, `*_pcg3_2`, `*_pcg6_3`,
 * solvers/pcg/solver_pcg.cpp) with its setTolerance (1e-6) / setAbsoluteTolerance (true) / setMaxIterations (-1: the number
 * of rows).  Call before b200_build_structure.  PCG trials run as plain launches (the stopping rule reports to the host
 * every 64 iterations), the Cholesky trials as CUDA-graph replays. */
enum { B200_LINEAR_SOLVER_CHOLESKY = 0, B200_LINEAR_SOLVER_PCG = 1 };
int b200_set_linear_solver(b200_ctx* ctx, int kind, double tolerance, int absolute_tolerance, int max_iterations);
/* G2OBatchStatistics::iterationsLinearSolver (core/batch_stats.h:61): CG iterations of the last PCG solve (0 with the Cholesky) */
int b200_get_linear_solver_iterations(b200_ctx* ctx);
/* robust kernel applied to every edge, like `g2o -robustKernel NAME -robustKernelWidth W`
 * (apps/g2o_cli/g2o.cpp:322-336; kernels: core/robust_kernel_impl.cpp:65-126; use sites:
 * core/base_binary_edge.hpp:91-113 first-order weight rho' on information and omega_r,
 * core/sparse_optimizer.cpp:100-114 chi2 = sum of rho).  delta = kernel width (DCS: phi).  Default: none. */
enum { B200_ROBUST_NONE = 0, B200_ROBUST_HUBER = 1, B200_ROBUST_PSEUDO_HUBER = 2, B200_ROBUST_CAUCHY = 3,
       B200_ROBUST_SATURATED = 4, B200_ROBUST_DCS = 5 };
int b200_set_robust_kernel(b200_ctx* ctx, int kind, double delta);
/* per-edge robust kernels - Edge::setRobustKernel on individual edges (core/optimizable_graph.h:446-450), e.g. a kernel on
 * loop closures / sightings only: kinds[i], deltas[i] for edge i of the set given by b200_set_edges(edge_kind, ...) (same
 * order, n = its size; B200_ROBUST_NONE = no kernel on that edge).  Call after b200_set_edges; replaces the uniform kernel
 * for that set until the next b200_set_edges / b200_set_robust_kernel; kinds == NULL returns to the uniform one. */
int b200_set_edge_robust_kernels(b200_ctx* ctx, int edge_kind, int n, const uint8_t* kinds, const double* deltas);
/* Solver::computeMarginals (core/solver.h:100-106, core/block_solver.hpp:490-499, LinearSolver::solvePattern
 * solvers/csparse/linear_solver_csparse.h:190-225, core/marginal_covariance_cholesky.cpp): blocks (rows[q], cols[q])
 * of the inverse of the current Hpp (call after b200_build_system; lambda is not added), written to
 * out + q*d*d column-major.  Pose graphs only (B200_ERR_UNSUPPORTED with a Schur complement).
 * returns B200_OK / B200_NOT_POSITIVE_DEFINITE / <0 */
int b200_compute_marginals(b200_ctx* ctx, int nblocks, const int32_t* rows, const int32_t* cols, double* out);
/* LM properties (core/optimization_algorithm_levenberg.cpp:43-49) */
int b200_set_lm_params(b200_ctx* ctx, double user_lambda_init, int max_trials_after_failure);

/* ------------------------------------------------------------------ read-back (host mirrors the adapter needs) */
/* dims[8] = numPoses numLandmarks sizePoses sizeLandmarks numEdges numVertices poseDim landmarkDim */
int b200_get_dims(b200_ctx* ctx, int32_t* dims);
int b200_get_x(b200_ctx* ctx, double* x);           /* Solver::x(), length sizePoses+sizeLandmarks */
int b200_get_b(b200_ctx* ctx, double* b);           /* Solver::b() */
int b200_get_estimates(b200_ctx* ctx, int kind, double* estimates);  /* same layout/order as set_vertices */
/* overwrite the device-resident estimates (same layout/order as set_vertices): what a Level-3 adapter does when the
 * host graph changed between iterations (OptimizableGraph::Vertex::setEstimateData, core/optimizable_graph.h:191) */
int b200_set_estimates(b200_ctx* ctx, int kind, const double* estimates);
int b200_get_hessian_diagonal(b200_ctx* ctx, double* diag);          /* v->hessian(j,j), index order */
/* which: 0 Hpp 1 Hll 2 Hpl 3 Hschur; call with rows==NULL to get the block count.  Blocks are listed
 * column by column, ascending block row (SparseBlockMatrix order), values column-major. */
int b200_get_blocks(b200_ctx* ctx, int which, int32_t* rows, int32_t* cols, double* values);
int b200_get_bschur(b200_ctx* ctx, double* out);
/* fill-reducing block ordering actually used (bit-exact twin of cs_amd(1, blockPattern),
 * EXTERNAL/csparse/cs_amd.c:18-364 as called from linear_solver_csparse.h:268) and nnz(L) of the scalar
 * factor it implies (css::lnz, linear_solver_csparse.h:292-293) */
int b200_get_block_ordering(b200_ctx* ctx, int32_t* perm);   /* returns #blocks */
int64_t b200_get_factor_nnz(b200_ctx* ctx);
/* schedule facts, out[0..23] (out[16] = 1 when the update plan uses the wide 96 x 72 tiles of the FP64 tensor path,
 * out[17] = update work items, out[18] = flops their rectangular products execute, out[19] = scratch slots of split tiles,
 * out[20..23] reserved; out[12] = flops of one factorisation of the stored structure, out[13] = links of the tail
 * chain - supernodes factored by the one-CTA chain kernel -, out[14] = their flops, out[15] reserved): #supernodes, #tasks, #levels, max panel rows, max panel cols, stored factor doubles,
 * #dataflow tasks of the factorisation kernel; Schur plan (0 without landmarks): #landmark ranges, #segments
 * (partial sums), #contributions (block products), #Hpl slots, shared-memory bytes of a range CTA */
int b200_get_factor_info(b200_ctx* ctx, int64_t* out);
/* kernels launched by this context since creation (bench "gpu_launches") */
int64_t b200_get_launch_count(b200_ctx* ctx);
/* seconds of the dominant kernels accumulated with CUDA events when profiling is on */
int b200_set_profiling(b200_ctx* ctx, int on);
/* kernel-group id: 0 errors+chi2, 1 linearize (edges / per-landmark), 2 schur landmark ranges, 3 cholesky factor,
 * 4 triangular solves, 5 oplus update, 6 landmark back-substitution, 7 linearize per-camera, 8 ordered gather,
 * 9 landmark inverses, 10 LM scale, 11 collective, 14 schur finish; inside the Cholesky: 12 scatter of the input
 * blocks, 13 the dataflow factorisation kernel (updates + panels + forward solve), 19 the dataflow backward sweep.
 * Profiling turns CUDA-graph replay off.  returns accumulated seconds and #occurrences */
int b200_get_phase_time(b200_ctx* ctx, int phase, double* seconds, int64_t* count);

/* ------------------------------------------------------------------ Level 1: g2o::LinearSolver<MatrixType>
 * (core/linear_solver.h:40-81): solve A x = b for an upper-triangular block-CCS SparseBlockMatrix
 * (core/sparse_block_matrix.h:61-220) with uniform block size.  Pattern is analysed at the first solve after
 * b200_ls_init(), like LinearSolverCSparse (linear_solver_csparse.h:106-142). */
typedef struct b200_linear_solver b200_linear_solver;
int b200_ls_create(int device, b200_linear_solver** out);
void b200_ls_destroy(b200_linear_solver* ls);
int b200_ls_init(b200_linear_solver* ls);            /* LinearSolver::init(): drop the symbolic factor */
/* colptr[nblocks+1], rowidx[]: upper block pattern incl. diagonal, ascending rows per column
 * (SparseBlockMatrix::fillBlockStructure, core/sparse_block_matrix.hpp:519-545).
 * values: blocks in the same order, block_dim^2 doubles each, column-major.  x,b: HOST buffers. */
int b200_ls_solve(b200_linear_solver* ls, int nblocks, int block_dim, const int32_t* colptr,
                  const int32_t* rowidx, const double* values, double* x, const double* b);
/* LinearSolverPCG<MatrixType>::solve (solvers/pcg/linear_solver_pcg.hpp:79-160, linear_solver_pcg.h:47-98): conjugate
 * gradients with the block-Jacobi preconditioner on the same upper-triangular block CCS matrix.  tolerance /
 * absolute_tolerance / max_iterations are the reference's setTolerance (1e-6), setAbsoluteTolerance (true) and
 * setMaxIterations (-1: the number of rows); the absolute residual of one solve carries over to the next until
 * b200_ls_init(), like _residual.  *iterations = G2OBatchStatistics::iterationsLinearSolver, *residual = _residual.
 * Returns B200_OK like the reference's unconditional `return true`, B200_NOT_POSITIVE_DEFINITE when a diagonal block
 * is not positive definite (the reference would invert it silently). */
int b200_ls_solve_pcg(b200_linear_solver* ls, int nblocks, int block_dim, const int32_t* colptr, const int32_t* rowidx,
                      const double* values, double* x, const double* b, double tolerance, int absolute_tolerance,
                      int max_iterations, int32_t* iterations, double* residual);
int b200_ls_get_block_ordering(b200_linear_solver* ls, int32_t* perm);
int64_t b200_ls_get_factor_nnz(b200_linear_solver* ls);
const char* b200_ls_last_error(const b200_linear_solver* ls);

/* FNV-1a digest of every array host-only contexts (device -1) of the calling thread have prepared for the device since
 * the last reset: the complete plan of the structure phase (edge order, gather lists, Schur ranges, Cholesky task
 * list ...).  Regression aid: host-side changes of b200_build_structure must leave it unchanged. */
uint64_t b200_debug_upload_digest(int reset);

/* ------------------------------------------------------------------ host-only helpers (no GPU needed)
 * The ordering by itself, for parity checks against cs_amd. */
int b200_block_amd(int nblocks, const int32_t* colptr, const int32_t* rowidx, int32_t* perm);

/* ------------------------------------------------------------------ standalone host
 * .g2o text -> SoA ingest without the g2o object graph: OptimizableGraph::load
 * (core/optimizable_graph.cpp:356-569) for the configured tags, the CLI's gauge + marginalisation
 * (apps/g2o_cli/g2o.cpp:272-320) and SparseOptimizer::initializeOptimization index mapping
 * (core/sparse_optimizer.cpp:166-267), then b200_set_vertices/b200_set_edges on ctx. */
typedef struct b200_graph b200_graph;
int b200_graph_create(b200_graph** out);
void b200_graph_destroy(b200_graph* g);
int b200_graph_load(b200_graph* g, const char* path);
int b200_graph_add_vertex(b200_graph* g, int kind, int id, const double* payload, int n);
int b200_graph_add_edge(b200_graph* g, int kind, int id1, int id2, const double* payload, int n);
/* bulk ingest: payload row-major [n x stride] */
int b200_graph_add_vertices(b200_graph* g, int kind, int n, const int32_t* ids, const double* payload, int stride);
int b200_graph_add_edges(b200_graph* g, int kind, int n, const int32_t* id1, const int32_t* id2, const double* payload, int stride);
int b200_graph_set_fixed(b200_graph* g, int id, int fixed);
/* Edge::setRobustKernel on one edge: edge_index counts the edges in the order they were added / read (internalId);
 * edges without a kernel of their own stay without one as soon as any edge has one (else the context's uniform kernel) */
int b200_graph_set_edge_robust_kernel(b200_graph* g, int edge_index, int kind, double delta);
/* PARAMS_CAMERAPARAMETERS id focal_length cx cy baseline (types/sba/types_six_dof_expmap.h:45-80); has to precede the
 * XYZ2UV edges that name it (payload of such an edge: paramId u v i00 i01 i11, types_six_dof_expmap.cpp:241-256).
 * All edges of one pose must name parameters with equal values (the intrinsics ride in the pose's estimate row). */
int b200_graph_add_camera_parameters(b200_graph* g, int id, double focal_length, double cx, double cy, double baseline);
/* PARAMS_SE3OFFSET id x y z qx qy qz qw (types/slam3d/parameter_se3_offset.cpp:47-56); has to precede the SE3_XYZ
 * edges (EDGE_SE3_TRACKXYZ; payload paramId x y z + upper triangle of the information, edge_se3_pointxyz.cpp:62-84)
 * that name it.  All SE3_XYZ edges of a graph must name offsets with equal values. */
int b200_graph_add_se3_offset(b200_graph* g, int id, const double* xyz_qxyzw);
/* returns the gauge vertex id fixed (or -1 if none needed) */
int b200_graph_setup_cli(b200_graph* g, int requires_marginalize);
int b200_graph_initialize(b200_graph* g);
/* vertex_counts[B200_NUM_VERTEX_KINDS] = #vertices by kind; edge_counts[B200_NUM_EDGE_KINDS] */
int b200_graph_counts(b200_graph* g, int32_t* vertex_counts, int32_t* edge_counts);
/* upload to a context.  shard/num_shards: landmark sharding (0,1 = everything) */
int b200_graph_upload(b200_graph* g, b200_ctx* ctx, int shard, int num_shards);
/* write estimates from ctx back into the graph; b200_graph_get_estimate returns the canonical layout */
int b200_graph_download(b200_graph* g, b200_ctx* ctx);
int b200_graph_get_estimate(b200_graph* g, int id, double* out);
/* info[4] = kind, hessianIndex, fixed, marginalized */
int b200_graph_get_vertex_info(b200_graph* g, int id, int32_t* info);
int b200_graph_save(b200_graph* g, const char* path);   /* OptimizableGraph::save (optimizable_graph.cpp:589-622) */
const char* b200_graph_last_error(const b200_graph* g);

#ifdef __cplusplus
}
#endif
#endif
