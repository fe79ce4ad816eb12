// solver.h - the device-resident solver context behind the C-ABI (include/g2o_b200.h).
//
// Host-side mirror of the reference's BlockSolver<Traits> + OptimizationAlgorithm{GaussNewton,Levenberg}
// for the three configured graph families (SE2 pose graph, SE3 pose graph, CAM+XYZ bundle adjustment):
// same phases, same names, same error behaviour (bool / status, never exceptions across the ABI), with all
// per-edge / per-vertex / per-landmark loops and the sparse Cholesky running as sm_100a kernels.
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "../../include/g2o_b200.h"
#include "chol.h"
#include "common.h"
#include "nccl_comm.h"
#include "pcg_host.h"

struct b200_ctx {
  int device = 0;
  bool host_only = false;  // created with device -1: structure phase only, no compute
  cudaStream_t stream = nullptr;
  // side stream of the linearisation: the per-camera pass (Hpp, b_p) runs beside the per-landmark pass (Hll, Hpl, b_l) -
  // forked from and joined back into `stream` with two events, inside the captured graph as well
  cudaStream_t stream2 = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool overlap_linearize = false;   // measured: no gain (the per-landmark pass fills every SM's register file), kept for experiments (G2O_B200_OVERLAP=1)
  bool lin_packets = true;   // BA linearisation with one lane per observation (ba_linearize_packets_kernel)
  int n_lin_packets = 0;
  int cams_minb = 3;         // ba_linearize_cams_kernel: CTAs per SM it is compiled for (3: 168 registers, 4: 128, 6: 80)
  int lin_minb = 4;          // register budget of that kernel: 4 / 5 / 6 CTAs per SM (128 / 96 / 80 registers)
  g2o_b200::DevBuf<int> d_pk_rank0;
  std::string err;
  g2o_b200::LaunchCounter lc;

  // ---------------- inputs as handed over by the adapter (host copies)
  struct VertexSet {
    bool set = false;
    int n = 0;
    std::vector<double> est;
    std::vector<int> hidx;
    std::vector<unsigned char> marg;
  } vs[B200_NUM_VERTEX_KINDS];   // slot 4 (SE3_EXPMAP) stays empty: that kind lives in the CAM slot
  int edge_kind = -1;
  int nE = 0;
  // landmark SLAM (SE2 + XY / SE3 + XYZ, nothing marginalized: the reference's variable-block-size `*_var` path): the
  // pose-landmark edges (B200_EDGE_SE2_XY / B200_EDGE_SE3_XYZ) beside the pose-pose edges above; the landmarks share the
  // poses' index space, their blocks are padded to the pose dimension (g2o_b200.h)
  int l_edge_kind = -1;
  int nLE = 0;
  std::vector<int> l_vi, l_vj;
  std::vector<double> l_meas, l_info;
  double sensor_offset[12] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0};  // ParameterSE3Offset of the SE3_XYZ edges
  bool var_lm = false;   // structure: landmarks live in the pose index space (no Schur complement)
  int lm_kind = -1;      // B200_VERTEX_XY / B200_VERTEX_XYZ of the landmarks of a var_lm structure
  // camera model of the BA family: 0 VertexCam / EdgeProjectP2MC, 1 VertexSE3Expmap / EdgeProjectXYZ2UV.  Both live in
  // the vs[B200_VERTEX_CAM] slot / edge_kind B200_EDGE_P2MC internally (same block sizes 6, 3, 2; same Schur plan)
  int cam_model = 0, edge_model = 0;
  std::vector<int> e_vi, e_vj;
  std::vector<double> e_meas, e_info;
  std::vector<long long> extra_schur_keys;  // (col<<32|row) blocks other shards contribute to Hschur

  // ---------------- structure (host)
  bool structured = false;
  int pose_kind = -1;
  bool schur = false;
  int np = 0, nl = 0, pd = 0, ld = 0, sizeP = 0, sizeL = 0;
  int n_pose_v = 0, n_lm_v = 0;
  std::vector<int> pose_vertex, lm_vertex;     // index -> vertex slot
  std::vector<int> hpp_colptr, hpp_rowidx;     // upper block pattern of Hpp (CCS)
  std::vector<int> hpp_diag_block;             // pose i -> block index of (i,i)
  std::vector<int> hs_colptr, hs_rowidx;       // Hschur pattern
  std::vector<int> e_order;                    // device edge order -> input edge index (BA)
  std::vector<int> hpl_row, hpl_col;           // per Hpl slot
  std::vector<int> hpl_export;                 // SparseBlockMatrix (landmark-major) position -> slot
  int n_hpl = 0, n_hs = 0, n_hpp = 0;

  // ---------------- device state
  g2o_b200::DevBuf<double> d_pose_est, d_lm_est, d_cam_der, d_pose_bak, d_lm_bak, d_cam_der_bak;
  g2o_b200::DevBuf<int> d_pose_hidx, d_lm_lidx, d_pose_vertex, d_lm_vertex;
  g2o_b200::DevBuf<int> d_ev0, d_ev1, d_e_pose, d_e_hpl;
  g2o_b200::DevBuf<unsigned char> d_e_flag;  // pose graphs: transposed; BA: first-occurrence of its Hpl block
  g2o_b200::DevBuf<double> d_meas, d_info, d_stage;
  g2o_b200::DevBuf<int> d_lev0, d_lev1;          // var_lm: pose / landmark vertex of every pose-landmark edge
  g2o_b200::DevBuf<double> d_lmeas, d_linfo;     // ... their measurements / information (SoA)
  g2o_b200::DevBuf<double> d_pad_diag;           // ... 1 at the padding unknowns of the landmark blocks, 0 elsewhere
  g2o_b200::DevBuf<int> d_hsrc_ptr, d_hsrc_id, d_bsrc_ptr, d_bsrc_id;
  g2o_b200::DevBuf<int> d_lm_order;  // landmark rank (processing order) -> Hessian landmark index
  g2o_b200::DevBuf<int> d_lm_eptr, d_cam_eptr, d_cam_eidx, d_hpp_diag_block;
  g2o_b200::DevBuf<double> d_Hpp, d_Hll, d_Hpl, d_Hschur, d_Dinv, d_Wu, d_b, d_x, d_bschur, d_diag;
  g2o_b200::DevBuf<int> d_t_row, d_t_col, d_t_hpp;
  // Schur plan: landmark ranges, their segments (contributions to one Hschur block), per-block segment lists
  g2o_b200::DevBuf<int> d_sr_slot0, d_sr_lm_ptr, d_sr_lm_ids, d_sr_lm_slot, d_sr_seg_ptr, d_sr_seg_t, d_sr_seg_cb, d_sr_seg_ce, d_tseg_ptr, d_tseg_idx;
  g2o_b200::DevBuf<unsigned short> d_sr_a, d_sr_b, d_sr_l;
  g2o_b200::DevBuf<unsigned char> d_t_diag;
  g2o_b200::DevBuf<double> d_sr_partial;
  int sr_n = 0, sr_nseg = 0, sr_cap_slots = 0, sr_cap_lms = 0, sr_cap_contrib = 0;
  // landmarks seen by more cameras than a range CTA can stage (schur_wide_kernel): pair / segment offsets, Wu index, slots
  int sr_nwide = 0;
  long long sr_wide_pairs = 0;
  g2o_b200::DevBuf<long long> d_sw_pair0, d_sw_seg0;
  g2o_b200::DevBuf<int> d_sw_lm, d_sw_slot0, d_sw_deg;
  long long sr_ncontrib = 0;
  g2o_b200::DevBuf<double> d_stage_est;            // dense staging for host<->device estimate copies
  // scalars: [0] chi2 [1] scale (landmark part; sharded: + pose part) [2] maxdiag [3] lambda [4] scale (pose part)
  //          [5] sharded: chi2 before the trial, summed over the ranks [6] constant 0 [7] sharded: this rank's part of [5]
  g2o_b200::DevBuf<double> d_partials2;  // block sums of the fused trial tail (landmark part of the LM scale)
  int n_edges_free_lm = 0;               // observations of free landmarks (the device edge order lists them first)
  bool fuse_tail = true;                 // BA trials: back-substitution + landmark update + chi2 + scale in one kernel
  g2o_b200::DevBuf<double> d_partials, d_scalars;
  double* h_scalars = nullptr;                     // pinned mirror of d_scalars (+ status as double)
  int* h_status = nullptr;
  int backup_depth = 0;
  g2o_b200::CholeskyGpu chol;
  // linear solver of the (reduced) pose system: 0 = supernodal Cholesky (LinearSolverCSparse / Cholmod), 1 = block-Jacobi
  // PCG (LinearSolverPCG, solvers/pcg): b200_set_linear_solver.  PCG stops on the device but reports to the host every 64
  // iterations, so its trials run as plain launches (no CUDA-graph replay)
  int linear_solver = 0;
  double pcg_tolerance = 1e-6;
  int pcg_absolute = 1, pcg_max_iterations = -1;
  int pcg_last_iterations = 0;
  g2o_b200::PcgGpu pcg;
  // b200_compute_marginals: buffers kept between calls (a front end asks for marginals after every optimisation)
  g2o_b200::DevBuf<double> d_marg_rhs, d_marg_x, d_marg_out;
  g2o_b200::DevBuf<long long> d_marg_off;
  g2o_b200::DevBuf<int> d_marg_ld;
  g2o_b200::DevBuf<unsigned char> d_marg_trans;
  g2o_b200::DevBuf<double> d_pcg_A;   // pose graphs: Hpp + lambda I (+ the unit diagonal of padding unknowns)

  int nd_levels = 0;                           // ordering: 0 = block AMD (reference), k = nested dissection, 2^k parts
  g2o_b200::Robust robust{0, 1.0, nullptr, nullptr};   // robust kernel applied to every edge (b200_set_robust_kernel) ...
  g2o_b200::Robust robust_l{0, 1.0, nullptr, nullptr}; // ... the same for the pose-landmark edge set of a landmark-SLAM graph
  // per-edge kernels (b200_set_edge_robust_kernels): input order on the host, device edge order on the device
  int rk_uniform_kind = 0;
  double rk_uniform_delta = 1.0;
  std::vector<unsigned char> rk_kinds, l_rk_kinds;
  std::vector<double> rk_deltas, l_rk_deltas;
  g2o_b200::DevBuf<unsigned char> d_rk_kinds, d_lrk_kinds;
  g2o_b200::DevBuf<double> d_rk_deltas, d_lrk_deltas;

  // ---------------- algorithm state (core/optimization_algorithm_levenberg.h)
  double lambda = -1.0, ni = 2.0;
  double lambda_for_solve = 0.0;
  int levenberg_iterations = 0;
  double user_lambda_init = 0.0;
  int max_trials_after_failure = 10;
  int num_oplus_calls = 0;
  double last_chi2 = 0.0;

  // ---------------- sharding (landmark shards, cameras replicated).  Two collectives per LM trial:
  // ONE ncclAllReduce over [Hschur | bschur | chi2 of the state before the trial] (every rank contributes its partial
  // Hpp / b_p inside Hschur / bschur, rank 0 the lambda term) and a 2-double one after the update (chi2 of the new
  // state, LM scale); iteration 0 adds one over the Hpp diagonal for the initial lambda.
  g2o_b200::NcclComm nccl;            // native communicator (b200_comm_init); preferred
  b200_allreduce_fn allreduce = nullptr;  // host-supplied callback (b200_set_allreduce): same protocol, no CUDA graphs
  void* allreduce_user = nullptr;
  int rank = 0, world = 1;
  bool comm_warm = false;             // the first sharded trial runs uncaptured (NCCL sets up its channels lazily)
  g2o_b200::DevBuf<double> d_comm_diag;  // iteration 0: [Hpp diagonal partial sums | per-rank landmark maxima]

  // ---------------- CUDA graphs of the two launch-bound sequences of an LM iteration (single GPU, profiling off)
  cudaGraphExec_t graph_prologue = nullptr, graph_trial = nullptr, graph_build = nullptr;
  long long graph_prologue_launches = 0, graph_trial_launches = 0, graph_build_launches = 0;
  // chi2 of the estimates currently on the device, when known: the last LM iteration ended with exactly this state
  // (accepted trial: its chi2; rejected: the restored backup's), so the next iteration's computeActiveErrors pass
  // (optimization_algorithm_levenberg.cpp:71-78) would recompute the same number bit for bit and is skipped.  Any
  // call that changes estimates, edges or the robust kernel from outside invalidates it.
  bool state_chi2_valid = false;
  double state_chi2 = 0.0;
  bool use_graphs = true;

  // SparseOptimizer::terminate() (core/sparse_optimizer.h:189): polled between LM trials and between iterations
  b200_terminate_fn terminate = nullptr;
  void* terminate_user = nullptr;

  // ---------------- profiling
  g2o_b200::EventProfiler prof;
  double time_symbolic = 0.0;
};
