// solver.cu - implementation of the solver context (see solver.h) and of the Level-2 / Level-3 C-ABI.
#include "solver.h"

#include <algorithm>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "host_parallel.h"
#include "kernels.cuh"

using namespace g2o_b200;

namespace {

double wall() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int vdim(int kind) { return kind == B200_VERTEX_SE2 ? 3 : kind == B200_VERTEX_XYZ ? 3 : kind == B200_VERTEX_XY ? 2 : 6; }
int vest(int kind) { return kind == B200_VERTEX_SE2 ? 3 : kind == B200_VERTEX_XYZ ? 3 : kind == B200_VERTEX_XY ? 2 : 12; }   // doubles in the ABI layout
int vstride(int kind) { return (kind == B200_VERTEX_SE2 || kind == B200_VERTEX_XYZ || kind == B200_VERTEX_XY) ? 4 : 12; }  // doubles on the device
int edim(int kind) { return (kind == B200_EDGE_SE2 || kind == B200_EDGE_SE3_XYZ) ? 3 : kind == B200_EDGE_SE3 ? 6 : 2; }
int emeas(int kind) { return (kind == B200_EDGE_SE2 || kind == B200_EDGE_SE3_XYZ) ? 3 : kind == B200_EDGE_SE3 ? 12 : 2; }
bool is_landmark_kind(int kind) { return kind == B200_VERTEX_XYZ || kind == B200_VERTEX_XY; }

// kernel groups for the profiling counters (b200_get_phase_time ids)
enum { PH_ERRORS = 0, PH_LINEARIZE = 1, PH_SCHUR = 2, PH_FACTOR = 3, PH_TRISOLVE = 4, PH_UPDATE = 5, PH_BACKSUB = 6,
       PH_LINEARIZE_CAMS = 7, PH_GATHER = 8, PH_SCHUR_INV = 9, PH_SCALE = 10, PH_COLLECTIVE = 11,
       /* 12, 13, 19: inside the Cholesky, see chol.cu */ PH_SCHUR_FINISH = 14, PH_COUNT = 24 };

// camera-model dispatch of the templated BA kernels
#define BA_MODEL_LAUNCH(c, KERNEL, ...) do { if ((c)->cam_model == 0) k::KERNEL<0> __VA_ARGS__; else k::KERNEL<1> __VA_ARGS__; } while (0)

struct PhaseTimer : ScopedPhase {
  PhaseTimer(b200_ctx* c, int ph) : ScopedPhase(&c->prof, ph) {}
};

template <typename F>
int guarded(b200_ctx* c, F&& f) {
  if (!c) return B200_ERR_INVALID;
  host_only_flag() = c->host_only;
  try {
    return f();
  } catch (const CudaError& e) {
    c->err = describe(e);
    if (e.code == cudaErrorNoDevice || e.code == cudaErrorInsufficientDriver) return B200_ERR_NO_DEVICE;
    return B200_ERR_CUDA;
  } catch (const std::exception& e) {
    c->err = std::string("exception: ") + e.what();
    return B200_ERR_EXCEPTION;
  }
}
int fail(b200_ctx* c, int code, const std::string& msg) {
  c->err = msg;
  return code;
}

std::string g_create_error;
#define NEED_DEVICE_(c) \
  if ((c)->host_only) return fail((c), B200_ERR_NO_DEVICE, "host-only context: the B200 solve path has no CPU fallback")

// ---- host math for the one-time ingest (same formulas as geometry.cuh; runs where the reference's read() runs)
void host_se2_inverse(const double* z, double* zi) {
  geo::SE2 r = geo::se2_inv(geo::SE2{z[0], z[1], z[2]});
  zi[0] = r.x; zi[1] = r.y; zi[2] = r.th;
}
void host_iso_inverse(const double* Z, double* Zi) {
  geo::Iso a;
  memcpy(a.R, Z, 9 * sizeof(double)); memcpy(a.t, Z + 9, 3 * sizeof(double));
  geo::Iso r = geo::iso_inverse(a);
  memcpy(Zi, r.R, 9 * sizeof(double)); memcpy(Zi + 9, r.t, 3 * sizeof(double));
}

void sync_scalars(b200_ctx* c) {
  B200_CUDA(cudaMemcpyAsync(c->h_scalars, c->d_scalars.p, 8 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  B200_CUDA(cudaMemcpyAsync(c->h_status, c->chol.status_ptr(), sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  B200_CUDA(cudaStreamSynchronize(c->stream));
}
void set_device_lambda(b200_ctx* c, double lam) {
  c->h_scalars[8] = lam;  // staging slot beyond the mirrored range
  B200_CUDA(cudaMemcpyAsync(c->d_scalars.p + 3, c->h_scalars + 8, sizeof(double), cudaMemcpyHostToDevice, c->stream));
}

// ------------------------------------------------------------------------------------------------
// structure
// ------------------------------------------------------------------------------------------------
void drop_graphs(b200_ctx* c) {
  if (c->graph_prologue) { cudaGraphExecDestroy(c->graph_prologue); c->graph_prologue = nullptr; }
  if (c->graph_trial) { cudaGraphExecDestroy(c->graph_trial); c->graph_trial = nullptr; }
  if (c->graph_build) { cudaGraphExecDestroy(c->graph_build); c->graph_build = nullptr; }
}

static size_t schur_range_smem(const b200_ctx* c) {
  return ((size_t)c->sr_cap_slots * 18 + (size_t)c->sr_cap_lms * k::kDinvStride) * sizeof(double) +
         (size_t)((c->sr_cap_lms + 1 + 3) & ~3) * sizeof(int) + (size_t)c->sr_cap_contrib * 3 * sizeof(unsigned short) +
         (size_t)((c->sr_cap_slots + 7) & ~7) * sizeof(unsigned short);   // + the landmark of every slot
}

// per-edge robust kernels -> device edge order (BA: the landmark-rank order of e_order; pose graphs: input order)
void upload_robust(b200_ctx* c) {
  cudaStream_t s = c->stream;
  c->robust = Robust{c->rk_uniform_kind, c->rk_uniform_delta, nullptr, nullptr};
  c->robust_l = c->robust;
  if ((int)c->rk_kinds.size() == c->nE && c->nE > 0) {
    std::vector<unsigned char> kd(c->nE);
    std::vector<double> dl(c->nE);
    for (int q = 0; q < c->nE; ++q) {
      const int e = c->schur ? c->e_order[q] : q;
      kd[q] = c->rk_kinds[e]; dl[q] = c->rk_deltas[e];
    }
    c->d_rk_kinds.upload(kd, s); c->d_rk_deltas.upload(dl, s);
    if (!c->host_only) B200_CUDA(cudaStreamSynchronize(s));
    c->robust.kinds = c->d_rk_kinds.p; c->robust.deltas = c->d_rk_deltas.p;
    c->robust.kind = 1;   // "some kernel": the kernels look the edge's own kind up
  }
  if (c->var_lm && (int)c->l_rk_kinds.size() == c->nLE && c->nLE > 0) {
    c->d_lrk_kinds.upload(c->l_rk_kinds, s); c->d_lrk_deltas.upload(c->l_rk_deltas, s);
    if (!c->host_only) B200_CUDA(cudaStreamSynchronize(s));
    c->robust_l.kinds = c->d_lrk_kinds.p; c->robust_l.deltas = c->d_lrk_deltas.p;
    c->robust_l.kind = 1;
  }
}

int build_structure_impl(b200_ctx* c) {
  double _t_last = wall();
  const bool _tv = getenv("G2O_B200_STRUCT_VERBOSE") != nullptr;
#define STAMP(name) do { if (_tv) { double _n = wall(); fprintf(stderr, "  structure %-28s %.3f s\n", name, _n - _t_last); _t_last = _n; } } while (0)

  double t0 = wall();
  drop_graphs(c);
  cudaStream_t s = c->stream;
  // which graph family?
  int pose_kind = -1;
  for (int k = 0; k < 3; ++k)
    if (c->vs[k].set && c->vs[k].n > 0) {
      if (pose_kind >= 0) return fail(c, B200_ERR_UNSUPPORTED, "more than one pose vertex kind in one context");
      pose_kind = k;
    }
  if (pose_kind < 0) return fail(c, B200_ERR_INVALID, "no pose vertices set");
  // landmark SLAM: SE2 + XY / SE3 + XYZ with pose-landmark edges, nothing marginalized (variable-block `*_var` path)
  const bool var_lm = c->nLE > 0;
  const int lm_kind = var_lm ? (c->l_edge_kind == B200_EDGE_SE2_XY ? B200_VERTEX_XY : B200_VERTEX_XYZ) : B200_VERTEX_XYZ;
  if (!var_lm && c->vs[B200_VERTEX_XY].set && c->vs[B200_VERTEX_XY].n > 0) return fail(c, B200_ERR_UNSUPPORTED, "XY vertices need SE2_XY edges");
  const bool has_lm = !var_lm && c->vs[B200_VERTEX_XYZ].set && c->vs[B200_VERTEX_XYZ].n > 0;
  if (var_lm) {
    const int want = c->l_edge_kind == B200_EDGE_SE2_XY ? B200_VERTEX_SE2 : B200_VERTEX_SE3;
    if (want != pose_kind) return fail(c, B200_ERR_UNSUPPORTED, "SE2_XY edges go with SE2 poses, SE3_XYZ edges with SE3 poses");
    if (c->nE > 0 && c->edge_kind != (want == B200_VERTEX_SE2 ? B200_EDGE_SE2 : B200_EDGE_SE3)) return fail(c, B200_ERR_UNSUPPORTED, "edge kind does not match the pose vertex kind");
    if (!c->vs[lm_kind].set || c->vs[lm_kind].n <= 0) return fail(c, B200_ERR_INVALID, "pose-landmark edges without landmark vertices");
    if (c->world > 1) return fail(c, B200_ERR_UNSUPPORTED, "landmark SLAM graphs are not sharded");
    if (c->nE <= 0) c->edge_kind = want == B200_VERTEX_SE2 ? B200_EDGE_SE2 : B200_EDGE_SE3;
  } else {
    if (c->edge_kind < 0 || c->nE <= 0) return fail(c, B200_ERR_INVALID, "no edges set");
    const int want_pose = c->edge_kind == B200_EDGE_SE2 ? B200_VERTEX_SE2 : c->edge_kind == B200_EDGE_SE3 ? B200_VERTEX_SE3 : B200_VERTEX_CAM;
    if (want_pose != pose_kind) return fail(c, B200_ERR_UNSUPPORTED, "edge kind does not match the pose vertex kind");
    if ((c->edge_kind == B200_EDGE_P2MC) != has_lm) return fail(c, B200_ERR_UNSUPPORTED, "P2MC / XYZ2UV edges need XYZ vertices (and only they do)");
    if (c->edge_kind == B200_EDGE_P2MC && c->edge_model != c->cam_model) return fail(c, B200_ERR_UNSUPPORTED, "P2MC edges go with CAM vertices, XYZ2UV edges with SE3_EXPMAP vertices");
  }
  c->pose_kind = pose_kind;
  c->schur = has_lm;
  c->var_lm = var_lm;
  c->lm_kind = var_lm ? lm_kind : -1;
  b200_ctx::VertexSet& PV = c->vs[pose_kind];
  b200_ctx::VertexSet& LV = c->vs[lm_kind];
  c->n_pose_v = PV.n;
  c->n_lm_v = (has_lm || var_lm) ? LV.n : 0;
  c->pd = vdim(pose_kind);
  c->ld = var_lm ? vdim(lm_kind) : 3;
  // index mapping as assigned by SparseOptimizer::buildIndexMapping: poses first, then landmarks
  int np = 0, nl = 0;
  for (int v = 0; v < PV.n; ++v) {
    if (PV.hidx[v] >= 0) ++np;
    if (!PV.marg.empty() && PV.marg[v]) return fail(c, B200_ERR_UNSUPPORTED, "marginalized pose vertices are not supported");
  }
  if (has_lm)
    for (int v = 0; v < LV.n; ++v) {
      if (LV.hidx[v] >= 0) {
        ++nl;
        if (LV.marg.empty() || !LV.marg[v]) return fail(c, B200_ERR_UNSUPPORTED, "XYZ vertices must be marginalized (Schur) for this solver");
      }
    }
  if (var_lm)
    for (int v = 0; v < LV.n; ++v) {  // the landmarks are numbered with the poses (buildIndexMapping: nothing marginalized)
      if (LV.hidx[v] >= 0) ++np;
      if (!LV.marg.empty() && LV.marg[v]) return fail(c, B200_ERR_UNSUPPORTED, "marginalized XY / XYZ landmarks of a landmark-SLAM graph (Schur complement) are not supported: use the variable-block path (nothing marginalized)");
    }
  if (np == 0) return fail(c, B200_ERR_INVALID, "0 vertices to optimize");
  c->np = np; c->nl = nl;
  c->sizeP = np * c->pd; c->sizeL = nl * c->ld;
  c->pose_vertex.assign(np, -1);
  for (int v = 0; v < PV.n; ++v) {
    int h = PV.hidx[v];
    if (h < 0) continue;
    if (h >= np || c->pose_vertex[h] != -1) return fail(c, B200_ERR_INVALID, "pose hessian indices must be a permutation of [0,numPoses)");
    c->pose_vertex[h] = v;
  }
  if (var_lm)
    for (int v = 0; v < LV.n; ++v) {
      const int h = LV.hidx[v];
      if (h < 0) continue;
      if (h >= np || c->pose_vertex[h] != -1) return fail(c, B200_ERR_INVALID, "pose hessian indices must be a permutation of [0,numPoses)");
      c->pose_vertex[h] = -2 - v;  // a landmark
    }
  c->lm_vertex.assign(nl, -1);
  std::vector<int> lm_lidx(c->n_lm_v, -1);
  if (has_lm)
    for (int v = 0; v < LV.n; ++v) {
      int h = LV.hidx[v];
      if (h < 0) continue;
      int l = h - np;
      if (l < 0 || l >= nl || c->lm_vertex[l] != -1) return fail(c, B200_ERR_INVALID, "landmark hessian indices must follow the poses contiguously");
      c->lm_vertex[l] = v;
      lm_lidx[v] = l;
    }
  const int E = c->nE;
  for (int e = 0; e < E; ++e) {
    const int nvi = has_lm ? LV.n : PV.n;
    if (c->e_vi[e] < 0 || c->e_vi[e] >= nvi || c->e_vj[e] < 0 || c->e_vj[e] >= PV.n) return fail(c, B200_ERR_INVALID, "edge vertex index out of range");
  }
  const int LE = var_lm ? c->nLE : 0;
  for (int e = 0; e < LE; ++e)
    if (c->l_vi[e] < 0 || c->l_vi[e] >= PV.n || c->l_vj[e] < 0 || c->l_vj[e] >= LV.n) return fail(c, B200_ERR_INVALID, "edge vertex index out of range");

  // ---- vertex state on the device
  {
    const int st = vstride(pose_kind), ne = vest(pose_kind);
    std::vector<double> buf((size_t)PV.n * st, 0.0);
    for (int v = 0; v < PV.n; ++v) memcpy(&buf[(size_t)v * st], &PV.est[(size_t)v * ne], ne * sizeof(double));
    c->d_pose_est.upload(buf, s);
    c->d_pose_bak.alloc(buf.size());
    c->d_pose_hidx.upload(PV.hidx, s);
    c->d_pose_vertex.upload(c->pose_vertex, s);
    if (pose_kind == B200_VERTEX_CAM) {
      c->d_cam_der.alloc((size_t)PV.n * 16);
      c->d_cam_der_bak.alloc((size_t)PV.n * 16);
      if (!c->host_only) {
        BA_MODEL_LAUNCH(c, cam_derive_kernel, <<<ceil_div(PV.n, 128), 128, 0, s>>>(PV.n, c->d_pose_est.p, c->d_cam_der.p));
        c->lc.n++;
      }
    }
    if (has_lm) {
      std::vector<double> lb((size_t)LV.n * 4, 0.0);
      for (int v = 0; v < LV.n; ++v) memcpy(&lb[(size_t)v * 4], &LV.est[(size_t)v * 3], 3 * sizeof(double));
      c->d_lm_est.upload(lb, s);
      c->d_lm_bak.alloc(lb.size());
      c->d_lm_lidx.upload(lm_lidx, s);
      c->d_lm_vertex.upload(c->lm_vertex, s);
    }
    if (var_lm) {
      const int lne = vest(lm_kind);
      std::vector<double> lb((size_t)LV.n * 4, 0.0);
      for (int v = 0; v < LV.n; ++v) memcpy(&lb[(size_t)v * 4], &LV.est[(size_t)v * lne], lne * sizeof(double));
      c->d_lm_est.upload(lb, s);
      c->d_lm_bak.alloc(lb.size());
      c->d_lm_lidx.upload(LV.hidx, s);   // index in the common pose / landmark numbering
    }
    if (!c->host_only) B200_CUDA(cudaStreamSynchronize(s));
  }
  STAMP("checks + vertex upload");
  const int ntot = c->sizeP + c->sizeL;
  c->d_b.alloc(ntot); c->d_x.alloc(ntot);
  c->d_b.zero(s); c->d_x.zero(s);
  c->d_diag.alloc(ntot);
  c->d_scalars.alloc(16); c->d_scalars.zero(s);
  c->d_partials.alloc((size_t)ceil_div(std::max(E, ntot), 256) + (size_t)ceil_div(std::max(LE, 1), 256) + (size_t)ceil_div(std::max(nl, 1), 128) + 64);

  const int D = edim(c->edge_kind);
  const int pd = c->pd;
  const int ET = E + LE;   // pose graphs: records of the per-edge staging (pose-pose edges first, then pose-landmark)
  std::vector<int> bp_colptr, bp_rowidx;  // pattern handed to the Cholesky

  if (!has_lm) {
    // =========================== pose graph ===========================
    // Hpp pattern: diagonal blocks + (min,max) per edge with two free vertices (block_solver.hpp:204-232)
    auto h_i = [&](int e) { return e < E ? PV.hidx[c->e_vi[e]] : PV.hidx[c->l_vi[e - E]]; };
    auto h_j = [&](int e) { return e < E ? PV.hidx[c->e_vj[e]] : LV.hidx[c->l_vj[e - E]]; };
    std::vector<long long> keys;
    keys.reserve((size_t)np + ET);
    for (int i = 0; i < np; ++i) keys.push_back(((long long)i << 32) | i);
    for (int e = 0; e < ET; ++e) {
      int hi = h_i(e), hj = h_j(e);
      if (hi < 0 || hj < 0) continue;
      if (hi == hj) return fail(c, B200_ERR_UNSUPPORTED, "self-loop edge");
      int lo = std::min(hi, hj), hi2 = std::max(hi, hj);
      keys.push_back(((long long)hi2 << 32) | lo);
    }
    std::sort(keys.begin(), keys.end());
    keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
    const int nblk = (int)keys.size();
    c->n_hpp = nblk;
    c->hpp_colptr.assign(np + 1, 0);
    c->hpp_rowidx.resize(nblk);
    for (int k2 = 0; k2 < nblk; ++k2) { c->hpp_colptr[(keys[k2] >> 32) + 1]++; c->hpp_rowidx[k2] = (int)(keys[k2] & 0xffffffff); }
    for (int i = 0; i < np; ++i) c->hpp_colptr[i + 1] += c->hpp_colptr[i];
    auto find_block = [&](int row, int col) {
      long long key = ((long long)col << 32) | row;
      return (int)(std::lower_bound(keys.begin(), keys.end(), key) - keys.begin());
    };
    c->hpp_diag_block.resize(np);
    for (int i = 0; i < np; ++i) c->hpp_diag_block[i] = find_block(i, i);
    // per-edge flags and ordered gather lists
    std::vector<unsigned char> transposed(ET, 0);
    std::vector<int> hcnt(nblk + 1, 0), bcnt(np + 1, 0);
    std::vector<int> eb_ii(ET, -1), eb_jj(ET, -1), eb_ij(ET, -1);
    for (int e = 0; e < ET; ++e) {
      int hi = h_i(e), hj = h_j(e);
      if (hi >= 0) { eb_ii[e] = c->hpp_diag_block[hi]; hcnt[eb_ii[e] + 1]++; bcnt[hi + 1]++; }
      if (hj >= 0) { eb_jj[e] = c->hpp_diag_block[hj]; hcnt[eb_jj[e] + 1]++; bcnt[hj + 1]++; }
      if (hi >= 0 && hj >= 0) {
        transposed[e] = hi > hj;
        eb_ij[e] = find_block(std::min(hi, hj), std::max(hi, hj));
        hcnt[eb_ij[e] + 1]++;
      }
    }
    for (int k2 = 0; k2 < nblk; ++k2) hcnt[k2 + 1] += hcnt[k2];
    for (int i = 0; i < np; ++i) bcnt[i + 1] += bcnt[i];
    std::vector<int> hsrc(hcnt[nblk]), bsrc(bcnt[np]);
    {
      std::vector<int> hf(hcnt.begin(), hcnt.end() - 1), bf(bcnt.begin(), bcnt.end() - 1);
      if ((long long)ET * 5 > 0x7fffffffll) return fail(c, B200_ERR_UNSUPPORTED, "more than 4e8 edges in one pose graph");
      for (int e = 0; e < ET; ++e) {
        int hi = h_i(e), hj = h_j(e);
        if (eb_ii[e] >= 0) { hsrc[hf[eb_ii[e]]++] = e * 5 + 0; bsrc[bf[hi]++] = e * 5 + 3; }
        if (eb_jj[e] >= 0) { hsrc[hf[eb_jj[e]]++] = e * 5 + 1; bsrc[bf[hj]++] = e * 5 + 4; }
        if (eb_ij[e] >= 0) hsrc[hf[eb_ij[e]]++] = e * 5 + 2;
      }
    }
    // edge data, SoA
    const int MS = c->edge_kind == B200_EDGE_SE2 ? 3 : 12;
    const int IS = D * (D + 1) / 2;
    std::vector<double> meas((size_t)MS * E), info((size_t)IS * E);
    for (int e = 0; e < E; ++e) {
      double zi[12];
      if (c->edge_kind == B200_EDGE_SE2) host_se2_inverse(&c->e_meas[(size_t)e * 3], zi);
      else host_iso_inverse(&c->e_meas[(size_t)e * 12], zi);
      for (int f = 0; f < MS; ++f) meas[(size_t)f * E + e] = zi[f];
      int f = 0;
      const double* W = &c->e_info[(size_t)e * D * D];
      for (int i = 0; i < D; ++i) for (int j = i; j < D; ++j) info[(size_t)(f++) * E + e] = W[i + D * j];
    }
    c->d_ev0.upload(c->e_vi, s); c->d_ev1.upload(c->e_vj, s);
    c->d_meas.upload(meas, s); c->d_info.upload(info, s);
    c->d_e_flag.upload(transposed, s);
    c->d_hsrc_ptr.upload(hcnt, s); c->d_hsrc_id.upload(hsrc, s);
    c->d_bsrc_ptr.upload(bcnt, s); c->d_bsrc_id.upload(bsrc, s);
    c->d_hpp_diag_block.upload(c->hpp_diag_block, s);
    c->d_stage.alloc((size_t)ET * (3 * D * D + 2 * D));
    c->d_Hpp.alloc((size_t)nblk * pd * pd);
    if (var_lm) {
      // pose-landmark edges (SoA) and the unit diagonal of the padding unknowns of the landmark blocks
      const int LDm = edim(c->l_edge_kind), LMS = emeas(c->l_edge_kind), LIS = LDm * (LDm + 1) / 2;
      std::vector<double> lmeas((size_t)LMS * LE), linfo((size_t)LIS * LE);
      for (int e = 0; e < LE; ++e) {
        for (int f = 0; f < LMS; ++f) lmeas[(size_t)f * LE + e] = c->l_meas[(size_t)e * LMS + f];
        int f = 0;
        const double* W = &c->l_info[(size_t)e * LDm * LDm];
        for (int i = 0; i < LDm; ++i) for (int j = i; j < LDm; ++j) linfo[(size_t)(f++) * LE + e] = W[i + LDm * j];
      }
      c->d_lev0.upload(c->l_vi, s); c->d_lev1.upload(c->l_vj, s);
      c->d_lmeas.upload(lmeas, s); c->d_linfo.upload(linfo, s);
      std::vector<double> pad((size_t)np * pd, 0.0);
      for (int v = 0; v < LV.n; ++v)
        if (LV.hidx[v] >= 0) for (int k2 = c->ld; k2 < pd; ++k2) pad[(size_t)LV.hidx[v] * pd + k2] = 1.0;
      c->d_pad_diag.upload(pad, s);
    }
    if (!c->host_only) B200_CUDA(cudaStreamSynchronize(s));
    bp_colptr = c->hpp_colptr; bp_rowidx = c->hpp_rowidx;
  } else {
    // =========================== bundle adjustment ===========================
    // Landmark processing order ("rank"): lexicographic by the list of observing cameras, so that neighbouring
    // landmarks feed the same Hschur blocks and a contiguous range of Hpl is one Schur work unit.  Every
    // landmark-major device structure (edge order, lm_eptr, Hpl slots) follows it; arrays indexed by the landmark's
    // Hessian index (estimates, Hll, b, x, Dinv) are reached through lm_order.
    // edges grouped by landmark (fixed points = group nl, last), inside a group ascending camera pose index, ties in
    // input order: a counting sort over the landmarks + a small stable sort per group (k is 2..10 for most landmarks)
    auto lkey = [&](int e) { int l = lm_lidx[c->e_vi[e]]; return l < 0 ? nl : l; };
    std::vector<int> order(E), grp_ptr(nl + 2, 0);
    std::vector<int> e_cam_hidx(E);
    for (int e = 0; e < E; ++e) { grp_ptr[lkey(e) + 1]++; e_cam_hidx[e] = PV.hidx[c->e_vj[e]]; }
    for (int l = 0; l <= nl; ++l) grp_ptr[l + 1] += grp_ptr[l];
    {
      std::vector<int> fill(grp_ptr.begin(), grp_ptr.end() - 1);
      for (int e = 0; e < E; ++e) order[fill[lkey(e)]++] = e;
      parallel_ranges((size_t)nl + 1, range_count((size_t)nl + 1, (size_t)1 << 14), [&](int, size_t lb, size_t le) {
      for (size_t l = lb; l < le; ++l) {
        int* b0 = order.data() + grp_ptr[l];
        const int k2 = grp_ptr[l + 1] - grp_ptr[l];
        if (k2 <= 24) {  // stable insertion sort
          for (int i = 1; i < k2; ++i) {
            const int v = b0[i], kv = e_cam_hidx[v];
            int j = i - 1;
            for (; j >= 0 && e_cam_hidx[b0[j]] > kv; --j) b0[j + 1] = b0[j];
            b0[j + 1] = v;
          }
        } else {
          std::stable_sort(b0, b0 + k2, [&](int a, int b) { return e_cam_hidx[a] < e_cam_hidx[b]; });
        }
      }
      });
    }
  STAMP("edge sort by landmark");
    std::vector<int> lm_order(nl), lm_rank(nl);
    {
      std::vector<int> cl_ptr(nl + 1, 0), cl;  // per landmark: ascending distinct free cameras
      cl.reserve(E);
      for (int l = 0; l < nl; ++l) {
        int prev = -1;
        for (int q = grp_ptr[l]; q < grp_ptr[l + 1]; ++q) {
          const int pz = e_cam_hidx[order[q]];
          if (pz >= 0 && pz != prev) { cl.push_back(pz); prev = pz; }
        }
        cl_ptr[l + 1] = (int)cl.size();
      }
      // lexicographic order of the camera lists, landmarks without a free camera last, ties by landmark index.  The
      // first cameras of a list are packed into one integer (as many as fit into 63 bits; a missing entry sorts
      // first), so most comparisons are one 64-bit compare; only equal prefixes look at the lists.
      struct RankKey { unsigned long long key; int l; };
      std::vector<RankKey> rk(nl);
      int bits = 1;
      while ((1ll << bits) < (long long)np + 2) ++bits;  // values 0 (missing) .. np
      const int npack = 63 / bits;
      for (int l = 0; l < nl; ++l) {
        const int n = cl_ptr[l + 1] - cl_ptr[l];
        const int* r = cl.data() + cl_ptr[l];
        unsigned long long key = ~0ull;
        if (n > 0) {
          key = 0;
          for (int q = 0; q < npack; ++q) key = (key << bits) | (unsigned long long)(q < n ? r[q] + 1 : 0);
        }
        rk[l] = RankKey{key, l};
      }
      parallel_sort(rk.begin(), rk.end(), [&](const RankKey& X, const RankKey& Y) {
        if (X.key != Y.key) return X.key < Y.key;
        const int x = X.l, y = Y.l;
        const int nx = cl_ptr[x + 1] - cl_ptr[x], ny = cl_ptr[y + 1] - cl_ptr[y];
        const int* rx = cl.data() + cl_ptr[x];
        const int* ry = cl.data() + cl_ptr[y];
        for (int q = npack; q < nx && q < ny; ++q) if (rx[q] != ry[q]) return rx[q] < ry[q];
        if (nx != ny) return nx < ny;
        return x < y;
      });
      for (int i = 0; i < nl; ++i) { lm_order[i] = rk[i].l; lm_rank[rk[i].l] = i; }
    }
  STAMP("landmark ranking");
    // device edge order: by landmark rank (fixed points last), then camera pose index, then input order
    {
      std::vector<int> ranked(E);
      int q = 0;
      for (int i = 0; i <= nl; ++i) {
        const int l = i < nl ? lm_order[i] : nl;
        for (int a = grp_ptr[l]; a < grp_ptr[l + 1]; ++a) ranked[q++] = order[a];
      }
      order.swap(ranked);
    }
  STAMP("edge sort by rank");
    c->e_order = order;
    std::vector<int> e_pt(E), e_cam(E), e_pose(E), e_hpl(E, -1), lm_eptr(nl + 1, 0);
    std::vector<unsigned char> e_first(E, 0);
    std::vector<double> meas((size_t)2 * E), info((size_t)3 * E);
    // An observation opens a new Hpl slot unless the previous edge (device order) is the same landmark seen by the same
    // free camera (duplicate observation: both feed one block).  Slots are numbered in device order, so a slot index
    // is a prefix count of "opens a slot": counted per contiguous range of edges, offset by a prefix sum, filled
    // concurrently - identical for any thread count.
    auto e_l = [&](int q) { return lm_lidx[c->e_vi[order[q]]]; };
    auto opens_slot = [&](int q, int l, int pose) {
      if (l < 0 || pose < 0) return false;
      return !(q > 0 && e_l(q - 1) == l && e_cam_hidx[order[q - 1]] == pose);
    };
    const int eparts = range_count((size_t)E, (size_t)1 << 16);
    std::vector<int> slot_base(eparts + 1, 0);
    parallel_ranges((size_t)E, eparts, [&](int t, size_t b, size_t e2) {
      int cnt = 0;
      for (size_t q = b; q < e2; ++q) {
        const int e = order[q];
        cnt += opens_slot((int)q, lm_lidx[c->e_vi[e]], e_cam_hidx[e]);
      }
      slot_base[t + 1] = cnt;
    });
    for (int t = 0; t < eparts; ++t) slot_base[t + 1] += slot_base[t];
    const int nslot = slot_base[eparts];
    c->hpl_row.assign(nslot, 0); c->hpl_col.assign(nslot, 0);
    parallel_ranges((size_t)E, eparts, [&](int t, size_t b, size_t e2) {
      int next = slot_base[t];
      for (size_t q = b; q < e2; ++q) {
        const int e = order[q];
        e_pt[q] = c->e_vi[e]; e_cam[q] = c->e_vj[e];
        const int pose = e_cam_hidx[e], l = lm_lidx[c->e_vi[e]];
        e_pose[q] = pose;
        if (l >= 0 && pose >= 0) {
          if (opens_slot((int)q, l, pose)) { e_hpl[q] = next; e_first[q] = 1; c->hpl_row[next] = pose; c->hpl_col[next] = l; ++next; }
          else e_hpl[q] = next - 1;  // duplicate: the slot the run opened (possibly in the previous range)
        }
        meas[q] = c->e_meas[(size_t)e * 2]; meas[(size_t)E + q] = c->e_meas[(size_t)e * 2 + 1];
        const double* W = &c->e_info[(size_t)e * 4];
        info[q] = W[0]; info[(size_t)E + q] = W[2]; info[(size_t)2 * E + q] = W[3];
      }
    });
    // edges are in rank order: a landmark's edges are its group
    for (int i = 0; i < nl; ++i) lm_eptr[i + 1] = lm_eptr[i] + (grp_ptr[lm_order[i] + 1] - grp_ptr[lm_order[i]]);
    c->n_hpl = nslot;
  STAMP("edge arrays + Hpl slots");
    // camera observation lists (ascending device edge index)
    std::vector<int> cam_eptr(np + 1, 0), cam_eidx;
    for (int q = 0; q < E; ++q) if (e_pose[q] >= 0) cam_eptr[e_pose[q] + 1]++;
    for (int i = 0; i < np; ++i) cam_eptr[i + 1] += cam_eptr[i];
    cam_eidx.resize(cam_eptr[np]);
    {
      std::vector<int> f(cam_eptr.begin(), cam_eptr.end() - 1);
      for (int q = 0; q < E; ++q) if (e_pose[q] >= 0) cam_eidx[f[e_pose[q]]++] = q;
    }
    // Hpp: P2MC graphs have no camera-camera edges -> block diagonal
    c->n_hpp = np;
    c->hpp_colptr.resize(np + 1); c->hpp_rowidx.resize(np); c->hpp_diag_block.resize(np);
    for (int i = 0; i < np; ++i) { c->hpp_colptr[i] = i; c->hpp_rowidx[i] = i; c->hpp_diag_block[i] = i; }
    c->hpp_colptr[np] = np;
  STAMP("camera lists");
    // Hschur pattern (block_solver.hpp:262-288): Hpp pattern + co-observation pairs (i1<=i2)
    std::vector<long long> keys;
    auto compact = [&]() { std::sort(keys.begin(), keys.end()); keys.erase(std::unique(keys.begin(), keys.end()), keys.end()); };
    for (int i = 0; i < np; ++i) keys.push_back(((long long)i << 32) | i);
    for (long long k2 : c->extra_schur_keys) {  // blocks other shards contribute (b200_add_schur_pattern)
      if ((int)(k2 >> 32) >= np) return fail(c, B200_ERR_INVALID, "b200_add_schur_pattern: block index beyond the number of free poses");
      keys.push_back(k2);
    }
    size_t next_compact = std::max<size_t>(keys.size() * 2, (size_t)1 << 24);
    // distinct Hpl slots per landmark rank (slots are numbered in rank order; hpl_row[slot] = camera, ascending per landmark)
    std::vector<int> lm_s0(nl + 1, 0);
    {
      int prev = -1;
      for (int i = 0; i < nl; ++i) {
        for (int a = lm_eptr[i]; a < lm_eptr[i + 1]; ++a)
          if (e_hpl[a] >= 0 && e_hpl[a] != prev) { prev = e_hpl[a]; lm_s0[i + 1]++; }
      }
      for (int i = 0; i < nl; ++i) lm_s0[i + 1] += lm_s0[i];
    }
    // neighbours in rank order mostly see the same cameras: a landmark whose camera list equals its predecessor's adds
    // nothing to the pattern (and reuses its block indices in the Schur plan below)
    std::vector<unsigned char> same_as_prev(nl, 0);
    for (int i = 1; i < nl; ++i) {
      const int k2 = lm_s0[i + 1] - lm_s0[i];
      same_as_prev[i] = k2 > 0 && k2 == lm_s0[i] - lm_s0[i - 1] &&
                        std::equal(c->hpl_row.begin() + lm_s0[i], c->hpl_row.begin() + lm_s0[i + 1], c->hpl_row.begin() + lm_s0[i - 1]);
    }
    for (int i = 0; i < nl; ++i) {
      if (same_as_prev[i]) continue;
      for (int a = lm_s0[i]; a < lm_s0[i + 1]; ++a)
        for (int b2 = a; b2 < lm_s0[i + 1]; ++b2) keys.push_back(((long long)c->hpl_row[b2] << 32) | c->hpl_row[a]);
      if (keys.size() > next_compact) { compact(); next_compact = std::max<size_t>(keys.size() * 2, (size_t)1 << 24); }
    }
    compact();
    const int nT = (int)keys.size();
    c->n_hs = nT;
    c->hs_colptr.assign(np + 1, 0); c->hs_rowidx.resize(nT);
    std::vector<int> t_row(nT), t_col(nT), t_hpp(nT, -1);
    for (int t = 0; t < nT; ++t) {
      t_col[t] = (int)(keys[t] >> 32); t_row[t] = (int)(keys[t] & 0xffffffff);
      c->hs_colptr[t_col[t] + 1]++; c->hs_rowidx[t] = t_row[t];
      if (t_row[t] == t_col[t]) t_hpp[t] = t_row[t];
    }
    for (int i = 0; i < np; ++i) c->hs_colptr[i + 1] += c->hs_colptr[i];
    auto find_t = [&](int row, int col) {
      long long key = ((long long)col << 32) | row;
      return (int)(std::lower_bound(keys.begin(), keys.end(), key) - keys.begin());
    };
  STAMP("Hschur pattern");
    // ---- Schur plan (kernels.cuh: schur_range_kernel / schur_finish_kernel)
    // landmarks seen by more cameras than a range CTA can stage go through schur_wide_kernel (one thread per camera pair)
    int wide_thr = 1400;
    if (const char* e = getenv("G2O_B200_SR_WIDE")) wide_thr = std::max(1, std::min(1400, atoi(e)));   // tests: force the wide path
    std::vector<int> pi, wide;  // ranks with at least one slot: range landmarks / wide landmarks
    for (int i = 0; i < nl; ++i)
      if (lm_s0[i + 1] > lm_s0[i]) (lm_s0[i + 1] - lm_s0[i] > wide_thr ? wide : pi).push_back(i);
    {
      // SparseBlockMatrix (landmark-major, ascending camera) position -> slot, for the exported Hpl pattern
      c->hpl_export.resize(nslot);
      std::vector<int> colfill(nl + 1, 0);  // stable counting sort by landmark column
      for (int q = 0; q < nslot; ++q) colfill[c->hpl_col[q] + 1]++;
      for (int l = 0; l < nl; ++l) colfill[l + 1] += colfill[l];
      for (int q = 0; q < nslot; ++q) c->hpl_export[colfill[c->hpl_col[q]]++] = q;
    }
    {
  STAMP("hpl export order");
      // shared memory per range CTA (kernels.cuh): 3 CTAs per SM unless one landmark alone needs more slots
      int kmax = 0;
      for (int l : pi) kmax = std::max(kmax, lm_s0[l + 1] - lm_s0[l]);
      const int cap_slots = std::max(352, kmax), cap_lms = 160, cap_contrib = 1536;
      // (1) range boundaries: a cheap sequential pass - a range closes when the next landmark would exceed the slot,
      //     landmark or contribution budget of one CTA
      std::vector<int> rg_first{0};  // index into pi of the first landmark of each range
      std::vector<long long> rg_contrib;
      long long npairs = 0;
      {
        int cur_slots = 0, cur_lms = 0, prev_end = -1;
        long long cur_contrib = 0;
        for (size_t j = 0; j < pi.size(); ++j) {
          const int l = pi[j], k2 = lm_s0[l + 1] - lm_s0[l];
          const long long pairs = (long long)k2 * (k2 + 1) / 2;
          npairs += pairs;
          // a range is ONE contiguous run of Hpl slots: a wide landmark between two range landmarks closes the range
          const bool gap = lm_s0[l] != prev_end;
          prev_end = lm_s0[l + 1];
          if (cur_lms > 0 && (gap || cur_slots + k2 > cap_slots || cur_lms + 1 > cap_lms || cur_contrib + pairs > cap_contrib)) {
            rg_first.push_back((int)j); rg_contrib.push_back(cur_contrib);
            cur_slots = 0; cur_lms = 0; cur_contrib = 0;
          }
          cur_slots += k2; cur_lms += 1; cur_contrib += pairs;
        }
        if (cur_lms > 0) { rg_first.push_back((int)pi.size()); rg_contrib.push_back(cur_contrib); }
      }
      const int nr = (int)rg_first.size() - 1;
      // (2) everything whose position follows from the boundaries
      std::vector<int> r_slot0(2 * (size_t)nr + 2, 0), r_lm_ptr(nr + 1, 0), r_lm_ids(pi.size()), r_lm_slot(pi.size() + nr), r_seg_ptr(nr + 1, 0);
      std::vector<long long> sc_off(nr + 1, 0);  // contributions of a range start 16-byte aligned (bulk copies)
      for (int r = 0; r < nr; ++r) {
        r_slot0[2 * r] = lm_s0[pi[rg_first[r]]];                  // [first slot, end slot) of the range
        r_slot0[2 * r + 1] = lm_s0[pi[rg_first[r + 1] - 1] + 1];
        r_lm_ptr[r + 1] = rg_first[r + 1];
        sc_off[r + 1] = sc_off[r] + ((rg_contrib[r] + 7) & ~7ll);
      }
      if (sc_off[nr] + 8 > 0x7fffffffll) return fail(c, B200_ERR_UNSUPPORTED, "more than 2^31 Schur contributions on one GPU: shard the landmarks");
      std::vector<unsigned short> sc_a((size_t)sc_off[nr] + 8, 0), sc_b((size_t)sc_off[nr] + 8, 0), sc_l((size_t)sc_off[nr] + 8, 0);  // + 8: bulk copies round sizes up
      // (3) contributions and segments of the ranges, a contiguous block of ranges per thread
      struct Contrib { int t; unsigned short l, a, b; };
      struct RangeOut { std::vector<int> seg_t, seg_cb, seg_ce; };
      const int rparts = range_count((size_t)nr, 64);
      std::vector<RangeOut> outs(rparts);
      std::vector<int> r_nseg(nr, 0);
      parallel_ranges((size_t)nr, rparts, [&](int tid, size_t rb, size_t re) {
        RangeOut& out = outs[tid];
        std::vector<Contrib> rc;
        std::vector<int> pair_t;
        bool have_pairs = false;  // pair_t holds the block indices of the previous landmark's camera list
        for (size_t r = rb; r < re; ++r) {
          const int slot0 = r_slot0[2 * r];
          rc.clear();
          for (int j = rg_first[r]; j < rg_first[r + 1]; ++j) {
            const int l = pi[j], k2 = lm_s0[l + 1] - lm_s0[l], slot = lm_s0[l];
            const int base = slot - slot0, ll = j - rg_first[r];
            if (!same_as_prev[l] || !have_pairs) {  // Hschur block of every camera pair of this list
              pair_t.clear();
              for (int a = 0; a < k2; ++a)
                for (int b2 = a; b2 < k2; ++b2) pair_t.push_back(find_t(c->hpl_row[slot + a], c->hpl_row[slot + b2]));
              have_pairs = true;
            }
            for (int a = 0, z = 0; a < k2; ++a)
              for (int b2 = a; b2 < k2; ++b2, ++z)
                rc.push_back({pair_t[z], (unsigned short)ll, (unsigned short)(base + a), (unsigned short)(base + b2)});
            r_lm_ids[j] = lm_order[l];
            r_lm_slot[j + r] = base;
          }
          r_lm_slot[rg_first[r + 1] + r] = r_slot0[2 * r + 1] - slot0;  // end entry of the range
          std::stable_sort(rc.begin(), rc.end(), [](const Contrib& x, const Contrib& y) { return x.t < y.t; });
          size_t pos = (size_t)sc_off[r];
          int nseg_r = 0;
          for (size_t q = 0; q < rc.size();) {
            size_t e2 = q;
            while (e2 < rc.size() && rc[e2].t == rc[q].t && e2 - q < (size_t)k::kSrSegMax) ++e2;
            out.seg_t.push_back(rc[q].t);
            out.seg_cb.push_back((int)pos);
            for (size_t z = q; z < e2; ++z, ++pos) { sc_a[pos] = rc[z].a; sc_b[pos] = rc[z].b; sc_l[pos] = rc[z].l; }
            out.seg_ce.push_back((int)pos);
            ++nseg_r;
            q = e2;
          }
          r_nseg[r] = nseg_r;
        }
      });
      for (int r = 0; r < nr; ++r) r_seg_ptr[r + 1] = r_seg_ptr[r] + r_nseg[r];
      std::vector<int> seg_t, seg_cb, seg_ce;
      seg_t.reserve(r_seg_ptr[nr]); seg_cb.reserve(r_seg_ptr[nr]); seg_ce.reserve(r_seg_ptr[nr]);
      for (const RangeOut& o : outs) {  // thread blocks are contiguous in range order
        seg_t.insert(seg_t.end(), o.seg_t.begin(), o.seg_t.end());
        seg_cb.insert(seg_cb.end(), o.seg_cb.begin(), o.seg_cb.end());
        seg_ce.insert(seg_ce.end(), o.seg_ce.begin(), o.seg_ce.end());
      }
      const int nseg_ranges = (int)seg_t.size();
      // wide landmarks: one segment per camera pair, numbered behind the segments of the ranges
      std::vector<long long> w_pair0{0}, w_seg0;
      std::vector<int> w_lm, w_slot0, w_deg;
      for (int l : wide) {
        const int k2 = lm_s0[l + 1] - lm_s0[l], slot = lm_s0[l];
        if (k2 > 65535) return fail(c, B200_ERR_UNSUPPORTED, "a landmark observed by more than 65535 cameras");
        const long long pairs = (long long)k2 * (k2 + 1) / 2;
        if ((long long)seg_t.size() + pairs > 0x7fffffffll) return fail(c, B200_ERR_UNSUPPORTED, "more than 2^31 Schur segments on one GPU: shard the landmarks");
        npairs += pairs;
        w_lm.push_back(lm_order[l]); w_slot0.push_back(slot); w_deg.push_back(k2);
        w_seg0.push_back((long long)seg_t.size()); w_pair0.push_back(w_pair0.back() + pairs);
        for (int a = 0; a < k2; ++a)
          for (int b2 = a; b2 < k2; ++b2) seg_t.push_back(find_t(c->hpl_row[slot + a], c->hpl_row[slot + b2]));
      }
      c->sr_nwide = (int)wide.size(); c->sr_wide_pairs = w_pair0.back();
      c->d_sw_pair0.upload(w_pair0, s); c->d_sw_seg0.upload(w_seg0, s);
      c->d_sw_lm.upload(w_lm, s); c->d_sw_slot0.upload(w_slot0, s); c->d_sw_deg.upload(w_deg, s);
      const int nseg = (int)seg_t.size();
      (void)nseg_ranges;
      // per block: its segments in ascending order (= fixed summation order of the finish kernel)
      std::vector<int> tseg_ptr(nT + 1, 0), tseg_idx(nseg);
      for (int sg = 0; sg < nseg; ++sg) tseg_ptr[seg_t[sg] + 1]++;
      for (int t = 0; t < nT; ++t) tseg_ptr[t + 1] += tseg_ptr[t];
      {
        std::vector<int> f(tseg_ptr.begin(), tseg_ptr.end() - 1);
        for (int sg = 0; sg < nseg; ++sg) tseg_idx[f[seg_t[sg]]++] = sg;
      }
      std::vector<unsigned char> t_diag(nT);
      for (int t = 0; t < nT; ++t) t_diag[t] = t_row[t] == t_col[t];
      c->sr_n = nr; c->sr_nseg = nseg; c->sr_ncontrib = npairs;
      c->sr_cap_slots = cap_slots; c->sr_cap_lms = cap_lms; c->sr_cap_contrib = cap_contrib;
      c->d_sr_slot0.upload(r_slot0, s); c->d_sr_lm_ptr.upload(r_lm_ptr, s); c->d_sr_lm_ids.upload(r_lm_ids, s); c->d_sr_lm_slot.upload(r_lm_slot, s);
      c->d_sr_seg_ptr.upload(r_seg_ptr, s); c->d_sr_seg_t.upload(seg_t, s); c->d_sr_seg_cb.upload(seg_cb, s); c->d_sr_seg_ce.upload(seg_ce, s);
      c->d_sr_a.upload(sc_a, s); c->d_sr_b.upload(sc_b, s); c->d_sr_l.upload(sc_l, s);
      c->d_t_diag.upload(t_diag, s); c->d_tseg_ptr.upload(tseg_ptr, s); c->d_tseg_idx.upload(tseg_idx, s);
      c->d_sr_partial.alloc((size_t)std::max(nseg, 1) * k::kSrPartial);
      if (!c->host_only) {
        B200_CUDA(cudaStreamSynchronize(s));
        // always the device maximum (not this context's size): a second context on the same device with a smaller
        // range capacity must not lower the opt-in under this one
        if (schur_range_smem(c) > (size_t)k::kSrMaxDynSmem) return fail(c, B200_ERR_UNSUPPORTED, "Schur range larger than shared memory");
        B200_CUDA(cudaFuncSetAttribute(k::schur_range_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, k::kSrMaxDynSmem));
      }
    }
  STAMP("Schur plan");
    c->d_ev0.upload(e_pt, s); c->d_ev1.upload(e_cam, s); c->d_e_pose.upload(e_pose, s); c->d_e_hpl.upload(e_hpl, s);
    c->d_e_flag.upload(e_first, s);
    c->d_meas.upload(meas, s); c->d_info.upload(info, s);
    c->n_edges_free_lm = lm_eptr[nl];
    {
      // packets of the lane-per-observation linearisation (kernels.cuh: ba_linearize_packets_kernel): consecutive ranks
      // with at most 32 observations (and at most 32 landmarks) in total; a landmark with more is a packet of its own
      std::vector<int> pk{0};
      int cur_e = 0, cur_n = 0;
      for (int i = 0; i < nl; ++i) {
        const int k2 = lm_eptr[i + 1] - lm_eptr[i];
        if (cur_n > 0 && (cur_e + k2 > 32 || cur_n + 1 > 32)) { pk.push_back(i); cur_e = 0; cur_n = 0; }
        cur_e += k2; ++cur_n;
        if (k2 > 32) { pk.push_back(i + 1); cur_e = 0; cur_n = 0; }
      }
      if (cur_n > 0) pk.push_back(nl);
      c->n_lin_packets = (int)pk.size() - 1;
      std::vector<int> pk4((size_t)c->n_lin_packets * 4);   // {first rank, #ranks, first edge, end edge}
      for (int q = 0; q < c->n_lin_packets; ++q) {
        pk4[4 * q] = pk[q]; pk4[4 * q + 1] = pk[q + 1] - pk[q]; pk4[4 * q + 2] = lm_eptr[pk[q]]; pk4[4 * q + 3] = lm_eptr[pk[q + 1]];
      }
      c->d_pk_rank0.upload(pk4, s);
    }
    c->d_partials2.alloc((size_t)ceil_div(std::max(nl, 1), 128) + 8);
    c->d_lm_eptr.upload(lm_eptr, s); c->d_lm_order.upload(lm_order, s); c->d_cam_eptr.upload(cam_eptr, s); c->d_cam_eidx.upload(cam_eidx, s);
    c->d_hpp_diag_block.upload(c->hpp_diag_block, s);
    c->d_t_row.upload(t_row, s); c->d_t_col.upload(t_col, s); c->d_t_hpp.upload(t_hpp, s);
    c->d_Hpp.alloc((size_t)np * 36 + (size_t)c->sizeP);  // [Hpp | b_p staging] contiguous for one all-reduce
    c->d_Hll.alloc((size_t)std::max(nl, 1) * 9); c->d_Hpl.alloc((size_t)std::max(nslot, 1) * 18);
    c->d_Dinv.alloc((size_t)std::max(nl, 1) * k::kDinvStride); c->d_Dinv.zero(s); c->d_Wu.alloc((size_t)std::max(nl, 1) * k::kDinvStride); c->d_Wu.zero(s);
    c->d_Hschur.alloc((size_t)nT * 36 + (size_t)c->sizeP + 8);  // [Hschur | bschur | scalars] contiguous
    if (!c->host_only) B200_CUDA(cudaStreamSynchronize(s));
    bp_colptr = c->hs_colptr; bp_rowidx = c->hs_rowidx;
  }
  STAMP("uploads");
  // ---- symbolic phase of the linear solver
  SymbolicOptions opt;
  // tuning knobs (the adapter exposes them as solver properties; the environment overrides are for experiments)
  if (const char* e = getenv("G2O_B200_PANEL_COLS")) opt.max_panel_cols_scalar = std::max(pd, std::min(72, atoi(e)));
  if (const char* e = getenv("G2O_B200_SUBTREE_FLOPS")) opt.subtree_min_flops = atof(e);
  if (const char* e = getenv("G2O_B200_RELAX")) opt.relax = atoi(e) != 0;
  if (const char* e = getenv("G2O_B200_CHAIN")) opt.chain = atoi(e) != 0;   // 0: every supernode through the dataflow kernel
  if (const char* e = getenv("G2O_B200_GROUP_ITEMS")) opt.group_items = std::max(1, atoi(e));
  if (const char* e = getenv("G2O_B200_SORT_ITEMS")) opt.sort_items_by_level = atoi(e) != 0;
  if (const char* e = getenv("G2O_B200_RELAX_FRAC")) opt.relax_frac = atof(e);
  if (const char* e = getenv("G2O_B200_WIDE_TILES")) opt.wide_tiles = atoi(e);
  if (const char* e = getenv("G2O_B200_SUBTREE_MAX_FLOPS")) opt.subtree_max_flops = atof(e);
  if (const char* e = getenv("G2O_B200_SPLIT_LATE")) opt.split_late_items = atoi(e) != 0;
  if (const char* e = getenv("G2O_B200_GROUP_SLACK")) opt.group_slack = atoi(e);
  if (const char* e = getenv("G2O_B200_GROUPS_ASAP")) opt.groups_asap = atoi(e) != 0;   // -1 auto (by flops), 0 off, 1 on
  opt.nd_levels = c->nd_levels;
  if (const char* e = getenv("G2O_B200_ND_LEVELS")) opt.nd_levels = std::max(0, atoi(e));
  c->chol.analyze(np, pd, bp_colptr.data(), bp_rowidx.data(), opt, s);
  c->chol.set_diagonal_extra(var_lm ? c->d_pad_diag.p : nullptr);
  c->pcg.init();   // a new structure: LinearSolverPCG::init() (BlockSolver::init -> _linearSolver->init())
  if (c->linear_solver == 1) {
    if (!c->pcg.analyze(np, pd, bp_colptr.data(), bp_rowidx.data(), s, &c->err)) return B200_ERR_INVALID;
    if (!has_lm) c->d_pcg_A.alloc((size_t)c->n_hpp * pd * pd);
  }
  STAMP("symbolic (ordering + plan)");
  c->structured = true;
  upload_robust(c);
  c->state_chi2_valid = false;  // new graph / new estimates on the device
  c->backup_depth = 0;
  c->time_symbolic = wall() - t0;
  return B200_OK;
}

double* bschur_ptr(b200_ctx* c) { return c->d_Hschur.p + (size_t)c->n_hs * 36; }

// ------------------------------------------------------------------------------------------------
// per-iteration phases (all asynchronous on c->stream)
// ------------------------------------------------------------------------------------------------
void enqueue_chi2(b200_ctx* c) {  // result -> d_scalars[0]
  PhaseTimer pt(c, PH_ERRORS);
  const int E = c->nE;
  int nb = ceil_div(E, 256);
  cudaStream_t s = c->stream;
  if (E > 0) {
    if (c->edge_kind == B200_EDGE_SE2)
      k::pg_chi2_kernel<0><<<nb, 256, 0, s>>>(E, c->d_ev0.p, c->d_ev1.p, c->d_pose_est.p, c->d_meas.p, c->d_info.p, c->robust, c->d_partials.p);
    else if (c->edge_kind == B200_EDGE_SE3)
      k::pg_chi2_kernel<1><<<nb, 256, 0, s>>>(E, c->d_ev0.p, c->d_ev1.p, c->d_pose_est.p, c->d_meas.p, c->d_info.p, c->robust, c->d_partials.p);
    else
      BA_MODEL_LAUNCH(c, ba_chi2_kernel, <<<nb, 256, 0, s>>>(E, c->d_ev0.p, c->d_ev1.p, c->d_lm_est.p, c->d_cam_der.p, c->d_meas.p, c->d_info.p, c->robust, c->d_partials.p, 0));
    c->lc.n++;
  }
  if (c->var_lm) {  // the pose-landmark edges: their block sums follow the pose-pose edges'
    const int LE = c->nLE, nb2 = ceil_div(LE, 256);
    k::SensorOffset off;
    memcpy(off.m, c->sensor_offset, sizeof(off.m));
    if (c->l_edge_kind == B200_EDGE_SE2_XY)
      k::pl_chi2_kernel<0><<<nb2, 256, 0, s>>>(LE, c->d_lev0.p, c->d_lev1.p, c->d_pose_est.p, c->d_lm_est.p, c->d_lmeas.p, c->d_linfo.p, off, c->robust_l, c->d_partials.p + nb);
    else
      k::pl_chi2_kernel<1><<<nb2, 256, 0, s>>>(LE, c->d_lev0.p, c->d_lev1.p, c->d_pose_est.p, c->d_lm_est.p, c->d_lmeas.p, c->d_linfo.p, off, c->robust_l, c->d_partials.p + nb);
    nb += nb2;
    c->lc.n++;
  }
  k::reduce_partials_kernel<<<1, 1024, 0, s>>>(c->d_partials.p, nb, c->d_scalars.p + 0);
  c->lc.n++;
  B200_CUDA(cudaGetLastError());
}

bool sharded(const b200_ctx* c) { return c->world > 1 && (c->nccl.active() || c->allreduce); }

// in-place all-reduce (op 0 = sum, 1 = max) of device doubles, ordered on the solver stream: ncclAllReduce on the native
// communicator (b200_comm_init), else the host-supplied callback (b200_set_allreduce)
int allreduce_dev(b200_ctx* c, double* p, long long count, int op = 0) {
  if (!sharded(c)) return 0;
  PhaseTimer pt(c, PH_COLLECTIVE);
  if (c->nccl.active()) {
    const int rc = op == 0 ? c->nccl.allreduce_sum(p, count, c->stream, &c->err) : c->nccl.allreduce_max(p, count, c->stream, &c->err);
    c->lc.n++;  // one NCCL kernel
    return rc ? B200_ERR_COLLECTIVE : 0;
  }
  int rc = c->allreduce(p, count, op, (void*)c->stream, c->allreduce_user);
  if (rc != 0) { c->err = "all-reduce callback failed"; return B200_ERR_COLLECTIVE; }
  return 0;
}

int enqueue_build_system(b200_ctx* c) {
  cudaStream_t s = c->stream;
  const int E = c->nE, np = c->np;
  if (!c->schur) {
    k::SensorOffset off;
    memcpy(off.m, c->sensor_offset, sizeof(off.m));
    if (c->edge_kind == B200_EDGE_SE2) {
      { PhaseTimer pt(c, PH_LINEARIZE);
      if (E > 0) k::pg_linearize_kernel<0><<<ceil_div(E, 128), 128, 0, s>>>(E, c->d_ev0.p, c->d_ev1.p, c->d_pose_est.p, c->d_meas.p, c->d_info.p, c->d_e_flag.p, c->robust, c->d_stage.p);
      if (c->var_lm) { k::pl_linearize_kernel<0><<<ceil_div(c->nLE, 128), 128, 0, s>>>(c->nLE, E, c->d_lev0.p, c->d_lev1.p, c->d_pose_est.p, c->d_lm_est.p, c->d_lmeas.p, c->d_linfo.p, c->d_e_flag.p, off, c->robust_l, c->d_stage.p); c->lc.n++; } }
      PhaseTimer pt(c, PH_GATHER);
      k::gather_segments_kernel<3, 9><<<ceil_div((long long)c->n_hpp * 9, 256), 256, 0, s>>>(c->n_hpp, c->d_hsrc_ptr.p, c->d_hsrc_id.p, c->d_stage.p, c->d_Hpp.p);
      k::gather_segments_kernel<3, 3><<<ceil_div((long long)np * 3, 256), 256, 0, s>>>(np, c->d_bsrc_ptr.p, c->d_bsrc_id.p, c->d_stage.p, c->d_b.p);
    } else {
      { PhaseTimer pt(c, PH_LINEARIZE);
      if (E > 0) k::pg_linearize_kernel<1><<<ceil_div(E, 128), 128, 0, s>>>(E, c->d_ev0.p, c->d_ev1.p, c->d_pose_est.p, c->d_meas.p, c->d_info.p, c->d_e_flag.p, c->robust, c->d_stage.p);
      if (c->var_lm) { k::pl_linearize_kernel<1><<<ceil_div(c->nLE, 128), 128, 0, s>>>(c->nLE, E, c->d_lev0.p, c->d_lev1.p, c->d_pose_est.p, c->d_lm_est.p, c->d_lmeas.p, c->d_linfo.p, c->d_e_flag.p, off, c->robust_l, c->d_stage.p); c->lc.n++; } }
      PhaseTimer pt(c, PH_GATHER);
      k::gather_segments_kernel<6, 36><<<ceil_div((long long)c->n_hpp * 36, 256), 256, 0, s>>>(c->n_hpp, c->d_hsrc_ptr.p, c->d_hsrc_id.p, c->d_stage.p, c->d_Hpp.p);
      k::gather_segments_kernel<6, 6><<<ceil_div((long long)np * 6, 256), 256, 0, s>>>(np, c->d_bsrc_ptr.p, c->d_bsrc_id.p, c->d_stage.p, c->d_b.p);
    }
    c->lc.n += 3;
  } else {
    double* b_p_stage = c->d_Hpp.p + (size_t)np * 36;
    // the per-camera pass reads the same estimates and writes disjoint outputs: it runs on the side stream beside the
    // per-landmark pass (event fork / join; per-phase profiling keeps the serial order so that its intervals stay meaningful)
    const bool fork = c->overlap_linearize && !c->prof.on && c->nl > 0 && c->stream2;
    cudaStream_t sc = fork ? c->stream2 : s;
    if (fork) {
      B200_CUDA(cudaEventRecord(c->ev_fork, s));
      B200_CUDA(cudaStreamWaitEvent(sc, c->ev_fork, 0));
    }
    if (c->nl > 0) {
      PhaseTimer pt(c, PH_LINEARIZE);
      if (c->lin_packets) {
#define LIN_PACKETS_LAUNCH(M, B) k::ba_linearize_packets_kernel<M, B><<<ceil_div(c->n_lin_packets, k::kLinWarps), 32 * k::kLinWarps, 0, s>>>(c->n_lin_packets, reinterpret_cast<const int4*>(c->d_pk_rank0.p), c->d_lm_eptr.p, c->d_lm_order.p, c->d_ev0.p, c->d_ev1.p, c->d_e_hpl.p, c->d_e_flag.p, c->d_lm_est.p, c->d_pose_est.p, c->d_cam_der.p, c->d_meas.p, c->d_info.p, E, c->robust, c->d_Hll.p, c->d_Hpl.p, c->d_b.p + c->sizeP)
        if (c->cam_model == 0) { if (c->lin_minb == 6) LIN_PACKETS_LAUNCH(0, 6); else if (c->lin_minb == 5) LIN_PACKETS_LAUNCH(0, 5); else LIN_PACKETS_LAUNCH(0, 4); }
        else { if (c->lin_minb == 6) LIN_PACKETS_LAUNCH(1, 6); else if (c->lin_minb == 5) LIN_PACKETS_LAUNCH(1, 5); else LIN_PACKETS_LAUNCH(1, 4); }
#undef LIN_PACKETS_LAUNCH
      }
      else BA_MODEL_LAUNCH(c, ba_linearize_points_kernel, <<<ceil_div(c->nl, 128), 128, 0, s>>>(c->nl, c->d_lm_eptr.p, c->d_lm_order.p, c->d_lm_vertex.p, c->d_ev1.p, c->d_e_hpl.p, c->d_e_flag.p, c->d_lm_est.p, c->d_pose_est.p, c->d_cam_der.p, c->d_meas.p, c->d_info.p, E, c->robust, c->d_Hll.p, c->d_Hpl.p, c->d_b.p + c->sizeP));
      c->lc.n++;
    }
    { PhaseTimer pt(c, PH_LINEARIZE_CAMS);
#define LIN_CAMS_LAUNCH(M, B) k::ba_linearize_cams_kernel<M, B><<<np, 128, 0, sc>>>(c->d_cam_eptr.p, c->d_cam_eidx.p, c->d_pose_vertex.p, c->d_ev0.p, c->d_lm_est.p, c->d_pose_est.p, c->d_cam_der.p, c->d_meas.p, c->d_info.p, E, c->robust, c->d_hpp_diag_block.p, c->d_Hpp.p, b_p_stage)
    if (c->cam_model == 0) { if (c->cams_minb == 6) LIN_CAMS_LAUNCH(0, 6); else if (c->cams_minb == 4) LIN_CAMS_LAUNCH(0, 4); else LIN_CAMS_LAUNCH(0, 3); }
    else { if (c->cams_minb == 6) LIN_CAMS_LAUNCH(1, 6); else if (c->cams_minb == 4) LIN_CAMS_LAUNCH(1, 4); else LIN_CAMS_LAUNCH(1, 3); }
#undef LIN_CAMS_LAUNCH
    }
    c->lc.n++;
    B200_CUDA(cudaGetLastError());
    if (fork) {
      B200_CUDA(cudaEventRecord(c->ev_join, sc));
      B200_CUDA(cudaStreamWaitEvent(s, c->ev_join, 0));
    }
    // sharded: Hpp and b_p stay this rank's PARTIAL sums (its own observations); they enter the reduced system through
    // schur_finish_kernel and are summed by the one all-reduce of [Hschur | bschur] (enqueue_solve)
    B200_CUDA(cudaMemcpyAsync(c->d_b.p, b_p_stage, (size_t)c->sizeP * sizeof(double), cudaMemcpyDeviceToDevice, s));
  }
  B200_CUDA(cudaGetLastError());
  return 0;
}

int enqueue_max_diag(b200_ctx* c) {  // d_scalars[2] = max_j |H_jj| over poses and landmarks
  cudaStream_t s = c->stream;
  const int np = c->np, nl = c->nl;
  if (sharded(c) && c->schur) {
    // landmark-sharded: the camera diagonal is a sum over the ranks, the landmark diagonals live on their owners.
    // One sum all-reduce over [Hpp diagonal partial sums (6 np) | per-rank landmark maximum in slot `rank`], then the
    // maximum of the reduced vector (iteration 0 only)
    const int n = c->sizeP + c->world;
    c->d_comm_diag.alloc((size_t)n);
    B200_CUDA(cudaMemsetAsync(c->d_comm_diag.p, 0, (size_t)n * sizeof(double), s));
    k::extract_diag_kernel<6><<<ceil_div((long long)np * 6, 256), 256, 0, s>>>(np, c->d_hpp_diag_block.p, c->d_Hpp.p, c->d_comm_diag.p);
    c->lc.n++;
    if (nl > 0) {
      const int nb = ceil_div((long long)nl * 3, 256);
      k::max_diag_kernel<3><<<nb, 256, 0, s>>>(nl, nullptr, c->d_Hll.p, c->d_partials.p);
      k::reduce_max_kernel<<<1, 1024, 0, s>>>(c->d_partials.p, nb, 1.0, c->d_comm_diag.p + c->sizeP + c->rank, 0);
      c->lc.n += 2;
    }
    B200_CUDA(cudaGetLastError());
    if (int rc = allreduce_dev(c, c->d_comm_diag.p, n)) return rc;
    const int nb = ceil_div(n, 256);
    k::absmax_kernel<<<nb, 256, 0, s>>>(n, c->d_comm_diag.p, c->d_partials.p);
    k::reduce_max_kernel<<<1, 1024, 0, s>>>(c->d_partials.p, nb, 1.0, c->d_scalars.p + 2, 0);
    c->lc.n += 2;
    B200_CUDA(cudaGetLastError());
    return 0;
  }
  int nb = ceil_div((long long)np * c->pd, 256);
  if (c->pd == 3) k::max_diag_kernel<3><<<nb, 256, 0, s>>>(np, c->d_hpp_diag_block.p, c->d_Hpp.p, c->d_partials.p);
  else k::max_diag_kernel<6><<<nb, 256, 0, s>>>(np, c->d_hpp_diag_block.p, c->d_Hpp.p, c->d_partials.p);
  k::reduce_max_kernel<<<1, 1024, 0, s>>>(c->d_partials.p, nb, 1.0, c->d_scalars.p + 2, 0);
  c->lc.n += 2;
  if (c->schur && nl > 0) {
    nb = ceil_div((long long)nl * 3, 256);
    k::max_diag_kernel<3><<<nb, 256, 0, s>>>(nl, nullptr, c->d_Hll.p, c->d_partials.p);
    k::reduce_max_kernel<<<1, 1024, 0, s>>>(c->d_partials.p, nb, 1.0, c->d_scalars.p + 2, 1);
    c->lc.n += 2;
  }
  B200_CUDA(cudaGetLastError());
  return 0;
}


// Solver::solve with the lambda currently stored at d_scalars[3]
int enqueue_solve(b200_ctx* c, bool skip_backsub = false) {
  cudaStream_t s = c->stream;
  const double* d_lambda = c->d_scalars.p + 3;
  if (!c->schur && c->linear_solver == 1) {
    // LinearSolverPCG on Hpp + lambda I: the damped matrix is materialised once (the Cholesky adds lambda while it scatters)
    PhaseTimer pt(c, PH_FACTOR);
    B200_CUDA(cudaMemcpyAsync(c->d_pcg_A.p, c->d_Hpp.p, (size_t)c->n_hpp * c->pd * c->pd * sizeof(double), cudaMemcpyDeviceToDevice, s));
    if (c->pd == 3) k::add_block_diagonal_kernel<3><<<ceil_div((long long)c->np * 3, 256), 256, 0, s>>>(c->np, c->d_hpp_diag_block.p, d_lambda, c->var_lm ? c->d_pad_diag.p : nullptr, c->d_pcg_A.p);
    else k::add_block_diagonal_kernel<6><<<ceil_div((long long)c->np * 6, 256), 256, 0, s>>>(c->np, c->d_hpp_diag_block.p, d_lambda, c->var_lm ? c->d_pad_diag.p : nullptr, c->d_pcg_A.p);
    c->lc.n++;
    B200_CUDA(cudaMemsetAsync(c->chol.status_ptr(), 0, sizeof(int), s));
    const int rc = c->pcg.solve(c->d_pcg_A.p, c->d_b.p, c->pcg_tolerance, c->pcg_absolute, c->pcg_max_iterations, s, &c->lc, &c->pcg_last_iterations, nullptr);
    if (rc) { const int one = 1; B200_CUDA(cudaMemcpyAsync(c->chol.status_ptr(), &one, sizeof(int), cudaMemcpyHostToDevice, s)); }
    else B200_CUDA(cudaMemcpyAsync(c->d_x.p, c->pcg.x(), (size_t)c->sizeP * sizeof(double), cudaMemcpyDeviceToDevice, s));
    B200_CUDA(cudaGetLastError());
    return 0;
  } else if (!c->schur) {
    { PhaseTimer pt(c, PH_FACTOR); c->chol.factor(c->d_Hpp.p, d_lambda, c->d_b.p, s, &c->lc, &c->prof); }
    { PhaseTimer pt(c, PH_TRISOLVE); c->chol.solve(c->d_b.p, c->d_x.p, s, &c->lc, &c->prof); }
    return 0;
  }
  {
    if (c->nl > 0) {
      PhaseTimer pt(c, PH_SCHUR_INV);
      k::schur_landmark_inverse_kernel<<<ceil_div(c->nl, 128), 128, 0, s>>>(c->nl, c->d_Hll.p, c->d_b.p + c->sizeP, d_lambda, c->d_Dinv.p, c->d_Wu.p);
      c->lc.n++;
    }
    const double lambda_scale = (sharded(c) && c->rank != 0) ? 0.0 : 1.0;  // the lambda term enters the sum once
    {
      PhaseTimer pt(c, PH_SCHUR);
      if (c->sr_n > 0) {
        k::SchurRanges R{c->d_sr_slot0.p, c->d_sr_lm_ptr.p, c->d_sr_lm_ids.p, c->d_sr_lm_slot.p, c->d_sr_seg_ptr.p, c->d_sr_seg_t.p, c->d_sr_seg_cb.p, c->d_sr_seg_ce.p,
                         c->d_sr_a.p, c->d_sr_b.p, c->d_sr_l.p, c->d_t_diag.p, c->sr_cap_slots, c->sr_cap_lms, c->sr_cap_contrib};
        k::schur_range_kernel<<<c->sr_n, k::kSrThreads, schur_range_smem(c), s>>>(R, c->d_Hpl.p, c->d_Wu.p, c->d_sr_partial.p);
        c->lc.n++;
      }
      if (c->sr_nwide > 0) {  // landmarks too wide for a range CTA: one thread per camera pair
        k::SchurWide Wd{c->sr_nwide, c->d_sw_pair0.p, c->d_sw_lm.p, c->d_sw_slot0.p, c->d_sw_deg.p, c->d_sw_seg0.p};
        k::schur_wide_kernel<<<ceil_div(c->sr_wide_pairs, 128), 128, 0, s>>>(Wd, c->sr_wide_pairs, c->d_Hpl.p, c->d_Wu.p, c->d_sr_partial.p);
        c->lc.n++;
      }
    }
    {
      PhaseTimer pt(c, PH_SCHUR_FINISH);
      k::schur_finish_kernel<<<ceil_div(c->n_hs, 4), 256, 0, s>>>(c->n_hs, c->d_t_row.p, c->d_t_col.p, c->d_t_hpp.p, c->d_tseg_ptr.p, c->d_tseg_idx.p,
                                                               c->d_sr_partial.p, c->d_Hpp.p, c->d_b.p, d_lambda, 1.0, lambda_scale, c->d_Hschur.p, bschur_ptr(c));
      c->lc.n++;
    }
    B200_CUDA(cudaGetLastError());
    if (sharded(c)) {
      // THE collective of the trial: [Hschur | bschur | chi2 of the state before the trial] in one ncclAllReduce
      double* tail = bschur_ptr(c) + c->sizeP;
      B200_CUDA(cudaMemcpyAsync(tail, c->d_scalars.p + 7, sizeof(double), cudaMemcpyDeviceToDevice, s));
      int rc = allreduce_dev(c, c->d_Hschur.p, (long long)c->n_hs * 36 + c->sizeP + 1);
      if (rc) return rc;
      B200_CUDA(cudaMemcpyAsync(c->d_scalars.p + 5, tail, sizeof(double), cudaMemcpyDeviceToDevice, s));
    }
  }
  if (c->linear_solver == 1) {  // LinearSolverPCG on the reduced camera system (lambda is already on its diagonal)
    PhaseTimer pt(c, PH_FACTOR);
    B200_CUDA(cudaMemsetAsync(c->chol.status_ptr(), 0, sizeof(int), s));
    const int rc = c->pcg.solve(c->d_Hschur.p, bschur_ptr(c), c->pcg_tolerance, c->pcg_absolute, c->pcg_max_iterations, s, &c->lc, &c->pcg_last_iterations, nullptr);
    if (rc) { const int one = 1; B200_CUDA(cudaMemcpyAsync(c->chol.status_ptr(), &one, sizeof(int), cudaMemcpyHostToDevice, s)); }
    else B200_CUDA(cudaMemcpyAsync(c->d_x.p, c->pcg.x(), (size_t)c->sizeP * sizeof(double), cudaMemcpyDeviceToDevice, s));
  } else {
  { PhaseTimer pt(c, PH_FACTOR); c->chol.factor(c->d_Hschur.p, nullptr, bschur_ptr(c), s, &c->lc, &c->prof); }
  { PhaseTimer pt(c, PH_TRISOLVE); c->chol.solve(bschur_ptr(c), c->d_x.p, s, &c->lc, &c->prof); }
  }
  if (c->nl > 0 && !skip_backsub) {
    PhaseTimer pt(c, PH_BACKSUB);
    // on a failed factorisation the reference returns before touching the landmark part of x; the pose part
    // is left untouched by chol.solve, the landmark part is recomputed from the stale pose part (harmless:
    // LM discards the step, GN reports Fail).
    k::ba_backsub_kernel<<<ceil_div(c->nl, 128), 128, 0, s>>>(c->nl, c->d_lm_eptr.p, c->d_lm_order.p, c->d_e_hpl.p, c->d_e_pose.p, c->d_Hpl.p, c->d_Dinv.p, c->d_b.p + c->sizeP, c->d_x.p, c->d_x.p + c->sizeP);
    c->lc.n++;
    B200_CUDA(cudaGetLastError());
  }
  return 0;
}

void enqueue_update(b200_ctx* c) {
  PhaseTimer pt(c, PH_UPDATE);
  cudaStream_t s = c->stream;
  const int n = c->n_pose_v;
  if (c->pose_kind == B200_VERTEX_SE2) k::oplus_se2_kernel<<<ceil_div(n, 128), 128, 0, s>>>(n, c->d_pose_hidx.p, c->d_x.p, c->d_pose_est.p);
  else if (c->pose_kind == B200_VERTEX_SE3) {
    int orth = 0;
    if (++c->num_oplus_calls > 1000) { c->num_oplus_calls = 0; orth = 1; }  // VertexSE3::orthogonalizeAfter (vertex_se3.h:56,111)
    k::oplus_se3_kernel<<<ceil_div(n, 128), 128, 0, s>>>(n, c->d_pose_hidx.p, c->d_x.p, c->d_pose_est.p, orth);
  } else {
    BA_MODEL_LAUNCH(c, oplus_cam_kernel, <<<ceil_div(n, 128), 128, 0, s>>>(n, c->d_pose_hidx.p, c->d_x.p, c->d_pose_est.p, c->d_cam_der.p));
  }
  c->lc.n++;
  if (c->var_lm && c->n_lm_v > 0) {
    if (c->lm_kind == B200_VERTEX_XY) k::oplus_point_kernel<3, 2><<<ceil_div(c->n_lm_v, 128), 128, 0, s>>>(c->n_lm_v, c->d_lm_lidx.p, c->d_x.p, c->d_lm_est.p);
    else k::oplus_point_kernel<6, 3><<<ceil_div(c->n_lm_v, 128), 128, 0, s>>>(c->n_lm_v, c->d_lm_lidx.p, c->d_x.p, c->d_lm_est.p);
    c->lc.n++;
  }
  if (c->schur && c->n_lm_v > 0) {
    k::oplus_xyz_kernel<<<ceil_div(c->n_lm_v, 128), 128, 0, s>>>(c->n_lm_v, c->d_lm_lidx.p, c->d_x.p + c->sizeP, c->d_lm_est.p);
    c->lc.n++;
  }
  B200_CUDA(cudaGetLastError());
}

// sum_j x_j (lambda x_j + b_j): pose part -> d_scalars[4] (identical on every rank), landmark part -> d_scalars[1]
// (a partial sum when landmarks are sharded; all-reduced together with chi2 in d_scalars[0])
void enqueue_scale(b200_ctx* c) {
  PhaseTimer pt(c, PH_SCALE);
  cudaStream_t s = c->stream;
  int nb = ceil_div(c->sizeP, 256);
  // sharded: b_p is this rank's partial sum, so sum_j x_j b_j is partial as well; the lambda x_j^2 term is rank 0's
  const double* pose_lambda = (sharded(c) && c->rank != 0) ? c->d_scalars.p + 6 /* constant 0 */ : c->d_scalars.p + 3;
  k::lm_scale_kernel<<<nb, 256, 0, s>>>(c->sizeP, c->d_x.p, c->d_b.p, pose_lambda, c->d_partials.p);
  k::reduce_partials_kernel<<<1, 1024, 0, s>>>(c->d_partials.p, nb, c->d_scalars.p + 4);
  c->lc.n += 2;
  if (c->sizeL > 0) {
    nb = ceil_div(c->sizeL, 256);
    k::lm_scale_kernel<<<nb, 256, 0, s>>>(c->sizeL, c->d_x.p + c->sizeP, c->d_b.p + c->sizeP, c->d_scalars.p + 3, c->d_partials.p);
    k::reduce_partials_kernel<<<1, 1024, 0, s>>>(c->d_partials.p, nb, c->d_scalars.p + 1);
    c->lc.n += 2;
  } else {
    B200_CUDA(cudaMemsetAsync(c->d_scalars.p + 1, 0, sizeof(double), s));
  }
  if (sharded(c)) { k::fold_scale_kernel<<<1, 1, 0, s>>>(c->d_scalars.p); c->lc.n++; }  // [1] += [4]: one slot to reduce
  B200_CUDA(cudaGetLastError());
}

void do_push(b200_ctx* c) {
  cudaStream_t s = c->stream;
  B200_CUDA(cudaMemcpyAsync(c->d_pose_bak.p, c->d_pose_est.p, c->d_pose_est.n * sizeof(double), cudaMemcpyDeviceToDevice, s));
  if (c->pose_kind == B200_VERTEX_CAM) B200_CUDA(cudaMemcpyAsync(c->d_cam_der_bak.p, c->d_cam_der.p, c->d_cam_der.n * sizeof(double), cudaMemcpyDeviceToDevice, s));
  if ((c->schur || c->var_lm) && c->n_lm_v > 0) B200_CUDA(cudaMemcpyAsync(c->d_lm_bak.p, c->d_lm_est.p, c->d_lm_est.n * sizeof(double), cudaMemcpyDeviceToDevice, s));
}
void do_pop(b200_ctx* c) {
  cudaStream_t s = c->stream;
  B200_CUDA(cudaMemcpyAsync(c->d_pose_est.p, c->d_pose_bak.p, c->d_pose_est.n * sizeof(double), cudaMemcpyDeviceToDevice, s));
  if (c->pose_kind == B200_VERTEX_CAM) B200_CUDA(cudaMemcpyAsync(c->d_cam_der.p, c->d_cam_der_bak.p, c->d_cam_der.n * sizeof(double), cudaMemcpyDeviceToDevice, s));
  if ((c->schur || c->var_lm) && c->n_lm_v > 0) B200_CUDA(cudaMemcpyAsync(c->d_lm_est.p, c->d_lm_bak.p, c->d_lm_est.n * sizeof(double), cudaMemcpyDeviceToDevice, s));
}

// BA trial tail, fused (kernels.cuh: ba_backsub_update_kernel): cameras first, then ONE pass over the landmarks' observations
// does back-substitution, landmark update, chi2 of the new state and the landmark part of the LM scale
bool fused_tail(const b200_ctx* c) { return c->fuse_tail && c->schur && c->nl > 0; }
void enqueue_fused_tail(b200_ctx* c) {
  cudaStream_t s = c->stream;
  const int n = c->n_pose_v, E = c->nE;
  { PhaseTimer pt(c, PH_UPDATE);
    BA_MODEL_LAUNCH(c, oplus_cam_kernel, <<<ceil_div(n, 128), 128, 0, s>>>(n, c->d_pose_hidx.p, c->d_x.p, c->d_pose_est.p, c->d_cam_der.p));
    c->lc.n++; }
  const int nb = ceil_div(c->nl, 128);
  { PhaseTimer pt(c, PH_BACKSUB);
    BA_MODEL_LAUNCH(c, ba_backsub_update_kernel, <<<nb, 128, 0, s>>>(c->nl, c->d_lm_eptr.p, c->d_lm_order.p, c->d_lm_vertex.p, c->d_ev1.p, c->d_e_hpl.p, c->d_e_pose.p,
        c->d_Hpl.p, c->d_Dinv.p, c->d_b.p + c->sizeP, c->d_x.p, c->d_x.p + c->sizeP, c->d_scalars.p + 3, c->d_lm_est.p, c->d_cam_der.p,
        c->d_meas.p, c->d_info.p, E, c->robust, c->d_partials.p, c->d_partials2.p));
    c->lc.n++; }
  int nchi = nb;
  if (c->n_edges_free_lm < E) {  // observations of fixed points: the plain error kernel on the tail of the edge list
    PhaseTimer pt(c, PH_ERRORS);
    const int nt = ceil_div(E - c->n_edges_free_lm, 256);
    BA_MODEL_LAUNCH(c, ba_chi2_kernel, <<<nt, 256, 0, s>>>(E, c->d_ev0.p, c->d_ev1.p, c->d_lm_est.p, c->d_cam_der.p, c->d_meas.p, c->d_info.p, c->robust, c->d_partials.p + nb, c->n_edges_free_lm));
    nchi += nt;
    c->lc.n++;
  }
  { PhaseTimer pt(c, PH_SCALE);
    k::reduce_partials_kernel<<<1, 1024, 0, s>>>(c->d_partials.p, nchi, c->d_scalars.p + 0);
    k::reduce_partials_kernel<<<1, 1024, 0, s>>>(c->d_partials2.p, nb, c->d_scalars.p + 1);
    // pose part of the scale (sharded: partial, see enqueue_scale)
    const int np = ceil_div(c->sizeP, 256);
    const double* pose_lambda = (sharded(c) && c->rank != 0) ? c->d_scalars.p + 6 : c->d_scalars.p + 3;
    k::lm_scale_kernel<<<np, 256, 0, s>>>(c->sizeP, c->d_x.p, c->d_b.p, pose_lambda, c->d_partials.p);
    k::reduce_partials_kernel<<<1, 1024, 0, s>>>(c->d_partials.p, np, c->d_scalars.p + 4);
    c->lc.n += 4;
    if (sharded(c)) { k::fold_scale_kernel<<<1, 1, 0, s>>>(c->d_scalars.p); c->lc.n++; } }
  B200_CUDA(cudaGetLastError());
}

// sharded runs: chi2 and the landmark part of the LM scale are partial sums -> one tiny all-reduce
int reduce_trial_scalars(b200_ctx* c, int count = 1) {
  if (!sharded(c)) return 0;
  // [0] chi2 partial, [1] landmark part of the LM scale; the pose part (slot 4) is identical on every rank
  return allreduce_dev(c, c->d_scalars.p + 0, count);
}


// ---- CUDA graph replay of the launch-bound sequences (about 150-250 small kernels each)
// sharded: the native ncclAllReduce is captured like any kernel; the host callback cannot be
bool graphs_usable(b200_ctx* c) { return c->use_graphs && !c->prof.on && c->linear_solver == 0 && (!sharded(c) || (c->nccl.active() && c->comm_warm)); }

template <typename F>
int run_captured(b200_ctx* c, cudaGraphExec_t* exec, long long* launches, F&& enqueue) {
  if (!*exec) {
    const long long before = c->lc.n;
    cudaGraph_t graph = nullptr;
    B200_CUDA(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    int rc = 0;
    try {
      rc = enqueue();
    } catch (...) {
      cudaStreamEndCapture(c->stream, &graph);
      if (graph) cudaGraphDestroy(graph);
      throw;
    }
    B200_CUDA(cudaStreamEndCapture(c->stream, &graph));
    if (rc) { cudaGraphDestroy(graph); return rc; }
    cudaError_t e = cudaGraphInstantiate(exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { *exec = nullptr; throw CudaError{e, "cudaGraphInstantiate", __FILE__, __LINE__}; }
    *launches = c->lc.n - before;
    c->lc.n = before;  // counted per replay below
  }
  B200_CUDA(cudaGraphLaunch(*exec, c->stream));
  c->lc.n += *launches;
  return 0;
}

// errors + chi2 + buildSystem of the current state; the error pass is skipped when the chi2 of this very state is known
int run_prologue(b200_ctx* c, bool need_chi2) {
  auto body = [&]() -> int {
    if (need_chi2) {
      enqueue_chi2(c);
      // sharded: this rank's part of chi2 rides in the first trial's all-reduce (slot 7 -> tail of [Hschur | bschur])
      if (sharded(c)) B200_CUDA(cudaMemcpyAsync(c->d_scalars.p + 7, c->d_scalars.p + 0, sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    }
    return enqueue_build_system(c);
  };
  if (!graphs_usable(c)) return body();
  if (need_chi2) return run_captured(c, &c->graph_prologue, &c->graph_prologue_launches, body);
  return run_captured(c, &c->graph_build, &c->graph_build_launches, body);
}

// one LM trial: lambda -> solve -> update -> errors + chi2 -> scale
int run_trial(b200_ctx* c) {
  c->h_scalars[8] = c->lambda;
  bool orth_now = false;
  if (c->pose_kind == B200_VERTEX_SE3 && c->num_oplus_calls + 1 > 1000) orth_now = true;  // rare: take the plain path
  const bool fuse = fused_tail(c);
  if (!graphs_usable(c) || orth_now) {
    set_device_lambda(c, c->lambda);
    int rc = enqueue_solve(c, fuse);
    if (rc) return rc;
    if (fuse) enqueue_fused_tail(c);
    else { enqueue_update(c); enqueue_chi2(c); enqueue_scale(c); }
    rc = reduce_trial_scalars(c, 2);
    if (!rc && sharded(c)) c->comm_warm = true;
    return rc;
  }
  const int saved = c->num_oplus_calls;
  auto body = [&]() -> int {
    B200_CUDA(cudaMemcpyAsync(c->d_scalars.p + 3, c->h_scalars + 8, sizeof(double), cudaMemcpyHostToDevice, c->stream));
    int rc = enqueue_solve(c, fuse);
    if (rc) return rc;
    if (fuse) enqueue_fused_tail(c);
    else { enqueue_update(c); enqueue_chi2(c); enqueue_scale(c); }
    return reduce_trial_scalars(c, 2);  // sharded: chi2 of the new state + LM scale, 2 doubles
  };
  int rc = run_captured(c, &c->graph_trial, &c->graph_trial_launches, body);
  c->num_oplus_calls = saved + 1;  // the captured enqueue_update only counts once
  return rc;
}

}  // namespace

// ================================================================================================
// C-ABI
// ================================================================================================
extern "C" {

const char* b200_version(void) { return "g2o_b200 0.1 (sm_100a)"; }

int b200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int b200_create(int device, b200_ctx** out) {
  if (!out) return B200_ERR_INVALID;
  *out = nullptr;
  if (device == -1) {  // host-only context: structure phase for CPU-side tests, compute calls fail with NO_DEVICE
    b200_ctx* c = new b200_ctx();
    c->device = -1;
    c->host_only = true;
    *out = c;
    return B200_OK;
  }
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    g_create_error = "no CUDA device available: the B200 solve path has no CPU fallback";
    return B200_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= n) { g_create_error = "invalid device index"; return B200_ERR_INVALID; }
  b200_ctx* c = new b200_ctx();
  c->device = device;
  if (const char* e = getenv("G2O_B200_FUSE")) c->fuse_tail = atoi(e) != 0;  // 0: separate back-substitution / update / chi2 / scale kernels
  try {
    B200_CUDA(cudaSetDevice(device));
    B200_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    B200_CUDA(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
    B200_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    B200_CUDA(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    if (const char* e2 = getenv("G2O_B200_OVERLAP")) c->overlap_linearize = atoi(e2) != 0;
    if (const char* e2 = getenv("G2O_B200_LIN_PACKETS")) c->lin_packets = atoi(e2) != 0;   // 0: the thread-per-landmark kernel
    if (const char* e2 = getenv("G2O_B200_CAMS_MINB")) c->cams_minb = atoi(e2);            // ... and the per-camera pass (1 / 4 / 6)
    if (const char* e2 = getenv("G2O_B200_LIN_MINB")) c->lin_minb = atoi(e2);              // CTAs per SM the packets kernel is compiled for
    B200_CUDA(cudaMallocHost((void**)&c->h_scalars, 16 * sizeof(double)));
    B200_CUDA(cudaMallocHost((void**)&c->h_status, sizeof(int)));
  } catch (const CudaError& err) {
    g_create_error = describe(err);
    delete c;
    return B200_ERR_CUDA;
  }
  *out = c;
  return B200_OK;
}

void b200_destroy(b200_ctx* c) {
  if (!c) return;
  if (c->host_only) { host_only_flag() = true; delete c; host_only_flag() = false; return; }
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  c->prof.destroy();
  drop_graphs(c);
  c->nccl.destroy();
  if (c->h_scalars) cudaFreeHost(c->h_scalars);
  if (c->h_status) cudaFreeHost(c->h_status);
  cudaStream_t s = c->stream, s2 = c->stream2;
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  delete c;
  if (s2) cudaStreamDestroy(s2);
  if (s) cudaStreamDestroy(s);
}

const char* b200_last_error(const b200_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int b200_set_vertices(b200_ctx* c, int kind, int n, const double* est, const int32_t* hidx, const uint8_t* marg) {
  if (!c || kind < 0 || kind >= B200_NUM_VERTEX_KINDS || n < 0 || (n > 0 && (!est || !hidx))) return B200_ERR_INVALID;
  if (kind == B200_VERTEX_CAM) c->cam_model = 0;
  if (kind == B200_VERTEX_SE3_EXPMAP) { c->cam_model = 1; kind = B200_VERTEX_CAM; }  // same slot, same 12-double rows
  b200_ctx::VertexSet& V = c->vs[kind];
  V.set = true; V.n = n;
  V.est.assign(est, est + (size_t)n * vest(kind));
  V.hidx.assign(hidx, hidx + n);
  if (marg) V.marg.assign(marg, marg + n); else V.marg.clear();
  c->structured = false;
  // a new vertex set renumbers the poses: pattern keys handed over for the previous graph are stale (the uploader
  // sends vertices first, then the foreign Schur pattern, then the edges)
  c->extra_schur_keys.clear();
  return B200_OK;
}

int b200_set_edges(b200_ctx* c, int kind, int n, const int32_t* vi, const int32_t* vj, const double* meas, const double* info) {
  if (!c || kind < 0 || kind >= B200_NUM_EDGE_KINDS || n < 0 || (n > 0 && (!vi || !vj || !meas || !info))) return B200_ERR_INVALID;
  if (kind == B200_EDGE_SE2_XY || kind == B200_EDGE_SE3_XYZ) {  // the pose-landmark set of a landmark-SLAM graph
    c->l_edge_kind = n > 0 ? kind : -1; c->nLE = n;
    c->l_vi.assign(vi, vi + n); c->l_vj.assign(vj, vj + n);
    c->l_meas.assign(meas, meas + (size_t)n * emeas(kind));
    const int D = edim(kind);
    c->l_info.assign(info, info + (size_t)n * D * D);
    c->l_rk_kinds.clear(); c->l_rk_deltas.clear();   // per-edge kernels belong to the edge set they were given for
    c->structured = false;
    return B200_OK;
  }
  if (kind == B200_EDGE_P2MC) c->edge_model = 0;
  if (kind == B200_EDGE_XYZ2UV) { c->edge_model = 1; kind = B200_EDGE_P2MC; }  // same sizes (2 | 2x2), same structure
  c->edge_kind = kind; c->nE = n;
  c->e_vi.assign(vi, vi + n); c->e_vj.assign(vj, vj + n);
  c->e_meas.assign(meas, meas + (size_t)n * emeas(kind));
  const int D = edim(kind);
  c->e_info.assign(info, info + (size_t)n * D * D);
  c->rk_kinds.clear(); c->rk_deltas.clear();
  c->structured = false;
  return B200_OK;
}

int b200_set_sensor_offset(b200_ctx* c, const double* iso) {
  if (!c || !iso) return B200_ERR_INVALID;
  memcpy(c->sensor_offset, iso, sizeof(c->sensor_offset));
  c->state_chi2_valid = false;
  drop_graphs(c);  // the offset is a kernel argument of the captured launches
  return B200_OK;
}

int b200_set_allreduce(b200_ctx* c, b200_allreduce_fn fn, void* user, int rank, int world) {
  if (!c || world < 1 || rank < 0 || rank >= world) return B200_ERR_INVALID;
  c->allreduce = fn; c->allreduce_user = user; c->rank = rank; c->world = world;
  return B200_OK;
}

int b200_comm_unique_id(void* out, int capacity) {
  if (!out || capacity < B200_COMM_ID_BYTES) return B200_ERR_INVALID;
  std::string err;
  if (NcclComm::unique_id(out, &err)) { g_create_error = err; return B200_ERR_COLLECTIVE; }
  return B200_OK;
}
int b200_comm_init(b200_ctx* c, const void* unique_id, int rank, int world) {
  if (!c || !unique_id || world < 1 || rank < 0 || rank >= world) return B200_ERR_INVALID;
  return guarded(c, [&]() -> int {
    NEED_DEVICE_(c);
    B200_CUDA(cudaSetDevice(c->device));
    drop_graphs(c);
    if (c->nccl.init(unique_id, rank, world, &c->err)) return B200_ERR_COLLECTIVE;
    c->rank = rank; c->world = world; c->comm_warm = false;
    return (int)B200_OK;
  });
}
int b200_comm_destroy(b200_ctx* c) {
  if (!c) return B200_ERR_INVALID;
  if (!c->host_only) { cudaSetDevice(c->device); if (c->stream) cudaStreamSynchronize(c->stream); drop_graphs(c); }
  c->nccl.destroy();
  if (!c->allreduce) { c->rank = 0; c->world = 1; }
  return B200_OK;
}
int b200_comm_version(void) { return NcclComm::version(); }

int b200_add_schur_pattern(b200_ctx* c, int n, const int32_t* rows, const int32_t* cols) {
  if (!c || n < 0 || (n > 0 && (!rows || !cols))) return B200_ERR_INVALID;
  for (int i = 0; i < n; ++i)
    if (rows[i] < 0 || cols[i] < 0 || rows[i] > cols[i]) return fail(c, B200_ERR_INVALID, "Schur pattern keys must be upper-triangular block indices (0 <= row <= col)");
  for (int i = 0; i < n; ++i) c->extra_schur_keys.push_back(((long long)cols[i] << 32) | (unsigned)rows[i]);
  c->structured = false;
  return B200_OK;
}

int b200_build_structure(b200_ctx* c) {
  return guarded(c, [&]() {
    if (!c->host_only) B200_CUDA(cudaSetDevice(c->device));
    return build_structure_impl(c);
  });
}

#define NEED_STRUCTURE(c) \
  if (!(c)->structured) return fail((c), B200_ERR_INVALID, "call b200_build_structure first")
#define NEED_DEVICE(c) \
  if ((c)->host_only) return fail((c), B200_ERR_NO_DEVICE, "host-only context: the B200 solve path has no CPU fallback")

int b200_compute_active_errors(b200_ctx* c, double* chi2) {
  return guarded(c, [&]() {
    NEED_DEVICE(c);
    NEED_STRUCTURE(c);
    B200_CUDA(cudaSetDevice(c->device));
    enqueue_chi2(c);
    int rc = reduce_trial_scalars(c);
    if (rc) return rc;
    sync_scalars(c);
    c->last_chi2 = c->h_scalars[0];
    if (chi2) *chi2 = c->h_scalars[0];
    return (int)B200_OK;
  });
}

int b200_build_system(b200_ctx* c) {
  return guarded(c, [&]() {
    NEED_DEVICE(c);
    NEED_STRUCTURE(c);
    B200_CUDA(cudaSetDevice(c->device));
    return enqueue_build_system(c);
  });
}

int b200_set_lambda(b200_ctx* c, double lambda, int /*backup*/) {
  return guarded(c, [&]() {
    NEED_STRUCTURE(c);
    // the diagonal is never modified in place: lambda is applied where H is consumed (Cholesky scatter,
    // landmark inverses, Schur diagonal), which is what backup + restoreDiagonal achieve in the reference
    c->lambda_for_solve = lambda;
    return (int)B200_OK;
  });
}
int b200_restore_diagonal(b200_ctx* c) {
  return guarded(c, [&]() {
    NEED_STRUCTURE(c);
    c->lambda_for_solve = 0.0;
    return (int)B200_OK;
  });
}

int b200_solve(b200_ctx* c) {
  return guarded(c, [&]() {
    NEED_DEVICE(c);
    NEED_STRUCTURE(c);
    B200_CUDA(cudaSetDevice(c->device));
    set_device_lambda(c, c->lambda_for_solve);
    int rc = enqueue_solve(c);
    if (rc) return rc;
    sync_scalars(c);
    return *c->h_status ? (int)B200_NOT_POSITIVE_DEFINITE : (int)B200_OK;
  });
}

// Solver::computeMarginals (core/block_solver.hpp:490-499 -> LinearSolver::solvePattern,
// solvers/csparse/linear_solver_csparse.h:190-225 -> MarginalCovarianceCholesky, core/marginal_covariance_cholesky.cpp):
// selected blocks of Hpp^-1.  The reference walks the scalar factor with a recursive formula and a cache; here ONE
// factorisation is followed by the same recursion in supernodal form on the GPU (sparse_inverse.cuh), which yields every
// block on the pattern of L at the cost of one more factorisation; a requested block outside that pattern (no fill
// between the two poses) falls back to d unit solves for its block column.
int b200_compute_marginals(b200_ctx* c, int nblocks, const int32_t* rows, const int32_t* cols, double* out) {
  return guarded(c, [&]() -> int {
    NEED_DEVICE(c);
    NEED_STRUCTURE(c);
    if (nblocks < 0 || (nblocks > 0 && (!rows || !cols || !out))) return fail(c, B200_ERR_INVALID, "bad arguments");
    if (c->schur) return fail(c, B200_ERR_UNSUPPORTED, "marginals are provided for pose graphs (no Schur complement): the reference inverts Hpp there");
    B200_CUDA(cudaSetDevice(c->device));
    const int d = c->pd, n = c->sizeP;
    for (int q = 0; q < nblocks; ++q)
      if (rows[q] < 0 || rows[q] >= c->np || cols[q] < 0 || cols[q] >= c->np) return fail(c, B200_ERR_INVALID, "block index out of range");
    cudaStream_t s = c->stream;
    DevBuf<double>& rhs = c->d_marg_rhs;
    DevBuf<double>& xs = c->d_marg_x;
    rhs.alloc(n); xs.alloc(n);
    // (1) ONE factorisation of Hpp (lambda = 0; the forward substitution that rides along gets a zero right-hand side),
    //     then the sparse inverse subset on the factor: every requested block on the pattern of L comes out of one
    //     sweep down the supernodal tree (chol.h: sparse_inverse)
    {
      B200_CUDA(cudaMemsetAsync(rhs.p, 0, n * sizeof(double), s));
      c->chol.keep_chain_inverses(true);
      c->chol.factor(c->d_Hpp.p, nullptr, rhs.p, s, &c->lc, nullptr);
      c->chol.keep_chain_inverses(false);
      int status = 0;
      B200_CUDA(cudaMemcpyAsync(&status, c->chol.status_ptr(), sizeof(int), cudaMemcpyDeviceToHost, s));
      B200_CUDA(cudaStreamSynchronize(s));
      if (status) return (int)B200_NOT_POSITIVE_DEFINITE;
    }
    std::vector<int> in_pattern, outside;
    std::vector<long long> off;
    std::vector<int> ld;
    std::vector<unsigned char> trans;
    for (int q = 0; q < nblocks; ++q) {
      long long o; int l; bool t;
      if (c->chol.locate_inverse_block(rows[q], cols[q], &o, &l, &t)) { in_pattern.push_back(q); off.push_back(o); ld.push_back(l); trans.push_back(t); }
      else outside.push_back(q);
    }
    if (!in_pattern.empty()) {
      c->chol.sparse_inverse(s, &c->lc);
      const int m = (int)in_pattern.size();
      DevBuf<long long>& d_off = c->d_marg_off; DevBuf<int>& d_ld = c->d_marg_ld; DevBuf<unsigned char>& d_trans = c->d_marg_trans; DevBuf<double>& d_out = c->d_marg_out;
      d_off.upload(off, s); d_ld.upload(ld, s); d_trans.upload(trans, s); d_out.alloc((size_t)m * d * d);
      c->chol.gather_inverse_blocks(m, d_off.p, d_ld.p, d_trans.p, d_out.p, s, &c->lc);
      std::vector<double> h((size_t)m * d * d);
      B200_CUDA(cudaMemcpyAsync(h.data(), d_out.p, h.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
      B200_CUDA(cudaStreamSynchronize(s));
      for (int k = 0; k < m; ++k) memcpy(out + (size_t)in_pattern[k] * d * d, &h[(size_t)k * d * d], (size_t)d * d * sizeof(double));
    }
    // (2) blocks outside the pattern of L (the reference's recursion reaches them through entries it has to create on the
    //     way): columns of the inverse by unit solves, one block column at a time
    const int nout = (int)outside.size();
    std::vector<int> order(outside);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b2) { return cols[a] < cols[b2]; });
    std::vector<double> hx(n);
    const double one = 1.0;
    for (int q0 = 0; q0 < nout;) {
      const int cb = cols[order[q0]];
      int q1 = q0;
      while (q1 < nout && cols[order[q1]] == cb) ++q1;
      for (int k = 0; k < d; ++k) {
        B200_CUDA(cudaMemsetAsync(rhs.p, 0, n * sizeof(double), s));
        B200_CUDA(cudaMemcpyAsync(rhs.p + (size_t)cb * d + k, &one, sizeof(double), cudaMemcpyHostToDevice, s));
        c->chol.factor(c->d_Hpp.p, nullptr, rhs.p, s, &c->lc, nullptr);
        c->chol.solve(rhs.p, xs.p, s, &c->lc, nullptr);
        B200_CUDA(cudaMemcpyAsync(hx.data(), xs.p, n * sizeof(double), cudaMemcpyDeviceToHost, s));
        B200_CUDA(cudaStreamSynchronize(s));
        for (int q = q0; q < q1; ++q) {
          const int req = order[q];
          for (int i = 0; i < d; ++i) out[(size_t)req * d * d + i + (size_t)k * d] = hx[(size_t)rows[req] * d + i];
        }
      }
      q0 = q1;
    }
    return (int)B200_OK;
  });
}

int b200_update(b200_ctx* c) {
  return guarded(c, [&]() {
    NEED_DEVICE(c);
    NEED_STRUCTURE(c);
    B200_CUDA(cudaSetDevice(c->device));
    c->state_chi2_valid = false;
    enqueue_update(c);
    return (int)B200_OK;
  });
}
int b200_push(b200_ctx* c) {
  return guarded(c, [&]() {
    NEED_STRUCTURE(c);
    NEED_DEVICE(c);
    if (c->backup_depth >= 1) return fail(c, B200_ERR_UNSUPPORTED, "backup stack depth > 1 (LM needs 1)");
    B200_CUDA(cudaSetDevice(c->device));
    do_push(c);
    c->backup_depth = 1;
    return (int)B200_OK;
  });
}
int b200_pop(b200_ctx* c) {
  return guarded(c, [&]() {
    NEED_STRUCTURE(c);
    if (c->backup_depth < 1) return fail(c, B200_ERR_INVALID, "pop on an empty backup stack");
    B200_CUDA(cudaSetDevice(c->device));
    c->state_chi2_valid = false;
    do_pop(c);
    c->backup_depth = 0;
    return (int)B200_OK;
  });
}
int b200_discard_top(b200_ctx* c) {
  return guarded(c, [&]() {
    NEED_STRUCTURE(c);
    if (c->backup_depth < 1) return fail(c, B200_ERR_INVALID, "discardTop on an empty backup stack");
    c->backup_depth = 0;
    return (int)B200_OK;
  });
}

int b200_set_ordering(b200_ctx* c, int nd_levels) {
  if (!c || nd_levels < 0 || nd_levels > 16) return B200_ERR_INVALID;
  c->nd_levels = nd_levels;
  c->structured = false;  // takes effect at the next b200_build_structure
  return B200_OK;
}

int b200_set_linear_solver(b200_ctx* c, int kind, double tolerance, int absolute_tolerance, int max_iterations) {
  if (!c || (kind != B200_LINEAR_SOLVER_CHOLESKY && kind != B200_LINEAR_SOLVER_PCG)) return B200_ERR_INVALID;
  if (kind == B200_LINEAR_SOLVER_PCG && !(tolerance > 0)) return fail(c, B200_ERR_INVALID, "PCG tolerance must be positive");
  if (kind != c->linear_solver) c->structured = false;   // the PCG lists are built with the structure
  c->linear_solver = kind;
  if (kind == B200_LINEAR_SOLVER_PCG) { c->pcg_tolerance = tolerance; c->pcg_absolute = absolute_tolerance != 0; c->pcg_max_iterations = max_iterations; }
  drop_graphs(c);
  return B200_OK;
}

int b200_get_linear_solver_iterations(b200_ctx* c) { return c ? (c->linear_solver == 1 ? c->pcg_last_iterations : 0) : B200_ERR_INVALID; }

int b200_set_robust_kernel(b200_ctx* c, int kind, double delta) {
  if (!c || kind < B200_ROBUST_NONE || kind > B200_ROBUST_DCS || !(delta > 0.0)) return B200_ERR_INVALID;
  c->rk_uniform_kind = kind;
  c->rk_uniform_delta = delta;
  c->rk_kinds.clear(); c->rk_deltas.clear(); c->l_rk_kinds.clear(); c->l_rk_deltas.clear();   // one kernel for every edge again
  c->state_chi2_valid = false;
  if (!c->host_only) { cudaSetDevice(c->device); drop_graphs(c); }  // captured launches carry the kernel by value
  return guarded(c, [&]() { upload_robust(c); return (int)B200_OK; });
}

int b200_set_edge_robust_kernels(b200_ctx* c, int edge_kind, int n, const uint8_t* kinds, const double* deltas) {
  if (!c || edge_kind < 0 || edge_kind >= B200_NUM_EDGE_KINDS || n < 0) return B200_ERR_INVALID;
  const bool lset = edge_kind == B200_EDGE_SE2_XY || edge_kind == B200_EDGE_SE3_XYZ;
  std::vector<unsigned char>& K = lset ? c->l_rk_kinds : c->rk_kinds;
  std::vector<double>& D = lset ? c->l_rk_deltas : c->rk_deltas;
  if (!kinds) { K.clear(); D.clear(); }
  else {
    if (!deltas || n != (lset ? c->nLE : c->nE)) return fail(c, B200_ERR_INVALID, "b200_set_edge_robust_kernels: one (kind, width) per edge of the set given to b200_set_edges");
    for (int i = 0; i < n; ++i)
      if (kinds[i] > B200_ROBUST_DCS || (kinds[i] != B200_ROBUST_NONE && !(deltas[i] > 0.0))) return fail(c, B200_ERR_INVALID, "b200_set_edge_robust_kernels: unknown kernel or non-positive width");
    K.assign(kinds, kinds + n); D.assign(deltas, deltas + n);
  }
  c->state_chi2_valid = false;
  if (!c->host_only) { cudaSetDevice(c->device); drop_graphs(c); }
  if (!c->structured) return B200_OK;   // uploaded (in device edge order) by b200_build_structure
  return guarded(c, [&]() { upload_robust(c); return (int)B200_OK; });
}

int b200_set_terminate(b200_ctx* c, b200_terminate_fn fn, void* user) {
  if (!c) return B200_ERR_INVALID;
  c->terminate = fn; c->terminate_user = user;
  return B200_OK;
}

int b200_set_lm_params(b200_ctx* c, double user_lambda_init, int max_trials) {
  if (!c || max_trials < 1) return B200_ERR_INVALID;
  c->user_lambda_init = user_lambda_init;
  c->max_trials_after_failure = max_trials;
  return B200_OK;
}

// one OptimizationAlgorithm::solve(iteration); returns SolverResult in st->result and as return value
int b200_algorithm_solve(b200_ctx* c, int algorithm, int iteration, b200_iter_stats* st) {
  return guarded(c, [&]() -> int {
    NEED_DEVICE(c);
    B200_CUDA(cudaSetDevice(c->device));
    b200_iter_stats local;
    if (!st) st = &local;
    memset(st, 0, sizeof(*st));
    st->iteration = iteration;
    const double t_start = wall();
    if (iteration == 0 && !c->structured) {
      int rc = build_structure_impl(c);
      if (rc) return rc;
    }
    NEED_STRUCTURE(c);
    if (iteration == 0) st->time_symbolic = c->time_symbolic;
    // G2OBatchStatistics phase fields (core/batch_stats.h:49-58) from the CUDA-event profiler, when it is on
    double ph0[PH_COUNT];
    if (c->prof.on) { c->prof.flush(); for (int i = 0; i < PH_COUNT; ++i) ph0[i] = c->prof.seconds[i]; }
    auto fill_phase_stats = [&]() {
      if (!c->prof.on) return;
      c->prof.flush();
      auto d = [&](int ph) { return c->prof.seconds[ph] - ph0[ph]; };
      st->time_residuals = d(PH_ERRORS);                                             // computeActiveErrors (+ chi2)
      st->time_quadratic_form = d(PH_LINEARIZE) + d(PH_LINEARIZE_CAMS) + d(PH_GATHER);  // buildSystem
      st->time_schur = d(PH_SCHUR_INV) + d(PH_SCHUR) + d(PH_SCHUR_FINISH) + d(PH_COLLECTIVE);  // block_solver.hpp:370-441
      st->time_numeric = d(PH_FACTOR);                                               // numeric decomposition (+ forward substitution)
      st->time_linear_solution = d(PH_TRISOLVE);                                     // backward substitution
      st->time_linear_solver = d(PH_FACTOR) + d(PH_TRISOLVE) + d(PH_BACKSUB);        // block_solver.hpp:445-486
      st->time_update = d(PH_UPDATE);
    };
    int rc = 0;
    if (algorithm == B200_GAUSS_NEWTON) {
      // core/optimization_algorithm_gauss_newton.cpp:50-93 (computeActiveErrors only caches edge errors there)
      set_device_lambda(c, 0.0);
      c->state_chi2_valid = false;
      if ((rc = enqueue_build_system(c))) return rc;
      if ((rc = enqueue_solve(c))) return rc;
      enqueue_update(c);
      sync_scalars(c);
      st->result = *c->h_status ? B200_RESULT_FAIL : B200_RESULT_OK;
      fill_phase_stats();
      st->time_iteration = wall() - t_start;
      return st->result == B200_RESULT_FAIL ? B200_SOLVE_FAIL : st->result;
    }
    // ---- Levenberg-Marquardt: core/optimization_algorithm_levenberg.cpp:57-147
    const bool shard = sharded(c);
    const bool know_chi2 = c->state_chi2_valid && iteration > 0;
    c->state_chi2_valid = false;
    if ((rc = run_prologue(c, !know_chi2))) return rc;
    if (iteration == 0 && (rc = enqueue_max_diag(c))) return rc;
    // sharded: chi2 of the current state is a partial sum here; the total arrives with the first trial's all-reduce
    if (!know_chi2 && (!shard || iteration == 0)) sync_scalars(c);
    double currentChi = know_chi2 ? c->state_chi2 : shard ? 0.0 : c->h_scalars[0];
    double tempChi = currentChi;
    if (iteration == 0) {
      c->lambda = c->user_lambda_init > 0 ? c->user_lambda_init : 1e-5 * c->h_scalars[2];
      c->ni = 2;
    }
    double rho = 0;
    int& qmax = c->levenberg_iterations;
    qmax = 0;
    do {
      do_push(c);
      if ((rc = run_trial(c))) return rc;
      sync_scalars(c);
      if (shard && qmax == 0 && !know_chi2) currentChi = c->h_scalars[5];
      const bool ok2 = *c->h_status == 0;
      tempChi = c->h_scalars[0];
      if (!ok2) tempChi = DBL_MAX;
      rho = currentChi - tempChi;
      double scale = c->h_scalars[4] + c->h_scalars[1];
      scale += 1e-3;
      rho /= scale;
      if (rho > 0 && std::isfinite(tempChi)) {
        double alpha = 1. - pow((2 * rho - 1), 3);
        alpha = std::min(alpha, 2. / 3.);
        double scaleFactor = std::max(1. / 3., alpha);
        c->lambda *= scaleFactor;
        c->ni = 2;
        currentChi = tempChi;
      } else {
        c->lambda *= c->ni;
        c->ni *= 2;
        do_pop(c);
      }
      qmax++;
    } while (rho < 0 && qmax < c->max_trials_after_failure && !(c->terminate && c->terminate(c->terminate_user)));
    c->last_chi2 = currentChi;
    c->state_chi2 = currentChi;      // accepted: the trial's chi2; rejected: the backup was restored
    c->state_chi2_valid = true;
    st->chi2 = currentChi;
    st->lambda = c->lambda;
    st->levenberg_iterations = qmax;
    st->result = (qmax == c->max_trials_after_failure || rho == 0) ? B200_RESULT_TERMINATE : B200_RESULT_OK;
    fill_phase_stats();
    st->time_iteration = wall() - t_start;
    return st->result;
  });
}

int b200_optimize(b200_ctx* c, int algorithm, int max_iterations, b200_iter_stats* stats) {
  if (!c) return B200_ERR_INVALID;
  // OptimizationAlgorithmWithHessian::init -> BlockSolver::init -> LinearSolver::init would drop the symbolic
  // factor here; the structure is a pure function of the graph handed over, so it is kept until the graph changes
  int done = 0, result = B200_RESULT_OK;
  bool ok = true;
  for (int i = 0; i < max_iterations && ok && !(c->terminate && c->terminate(c->terminate_user)); ++i) {
    b200_iter_stats local;
    b200_iter_stats* st = stats ? &stats[i] : &local;
    result = b200_algorithm_solve(c, algorithm, i, st);
    if (result < 0) return result;  // hard error (never an LM/GN outcome: those are 1, 2, B200_SOLVE_FAIL)
    ok = result == B200_RESULT_OK;
    // what SparseOptimizer::optimize does when statistics are on: chi2 of the state after the iteration
    // (LM already knows it; GN needs one more error pass)
    if (stats && algorithm == B200_GAUSS_NEWTON) {
      double chi = 0;
      int rc = b200_compute_active_errors(c, &chi);
      if (rc) return rc;
      st->chi2 = chi;
    }
    ++done;
  }
  if (result == B200_SOLVE_FAIL) return 0;
  return done;
}

// ------------------------------------------------------------------------------------------------ read-back
int b200_get_dims(b200_ctx* c, int32_t* d) {
  if (!c || !d) return B200_ERR_INVALID;
  d[0] = c->np; d[1] = c->nl; d[2] = c->sizeP; d[3] = c->sizeL; d[4] = c->nE + (c->var_lm ? c->nLE : 0); d[5] = c->n_pose_v + c->n_lm_v; d[6] = c->pd; d[7] = c->schur ? c->ld : 0;
  return B200_OK;
}
static int copy_out(b200_ctx* c, const double* dev, double* host, size_t n) {
  NEED_DEVICE(c);
  B200_CUDA(cudaSetDevice(c->device));
  B200_CUDA(cudaMemcpyAsync(host, dev, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  B200_CUDA(cudaStreamSynchronize(c->stream));
  return B200_OK;
}
int b200_get_x(b200_ctx* c, double* x) {
  return guarded(c, [&]() { NEED_STRUCTURE(c); return copy_out(c, c->d_x.p, x, (size_t)c->sizeP + c->sizeL); });
}
int b200_get_b(b200_ctx* c, double* b) {
  return guarded(c, [&]() { NEED_STRUCTURE(c); return copy_out(c, c->d_b.p, b, (size_t)c->sizeP + c->sizeL); });
}
int b200_get_bschur(b200_ctx* c, double* out) {
  return guarded(c, [&]() { NEED_STRUCTURE(c); if (!c->schur) return fail(c, B200_ERR_INVALID, "no Schur complement in this problem"); return copy_out(c, bschur_ptr(c), out, (size_t)c->sizeP); });
}
int b200_get_estimates(b200_ctx* c, int kind, double* out) {
  return guarded(c, [&]() {
    NEED_STRUCTURE(c);
    if (kind == B200_VERTEX_SE3_EXPMAP || kind == B200_VERTEX_CAM) kind = ((kind == B200_VERTEX_SE3_EXPMAP) == (c->cam_model == 1)) ? B200_VERTEX_CAM : -1;
    const bool lm = is_landmark_kind(kind);
    if (!lm && kind != c->pose_kind) return fail(c, B200_ERR_INVALID, "vertex kind not present");
    if (lm && !(c->schur ? kind == B200_VERTEX_XYZ : (c->var_lm && kind == c->lm_kind))) return fail(c, B200_ERR_INVALID, "vertex kind not present");
    const int n = lm ? c->n_lm_v : c->n_pose_v, st = vstride(kind), ne = vest(kind);
    if (c->host_only) {  // no device, nothing was optimised: the estimates as ingested (structure-phase tests on the CPU box)
      const b200_ctx::VertexSet& V = c->vs[lm ? kind : c->pose_kind];
      std::copy(V.est.begin(), V.est.begin() + (size_t)n * ne, out);
      return (int)B200_OK;
    }
    B200_CUDA(cudaSetDevice(c->device));
    const double* dev = lm ? c->d_lm_est.p : c->d_pose_est.p;
    if (st == ne) {
      B200_CUDA(cudaMemcpyAsync(out, dev, (size_t)n * ne * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    } else {
      c->d_stage_est.alloc((size_t)n * ne);
      k::pack_rows_kernel<<<ceil_div((long long)n * ne, 256), 256, 0, c->stream>>>(n, ne, st, dev, c->d_stage_est.p);
      c->lc.n++;
      B200_CUDA(cudaMemcpyAsync(out, c->d_stage_est.p, (size_t)n * ne * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    }
    B200_CUDA(cudaStreamSynchronize(c->stream));
    return (int)B200_OK;
  });
}
int b200_set_estimates(b200_ctx* c, int kind, const double* est) {
  return guarded(c, [&]() {
    NEED_DEVICE(c);
    NEED_STRUCTURE(c);
    c->state_chi2_valid = false;
    if (kind == B200_VERTEX_SE3_EXPMAP || kind == B200_VERTEX_CAM) kind = ((kind == B200_VERTEX_SE3_EXPMAP) == (c->cam_model == 1)) ? B200_VERTEX_CAM : -1;
    const bool lm = is_landmark_kind(kind);
    if ((!lm && kind != c->pose_kind) || (lm && !(c->schur ? kind == B200_VERTEX_XYZ : (c->var_lm && kind == c->lm_kind))) || !est) return fail(c, B200_ERR_INVALID, "vertex kind not present");
    B200_CUDA(cudaSetDevice(c->device));
    const int n = lm ? c->n_lm_v : c->n_pose_v, st = vstride(kind), ne = vest(kind);
    double* dev = lm ? c->d_lm_est.p : c->d_pose_est.p;
    if (st == ne) {
      B200_CUDA(cudaMemcpyAsync(dev, est, (size_t)n * ne * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    } else {  // padded rows on the device (3 -> 4 doubles): one contiguous copy, then expand on the device
      c->d_stage_est.alloc((size_t)n * ne);
      B200_CUDA(cudaMemcpyAsync(c->d_stage_est.p, est, (size_t)n * ne * sizeof(double), cudaMemcpyHostToDevice, c->stream));
      k::unpack_rows_kernel<<<ceil_div((long long)n * ne, 256), 256, 0, c->stream>>>(n, ne, st, c->d_stage_est.p, dev);
      c->lc.n++;
    }
    if (kind == B200_VERTEX_CAM) {
      BA_MODEL_LAUNCH(c, cam_derive_kernel, <<<ceil_div(n, 128), 128, 0, c->stream>>>(n, c->d_pose_est.p, c->d_cam_der.p));
      c->lc.n++;
    }
    B200_CUDA(cudaStreamSynchronize(c->stream));
    return (int)B200_OK;
  });
}
int b200_get_hessian_diagonal(b200_ctx* c, double* diag) {
  return guarded(c, [&]() {
    NEED_DEVICE(c);
    NEED_STRUCTURE(c);
    B200_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    if (c->pd == 3) k::extract_diag_kernel<3><<<ceil_div((long long)c->np * 3, 256), 256, 0, s>>>(c->np, c->d_hpp_diag_block.p, c->d_Hpp.p, c->d_diag.p);
    else k::extract_diag_kernel<6><<<ceil_div((long long)c->np * 6, 256), 256, 0, s>>>(c->np, c->d_hpp_diag_block.p, c->d_Hpp.p, c->d_diag.p);
    c->lc.n++;
    if (c->schur && c->nl > 0) { k::extract_diag_kernel<3><<<ceil_div((long long)c->nl * 3, 256), 256, 0, s>>>(c->nl, nullptr, c->d_Hll.p, c->d_diag.p + c->sizeP); c->lc.n++; }
    B200_CUDA(cudaGetLastError());
    return copy_out(c, c->d_diag.p, diag, (size_t)c->sizeP + c->sizeL);
  });
}
int b200_get_blocks(b200_ctx* c, int which, int32_t* rows, int32_t* cols, double* values) {
  return guarded(c, [&]() -> int {
    NEED_STRUCTURE(c);
    const int pd = c->pd;
    if (which == 0) {
      if (!rows) return c->n_hpp;
      int k2 = 0;
      for (int col = 0; col < c->np; ++col) for (int p = c->hpp_colptr[col]; p < c->hpp_colptr[col + 1]; ++p, ++k2) { rows[k2] = c->hpp_rowidx[p]; cols[k2] = col; }
      if (values) { int rc2 = copy_out(c, c->d_Hpp.p, values, (size_t)c->n_hpp * pd * pd); if (rc2) return rc2; }
      return c->n_hpp;
    }
    if (!c->schur) return fail(c, B200_ERR_INVALID, "no landmark blocks in this problem");
    if (which == 1) {
      if (!rows) return c->nl;
      for (int l = 0; l < c->nl; ++l) rows[l] = cols[l] = l;
      if (values) { int rc2 = copy_out(c, c->d_Hll.p, values, (size_t)c->nl * 9); if (rc2) return rc2; }
      return c->nl;
    }
    if (which == 2) {  // exported in SparseBlockMatrix order (landmark-major, ascending camera); stored in Schur-range order
      if (!rows) return c->n_hpl;
      for (int q = 0; q < c->n_hpl; ++q) { const int sl = c->hpl_export[q]; rows[q] = c->hpl_row[sl]; cols[q] = c->hpl_col[sl]; }
      if (values) {
        std::vector<double> tmp((size_t)c->n_hpl * 18);
        int rc2 = copy_out(c, c->d_Hpl.p, tmp.data(), tmp.size());
        if (rc2) return rc2;
        for (int q = 0; q < c->n_hpl; ++q) std::copy_n(tmp.data() + 18 * (size_t)c->hpl_export[q], 18, values + 18 * (size_t)q);
      }
      return c->n_hpl;
    }
    if (which == 3) {
      if (!rows) return c->n_hs;
      int k2 = 0;
      for (int col = 0; col < c->np; ++col) for (int p = c->hs_colptr[col]; p < c->hs_colptr[col + 1]; ++p, ++k2) { rows[k2] = c->hs_rowidx[p]; cols[k2] = col; }
      if (values) { int rc2 = copy_out(c, c->d_Hschur.p, values, (size_t)c->n_hs * 36); if (rc2) return rc2; }
      return c->n_hs;
    }
    return fail(c, B200_ERR_INVALID, "which must be 0..3");
  });
}
int b200_get_block_ordering(b200_ctx* c, int32_t* perm) {
  if (!c || !c->structured) return B200_ERR_INVALID;
  const std::vector<int>& P = c->chol.symbolic().perm;
  if (perm) memcpy(perm, P.data(), P.size() * sizeof(int));
  return (int)P.size();
}
int64_t b200_get_factor_nnz(b200_ctx* c) { return (c && c->structured) ? c->chol.symbolic().scalar_lnz : -1; }
int b200_get_factor_info(b200_ctx* c, int64_t* out) {
  if (!c || !c->structured || !out) return B200_ERR_INVALID;
  const SymbolicFactor& S = c->chol.symbolic();
  out[0] = S.nsn; out[1] = (int64_t)S.task_ptr.size() - 1; out[2] = S.nlevels; out[3] = S.max_nrow; out[4] = S.max_ncol; out[5] = S.factor_doubles;
  out[6] = (int64_t)S.flow_kind.size();
  out[12] = (int64_t)S.flops;
  out[13] = (int64_t)S.chain_sn.size(); out[14] = (int64_t)S.chain_flops; out[15] = 0;
  // out[16..]: update plan of the factorisation: tile geometry, work items, flops the items execute (incl. the parts
  // of their rectangular products that fall above the diagonal)
  out[20] = (int64_t)S.subtree_flops;
  out[16] = S.wide ? 1 : 0; out[17] = (int64_t)S.work_u.size(); out[19] = S.max_group_slots;
  {
    double ex = 0;
    for (size_t q = 0; q < S.work_u.size(); ++q)
      ex += 2.0 * (S.work_a1[q] - S.work_a0[q]) * S.d * (double)((S.work_b1[q] - S.work_b0[q]) * S.d) * S.work_nk[q];
    out[18] = (int64_t)ex;
  }
  out[7] = c->sr_n; out[8] = c->sr_nseg; out[9] = c->sr_ncontrib; out[10] = c->n_hpl; out[11] = c->sr_n > 0 ? (int64_t)schur_range_smem(c) : 0;
  return B200_OK;
}
int64_t b200_get_launch_count(b200_ctx* c) { return c ? c->lc.n : -1; }
uint64_t b200_debug_upload_digest(int reset) {
  const uint64_t h = upload_digest();
  if (reset) upload_digest() = 1469598103934665603ull;
  return h;
}
int b200_set_profiling(b200_ctx* c, int on) {
  if (!c) return B200_ERR_INVALID;
  if (!c->host_only) { cudaSetDevice(c->device); c->prof.stream = c->stream; c->prof.reset(); }
  c->prof.on = on != 0;
  return B200_OK;
}
int b200_get_phase_time(b200_ctx* c, int phase, double* seconds, int64_t* count) {
  if (!c || phase < 0 || phase >= PH_COUNT) return B200_ERR_INVALID;
  if (!c->host_only) { cudaSetDevice(c->device); c->prof.flush(); }
  if (seconds) *seconds = c->prof.seconds[phase];
  if (count) *count = c->prof.count[phase];
  return B200_OK;
}
void* b200_get_stream(b200_ctx* c) { return c ? (void*)c->stream : nullptr; }
int b200_synchronize(b200_ctx* c) {
  return guarded(c, [&]() { NEED_DEVICE(c); B200_CUDA(cudaSetDevice(c->device)); B200_CUDA(cudaStreamSynchronize(c->stream)); return (int)B200_OK; });
}

}  // extern "C"
