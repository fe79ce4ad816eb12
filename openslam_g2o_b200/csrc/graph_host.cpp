// graph_host.cpp - standalone host side above the C-ABI: .g2o text <-> SoA arrays without the g2o object graph.
//
// Mirrors, for the configured tags only (VERTEX_SE2, EDGE_SE2, VERTEX_SE3:QUAT, EDGE_SE3:QUAT, VERTEX_CAM,
// VERTEX_XYZ, EDGE_PROJECT_P2MC, FIX):
//   OptimizableGraph::load                      core/optimizable_graph.cpp:356-569
//   per-type read()                             types/slam2d/{vertex_se2,edge_se2}.cpp, types/slam3d/{vertex_se3,edge_se3}.cpp,
//                                               types/sba/types_sba.cpp:74-112,180-186,204-213
//   g2o CLI gauge + marginalisation             apps/g2o_cli/g2o.cpp:272-320, core/sparse_optimizer.cpp:116-164
//   SparseOptimizer::initializeOptimization     core/sparse_optimizer.cpp:166-267
//   OptimizableGraph::save                      core/optimizable_graph.cpp:589-622
// Host-only, pointer-free after load; everything numeric per iteration happens behind b200_ctx.
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <climits>
#include <cstring>
#include <map>
#include <memory>
#include <charconv>
#include <chrono>
#include <functional>
#include <string>
#include <thread>
#include <tr1/unordered_map>
#include <vector>

#include "../../include/g2o_b200.h"
#include "geometry.cuh"
#include "host_parallel.h"

using namespace g2o_b200;

namespace {
struct HVertex {
  int kind = 0, id = 0;
  bool fixed = false, marginalized = false, active = false;
  int hidx = -1;
  int slot = -1;  // index inside the per-kind array handed to the context
  double est[12];
};
struct HEdge {
  int param = -1;  // XYZ2UV: id of its CameraParameters
  int kind = 0;
  int v0 = 0, v1 = 0;  // indices into vertices
  bool active = false;
  double* meas = nullptr;  // emeas(kind) doubles, then info: D x D col-major - both inside b200_graph::edge_pool
  double* info = nullptr;  // (a P2MC edge is 88 bytes in total: 20M-edge inputs stay below 2 GB of host memory)
};
// bump allocator whose blocks never move (edges keep plain pointers into it)
struct DoublePool {
  std::vector<std::unique_ptr<double[]>> blocks;
  size_t used = 0, cap = 0;
  double* alloc(size_t n) {
    if (used + n > cap) {
      cap = std::max<size_t>(n, (size_t)1 << 20);
      blocks.emplace_back(new double[cap]);
      used = 0;
    }
    double* r = blocks.back().get() + used;
    used += n;
    return r;
  }
};
int vdim(int kind) { return (kind == B200_VERTEX_SE2 || kind == B200_VERTEX_XYZ) ? 3 : kind == B200_VERTEX_XY ? 2 : 6; }
int vest(int kind) { return (kind == B200_VERTEX_SE2 || kind == B200_VERTEX_XYZ) ? 3 : kind == B200_VERTEX_XY ? 2 : 12; }
int edim(int kind) { return (kind == B200_EDGE_SE2 || kind == B200_EDGE_SE3_XYZ) ? 3 : kind == B200_EDGE_SE3 ? 6 : 2; }
int emeas(int kind) { return (kind == B200_EDGE_SE2 || kind == B200_EDGE_SE3_XYZ) ? 3 : kind == B200_EDGE_SE3 ? 12 : 2; }
bool is_ba(int ekind) { return ekind == B200_EDGE_P2MC || ekind == B200_EDGE_XYZ2UV; }
bool is_landmark_edge(int ekind) { return ekind == B200_EDGE_SE2_XY || ekind == B200_EDGE_SE3_XYZ; }
// SE3Quat::inverse (types/slam3d/se3quat.h:125-130) on [t3 | q(xyzw)4]: r = conj(q), t = r * (-t) with Eigen's
// quaternion * vector (v + w uv + qv x uv, uv = 2 qv x v); no normalisation
void se3quat_inverse(const double* a, double* r) {
  const double q[4] = {-a[3], -a[4], -a[5], a[6]};
  const double v[3] = {a[0] * -1., a[1] * -1., a[2] * -1.};
  double uv[3] = {q[1] * v[2] - q[2] * v[1], q[2] * v[0] - q[0] * v[2], q[0] * v[1] - q[1] * v[0]};
  for (double& d : uv) d += d;
  const double c[3] = {q[1] * uv[2] - q[2] * uv[1], q[2] * uv[0] - q[0] * uv[2], q[0] * uv[1] - q[1] * uv[0]};
  for (int i = 0; i < 3; ++i) r[i] = v[i] + q[3] * uv[i] + c[i];
  for (int i = 0; i < 4; ++i) r[3 + i] = q[i];
}

void quat_normalize(double* q) {
  double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int i = 0; i < 4; ++i) q[i] /= n;
}
// Eigen Quaterniond(Matrix3d), for save() (toVectorQT, isometry3d_mappings.cpp:101-108)
void R_to_quat(const double* R, double* q) {
  auto m = [&](int r, int c) { return R[r + 3 * c]; };
  double t = m(0, 0) + m(1, 1) + m(2, 2);
  if (t > 0) {
    t = std::sqrt(t + 1.0); q[3] = 0.5 * t; t = 0.5 / t;
    q[0] = (m(2, 1) - m(1, 2)) * t; q[1] = (m(0, 2) - m(2, 0)) * t; q[2] = (m(1, 0) - m(0, 1)) * t;
  } else {
    int i = 0;
    if (m(1, 1) > m(0, 0)) i = 1;
    if (m(2, 2) > m(i, i)) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + 1.0);
    q[i] = 0.5 * t; t = 0.5 / t;
    q[3] = (m(k, j) - m(j, k)) * t; q[j] = (m(j, i) + m(i, j)) * t; q[k] = (m(k, i) + m(i, k)) * t;
  }
}
}  // namespace

struct b200_graph {
  std::tr1::unordered_map<int, int> idmap;  // same container family as HyperGraph::VertexIDMap (gauge order!)
  std::vector<HVertex> vertices;
  std::vector<HEdge> edges;
  DoublePool edge_pool;
  std::vector<int> active_edges;            // edge indices, internalId order
  std::vector<int> kind_slots[B200_NUM_VERTEX_KINDS];  // per kind: vertex indices handed to the context (ascending id)
  std::map<int, std::array<double, 4>> camera_parameters;  // PARAMS_CAMERAPARAMETERS id -> f cx cy baseline
  std::map<int, std::array<double, 12>> se3_offsets;       // PARAMS_SE3OFFSET id -> isometry [R | t]
  std::map<int, std::array<double, 7>> se3_offsets_text;   // ... as read (x y z qx qy qz qw), for save()
  // landmark sharding of the last upload: per XYZ slot the row handed to the context (-1: another shard owns it).
  // Empty = nothing sharded (row i of the context is slot i).  b200_graph_download scatters through it.
  std::vector<int> uploaded_lm_row;
  // per-edge robust kernels (b200_graph_set_edge_robust_kernel), indexed like `edges`; empty = none set (255 = unset entry)
  std::vector<unsigned char> edge_rk_kind;
  std::vector<double> edge_rk_delta;
  std::string err;
  int find(int id) const { auto it = idmap.find(id); return it == idmap.end() ? -1 : it->second; }
};

namespace {

int add_vertex(b200_graph* g, int kind, int id) {
  if (g->find(id) >= 0) return -1;  // HyperGraph::addVertex refuses duplicate ids
  HVertex v;
  v.kind = kind; v.id = id;
  for (double& d : v.est) d = 0;
  g->vertices.push_back(v);
  g->idmap[id] = (int)g->vertices.size() - 1;
  return (int)g->vertices.size() - 1;
}
void set_to_origin(HVertex& v) {
  for (double& d : v.est) d = 0;
  if (v.kind == B200_VERTEX_SE3) v.est[0] = v.est[4] = v.est[8] = 1;
  if (v.kind == B200_VERTEX_CAM) { v.est[6] = 1; v.est[7] = 1; v.est[8] = 1; v.est[9] = 0.5; v.est[10] = 0.5; }
  if (v.kind == B200_VERTEX_SE3_EXPMAP) v.est[6] = 1;  // SE3Quat(); intrinsics arrive with the first edge
}
bool vertex_read(HVertex& v, const double* p, int n) {
  switch (v.kind) {
    case B200_VERTEX_SE2:
    case B200_VERTEX_XYZ:
      if (n < 3) return false;
      v.est[0] = p[0]; v.est[1] = p[1]; v.est[2] = p[2];
      return true;
    case B200_VERTEX_XY:  // types/slam2d/vertex_point_xy.cpp read
      if (n < 2) return false;
      v.est[0] = p[0]; v.est[1] = p[1];
      return true;
    case B200_VERTEX_SE3: {  // fromVectorQT: quaternion used un-normalised (isometry3d_mappings.cpp:131-136)
      if (n < 7) return false;
      const double q[4] = {p[3], p[4], p[5], p[6]};
      geo::quat_to_R(q, v.est);
      v.est[9] = p[0]; v.est[10] = p[1]; v.est[11] = p[2];
      return true;
    }
    case B200_VERTEX_CAM: {  // types_sba.cpp:74-112 + SE3Quat::normalizeRotation (se3quat.h:280-285)
      if (n < 7) return false;
      v.est[0] = p[0]; v.est[1] = p[1]; v.est[2] = p[2];
      double q[4] = {p[3], p[4], p[5], p[6]};
      quat_normalize(q);
      if (q[3] < 0) for (double& d : q) d *= -1;
      quat_normalize(q);
      for (int i = 0; i < 4; ++i) v.est[3 + i] = q[i];
      if (n >= 12) for (int i = 0; i < 5; ++i) v.est[7 + i] = p[7 + i];
      else { v.est[7] = 300; v.est[8] = 300; v.est[9] = 320; v.est[10] = 320; v.est[11] = 0.1; }
      return true;
    }
    case B200_VERTEX_SE3_EXPMAP: {  // types_six_dof_expmap.cpp:76-84: the file holds cam2world (fromVector, no
      if (n < 7) return false;      // normalisation), the estimate is its inverse
      se3quat_inverse(p, v.est);
      for (int i = 7; i < 12; ++i) v.est[i] = 0;  // intrinsics: set by the first XYZ2UV edge (f = 0 marks 'unset')
      return true;
    }
  }
  return false;
}
bool edge_read(HEdge& e, const double* p, int n) {
  const int D = edim(e.kind);
  for (int i = 0; i < D * D; ++i) e.info[i] = 0;
  for (int i = 0; i < D; ++i) e.info[i + D * i] = 1;
  switch (e.kind) {
    case B200_EDGE_SE2: {
      if (n < 9) return false;
      e.meas[0] = p[0]; e.meas[1] = p[1]; e.meas[2] = p[2];
      int k = 3;
      for (int i = 0; i < 3; ++i) for (int j = i; j < 3; ++j) { e.info[i + 3 * j] = p[k]; e.info[j + 3 * i] = p[k]; ++k; }
      return true;
    }
    case B200_EDGE_SE3: {  // measurement quaternion IS normalised (edge_se3.cpp:18-20)
      if (n < 7) return false;
      double q[4] = {p[3], p[4], p[5], p[6]};
      quat_normalize(q);
      geo::quat_to_R(q, e.meas);
      e.meas[9] = p[0]; e.meas[10] = p[1]; e.meas[11] = p[2];
      int k = 7;
      for (int i = 0; i < 6 && k < n; ++i) for (int j = i; j < 6 && k < n; ++j) { e.info[i + 6 * j] = p[k]; e.info[j + 6 * i] = p[k]; ++k; }
      return true;
    }
    case B200_EDGE_P2MC:  // information forced to identity on read (types_sba.cpp:204-213)
      if (n < 2) return false;
      e.meas[0] = p[0]; e.meas[1] = p[1];
      return true;
    case B200_EDGE_XYZ2UV: {  // types_six_dof_expmap.cpp:241-256: paramId u v i00 i01 i11
      if (n < 6) return false;
      e.param = (int)p[0];
      e.meas[0] = p[1]; e.meas[1] = p[2];
      e.info[0] = p[3]; e.info[1] = e.info[2] = p[4]; e.info[3] = p[5];
      return true;
    }
    case B200_EDGE_SE2_XY: {  // types/slam2d/edge_se2_pointxy.cpp:41-47: x y i00 i01 i11
      if (n < 5) return false;
      e.meas[0] = p[0]; e.meas[1] = p[1];
      e.info[0] = p[2]; e.info[1] = e.info[2] = p[3]; e.info[3] = p[4];
      return true;
    }
    case B200_EDGE_SE3_XYZ: {  // types/slam3d/edge_se3_pointxyz.cpp:62-84: paramId x y z + upper triangle (identity if absent)
      if (n < 4) return false;
      e.param = (int)p[0];
      e.meas[0] = p[1]; e.meas[1] = p[2]; e.meas[2] = p[3];
      int k = 4;
      for (int i = 0; i < 3 && k < n; ++i) for (int j = i; j < 3 && k < n; ++j) { e.info[i + 3 * j] = p[k]; e.info[j + 3 * i] = p[k]; ++k; }
      return true;
    }
  }
  return false;
}
// EdgeSE2/EdgeSE3::initialEstimate for vertices first seen in an edge line (load with createEdges = true)
void initial_estimate(b200_graph* g, const HEdge& e, bool to_from_from) {
  HVertex& a = g->vertices[e.v0];
  HVertex& b = g->vertices[e.v1];
  if (e.kind == B200_EDGE_SE2) {
    geo::SE2 z{e.meas[0], e.meas[1], e.meas[2]};
    if (to_from_from) { geo::SE2 r = geo::se2_mul(geo::SE2{a.est[0], a.est[1], a.est[2]}, z); b.est[0] = r.x; b.est[1] = r.y; b.est[2] = r.th; }
    else { geo::SE2 r = geo::se2_mul(geo::SE2{b.est[0], b.est[1], b.est[2]}, geo::se2_inv(z)); a.est[0] = r.x; a.est[1] = r.y; a.est[2] = r.th; }
  } else if (e.kind == B200_EDGE_SE3) {
    geo::Iso Z, A, B;
    memcpy(Z.R, e.meas, 72); memcpy(Z.t, e.meas + 9, 24);
    memcpy(A.R, a.est, 72); memcpy(A.t, a.est + 9, 24);
    memcpy(B.R, b.est, 72); memcpy(B.t, b.est + 9, 24);
    if (to_from_from) { geo::Iso r = geo::iso_mul(A, Z); memcpy(b.est, r.R, 72); memcpy(b.est + 9, r.t, 24); }
    else { geo::Iso r = geo::iso_mul(B, geo::iso_inverse(Z)); memcpy(a.est, r.R, 72); memcpy(a.est + 9, r.t, 24); }
  } else if (e.kind == B200_EDGE_SE2_XY && to_from_from) {  // edge_se2_pointxy.cpp:55-64: point = pose * measurement
    double sn, cs;
    sincos(a.est[2], &sn, &cs);
    b.est[0] = cs * e.meas[0] - sn * e.meas[1] + a.est[0];
    b.est[1] = sn * e.meas[0] + cs * e.meas[1] + a.est[1];
  } else if (e.kind == B200_EDGE_SE3_XYZ && to_from_from) {  // edge_se3_pointxyz.cpp initialEstimate: pose * (offset * z)
    geo::Iso A, O;
    memcpy(A.R, a.est, 72); memcpy(A.t, a.est + 9, 24);
    const std::array<double, 12>& of = g->se3_offsets[e.param];
    memcpy(O.R, of.data(), 72); memcpy(O.t, of.data() + 9, 24);
    const geo::Iso n2w = geo::iso_mul(A, O);
    for (int r = 0; r < 3; ++r) b.est[r] = n2w.R[r] * e.meas[0] + n2w.R[r + 3] * e.meas[1] + n2w.R[r + 6] * e.meas[2] + n2w.t[r];
  }
}
int add_edge(b200_graph* g, int kind, int id1, int id2, const double* payload, int n) {
  static const int vk0[B200_NUM_EDGE_KINDS] = {B200_VERTEX_SE2, B200_VERTEX_SE3, B200_VERTEX_XYZ, B200_VERTEX_XYZ, B200_VERTEX_SE2, B200_VERTEX_SE3};
  static const int vk1[B200_NUM_EDGE_KINDS] = {B200_VERTEX_SE2, B200_VERTEX_SE3, B200_VERTEX_CAM, B200_VERTEX_SE3_EXPMAP, B200_VERTEX_XY, B200_VERTEX_XYZ};
  const std::array<double, 4>* cp = nullptr;
  if (kind == B200_EDGE_SE3_XYZ) {  // resolveParameters: unknown ParameterSE3Offset id rejects the edge
    if (n < 1 || !g->se3_offsets.count((int)payload[0])) { g->err = "SE3_XYZ edge names unknown ParameterSE3Offset"; return B200_ERR_INVALID; }
  }
  if (kind == B200_EDGE_XYZ2UV) {  // OptimizableGraph::addEdge -> resolveParameters: unknown parameter id rejects the edge
    auto it = n >= 1 ? g->camera_parameters.find((int)payload[0]) : g->camera_parameters.end();
    if (it == g->camera_parameters.end()) { g->err = "XYZ2UV edge names unknown CameraParameters"; return B200_ERR_INVALID; }
    cp = &it->second;
  }
  int a = g->find(id1), b = g->find(id2);
  int doInit = 0;
  if (a < 0) { a = add_vertex(g, vk0[kind], id1); set_to_origin(g->vertices[a]); doInit = 2; }
  if (b < 0) { b = add_vertex(g, vk1[kind], id2); set_to_origin(g->vertices[b]); doInit = 1; }
  if (g->vertices[a].kind != vk0[kind] || g->vertices[b].kind != vk1[kind]) { g->err = "edge connects vertices of the wrong type"; return B200_ERR_UNSUPPORTED; }
  HEdge e;
  e.kind = kind; e.v0 = a; e.v1 = b;
  e.meas = g->edge_pool.alloc(emeas(kind) + edim(kind) * edim(kind));
  e.info = e.meas + emeas(kind);
  if (!edge_read(e, payload, n)) { g->err = "short edge payload"; return B200_ERR_INVALID; }
  if (cp) {  // the pose row carries the intrinsics of its edges: f f cx cy baseline
    HVertex& pv = g->vertices[b];
    const double want[5] = {(*cp)[0], (*cp)[0], (*cp)[1], (*cp)[2], (*cp)[3]};
    if (pv.est[7] == 0) for (int i = 0; i < 5; ++i) pv.est[7 + i] = want[i];
    else for (int i = 0; i < 5; ++i) if (pv.est[7 + i] != want[i]) { g->err = "edges of one pose name different CameraParameters"; return B200_ERR_UNSUPPORTED; }
  }
  g->edges.push_back(e);
  if (doInit == 1) initial_estimate(g, e, true);
  if (doInit == 2) initial_estimate(g, e, false);
  return B200_OK;
}

// tokenizer over an in-memory copy of the file (20M-edge inputs: no iostreams on the hot parse path)
struct LineParser {
  const char* p;
  const char* end;
  bool next_token(const char*& b, const char*& e) {
    while (p < end && (*p == ' ' || *p == '\t' || *p == '\r')) ++p;
    if (p >= end || *p == '\n') return false;
    b = p;
    while (p < end && *p != ' ' && *p != '\t' && *p != '\n' && *p != '\r') ++p;
    e = p;
    return true;
  }
  void skip_line() {
    while (p < end && *p != '\n') ++p;
    if (p < end) ++p;
  }
};
enum { REC_VERTEX = 0, REC_EDGE = 1, REC_FIX = 2, REC_PARAMS = 3 };
struct ParsedRecord {
  int type, kind, id0, id1, n;  // n numbers at nums[off ...] (FIX: the ids; PARAMS: id0 + 4 numbers)
  size_t off;
};
struct ParsedChunk {
  std::vector<ParsedRecord> recs;
  std::vector<double> nums;
};
inline bool tag_is(const char* b, const char* e, const char* lit) {
  const size_t n = strlen(lit);
  return (size_t)(e - b) == n && memcmp(b, lit, n) == 0;
}
// correctly rounded like strtod / operator>>(double), ~3x faster; anything from_chars refuses (leading '+', hex) goes
// through strtod
inline double parse_double(const char* b, const char* e) {
  double v;
  auto r = std::from_chars(b, e, v);
  if (r.ec == std::errc() && r.ptr == e) return v;
  return strtod(b, nullptr);
}
void parse_chunk(const char* begin, const char* end, ParsedChunk& out) {
  LineParser lp{begin, end};
  out.nums.reserve((size_t)(end - begin) / 12);
  out.recs.reserve((size_t)(end - begin) / 64);
  while (lp.p < lp.end) {
    const char *tb, *te;
    if (!lp.next_token(tb, te)) { lp.skip_line(); continue; }
    if (*tb == '#') { lp.skip_line(); continue; }
    ParsedRecord r{-1, -1, 0, 0, 0, out.nums.size()};
    if (tag_is(tb, te, "EDGE_PROJECT_P2MC")) { r.type = REC_EDGE; r.kind = B200_EDGE_P2MC; }
    else if (tag_is(tb, te, "EDGE_SE3:QUAT")) { r.type = REC_EDGE; r.kind = B200_EDGE_SE3; }
    else if (tag_is(tb, te, "EDGE_SE2")) { r.type = REC_EDGE; r.kind = B200_EDGE_SE2; }
    else if (tag_is(tb, te, "EDGE_PROJECT_XYZ2UV:EXPMAP")) { r.type = REC_EDGE; r.kind = B200_EDGE_XYZ2UV; }
    else if (tag_is(tb, te, "VERTEX_XYZ")) { r.type = REC_VERTEX; r.kind = B200_VERTEX_XYZ; }
    else if (tag_is(tb, te, "VERTEX_SE3:QUAT")) { r.type = REC_VERTEX; r.kind = B200_VERTEX_SE3; }
    else if (tag_is(tb, te, "VERTEX_SE2")) { r.type = REC_VERTEX; r.kind = B200_VERTEX_SE2; }
    else if (tag_is(tb, te, "VERTEX_CAM")) { r.type = REC_VERTEX; r.kind = B200_VERTEX_CAM; }
    else if (tag_is(tb, te, "VERTEX_SE3:EXPMAP")) { r.type = REC_VERTEX; r.kind = B200_VERTEX_SE3_EXPMAP; }
    else if (tag_is(tb, te, "EDGE_SE2_XY")) { r.type = REC_EDGE; r.kind = B200_EDGE_SE2_XY; }
    else if (tag_is(tb, te, "EDGE_SE3_TRACKXYZ")) { r.type = REC_EDGE; r.kind = B200_EDGE_SE3_XYZ; }
    else if (tag_is(tb, te, "VERTEX_XY")) { r.type = REC_VERTEX; r.kind = B200_VERTEX_XY; }
    else if (tag_is(tb, te, "VERTEX_TRACKXYZ")) { r.type = REC_VERTEX; r.kind = B200_VERTEX_XYZ; }
    else if (tag_is(tb, te, "FIX")) r.type = REC_FIX;
    else if (tag_is(tb, te, "PARAMS_CAMERAPARAMETERS")) r.type = REC_PARAMS;
    else if (tag_is(tb, te, "PARAMS_SE3OFFSET")) { r.type = REC_PARAMS; r.kind = 1; }
    else { lp.skip_line(); continue; }  // unknown tags are skipped (optimizable_graph.cpp:417-423)
    const int nid = r.type == REC_EDGE ? 2 : r.type == REC_FIX ? 0 : 1;
    bool ok = true;
    for (int i = 0; i < nid; ++i) {
      if (!lp.next_token(tb, te)) { ok = false; break; }
      (i == 0 ? r.id0 : r.id1) = (int)strtol(tb, nullptr, 10);
    }
    while (ok && lp.next_token(tb, te)) out.nums.push_back(r.type == REC_FIX ? (double)strtol(tb, nullptr, 10) : parse_double(tb, te));
    lp.skip_line();
    if (!ok) { out.nums.resize(r.off); continue; }
    r.n = (int)(out.nums.size() - r.off);
    out.recs.push_back(r);
  }
}
}  // namespace

extern "C" {

int b200_graph_create(b200_graph** out) {
  if (!out) return B200_ERR_INVALID;
  *out = new b200_graph();
  return B200_OK;
}
void b200_graph_destroy(b200_graph* g) { delete g; }
const char* b200_graph_last_error(const b200_graph* g) { return g ? g->err.c_str() : ""; }

int b200_graph_add_vertex(b200_graph* g, int kind, int id, const double* payload, int n) {
  if (!g || kind < 0 || kind >= B200_NUM_VERTEX_KINDS) return B200_ERR_INVALID;
  int v = add_vertex(g, kind, id);
  if (v < 0) { g->err = "duplicate vertex id"; return B200_ERR_INVALID; }
  if (!vertex_read(g->vertices[v], payload, n)) { g->err = "short vertex payload"; return B200_ERR_INVALID; }
  return B200_OK;
}
int b200_graph_add_edge(b200_graph* g, int kind, int id1, int id2, const double* payload, int n) {
  if (!g || kind < 0 || kind >= B200_NUM_EDGE_KINDS) return B200_ERR_INVALID;
  return add_edge(g, kind, id1, id2, payload, n);
}
// bulk variants: payload is row-major [n x stride]
int b200_graph_add_vertices(b200_graph* g, int kind, int n, const int32_t* ids, const double* payload, int stride) {
  if (!g || kind < 0 || kind >= B200_NUM_VERTEX_KINDS || n < 0 || !ids || !payload) return B200_ERR_INVALID;
  g->vertices.reserve(g->vertices.size() + n);
  for (int i = 0; i < n; ++i) {
    int rc = b200_graph_add_vertex(g, kind, ids[i], payload + (size_t)i * stride, stride);
    if (rc) return rc;
  }
  return B200_OK;
}
int b200_graph_add_edges(b200_graph* g, int kind, int n, const int32_t* id1, const int32_t* id2, const double* payload, int stride) {
  if (!g || kind < 0 || kind >= B200_NUM_EDGE_KINDS || n < 0 || !id1 || !id2 || !payload) return B200_ERR_INVALID;
  g->edges.reserve(g->edges.size() + n);
  for (int i = 0; i < n; ++i) {
    int rc = add_edge(g, kind, id1[i], id2[i], payload + (size_t)i * stride, stride);
    if (rc) return rc;
  }
  return B200_OK;
}
int b200_graph_add_camera_parameters(b200_graph* g, int id, double focal_length, double cx, double cy, double baseline) {
  if (!g) return B200_ERR_INVALID;
  if (g->camera_parameters.count(id)) { g->err = "duplicate parameter id"; return B200_ERR_INVALID; }  // ParameterContainer::addParameter
  g->camera_parameters[id] = {focal_length, cx, cy, baseline};
  return B200_OK;
}
int b200_graph_add_se3_offset(b200_graph* g, int id, const double* o) {
  if (!g || !o) return B200_ERR_INVALID;
  if (g->se3_offsets.count(id)) { g->err = "duplicate parameter id"; return B200_ERR_INVALID; }
  double q[4] = {o[3], o[4], o[5], o[6]};
  quat_normalize(q);  // parameter_se3_offset.cpp:52-53
  std::array<double, 12> iso;
  geo::quat_to_R(q, iso.data());
  iso[9] = o[0]; iso[10] = o[1]; iso[11] = o[2];
  g->se3_offsets[id] = iso;
  g->se3_offsets_text[id] = {o[0], o[1], o[2], q[0], q[1], q[2], q[3]};
  return B200_OK;
}
int b200_graph_set_edge_robust_kernel(b200_graph* g, int edge_index, int kind, double delta) {
  if (!g || edge_index < 0 || edge_index >= (int)g->edges.size() || kind < B200_ROBUST_NONE || kind > B200_ROBUST_DCS ||
      (kind != B200_ROBUST_NONE && !(delta > 0.0))) return B200_ERR_INVALID;
  if (g->edge_rk_kind.size() != g->edges.size()) { g->edge_rk_kind.resize(g->edges.size(), 255); g->edge_rk_delta.resize(g->edges.size(), 1.0); }
  g->edge_rk_kind[edge_index] = (unsigned char)kind;
  g->edge_rk_delta[edge_index] = delta;
  return B200_OK;
}
int b200_graph_set_fixed(b200_graph* g, int id, int fixed) {
  if (!g) return B200_ERR_INVALID;
  int v = g->find(id);
  if (v < 0) { g->err = "unable to fix vertex: not found"; return B200_ERR_INVALID; }
  g->vertices[v].fixed = fixed != 0;
  return B200_OK;
}

int b200_graph_load(b200_graph* g, const char* path) {
  if (!g || !path) return B200_ERR_INVALID;
  FILE* f = fopen(path, "rb");
  if (!f) { g->err = std::string("cannot open ") + path; return B200_ERR_INVALID; }
  long sz = -1;
  if (fseek(f, 0, SEEK_END) == 0) sz = ftell(f);
  if (sz < 0 || sz == LONG_MAX || fseek(f, 0, SEEK_SET) != 0) {  // a directory, a pipe, ...
    fclose(f);
    g->err = std::string("cannot read ") + path + " (not a regular file)";
    return B200_ERR_INVALID;
  }
  std::vector<char> buf;
  try {
    buf.resize((size_t)sz + 1);
  } catch (const std::exception&) {
    fclose(f);
    g->err = std::string("out of memory reading ") + path;
    return B200_ERR_INVALID;
  }
  size_t rd = fread(buf.data(), 1, (size_t)sz, f);
  const bool read_error = ferror(f) != 0;
  fclose(f);
  if (read_error) { g->err = std::string("read error on ") + path; return B200_ERR_INVALID; }
  buf[rd] = '\n';
  const char* const base = buf.data();
  const char* const end = base + rd + 1;
  // Peak host memory: the file image + the parsed numbers (8 bytes per number) + 32 bytes per record, next to the graph
  // itself - roughly 3x the file size for bundle-adjustment inputs.
  // phase 1 (parallel): the text is cut at line starts into one chunk per thread; every chunk is tokenised into records
  // (tag, ids, numbers).  phase 2 (sequential, file order): the records are applied to the graph - id map inserts,
  // vertices created by an edge that precedes their VERTEX line, duplicates, FIX - exactly like a line-by-line reader.
  int nthreads = g2o_b200::host_threads();  // cores this process may use, at most 16
  if (const char* e = getenv("G2O_B200_LOADER_THREADS")) nthreads = std::max(1, atoi(e));
  nthreads = (int)std::max<size_t>(1, std::min<size_t>(nthreads, rd / ((size_t)1 << 20) + 1));  // >= 1 MiB per thread
  std::vector<const char*> cut(nthreads + 1);
  cut[0] = base;
  cut[nthreads] = end;
  for (int t = 1; t < nthreads; ++t) {
    const char* p = base + (size_t)((double)rd * t / nthreads);
    p = std::max(p, cut[t - 1]);
    while (p < end && *p != '\n') ++p;
    cut[t] = p < end ? p + 1 : end;
  }
  const bool verbose = getenv("G2O_B200_LOADER_VERBOSE") != nullptr;
  auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t0 = now();
  std::vector<ParsedChunk> chunks(nthreads);
  g2o_b200::parallel_ranges((size_t)nthreads, nthreads, [&](int, size_t tb, size_t te) {
    for (size_t t = tb; t < te; ++t) parse_chunk(cut[t], cut[t + 1], chunks[t]);
  });
  const double t1 = now();
  size_t nv = 0, ne = 0;
  for (const ParsedChunk& c : chunks) for (const ParsedRecord& r : c.recs) { if (r.type == REC_VERTEX) ++nv; else if (r.type == REC_EDGE) ++ne; }
  g->vertices.reserve(g->vertices.size() + nv);
  g->edges.reserve(g->edges.size() + ne);
  // no idmap.rehash() here: the gauge is the first max-dimension vertex in VertexIDMap ITERATION order
  // (b200_graph_setup_cli), and that order depends on the bucket count - the map must grow exactly like the
  // reference's naturally grown tr1::unordered_map (and like this graph's own add_vertices path)
  for (const ParsedChunk& c : chunks)
    for (const ParsedRecord& r : c.recs) {
      const double* nums = c.nums.data() + r.off;
      switch (r.type) {
        case REC_FIX:
          for (int i = 0; i < r.n; ++i) { int v = g->find((int)nums[i]); if (v >= 0) g->vertices[v].fixed = true; }
          break;
        case REC_PARAMS:  // optimizable_graph.cpp:398-415 + CameraParameters::read
          if (r.kind == 1) { if (r.n >= 7) b200_graph_add_se3_offset(g, r.id0, nums); }
          else if (r.n >= 4) b200_graph_add_camera_parameters(g, r.id0, nums[0], nums[1], nums[2], nums[3]);
          break;
        case REC_VERTEX: {
          int v = add_vertex(g, r.kind, r.id0);
          if (v >= 0) vertex_read(g->vertices[v], nums, r.n);
          break;
        }
        case REC_EDGE: {
          int rc = add_edge(g, r.kind, r.id0, r.id1, nums, r.n);
          if (rc == B200_ERR_UNSUPPORTED) return rc;
          break;
        }
      }
    }
  if (verbose) fprintf(stderr, "b200_graph_load: %d threads, tokenise %.3f s, apply %.3f s\n", nthreads, t1 - t0, now() - t1);
  return B200_OK;
}

int b200_graph_setup_cli(b200_graph* g, int requires_marginalize) {
  if (!g || g->vertices.empty()) return -2;
  int maxDim = 0, minDim = 1 << 30;
  for (const HVertex& v : g->vertices) { maxDim = std::max(maxDim, vdim(v.kind)); minDim = std::min(minDim, vdim(v.kind)); }
  // gaugeFreedom(): false as soon as a max-dimension vertex is fixed (no unary priors among the configured edges)
  bool gaugeFreedom = true;
  for (const HVertex& v : g->vertices) if (vdim(v.kind) == maxDim && v.fixed) { gaugeFreedom = false; break; }
  int ret = -1;
  if (gaugeFreedom) {
    // findGauge(): first max-dimension vertex in VertexIDMap iteration order
    for (auto it = g->idmap.begin(); it != g->idmap.end(); ++it) {
      HVertex& v = g->vertices[it->second];
      if (vdim(v.kind) == maxDim) { v.fixed = true; ret = v.id; break; }
    }
  }
  if (requires_marginalize && maxDim != minDim)
    for (HVertex& v : g->vertices) if (vdim(v.kind) != maxDim) v.marginalized = true;
  return ret;
}

int b200_graph_initialize(b200_graph* g) {
  if (!g) return B200_ERR_INVALID;
  if (g->edges.empty()) { g->err = "attempt to initialize an empty graph"; return B200_ERR_INVALID; }
  for (HVertex& v : g->vertices) { v.active = false; v.hidx = -1; v.slot = -1; }
  g->active_edges.clear();
  for (size_t k = 0; k < g->edges.size(); ++k) {
    HEdge& e = g->edges[k];
    e.active = !(g->vertices[e.v0].fixed && g->vertices[e.v1].fixed);
    if (e.active) { g->active_edges.push_back((int)k); g->vertices[e.v0].active = true; g->vertices[e.v1].active = true; }
  }
  std::vector<int> act;
  for (size_t i = 0; i < g->vertices.size(); ++i) if (g->vertices[i].active) act.push_back((int)i);
  std::sort(act.begin(), act.end(), [&](int a, int b) { return g->vertices[a].id < g->vertices[b].id; });
  if (act.empty()) { g->err = "no active vertices"; return B200_ERR_INVALID; }
  int idx = 0;
  for (int k = 0; k < 2; ++k)
    for (int i : act) {
      HVertex& v = g->vertices[i];
      if (!v.fixed && (int)v.marginalized == k) v.hidx = idx++;
    }
  for (int k = 0; k < B200_NUM_VERTEX_KINDS; ++k) g->kind_slots[k].clear();
  for (int i : act) { HVertex& v = g->vertices[i]; v.slot = (int)g->kind_slots[v.kind].size(); g->kind_slots[v.kind].push_back(i); }
  return B200_OK;
}

int b200_graph_counts(b200_graph* g, int32_t* vc, int32_t* ec) {
  if (!g) return B200_ERR_INVALID;
  if (vc) { for (int k = 0; k < B200_NUM_VERTEX_KINDS; ++k) vc[k] = 0; for (const HVertex& v : g->vertices) vc[v.kind]++; }
  if (ec) { for (int k = 0; k < B200_NUM_EDGE_KINDS; ++k) ec[k] = 0; for (const HEdge& e : g->edges) ec[e.kind]++; }
  return B200_OK;
}

// per-edge robust kernels of the edges just handed over (`sel`: their indices into g->edges, in hand-over order); edges
// without an entry of their own carry "no kernel"
static int upload_edge_kernels(b200_graph* g, b200_ctx* ctx, int ekind, const std::vector<int>& sel) {
  if (g->edge_rk_kind.empty()) return B200_OK;
  bool any = false;
  for (int k : sel) if (k < (int)g->edge_rk_kind.size() && g->edge_rk_kind[k] != 255) { any = true; break; }
  if (!any) return B200_OK;
  std::vector<uint8_t> kd(sel.size());
  std::vector<double> dl(sel.size());
  for (size_t i = 0; i < sel.size(); ++i) {
    const int k = sel[i];
    const bool has = k < (int)g->edge_rk_kind.size() && g->edge_rk_kind[k] != 255;
    kd[i] = has ? g->edge_rk_kind[k] : (uint8_t)B200_ROBUST_NONE;
    dl[i] = has ? g->edge_rk_delta[k] : 1.0;
  }
  int rc = b200_set_edge_robust_kernels(ctx, ekind, (int)sel.size(), kd.data(), dl.data());
  if (rc) g->err = b200_last_error(ctx);
  return rc;
}

int b200_graph_upload(b200_graph* g, b200_ctx* ctx, int shard, int num_shards) {
  if (!g || !ctx || num_shards < 1 || shard < 0 || shard >= num_shards) return B200_ERR_INVALID;
  if (g->active_edges.empty()) { g->err = "call b200_graph_initialize first"; return B200_ERR_INVALID; }
  // one pose-pose / projection edge kind, and - landmark SLAM - one pose-landmark kind beside it
  int ekind = -1, lkind = -1;
  for (int k : g->active_edges) {
    const int kd = g->edges[k].kind;
    int& slot = is_landmark_edge(kd) ? lkind : ekind;
    if (slot < 0) slot = kd;
    else if (slot != kd) { g->err = "mixed edge types are not supported"; return B200_ERR_UNSUPPORTED; }
  }
  if (lkind >= 0 && ekind >= 0 && ekind != (lkind == B200_EDGE_SE2_XY ? B200_EDGE_SE2 : B200_EDGE_SE3)) { g->err = "mixed edge types are not supported"; return B200_ERR_UNSUPPORTED; }
  const bool ba = ekind >= 0 && is_ba(ekind);
  if (num_shards > 1 && !ba) { g->err = "only bundle adjustment shards (pose graphs stay single-GPU)"; return B200_ERR_UNSUPPORTED; }
  // numPoses = #free non-marginalized vertices
  int np = 0;
  for (const HVertex& v : g->vertices) if (v.hidx >= 0 && !v.marginalized) ++np;
  // ---- landmark sharding: contiguous landmark-index ranges balanced by edge count
  std::vector<int> lm_shard;  // per XYZ slot
  std::vector<int> local_slot, local_hidx;
  if (ba) {
    const std::vector<int>& ls = g->kind_slots[B200_VERTEX_XYZ];
    lm_shard.assign(ls.size(), 0);
    if (num_shards > 1) {
      std::vector<long long> deg(ls.size(), 0);
      for (int k : g->active_edges) deg[g->vertices[g->edges[k].v0].slot]++;
      std::vector<int> order(ls.size());
      for (size_t i = 0; i < ls.size(); ++i) order[i] = (int)i;
      std::sort(order.begin(), order.end(), [&](int a, int b) {
        int ha = g->vertices[ls[a]].hidx, hb = g->vertices[ls[b]].hidx;
        if ((ha < 0) != (hb < 0)) return hb < 0;  // free landmarks first, by hessian index
        return ha != hb ? ha < hb : a < b;
      });
      long long total = 0;
      for (long long d : deg) total += d;
      long long acc = 0;
      for (int i : order) {
        int sidx = (int)std::min<long long>(num_shards - 1, acc * num_shards / std::max<long long>(total, 1));
        if (g->vertices[ls[i]].hidx < 0) sidx = 0;
        lm_shard[i] = sidx;
        acc += deg[i];
      }
    }
  }
  for (int kind = 0; kind < B200_NUM_VERTEX_KINDS; ++kind) {
    const std::vector<int>& sl = g->kind_slots[kind];
    if (sl.empty()) continue;
    const int ne = vest(kind);
    std::vector<double> est;
    std::vector<int32_t> hidx;
    std::vector<uint8_t> marg;
    if (kind == B200_VERTEX_XYZ) g->uploaded_lm_row.clear();
    if (kind == B200_VERTEX_XYZ && ba) {
      local_slot.assign(sl.size(), -1);
      int nloc = 0, nfree = 0;
      for (size_t i = 0; i < sl.size(); ++i) {
        if (lm_shard[i] != shard) continue;
        const HVertex& v = g->vertices[sl[i]];
        local_slot[i] = nloc++;
        est.insert(est.end(), v.est, v.est + ne);
        hidx.push_back(v.hidx >= 0 ? np + nfree++ : -1);
        marg.push_back(v.marginalized ? 1 : 0);
      }
      if (num_shards == 1) {  // keep g2o's own numbering when nothing is sharded
        for (size_t i = 0, q = 0; i < sl.size(); ++i) if (lm_shard[i] == shard) hidx[q++] = g->vertices[sl[i]].hidx;
      } else {
        g->uploaded_lm_row = local_slot;
      }
    } else {
      for (int i : sl) {
        const HVertex& v = g->vertices[i];
        est.insert(est.end(), v.est, v.est + ne);
        hidx.push_back(v.hidx);
        marg.push_back(v.marginalized ? 1 : 0);
      }
    }
    int rc = b200_set_vertices(ctx, kind, (int)hidx.size(), est.data(), hidx.data(), marg.data());
    if (rc) { g->err = b200_last_error(ctx); return rc; }
  }
  if (lkind >= 0) {
    // pose-landmark edges of a landmark-SLAM graph, active-edge order; one ParameterSE3Offset value for all of them
    const int D = edim(lkind), nm = emeas(lkind);
    std::vector<int32_t> vi, vj;
    std::vector<double> meas, info;
    std::vector<int> sel;
    const std::array<double, 12>* off = nullptr;
    for (int k : g->active_edges) {
      const HEdge& e = g->edges[k];
      if (e.kind != lkind) continue;
      sel.push_back(k);
      vi.push_back(g->vertices[e.v0].slot); vj.push_back(g->vertices[e.v1].slot);
      meas.insert(meas.end(), e.meas, e.meas + nm);
      info.insert(info.end(), e.info, e.info + D * D);
      if (lkind == B200_EDGE_SE3_XYZ) {
        const std::array<double, 12>& o = g->se3_offsets[e.param];
        if (!off) off = &o;
        else if (o != *off) { g->err = "SE3_XYZ edges name ParameterSE3Offsets with different values"; return B200_ERR_UNSUPPORTED; }
      }
    }
    if (off) { int rc = b200_set_sensor_offset(ctx, off->data()); if (rc) { g->err = b200_last_error(ctx); return rc; } }
    int rc = b200_set_edges(ctx, lkind, (int)vi.size(), vi.data(), vj.data(), meas.data(), info.data());
    if (rc) { g->err = b200_last_error(ctx); return rc; }
    if ((rc = upload_edge_kernels(g, ctx, lkind, sel)) != B200_OK) return rc;
    if (ekind < 0) {  // no odometry at all: empty pose-pose set of the matching kind
      rc = b200_set_edges(ctx, lkind == B200_EDGE_SE2_XY ? B200_EDGE_SE2 : B200_EDGE_SE3, 0, nullptr, nullptr, nullptr, nullptr);
      if (rc) { g->err = b200_last_error(ctx); return rc; }
    }
  } else {
    int rc = b200_set_edges(ctx, B200_EDGE_SE2_XY, 0, nullptr, nullptr, nullptr, nullptr);  // forget a previous graph's set
    if (rc) { g->err = b200_last_error(ctx); return rc; }
  }
  if (ekind >= 0) {
    const int D = edim(ekind), nm = emeas(ekind);
    std::vector<int32_t> vi, vj;
    std::vector<double> meas, info;
    std::vector<int32_t> xr, xc;  // Hschur blocks contributed by other shards
    std::vector<int> vslot(g->vertices.size());  // compact copy: the edge loop gathers two slots per edge
    for (size_t i = 0; i < vslot.size(); ++i) vslot[i] = g->vertices[i].slot;
    // this shard's edges in active-edge order: counted per contiguous range, offset by a prefix sum, copied concurrently
    const size_t nact = g->active_edges.size();
    auto mine = [&](const HEdge& e) { return e.kind == ekind && (!ba || lm_shard[vslot[e.v0]] == shard); };
    const int parts = g2o_b200::range_count(nact, (size_t)1 << 16);
    std::vector<size_t> base(parts + 1, 0);
    g2o_b200::parallel_ranges(nact, parts, [&](int t, size_t b, size_t e2) {
      size_t cnt = 0;
      for (size_t q = b; q < e2; ++q) cnt += mine(g->edges[g->active_edges[q]]);
      base[t + 1] = cnt;
    });
    for (int t = 0; t < parts; ++t) base[t + 1] += base[t];
    const size_t na = base[parts];
    vi.resize(na); vj.resize(na); meas.resize(na * nm); info.resize(na * D * D);
    g2o_b200::parallel_ranges(nact, parts, [&](int t, size_t b, size_t e2) {
      size_t o = base[t];
      for (size_t q = b; q < e2; ++q) {
        const HEdge& e = g->edges[g->active_edges[q]];
        if (!mine(e)) continue;
        const int s0 = vslot[e.v0];
        vi[o] = ba ? local_slot[s0] : s0;
        vj[o] = vslot[e.v1];
        memcpy(&meas[o * nm], e.meas, nm * sizeof(double));
        memcpy(&info[o * D * D], e.info, (size_t)D * D * sizeof(double));
        ++o;
      }
    });
    if (ba && num_shards > 1) {
      // per foreign landmark: the camera pairs it couples
      std::vector<std::vector<int>> cams(g->kind_slots[B200_VERTEX_XYZ].size());
      for (int k : g->active_edges) {
        const HEdge& e = g->edges[k];
        const HVertex& p = g->vertices[e.v0];
        const HVertex& c = g->vertices[e.v1];
        if (lm_shard[p.slot] == shard || p.hidx < 0 || c.hidx < 0) continue;
        cams[p.slot].push_back(c.hidx);
      }
      std::vector<long long> keys;
      for (auto& cl : cams) {
        std::sort(cl.begin(), cl.end());
        cl.erase(std::unique(cl.begin(), cl.end()), cl.end());
        for (size_t a = 0; a < cl.size(); ++a) for (size_t b = a; b < cl.size(); ++b) keys.push_back(((long long)cl[b] << 32) | cl[a]);
        if (keys.size() > ((size_t)1 << 24)) { std::sort(keys.begin(), keys.end()); keys.erase(std::unique(keys.begin(), keys.end()), keys.end()); }
      }
      std::sort(keys.begin(), keys.end());
      keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
      for (long long kk : keys) { xc.push_back((int)(kk >> 32)); xr.push_back((int)(kk & 0xffffffff)); }
      int rc = b200_add_schur_pattern(ctx, (int)xr.size(), xr.data(), xc.data());
      if (rc) return rc;
    }
    int rc = b200_set_edges(ctx, ekind, (int)vi.size(), vi.data(), vj.data(), meas.data(), info.data());
    if (rc) { g->err = b200_last_error(ctx); return rc; }
    if (!g->edge_rk_kind.empty()) {  // this shard's edges of the main set, in hand-over order
      std::vector<int> sel;
      sel.reserve(na);
      for (size_t q = 0; q < nact; ++q) if (mine(g->edges[g->active_edges[q]])) sel.push_back(g->active_edges[q]);
      if ((rc = upload_edge_kernels(g, ctx, ekind, sel)) != B200_OK) return rc;
    }
  }
  return B200_OK;
}

int b200_graph_download(b200_graph* g, b200_ctx* ctx) {
  if (!g || !ctx) return B200_ERR_INVALID;
  for (int kind = 0; kind < B200_NUM_VERTEX_KINDS; ++kind) {
    const std::vector<int>& sl = g->kind_slots[kind];
    if (sl.empty()) continue;
    const int ne = vest(kind);
    std::vector<double> est(sl.size() * ne);
    int rc = b200_get_estimates(ctx, kind, est.data());
    if (rc) { g->err = b200_last_error(ctx); return rc; }
    if (kind == B200_VERTEX_XYZ && !g->uploaded_lm_row.empty()) {
      // landmark-sharded upload: the context holds only this shard's landmarks, in local row order; the landmarks of
      // the other shards keep their host estimates (their owners hold the optimised values)
      if (g->uploaded_lm_row.size() != sl.size()) { g->err = "graph changed since the sharded upload"; return B200_ERR_INVALID; }
      for (size_t i = 0; i < sl.size(); ++i) {
        const int row = g->uploaded_lm_row[i];
        if (row >= 0) memcpy(g->vertices[sl[i]].est, &est[(size_t)row * ne], ne * sizeof(double));
      }
      continue;
    }
    for (size_t i = 0; i < sl.size(); ++i) memcpy(g->vertices[sl[i]].est, &est[i * ne], ne * sizeof(double));
  }
  return B200_OK;
}

int b200_graph_get_estimate(b200_graph* g, int id, double* out) {
  if (!g || !out) return B200_ERR_INVALID;
  int v = g->find(id);
  if (v < 0) return B200_ERR_INVALID;
  const int ne = vest(g->vertices[v].kind);
  memcpy(out, g->vertices[v].est, ne * sizeof(double));
  return ne;
}

int b200_graph_get_vertex_info(b200_graph* g, int id, int32_t* out /* kind, hidx, fixed, marginalized */) {
  if (!g || !out) return B200_ERR_INVALID;
  int v = g->find(id);
  if (v < 0) return B200_ERR_INVALID;
  const HVertex& x = g->vertices[v];
  out[0] = x.kind; out[1] = x.hidx; out[2] = x.fixed; out[3] = x.marginalized;
  return B200_OK;
}

int b200_graph_save(b200_graph* g, const char* path) {
  if (!g || !path) return B200_ERR_INVALID;
  FILE* f = fopen(path, "w");
  if (!f) { g->err = std::string("cannot write ") + path; return B200_ERR_INVALID; }
  std::vector<int> order(g->vertices.size());
  for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
  std::sort(order.begin(), order.end(), [&](int a, int b) { return g->vertices[a].id < g->vertices[b].id; });
  static const char* vtag[B200_NUM_VERTEX_KINDS] = {"VERTEX_SE2", "VERTEX_SE3:QUAT", "VERTEX_CAM", "VERTEX_XYZ", "VERTEX_SE3:EXPMAP", "VERTEX_XY"};
  static const char* etag[B200_NUM_EDGE_KINDS] = {"EDGE_SE2", "EDGE_SE3:QUAT", "EDGE_PROJECT_P2MC", "EDGE_PROJECT_XYZ2UV:EXPMAP", "EDGE_SE2_XY", "EDGE_SE3_TRACKXYZ"};
  bool track = false;  // XYZ points of a landmark-SLAM graph are VertexPointXYZ (VERTEX_TRACKXYZ)
  for (const HEdge& e : g->edges) if (e.kind == B200_EDGE_SE3_XYZ) { track = true; break; }
  for (const auto& kv : g->se3_offsets_text) {
    fprintf(f, "PARAMS_SE3OFFSET %d", kv.first);
    for (double d : kv.second) fprintf(f, " %.17g", d);
    fprintf(f, "\n");
  }
  for (const auto& kv : g->camera_parameters)  // parameters first (optimizable_graph.cpp:591-594)
    fprintf(f, "PARAMS_CAMERAPARAMETERS %d %.17g %.17g %.17g %.17g\n", kv.first, kv.second[0], kv.second[1], kv.second[2], kv.second[3]);
  for (int i : order) {
    const HVertex& v = g->vertices[i];
    fprintf(f, "%s %d", (track && v.kind == B200_VERTEX_XYZ) ? "VERTEX_TRACKXYZ" : vtag[v.kind], v.id);
    if (v.kind == B200_VERTEX_SE2 || v.kind == B200_VERTEX_XYZ) fprintf(f, " %.17g %.17g %.17g", v.est[0], v.est[1], v.est[2]);
    else if (v.kind == B200_VERTEX_XY) fprintf(f, " %.17g %.17g", v.est[0], v.est[1]);
    else if (v.kind == B200_VERTEX_SE3) {
      double q[4];
      R_to_quat(v.est, q);
      quat_normalize(q);
      fprintf(f, " %.17g %.17g %.17g %.17g %.17g %.17g %.17g", v.est[9], v.est[10], v.est[11], q[0], q[1], q[2], q[3]);
    } else if (v.kind == B200_VERTEX_SE3_EXPMAP) {  // VertexSE3Expmap::write: cam2world = estimate^-1
      double c2w[7];
      se3quat_inverse(v.est, c2w);
      for (int k = 0; k < 7; ++k) fprintf(f, " %.17g", c2w[k]);
    } else {
      for (int k = 0; k < 12; ++k) fprintf(f, " %.17g", v.est[k]);
    }
    fprintf(f, "\n");
    if (v.fixed) fprintf(f, "FIX %d\n", v.id);
  }
  for (const HEdge& e : g->edges) {
    fprintf(f, "%s %d %d", etag[e.kind], g->vertices[e.v0].id, g->vertices[e.v1].id);
    if (e.kind == B200_EDGE_XYZ2UV || e.kind == B200_EDGE_SE3_XYZ) fprintf(f, " %d", e.param);
    const int D = edim(e.kind);
    if (e.kind == B200_EDGE_SE2 || e.kind == B200_EDGE_SE3_XYZ) fprintf(f, " %.17g %.17g %.17g", e.meas[0], e.meas[1], e.meas[2]);
    else if (e.kind == B200_EDGE_SE3) {
      double q[4];
      R_to_quat(e.meas, q);
      quat_normalize(q);
      fprintf(f, " %.17g %.17g %.17g %.17g %.17g %.17g %.17g", e.meas[9], e.meas[10], e.meas[11], q[0], q[1], q[2], q[3]);
    } else fprintf(f, " %.17g %.17g", e.meas[0], e.meas[1]);
    if (e.kind != B200_EDGE_P2MC)
      for (int i = 0; i < D; ++i) for (int j = i; j < D; ++j) fprintf(f, " %.17g", e.info[i + D * j]);
    fprintf(f, "\n");
  }
  fclose(f);
  return B200_OK;
}

}  // extern "C"
