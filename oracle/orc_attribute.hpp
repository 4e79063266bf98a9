// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_core.hpp header).
// Attribute path: portabilization (quantizers), prediction schemes, prediction
// transforms and the per-attribute encoder loop, restated from the reference.
// Paths relative to /root/reference/draco-oxide/src/.
//
// Build with -O2 -ffp-contract=off -fno-fast-math so every f32 operation is a
// separately rounded IEEE operation, like rustc emits.
#pragma once
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include "orc_connectivity.hpp"

namespace orc {

// ORC_TIMING=1 prints per-stage wall clock to stderr (used to split the CPU baseline
// into host-only and attribute stages).
struct StageTimer {
  bool on; std::chrono::steady_clock::time_point t;
  StageTimer() : on(getenv("ORC_TIMING") != nullptr), t(std::chrono::steady_clock::now()) {}
  void lap(const char* what) {
    if (!on) return;
    auto n = std::chrono::steady_clock::now();
    fprintf(stderr, "[orc] %-28s %9.3f ms\n", what, std::chrono::duration<double, std::milli>(n - t).count());
    t = n;
  }
};

struct OracleConfig {
  unsigned position_bits = 11, texcoord_bits = 10, generic_bits = 11;
  bool literal = false;  // true: O(V^2) `contains` and literal stack removal, for small-mesh cross checks
};

// Rust `f32 as i64` / `f32 as i32`: truncate toward zero, saturate, NaN -> 0.
inline int64_t f32_as_i64(float f) {
  if (f != f) return 0;
  if (f >= 9223372036854775808.0f) return std::numeric_limits<int64_t>::max();
  if (f <= -9223372036854775808.0f) return std::numeric_limits<int64_t>::min();
  return (int64_t)f;
}
inline int32_t f32_as_i32(float f) {
  if (f != f) return 0;
  if (f >= 2147483648.0f) return std::numeric_limits<int32_t>::max();
  if (f <= -2147483648.0f) return std::numeric_limits<int32_t>::min();
  return (int32_t)f;
}
inline int32_t wrap_i32(int64_t v) { return (int32_t)(uint32_t)(uint64_t)v; }  // Rust `as i32`
inline int32_t add32(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }  // release-mode wrapping
inline int32_t sub32(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }
inline int32_t mul32(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }
inline int64_t add64(int64_t a, int64_t b) { return (int64_t)((uint64_t)a + (uint64_t)b); }
inline int64_t sub64(int64_t a, int64_t b) { return (int64_t)((uint64_t)a - (uint64_t)b); }
inline int64_t mul64(int64_t a, int64_t b) { return (int64_t)((uint64_t)a * (uint64_t)b); }
inline int64_t abs64(int64_t a) { return a < 0 ? (int64_t)(0 - (uint64_t)a) : a; }
inline int32_t abs32(int32_t a) { return a < 0 ? (int32_t)(0 - (uint32_t)a) : a; }
inline int32_t signum32(int32_t a) { return a > 0 ? 1 : (a < 0 ? -1 : 0); }
inline int64_t div64(int64_t a, int64_t b) {  // Rust `/` on i64: truncating; b==0 panics
  if (b == 0) throw Panic(ST_INTERNAL, "attempt to divide by zero");
  if (a == std::numeric_limits<int64_t>::min() && b == -1) throw Panic(ST_INTERNAL, "attempt to divide with overflow");
  return a / b;
}

// ---------------------------------------------------------------------------
// geom.rs

// octahedral_transform for f32 input — encode/attribute/prediction_transform/geom.rs:40-91
inline void octahedral_transform_f32(float x, float y, float z, float& ou, float& ov) {
  if (x == 0.0f && y == 0.0f && z == 0.0f)
    throw Panic(ST_ZERO_NORMAL, "Zero vector cannot be transformed to octahedron space as it is not a unit vector.");
  float abs_sum = (std::fabs(x) + std::fabs(y)) + std::fabs(z);
  float u = y / abs_sum;
  float v = z / abs_sum;
  if (x < 0.0f) {
    float u_out = (u < 0.0f) ? (std::fabs(v) - 1.0f) : (1.0f - std::fabs(v));
    float v_out = (v < 0.0f) ? (std::fabs(u) - 1.0f) : (1.0f - std::fabs(u));
    u = u_out;
    v = v_out;
  }
  ou = u;
  ov = v;
}
// integer input: converted component-wise via f64 -> f32; `float_v.normalize()` discards
// its result (geom.rs:47-55, Appendix B.5)
inline void octahedral_transform_i32(int32_t x, int32_t y, int32_t z, float& ou, float& ov) {
  if (x == 0 && y == 0 && z == 0)
    throw Panic(ST_ZERO_NORMAL, "Zero vector cannot be transformed to octahedron space as it is not a unit vector.");
  octahedral_transform_f32((float)(double)x, (float)(double)y, (float)(double)z, ou, ov);
}
// into_faithful_oct_quantization — geom.rs:137-157
inline void into_faithful_oct_quantization(int32_t& x, int32_t& y) {
  const int32_t max = 255, half = max / 2;
  int32_t u = x, v = y;
  if ((u == 0 && v == 0) || (u == 255 && v == 0) || (u == 0 && v == 255)) { x = 255; y = 255; return; }
  else if (u == 0 && v > 127) y = half - (v - half);
  else if (u == max && v < half) y = half + (half - v);
  else if (v == max && u < half) x = half + (half - u);
  else if (v == 0 && u > half) x = half - (u - half);
}
// (octahedral_transform(v) + [1,1]) * 127 as i32, then faithful fix-up —
// octahedral_quantization.rs:49-64, mesh_normal_prediction.rs:120-127
inline void oct_quantize_uv(float u, float v, int32_t& qx, int32_t& qy) {
  float a = (u + 1.0f) * 127.0f;
  float b = (v + 1.0f) * 127.0f;
  qx = f32_as_i32(a);
  qy = f32_as_i32(b);
  into_faithful_oct_quantization(qx, qy);
}

// ---------------------------------------------------------------------------
// Portabilization — encode/attribute/portabilization/*

struct PortResult {
  PortAttribute att;
  Bytes info;  // metadata written to the side buffer (attribute_encoder.rs:293-298)
  std::vector<float> min_values;
  float range = 0;
};

// QuantizationCoordinateWise — quantization_coordinate_wise.rs:24-117
inline PortResult quantize_coordinate_wise(const Attribute& att, unsigned bits) {
  PortResult r;
  const size_t N = att.num_components, U = att.num_unique();
  std::vector<float> mn(N, 0.0f), mx(N, 0.0f);  // NdVector::zero(): min and max start at 0 (Appendix B.3)
  for (size_t i = 0; i < U; ++i)
    for (size_t k = 0; k < N; ++k) { float c = (float)att.comp_as_f64(i, k); if (c < mn[k]) mn[k] = c; }
  for (size_t i = 0; i < U; ++i)
    for (size_t k = 0; k < N; ++k) { float c = (float)att.comp_as_f64(i, k); if (c > mx[k]) mx[k] = c; }
  float delta_max = 0.0f;
  for (size_t k = 0; k < N; ++k) { float d = mx[k] - mn[k]; if (d > delta_max) delta_max = d; }
  for (size_t k = 0; k < N; ++k) w_f32(r.info, mn[k]);
  w_f32(r.info, delta_max);
  w_u8(r.info, (uint8_t)bits);
  r.min_values = mn;
  r.range = delta_max;
  const float maxq = (float)(uint64_t)(((uint64_t)1 << bits) - 1);  // f32::from_u64((1<<bits)-1)
  r.att.id = att.id; r.att.att_type = att.att_type; r.att.n = (uint32_t)N;
  r.att.vals.resize(U * N);
  for (size_t i = 0; i < U; ++i) {
    for (size_t k = 0; k < N; ++k) {  // portabilize_value — :70-91
      float val = (float)att.comp_as_f64(i, k);
      float diff = val - mn[k];
      float normalized = (delta_max == 0.0f) ? diff : diff / delta_max;
      float quantized = normalized * maxq;
      r.att.vals[i * N + k] = wrap_i32(f32_as_i64(quantized + 0.5f));
    }
  }
  r.att.has_map = att.has_map;
  r.att.map = att.map;
  return r;
}

// OctahedralQuantization — octahedral_quantization.rs:31-90
inline PortResult quantize_octahedral(const Attribute& att) {
  if (att.att_type != AT_NORMAL) throw Panic(ST_INTERNAL, "Octahedral quantization can only be applied to normal attributes.");
  if (att.num_components != 3) throw Panic(ST_INTERNAL, "assert!(N==3) in octahedral_transform");
  PortResult r;
  w_u8(r.info, 8);
  const size_t U = att.num_unique();
  r.att.id = att.id; r.att.att_type = att.att_type; r.att.n = 2;
  r.att.vals.resize(U * 2);
  const bool is_float = att.comp_type == CT_F32 || att.comp_type == CT_F64;
  for (size_t i = 0; i < U; ++i) {
    float u, v;
    if (att.comp_type == CT_F32) {
      octahedral_transform_f32((float)att.comp_as_f64(i, 0), (float)att.comp_as_f64(i, 1), (float)att.comp_as_f64(i, 2), u, v);
    } else if (is_float) {  // f64: arithmetic in f64, result cast to f32 (geom.rs:84-88)
      double x = att.comp_as_f64(i, 0), y = att.comp_as_f64(i, 1), z = att.comp_as_f64(i, 2);
      if (x == 0 && y == 0 && z == 0) throw Panic(ST_ZERO_NORMAL, "Zero vector cannot be transformed to octahedron space");
      double s = (std::fabs(x) + std::fabs(y)) + std::fabs(z);
      double uu = y / s, vv = z / s;
      if (x < 0) { double uo = uu < 0 ? std::fabs(vv) - 1.0 : 1.0 - std::fabs(vv); double vo = vv < 0 ? std::fabs(uu) - 1.0 : 1.0 - std::fabs(uu); uu = uo; vv = vo; }
      u = (float)uu; v = (float)vv;
    } else {
      double x = att.comp_as_f64(i, 0), y = att.comp_as_f64(i, 1), z = att.comp_as_f64(i, 2);
      if (x == 0 && y == 0 && z == 0) throw Panic(ST_ZERO_NORMAL, "Zero vector cannot be transformed to octahedron space");
      octahedral_transform_f32((float)x, (float)y, (float)z, u, v);
    }
    int32_t qx, qy;
    oct_quantize_uv(u, v, qx, qy);
    r.att.vals[i * 2] = qx;
    r.att.vals[i * 2 + 1] = qy;
  }
  r.att.has_map = att.has_map;
  r.att.map = att.map;
  return r;
}

// ToBits — to_bits.rs:29-49; buffer later read back as NdVector<N,i32>
// (attribute_encoder.rs:301-305, core/buffer/attribute.rs:49-63)
inline PortResult to_bits(const Attribute& att) {
  if (comp_size(att.comp_type) != 4) throw Panic(ST_UNSUPPORTED_DATA_TYPE, "ToBits buffer is reinterpreted as i32: component size must be 4");
  PortResult r;
  r.att.id = att.id; r.att.att_type = att.att_type; r.att.n = att.num_components;
  r.att.vals.resize(att.num_unique() * att.num_components);
  memcpy(r.att.vals.data(), att.buffer.data(), r.att.vals.size() * 4);
  r.att.has_map = att.has_map;
  r.att.map = att.map;
  return r;
}

// ---------------------------------------------------------------------------
// `vertices_up_till_now` helper. rank[v] = position in the sequence. In literal
// mode the reference's Vec::contains linear scan is used instead (Appendix C.1).
struct SeqRecord {
  std::vector<uint32_t> record;        // sequence_record (attribute_encoder.rs:330-335)
  std::vector<uint32_t> rank;          // rank[v] = index in record, NONE if absent
  bool literal = false;
  void init(size_t num_vertices, bool lit) { rank.assign(num_vertices, NONE); literal = lit; }
  bool contains(uint32_t v) const {
    if (literal) { for (uint32_t x : record) if (x == v) return true; return false; }
    return rank[v] != NONE;  // only vertices pushed so far have a rank
  }
  void push(uint32_t v) { if (rank[v] == NONE) rank[v] = (uint32_t)record.size(); record.push_back(v); }
  bool empty() const { return record.empty(); }
  uint32_t last() const { return record.back(); }
};

enum SchemeType { SCH_DELTA = 0, SCH_PARALLELOGRAM = 1, SCH_TEXCOORD = 5, SCH_NORMAL = 6 };   // prediction_scheme/mod.rs:74-86
enum TransformType { TR_DIFFERENCE = 0, TR_WRAPPED = 1, TR_OCT_ORTHOGONAL = 3 };              // prediction_transform/mod.rs:92-102

struct Predictor {
  const GenericCornerTable& ct;
  const PortAttribute* pos = nullptr;  // parent (quantized positions)
  unsigned N;
  std::vector<uint8_t> flips;         // MeshNormalPrediction::flips
  std::vector<uint8_t> orientation;   // MeshPredictionForTextureCoordinates::orientation
  Predictor(const GenericCornerTable& c, unsigned n) : ct(c), N(n) {}

  void last_value_or_zero(const SeqRecord& rec, const PortAttribute& att, int32_t* out) const {
    if (!rec.empty()) {
      const int32_t* p = att.get(ct.point_idx(ct.left_most_corner(rec.last())));
      for (unsigned k = 0; k < N; ++k) out[k] = p[k];
    } else for (unsigned k = 0; k < N; ++k) out[k] = 0;
  }

  // DeltaPrediction::predict — delta_prediction.rs:56-71
  void predict_delta(uint32_t, const SeqRecord& rec, const PortAttribute& att, int32_t* out) const { last_value_or_zero(rec, att, out); }

  // MeshParallelogramPrediction::predict — mesh_parallelogram_prediction.rs:186-237
  void predict_parallelogram(uint32_t c, const SeqRecord& rec, const PortAttribute& att, int32_t* out) const {
    uint32_t opp = ct.opposite(c);
    if (opp != NONE) {
      uint32_t opp_v = ct.vertex_idx(opp), next_v = ct.vertex_idx(c_next(c)), prev_v = ct.vertex_idx(c_prev(c));
      if (rec.contains(opp_v) && rec.contains(next_v) && rec.contains(prev_v)) {
        const int32_t* a = att.get(ct.point_idx(c_next(c)));
        const int32_t* b = att.get(ct.point_idx(c_prev(c)));
        const int32_t* d = att.get(ct.point_idx(opp));
        for (unsigned k = 0; k < N; ++k) out[k] = sub32(add32(a[k], b[k]), d[k]);
        return;
      }
    }
    last_value_or_zero(rec, att, out);
  }

  // MeshNormalPrediction — mesh_normal_prediction.rs:22-144
  void face_normal(uint32_t c, const int32_t* pos_c, int64_t* acc) const {  // compute_normal_of_face :22-44
    const int32_t* pn = pos->get(ct.point_idx(c_next(c)));
    const int32_t* pp = pos->get(ct.point_idx(c_prev(c)));
    int32_t dn[3], dp[3];
    for (int k = 0; k < 3; ++k) { dn[k] = sub32(pn[k], pos_c[k]); dp[k] = sub32(pp[k], pos_c[k]); }
    // cross product in i32 (wrapping), then widened — core/shared.rs:616-634
    int32_t cx = sub32(mul32(dn[1], dp[2]), mul32(dn[2], dp[1]));
    int32_t cy = sub32(mul32(dn[2], dp[0]), mul32(dn[0], dp[2]));
    int32_t cz = sub32(mul32(dn[0], dp[1]), mul32(dn[1], dp[0]));
    acc[0] = add64(acc[0], cx); acc[1] = add64(acc[1], cy); acc[2] = add64(acc[2], cz);
  }
  void predict_normal(uint32_t c, const SeqRecord&, const PortAttribute& att, int32_t* out) {
    if (!pos || pos->n != 3) throw Panic(ST_INTERNAL, "MeshNormalPrediction requires a 3-component position parent");
    const int32_t* pc = pos->get(ct.point_idx(c));
    int32_t pos_c[3] = {pc[0], pc[1], pc[2]};
    uint32_t curr = c;
    for (uint32_t l; (l = ct.swing_left(curr)) != NONE;) { curr = l; if (curr == c) break; }
    uint32_t start = curr;
    int64_t sum[3] = {0, 0, 0};
    face_normal(curr, pos_c, sum);
    for (uint32_t r; (r = ct.swing_right(curr)) != NONE;) { curr = r; if (curr == start) break; face_normal(curr, pos_c, sum); }
    const int64_t upper_bound = (int64_t)1 << 29;
    int64_t abs_sum = add64(add64(abs64(sum[0]), abs64(sum[1])), abs64(sum[2]));
    if (abs_sum > upper_bound) {
      int64_t quotient = abs_sum / upper_bound;
      for (int k = 0; k < 3; ++k) sum[k] = div64(sum[k], quotient);
    }
    int32_t n3[3] = {wrap_i32(sum[0]), wrap_i32(sum[1]), wrap_i32(sum[2])};
    int32_t o[2] = {0, 0};
    if (!(n3[0] == 0 && n3[1] == 0 && n3[2] == 0)) {
      float u, v;
      octahedral_transform_i32(n3[0], n3[1], n3[2], u, v);
      oct_quantize_uv(u, v, o[0], o[1]);
    }
    const int32_t* actual = att.get(ct.point_idx(c));
    // diff1 = out - actual; diff2 = out * -1 - actual; dot in i32 (wrapping)
    int32_t d1[2], d2[2];
    for (int k = 0; k < 2; ++k) { d1[k] = sub32(o[k], actual[k]); d2[k] = sub32(mul32(o[k], -1), actual[k]); }
    int32_t dot1 = add32(mul32(d1[0], d1[0]), mul32(d1[1], d1[1]));
    int32_t dot2 = add32(mul32(d2[0], d2[0]), mul32(d2[1], d2[1]));
    if (dot1 > dot2) { flips.push_back(1); o[0] = mul32(o[0], -1); o[1] = mul32(o[1], -1); }
    else flips.push_back(0);
    out[0] = o[0]; out[1] = o[1];
    for (unsigned k = 2; k < N; ++k) out[k] = 0;
  }

  // MeshPredictionForTextureCoordinates — mesh_prediction_for_texture_coordinates.rs:21-219
  static uint64_t int_sqrt(uint64_t value) {  // :33-49
    if (value == 0) return 0;
    uint64_t act = value, sqrt = 1;
    while (act >= 2) { sqrt *= 2; act /= 4; }
    sqrt = (sqrt + value / sqrt) / 2;
    while (sqrt * sqrt > value) sqrt = (sqrt + value / sqrt) / 2;
    return sqrt;
  }
  void position_for(uint32_t p, int64_t* o) const {  // get_position_for_vertex :21-30
    if (p < pos->len()) { const int32_t* q = pos->get(p); o[0] = q[0]; o[1] = q[1]; o[2] = q[2]; }
    else { o[0] = o[1] = o[2] = 0; }
  }
  void fallback_predict(uint32_t c, const SeqRecord& rec, const PortAttribute& att, int32_t* out) const {  // :52-82
    uint32_t nc = c_next(c);
    if (rec.contains(ct.vertex_idx(nc))) { const int32_t* p = att.get(ct.point_idx(nc)); out[0] = p[0]; out[1] = p[1]; return; }
    last_value_or_zero(rec, att, out);
  }
  void predict_texcoord(uint32_t i, const SeqRecord& rec, const PortAttribute& att, int32_t* out) {
    if (N != 2) throw Panic(ST_INTERNAL, "Texture coordinate prediction is only for 2D vectors");
    if (!pos || pos->n != 3) throw Panic(ST_INTERNAL, "texcoord prediction requires a 3-component position parent");
    uint32_t nc = c_next(i), pc = c_prev(i);
    uint32_t next_pt = ct.point_idx(nc), prev_pt = ct.point_idx(pc), curr_pt = ct.point_idx(i);
    uint32_t next_vertex = ct.vertex_idx(nc), prev_vertex = ct.vertex_idx(pc);
    if (rec.contains(next_vertex) && rec.contains(prev_vertex)) {
      const int32_t* cu = att.get(curr_pt); int64_t curr_uv[2] = {cu[0], cu[1]};
      const int32_t* nu = att.get(next_pt); int64_t next_uv[2] = {nu[0], nu[1]};
      const int32_t* pu = att.get(prev_pt); int64_t prev_uv[2] = {pu[0], pu[1]};
      if (next_uv[0] == prev_uv[0] && next_uv[1] == prev_uv[1]) { out[0] = pu[0]; out[1] = pu[1]; return; }
      int64_t curr_pos[3], next_pos[3], prev_pos[3];
      position_for(curr_pt, curr_pos); position_for(next_pt, next_pos); position_for(prev_pt, prev_pos);
      int64_t pn[3] = {sub64(prev_pos[0], next_pos[0]), sub64(prev_pos[1], next_pos[1]), sub64(prev_pos[2], next_pos[2])};
      uint64_t pn_norm2 = (uint64_t)add64(add64(mul64(pn[0], pn[0]), mul64(pn[1], pn[1])), mul64(pn[2], pn[2]));
      if (pn_norm2 != 0) {
        int64_t cn[3] = {sub64(curr_pos[0], next_pos[0]), sub64(curr_pos[1], next_pos[1]), sub64(curr_pos[2], next_pos[2])};
        int64_t cn_dot_pn = add64(add64(mul64(pn[0], cn[0]), mul64(pn[1], cn[1])), mul64(pn[2], cn[2]));
        int64_t pn_uv[2] = {sub64(prev_uv[0], next_uv[0]), sub64(prev_uv[1], next_uv[1])};
        const int64_t I64MAX = std::numeric_limits<int64_t>::max();
        int64_t n_uv_absmax = std::max(abs64(next_uv[0]), abs64(next_uv[1]));
        if (n_uv_absmax > div64(I64MAX, (int64_t)pn_norm2)) { fallback_predict(i, rec, att, out); return; }
        int64_t pn_uv_absmax = std::max(abs64(pn_uv[0]), abs64(pn_uv[1]));
        if (abs64(cn_dot_pn) > div64(I64MAX, pn_uv_absmax)) { fallback_predict(i, rec, att, out); return; }
        int64_t x_uv[2] = {add64(mul64(next_uv[0], (int64_t)pn_norm2), mul64(pn_uv[0], cn_dot_pn)),
                           add64(mul64(next_uv[1], (int64_t)pn_norm2), mul64(pn_uv[1], cn_dot_pn))};
        int64_t pn_absmax = std::max(std::max(abs64(pn[0]), abs64(pn[1])), abs64(pn[2]));
        if (abs64(cn_dot_pn) > div64(I64MAX, pn_absmax)) { fallback_predict(i, rec, att, out); return; }
        // x_pos = next_pos + pn * cn_dot_pn / pn_norm2 (left-to-right: (pn*cn_dot_pn)/pn_norm2)
        int64_t x_pos[3], cx[3];
        for (int k = 0; k < 3; ++k) { x_pos[k] = add64(next_pos[k], div64(mul64(pn[k], cn_dot_pn), (int64_t)pn_norm2)); cx[k] = sub64(curr_pos[k], x_pos[k]); }
        uint64_t cx_norm2 = (uint64_t)add64(add64(mul64(cx[0], cx[0]), mul64(cx[1], cx[1])), mul64(cx[2], cx[2]));
        int64_t cx_uv[2] = {pn_uv[1], (int64_t)(0 - (uint64_t)pn_uv[0])};
        uint64_t norm_squared = int_sqrt(cx_norm2 * pn_norm2);
        cx_uv[0] = mul64(cx_uv[0], (int64_t)norm_squared);
        cx_uv[1] = mul64(cx_uv[1], (int64_t)norm_squared);
        int64_t p0[2] = {div64(add64(x_uv[0], cx_uv[0]), (int64_t)pn_norm2), div64(add64(x_uv[1], cx_uv[1]), (int64_t)pn_norm2)};
        int64_t p1[2] = {div64(sub64(x_uv[0], cx_uv[0]), (int64_t)pn_norm2), div64(sub64(x_uv[1], cx_uv[1]), (int64_t)pn_norm2)};
        int64_t e0[2] = {sub64(curr_uv[0], p0[0]), sub64(curr_uv[1], p0[1])};
        int64_t e1[2] = {sub64(curr_uv[0], p1[0]), sub64(curr_uv[1], p1[1])};
        int64_t d0 = add64(mul64(e0[0], e0[0]), mul64(e0[1], e0[1]));
        int64_t d1 = add64(mul64(e1[0], e1[0]), mul64(e1[1], e1[1]));
        if (d0 < d1) { orientation.push_back(1); out[0] = wrap_i32(p0[0]); out[1] = wrap_i32(p0[1]); }
        else { orientation.push_back(0); out[0] = wrap_i32(p1[0]); out[1] = wrap_i32(p1[1]); }
        return;
      }
    }
    fallback_predict(i, rec, att, out);
  }

  // encode_prediction_metadtata for normals — mesh_normal_prediction.rs:147-163
  void encode_normal_metadata(Bytes& w) const {
    size_t n0 = 0;
    for (uint8_t f : flips) if (!f) n0++;
    uint8_t zero_prob = zero_prob_f32(n0, (float)flips.size());
    RabsCoder rc(zero_prob);
    w_u8(w, zero_prob);
    for (uint8_t f : flips) rc.write(f ? 1 : 0);
    Bytes b = rc.flush();
    leb128_write(b.size(), w);
    w.insert(w.end(), b.begin(), b.end());
  }
  // encode_prediction_metadtata for texcoords — mesh_prediction_for_texture_coordinates.rs:221-260
  void encode_texcoord_metadata(Bytes& w) const {
    size_t n0 = 0;
    { bool last = true; for (uint8_t o : orientation) { bool ob = o != 0; if (ob != last) { last = ob; n0++; } } }
    float len_f = (float)orientation.size() + 0.001f;
    uint8_t zero_prob = zero_prob_f32(n0, len_f);
    RabsCoder rc(zero_prob);
    w_u32(w, (uint32_t)orientation.size());
    w_u8(w, zero_prob);
    std::vector<uint8_t> bits(orientation.size());
    { bool last = true;
      for (size_t k = orientation.size(); k-- > 0;) { bool ob = orientation[k] != 0; if (ob == last) bits[k] = 1; else { last = ob; bits[k] = 0; } } }
    for (uint8_t b : bits) rc.write(b);
    Bytes b = rc.flush();
    leb128_write(b.size(), w);
    w.insert(w.end(), b.begin(), b.end());
  }
};

struct AttributeTrace {
  std::vector<uint32_t> sequence;     // corners
  std::vector<int32_t> quantized;     // port_att values (unique)
  std::vector<int32_t> predictions;   // per sequence element, N comps (pre-clamp)
  std::vector<uint32_t> symbols;      // interleaved
  std::vector<uint8_t> side_bits;     // flips / orientation
  int32_t wrap_min = 0, wrap_max = 0;
  SymbolStreamTrace stream;
};

// AttributeEncoder::encode_portabilized — attribute_encoder.rs:312-389
inline PortAttribute encode_one_attribute(const Attribute& att, size_t att_idx, const std::vector<PortAttribute>& port_atts,
                                          const GenericCornerTable& ct, const std::vector<uint32_t>& corners_of_edgebreaker,
                                          const OracleConfig& cfg, Bytes& w, AttributeTrace* tr) {
  (void)att_idx;
  // GroupConfig::default_for — attribute_encoder.rs:59-108
  int scheme, transform;
  switch (att.att_type) {
    case AT_POSITION: scheme = SCH_PARALLELOGRAM; transform = TR_WRAPPED; break;
    case AT_NORMAL: scheme = SCH_NORMAL; transform = TR_OCT_ORTHOGONAL; break;
    case AT_TEXCOORD: scheme = SCH_TEXCOORD; transform = TR_WRAPPED; break;
    case AT_CUSTOM: scheme = SCH_PARALLELOGRAM; transform = TR_WRAPPED; break;
    default: scheme = SCH_DELTA; transform = TR_DIFFERENCE; break;
  }
  w_u8(w, (uint8_t)scheme);     // :159
  w_u8(w, (uint8_t)transform);  // :160
  if (comp_size(att.comp_type) == 0) throw EncodeError(ST_UNSUPPORTED_DATA_TYPE, "Unsupported data type.");
  if (att.num_components == 0) throw Panic(ST_INTERNAL, "Vector of dimension 0 is not allowed");
  if (att.num_components > 4) throw EncodeError(ST_UNSUPPORTED_NUM_COMPONENTS, "Attribute data has too many components");

  StageTimer tm;
  std::vector<uint32_t> sequence = compute_sequence(ct, corners_of_edgebreaker, cfg.literal);  // :240-251
  tm.lap("  sequencer");

  // portabilization::Config::default_for — portabilization/mod.rs:116-142
  PortResult pr;
  if (att.att_type == AT_NORMAL) pr = quantize_octahedral(att);
  else if (att.att_type == AT_CUSTOM) pr = to_bits(att);
  else pr = quantize_coordinate_wise(att, att.att_type == AT_TEXCOORD ? cfg.texcoord_bits
                                        : att.att_type == AT_POSITION ? cfg.position_bits : cfg.generic_bits);
  tm.lap("  portabilize");
  const PortAttribute& port = pr.att;
  const unsigned N = port.n;
  if (N < 1 || N > 4) throw EncodeError(ST_UNSUPPORTED_NUM_COMPONENTS, "Attribute data has too many components");

  // parents: already-portabilized attributes by id (attribute/mod.rs:63-66)
  Predictor pred(ct, N);
  std::vector<const PortAttribute*> parents;
  for (uint32_t pid : att.parents) {
    const PortAttribute* found = nullptr;
    for (auto& pa : port_atts) if (pa.id == pid) { found = &pa; break; }
    if (!found) throw Panic(ST_INVALID_ARGUMENT, "unwrap on None: parent attribute not yet encoded");
    parents.push_back(found);
  }
  if (scheme == SCH_NORMAL) {
    if (parents.size() != 1) throw Panic(ST_INVALID_ARGUMENT, "MeshNormalPrediction requires exactly one parent attribute for position.");
    if (parents[0]->att_type != AT_POSITION) throw Panic(ST_INVALID_ARGUMENT, "MeshNormalPrediction requires the first parent attribute to be of type Position.");
    pred.pos = parents[0];
  } else if (scheme == SCH_TEXCOORD) {
    if (parents.empty()) throw Panic(ST_INVALID_ARGUMENT, "index out of bounds: parents[0]");
    pred.pos = parents[0];
  }

  SeqRecord rec;
  rec.init(ct.num_vertices(), cfg.literal);
  std::vector<int32_t> origs, preds;
  origs.reserve(sequence.size() * N);
  preds.reserve(sequence.size() * N);
  for (uint32_t c : sequence) {  // :332-338
    int32_t p[4] = {0, 0, 0, 0};
    switch (scheme) {
      case SCH_PARALLELOGRAM: pred.predict_parallelogram(c, rec, port, p); break;
      case SCH_NORMAL: pred.predict_normal(c, rec, port, p); break;
      case SCH_TEXCOORD: pred.predict_texcoord(c, rec, port, p); break;
      default: pred.predict_delta(c, rec, port, p); break;
    }
    rec.push(ct.vertex_idx(c));
    const int32_t* o = port.get(ct.point_idx(c));
    for (unsigned k = 0; k < N; ++k) { origs.push_back(o[k]); preds.push_back(p[k]); }
  }

  tm.lap("  predict");
  // transforms
  Bytes transform_info;
  std::vector<uint32_t> symbols(origs.size());
  const size_t M = sequence.size();
  int32_t wmin = std::numeric_limits<int32_t>::max(), wmax = std::numeric_limits<int32_t>::min();
  if (transform == TR_WRAPPED) {  // WrappedDifference — wrapped_difference.rs:36-98
    for (int32_t v : origs) { if (v > wmax) wmax = v; if (v < wmin) wmin = v; }
    int32_t diff = sub32(wmax, wmin);
    int32_t max_diff = add32(1, diff);
    int32_t max_corr = max_diff / 2;
    int32_t min_corr = -max_corr;
    if ((max_diff & 1) == 0) max_corr -= 1;
    for (size_t i = 0; i < origs.size(); ++i) {
      int32_t pv = preds[i];
      if (M > 0 && wmin > wmax) throw Panic(ST_INTERNAL, "clamp: min > max");
      pv = pv < wmin ? wmin : (pv > wmax ? wmax : pv);
      int32_t val = sub32(origs[i], pv);
      int32_t corr = val > max_corr ? sub32(val, max_diff) : (val < min_corr ? add32(val, max_diff) : val);
      symbols[i] = (uint32_t)to_positive_i32(corr);
    }
    w_i32(transform_info, wmin);
    w_i32(transform_info, wmax);
  } else if (transform == TR_OCT_ORTHOGONAL) {  // oct_orthogonal.rs:23-85
    if (N != 2) throw Panic(ST_INTERNAL, "assert!(N==2) in OctahedronOrthogonalTransform");
    for (size_t i = 0; i < M; ++i) {
      int32_t o0 = origs[2 * i], o1 = origs[2 * i + 1], p0 = preds[2 * i], p1 = preds[2 * i + 1];
      const int32_t one = 255 / 2;
      p0 = sub32(p0, one); p1 = sub32(p1, one); o0 = sub32(o0, one); o1 = sub32(o1, one);
      if (add32(abs32(p0), abs32(p1)) > one) {
        int32_t pp0 = p0;
        int32_t qs = -signum32(mul32(p0, p1));
        p0 = add32(mul32(qs, p1), mul32(signum32(p0), one));
        p1 = add32(mul32(qs, pp0), mul32(signum32(p1), one));
        int32_t oo0 = o0;
        int32_t qs2 = -signum32(mul32(o0, o1));
        o0 = add32(mul32(qs2, o1), mul32(signum32(o0), one));
        o1 = add32(mul32(qs2, oo0), mul32(signum32(o1), one));
      }
      if (!(p0 == 0 && p1 == 0)) {
        while (p0 >= 0 || p1 > 0) {
          int32_t t = p0; p0 = -p1; p1 = t;
          t = o0; o0 = -o1; o1 = t;
        }
      }
      int32_t c0 = sub32(o0, p0), c1 = sub32(o1, p1);
      if (c0 < 0) c0 = add32(c0, 255);
      if (c1 < 0) c1 = add32(c1, 255);
      symbols[2 * i] = (uint32_t)c0;
      symbols[2 * i + 1] = (uint32_t)c1;
    }
    w_u32(transform_info, 255);
    w_u32(transform_info, 255 / 2);
  } else {  // Difference — difference.rs:26-49
    for (size_t i = 0; i < origs.size(); ++i) symbols[i] = (uint32_t)to_positive_i32(sub32(origs[i], preds[i]));
  }

  tm.lap("  transform");
  w_u8(w, 1);  // rans_encoding (:344)
  SymbolStreamTrace* st = tr ? &tr->stream : nullptr;
  // symbols widened with `as u64` from i32 (sign-extending) in the reference (:347-350);
  // a negative residual would index the histogram out of range there -> treat as invalid symbol.
  for (uint32_t s : symbols) if (s & 0x80000000u) throw EncodeError(ST_RANS_INVALID_SYMBOL, "negative symbol");
  encode_symbols_direct(symbols, w, st);
  tm.lap("  entropy (hist+table+rANS)");

  // metadata order — :362-382
  if (scheme == SCH_NORMAL) {
    w.insert(w.end(), transform_info.begin(), transform_info.end());
    pred.encode_normal_metadata(w);
  } else if (scheme == SCH_TEXCOORD) {
    pred.encode_texcoord_metadata(w);
    w.insert(w.end(), transform_info.begin(), transform_info.end());
  } else {
    w.insert(w.end(), transform_info.begin(), transform_info.end());
  }
  w.insert(w.end(), pr.info.begin(), pr.info.end());  // :384-386
  tm.lap("  metadata (rABS)");

  if (tr) {
    tr->sequence = sequence;
    tr->quantized = port.vals;
    tr->predictions = preds;
    tr->symbols = symbols;
    tr->side_bits = scheme == SCH_NORMAL ? pred.flips : pred.orientation;
    tr->wrap_min = wmin; tr->wrap_max = wmax;
  }
  return pr.att;
}

struct EncodeTrace {
  std::vector<uint32_t> opposite, corner_to_vertex, left_most;
  size_t num_vertices = 0;
  std::vector<uint8_t> eb_symbols;
  std::vector<uint32_t> corners_of_edgebreaker;
  size_t connectivity_end = 0;  // byte offset where the attribute section starts
  std::vector<AttributeTrace> atts;
  std::vector<std::vector<uint32_t>> att_corner_to_vertex, att_left_most;
  std::vector<std::vector<uint8_t>> att_seam;
};

// encode::encode — encode/mod.rs:59-97 (+ header/mod.rs:26-54, attribute/mod.rs:13-93)
inline Bytes encode_mesh(const Mesh& mesh, const OracleConfig& cfg, EncodeTrace* tr = nullptr) {
  Bytes w;
  // header
  for (char ch : std::string("DRACO")) w_u8(w, (uint8_t)ch);
  w_u8(w, 2); w_u8(w, 2);
  w_u8(w, 1);  // TrianglarMesh
  w_u8(w, 1);  // Edgebreaker
  w_u16(w, 0); // flags (metadata off)

  // connectivity — connectivity/mod.rs:17-79, edgebreaker.rs:128-193
  const Attribute* pos_att = nullptr;
  for (auto& a : mesh.atts) if (a.att_type == AT_POSITION) { pos_att = &a; break; }
  if (!pos_att) throw Panic(ST_INVALID_ARGUMENT, "unwrap on None: no position attribute");
  StageTimer tm;
  CornerTable ct(mesh.faces, *pos_att);
  tm.lap("corner table");
  std::vector<AttributeCornerTable> att_tables;
  for (auto& a : mesh.atts) { if (a.att_type == AT_POSITION) continue; att_tables.emplace_back(ct, a); }
  tm.lap("attribute corner tables");
  Edgebreaker eb(ct, att_tables);
  EdgebreakerOutput eo = eb.encode_connectivity(mesh.faces.size(), w);
  tm.lap("edgebreaker");
  if (tr) {
    tr->opposite = ct.opposite_corners;
    tr->left_most = ct.left_most_corners;
    tr->num_vertices = ct.num_vertices();
    tr->corner_to_vertex.resize(ct.num_corners());
    for (uint32_t c = 0; c < ct.num_corners(); ++c) tr->corner_to_vertex[c] = ct.vertex_idx(c);
    tr->eb_symbols = eo.symbols;
    tr->corners_of_edgebreaker = eo.corners_of_edgebreaker;
    tr->connectivity_end = w.size();
    for (auto& t : att_tables) { tr->att_corner_to_vertex.push_back(t.corner_to_vertex); tr->att_left_most.push_back(t.left_most_corners); tr->att_seam.push_back(t.is_edge_on_seam); }
    tr->atts.resize(mesh.atts.size());
  }

  // attributes — attribute/mod.rs:13-93
  const auto& atts = mesh.atts;
  w_u8(w, (uint8_t)atts.size());
  for (size_t i = 0; i < atts.size(); ++i) {
    w_u8(w, (uint8_t)((uint8_t)i - 1));  // (i as u8).wrapping_sub(1)
    w_u8(w, (uint8_t)atts[i].domain);
    w_u8(w, 0);  // TraversalType::DepthFirst
  }
  for (auto& a : atts) {
    w_u8(w, 1);
    w_u8(w, (uint8_t)a.att_type);
    w_u8(w, (uint8_t)a.comp_type);
    w_u8(w, (uint8_t)a.num_components);
    w_u8(w, 0);
    w_u8(w, (uint8_t)a.id);
    // PortabilizationType::default_for(..).get_id() — portabilization/mod.rs:85-110
    w_u8(w, a.att_type == AT_NORMAL ? 3 : (a.att_type == AT_CUSTOM ? 1 : 2));
  }
  std::vector<PortAttribute> port_atts;
  for (size_t i = 0; i < atts.size(); ++i) {
    // encode_typed: AllInclusiveCornerTable::attribute_corner_table(i) — all_inclusive_corner_table.rs:33-47
    PortAttribute pa;
    if (i > 0 && i - 1 < att_tables.size()) {
      RefAttributeCornerTable rct(ct, att_tables[i - 1]);
      pa = encode_one_attribute(atts[i], i, port_atts, rct, eo.corners_of_edgebreaker, cfg, w, tr ? &tr->atts[i] : nullptr);
    } else {
      pa = encode_one_attribute(atts[i], i, port_atts, ct, eo.corners_of_edgebreaker, cfg, w, tr ? &tr->atts[i] : nullptr);
    }
    port_atts.push_back(std::move(pa));
    tm.lap("attribute total");
  }
  return w;
}

}  // namespace orc
