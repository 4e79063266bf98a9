// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_core.hpp header).
// C entry points used by tests/ (ctypes), __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs. Never linked into the product library.
#include <chrono>
#include <cstdlib>
#include <map>
#include <memory>
#include <thread>
#include "../include/dxo.h"
#include "orc_attribute.hpp"
#include "orc_decode.hpp"
#include "orc_inverse.hpp"
#include "orc_mesh.hpp"

using namespace orc;

namespace {

struct Trace { std::map<std::string, std::vector<uint8_t>> items; };

template <class T> void put(Trace& t, const std::string& k, const std::vector<T>& v) {
  auto& dst = t.items[k];
  dst.resize(v.size() * sizeof(T));
  if (!v.empty()) memcpy(dst.data(), v.data(), dst.size());
}
template <class T> void put_scalar(Trace& t, const std::string& k, T v) { put(t, k, std::vector<T>{v}); }

Mesh mesh_from_c(const dxo_mesh* m) {
  if (!m || (!m->faces && m->num_faces) || (!m->attributes && m->num_attributes)) throw EncodeError(ST_INVALID_ARGUMENT, "null mesh");
  Mesh out;
  out.faces.resize(m->num_faces);
  for (uint64_t f = 0; f < m->num_faces; ++f) out.faces[f] = {m->faces[3 * f], m->faces[3 * f + 1], m->faces[3 * f + 2]};
  for (uint32_t i = 0; i < m->num_attributes; ++i) {
    const dxo_attribute& a = m->attributes[i];
    Attribute o;
    o.id = a.unique_id; o.att_type = a.att_type; o.comp_type = a.component_type; o.num_components = a.num_components; o.domain = a.domain;
    o.parents.assign(a.parent_ids, a.parent_ids + a.num_parents);
    size_t bytes = (size_t)a.num_unique_values * a.num_components * comp_size(a.component_type);
    o.buffer.assign((const uint8_t*)a.values, (const uint8_t*)a.values + bytes);
    if (a.point_to_value) { o.has_map = true; o.map.assign(a.point_to_value, a.point_to_value + a.num_points); }
    out.atts.push_back(std::move(o));
  }
  return out;
}

void fill_trace(Trace& t, const EncodeTrace& e) {
  put(t, "opposite", e.opposite);
  put(t, "corner_to_vertex", e.corner_to_vertex);
  put(t, "left_most", e.left_most);
  put_scalar<uint64_t>(t, "num_vertices", e.num_vertices);
  put(t, "eb_symbols", e.eb_symbols);
  put(t, "corners_of_edgebreaker", e.corners_of_edgebreaker);
  put_scalar<uint64_t>(t, "connectivity_end", e.connectivity_end);
  for (size_t j = 0; j < e.att_corner_to_vertex.size(); ++j) {
    std::string p = "att" + std::to_string(j + 1) + ".";
    put(t, p + "c2v", e.att_corner_to_vertex[j]);
    put(t, p + "left_most", e.att_left_most[j]);
    put(t, p + "seam", e.att_seam[j]);
  }
  for (size_t i = 0; i < e.atts.size(); ++i) {
    const AttributeTrace& a = e.atts[i];
    std::string p = "att" + std::to_string(i) + ".";
    put(t, p + "sequence", a.sequence);
    put(t, p + "quantized", a.quantized);
    put(t, p + "predictions", a.predictions);
    put(t, p + "symbols", a.symbols);
    put(t, p + "side_bits", a.side_bits);
    put(t, p + "wrap_minmax", std::vector<int32_t>{a.wrap_min, a.wrap_max});
    put(t, p + "histogram", a.stream.histogram);
    put(t, p + "distribution", a.stream.table.distribution);
    put(t, p + "table_bytes", a.stream.table.serialized);
    put(t, p + "payload", a.stream.payload);
    put(t, p + "bit_length", std::vector<uint32_t>{a.stream.bit_length, a.stream.precision});
  }
}

thread_local std::string g_last_error;

template <class F> int guarded(F&& f) {
  try { f(); return ST_OK; }
  catch (const Panic& p) { g_last_error = std::string("panic: ") + p.what(); return p.code == ST_ZERO_NORMAL || p.code == ST_UNUSED_VERTICES ? p.code : ST_UNSUPPORTED_INPUT; }
  catch (const EncodeError& e) { g_last_error = e.what(); return e.code; }
  catch (const std::bad_alloc&) { g_last_error = "out of memory"; return -22; }
  catch (const std::exception& e) { g_last_error = e.what(); return ST_INTERNAL; }
}

OracleConfig cfg_from_c(const dxo_config* c, int literal) {
  OracleConfig o;
  if (c) { o.position_bits = c->position_bits; o.texcoord_bits = c->texcoord_bits; o.generic_bits = c->generic_bits; }
  o.literal = literal != 0;
  return o;
}

struct MeshHandle {
  Mesh mesh;
  std::vector<uint32_t> faces_flat;
  std::vector<dxo_attribute> atts;
};

}  // namespace

#include <cmath>
extern "C" {

const char* orc_last_error() { return g_last_error.c_str(); }

// encode::encode restated. literal != 0 selects the reference's O(V^2) loops.
int orc_encode(const dxo_mesh* mesh, const dxo_config* cfg, int literal, dxo_bytes* out, void** trace) {
  if (!out) return ST_INVALID_ARGUMENT;
  out->data = nullptr; out->len = 0;
  if (trace) *trace = nullptr;
  return guarded([&] {
    Mesh m = mesh_from_c(mesh);
    EncodeTrace et;
    Bytes b = encode_mesh(m, cfg_from_c(cfg, literal), trace ? &et : nullptr);
    out->data = (uint8_t*)malloc(b.size() ? b.size() : 1);
    if (!out->data) throw std::bad_alloc();
    memcpy(out->data, b.data(), b.size());
    out->len = b.size();
    if (trace) { auto* t = new Trace(); fill_trace(*t, et); *trace = t; }
  });
}

// Repeats the oracle `reps` times on `threads` host threads (each thread encodes the
// same mesh; used for the multi-core CPU baseline). Returns seconds of wall clock.
double orc_encode_timed(const dxo_mesh* mesh, const dxo_config* cfg, int reps, int threads, int* status) {
  int st = ST_OK;
  double secs = 0;
  st = guarded([&] {
    Mesh m = mesh_from_c(mesh);
    OracleConfig oc = cfg_from_c(cfg, 0);
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    std::vector<int> sts((size_t)std::max(1, threads), ST_OK);
    for (int k = 0; k < std::max(1, threads); ++k)
      th.emplace_back([&, k] { sts[k] = guarded([&] { for (int r = 0; r < reps; ++r) { Bytes b = encode_mesh(m, oc, nullptr); if (b.empty()) throw EncodeError(ST_INTERNAL, "empty"); } }); });
    for (auto& t : th) t.join();
    secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (int s : sts) if (s != ST_OK) throw EncodeError(s, "worker failed");
  });
  if (status) *status = st;
  return secs;
}

// Corner tables only (CornerTable::new + AttributeCornerTable::new per non-position attribute).
int orc_corner_tables(const dxo_mesh* mesh, void** trace) {
  if (!trace) return ST_INVALID_ARGUMENT;
  *trace = nullptr;
  return guarded([&] {
    Mesh m = mesh_from_c(mesh);
    const Attribute* pos = nullptr;
    for (auto& a : m.atts) if (a.att_type == AT_POSITION) { pos = &a; break; }
    if (!pos) throw Panic(ST_INVALID_ARGUMENT, "no position attribute");
    CornerTable ct(m.faces, *pos);
    auto* t = new Trace();
    put(*t, "opposite", ct.opposite_corners);
    put(*t, "left_most", ct.left_most_corners);
    put_scalar<uint64_t>(*t, "num_vertices", ct.num_vertices());
    std::vector<uint32_t> c2v(ct.num_corners());
    for (uint32_t c = 0; c < ct.num_corners(); ++c) c2v[c] = ct.vertex_idx(c);
    put(*t, "corner_to_vertex", c2v);
    put_scalar<uint32_t>(*t, "non_manifold_edges", CornerTable::contains_non_manifold_edges(ct.conn_faces) ? 1 : 0);
    size_t j = 1;
    for (auto& a : m.atts) {
      if (a.att_type == AT_POSITION) continue;
      AttributeCornerTable act(ct, a);
      std::string p = "att" + std::to_string(j++) + ".";
      put(*t, p + "c2v", act.corner_to_vertex);
      put(*t, p + "left_most", act.left_most_corners);
      put(*t, p + "seam", act.is_edge_on_seam);
      put(*t, p + "vertex_on_seam", act.is_vertex_on_seam);
      put(*t, p + "vertex_to_value", act.vertex_to_attribute_map);
      put_scalar<uint64_t>(*t, p + "num_vertices", act.n_vertices);
      std::vector<uint32_t> sl(ct.num_corners()), sr(ct.num_corners());
      for (uint32_t c = 0; c < ct.num_corners(); ++c) { sl[c] = act.swing_left(c, ct); sr[c] = act.swing_right(c, ct); }
      put(*t, p + "swing_left", sl);
      put(*t, p + "swing_right", sr);
    }
    *trace = t;
  });
}

int orc_trace_get(void* trace, const char* key, const void** data, uint64_t* nbytes) {
  if (!trace || !key || !data || !nbytes) return ST_INVALID_ARGUMENT;
  auto* t = (Trace*)trace;
  auto it = t->items.find(key);
  if (it == t->items.end()) return ST_INVALID_ARGUMENT;
  *data = it->second.data();
  *nbytes = it->second.size();
  return ST_OK;
}
void orc_trace_free(void* trace) { delete (Trace*)trace; }
void orc_free_bytes(dxo_bytes* b) { if (b && b->data) { free(b->data); b->data = nullptr; b->len = 0; } }

// ---- caller side: mesh construction -------------------------------------------------
static void finish_handle(MeshHandle* h) {
  h->faces_flat.clear();
  for (auto& f : h->mesh.faces) { h->faces_flat.push_back(f[0]); h->faces_flat.push_back(f[1]); h->faces_flat.push_back(f[2]); }
  h->atts.clear();
  for (auto& a : h->mesh.atts) {
    dxo_attribute d{};
    d.att_type = a.att_type; d.component_type = a.comp_type; d.num_components = a.num_components; d.domain = a.domain;
    d.unique_id = a.id; d.num_parents = (uint32_t)a.parents.size(); d.parent_ids = a.parents.data();
    d.num_unique_values = a.num_unique(); d.values = a.buffer.data();
    d.num_points = a.len(); d.point_to_value = a.has_map ? a.map.data() : nullptr;
    h->atts.push_back(d);
  }
}

void* orc_mesh_from_obj(const char* path, int* status) {
  MeshHandle* h = nullptr;
  int st = guarded([&] { h = new MeshHandle(); h->mesh = load_obj(path); finish_handle(h); });
  if (st != ST_OK) { delete h; h = nullptr; }
  if (status) *status = st;
  return h;
}

// Raw per-point arrays -> MeshBuilder::add_attribute (value dedup) -> build().
// atts[i].values holds num_unique_values per-point values; point_to_value is ignored.
void* orc_mesh_build(const uint32_t* faces, uint64_t num_faces, const dxo_attribute* atts, uint32_t num_atts, int* status) {
  MeshHandle* h = nullptr;
  int st = guarded([&] {
    MeshBuilder b;
    for (uint64_t f = 0; f < num_faces; ++f) b.faces.push_back({faces[3 * f], faces[3 * f + 1], faces[3 * f + 2]});
    for (uint32_t i = 0; i < num_atts; ++i) {
      const dxo_attribute& a = atts[i];
      b.add_attribute(a.values, a.num_unique_values, a.component_type, a.num_components, a.att_type, a.domain,
                      std::vector<uint32_t>(a.parent_ids, a.parent_ids + a.num_parents));
    }
    h = new MeshHandle();
    h->mesh = b.build();
    finish_handle(h);
  });
  if (st != ST_OK) { delete h; h = nullptr; }
  if (status) *status = st;
  return h;
}
int orc_mesh_view(void* handle, dxo_mesh* out) {
  if (!handle || !out) return ST_INVALID_ARGUMENT;
  auto* h = (MeshHandle*)handle;
  out->num_faces = h->mesh.faces.size();
  out->faces = h->faces_flat.data();
  out->num_attributes = (uint32_t)h->atts.size();
  out->attributes = h->atts.data();
  return ST_OK;
}
void orc_mesh_free(void* handle) { delete (MeshHandle*)handle; }

// Attribute::from (dedup) followed by Attribute::remove of `removed` points, f32 x ncomp.
// out_map receives the point->value map (identity when the attribute has none).
int orc_dedup_and_remove(const float* values, uint64_t n, uint32_t ncomp, const uint32_t* removed, uint64_t nrem,
                         uint32_t* out_map, uint64_t* out_len, uint64_t* out_num_unique, int* out_has_map) {
  return guarded([&] {
    Attribute a;
    a.comp_type = CT_F32; a.num_components = ncomp;
    a.buffer.assign((const uint8_t*)values, (const uint8_t*)values + n * ncomp * 4);
    remove_duplicate_values(a);
    for (uint64_t i = 0; i < nrem; ++i) remove_points(a, std::vector<uint32_t>{removed[i]});
    *out_len = a.len(); *out_num_unique = a.num_unique(); *out_has_map = a.has_map ? 1 : 0;
    for (size_t p = 0; p < a.len(); ++p) out_map[p] = a.unique_val_idx((uint32_t)p);
  });
}

// ---- unit-level entry points for the known-answer tests --------------------------------
int orc_leb128(uint64_t v, uint8_t* out, uint64_t* n) { Bytes b; leb128_write(v, b); memcpy(out, b.data(), b.size()); *n = b.size(); return ST_OK; }

int orc_bitwriter(int msb_first, const uint8_t* sizes, const uint64_t* values, uint64_t n, uint8_t* out, uint64_t* nout) {
  Bytes b;
  if (msb_first) { BitWriterMsb w(b); for (uint64_t i = 0; i < n; ++i) w.write_bits(sizes[i], values[i]); w.finish(); }
  else { BitWriterLsb w(b); for (uint64_t i = 0; i < n; ++i) w.write_bits(sizes[i], values[i]); w.finish(); }
  memcpy(out, b.data(), b.size());
  *nout = b.size();
  return ST_OK;
}

static int bytes_out(const Bytes& b, dxo_bytes* out) {
  out->data = (uint8_t*)malloc(b.size() ? b.size() : 1);
  if (!out->data) return -22;
  memcpy(out->data, b.data(), b.size());
  out->len = b.size();
  return ST_OK;
}

// RansCoder with a given (already normalised) table, symbols written in the given order.
int orc_rans_encode_raw(const uint64_t* freqs, uint64_t nfreq, uint32_t precision, const uint32_t* symbols, uint64_t n, dxo_bytes* out) {
  return guarded([&] {
    RansCoder c(std::vector<uint64_t>(freqs, freqs + nfreq), precision);
    for (uint64_t i = 0; i < n; ++i) c.write(symbols[i]);
    Bytes b = c.flush();
    if (bytes_out(b, out)) throw std::bad_alloc();
  });
}
int orc_rans_decode_raw(const uint64_t* freqs, uint64_t nfreq, uint32_t precision, const uint8_t* buf, uint64_t len, uint64_t n, uint32_t* out) {
  return guarded([&] {
    std::vector<uint32_t> cum(nfreq), slot((size_t)1 << precision);
    uint64_t c = 0;
    for (uint64_t i = 0; i < nfreq; ++i) { cum[i] = (uint32_t)c; for (uint64_t k = 0; k < freqs[i]; ++k) slot[c + k] = (uint32_t)i; c += freqs[i]; }
    size_t end = len;
    const uint64_t l_base = ((uint64_t)1 << precision) << 2;
    uint64_t state = ans_read_tail(buf, end) + l_base;
    for (uint64_t i = 0; i < n; ++i) {
      while (state < l_base) { if (end == 0) throw EncodeError(ST_INVALID_ARGUMENT, "NotEnoughData"); state = state * 256 + buf[--end]; }
      uint64_t q = state >> precision, r = state & (((uint64_t)1 << precision) - 1);
      uint32_t s = slot[r];
      state = q * freqs[s] + r - cum[s];
      out[i] = s;
    }
    while (end > 0) state = state * 256 + buf[--end];
    if (state != l_base) throw EncodeError(ST_INVALID_ARGUMENT, "rANS stream does not unwind to the initial state");
  });
}
int orc_rabs_encode(uint32_t zero_prob, const uint8_t* bits, uint64_t n, dxo_bytes* out) {
  return guarded([&] { RabsCoder c(zero_prob); for (uint64_t i = 0; i < n; ++i) c.write(bits[i]); Bytes b = c.flush(); if (bytes_out(b, out)) throw std::bad_alloc(); });
}
int orc_rabs_decode(uint32_t zero_prob, const uint8_t* buf, uint64_t len, uint64_t n, uint8_t* out) {
  return guarded([&] { auto v = rabs_decode(buf, len, zero_prob, n); memcpy(out, v.data(), n); });
}
// encode_symbols(symbols, _, DirectCoded, writer)
int orc_encode_symbols(const uint32_t* symbols, uint64_t n, dxo_bytes* out) {
  return guarded([&] { Bytes b; encode_symbols_direct(std::vector<uint32_t>(symbols, symbols + n), b); if (bytes_out(b, out)) throw std::bad_alloc(); });
}
int orc_decode_symbols(const uint8_t* buf, uint64_t len, uint64_t n, uint32_t* out, uint64_t* consumed) {
  return guarded([&] { size_t pos = 0; auto v = decode_symbols_direct(buf, len, n, pos); memcpy(out, v.data(), n * 4); *consumed = pos; });
}
// Traverser on the universal / attribute tables with literal stack removal toggled
// The inverse of the attribute path (orc_inverse.hpp): decodes `drc` against `mesh`. counts: 3 + 5 per attribute
// ([0] attributes, [1] first differing byte of header + connectivity or ~0, [2] bytes consumed; per attribute: values
// checked, mismatches, inconsistent writes, not invertible, unreferenced values); errors: per attribute max |dequantised -
// original| and its bound (0, 0 where the attribute is not coordinate-quantised).
int orc_decode_check(const dxo_mesh* mesh, const dxo_config* cfg, const uint8_t* drc, uint64_t len, uint64_t* counts, double* errors, uint32_t max_attributes) {
  return guarded([&] {
    Mesh m = mesh_from_c(mesh);
    OracleConfig oc = cfg_from_c(cfg, 0);
    InverseReport rep = decode_and_check(m, oc, drc, (size_t)len);
    counts[0] = rep.num_attributes; counts[1] = rep.prefix_mismatch_at; counts[2] = rep.consumed;
    for (size_t i = 0; i < rep.atts.size() && i < max_attributes; ++i) {
      const auto& a = rep.atts[i];
      uint64_t* c = counts + 3 + 5 * i;
      c[0] = a.values_checked; c[1] = a.mismatches; c[2] = a.inconsistent_writes; c[3] = a.not_invertible; c[4] = a.unreferenced;
      errors[2 * i] = a.max_abs_error; errors[2 * i + 1] = a.error_bound;
    }
  });
}
int orc_zero_prob(uint64_t n0, uint64_t len, int texcoord_variant) {
  float lf = texcoord_variant ? (float)len + 0.001f : (float)len;
  return zero_prob_f32(n0, lf);
}
int32_t orc_to_positive_i32(int32_t v) { return to_positive_i32(v); }
void orc_oct_quantize(float x, float y, float z, int32_t* out, int* status) {
  int st = guarded([&] { float u, v; octahedral_transform_f32(x, y, z, u, v); oct_quantize_uv(u, v, out[0], out[1]); });
  if (status) *status = st;
}
void orc_oct_transform(float x, float y, float z, float* out) { octahedral_transform_f32(x, y, z, out[0], out[1]); }

// compute_vec3_bounds / compute_vec4_bounds — io/gltf/encode.rs:815-899: the bounds start from point 0's value and fold
// f32::min / f32::max over points 1.. in order (Attribute::get maps the point through point_to_att_val_map,
// core/attribute/mod.rs:122-128, 216-230). f32::min / max return the other operand when one is NaN. For -0.0 against
// +0.0 the reference's result depends on LLVM's lowering of minnum / maxnum (unpinned); this restatement orders
// -0.0 below +0.0. Returns the number of points folded (0: the reference returns empty vectors, outputs untouched).
uint64_t orc_attribute_bounds(const float* values, uint64_t num_values, uint32_t ncomp, const uint32_t* point_to_value, uint64_t num_points,
                              float* out_min, float* out_max) {
  if (!point_to_value) num_points = num_values;
  if (num_points == 0) return 0;
  auto value_of = [&](uint64_t p, uint32_t k) { return values[(point_to_value ? point_to_value[p] : p) * ncomp + k]; };
  auto rust_min = [](float a, float b) {
    if (a != a) return b;
    if (b != b) return a;
    if (a == b) return std::signbit(a) ? a : b;  // -0.0 before +0.0
    return a < b ? a : b;
  };
  auto rust_max = [](float a, float b) {
    if (a != a) return b;
    if (b != b) return a;
    if (a == b) return std::signbit(a) ? b : a;
    return a > b ? a : b;
  };
  for (uint32_t k = 0; k < ncomp; ++k) { out_min[k] = value_of(0, k); out_max[k] = value_of(0, k); }
  for (uint64_t p = 1; p < num_points; ++p)
    for (uint32_t k = 0; k < ncomp; ++k) {
      out_min[k] = rust_min(out_min[k], value_of(p, k));
      out_max[k] = rust_max(out_max[k], value_of(p, k));
    }
  return num_points;
}

}  // extern "C"
