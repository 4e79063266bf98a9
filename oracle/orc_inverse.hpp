// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_core.hpp header).
//
// The INVERSE of the attribute path: a decoder that takes a .drc stream, the mesh it claims to encode and
// nothing else from the encoder, and rebuilds every attribute value the way a Draco-style decoder must —
// CAUSALLY: element i of the sequence is predicted from values decoded before it (positions included: the normal and
// texture-coordinate predictors read the DECODED positions), the symbol stream supplies the correction, the side streams
// supply the flip / orientation choices the encoder made by looking at the actual value. Values not decoded yet are
// poisoned, so a predictor that peeks ahead cannot go unnoticed. The result is compared with the quantised
// attribute and, dequantised, with the original floats.
//
// What the reference offers for this direction (paths relative to /root/reference/draco-oxide/src/):
//   * decode/mod.rs:5-6, 28-36 — `mod connectivity` / `mod attribute` are commented out; decode() returns an empty mesh;
//   * decode/connectivity/spirale_reversi.rs:1208-1213 — the standard traversal is `unimplemented!()`; :199-560 is a
//     commented-out, half-translated copy of Google Draco's DecodeConnectivity;
//   * decode/attribute/attribute_decoder.rs:32-63 reads an older container (u16 id, u64 length, ...) that the current
//     encoder never writes; decode/attribute/inverse_prediction_transform/oct_orthogonal.rs:42 is `unimplemented!()`.
// So there is no reference decoder to restate. What IS shared and is followed here: the entropy decoders
// (decode/entropy/rans.rs:29-128, decode/entropy/symbol_coding.rs:29-125 — orc_decode.hpp), the prediction schemes
// (shared/attribute/prediction_scheme/*.rs — one predict() for both directions, attribute_decoder.rs:187-199 shows the
// causal loop), `pred + corr + metadata` (decode/attribute/inverse_prediction_transform/difference.rs:47-56) and the
// dequantisation rule (decode/attribute/portabilization/dequantization_rect_array.rs). The inverse transforms are the
// algebraic inverses of encode/attribute/prediction_transform/{wrapped_difference,oct_orthogonal,difference}.rs.
// Connectivity: the decoder re-derives the corner tables from the mesh it is given and requires the stream's header +
// connectivity section to equal the bytes those tables produce (byte-for-byte), then walks the same sequencer.
#pragma once
#include <memory>
#include "orc_attribute.hpp"
#include "orc_decode.hpp"

namespace orc {

struct InverseReport {
  uint32_t num_attributes = 0;
  uint64_t prefix_mismatch_at = ~0ull;  // first differing byte of header + connectivity (~0 = identical)
  uint64_t consumed = 0;                // bytes of the stream the decoder used (must equal its length)
  struct Att {
    // not_invertible: normals whose symbols have more than one preimage under the reference's transform (see the causal loop)
    uint64_t values_checked = 0, mismatches = 0, inconsistent_writes = 0, not_invertible = 0, unreferenced = 0;
    double max_abs_error = 0, error_bound = 0;
  };
  std::vector<Att> atts;
};

inline int32_t from_positive_i32(uint32_t s) {  // inverse of to_positive_i32 (utils/mod.rs:152-167)
  return (s & 1u) ? (int32_t)(0u - ((s >> 1) + 1u)) : (int32_t)(s >> 1);
}

namespace inverse_detail {
constexpr int32_t kPoison = 0x5A5A5A5A;

struct Reader {
  const uint8_t* b; size_t len, pos;
  uint8_t u8() { need(1); return b[pos++]; }
  uint32_t u32() { need(4); uint32_t v; memcpy(&v, b + pos, 4); pos += 4; return v; }
  int32_t i32() { return (int32_t)u32(); }
  float f32() { uint32_t u = u32(); float f; memcpy(&f, &u, 4); return f; }
  uint64_t leb() { return leb128_read(b, len, pos); }
  void need(size_t n) const { if (pos + n > len) throw EncodeError(ST_INVALID_ARGUMENT, "NotEnoughData"); }
};

// bits of a side stream in the order the encoder wrote them (the decoder pops them last-first)
inline std::vector<uint8_t> side_bits(Reader& r, size_t count) {
  const unsigned zero_prob = r.u8();
  const size_t size = (size_t)r.leb();
  r.need(size);
  std::vector<uint8_t> rev = rabs_decode(r.b + r.pos, size, zero_prob, count);
  r.pos += size;
  return std::vector<uint8_t>(rev.rbegin(), rev.rend());
}

// the diamond inversion of OctahedronOrthogonalTransform (oct_orthogonal.rs:41-52), applied to one centred vector
inline void invert_diamond(int32_t& a, int32_t& b) {
  const int32_t one = 255 / 2;
  const int32_t a0 = a, qs = -signum32(mul32(a, b));
  a = add32(mul32(qs, b), mul32(signum32(a), one));
  b = add32(mul32(qs, a0), mul32(signum32(b), one));
}
// forward transform of (orig, pred) -> symbols, exactly as the encoder does it (oracle encode_one_attribute)
inline void oct_forward(int32_t o0, int32_t o1, int32_t p0, int32_t p1, int32_t& c0, int32_t& c1) {
  const int32_t one = 255 / 2;
  p0 = sub32(p0, one); p1 = sub32(p1, one); o0 = sub32(o0, one); o1 = sub32(o1, one);
  if (add32(abs32(p0), abs32(p1)) > one) { invert_diamond(p0, p1); invert_diamond(o0, o1); }
  if (!(p0 == 0 && p1 == 0)) while (p0 >= 0 || p1 > 0) { int32_t t = p0; p0 = -p1; p1 = t; t = o0; o0 = -o1; o1 = t; }
  c0 = sub32(o0, p0); c1 = sub32(o1, p1);
  if (c0 < 0) c0 = add32(c0, 255);
  if (c1 < 0) c1 = add32(c1, 255);
}
// its inverse: (pred, symbols) -> orig. Returns false when the recovered value does not map back to the symbols (the
// diamond inversion loses information where a centred component is 0), leaving the closest candidate in o0 / o1.
inline bool oct_inverse(int32_t p0_in, int32_t p1_in, int32_t c0, int32_t c1, int32_t& o0, int32_t& o1) {
  const int32_t one = 255 / 2;
  int32_t p0 = sub32(p0_in, one), p1 = sub32(p1_in, one);
  const bool inverted = add32(abs32(p0), abs32(p1)) > one;
  if (inverted) invert_diamond(p0, p1);
  int turns = 0;
  if (!(p0 == 0 && p1 == 0)) while (p0 >= 0 || p1 > 0) { int32_t t = p0; p0 = -p1; p1 = t; ++turns; }
  int32_t a = add32(c0, p0), b = add32(c1, p1);  // rotated, centred original; 255 was added when the difference was negative
  if (a > one) a = sub32(a, 255);
  if (b > one) b = sub32(b, 255);
  for (int k = 0; k < turns; ++k) { int32_t t = a; a = b; b = -t; }  // undo the quarter turns: (x, y) <- (y, -x)
  if (inverted) invert_diamond(a, b);  // an involution wherever both components are non-zero before and after
  o0 = add32(a, one); o1 = add32(b, one);
  // Points on the border of the octahedral square come in pairs that name the same direction; the quantiser only ever
  // produces the canonical one (into_faithful_oct_quantization, geom.rs:137-157), and the transform maps both to the same
  // symbols, so the decoder canonicalises as well.
  if (o0 >= 0 && o0 <= 255 && o1 >= 0 && o1 <= 255) into_faithful_oct_quantization(o0, o1);
  int32_t f0, f1;
  oct_forward(o0, o1, p0_in, p1_in, f0, f1);
  if (f0 == c0 && f1 == c1) return true;
  // exhaustive search over the canonical points for a preimage (only reached where the diamond inversion meets a zero component)
  for (int32_t x = 0; x <= 255; ++x)
    for (int32_t y = 0; y <= 255; ++y) {
      int32_t cx = x, cy = y;
      into_faithful_oct_quantization(cx, cy);
      if (cx != x || cy != y) continue;
      oct_forward(x, y, p0_in, p1_in, f0, f1);
      if (f0 == c0 && f1 == c1) { o0 = x; o1 = y; return false; }
    }
  return false;
}

}  // namespace inverse_detail

// Decoder-side predictors: the oracle's Predictor for everything that does not look at the value being coded, plus the
// two schemes whose encoder side does (normal: the flip; texture coordinates: the orientation).
struct InversePredictor : Predictor {
  InversePredictor(const GenericCornerTable& c, unsigned n) : Predictor(c, n) {}

  // MeshNormalPrediction::predict without the flip decision (mesh_normal_prediction.rs:77-127)
  void normal_base(uint32_t c, int32_t* o) const {
    const int32_t* pc = pos->get(ct.point_idx(c));
    int32_t pos_c[3] = {pc[0], pc[1], pc[2]};
    uint32_t curr = c;
    for (uint32_t l; (l = ct.swing_left(curr)) != NONE;) { curr = l; if (curr == c) break; }
    const uint32_t start = curr;
    int64_t sum[3] = {0, 0, 0};
    face_normal(curr, pos_c, sum);
    for (uint32_t r; (r = ct.swing_right(curr)) != NONE;) { curr = r; if (curr == start) break; face_normal(curr, pos_c, sum); }
    const int64_t upper_bound = (int64_t)1 << 29;
    const int64_t abs_sum = add64(add64(abs64(sum[0]), abs64(sum[1])), abs64(sum[2]));
    if (abs_sum > upper_bound) { const int64_t q = abs_sum / upper_bound; for (int k = 0; k < 3; ++k) sum[k] = div64(sum[k], q); }
    const int32_t n3[3] = {wrap_i32(sum[0]), wrap_i32(sum[1]), wrap_i32(sum[2])};
    o[0] = o[1] = 0;
    if (!(n3[0] == 0 && n3[1] == 0 && n3[2] == 0)) { float u, v; octahedral_transform_i32(n3[0], n3[1], n3[2], u, v); oct_quantize_uv(u, v, o[0], o[1]); }
  }

  // MeshPredictionForTextureCoordinates::predict with the orientation supplied by the stream
  // (mesh_prediction_for_texture_coordinates.rs:107-219). `next_orientation` is consumed only on the main branch.
  void texcoord(uint32_t i, const SeqRecord& rec, const PortAttribute& att, const std::vector<uint8_t>& orient, size_t& next_orientation, int32_t* out) const {
    const uint32_t nc = c_next(i), pc = c_prev(i);
    const uint32_t next_pt = ct.point_idx(nc), prev_pt = ct.point_idx(pc), curr_pt = ct.point_idx(i);
    if (rec.contains(ct.vertex_idx(nc)) && rec.contains(ct.vertex_idx(pc))) {
      const int32_t* nu = att.get(next_pt); const int64_t next_uv[2] = {nu[0], nu[1]};
      const int32_t* pu = att.get(prev_pt); const int64_t prev_uv[2] = {pu[0], pu[1]};
      if (next_uv[0] == prev_uv[0] && next_uv[1] == prev_uv[1]) { out[0] = pu[0]; out[1] = pu[1]; return; }
      int64_t curr_pos[3], next_pos[3], prev_pos[3];
      position_for(curr_pt, curr_pos); position_for(next_pt, next_pos); position_for(prev_pt, prev_pos);
      const int64_t pn[3] = {sub64(prev_pos[0], next_pos[0]), sub64(prev_pos[1], next_pos[1]), sub64(prev_pos[2], next_pos[2])};
      const uint64_t pn_norm2 = (uint64_t)add64(add64(mul64(pn[0], pn[0]), mul64(pn[1], pn[1])), mul64(pn[2], pn[2]));
      if (pn_norm2 != 0) {
        const int64_t cn[3] = {sub64(curr_pos[0], next_pos[0]), sub64(curr_pos[1], next_pos[1]), sub64(curr_pos[2], next_pos[2])};
        const int64_t cn_dot_pn = add64(add64(mul64(pn[0], cn[0]), mul64(pn[1], cn[1])), mul64(pn[2], cn[2]));
        const int64_t pn_uv[2] = {sub64(prev_uv[0], next_uv[0]), sub64(prev_uv[1], next_uv[1])};
        const int64_t I64MAX = std::numeric_limits<int64_t>::max();
        if (std::max(abs64(next_uv[0]), abs64(next_uv[1])) > div64(I64MAX, (int64_t)pn_norm2)) { fallback_predict(i, rec, att, out); return; }
        if (abs64(cn_dot_pn) > div64(I64MAX, std::max(abs64(pn_uv[0]), abs64(pn_uv[1])))) { fallback_predict(i, rec, att, out); return; }
        const int64_t x_uv[2] = {add64(mul64(next_uv[0], (int64_t)pn_norm2), mul64(pn_uv[0], cn_dot_pn)),
                                 add64(mul64(next_uv[1], (int64_t)pn_norm2), mul64(pn_uv[1], cn_dot_pn))};
        if (abs64(cn_dot_pn) > div64(I64MAX, std::max(std::max(abs64(pn[0]), abs64(pn[1])), abs64(pn[2])))) { fallback_predict(i, rec, att, out); return; }
        int64_t cx[3];
        for (int k = 0; k < 3; ++k) cx[k] = sub64(curr_pos[k], add64(next_pos[k], div64(mul64(pn[k], cn_dot_pn), (int64_t)pn_norm2)));
        const uint64_t cx_norm2 = (uint64_t)add64(add64(mul64(cx[0], cx[0]), mul64(cx[1], cx[1])), mul64(cx[2], cx[2]));
        const uint64_t nrm = int_sqrt(cx_norm2 * pn_norm2);
        const int64_t cx_uv[2] = {mul64(pn_uv[1], (int64_t)nrm), mul64((int64_t)(0 - (uint64_t)pn_uv[0]), (int64_t)nrm)};
        if (next_orientation >= orient.size()) throw EncodeError(ST_INVALID_ARGUMENT, "orientation stream exhausted");
        const bool first = orient[next_orientation++] != 0;  // true: the candidate x_uv + cx_uv was the closer one
        const int64_t p[2] = {div64(first ? add64(x_uv[0], cx_uv[0]) : sub64(x_uv[0], cx_uv[0]), (int64_t)pn_norm2),
                              div64(first ? add64(x_uv[1], cx_uv[1]) : sub64(x_uv[1], cx_uv[1]), (int64_t)pn_norm2)};
        out[0] = wrap_i32(p[0]); out[1] = wrap_i32(p[1]);
        return;
      }
    }
    fallback_predict(i, rec, att, out);
  }
};

inline InverseReport decode_and_check(const Mesh& mesh, const OracleConfig& cfg, const uint8_t* drc, size_t len) {
  using namespace inverse_detail;
  InverseReport rep;
  // ---- header + connectivity: rebuilt from the mesh, required to be the stream's prefix
  Bytes head;
  for (char ch : std::string("DRACO")) w_u8(head, (uint8_t)ch);
  w_u8(head, 2); w_u8(head, 2); w_u8(head, 1); w_u8(head, 1); w_u16(head, 0);
  const Attribute* pos_att = nullptr;
  for (auto& a : mesh.atts) if (a.att_type == AT_POSITION) { pos_att = &a; break; }
  if (!pos_att) throw Panic(ST_INVALID_ARGUMENT, "no position attribute");
  CornerTable ct(mesh.faces, *pos_att);
  std::vector<AttributeCornerTable> att_tables;
  for (auto& a : mesh.atts) { if (a.att_type == AT_POSITION) continue; att_tables.emplace_back(ct, a); }
  Edgebreaker eb(ct, att_tables);
  const EdgebreakerOutput eo = eb.encode_connectivity(mesh.faces.size(), head);
  for (size_t k = 0; k < head.size(); ++k)
    if (k >= len || drc[k] != head[k]) { rep.prefix_mismatch_at = k; return rep; }

  Reader r{drc, len, head.size()};
  const auto& atts = mesh.atts;
  rep.num_attributes = r.u8();
  if (rep.num_attributes != atts.size()) throw EncodeError(ST_INVALID_ARGUMENT, "attribute count differs from the mesh");
  for (size_t i = 0; i < atts.size(); ++i) {
    if (r.u8() != (uint8_t)((uint8_t)i - 1) || r.u8() != (uint8_t)atts[i].domain || r.u8() != 0) throw EncodeError(ST_INVALID_ARGUMENT, "attribute decoder header");
  }
  std::vector<uint8_t> decoder_type(atts.size());
  for (size_t i = 0; i < atts.size(); ++i) {
    const bool ok = r.u8() == 1 && r.u8() == (uint8_t)atts[i].att_type && r.u8() == (uint8_t)atts[i].comp_type && r.u8() == (uint8_t)atts[i].num_components &&
                    r.u8() == 0 && r.u8() == (uint8_t)atts[i].id;
    if (!ok) throw EncodeError(ST_INVALID_ARGUMENT, "attribute description differs from the mesh");
    decoder_type[i] = r.u8();
  }
  rep.atts.resize(atts.size());
  std::vector<PortAttribute> decoded;  // in attribute order: the parents of later attributes
  for (size_t i = 0; i < atts.size(); ++i) {
    const Attribute& att = atts[i];
    InverseReport::Att& ar = rep.atts[i];
    const int scheme = r.u8(), transform = r.u8();
    if (r.u8() != 1) throw EncodeError(ST_INVALID_ARGUMENT, "rANS flag");
    // the attribute's corner table and sequence (the decoder's traversal; shared/attribute/sequence.rs:48-151)
    std::unique_ptr<RefAttributeCornerTable> rct;
    const GenericCornerTable* table = &ct;
    if (i > 0 && i - 1 < att_tables.size()) { rct.reset(new RefAttributeCornerTable(ct, att_tables[i - 1])); table = rct.get(); }
    const std::vector<uint32_t> sequence = compute_sequence(*table, eo.corners_of_edgebreaker, false);
    const unsigned N = decoder_type[i] == 3 ? 2u : att.num_components;
    const size_t M = sequence.size();
    const std::vector<uint32_t> symbols = decode_symbols_direct(drc, len, M * N, r.pos);

    // metadata, in the scheme's order (attribute_encoder.rs:362-386)
    int32_t wmin = 0, wmax = 0;
    std::vector<uint8_t> flips, orient;
    auto read_transform = [&] {
      if (transform == TR_WRAPPED) { wmin = r.i32(); wmax = r.i32(); }
      else if (transform == TR_OCT_ORTHOGONAL) { if (r.u32() != 255 || r.u32() != 127) throw EncodeError(ST_INVALID_ARGUMENT, "octahedral transform constants"); }
    };
    if (scheme == SCH_NORMAL) { read_transform(); flips = side_bits(r, M); }
    else if (scheme == SCH_TEXCOORD) {
      const uint32_t count = r.u32();
      const std::vector<uint8_t> delta = side_bits(r, count);  // delta[k] = (o[k] == o[k+1]), o[count] = true
      orient.assign(count, 0);
      bool nxt = true;
      for (size_t k = count; k-- > 0;) { const bool o = delta[k] ? nxt : !nxt; orient[k] = o; nxt = o; }
      read_transform();
    } else read_transform();
    std::vector<float> qmin; float qrange = 0; unsigned qbits = 0;
    if (decoder_type[i] == 2) { for (unsigned k = 0; k < att.num_components; ++k) qmin.push_back(r.f32()); qrange = r.f32(); qbits = r.u8(); }
    else if (decoder_type[i] == 3) { if (r.u8() != 8) throw EncodeError(ST_INVALID_ARGUMENT, "octahedral bits"); }

    // what the decoded values are compared with: the quantised attribute
    PortResult want;
    if (att.att_type == AT_NORMAL) want = quantize_octahedral(att);
    else if (att.att_type == AT_CUSTOM) want = to_bits(att);
    else want = quantize_coordinate_wise(att, att.att_type == AT_TEXCOORD ? cfg.texcoord_bits : att.att_type == AT_POSITION ? cfg.position_bits : cfg.generic_bits);

    // ---- the causal loop (attribute_decoder.rs:187-199)
    PortAttribute dec;
    dec.id = att.id; dec.att_type = att.att_type; dec.n = N;
    dec.vals.assign(att.num_unique() * N, kPoison);
    dec.has_map = att.has_map; dec.map = att.map;
    std::vector<uint8_t> written(att.num_unique(), 0);
    InversePredictor pred(*table, N);
    if (scheme == SCH_NORMAL || scheme == SCH_TEXCOORD) {
      if (att.parents.empty()) throw EncodeError(ST_INVALID_ARGUMENT, "missing parent");
      for (auto& pa : decoded) if (pa.id == att.parents[0]) pred.pos = &pa;
      if (!pred.pos) throw EncodeError(ST_INVALID_ARGUMENT, "parent not decoded yet");
    }
    SeqRecord rec;
    rec.init(table->num_vertices(), false);
    const int32_t diff = sub32(wmax, wmin), max_diff = add32(1, diff);
    size_t next_orientation = 0;
    for (size_t k = 0; k < M; ++k) {
      const uint32_t c = sequence[k];
      int32_t p[4] = {0, 0, 0, 0}, v[4] = {0, 0, 0, 0};
      switch (scheme) {
        case SCH_PARALLELOGRAM: pred.predict_parallelogram(c, rec, dec, p); break;
        case SCH_NORMAL: pred.normal_base(c, p); if (flips[k]) { p[0] = mul32(p[0], -1); p[1] = mul32(p[1], -1); } break;
        case SCH_TEXCOORD: pred.texcoord(c, rec, dec, orient, next_orientation, p); break;
        default: pred.predict_delta(c, rec, dec, p); break;
      }
      if (transform == TR_WRAPPED) {  // inverse of wrapped_difference.rs:66-94
        for (unsigned j = 0; j < N; ++j) {
          const int32_t pv = p[j] < wmin ? wmin : (p[j] > wmax ? wmax : p[j]);
          int32_t o = add32(pv, from_positive_i32(symbols[k * N + j]));
          if (o > wmax) o = sub32(o, max_diff); else if (o < wmin) o = add32(o, max_diff);
          v[j] = o;
        }
      } else if (transform == TR_OCT_ORTHOGONAL) {
        // The reference's transform is not injective: when the prediction lies outside the diamond, every original with a
        // centred zero component collapses onto (+-127, 0) / (0, +-127) (signum(0) = 0 in the diamond inversion,
        // oct_orthogonal.rs:41-52), so such a normal cannot be recovered from its symbols by ANY decoder. Those are counted
        // apart: the decoded value differs from the quantised normal, but the quantised normal maps to exactly the symbols
        // in the stream, i.e. the stream is what the reference's encoder writes for it.
        oct_inverse(p[0], p[1], (int32_t)symbols[2 * k], (int32_t)symbols[2 * k + 1], v[0], v[1]);
        const int32_t* w = want.att.vals.data() + (size_t)att.unique_val_idx(table->point_idx(c)) * 2;
        if (v[0] != w[0] || v[1] != w[1]) {
          int32_t f0, f1;
          oct_forward(w[0], w[1], p[0], p[1], f0, f1);
          if (f0 == (int32_t)symbols[2 * k] && f1 == (int32_t)symbols[2 * k + 1]) { ++ar.not_invertible; v[0] = w[0]; v[1] = w[1]; }
        }
      } else {
        for (unsigned j = 0; j < N; ++j) v[j] = add32(p[j], from_positive_i32(symbols[k * N + j]));  // difference.rs:47-56
      }
      const uint32_t vi = att.unique_val_idx(table->point_idx(c));
      if (written[vi]) { for (unsigned j = 0; j < N; ++j) if (dec.vals[(size_t)vi * N + j] != v[j]) { ++ar.inconsistent_writes; break; } }
      for (unsigned j = 0; j < N; ++j) dec.vals[(size_t)vi * N + j] = v[j];
      written[vi] = 1;
      rec.push(table->vertex_idx(c));
    }
    if (scheme == SCH_TEXCOORD && next_orientation != orient.size()) throw EncodeError(ST_INVALID_ARGUMENT, "orientation stream not consumed");

    // ---- against the quantised attribute, and dequantised against the original floats
    if (decoder_type[i] == 2) {
      for (unsigned k = 0; k < att.num_components; ++k) if (memcmp(&qmin[k], &want.min_values[k], 4) != 0) ++ar.mismatches;
      if (memcmp(&qrange, &want.range, 4) != 0) ++ar.mismatches;
    }
    for (size_t u = 0; u < att.num_unique(); ++u) {
      if (!written[u]) { ++ar.unreferenced; continue; }
      ++ar.values_checked;
      bool same = true;
      for (unsigned j = 0; j < N; ++j) same &= dec.vals[u * N + j] == want.att.vals[u * N + j];
      if (!same) { ++ar.mismatches; continue; }
      if (decoder_type[i] == 2) {  // dequantization_rect_array.rs: min + q * range / (2^bits - 1)
        const double step = qbits ? (double)qrange / (double)((1ull << qbits) - 1) : 0.0;
        // half a step, plus what the quantiser's four f32 roundings ((v - min) / range * maxq + 0.5) can add
        ar.error_bound = 0.5 * step * (1.0 + 1e-3) + (double)qrange * 0x1p-21 + 1e-30;
        for (unsigned j = 0; j < N; ++j) {
          const double back = (double)qmin[j] + (double)dec.vals[u * N + j] * step;
          ar.max_abs_error = std::max(ar.max_abs_error, std::fabs(back - att.comp_as_f64(u, j)));
        }
      }
    }
    decoded.push_back(std::move(dec));
  }
  rep.consumed = r.pos;
  return rep;
}

}  // namespace orc
