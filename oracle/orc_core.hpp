// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the shipped product path.
//
// CPU restatement of draco-oxide's encoder data model, byte sinks and entropy
// coders. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may build, link or call anything in oracle/.
//
// PARITY STATUS: "parity unpinned" at whole-stream level. The reference cannot
// be compiled in this image (no cargo/rustc) and its own tests pin no .drc
// bytes. This restatement is pinned against every
// unit-level known answer the reference's tests hold (LEB128 bytes, BitWriter
// bytes, dedup maps, corner-table opposites / left-most corners, UV-seam
// corners, sequencer orders, OBJ loader faces) — see tests/test_oracle_known_answers.py.
//
// Every function cites the reference file:line it follows (paths relative to
// /root/reference/draco-oxide/src/).
#pragma once
#include <array>
#include <cstdint>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace orc {

constexpr uint32_t NONE = 0xFFFFFFFFu;

// Reference panics / unimplemented!() / assert!() map to Panic; Result::Err maps to EncodeError.
struct Panic : std::runtime_error {
  int code;
  Panic(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
struct EncodeError : std::runtime_error {
  int code;
  EncodeError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

// status codes shared with include/dxo.h (kept numerically identical)
enum {
  ST_OK = 0, ST_INVALID_ARGUMENT = -1, ST_UNSUPPORTED_INPUT = -2, ST_UNSUPPORTED_DATA_TYPE = -3,
  ST_UNSUPPORTED_NUM_COMPONENTS = -4, ST_TOO_MANY_ATTRIBUTES = -5, ST_RANS_INVALID_SYMBOL = -6,
  ST_RANS_STATE_TOO_LARGE = -7, ST_RANS_FREQ_TABLE = -8, ST_ZERO_NORMAL = -9, ST_UNUSED_VERTICES = -10,
  ST_INTERNAL = -99
};

// ---------------------------------------------------------------------------
// ids — core/attribute/mod.rs:527-716
enum AttType { AT_POSITION = 0, AT_NORMAL = 1, AT_COLOR = 2, AT_TEXCOORD = 3, AT_CUSTOM = 4, AT_TANGENT = 5,
               AT_MATERIAL = 6, AT_JOINT = 7, AT_WEIGHT = 8 };
enum CompType { CT_U8 = 1, CT_I8 = 2, CT_U16 = 3, CT_I16 = 4, CT_U32 = 5, CT_I32 = 6, CT_U64 = 7, CT_I64 = 8,
                CT_F32 = 9, CT_F64 = 10 };
inline size_t comp_size(uint32_t ct) {  // ComponentDataType::size, core/attribute/mod.rs:543-557
  switch (ct) {
    case CT_U8: case CT_I8: return 1;
    case CT_U16: case CT_I16: return 2;
    case CT_U32: case CT_I32: case CT_F32: return 4;
    case CT_U64: case CT_I64: case CT_F64: return 8;
    default: return 0;
  }
}

// Attribute — core/attribute/mod.rs:26-49. Raw AoS buffer of unique values plus
// the optional point -> value map.
struct Attribute {
  uint32_t id = 0;
  uint32_t att_type = AT_POSITION;
  uint32_t comp_type = CT_F32;
  uint32_t num_components = 3;
  uint32_t domain = 0;
  std::vector<uint32_t> parents;
  std::vector<uint8_t> buffer;  // num_unique * num_components * comp_size
  bool has_map = false;
  std::vector<uint32_t> map;

  size_t value_size() const { return comp_size(comp_type) * num_components; }
  size_t num_unique() const { size_t vs = value_size(); return vs ? buffer.size() / vs : 0; }
  // Attribute::len — core/attribute/mod.rs:195-202
  size_t len() const { return has_map ? map.size() : num_unique(); }
  // Attribute::get_unique_val_idx — core/attribute/mod.rs:209-224 (assert -> Panic)
  uint32_t unique_val_idx(uint32_t p) const {
    if (p >= len()) throw Panic(ST_INVALID_ARGUMENT, "Index out of bounds in get_unique_val_idx");
    return has_map ? map[p] : p;
  }
  // typed component read with DataValue::to_f64 semantics — core/shared.rs:375-470
  double comp_as_f64(size_t val, size_t c) const {
    const uint8_t* p = buffer.data() + val * value_size() + c * comp_size(comp_type);
    switch (comp_type) {
      case CT_F32: { float v; memcpy(&v, p, 4); return (double)v; }
      case CT_F64: { double v; memcpy(&v, p, 8); return v; }
      case CT_U8: return (double)*p;
      case CT_I8: return (double)*(const int8_t*)p;
      case CT_U16: { uint16_t v; memcpy(&v, p, 2); return (double)v; }
      case CT_I16: { int16_t v; memcpy(&v, p, 2); return (double)v; }
      case CT_U32: { uint32_t v; memcpy(&v, p, 4); return (double)v; }
      case CT_I32: { int32_t v; memcpy(&v, p, 4); return (double)v; }
      case CT_U64: { uint64_t v; memcpy(&v, p, 8); return (double)v; }
      case CT_I64: { int64_t v; memcpy(&v, p, 8); return (double)v; }
      default: throw EncodeError(ST_UNSUPPORTED_DATA_TYPE, "Unsupported data type");
    }
  }
};

struct Mesh {
  std::vector<std::array<uint32_t, 3>> faces;
  std::vector<Attribute> atts;
};

// Portabilized attribute: i32 AoS values + the original's point map.
// (Attribute::from_without_removing_duplicates + set_point_to_att_val_map,
//  encode/attribute/portabilization/quantization_coordinate_wise.rs:100-116)
struct PortAttribute {
  uint32_t id = 0;
  uint32_t att_type = 0;
  uint32_t n = 0;  // components
  std::vector<int32_t> vals;
  bool has_map = false;
  std::vector<uint32_t> map;
  size_t num_unique() const { return n ? vals.size() / n : 0; }
  size_t len() const { return has_map ? map.size() : num_unique(); }
  const int32_t* get(uint32_t p) const {  // Attribute::get, core/attribute/mod.rs:125-131
    if (p >= len()) throw Panic(ST_INVALID_ARGUMENT, "Index out of bounds in get");
    uint32_t v = has_map ? map[p] : p;
    return vals.data() + (size_t)v * n;
  }
};

// ---------------------------------------------------------------------------
// Byte sinks — core/bit_coder.rs:7-48 (Vec<u8> impl), utils/bit_coder.rs:20-33
using Bytes = std::vector<uint8_t>;
inline void w_u8(Bytes& b, uint8_t v) { b.push_back(v); }
inline void w_u16(Bytes& b, uint16_t v) { b.push_back((uint8_t)v); b.push_back((uint8_t)(v >> 8)); }
inline void w_u24(Bytes& b, uint32_t v) { b.push_back((uint8_t)v); b.push_back((uint8_t)(v >> 8)); b.push_back((uint8_t)(v >> 16)); }
inline void w_u32(Bytes& b, uint32_t v) { w_u16(b, (uint16_t)v); w_u16(b, (uint16_t)(v >> 16)); }
inline void w_i32(Bytes& b, int32_t v) { w_u32(b, (uint32_t)v); }
inline void w_f32(Bytes& b, float f) { uint32_t u; memcpy(&u, &f, 4); w_u32(b, u); }  // Portable for f32: core/shared.rs:476-481
inline void leb128_write(uint64_t value, Bytes& b) {  // utils/bit_coder.rs:20-33
  for (;;) {
    uint8_t byte = (uint8_t)(value & 0x7F);
    value >>= 7;
    if (value == 0) { b.push_back(byte); break; }
    b.push_back(byte | 0x80);
  }
}

// BitWriter<_, LsbFirst> — core/bit_coder.rs:90-188 (LsbFirst branches only; the
// encoder never instantiates MsbFirst on the hot path).
struct BitWriterLsb {
  Bytes& buf;
  uint8_t pos = 0;   // pos_in_curr_byte
  uint8_t curr = 0;  // curr_byte
  explicit BitWriterLsb(Bytes& b) : buf(b) {}
  void write_bits(uint8_t size, uint64_t value) {  // :113-178
    uint8_t offset = 0;
    if (pos != 0) {
      uint8_t rem = (uint8_t)(8 - pos);
      if (size <= rem) {
        curr |= (uint8_t)(value << pos);
        if (size == rem) { buf.push_back(curr); curr = 0; pos = 0; }
        else pos = (uint8_t)(pos + size);
        return;
      }
      curr |= (uint8_t)(value << pos);
      buf.push_back(curr);
      curr = 0;
      offset = rem;
    }
    unsigned iters = (unsigned)(size - offset) >> 3;
    for (unsigned i = 0; i < iters; ++i) { buf.push_back((uint8_t)(value >> offset)); offset = (uint8_t)(offset + 8); }
    curr = (uint8_t)(offset >= 64 ? 0 : (value >> offset));
    pos = (uint8_t)((size - offset) & 7);
  }
  void finish() { if (pos > 0) { buf.push_back(curr); pos = 0; curr = 0; } }  // Drop: :181-188
};

// BitWriter<_, MsbFirst> — only used by the known-answer tests (core/bit_coder.rs:514-590).
struct BitWriterMsb {
  Bytes& buf;
  uint8_t pos = 0, curr = 0;
  explicit BitWriterMsb(Bytes& b) : buf(b) {}
  void write_bits(uint8_t size, uint64_t value) {
    uint8_t offset = size;
    if (pos != 0) {
      uint8_t rem = (uint8_t)(8 - pos);
      if (size <= rem) {
        curr |= (uint8_t)((value & ((1u << rem) - 1)) << (rem - size));
        if (size == rem) { buf.push_back(curr); curr = 0; pos = 0; }
        else pos = (uint8_t)(pos + size);
        return;
      }
      curr |= (uint8_t)((value >> (size - rem)) & ((1u << rem) - 1));
      buf.push_back(curr);
      curr = 0;
      offset = (uint8_t)(size - rem);
    }
    unsigned iters = (unsigned)offset >> 3;
    for (unsigned i = 0; i < iters; ++i) { offset = (uint8_t)(offset - 8); buf.push_back((uint8_t)(value >> offset)); }
    curr = (uint8_t)((value & ((1ull << offset) - 1)) << (8 - offset));
    pos = offset;
  }
  void finish() { if (pos > 0) { buf.push_back(curr); pos = 0; curr = 0; } }
};

// ---------------------------------------------------------------------------
// Entropy coders.

// Rust `(f as u16)` etc: saturating, NaN -> 0.
inline uint16_t f32_as_u16(float f) {
  if (!(f == f)) return 0;
  if (f <= 0.0f) return 0;
  if (f >= 65535.0f) return 65535;
  return (uint16_t)f;
}
// zero_prob computation shared by the side streams:
// (((n0 as f32 / len as f32) * 256.0 + 0.5) as u16).clamp(1,255) as u8
// — edgebreaker.rs:595,640; mesh_normal_prediction.rs:151
inline uint8_t zero_prob_f32(size_t n0, float len_f) {
  float p = ((float)n0 / len_f) * 256.0f + 0.5f;
  uint16_t v = f32_as_u16(p);
  if (v < 1) v = 1;
  if (v > 255) v = 255;
  return (uint8_t)v;
}

// rANS / rABS flush tail — encode/entropy/rans.rs:48-68, :109-127
inline void ans_flush_state(uint64_t state, Bytes& out) {
  if (state < (1u << 6)) w_u8(out, (uint8_t)state);
  else if (state < (1u << 14)) w_u16(out, (uint16_t)((1u << 14) + (uint16_t)state));
  else if (state < (1u << 22)) w_u24(out, (2u << 22) + (uint32_t)state);
  else if (state < (1u << 30)) w_u32(out, (3u << 30) + (uint32_t)state);
  else throw EncodeError(ST_RANS_STATE_TOO_LARGE, "State too large for RANS coder");
}

// RabsCoder — encode/entropy/rans.rs:71-127 (RABS_PRECISION = 8, l_base = 4096)
struct RabsCoder {
  uint64_t state;
  uint64_t freq0;
  Bytes out;
  static constexpr unsigned P = 8;
  static constexpr uint64_t L = 4096;
  explicit RabsCoder(uint64_t freq_count_0) : state(L), freq0(freq_count_0) {}
  void write(uint8_t value) {  // :90-107
    uint64_t freq1 = (1u << P) - freq0;
    uint64_t freq = value > 0 ? freq1 : freq0;
    if (freq == 0) throw Panic(ST_INTERNAL, "attempt to divide by zero in RabsCoder");
    if (state >= (((L >> P) * freq) << 8)) { out.push_back((uint8_t)(state & 0xFF)); state >>= 8; }
    uint64_t q = state / freq, r = state % freq;
    state = (q << P) + r + (value > 0 ? 0 : freq1);
  }
  Bytes flush() {  // :109-127
    state -= L;
    ans_flush_state(state, out);
    return std::move(out);
  }
};

// RansCoder — encode/entropy/rans.rs:10-68; table: shared/entropy/mod.rs:41-64
struct RansCoder {
  unsigned P;
  uint64_t state, l_base;
  std::vector<uint32_t> freq, cum;
  Bytes out;
  RansCoder(const std::vector<uint64_t>& freq_counts, unsigned precision) : P(precision) {
    l_base = ((uint64_t)1 << P) << 2;
    state = l_base;
    uint64_t c = 0;
    freq.reserve(freq_counts.size()); cum.reserve(freq_counts.size());
    for (uint64_t f : freq_counts) { freq.push_back((uint32_t)f); cum.push_back((uint32_t)c); c += f; }
    if (c != ((uint64_t)1 << P))
      throw EncodeError(ST_RANS_FREQ_TABLE, "Frequency count not compatible with RANS precision");
  }
  void write(size_t idx) {  // :33-46
    if (idx >= freq.size()) throw EncodeError(ST_RANS_INVALID_SYMBOL, "Invalid symbol index");
    uint64_t f = freq[idx];
    if (f == 0) throw Panic(ST_INTERNAL, "attempt to divide by zero in RansCoder");
    while (state >= (((l_base >> P) * f) << 8)) { out.push_back((uint8_t)(state & 0xFF)); state >>= 8; }
    state = ((state / f) << P) + state % f + cum[idx];
  }
  Bytes flush() { state -= l_base; ans_flush_state(state, out); return std::move(out); }
};

// Result of RansSymbolEncoder::new's table construction — encode/entropy/rans.rs:146-230
struct RansTable {
  size_t num_symbols = 0;
  std::vector<uint64_t> distribution;  // normalised, sums to 2^P
  Bytes serialized;                    // leb128 #symbols + table bytes
};

inline RansTable rans_build_table(const std::vector<uint64_t>& freq_counts, unsigned P) {
  RansTable t;
  double total_freq = 0;  // iter().sum::<usize>() as f64
  { uint64_t s = 0; for (uint64_t f : freq_counts) s += f; total_freq = (double)s; }
  size_t num_symbols = 0;
  { bool found = false;
    for (size_t i = freq_counts.size(); i-- > 0;) if (freq_counts[i] > 0) { num_symbols = i + 1; found = true; break; }
    if (!found) throw Panic(ST_INTERNAL, "called Option::unwrap() on a None value (no symbols)"); }
  std::vector<uint64_t>& dist = t.distribution;
  dist.reserve(num_symbols);
  const uint64_t rans_precision = (uint64_t)1 << P;
  uint64_t total = 0;
  for (size_t i = 0; i < num_symbols; ++i) {
    uint64_t freq = freq_counts[i];
    double prob = (double)freq / total_freq;
    double x = prob * (double)rans_precision + 0.5;
    uint64_t nf = (x != x || x <= 0.0) ? 0 : (uint64_t)x;  // `as usize` saturating; values are small here
    if (nf == 0 && freq > 0) nf = 1;
    dist.push_back(nf);
    total += nf;
  }
  if (total != rans_precision) {
    // sort_by_key is a stable sort — rans.rs:171-175
    std::vector<uint32_t> sorted(num_symbols);
    for (size_t i = 0; i < num_symbols; ++i) sorted[i] = (uint32_t)i;
    std::stable_sort(sorted.begin(), sorted.end(), [&](uint32_t a, uint32_t b) { return dist[a] < dist[b]; });
    if (total < rans_precision) {
      dist[sorted.back()] += rans_precision - total;
    } else {
      uint64_t err = total - rans_precision;
      size_t i = dist.size() - 1;
      while (err > 0) {
        if (i == (size_t)-1) throw Panic(ST_INTERNAL, "index underflow in rANS normalisation");
        if (dist[sorted[i]] == 0) throw Panic(ST_INTERNAL, "attempt to subtract with overflow in rANS normalisation");
        dist[sorted[i]] -= 1;
        i -= 1;
        err -= 1;
      }
    }
  }
  t.num_symbols = num_symbols;
  // serialisation — rans.rs:193-230
  Bytes& w = t.serialized;
  leb128_write(num_symbols, w);
  size_t i = 0;
  while (i < num_symbols) {
    uint64_t freq = dist[i];
    if (freq == 0) {
      size_t offset = 0;
      while (offset < (1u << 6)) {
        if (i + offset + 1 >= dist.size()) throw Panic(ST_INTERNAL, "index out of bounds in zero-run scan");
        uint64_t next_prob = dist[i + offset + 1];
        if (next_prob > 0) { i += offset; break; }
        offset += 1;
      }
      w_u8(w, (uint8_t)(((uint8_t)offset << 2) | 3));  // (64u8 << 2) wraps to 0 -> byte 3 (Appendix B.7)
    } else {
      unsigned extra = 0;
      if (freq >= (1u << 6)) { extra++; if (freq >= (1u << 14)) { extra++; if (freq >= (1u << 22)) throw Panic(ST_INTERNAL, "RANS precision too high"); } }
      w_u8(w, (uint8_t)((freq << 2) | (extra & 3)));
      for (unsigned b = 0; b < extra; ++b) w_u8(w, (uint8_t)(freq >> (8 * (b + 1) - 2)));
    }
    i += 1;
  }
  return t;
}

// rANS precision table — encode/entropy/symbol_coding.rs:120-140
inline unsigned rans_precision_for_bit_length(unsigned bl) {
  static const unsigned tbl[19] = {0, 12, 12, 12, 12, 12, 12, 12, 12, 13, 15, 16, 18, 19, 20, 20, 20, 20, 20};
  return tbl[bl];
}

struct SymbolStreamTrace {
  unsigned bit_length = 0, precision = 0;
  std::vector<uint64_t> histogram;
  RansTable table;
  Bytes payload;
};

// encode_symbols(.., DirectCoded, ..) — encode/entropy/symbol_coding.rs:17-55,109-166
inline void encode_symbols_direct(const std::vector<uint32_t>& symbols, Bytes& w, SymbolStreamTrace* tr = nullptr) {
  w_u8(w, 1);  // SymbolEncodingMethod::DirectCoded — shared/entropy/mod.rs:30-36
  size_t num_nonzero = 0;
  for (uint32_t s : symbols) if (s > 0) num_nonzero++;
  // bit_length = (64 - leading_zeros(n) + 1).clamp(1,18) — symbol_coding.rs:118
  unsigned lz = num_nonzero == 0 ? 64 : (unsigned)__builtin_clzll((unsigned long long)num_nonzero);
  unsigned bit_length = 64 - lz + 1;
  if (bit_length < 1) bit_length = 1;
  if (bit_length > 18) bit_length = 18;
  w_u8(w, (uint8_t)bit_length);
  unsigned P = rans_precision_for_bit_length(bit_length);
  // histogram sized max_symbol + 1 — symbol_coding.rs:149-157
  std::vector<uint64_t> freq_counts;
  uint64_t max_symbol = 0;
  for (uint32_t s : symbols) {
    if (s >= max_symbol) { max_symbol = s; freq_counts.resize(max_symbol + 1, 0); }
    freq_counts[s] += 1;
  }
  RansTable table = rans_build_table(freq_counts, P);
  w.insert(w.end(), table.serialized.begin(), table.serialized.end());
  RansCoder coder(table.distribution, P);
  for (size_t i = symbols.size(); i-- > 0;) {  // fed last-to-first — symbol_coding.rs:161
    if (symbols[i] >= table.num_symbols) throw EncodeError(ST_RANS_INVALID_SYMBOL, "Invalid symbol index");
    coder.write(symbols[i]);
  }
  Bytes payload = coder.flush();
  leb128_write(payload.size(), w);  // RansSymbolEncoder::flush — rans.rs:248-255
  w.insert(w.end(), payload.begin(), payload.end());
  if (tr) { tr->bit_length = bit_length; tr->precision = P; tr->histogram = freq_counts; tr->table = table; tr->payload = payload; }
}

// to_positive_i32 — utils/mod.rs:152-158 (release-mode wrapping arithmetic)
inline int32_t to_positive_i32(int32_t val) {
  if (val >= 0) return (int32_t)((uint32_t)val << 1);
  uint32_t t = (uint32_t)(-(int64_t)((int64_t)val + 1));  // -(val+1), val+1 cannot overflow for val<0
  return (int32_t)((t << 1) + 1u);
}

}  // namespace orc
