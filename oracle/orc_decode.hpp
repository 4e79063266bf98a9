// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_core.hpp header).
// Entropy DEcoders, written independently of the encoder code path so that
// encoder output can be proven decodable (round-trip checker). They follow the
// inverse operations specified by the reference's (uncompiled) decoder:
// decode/entropy/rans.rs:29-66 (RansDecoder), :70-128 (RabsDecoder),
// :146-195 (table parsing), decode/entropy/symbol_coding.rs:29-125.
#pragma once
#include "orc_core.hpp"

namespace orc {

inline uint64_t leb128_read(const uint8_t* b, size_t len, size_t& pos) {  // utils/bit_coder.rs:4-18
  uint64_t result = 0;
  unsigned shift = 0;
  for (;;) {
    if (pos >= len) throw EncodeError(ST_INVALID_ARGUMENT, "NotEnoughData");
    uint8_t byte = b[pos++];
    result |= (uint64_t)(byte & 0x7F) << shift;
    if (!(byte & 0x80)) break;
    shift += 7;
  }
  return result;
}

// reads the ANS tail written by ans_flush_state, walking backwards from `end`
inline uint64_t ans_read_tail(const uint8_t* b, size_t& end) {
  if (end == 0) throw EncodeError(ST_INVALID_ARGUMENT, "NotEnoughData");
  uint8_t metadata = b[--end];
  unsigned flag = metadata >> 6;
  uint64_t state = 0;
  for (unsigned k = 0; k < flag; ++k) {  // read_u8/u16/u24_back: little-endian value read from the end
    if (end == 0) throw EncodeError(ST_INVALID_ARGUMENT, "NotEnoughData");
    state = (state << 8) | b[--end];
  }
  state |= (uint64_t)(metadata & 0x3F) << (flag << 3);
  return state;
}

// decode_symbols (DirectCoded) — returns the symbols in forward order
inline std::vector<uint32_t> decode_symbols_direct(const uint8_t* b, size_t len, size_t num_symbols, size_t& pos) {
  if (pos >= len || b[pos++] != 1) throw EncodeError(ST_INVALID_ARGUMENT, "not DirectCoded");
  if (pos >= len) throw EncodeError(ST_INVALID_ARGUMENT, "NotEnoughData");
  unsigned bit_length = b[pos++];
  if (bit_length < 1 || bit_length > 18) throw EncodeError(ST_INVALID_ARGUMENT, "InvalidBitLength");
  unsigned P = rans_precision_for_bit_length(bit_length);
  size_t nsym = (size_t)leb128_read(b, len, pos);
  std::vector<uint32_t> freq(nsym, 0);
  for (size_t i = 0; i < nsym;) {
    if (pos >= len) throw EncodeError(ST_INVALID_ARGUMENT, "NotEnoughData");
    unsigned count = b[pos++];
    unsigned token = count & 3;
    if (token == 3) {
      size_t offset = count >> 2;
      if (i + offset >= nsym) throw EncodeError(ST_INVALID_ARGUMENT, "Invalid offset for frequency counts");
      for (size_t j = 0; j <= offset; ++j) freq[i + j] = 0;
      i += offset;
    } else {
      uint32_t c = count >> 2;
      for (unsigned j = 0; j < token; ++j) { if (pos >= len) throw EncodeError(ST_INVALID_ARGUMENT, "NotEnoughData"); c |= (uint32_t)b[pos++] << (8 * (j + 1) - 2); }
      freq[i] = c;
    }
    i += 1;
  }
  size_t payload = (size_t)leb128_read(b, len, pos);
  if (pos + payload > len) throw EncodeError(ST_INVALID_ARGUMENT, "NotEnoughData");
  // slot table
  std::vector<uint32_t> cum(nsym), slot((size_t)1 << P);
  { uint64_t c = 0;
    for (size_t i = 0; i < nsym; ++i) { cum[i] = (uint32_t)c; for (uint32_t k = 0; k < freq[i]; ++k) { if (c + k >= slot.size()) throw EncodeError(ST_RANS_FREQ_TABLE, "freq sum too large"); slot[c + k] = (uint32_t)i; } c += freq[i]; }
    if (c != ((uint64_t)1 << P)) throw EncodeError(ST_RANS_FREQ_TABLE, "Frequency count not compatible with RANS precision"); }
  const uint8_t* pb = b + pos;
  size_t end = payload;
  const uint64_t l_base = ((uint64_t)1 << P) << 2;
  uint64_t state = ans_read_tail(pb, end) + l_base;
  std::vector<uint32_t> out(num_symbols);
  for (size_t i = 0; i < num_symbols; ++i) {
    while (state < l_base) { if (end == 0) throw EncodeError(ST_INVALID_ARGUMENT, "NotEnoughData"); state = state * 256 + pb[--end]; }
    uint64_t q = state >> P, r = state & (((uint64_t)1 << P) - 1);
    uint32_t s = slot[r];
    state = q * freq[s] + r - cum[s];
    out[i] = s;
  }
  // Bytes emitted while the encoder still held (part of) its initial state are not needed
  // to decode the symbols; shifting them back in must reproduce the initial state exactly.
  while (end > 0) state = state * 256 + pb[--end];
  if (state != l_base) throw EncodeError(ST_INVALID_ARGUMENT, "rANS stream does not unwind to the initial state");
  pos += payload;
  return out;
}

// RabsDecoder::read — returns bits in the order they come out (reverse of write order)
inline std::vector<uint8_t> rabs_decode(const uint8_t* pb, size_t payload, unsigned zero_prob, size_t nbits) {
  const unsigned P = 8;
  const uint64_t L = 4096;
  size_t end = payload;
  uint64_t state = ans_read_tail(pb, end) + L;
  const uint64_t f0 = zero_prob, f1 = (1u << P) - f0;
  std::vector<uint8_t> out(nbits);
  for (size_t i = 0; i < nbits; ++i) {
    if (state < L) { if (end == 0) throw EncodeError(ST_INVALID_ARGUMENT, "NotEnoughData"); state = (state << 8) + pb[--end]; }
    uint64_t x = state, q = x >> P, r = x & ((1u << P) - 1), xn = q * f1;
    if (r < f1) { state = xn + r; out[i] = 1; }
    else { state = x - xn - f1; out[i] = 0; }
  }
  while (end > 0) state = (state << 8) + pb[--end];
  if (state != L) throw EncodeError(ST_INVALID_ARGUMENT, "rABS stream does not unwind to the initial state");
  return out;
}

}  // namespace orc
