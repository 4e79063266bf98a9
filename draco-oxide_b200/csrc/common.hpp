// Product host code — shared definitions, byte sinks and the binary rANS (rABS) coder.
// Reference paths are relative to /root/reference/draco-oxide/src/.
#pragma once
#include <cstdint>
#include <cstdlib>
#if defined(__x86_64__)
#include <immintrin.h>
#endif
#include <cstring>
#include <memory>
#include <utility>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/dxo.h"

namespace dxo {

constexpr uint32_t kNone = 0xFFFFFFFFu;

// std::vector whose resize() leaves new elements uninitialised: the big connectivity arrays are
// filled right after being sized (by a copy from the device or by a pass that writes every entry),
// so the zero-fill of a plain vector would only add a memory pass.
// Large blocks (>= kHostBlockMin bytes) come from a process-wide pool of 2 MB-aligned, huge-page-advised blocks that are
// kept for reuse (connectivity.cpp): the per-mesh tables are tens of megabytes, and taking them from the C library means a
// fresh mmap, a page fault per 4 KB and a munmap with its TLB shoot-down on every call — which serialises concurrent
// encoder threads — while the order-dependent walks over them miss the TLB at every step unless the pages are large.
constexpr size_t kHostBlockMin = 256u << 10;
void* host_block_take(size_t bytes);
void host_block_give(void* p) noexcept;

template <class T>
struct NoInitAllocator : std::allocator<T> {
  template <class U> struct rebind { using other = NoInitAllocator<U>; };
  NoInitAllocator() = default;
  template <class U> NoInitAllocator(const NoInitAllocator<U>&) {}
  T* allocate(size_t n) {
    if (n * sizeof(T) >= kHostBlockMin) return static_cast<T*>(host_block_take(n * sizeof(T)));
    return static_cast<T*>(::operator new(n * sizeof(T)));
  }
  void deallocate(T* p, size_t n) noexcept {
    if (n * sizeof(T) >= kHostBlockMin) host_block_give(p);
    else ::operator delete(p);
  }
  template <class U, class... Args> void construct(U* p, Args&&... args) {
    if constexpr (sizeof...(Args) == 0) ::new ((void*)p) U;
    else ::new ((void*)p) U(std::forward<Args>(args)...);
  }
};
using U32Array = std::vector<uint32_t, NoInitAllocator<uint32_t>>;
using U8Array = std::vector<uint8_t, NoInitAllocator<uint8_t>>;

// Array of the connectivity tables. Normally it owns a no-init vector; with a memory source set it lives in
// caller-provided memory instead (pinned host blocks, so that the device copies of K12-K14 land in it by DMA
// and no second host copy is needed). The source outlives the array; growing past the block copies into an
// owned vector.
template <class T>
class HostArray {
 public:
  HostArray() = default;
  HostArray(const HostArray&) = delete;
  HostArray& operator=(const HostArray&) = delete;
  HostArray(HostArray&&) = default;             // a moved vector keeps its buffer, so p_ stays valid
  HostArray& operator=(HostArray&&) = default;
  using Source = void* (*)(void* user, size_t bytes);  // returns memory for `bytes` bytes or nullptr
  void set_source(Source fn, void* user) { source_ = fn; source_user_ = user; }
  // lives in caller-owned memory from now on (contents kept as they are); the memory outlives the array
  void adopt(T* p, size_t n) { own_.clear(); p_ = p; n_ = cap_ = n; }
  T* data() { return p_; }
  const T* data() const { return p_; }
  size_t size() const { return n_; }
  bool empty() const { return n_ == 0; }
  T& operator[](size_t i) { return p_[i]; }
  const T& operator[](size_t i) const { return p_[i]; }
  T* begin() { return p_; }
  T* end() { return p_ + n_; }
  const T* begin() const { return p_; }
  const T* end() const { return p_ + n_; }
  void resize(size_t n) {  // contents unspecified unless shrinking inside the current storage
    if (n <= cap_) { n_ = n; return; }
    if (source_) {
      if (void* ext = source_(source_user_, n * sizeof(T))) { own_.clear(); p_ = (T*)ext; n_ = cap_ = n; return; }
    }
    own_.resize(n);
    p_ = own_.data(); n_ = cap_ = n;
  }
  void assign(size_t n, T v) { resize(n); for (size_t i = 0; i < n; ++i) p_[i] = v; }
  void clear() { n_ = 0; }
  void reserve(size_t n) { if (n > cap_) { const size_t keep = n_; grow(n); n_ = keep; } }
  void push_back(T v) {
    if (n_ == cap_) grow(cap_ ? cap_ * 2 : 16);
    p_[n_++] = v;
  }
 private:
  void grow(size_t cap) {  // into an owned vector, keeping the contents
    std::vector<T, NoInitAllocator<T>> bigger(cap);
    if (n_) memcpy(bigger.data(), p_, n_ * sizeof(T));
    own_.swap(bigger);
    p_ = own_.data(); cap_ = cap;
  }
  T* p_ = nullptr;
  size_t n_ = 0, cap_ = 0;
  std::vector<T, NoInitAllocator<T>> own_;
  Source source_ = nullptr;
  void* source_user_ = nullptr;
};

struct Error : std::runtime_error {
  int status;
  Error(int st, const std::string& what) : std::runtime_error(what), status(st) {}
};

// corner navigation inside a triangle (core/corner_table/mod.rs:503-523)
static inline uint32_t corner_next(uint32_t c) { return (c % 3u == 2u) ? c - 2u : c + 1u; }
static inline uint32_t corner_prev(uint32_t c) { return (c % 3u == 0u) ? c + 2u : c - 1u; }

// Append-only little-endian byte sink (ByteWriter for Vec<u8>, core/bit_coder.rs:28-48).
class ByteSink {
 public:
  std::vector<uint8_t> data;
  void u8(uint8_t v) { data.push_back(v); }
  void u16(uint16_t v) { data.push_back((uint8_t)v); data.push_back((uint8_t)(v >> 8)); }
  void u24(uint32_t v) { u8((uint8_t)v); u8((uint8_t)(v >> 8)); u8((uint8_t)(v >> 16)); }
  void u32(uint32_t v) { u16((uint16_t)v); u16((uint16_t)(v >> 16)); }
  void i32(int32_t v) { u32((uint32_t)v); }
  void f32(float f) { uint32_t b; memcpy(&b, &f, 4); u32(b); }
  void bytes(const uint8_t* p, size_t n) { data.insert(data.end(), p, p + n); }
  void bytes(const std::vector<uint8_t>& v) { data.insert(data.end(), v.begin(), v.end()); }
  // LEB128 (utils/bit_coder.rs:20-33)
  void varint(uint64_t v) {
    do {
      uint8_t b = (uint8_t)(v & 0x7F);
      v >>= 7;
      data.push_back(v ? (uint8_t)(b | 0x80) : b);
    } while (v);
  }
  size_t size() const { return data.size(); }
};

// LSB-first bit packer with the reference BitWriter<LsbFirst> byte behaviour
// (core/bit_coder.rs:113-188): bits fill each byte from bit 0 upwards; a partial
// last byte is flushed by finish().
class BitPacker {
 public:
  explicit BitPacker(ByteSink& s) : sink_(s) {}
  void put(unsigned nbits, uint32_t value) {
    acc_ |= (uint64_t)value << fill_;
    fill_ += nbits;
    while (fill_ >= 8) { sink_.u8((uint8_t)acc_); acc_ >>= 8; fill_ -= 8; }
  }
  void finish() { if (fill_) { sink_.u8((uint8_t)acc_); acc_ = 0; fill_ = 0; } }
 private:
  ByteSink& sink_;
  uint64_t acc_ = 0;
  unsigned fill_ = 0;
};

// Tail of an ANS stream: state - base with a 2-bit length tag (encode/entropy/rans.rs:48-68).
static inline void ans_write_tail(uint32_t x, ByteSink& out) {
  if (x < (1u << 6)) out.u8((uint8_t)x);
  else if (x < (1u << 14)) out.u16((uint16_t)(0x4000u + x));
  else if (x < (1u << 22)) out.u24(0x800000u + x);
  else if (x < (1u << 30)) out.u32(0xC0000000u + x);
  else throw Error(DXO_ERR_RANS_STATE_TOO_LARGE, "rANS state too large");
}

// Probability of a zero bit used by all binary side streams (Appendix B.8 of SURVEY.md):
// (((n0 as f32 / len as f32) * 256.0 + 0.5) as u16).clamp(1, 255) — all in f32;
// an empty stream gives NaN -> 0 -> 1. edgebreaker.rs:595,640; mesh_normal_prediction.rs:151.
static inline uint8_t side_stream_zero_prob(uint64_t zeros, float len_as_f32) {
  volatile float ratio = (float)zeros / len_as_f32;  // volatile: keep the f32 rounding steps separate
  volatile float scaled = ratio * 256.0f;
  float p = scaled + 0.5f;
  uint32_t v;
  if (!(p == p) || p <= 0.0f) v = 0;
  else if (p >= 65535.0f) v = 65535;
  else v = (uint32_t)p;
  if (v < 1) v = 1;
  if (v > 255) v = 255;
  return (uint8_t)v;
}

// Largest element of an index array (the range checks of faces and point maps). The library is built without -march
// flags, so the AVX2 body is a per-function target picked at run time.
#if defined(__x86_64__)
__attribute__((target("avx2"))) static inline uint32_t max_u32_avx2(const uint32_t* p, size_t n) {
  __m256i m0 = _mm256_setzero_si256(), m1 = m0, m2 = m0, m3 = m0;
  size_t i = 0;
  for (; i + 32 <= n; i += 32) {
    m0 = _mm256_max_epu32(m0, _mm256_loadu_si256((const __m256i*)(p + i)));
    m1 = _mm256_max_epu32(m1, _mm256_loadu_si256((const __m256i*)(p + i + 8)));
    m2 = _mm256_max_epu32(m2, _mm256_loadu_si256((const __m256i*)(p + i + 16)));
    m3 = _mm256_max_epu32(m3, _mm256_loadu_si256((const __m256i*)(p + i + 24)));
  }
  m0 = _mm256_max_epu32(_mm256_max_epu32(m0, m1), _mm256_max_epu32(m2, m3));
  alignas(32) uint32_t t[8];
  _mm256_store_si256((__m256i*)t, m0);
  uint32_t mx = 0;
  for (int k = 0; k < 8; ++k) mx = t[k] > mx ? t[k] : mx;
  for (; i < n; ++i) mx = p[i] > mx ? p[i] : mx;
  return mx;
}
#endif
static inline uint32_t max_u32(const uint32_t* p, size_t n) {
#if defined(__x86_64__)
  static const bool avx2 = __builtin_cpu_supports("avx2") && !getenv("DXO_NO_AVX2");
  if (avx2) return max_u32_avx2(p, n);
#endif
  uint32_t m[4] = {0, 0, 0, 0};
  size_t i = 0;
  for (; i + 4 <= n; i += 4) for (int k = 0; k < 4; ++k) m[k] = p[i + k] > m[k] ? p[i + k] : m[k];
  uint32_t mx = 0;
  for (; i < n; ++i) mx = p[i] > mx ? p[i] : mx;
  for (int k = 0; k < 4; ++k) mx = m[k] > mx ? m[k] : mx;
  return mx;
}

// Binary rANS coder, precision 8, base 4096 (RabsCoder, encode/entropy/rans.rs:71-127).
// Encodes a whole bit sequence in the order given and returns the byte stream.
static inline void rabs_encode(const uint8_t* bits, size_t n, bool reversed, uint8_t zero_prob, ByteSink& out) {
  const uint32_t f0 = zero_prob, f1 = 256u - f0;
  uint32_t x = 4096u;
  auto step = [&](uint8_t bit) {
    const uint32_t f = bit ? f1 : f0;
    if (f == 0) throw Error(DXO_ERR_UNSUPPORTED_INPUT, "rABS: zero frequency");
    if (x >= ((16u * f) << 8)) { out.u8((uint8_t)x); x >>= 8; }  // single renormalisation step (:98-101)
    const uint32_t q = x / f, r = x - q * f;
    x = (q << 8) + r + (bit ? 0u : f1);
  };
  if (reversed) for (size_t i = n; i-- > 0;) step(bits[i]);
  else for (size_t i = 0; i < n; ++i) step(bits[i]);
  ans_write_tail(x - 4096u, out);
}

// Same coder, tuned for the long per-element side streams of the attribute path (flips,
// orientations): reciprocal multiplication instead of division (exact: x < 2^20, f <= 255,
// m = ceil(2^32 / f) => x * (m * f - 2^32) < 2^28 < 2^32), bytes written through a raw pointer.
// `bit(i)` yields the i-th bit to code (0 / 1), so a caller can derive the bits on the fly.
// Long streams that are almost all one value (rabs.cpp): runs of the common value are crossed by table look-ups, only the
// rare bits are coded one by one. Same bytes as the loops below.
bool rabs_sparse_applies(size_t n, uint8_t zero_prob, size_t rare_count);
void rabs_encode_sparse(size_t n, uint8_t zero_prob, uint32_t rare_bit, const uint32_t* rare_pos, size_t rare_count, std::vector<uint8_t>& out);

template <class BitAt>
static inline void rabs_encode_forward_fn(size_t n, uint8_t zero_prob, std::vector<uint8_t>& out, BitAt bit) {
  const uint32_t f0 = zero_prob, f1 = 256u - f0;
  if ((f0 == 0 || f1 == 0) && n) {  // only reachable with a probability outside [1,255]
    for (size_t i = 0; i < n; ++i) if ((bit(i) ? f1 : f0) == 0) throw Error(DXO_ERR_UNSUPPORTED_INPUT, "rABS: zero frequency");
  }
  if (f0 && f1 && rabs_sparse_applies(n, zero_prob, 0)) {  // skewed enough by its probability: where are the rare bits?
    const uint32_t rare_bit = f1 < f0 ? 1u : 0u;
    const size_t most = n / 64;
    // branch-free append (the slot is written for every bit, the count advances on a rare one); the budget is checked per block
    std::vector<uint32_t, NoInitAllocator<uint32_t>> rare(most + 4096 + 1);
    uint32_t* const buf = rare.data();
    size_t k = 0;
    for (size_t i0 = 0; i0 < n && k <= most; i0 += 4096) {
      const size_t i1 = i0 + 4096 < n ? i0 + 4096 : n;
      for (size_t i = i0; i < i1; ++i) { buf[k] = (uint32_t)i; k += (uint32_t)(bit(i) != 0) == rare_bit; }
    }
    if (k <= most && rabs_sparse_applies(n, zero_prob, k)) { rabs_encode_sparse(n, zero_prob, rare_bit, buf, k, out); return; }
  }
  // x' = (q << 8) + (x - q f) + cum = x + q (256 - f) + cum, q = floor(x / f) = (x * m) >> 32
  const uint32_t thr[2] = {f0 << 12, f1 << 12}, g[2] = {256u - f0, 256u - f1}, cum[2] = {f1, 0u};
  const uint64_t m[2] = {f0 ? ((1ull << 32) + f0 - 1) / f0 : 0, f1 ? ((1ull << 32) + f1 - 1) / f1 : 0};
  if (out.size() < n + 8) out.resize(n + 8);  // at most one byte per bit, plus the tail (capacity is kept across calls)
  uint8_t* p = out.data();
  uint32_t x = 4096u;
  if (g[0] == 1u || g[1] == 1u) {
    // One value has probability 255/256 (flags of smooth meshes): the loop branches on the bit (predictable) and
    // the common case needs no second multiply, x' = x + q + cum. The chain per bit drops from ~10 to ~6 cycles.
    const uint32_t common = g[0] == 1u ? 0u : 1u, rare = common ^ 1u;
    const uint32_t thr_c = thr[common], cum_c = cum[common], thr_r = thr[rare], g_r = g[rare], cum_r = cum[rare];
    const uint64_t m_c = m[common], m_r = m[rare];
    for (size_t i = 0; i < n; ++i) {
      if (__builtin_expect((uint32_t)(bit(i) != 0) == common, 1)) {
        if (x >= thr_c) { *p++ = (uint8_t)x; x >>= 8; }
        x = x + (uint32_t)(((uint64_t)x * m_c) >> 32) + cum_c;
      } else {
        if (x >= thr_r) { *p++ = (uint8_t)x; x >>= 8; }
        x = x + (uint32_t)(((uint64_t)x * m_r) >> 32) * g_r + cum_r;
      }
    }
  } else {
    for (size_t i = 0; i < n; ++i) {
      const uint32_t b = bit(i) != 0;
      if (x >= thr[b]) { *p++ = (uint8_t)x; x >>= 8; }
      const uint32_t q = (uint32_t)(((uint64_t)x * m[b]) >> 32);
      x = x + q * g[b] + cum[b];
    }
  }
  ByteSink tail;
  ans_write_tail(x - 4096u, tail);
  for (uint8_t b : tail.data) *p++ = b;
  out.resize((size_t)(p - out.data()));
}
static inline void rabs_encode_forward(const uint8_t* bits, size_t n, uint8_t zero_prob, std::vector<uint8_t>& out) {
  if (zero_prob && rabs_sparse_applies(n, zero_prob, 0)) {
    // flags in a byte array: eight at a time, only words that hold a rare flag are looked at (0.15 ns per flag)
    const uint32_t rare_bit = (256u - zero_prob) < zero_prob ? 1u : 0u;
    const uint64_t common_word = rare_bit ? 0ull : 0x0101010101010101ull;
    const size_t most = n / 64;
    std::vector<uint32_t, NoInitAllocator<uint32_t>> rare(most + 16);
    uint32_t* const buf = rare.data();
    size_t k = 0, i = 0;
    for (; i + 8 <= n && k <= most; i += 8) {
      uint64_t w;
      memcpy(&w, bits + i, 8);
      if (w == common_word) continue;
      for (size_t j = i; j < i + 8; ++j) { buf[k] = (uint32_t)j; k += (uint32_t)(bits[j] != 0) == rare_bit; }
    }
    for (; i < n && k <= most; ++i) { buf[k] = (uint32_t)i; k += (uint32_t)(bits[i] != 0) == rare_bit; }
    if (k <= most && rabs_sparse_applies(n, zero_prob, k)) { rabs_encode_sparse(n, zero_prob, rare_bit, buf, k, out); return; }
  }
  rabs_encode_forward_fn(n, zero_prob, out, [bits](size_t i) { return bits[i]; });
}

// Two independent streams coded in ONE loop: the coder is a serial chain of ~6 cycles per bit, two chains interleave in
// the pipeline and finish in little more than the time of one. Used when host threads are scarce (one thread per resident
// session instead of three). Same bytes as rabs_encode_forward_fn on each stream.
struct RabsChain {
  uint32_t x = 4096u;
  uint32_t thr[2], g[2], cum[2];
  uint64_t m[2];
  uint8_t* p = nullptr;
  void init(uint8_t zero_prob, uint8_t* out) {
    const uint32_t f0 = zero_prob, f1 = 256u - f0;
    thr[0] = f0 << 12; thr[1] = f1 << 12;
    g[0] = 256u - f0; g[1] = 256u - f1;
    cum[0] = f1; cum[1] = 0u;
    m[0] = f0 ? ((1ull << 32) + f0 - 1) / f0 : 0; m[1] = f1 ? ((1ull << 32) + f1 - 1) / f1 : 0;
    p = out;
  }
  inline void put(uint32_t b) {
    if (x >= thr[b]) { *p++ = (uint8_t)x; x >>= 8; }
    x = x + (uint32_t)(((uint64_t)x * m[b]) >> 32) * g[b] + cum[b];
  }
  void finish(std::vector<uint8_t>& out) {
    ByteSink tail;
    ans_write_tail(x - 4096u, tail);
    for (uint8_t b : tail.data) *p++ = b;
    out.resize((size_t)(p - out.data()));
  }
};
template <class BitA, class BitB>
static inline void rabs_encode_forward_pair(size_t na, uint8_t pa, std::vector<uint8_t>& outa, BitA bita, size_t nb, uint8_t pb, std::vector<uint8_t>& outb, BitB bitb) {
  if (pa == 0 || pb == 0) throw Error(DXO_ERR_UNSUPPORTED_INPUT, "rABS: zero frequency");  // zero_prob is clamped to [1, 255] by its callers
  if (outa.size() < na + 8) outa.resize(na + 8);
  if (outb.size() < nb + 8) outb.resize(nb + 8);
  RabsChain a, b;
  a.init(pa, outa.data());
  b.init(pb, outb.data());
  const size_t both = na < nb ? na : nb;
  for (size_t i = 0; i < both; ++i) { a.put(bita(i) != 0); b.put(bitb(i) != 0); }
  for (size_t i = both; i < na; ++i) a.put(bita(i) != 0);
  for (size_t i = both; i < nb; ++i) b.put(bitb(i) != 0);
  a.finish(outa);
  b.finish(outb);
}

// A stream of n zero bits (the seam flags of an attribute without interior seams, edgebreaker.rs:610-653): the coder's
// state walks the same cycle over and over — it grows by f0-th parts until the renormalisation threshold, emits one byte
// and drops back to one of at most 16 states x >> 8 — so the cycles are tabulated per start state (steps until the next
// renormalisation, byte emitted, state after it) by running the coder's own update rule once, and the run is then
// coded by table jumps plus a tail of single steps. Bytes are those of rabs_encode by construction.
static inline void rabs_encode_zero_run(size_t n, uint8_t zero_prob, std::vector<uint8_t>& out) {
  const uint32_t f0 = zero_prob, f1 = 256u - f0;
  if (f0 == 0) throw Error(DXO_ERR_UNSUPPORTED_INPUT, "rABS: zero frequency");
  const uint32_t thr = f0 << 12;
  auto step = [&](uint32_t x) { const uint32_t q = x / f0; return (q << 8) + (x - q * f0) + f1; };  // after the renormalisation check
  out.clear();
  uint32_t x = 4096u;
  size_t left = n;
  struct Cycle { uint32_t steps = 0, next = 0; uint8_t byte = 0; };
  // states right after a renormalisation lie in [thr >> 8, (2^20 - 1) >> 8]
  const uint32_t lo = thr >> 8, hi = ((1u << 20) - 1u) >> 8;
  std::vector<Cycle> table(hi - lo + 1);
  while (left) {
    if (x >= thr) {  // renormalise, then look the cycle of the new state up (or measure it)
      out.push_back((uint8_t)x);
      x >>= 8;
      for (;;) {
        Cycle& cy = table[x - lo];
        if (cy.steps == 0) {  // from x (just renormalised): steps taken before the next renormalisation is due
          uint32_t y = x, k = 0;
          while (y < thr) { y = step(y); ++k; }
          cy.steps = k; cy.byte = (uint8_t)y; cy.next = y >> 8;
        }
        if (left <= cy.steps) break;  // the run ends inside this cycle (a renormalisation happens only when another bit follows)
        left -= cy.steps;
        out.push_back(cy.byte);
        x = cy.next;
      }
    }
    x = step(x);
    --left;
  }
  ByteSink tail;
  ans_write_tail(x - 4096u, tail);
  out.insert(out.end(), tail.data.begin(), tail.data.end());
}

// zero_prob byte + leb128 length + rABS bytes: the framing every side stream uses.
static inline void write_side_stream(const uint8_t* bits, size_t n, bool reversed, uint8_t zero_prob, ByteSink& w) {
  w.u8(zero_prob);
  ByteSink tmp;
  rabs_encode(bits, n, reversed, zero_prob, tmp);
  w.varint(tmp.size());
  w.bytes(tmp.data);
}

}  // namespace dxo
