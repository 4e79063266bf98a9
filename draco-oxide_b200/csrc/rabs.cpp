// Binary rANS (RabsCoder, encode/entropy/rans.rs:71-127) for long, heavily skewed bit streams: seam flags, flips,
// orientation deltas of smooth meshes are >= 97 % one value. The bit-by-bit coder (common.hpp) is a serial chain of one
// multiply per bit; here the runs of the common value are crossed by table look-ups and only the rare bits are coded one by
// one. Same bytes as rabs_encode by construction: every table entry is produced by the coder's own update rule.
//
// Between two rare bits the state only ever follows "common steps" x -> T(x) = ((x / f) << 8) + x % f + cum, interrupted by
// a renormalisation (emit the low byte, x >>= 8) whenever x has reached f << 12 and another common bit follows. Such a walk
// starts from one of few states: the state right after a rare bit — it is a function of the value the rare step divided,
// which lies below 4096 * f_rare — or the state right after a renormalisation (below 4096), or the initial 4096. For every
// start state the table holds its trajectory t[0] = start, t[m + 1] = T(t[m]) up to the first t[L] >= f << 12, filled the
// first time the state is met. A run of r common bits from a start state is then: r <= L ? t[r] : (t[L], emit, continue
// from the post-renormalisation state's trajectory with r - L bits). The tables depend on the zero probability alone — not
// on any input — and are kept for the life of the process (one per probability met, pages committed as they are touched).
#include <atomic>
#include <memory>
#include <mutex>

#include "common.hpp"

namespace dxo {
namespace {

struct SparseTable {
  uint32_t f_c, f_r, cum_c, cum_r, thr_c, thr_r;
  uint32_t stride = 0;  // longest trajectory + 1
  size_t slots = 0;     // [0, 4096 f_r): after a rare bit, indexed by the value its step divided; then 4096 post-renormalisation states; then the initial state
  std::unique_ptr<std::atomic<uint16_t>[]> len;          // 0xFFFF = not filled yet
  std::unique_ptr<uint32_t[]> traj;                     // slot * stride + m (plain new[]: untouched pages stay uncommitted)
  std::mutex fill;

  uint32_t T(uint32_t x) const { const uint32_t q = x / f_c; return (q << 8) + (x - q * f_c) + cum_c; }
  uint32_t T_rare(uint32_t x) const { const uint32_t q = x / f_r; return (q << 8) + (x - q * f_r) + cum_r; }

  SparseTable(uint8_t zero_prob, uint32_t rare_bit) {
    const uint32_t f0 = zero_prob, f1 = 256u - f0;
    f_c = rare_bit ? f0 : f1; f_r = rare_bit ? f1 : f0;
    cum_c = rare_bit ? f1 : 0u;  // cum of bit 0 is f1, of bit 1 is 0
    cum_r = rare_bit ? 0u : f1;
    thr_c = f_c << 12; thr_r = f_r << 12;
    uint32_t x = 4096u, L = 0;  // the lowest state any walk starts from (after its first step a post-renormalisation state is above it)
    while (x < thr_c) { x = T(x); ++L; }
    stride = L + 3;
    slots = (size_t)4096 * f_r + 4096 + 1;
    len.reset(new std::atomic<uint16_t>[slots]);
    for (size_t i = 0; i < slots; ++i) len[i].store(0xFFFF, std::memory_order_relaxed);
    traj.reset(new uint32_t[slots * stride]);
  }
  size_t rare_slot(uint32_t divided) const { return divided; }
  size_t renorm_slot(uint32_t s) const { return (size_t)4096 * f_r + s; }
  size_t initial_slot() const { return (size_t)4096 * f_r + 4096; }
  // trajectory of `slot`, whose walk starts at `start`; returns its length L (t[L] >= thr_c)
  uint32_t ensure(size_t slot, uint32_t start) {
    uint32_t L = len[slot].load(std::memory_order_acquire);
    if (L != 0xFFFF) return L;
    std::lock_guard<std::mutex> lock(fill);
    L = len[slot].load(std::memory_order_acquire);
    if (L != 0xFFFF) return L;
    uint32_t* t = traj.get() + slot * stride;
    uint32_t x = start, m = 0;
    t[0] = x;
    while (x < thr_c) {
      x = T(x);
      if (++m >= stride) throw Error(DXO_ERR_INTERNAL, "rABS trajectory longer than its bound");
      t[m] = x;
    }
    len[slot].store((uint16_t)m, std::memory_order_release);
    return m;
  }
};

std::atomic<SparseTable*> g_tables[256][2];
std::mutex g_tables_mutex;

SparseTable& table_for(uint8_t zero_prob, uint32_t rare_bit) {
  SparseTable* t = g_tables[zero_prob][rare_bit].load(std::memory_order_acquire);
  if (t) return *t;
  std::lock_guard<std::mutex> lock(g_tables_mutex);
  t = g_tables[zero_prob][rare_bit].load(std::memory_order_acquire);
  if (!t) {
    t = new SparseTable(zero_prob, rare_bit);  // kept for the life of the process
    g_tables[zero_prob][rare_bit].store(t, std::memory_order_release);
  }
  return *t;
}

}  // namespace

bool rabs_sparse_applies(size_t n, uint8_t zero_prob, size_t rare_count) {
  static const bool off = getenv("DXO_NO_SPARSE_RABS") != nullptr;
  const uint32_t f0 = zero_prob, f1 = 256u - f0, f_c = f0 > f1 ? f0 : f1;
  return !off && n >= 16384 && f_c >= 248 && f_c <= 255 && rare_count * 64 <= n;
}

void rabs_encode_sparse(size_t n, uint8_t zero_prob, uint32_t rare_bit, const uint32_t* rare_pos, size_t rare_count, std::vector<uint8_t>& out) {
  SparseTable& tb = table_for(zero_prob, rare_bit);
  out.clear();
  uint32_t x = 4096u;
  size_t slot = tb.initial_slot();  // the trajectory slot whose t[0] is x
  size_t cur = 0;
  for (size_t k = 0; k <= rare_count; ++k) {
    const size_t stop = k < rare_count ? rare_pos[k] : n;
    if (stop < cur || stop > n) throw Error(DXO_ERR_INTERNAL, "rABS: rare positions out of order");
    size_t run = stop - cur;  // common bits up to the next rare bit (or the end)
    while (run) {
      const uint32_t L = tb.ensure(slot, x);
      const uint32_t* t = tb.traj.get() + slot * tb.stride;
      if (run <= L) { x = t[run]; break; }
      x = t[L];  // >= thr_c and another common bit follows: it renormalises first
      run -= L;
      out.push_back((uint8_t)x);
      x >>= 8;
      slot = tb.renorm_slot(x);
    }
    if (k == rare_count) break;
    if (x >= tb.thr_r) { out.push_back((uint8_t)x); x >>= 8; }  // the rare bit's own step
    slot = tb.rare_slot(x);
    x = tb.T_rare(x);
    cur = stop + 1;
  }
  ByteSink tail;
  ans_write_tail(x - 4096u, tail);
  out.insert(out.end(), tail.data.begin(), tail.data.end());
}

}  // namespace dxo
