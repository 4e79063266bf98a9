// Attribute hot path kernels for sm_100a (B200). Hand-written CUDA; compiled with
// -fmad=false (no FMA contraction) and default IEEE division/sqrt so that integer
// outputs are bit-identical to the reference's Rust arithmetic.
//
// Every kernel cites the reference loop it replaces (paths relative to
// /root/reference/draco-oxide/src/). None of this work is a dense contraction, so
// tensor cores are not used; the kernels are HBM-bound gathers/streams (K1-K8), one
// small single-CTA table kernel (K9) and a latency-bound serial coder (K10).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>

#include "kernels.cuh"

#include <cstdlib>
#include <mutex>
#include <vector>

namespace dxo {
namespace gpu {

namespace {

constexpr int kThreads = 256;
constexpr int kMaxBlocks = 148 * 8;  // 148 SMs x 8 resident CTAs of 256 threads

inline int grid_for(uint64_t work_items, int per_block = kThreads) {
  uint64_t b = (work_items + per_block - 1) / per_block;
  if (b < 1) b = 1;
  if (b > (uint64_t)kMaxBlocks) b = kMaxBlocks;
  return (int)b;
}

__device__ __forceinline__ uint32_t cnext(uint32_t c) { return (c % 3u == 2u) ? c - 2u : c + 1u; }
__device__ __forceinline__ uint32_t cprev(uint32_t c) { return (c % 3u == 0u) ? c + 2u : c - 1u; }

// streaming loads (read once): bypass L1 allocation
__device__ __forceinline__ uint32_t ld_stream(const uint32_t* p) { return __ldcs(p); }
__device__ __forceinline__ uint4 ld_stream4(const uint4* p) { return __ldcs(p); }

__device__ __forceinline__ uint32_t opp_of(const TableDev& t, uint32_t c) {
  if (t.seam && t.seam[c]) return kNoneDev;
  return __ldg(t.opposite + c);
}
__device__ __forceinline__ uint32_t value_index(const QuantDev& q, uint32_t point) { return q.map ? __ldg(q.map + point) : point; }

// Quantized values are stored with a power-of-two stride so one value is one vector load:
// N = 1 -> int, N = 2 -> int2, N = 3 -> int4 (w unused), N = 4 -> int4.
template <int N> __device__ __forceinline__ void load_q(const QuantDev& q, uint32_t vi, int32_t* out) {
  if (N == 1) { out[0] = __ldg(q.values + vi); }
  else if (N == 2) { const int2 v = __ldg(reinterpret_cast<const int2*>(q.values) + vi); out[0] = v.x; out[1] = v.y; }
  else { const int4 v = __ldg(reinterpret_cast<const int4*>(q.values) + vi); out[0] = v.x; out[1] = v.y; out[2] = v.z; if (N == 4) out[3] = v.w; }
}

// The three corners of a face are one 16-byte tuple {c0, c1, c2, -}: corner c = 3 f + k reads
// its own entry and those of next(c) / prev(c) with a single LDG.128.
struct Tri { uint32_t self, next, prev; };
__device__ __forceinline__ Tri load_tri(const uint4* __restrict__ tuples, uint32_t c) {
  const uint32_t f = c / 3u, k = c - 3u * f;
  const uint4 t = __ldg(tuples + f);
  Tri r;
  r.self = k == 0 ? t.x : (k == 1 ? t.y : t.z);
  r.next = k == 0 ? t.y : (k == 1 ? t.z : t.x);
  r.prev = k == 0 ? t.z : (k == 1 ? t.x : t.y);
  return r;
}

// to_positive_i32 — utils/mod.rs:152-158 (wrapping arithmetic)
__device__ __forceinline__ uint32_t zigzag(int32_t v) {
  return v >= 0 ? ((uint32_t)v << 1) : ((((uint32_t)(-(v + 1))) << 1) + 1u);
}

// Block-wide accumulation of (nonzero count, max symbol) into the stats block.
__device__ __forceinline__ void accumulate_symbol_stats(uint32_t nonzero, uint32_t maxsym, uint32_t errs, AttrStats* stats) {
  nonzero = __reduce_add_sync(0xFFFFFFFFu, nonzero);
  maxsym = __reduce_max_sync(0xFFFFFFFFu, maxsym);
  errs = __reduce_or_sync(0xFFFFFFFFu, errs);
  __shared__ uint32_t s_nz, s_mx, s_er;
  if (threadIdx.x == 0) { s_nz = 0; s_mx = 0; s_er = 0; }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) {
    if (nonzero) atomicAdd(&s_nz, nonzero);
    if (maxsym) atomicMax(&s_mx, maxsym);
    if (errs) atomicOr(&s_er, errs);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (s_nz) atomicAdd(&stats->nonzero_symbols, s_nz);
    if (s_mx) atomicMax(&stats->max_symbol, s_mx);
    if (s_er) atomicOr(&stats->error_flags, s_er);
  }
}

}  // namespace

// ---------------------------------------------------------------------------------------
__global__ void init_stats_kernel(AttrStats* st) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    for (int k = 0; k < 4; ++k) { st->vmin_bits[k] = 0; st->vmax_bits[k] = 0; }  // +0.0: min and max start at zero
    st->range = 0.0f;
    st->wrap_min = 0x7FFFFFFF;
    st->wrap_max = (int32_t)0x80000000;
    st->nonzero_symbols = 0; st->max_symbol = 0; st->error_flags = 0;
    st->bit_length = 0; st->precision = 0; st->num_table_symbols = 0; st->table_bytes = 0; st->payload_bytes = 0;
    st->pad[0] = st->pad[1] = st->pad[2] = 0;
  }
}
void init_stats(AttrStats* stats, cudaStream_t s) { init_stats_kernel<<<1, 32, 0, s>>>(stats); }

// ---------------------------------------------------------------------------------------
// K1 — per-component min / max starting from 0 with strict comparisons
// (quantization_coordinate_wise.rs:28-46). Results are +0.0 or strictly negative
// (min) / strictly positive (max), so unsigned atomicMax on the bit pattern orders
// both correctly and never produces -0.0.
template <int N>
__device__ __forceinline__ void minmax_body(const float* __restrict__ values, AttrStats* stats, uint64_t i0, uint64_t i1, uint64_t istep) {
  float mn[N], mx[N];
#pragma unroll
  for (int k = 0; k < N; ++k) { mn[k] = 0.0f; mx[k] = 0.0f; }
  for (uint64_t i = i0; i < i1; i += istep) {
#pragma unroll
    for (int k = 0; k < N; ++k) {
      const float c = __ldcs(values + i * N + k);
      if (c < mn[k]) mn[k] = c;
      if (c > mx[k]) mx[k] = c;
    }
  }
  __shared__ uint32_t s_mn[N], s_mx[N];
  if (threadIdx.x < N) { s_mn[threadIdx.x] = 0; s_mx[threadIdx.x] = 0; }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < N; ++k) {
    const uint32_t a = __reduce_max_sync(0xFFFFFFFFu, __float_as_uint(mn[k]));
    const uint32_t b = __reduce_max_sync(0xFFFFFFFFu, __float_as_uint(mx[k]));
    if ((threadIdx.x & 31) == 0) { if (a) atomicMax(&s_mn[k], a); if (b) atomicMax(&s_mx[k], b); }
  }
  __syncthreads();
  if (threadIdx.x < N) {
    if (s_mn[threadIdx.x]) atomicMax(&stats->vmin_bits[threadIdx.x], s_mn[threadIdx.x]);
    if (s_mx[threadIdx.x]) atomicMax(&stats->vmax_bits[threadIdx.x], s_mx[threadIdx.x]);
  }
}
template <int N>
__global__ void __launch_bounds__(kThreads) minmax_kernel(const float* __restrict__ values, uint64_t num_values, AttrStats* stats) {
  minmax_body<N>(values, stats, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, num_values, (uint64_t)gridDim.x * blockDim.x);
}

void launch_minmax(const float* values, uint64_t num_values, uint32_t ncomp, AttrStats* stats, cudaStream_t s) {
  const int g = grid_for(num_values, kThreads * 4);
  switch (ncomp) {
    case 1: minmax_kernel<1><<<g, kThreads, 0, s>>>(values, num_values, stats); break;
    case 2: minmax_kernel<2><<<g, kThreads, 0, s>>>(values, num_values, stats); break;
    case 3: minmax_kernel<3><<<g, kThreads, 0, s>>>(values, num_values, stats); break;
    default: minmax_kernel<4><<<g, kThreads, 0, s>>>(values, num_values, stats); break;
  }
}

// K2 — q = (i32)(i64)(((v - min) / range) * (2^bits - 1) + 0.5), four separate f32
// roundings (quantization_coordinate_wise.rs:70-91). One thread per value; the result is
// stored with the padded stride of load_q (N = 3 -> int4).
template <int N>
__device__ __forceinline__ void quantize_body(const float* __restrict__ values, uint32_t bits, int32_t* __restrict__ out, AttrStats* stats,
                                              bool writes_range, uint64_t i0, uint64_t i1, uint64_t istep, int32_t w_init = 0,
                                              const uint8_t* __restrict__ used = nullptr, const uint32_t* __restrict__ rank_w = nullptr) {
  int32_t wmn = 0x7FFFFFFF, wmx = (int32_t)0x80000000;
  float mn[N];
  float range = 0.0f;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    mn[k] = __uint_as_float(stats->vmin_bits[k]);
    const float d = __uint_as_float(stats->vmax_bits[k]) - mn[k];
    if (d > range) range = d;
  }
  if (writes_range && threadIdx.x == 0) stats->range = range;
  const float maxq = (float)(unsigned long long)((1ull << bits) - 1ull);
  for (uint64_t i = i0; i < i1; i += istep) {
    int32_t q[4] = {0, 0, 0, N == 3 ? w_init : 0};
#pragma unroll
    for (int k = 0; k < N; ++k) {
      const float v = __ldcs(values + i * N + k);
      const float diff = v - mn[k];
      const float normalized = (range == 0.0f) ? diff : (diff / range);
      const float quantized = normalized * maxq;
      const float r = quantized + 0.5f;
      const long long wide = (long long)r;  // cvt.rzi.s64.f32: truncates, saturates, NaN -> 0 (Rust `as i64`)
      q[k] = (int32_t)wide;                 // wrapping `as i32`
    }
    if (used && __ldcs(used + i)) {
#pragma unroll
      for (int k = 0; k < N; ++k) { wmn = min(wmn, q[k]); wmx = max(wmx, q[k]); }
    }
    if (N == 3 && rank_w) q[3] = (int32_t)__ldcs(rank_w + i);
    if (N == 1) out[i] = q[0];
    else if (N == 2) reinterpret_cast<int2*>(out)[i] = make_int2(q[0], q[1]);
    else reinterpret_cast<int4*>(out)[i] = make_int4(q[0], q[1], q[2], q[3]);
  }
  if (used) {  // warp -> CTA -> one global atomic pair per CTA
    __shared__ int32_t s_wmn, s_wmx;
    if (threadIdx.x == 0) { s_wmn = 0x7FFFFFFF; s_wmx = (int32_t)0x80000000; }
    __syncthreads();
    wmn = __reduce_min_sync(0xFFFFFFFFu, wmn);
    wmx = __reduce_max_sync(0xFFFFFFFFu, wmx);
    if ((threadIdx.x & 31) == 0) { atomicMin(&s_wmn, wmn); atomicMax(&s_wmx, wmx); }
    __syncthreads();
    if (threadIdx.x == 0) { atomicMin(&stats->wrap_min, s_wmn); atomicMax(&stats->wrap_max, s_wmx); }
  }
}
template <int N>
__global__ void __launch_bounds__(kThreads) quantize_kernel(const float* __restrict__ values, uint64_t num_values, uint32_t bits,
                                                            int32_t* __restrict__ out, AttrStats* stats, int32_t w_init, const uint8_t* __restrict__ used,
                                                            const uint32_t* __restrict__ rank_w) {
  quantize_body<N>(values, bits, out, stats, blockIdx.x == 0, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, num_values, (uint64_t)gridDim.x * blockDim.x, w_init,
                   used, rank_w);
}

void launch_quantize(const float* values, uint64_t num_values, uint32_t ncomp, uint32_t bits, int32_t* out, AttrStats* stats, cudaStream_t s, int32_t w_init,
                     const uint8_t* used, const uint32_t* rank_w) {
  const int g = grid_for(num_values, kThreads * 2);
  switch (ncomp) {
    case 1: quantize_kernel<1><<<g, kThreads, 0, s>>>(values, num_values, bits, out, stats, 0, used, nullptr); break;
    case 2: quantize_kernel<2><<<g, kThreads, 0, s>>>(values, num_values, bits, out, stats, 0, used, nullptr); break;
    case 3: quantize_kernel<3><<<g, kThreads, 0, s>>>(values, num_values, bits, out, stats, w_init, used, rank_w); break;
    default: quantize_kernel<4><<<g, kThreads, 0, s>>>(values, num_values, bits, out, stats, 0, used, nullptr); break;
  }
}

// rank / used tables of a sequence (static per mesh)
__global__ void __launch_bounds__(kThreads) sequence_tables_kernel(const uint32_t* __restrict__ seq, uint32_t n, TableDev t, const uint32_t* __restrict__ map,
                                                                   uint32_t* __restrict__ rank, uint8_t* __restrict__ used) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t c = ld_stream(seq + i);
    if (rank) rank[__ldg(t.corner_vertex + c)] = i;
    if (used) { const uint32_t p = __ldg(t.corner_point + c); used[map ? __ldg(map + p) : p] = 1; }
  }
}
void launch_sequence_tables(const uint32_t* seq, uint32_t n, TableDev t, const uint32_t* map, uint32_t* rank, uint8_t* used, cudaStream_t s) {
  sequence_tables_kernel<<<grid_for(n, kThreads * 2), kThreads, 0, s>>>(seq, n, t, map, rank, used);
}
__global__ void __launch_bounds__(kThreads) wrap_minmax_kernel(const int32_t* __restrict__ values, uint64_t num_values, uint32_t ncomp, uint32_t stride_v,
                                                               const uint8_t* __restrict__ used, AttrStats* stats) {
  int32_t mn = 0x7FFFFFFF, mx = (int32_t)0x80000000;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < num_values; i += stride) {
    if (!__ldcs(used + i)) continue;
    for (uint32_t k = 0; k < ncomp; ++k) { const int32_t v = __ldcs(values + i * stride_v + k); mn = min(mn, v); mx = max(mx, v); }
  }
  mn = __reduce_min_sync(0xFFFFFFFFu, mn);
  mx = __reduce_max_sync(0xFFFFFFFFu, mx);
  if ((threadIdx.x & 31) == 0) { atomicMin(&stats->wrap_min, mn); atomicMax(&stats->wrap_max, mx); }
}
void launch_wrap_minmax(const int32_t* values, uint64_t num_values, uint32_t ncomp, uint32_t stride, const uint8_t* used, AttrStats* stats, cudaStream_t s) {
  wrap_minmax_kernel<<<grid_for(num_values, kThreads * 8), kThreads, 0, s>>>(values, num_values, ncomp, stride, used, stats);
}

// 3-wide arrays -> 16-byte tuples: faces / corner_vertex (one tuple per face), ToBits values with 3 components.
__global__ void __launch_bounds__(kThreads) pad3_kernel(const uint32_t* __restrict__ in, uint64_t n_tuples, uint4* __restrict__ out) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_tuples; i += stride)
    out[i] = make_uint4(__ldcs(in + 3 * i), __ldcs(in + 3 * i + 1), __ldcs(in + 3 * i + 2), 0u);
}
void launch_pad3(const uint32_t* in, uint64_t n_tuples, uint4* out, cudaStream_t s) {
  pad3_kernel<<<grid_for(n_tuples, kThreads * 2), kThreads, 0, s>>>(in, n_tuples, out);
}

// ---------------------------------------------------------------------------------------
// Octahedral map of a float vector (geom.rs:57-91) followed by (+1)*127, truncation and
// the border fix-ups (octahedral_quantization.rs:49-64, geom.rs:137-157).
__device__ __forceinline__ void oct_quantize_f32(float x, float y, float z, int32_t& qx, int32_t& qy) {
  const float abs_sum = (fabsf(x) + fabsf(y)) + fabsf(z);
  float u = y / abs_sum;
  float v = z / abs_sum;
  if (x < 0.0f) {
    const float uo = (u < 0.0f) ? (fabsf(v) - 1.0f) : (1.0f - fabsf(v));
    const float vo = (v < 0.0f) ? (fabsf(u) - 1.0f) : (1.0f - fabsf(u));
    u = uo; v = vo;
  }
  const float a = (u + 1.0f) * 127.0f;
  const float b = (v + 1.0f) * 127.0f;
  int32_t qu = (int32_t)a, qv = (int32_t)b;  // cvt.rzi.s32.f32: truncate, saturate, NaN -> 0 (Rust `as i32`)
  const int32_t mx = 255, half = 127;
  if ((qu == 0 && qv == 0) || (qu == mx && qv == 0) || (qu == 0 && qv == mx)) { qx = mx; qy = mx; return; }
  if (qu == 0 && qv > half) qv = half - (qv - half);
  else if (qu == mx && qv < half) qv = half + (half - qv);
  else if (qv == mx && qu < half) qu = half + (half - qu);
  else if (qv == 0 && qu > half) qu = half - (qu - half);
  qx = qu; qy = qv;
}

// K3 — octahedral normal quantization, one thread per unique normal.
__device__ __forceinline__ void oct_quantize_body(const float* __restrict__ normals, int32_t* __restrict__ out, AttrStats* stats, uint64_t i0, uint64_t i1,
                                                  uint64_t istep) {
  uint32_t err = 0;
  for (uint64_t i = i0; i < i1; i += istep) {
    const float x = __ldcs(normals + 3 * i), y = __ldcs(normals + 3 * i + 1), z = __ldcs(normals + 3 * i + 2);
    int32_t qx = 0, qy = 0;
    if (x == 0.0f && y == 0.0f && z == 0.0f) err |= kErrZeroNormal;  // reference asserts (geom.rs:45)
    else oct_quantize_f32(x, y, z, qx, qy);
    reinterpret_cast<int2*>(out)[i] = make_int2(qx, qy);
  }
  if (__any_sync(0xFFFFFFFFu, err != 0) && (threadIdx.x & 31) == 0) atomicOr(&stats->error_flags, kErrZeroNormal);
}
__global__ void __launch_bounds__(kThreads) oct_quantize_kernel(const float* __restrict__ normals, uint64_t n, int32_t* __restrict__ out, AttrStats* stats) {
  oct_quantize_body(normals, out, stats, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, n, (uint64_t)gridDim.x * blockDim.x);
}

void launch_oct_quantize(const float* normals, uint64_t num_values, int32_t* out, AttrStats* stats, cudaStream_t s) {
  oct_quantize_kernel<<<grid_for(num_values, kThreads * 2), kThreads, 0, s>>>(normals, num_values, out, stats);
}

// ---------------------------------------------------------------------------------------
// Sequence preparation. rank[v] = index of v in the sequence turns the reference's
// `vertices_up_till_now.contains(v)` (O(V) scan per element) into `rank[v] < i`
// (SURVEY Appendix C.1). Also reduces WrappedDifference's min / max over all
// components of the visited originals (wrapped_difference.rs:41-49).
__device__ __forceinline__ void seq_prepare_body(const uint32_t* __restrict__ seq, const TableDev& t, const QuantDev& q, uint32_t* __restrict__ rank,
                                                 int want_minmax, AttrStats* stats, uint32_t i0, uint32_t i1, uint32_t istep) {
  int32_t mn = 0x7FFFFFFF, mx = (int32_t)0x80000000;
  for (uint32_t i = i0; i < i1; i += istep) {
    const uint32_t c = ld_stream(seq + i);
    const uint32_t vtx = __ldg(t.corner_vertex + c);
    if (q.rank_in_w) const_cast<int32_t*>(q.values)[4 * (size_t)vtx + 3] = (int32_t)i;  // vertex == value index on this path
    else rank[vtx] = i;
    if (want_minmax) {
      const uint32_t vi = value_index(q, __ldg(t.corner_point + c));
      int32_t o[4];
      const uint32_t nq = q.num_components;
      if (nq == 1) load_q<1>(q, vi, o); else if (nq == 2) load_q<2>(q, vi, o); else if (nq == 3) load_q<3>(q, vi, o); else load_q<4>(q, vi, o);
      for (uint32_t k = 0; k < nq; ++k) { mn = min(mn, o[k]); mx = max(mx, o[k]); }
    }
  }
  if (want_minmax) {  // warp -> CTA -> one global atomic pair per CTA (thousands of warps on one address serialise in L2)
    __shared__ int32_t s_mn, s_mx;
    if (threadIdx.x == 0) { s_mn = 0x7FFFFFFF; s_mx = (int32_t)0x80000000; }
    __syncthreads();
    mn = __reduce_min_sync(0xFFFFFFFFu, mn);
    mx = __reduce_max_sync(0xFFFFFFFFu, mx);
    if ((threadIdx.x & 31) == 0) { atomicMin(&s_mn, mn); atomicMax(&s_mx, mx); }
    __syncthreads();
    if (threadIdx.x == 0) { atomicMin(&stats->wrap_min, s_mn); atomicMax(&stats->wrap_max, s_mx); }
  }
}
__global__ void __launch_bounds__(kThreads) seq_prepare_kernel(const uint32_t* __restrict__ seq, uint32_t n, TableDev t, QuantDev q,
                                                               uint32_t* __restrict__ rank, int want_minmax, AttrStats* stats) {
  seq_prepare_body(seq, t, q, rank, want_minmax, stats, blockIdx.x * blockDim.x + threadIdx.x, n, gridDim.x * blockDim.x);
}

void launch_seq_prepare(const uint32_t* seq, uint32_t n, TableDev t, QuantDev q, uint32_t* rank, bool want_minmax, AttrStats* stats, cudaStream_t s) {
  seq_prepare_kernel<<<grid_for(n, kThreads * 2), kThreads, 0, s>>>(seq, n, t, q, rank, want_minmax ? 1 : 0, stats);
}

// ---------------------------------------------------------------------------------------
// WrappedDifference::squeeze — wrapped_difference.rs:66-94
struct WrapParams { int32_t mn, mx, max_diff, max_corr, min_corr; };
__device__ __forceinline__ WrapParams wrap_params(const AttrStats* stats) {
  WrapParams w;
  w.mn = stats->wrap_min; w.mx = stats->wrap_max;
  const int32_t diff = (int32_t)((uint32_t)w.mx - (uint32_t)w.mn);
  w.max_diff = (int32_t)(1u + (uint32_t)diff);
  w.max_corr = w.max_diff / 2;
  w.min_corr = -w.max_corr;
  if ((w.max_diff & 1) == 0) w.max_corr -= 1;
  return w;
}
__device__ __forceinline__ uint32_t wrapped_symbol(int32_t orig, int32_t pred, const WrapParams& w) {
  pred = pred < w.mn ? w.mn : (pred > w.mx ? w.mx : pred);
  const int32_t val = (int32_t)((uint32_t)orig - (uint32_t)pred);
  int32_t corr = val;
  if (val > w.max_corr) corr = (int32_t)((uint32_t)val - (uint32_t)w.max_diff);
  else if (val < w.min_corr) corr = (int32_t)((uint32_t)val + (uint32_t)w.max_diff);
  return zigzag(corr);
}

// value of the vertex sequenced just before element i (left_most_corner(last_v)), or zero
template <int N>
__device__ __forceinline__ void previous_value(const uint32_t* seq, uint32_t i, const TableDev& t, const QuantDev& q, int32_t* pred) {
  if (i == 0) {
#pragma unroll
    for (int k = 0; k < N; ++k) pred[k] = 0;
    return;
  }
  const uint32_t last_v = __ldg(t.corner_vertex + __ldg(seq + i - 1));
  const uint32_t lc = __ldg(t.left_most + last_v);
  const uint32_t vi = value_index(q, __ldg(t.corner_point + lc));
  load_q<N>(q, vi, pred);
}

// K4 — MeshParallelogramPrediction::predict (mesh_parallelogram_prediction.rs:186-237)
// fused with WrappedDifference and zig-zag. One thread per sequence element.
template <int N>
__device__ __forceinline__ void predict_parallelogram_body(const uint32_t* __restrict__ seq, const TableDev& t, const QuantDev& q,
                                                           const uint32_t* __restrict__ rank, uint32_t* __restrict__ symbols, AttrStats* stats,
                                                           uint32_t i0, uint32_t i1, uint32_t istep) {
  const WrapParams w = wrap_params(stats);
  uint32_t nz = 0, mxs = 0, err = 0;
  for (uint32_t i = i0; i < i1; i += istep) {
    const uint32_t c = ld_stream(seq + i);
    int32_t pred[N];
    if (N == 3 && q.rank_in_w && t.fan_link) {
      // Fast path (no point map, seams or splits: point == vertex == value index). The kernel is bound by L1 wavefronts —
      // every scattered gather of a warp costs up to 32 of them — so the number of gathers per element is what counts:
      // the link array delivers the opposite corner together with its point, and the int4 value entries carry their
      // vertex's rank in w. Six gathers instead of ten.
      const uint2 l = __ldg(t.fan_link + c);
      const Tri pts = load_tri(t.corner_point4, c);
      const int4* vals = reinterpret_cast<const int4*>(q.values);
      const int4 a = __ldg(vals + pts.next), b = __ldg(vals + pts.prev), self = __ldg(vals + pts.self);
      const int4 d = __ldg(vals + (l.x != kNoneDev ? l.y : pts.self));
      if (l.x != kNoneDev && (uint32_t)d.w < i && (uint32_t)a.w < i && (uint32_t)b.w < i) {
        pred[0] = (int32_t)((uint32_t)a.x + (uint32_t)b.x - (uint32_t)d.x);
        pred[1] = (int32_t)((uint32_t)a.y + (uint32_t)b.y - (uint32_t)d.y);
        pred[2] = (int32_t)((uint32_t)a.z + (uint32_t)b.z - (uint32_t)d.z);
      } else if (i == 0) {
        pred[0] = pred[1] = pred[2] = 0;
      } else {  // value of the vertex sequenced just before (left_most_corner(last_v))
        const uint32_t last_v = __ldg(t.corner_vertex + __ldg(seq + i - 1));
        const int4 p = __ldg(vals + __ldg(t.corner_point + __ldg(t.left_most + last_v)));
        pred[0] = p.x; pred[1] = p.y; pred[2] = p.z;
      }
      const int32_t orig3[3] = {self.x, self.y, self.z};
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const uint32_t s = wrapped_symbol(orig3[k], pred[k], w);
        symbols[(uint64_t)i * 3 + k] = s;
        nz += s != 0; mxs = max(mxs, s); err |= (s & 0x80000000u) ? kErrNegativeSymbol : 0u;
      }
      continue;
    }
    const uint32_t o = opp_of(t, c);
    const Tri pts = load_tri(t.corner_point4, c);  // points of c, next(c), prev(c): one 128-bit load
    // The gathers are latency-bound, so nothing waits for the rank test: the three values of the parallelogram are
    // requested together with the ranks (a corner without an opposite reads its own entries instead; the result is then
    // discarded). Dependent levels per element: seq -> {opposite, tuples} -> {ranks, values of next / prev / self, the
    // opposite's vertex} -> {its rank, its value}.
    const uint32_t os = o != kNoneDev ? o : c;
    // without non-manifold splits / seams / point maps a corner's vertex is its point: skip the vertex tuples
    const Tri vts = t.vertex_is_point ? pts : load_tri(t.corner_vertex4, c);
    const uint32_t op = __ldg(t.corner_point + os);
    const uint32_t ov = t.vertex_is_point ? op : __ldg(t.corner_vertex + os);
    const uint32_t r1 = __ldg(rank + vts.next), r2 = __ldg(rank + vts.prev);
    int32_t qa[N], qb[N], qd[N], orig[N];
    load_q<N>(q, value_index(q, pts.next), qa);
    load_q<N>(q, value_index(q, pts.prev), qb);
    load_q<N>(q, value_index(q, pts.self), orig);
    const uint32_t r0 = __ldg(rank + ov);
    load_q<N>(q, value_index(q, op), qd);
    if (o != kNoneDev && r0 < i && r1 < i && r2 < i) {
#pragma unroll
      for (int k = 0; k < N; ++k) pred[k] = (int32_t)((uint32_t)qa[k] + (uint32_t)qb[k] - (uint32_t)qd[k]);
    } else {
      previous_value<N>(seq, i, t, q, pred);
    }
#pragma unroll
    for (int k = 0; k < N; ++k) {
      const uint32_t s = wrapped_symbol(orig[k], pred[k], w);
      symbols[(uint64_t)i * N + k] = s;
      nz += s != 0; mxs = max(mxs, s); err |= (s & 0x80000000u) ? kErrNegativeSymbol : 0u;
    }
  }
  accumulate_symbol_stats(nz, mxs, err, stats);
}
template <int N>
__global__ void __launch_bounds__(kThreads) predict_parallelogram_kernel(const uint32_t* __restrict__ seq, uint32_t n, TableDev t, QuantDev q,
                                                                         const uint32_t* __restrict__ rank, uint32_t* __restrict__ symbols, AttrStats* stats) {
  predict_parallelogram_body<N>(seq, t, q, rank, symbols, stats, blockIdx.x * blockDim.x + threadIdx.x, n, gridDim.x * blockDim.x);
}

// K4 from records (resident sessions): whether an element has its parallelogram — an opposite corner whose vertex and the
// two neighbours are sequenced before it — and which values it then reads depend on the connectivity alone, so
// launch_parallelogram_records resolves them once per upload: {value index of the vertex, next, prev, opposite}, or
// {vertex, the vertex sequenced just before (0xFFFFFFFF = none), 0xFFFFFFFF, -} for the fallback. The step then costs one
// 16-byte record and two to four independent value gathers per element.
constexpr uint32_t kNoRecord = 0xFFFFFFFFu;
__global__ void __launch_bounds__(kThreads) parallelogram_records_kernel(const uint32_t* __restrict__ seq, uint32_t n, TableDev t, QuantDev q,
                                                                         const uint32_t* __restrict__ rank, uint4* __restrict__ rec) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t c = __ldg(seq + i);
    const uint32_t o = opp_of(t, c);
    const Tri pts = load_tri(t.corner_point4, c);
    const Tri vts = t.vertex_is_point ? pts : load_tri(t.corner_vertex4, c);
    bool para = false;
    uint32_t op = 0;
    if (o != kNoneDev) {
      op = __ldg(t.corner_point + o);
      const uint32_t ov = t.vertex_is_point ? op : __ldg(t.corner_vertex + o);
      para = __ldg(rank + ov) < i && __ldg(rank + vts.next) < i && __ldg(rank + vts.prev) < i;
    }
    const uint32_t self = value_index(q, pts.self);
    if (para) rec[i] = make_uint4(self, value_index(q, pts.next), value_index(q, pts.prev), value_index(q, op));
    else {
      uint32_t last = kNoRecord;
      if (i > 0) last = value_index(q, __ldg(t.corner_point + __ldg(t.left_most + __ldg(t.corner_vertex + __ldg(seq + i - 1)))));  // previous_value
      rec[i] = make_uint4(self, last, kNoRecord, 0u);
    }
  }
}
void launch_parallelogram_records(const uint32_t* seq, uint32_t n, TableDev t, QuantDev q, const uint32_t* rank, uint4* records, cudaStream_t s) {
  if (n) parallelogram_records_kernel<<<grid_for(n), kThreads, 0, s>>>(seq, n, t, q, rank, records);
}
template <int N>
__global__ void __launch_bounds__(kThreads) predict_parallelogram_rec_kernel(const uint4* __restrict__ rec, uint32_t n, QuantDev q,
                                                                             uint32_t* __restrict__ symbols, AttrStats* stats) {
  const WrapParams w = wrap_params(stats);
  uint32_t nz = 0, mxs = 0, err = 0;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint4 r = ld_stream4(rec + i);
    int32_t orig[N], pred[N];
    load_q<N>(q, r.x, orig);
    if (r.z != kNoRecord) {
      int32_t qa[N], qb[N], qd[N];
      load_q<N>(q, r.y, qa); load_q<N>(q, r.z, qb); load_q<N>(q, r.w, qd);
#pragma unroll
      for (int k = 0; k < N; ++k) pred[k] = (int32_t)((uint32_t)qa[k] + (uint32_t)qb[k] - (uint32_t)qd[k]);
    } else if (r.y != kNoRecord) {
      load_q<N>(q, r.y, pred);
    } else {
#pragma unroll
      for (int k = 0; k < N; ++k) pred[k] = 0;
    }
#pragma unroll
    for (int k = 0; k < N; ++k) {
      const uint32_t s = wrapped_symbol(orig[k], pred[k], w);
      symbols[(uint64_t)i * N + k] = s;
      nz += s != 0; mxs = max(mxs, s); err |= (s & 0x80000000u) ? kErrNegativeSymbol : 0u;
    }
  }
  accumulate_symbol_stats(nz, mxs, err, stats);
}
void launch_predict_parallelogram_records(const uint4* records, uint32_t n, QuantDev q, uint32_t* symbols, AttrStats* stats, cudaStream_t s) {
  if (!n) return;
  const int g = grid_for(n);
  switch (q.num_components) {
    case 1: predict_parallelogram_rec_kernel<1><<<g, kThreads, 0, s>>>(records, n, q, symbols, stats); break;
    case 2: predict_parallelogram_rec_kernel<2><<<g, kThreads, 0, s>>>(records, n, q, symbols, stats); break;
    case 3: predict_parallelogram_rec_kernel<3><<<g, kThreads, 0, s>>>(records, n, q, symbols, stats); break;
    default: predict_parallelogram_rec_kernel<4><<<g, kThreads, 0, s>>>(records, n, q, symbols, stats); break;
  }
}

void launch_predict_parallelogram(const uint32_t* seq, uint32_t n, TableDev t, QuantDev q, const uint32_t* rank, uint32_t* symbols, AttrStats* stats, cudaStream_t s) {
  const int g = grid_for(n);
  switch (q.num_components) {
    case 1: predict_parallelogram_kernel<1><<<g, kThreads, 0, s>>>(seq, n, t, q, rank, symbols, stats); break;
    case 2: predict_parallelogram_kernel<2><<<g, kThreads, 0, s>>>(seq, n, t, q, rank, symbols, stats); break;
    case 3: predict_parallelogram_kernel<3><<<g, kThreads, 0, s>>>(seq, n, t, q, rank, symbols, stats); break;
    default: predict_parallelogram_kernel<4><<<g, kThreads, 0, s>>>(seq, n, t, q, rank, symbols, stats); break;
  }
}

// K7 — DeltaPrediction (delta_prediction.rs:56-71) + Difference (difference.rs:26-34)
template <int N>
__device__ __forceinline__ void predict_delta_body(const uint32_t* __restrict__ seq, const TableDev& t, const QuantDev& q, uint32_t* __restrict__ symbols,
                                                   AttrStats* stats, uint32_t i0, uint32_t i1, uint32_t istep) {
  uint32_t nz = 0, mxs = 0, err = 0;
  for (uint32_t i = i0; i < i1; i += istep) {
    const uint32_t c = ld_stream(seq + i);
    int32_t pred[N];
    previous_value<N>(seq, i, t, q, pred);
    const uint32_t vi = value_index(q, __ldg(t.corner_point + c));
    int32_t orig[N];
    load_q<N>(q, vi, orig);
#pragma unroll
    for (int k = 0; k < N; ++k) {
      const uint32_t s = zigzag((int32_t)((uint32_t)orig[k] - (uint32_t)pred[k]));
      symbols[(uint64_t)i * N + k] = s;
      nz += s != 0; mxs = max(mxs, s); err |= (s & 0x80000000u) ? kErrNegativeSymbol : 0u;
    }
  }
  accumulate_symbol_stats(nz, mxs, err, stats);
}
template <int N>
__global__ void __launch_bounds__(kThreads) predict_delta_kernel(const uint32_t* __restrict__ seq, uint32_t n, TableDev t, QuantDev q,
                                                                 uint32_t* __restrict__ symbols, AttrStats* stats) {
  predict_delta_body<N>(seq, t, q, symbols, stats, blockIdx.x * blockDim.x + threadIdx.x, n, gridDim.x * blockDim.x);
}

void launch_predict_delta(const uint32_t* seq, uint32_t n, TableDev t, QuantDev q, uint32_t* symbols, AttrStats* stats, cudaStream_t s) {
  const int g = grid_for(n);
  switch (q.num_components) {
    case 1: predict_delta_kernel<1><<<g, kThreads, 0, s>>>(seq, n, t, q, symbols, stats); break;
    case 2: predict_delta_kernel<2><<<g, kThreads, 0, s>>>(seq, n, t, q, symbols, stats); break;
    case 3: predict_delta_kernel<3><<<g, kThreads, 0, s>>>(seq, n, t, q, symbols, stats); break;
    default: predict_delta_kernel<4><<<g, kThreads, 0, s>>>(seq, n, t, q, symbols, stats); break;
  }
}

// ---------------------------------------------------------------------------------------
// Fan walks of K5: one 8-byte load per swing. link[c] = {opposite corner of c with seam edges masked out, point of
// that opposite corner}. Swinging from a corner of vertex v across an edge lands in the neighbouring face, whose only
// vertex not on the shared edge is the tip of that opposite corner — so its point comes with the link.
__device__ __forceinline__ uint2 fan_link(const TableDev& t, uint32_t c) {
  if (t.fan_link) return __ldg(t.fan_link + c);
  const uint32_t o = opp_of(t, c);
  return make_uint2(o, o == kNoneDev ? 0u : __ldg(t.corner_point + o));
}
__global__ void __launch_bounds__(kThreads) fan_link_kernel(const uint32_t* __restrict__ opposite, const uint8_t* __restrict__ seam,
                                                            const uint32_t* __restrict__ corner_point, uint64_t n, uint2* __restrict__ out) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += stride) {
    const uint32_t o = (seam && __ldcs(seam + c)) ? kNoneDev : __ldcs(opposite + c);
    out[c] = make_uint2(o, o == kNoneDev ? 0u : __ldg(corner_point + o));
  }
}
void launch_fan_link(const uint32_t* opposite, const uint8_t* seam, const uint32_t* corner_point, uint64_t n, uint2* out, cudaStream_t s) {
  fan_link_kernel<<<grid_for(n, kThreads * 4), kThreads, 0, s>>>(opposite, seam, corner_point, n, out);
}

// K5 — MeshNormalPrediction::predict (mesh_normal_prediction.rs:77-144) fused with
// OctahedronOrthogonalTransform (oct_orthogonal.rs:23-74).
// compute_normal_of_face (:22-44): cross product in wrapping i32, then widened.
__device__ __forceinline__ void add_cross(const int32_t* pn, const int32_t* pp, const int32_t* pc, long long* sum) {
  uint32_t dn[3], dp[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) { dn[k] = (uint32_t)pn[k] - (uint32_t)pc[k]; dp[k] = (uint32_t)pp[k] - (uint32_t)pc[k]; }
  sum[0] += (long long)(int32_t)(dn[1] * dp[2] - dn[2] * dp[1]);
  sum[1] += (long long)(int32_t)(dn[2] * dp[0] - dn[0] * dp[2]);
  sum[2] += (long long)(int32_t)(dn[0] * dp[1] - dn[1] * dp[0]);
}
__device__ __forceinline__ int32_t isign(int32_t a) { return (a > 0) - (a < 0); }
__device__ __forceinline__ int32_t iabs_wrap(int32_t a) { return a < 0 ? (int32_t)(0u - (uint32_t)a) : a; }

// Flattens the fan walk of predict_normal_body (same links, same order of faces) into a polyline of position-value indices:
// [tips met swinging right, last first] + [next(c), prev(c)] + [tips met swinging left]; every pair of neighbours spans one
// face of the fan, in the orientation add_cross expects. A closed fan ends with the vertex it started with.
__global__ void __launch_bounds__(kThreads) normal_ring_kernel(const uint32_t* __restrict__ seq, uint32_t n, TableDev t, QuantDev q, QuantDev pos,
                                                               uint4* __restrict__ ring, uint2* __restrict__ head, uint8_t* __restrict__ count) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t c = __ldg(seq + i);
    const uint32_t point_c = __ldg(t.corner_point + c);
    head[i] = make_uint2(value_index(pos, point_c), value_index(q, point_c));
    const Tri pts = load_tri(t.corner_point4, c);
    uint32_t left[8], right[8];
    uint32_t nl = 0, nr = 0;
    bool over = false, closed = false;
    left[nl++] = value_index(pos, pts.next);
    left[nl++] = value_index(pos, pts.prev);
    uint32_t cur = c, guard = t.num_corners;
    uint2 l = fan_link(t, cnext(cur));
    while (l.x != kNoneDev) {
      cur = cnext(l.x);
      if (cur == c) { closed = true; break; }
      if (nl < 8) left[nl++] = value_index(pos, l.y); else { over = true; break; }
      l = fan_link(t, cnext(cur));
      if (--guard == 0) { over = true; break; }
    }
    if (!closed && !over) {
      cur = c;
      guard = t.num_corners;
      l = fan_link(t, cprev(cur));
      while (l.x != kNoneDev) {
        cur = cprev(l.x);
        if (cur == c) break;
        if (nr < 8) right[nr++] = value_index(pos, l.y); else { over = true; break; }
        l = fan_link(t, cprev(cur));
        if (--guard == 0) { over = true; break; }
      }
    }
    uint32_t out[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const uint32_t total = nl + nr;
    if (over || total > 8) count[i] = 0xFF;
    else {
      for (uint32_t k = 0; k < nr; ++k) out[k] = right[nr - 1 - k];
      for (uint32_t k = 0; k < nl; ++k) out[nr + k] = left[k];
      count[i] = (uint8_t)total;
    }
    ring[2 * (size_t)i] = make_uint4(out[0], out[1], out[2], out[3]);
    ring[2 * (size_t)i + 1] = make_uint4(out[4], out[5], out[6], out[7]);
  }
}
void launch_normal_rings(const uint32_t* seq, uint32_t n, TableDev t, QuantDev q, QuantDev pos, uint4* ring, uint2* ring_head, uint8_t* ring_count,
                         cudaStream_t s) {
  if (n) normal_ring_kernel<<<grid_for(n), kThreads, 0, s>>>(seq, n, t, q, pos, ring, ring_head, ring_count);
}

template <bool RING = false>
__device__ __forceinline__ void predict_normal_body(const uint32_t* __restrict__ seq, const TableDev& t, const QuantDev& q, const QuantDev& pos,
                                                    uint32_t* __restrict__ symbols, uint8_t* __restrict__ flips, AttrStats* stats, uint32_t i0, uint32_t i1,
                                                    uint32_t istep) {
  uint32_t nz = 0, mxs = 0, err = 0;
  for (uint32_t i = i0; i < i1; i += istep) {
    long long sum[3] = {0, 0, 0};
    uint32_t vi;
    const uint32_t ring_n = RING ? (uint32_t)__ldg(t.ring_count + i) : 0xFFu;
    if (ring_n != 0xFFu) {
      // Flattened fan (launch_normal_rings): all position gathers are independent of each other.
      const uint2 h = __ldg(t.ring_head + i);
      const uint4 ra = __ldg(t.ring + 2 * (size_t)i), rb = __ldg(t.ring + 2 * (size_t)i + 1);
      const uint32_t idx[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
      int32_t pc[3], p[8][3];
      load_q<3>(pos, h.x, pc);
#pragma unroll
      for (int k = 0; k < 8; ++k) if ((uint32_t)k < ring_n) load_q<3>(pos, idx[k], p[k]);
#pragma unroll
      for (int k = 0; k + 1 < 8; ++k) if ((uint32_t)(k + 1) < ring_n) add_cross(p[k], p[k + 1], pc, sum);
      vi = h.y;
    } else {
    const uint32_t c = ld_stream(seq + i);
    int32_t pc[3];
    const uint32_t point_c = __ldg(t.corner_point + c);
    load_q<3>(pos, value_index(pos, point_c), pc);
    // The reference swings left to the start of the fan (:86-92) and then right, summing the face
    // normals (:94-101). The sum is a wrapping i64 sum (order-independent), so one pass suffices:
    // accumulate while swinging left; only an open fan needs the walk to the right of c as well.
    // Neighbouring faces of the fan share an edge, i.e. one of the two outer vertices: its position is carried
    // over (equal vertex => equal position value), so a swing costs one link and one position load.
    const Tri pts = load_tri(t.corner_point4, c);
    int32_t first_next[3], shared[3], tip[3];
    load_q<3>(pos, value_index(pos, pts.next), first_next);
    load_q<3>(pos, value_index(pos, pts.prev), shared);
    add_cross(first_next, shared, pc, sum);  // face of c: (next - c) x (prev - c)
    uint32_t cur = c;
    uint32_t guard = t.num_corners;
    bool closed = false;
    // swing left: across the edge (c, prev); the new face has next = shared vertex, prev = the link's tip.
    // The link of the following swing is requested before the tip's position is consumed (two loads in flight).
    uint2 l = fan_link(t, cnext(cur));
    while (l.x != kNoneDev) {
      cur = cnext(l.x);
      if (cur == c) { closed = true; break; }
      const uint2 l_next = fan_link(t, cnext(cur));
      load_q<3>(pos, value_index(pos, l.y), tip);
      add_cross(shared, tip, pc, sum);
#pragma unroll
      for (int k = 0; k < 3; ++k) shared[k] = tip[k];
      l = l_next;
      if (--guard == 0) { err |= kErrFanWalk; break; }
    }
    if (!closed) {
      cur = c;
      guard = t.num_corners;
#pragma unroll
      for (int k = 0; k < 3; ++k) shared[k] = first_next[k];
      // swing right: across the edge (c, next); the new face has prev = shared vertex, next = the link's tip
      l = fan_link(t, cprev(cur));
      while (l.x != kNoneDev) {
        cur = cprev(l.x);
        if (cur == c) break;  // cannot happen for an open fan; kept as a guard
        const uint2 l_next = fan_link(t, cprev(cur));
        load_q<3>(pos, value_index(pos, l.y), tip);
        add_cross(tip, shared, pc, sum);
#pragma unroll
        for (int k = 0; k < 3; ++k) shared[k] = tip[k];
        l = l_next;
        if (--guard == 0) { err |= kErrFanWalk; break; }
      }
    }
    vi = value_index(q, point_c);
    }
    const long long upper = 1ll << 29;
    const long long abs_sum = llabs(sum[0]) + llabs(sum[1]) + llabs(sum[2]);
    if (abs_sum > upper) {
      const long long quot = abs_sum / upper;
      sum[0] /= quot; sum[1] /= quot; sum[2] /= quot;
    }
    const int32_t nx = (int32_t)sum[0], ny = (int32_t)sum[1], nzc = (int32_t)sum[2];
    int32_t p0 = 0, p1 = 0;
    if (!(nx == 0 && ny == 0 && nzc == 0)) {
      // integer vector -> f32 through f64 (geom.rs:47-52); the normalize() result is discarded there
      oct_quantize_f32((float)(double)nx, (float)(double)ny, (float)(double)nzc, p0, p1);
    }
    const int2 actual = __ldg(reinterpret_cast<const int2*>(q.values) + vi);
    // choose the sign closer to the actual value, in wrapping i32 (:128-143)
    const uint32_t d10 = (uint32_t)p0 - (uint32_t)actual.x, d11 = (uint32_t)p1 - (uint32_t)actual.y;
    const uint32_t d20 = (uint32_t)(0u - (uint32_t)p0) - (uint32_t)actual.x, d21 = (uint32_t)(0u - (uint32_t)p1) - (uint32_t)actual.y;
    const int32_t dot1 = (int32_t)(d10 * d10 + d11 * d11), dot2 = (int32_t)(d20 * d20 + d21 * d21);
    const bool flip = dot1 > dot2;
    if (flip) { p0 = (int32_t)(0u - (uint32_t)p0); p1 = (int32_t)(0u - (uint32_t)p1); }
    flips[i] = flip ? 1 : 0;

    // OctahedronOrthogonalTransform::map_with_tentative_metadata
    const int32_t one = 127;
    int32_t o0 = actual.x - one, o1 = actual.y - one;
    p0 -= one; p1 -= one;
    if (iabs_wrap(p0) + iabs_wrap(p1) > one) {
      const int32_t pp0 = p0, qs = -isign(p0 * p1);
      p0 = qs * p1 + isign(p0) * one;
      p1 = qs * pp0 + isign(p1) * one;
      const int32_t oo0 = o0, qs2 = -isign(o0 * o1);
      o0 = qs2 * o1 + isign(o0) * one;
      o1 = qs2 * oo0 + isign(o1) * one;
    }
    if (!(p0 == 0 && p1 == 0)) {
#pragma unroll 1
      for (int it = 0; it < 4 && (p0 >= 0 || p1 > 0); ++it) {  // at most 3 quarter turns are ever needed
        int32_t tmp = p0; p0 = -p1; p1 = tmp;
        tmp = o0; o0 = -o1; o1 = tmp;
      }
    }
    int32_t c0 = o0 - p0, c1 = o1 - p1;
    if (c0 < 0) c0 += 255;
    if (c1 < 0) c1 += 255;
    reinterpret_cast<uint2*>(symbols)[i] = make_uint2((uint32_t)c0, (uint32_t)c1);
    nz += (c0 != 0) + (c1 != 0);
    mxs = max(mxs, max((uint32_t)c0, (uint32_t)c1));
    err |= ((c0 | c1) < 0) ? kErrNegativeSymbol : 0u;
  }
  accumulate_symbol_stats(nz, mxs, err, stats);
}
__global__ void __launch_bounds__(kThreads, 8) predict_normal_kernel(const uint32_t* __restrict__ seq, uint32_t n, TableDev t, QuantDev q, QuantDev pos,
                                                                  uint32_t* __restrict__ symbols, uint8_t* __restrict__ flips, AttrStats* stats) {
  predict_normal_body<false>(seq, t, q, pos, symbols, flips, stats, blockIdx.x * blockDim.x + threadIdx.x, n, gridDim.x * blockDim.x);
}
// The flattened-fan form keeps eight positions in registers while their loads are in flight: half the occupancy of the
// walking form, which needs every warp it can get to hide its chain of dependent loads.
__global__ void __launch_bounds__(kThreads, 4) predict_normal_ring_kernel(const uint32_t* __restrict__ seq, uint32_t n, TableDev t, QuantDev q, QuantDev pos,
                                                                       uint32_t* __restrict__ symbols, uint8_t* __restrict__ flips, AttrStats* stats) {
  predict_normal_body<true>(seq, t, q, pos, symbols, flips, stats, blockIdx.x * blockDim.x + threadIdx.x, n, gridDim.x * blockDim.x);
}

void launch_predict_normal(const uint32_t* seq, uint32_t n, TableDev t, QuantDev q, QuantDev pos, uint32_t* symbols, uint8_t* flips, AttrStats* stats, cudaStream_t s) {
  if (t.ring) predict_normal_ring_kernel<<<grid_for(n), kThreads, 0, s>>>(seq, n, t, q, pos, symbols, flips, stats);
  else predict_normal_kernel<<<grid_for(n), kThreads, 0, s>>>(seq, n, t, q, pos, symbols, flips, stats);
}

// ---------------------------------------------------------------------------------------
// K6 — MeshPredictionForTextureCoordinates::predict
// (mesh_prediction_for_texture_coordinates.rs:107-219) fused with WrappedDifference.
// All arithmetic is wrapping i64 / u64 as in release-mode Rust.
__device__ __forceinline__ unsigned long long isqrt_ref(unsigned long long value) {  // int_sqrt :33-49
  if (value == 0) return 0;
  unsigned long long act = value, s = 1;
  while (act >= 2) { s *= 2; act /= 4; }
  s = (s + value / s) / 2;
  while (s * s > value) s = (s + value / s) / 2;
  return s;
}
// Same value as isqrt_ref, faster: for value < 2^62 the reference iteration cannot overflow and converges to
// floor(sqrt(value)) (integer Newton from above), which is computed here from the f64 root with an exact fix-up.
__device__ __forceinline__ unsigned long long isqrt_fast(unsigned long long value) {
  if (value >= (1ull << 62)) return isqrt_ref(value);
  unsigned long long r = (unsigned long long)__dsqrt_rn((double)value);
  while (r * r > value) --r;
  while ((r + 1) * (r + 1) <= value) ++r;
  return r;
}
// a > i64::MAX / d for d > 0, without the division: a * d > i64::MAX in 128 bits (a <= 0 never exceeds)
__device__ __forceinline__ bool exceeds_max_over(long long a, long long d) {
  if (a <= 0) return false;
  const unsigned long long hi = __umul64hi((unsigned long long)a, (unsigned long long)d), lo = (unsigned long long)a * (unsigned long long)d;
  return hi != 0 || lo > 0x7FFFFFFFFFFFFFFFull;
}
// n / d (truncating) with a shared f64 reciprocal; exact: for |n| < 2^52 and d < 2^52 the estimate is
// within one of the quotient and is corrected from the remainder; anything larger takes the native division.
struct DivBy {
  long long d; double rd; bool small;
  __device__ __forceinline__ explicit DivBy(long long d_) : d(d_), rd(1.0 / (double)d_), small(d_ > 0 && d_ < (1ll << 52)) {}
  __device__ __forceinline__ long long operator()(long long n) const {
    const unsigned long long an = n < 0 ? 0ull - (unsigned long long)n : (unsigned long long)n;
    if (!small || an >= (1ull << 52)) return n / d;
    long long q = (long long)((double)(long long)an * rd);
    long long r = (long long)an - q * d;
    if (r < 0) { --q; r += d; }
    else if (r >= d) { ++q; }
    return n < 0 ? -q : q;
  }
};
__device__ __forceinline__ long long labs64(long long a) { return a < 0 ? (long long)(0ull - (unsigned long long)a) : a; }
__device__ __forceinline__ long long mul64w(long long a, long long b) { return (long long)((unsigned long long)a * (unsigned long long)b); }
__device__ __forceinline__ long long add64w(long long a, long long b) { return (long long)((unsigned long long)a + (unsigned long long)b); }
__device__ __forceinline__ long long sub64w(long long a, long long b) { return (long long)((unsigned long long)a - (unsigned long long)b); }

// fallback_predict (:52-82): next vertex's value if it is already sequenced, else the last sequenced value
__device__ __forceinline__ void texcoord_fallback(const uint32_t* seq, uint32_t i, bool next_seen, uint32_t next_pt, const TableDev& t, const QuantDev& q,
                                                  int32_t* pred) {
  if (next_seen) {
    const int2 v = __ldg(reinterpret_cast<const int2*>(q.values) + value_index(q, next_pt));
    pred[0] = v.x; pred[1] = v.y;
    return;
  }
  previous_value<2>(seq, i, t, q, pred);
}

// The predictor proper (:118-219) for inputs below 2^13 (positions) / 2^12 (texture coordinates), all non-negative.
// Bounds, with P = 8191, T = 4095: |pn_k|, |cn_k| <= P; pn2, |cn.pn| <= 3 P^2 < 2^28 (i32); |x_uv| < 2^41;
// |pn_k (cn.pn)| < 2^41 and its quotient by pn2 is at most |cn| (Cauchy-Schwarz), so |cx_k| < 2^15 and cx2 < 2^32;
// cx2 pn2 < 2^59 (isqrt_fast is exact below 2^62); nrm < 2^30; |cx_uv| < 2^42; every dividend is below 2^52, where the
// reciprocal division is exact; |a|, |b| < 2^29, so the squared errors stay below 2^60. Nothing wraps, so the wrapping
// i64 arithmetic of the general path and this one agree; the guards (:139-160) compare against i64::MAX and cannot fire.
// Returns the orientation flag (1 / 2), or 0 when the reference takes its fallback (degenerate edge).
__device__ __forceinline__ uint8_t texcoord_small(const int32_t* c3, const int32_t* n3, const int32_t* p3, int2 nu, int2 pu, int2 cur, int32_t* pred) {
  const int32_t pn[3] = {p3[0] - n3[0], p3[1] - n3[1], p3[2] - n3[2]};
  const int32_t pn2 = pn[0] * pn[0] + pn[1] * pn[1] + pn[2] * pn[2];
  if (pn2 == 0) return 0;
  const int32_t cn[3] = {c3[0] - n3[0], c3[1] - n3[1], c3[2] - n3[2]};
  const int32_t cn_dot_pn = pn[0] * cn[0] + pn[1] * cn[1] + pn[2] * cn[2];
  const int32_t pn_uv[2] = {pu.x - nu.x, pu.y - nu.y};  // not both zero: the caller has dealt with nu == pu
  const long long d = pn2;
  const double rd = 1.0 / (double)pn2;
  auto div = [&](long long n) -> long long {  // truncating n / d, |n| < 2^52
    const long long an = n < 0 ? -n : n;
    long long q = (long long)((double)an * rd);
    const long long r = an - q * d;
    if (r < 0) --q; else if (r >= d) ++q;
    return n < 0 ? -q : q;
  };
  const long long x_uv[2] = {(long long)nu.x * pn2 + (long long)pn_uv[0] * cn_dot_pn, (long long)nu.y * pn2 + (long long)pn_uv[1] * cn_dot_pn};
  int32_t cx[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) cx[k] = c3[k] - (n3[k] + (int32_t)div((long long)pn[k] * cn_dot_pn));
  const unsigned long long cx2 = (unsigned long long)((long long)cx[0] * cx[0] + (long long)cx[1] * cx[1] + (long long)cx[2] * cx[2]);
  const int32_t nrm = (int32_t)isqrt_fast(cx2 * (unsigned long long)pn2);
  const long long cx_uv[2] = {(long long)pn_uv[1] * nrm, (long long)(-pn_uv[0]) * nrm};
  const int32_t a0 = (int32_t)div(x_uv[0] + cx_uv[0]), a1 = (int32_t)div(x_uv[1] + cx_uv[1]);
  const int32_t b0 = (int32_t)div(x_uv[0] - cx_uv[0]), b1 = (int32_t)div(x_uv[1] - cx_uv[1]);
  const int32_t ea0 = cur.x - a0, ea1 = cur.y - a1, eb0 = cur.x - b0, eb1 = cur.y - b1;
  const long long da = (long long)ea0 * ea0 + (long long)ea1 * ea1, db = (long long)eb0 * eb0 + (long long)eb1 * eb1;
  if (da < db) { pred[0] = a0; pred[1] = a1; return 2; }
  pred[0] = b0; pred[1] = b1;
  return 1;
}

// The prediction for an element whose next and previous vertices are sequenced (:118-219), from loaded operands.
// Returns -1 when the reference falls back, 0 when the prediction carries no orientation bit (equal neighbours), 1 / 2 otherwise.
__device__ __forceinline__ int texcoord_math(int2 cur, int2 nu, int2 pu, const int32_t* c3, const int32_t* n3, const int32_t* p3, int32_t* pred) {
  if (nu.x == pu.x && nu.y == pu.y) { pred[0] = pu.x; pred[1] = pu.y; return 0; }
  // Quantised inputs are small non-negative numbers (13 / 12 bits cover the usual settings): then no product of the
  // reference's i64 arithmetic can wrap and none of its overflow guards can fire, and most of it fits 32 x 32 -> 64
  // bit multiplies (texcoord_small). Checked per element on the values themselves, not assumed from the settings.
  const uint32_t pos_bits = (uint32_t)(c3[0] | c3[1] | c3[2] | n3[0] | n3[1] | n3[2] | p3[0] | p3[1] | p3[2]);
  const uint32_t uv_bits = (uint32_t)(nu.x | nu.y | pu.x | pu.y | cur.x | cur.y);
  if (pos_bits < 8192u && uv_bits < 4096u) {
    const uint8_t f = texcoord_small(c3, n3, p3, nu, pu, cur, pred);
    return f ? (int)f : -1;
  }
  const long long I64MAX = 0x7FFFFFFFFFFFFFFFll;
  uint8_t oflag = 0;
  bool done = false;
  const long long cp[3] = {c3[0], c3[1], c3[2]}, np[3] = {n3[0], n3[1], n3[2]}, pp[3] = {p3[0], p3[1], p3[2]};
  const long long pn[3] = {sub64w(pp[0], np[0]), sub64w(pp[1], np[1]), sub64w(pp[2], np[2])};
  const unsigned long long pn2 = (unsigned long long)add64w(add64w(mul64w(pn[0], pn[0]), mul64w(pn[1], pn[1])), mul64w(pn[2], pn[2]));
  if (pn2 != 0) {
    const long long cn[3] = {sub64w(cp[0], np[0]), sub64w(cp[1], np[1]), sub64w(cp[2], np[2])};
    const long long cn_dot_pn = add64w(add64w(mul64w(pn[0], cn[0]), mul64w(pn[1], cn[1])), mul64w(pn[2], cn[2]));
    const long long pn_uv[2] = {(long long)pu.x - (long long)nu.x, (long long)pu.y - (long long)nu.y};
    const long long n_uv_absmax = max(labs64((long long)nu.x), labs64((long long)nu.y));
    const long long pn_uv_absmax = max(labs64(pn_uv[0]), labs64(pn_uv[1]));
    const long long pn_absmax = max(max(labs64(pn[0]), labs64(pn[1])), labs64(pn[2]));
    // overflow guards (:139-160); a zero / -1 divisor would panic in the reference and cannot occur for
    // quantized inputs (pn2 > 0 as i64 for < 2^31-bit coordinates, pn_uv != 0, pn != 0)
    bool fallback = false;
    if ((long long)pn2 <= 0) fallback = (long long)pn2 == 0 || n_uv_absmax > I64MAX / (long long)pn2;  // unreachable for quantized input
    else if (exceeds_max_over(n_uv_absmax, (long long)pn2)) fallback = true;
    if (fallback) {}
    else if (pn_uv_absmax <= 0) fallback = pn_uv_absmax == 0 || labs64(cn_dot_pn) > I64MAX / pn_uv_absmax;
    else if (exceeds_max_over(labs64(cn_dot_pn), pn_uv_absmax)) fallback = true;
    if (fallback) {}
    else if (pn_absmax <= 0) fallback = pn_absmax == 0 || labs64(cn_dot_pn) > I64MAX / pn_absmax;
    else if (exceeds_max_over(labs64(cn_dot_pn), pn_absmax)) fallback = true;
    if (!fallback) {
      const DivBy div((long long)pn2);
      const long long d = (long long)pn2;
      const long long x_uv[2] = {add64w(mul64w((long long)nu.x, d), mul64w(pn_uv[0], cn_dot_pn)),
                                 add64w(mul64w((long long)nu.y, d), mul64w(pn_uv[1], cn_dot_pn))};
      long long cx[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const long long x_pos = add64w(np[k], div(mul64w(pn[k], cn_dot_pn)));
        cx[k] = sub64w(cp[k], x_pos);
      }
      const unsigned long long cx2 = (unsigned long long)add64w(add64w(mul64w(cx[0], cx[0]), mul64w(cx[1], cx[1])), mul64w(cx[2], cx[2]));
      const long long nrm = (long long)isqrt_fast(cx2 * pn2);
      const long long cx_uv[2] = {mul64w(pn_uv[1], nrm), mul64w((long long)(0ull - (unsigned long long)pn_uv[0]), nrm)};
      const long long a0 = div(add64w(x_uv[0], cx_uv[0])), a1 = div(add64w(x_uv[1], cx_uv[1]));
      const long long b0 = div(sub64w(x_uv[0], cx_uv[0])), b1 = div(sub64w(x_uv[1], cx_uv[1]));
      const long long ea0 = sub64w((long long)cur.x, a0), ea1 = sub64w((long long)cur.y, a1);
      const long long eb0 = sub64w((long long)cur.x, b0), eb1 = sub64w((long long)cur.y, b1);
      const long long da = add64w(mul64w(ea0, ea0), mul64w(ea1, ea1));
      const long long db = add64w(mul64w(eb0, eb0), mul64w(eb1, eb1));
      if (da < db) { oflag = 2; pred[0] = (int32_t)a0; pred[1] = (int32_t)a1; }
      else { oflag = 1; pred[0] = (int32_t)b0; pred[1] = (int32_t)b1; }
      done = true;
    }
  }
  return done ? (int)oflag : -1;
}

// K6 with everything that depends on the connectivity alone resolved beforehand (launch_texcoord_records): one 32-byte record
// per sequence element instead of the chain sequence -> corner -> point / vertex tuples -> ranks -> value indices.
__global__ void __launch_bounds__(kThreads) texcoord_records_kernel(const uint32_t* __restrict__ seq, uint32_t n, TableDev t, QuantDev q, QuantDev pos,
                                                                    uint32_t pos_num_points, const uint32_t* __restrict__ rank, uint4* __restrict__ rec) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t c = __ldg(seq + i);
    const Tri pts = load_tri(t.corner_point4, c), vts = load_tri(t.corner_vertex4, c);
    uint32_t last = 0xFFFFFFFFu;
    if (i > 0) {  // previous_value
      const uint32_t last_v = __ldg(t.corner_vertex + __ldg(seq + i - 1));
      last = value_index(q, __ldg(t.corner_point + __ldg(t.left_most + last_v)));
    }
    const uint32_t flags = (__ldg(rank + vts.next) < i ? 1u : 0u) | (__ldg(rank + vts.prev) < i ? 2u : 0u);
    auto pv = [&](uint32_t pt) { return pt < pos_num_points ? value_index(pos, pt) : 0xFFFFFFFFu; };
    rec[2 * (size_t)i] = make_uint4(value_index(q, pts.self), value_index(q, pts.next), value_index(q, pts.prev), last);
    rec[2 * (size_t)i + 1] = make_uint4(pv(pts.self), pv(pts.next), pv(pts.prev), flags);
  }
}
void launch_texcoord_records(const uint32_t* seq, uint32_t n, TableDev t, QuantDev q, QuantDev pos, uint32_t pos_num_points, const uint32_t* rank,
                             uint4* records, cudaStream_t s) {
  if (n) texcoord_records_kernel<<<grid_for(n), kThreads, 0, s>>>(seq, n, t, q, pos, pos_num_points, rank, records);
}

__global__ void __launch_bounds__(kThreads, 4) predict_texcoord_rec_kernel(const uint4* __restrict__ rec, uint32_t n, QuantDev q, QuantDev pos,
                                                                        uint32_t* __restrict__ symbols, uint8_t* __restrict__ orient, AttrStats* stats) {
  const WrapParams w = wrap_params(stats);
  uint32_t nz = 0, mxs = 0, err = 0;
  const uint32_t stride = gridDim.x * blockDim.x;
  const int2* const uv = reinterpret_cast<const int2*>(q.values);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint4 ra = ld_stream4(rec + 2 * (size_t)i), rb = ld_stream4(rec + 2 * (size_t)i + 1);
    const bool next_seen = rb.w & 1u, both = (rb.w & 3u) == 3u;
    // every gather is independent of the others: issued together
    const int2 cur = __ldg(uv + ra.x);
    int2 nu = make_int2(0, 0), pu = make_int2(0, 0), lu = make_int2(0, 0);
    int32_t c3[3] = {0, 0, 0}, n3[3] = {0, 0, 0}, p3[3] = {0, 0, 0};
    if (next_seen) nu = __ldg(uv + ra.y);
    else if (ra.w != 0xFFFFFFFFu) lu = __ldg(uv + ra.w);
    if (both) {
      pu = __ldg(uv + ra.z);
      if (rb.x != 0xFFFFFFFFu) load_q<3>(pos, rb.x, c3);
      if (rb.y != 0xFFFFFFFFu) load_q<3>(pos, rb.y, n3);
      if (rb.z != 0xFFFFFFFFu) load_q<3>(pos, rb.z, p3);
    }
    int32_t pred[2];
    int r = -1;
    if (both) r = texcoord_math(cur, nu, pu, c3, n3, p3, pred);
    if (r < 0) {  // fallback_predict (:52-82)
      if (next_seen) { pred[0] = nu.x; pred[1] = nu.y; }
      else { pred[0] = lu.x; pred[1] = lu.y; }
    }
    orient[i] = r > 0 ? (uint8_t)r : 0;
    const uint32_t s0 = wrapped_symbol(cur.x, pred[0], w), s1 = wrapped_symbol(cur.y, pred[1], w);
    reinterpret_cast<uint2*>(symbols)[i] = make_uint2(s0, s1);
    nz += (s0 != 0) + (s1 != 0);
    mxs = max(mxs, max(s0, s1));
    err |= ((s0 | s1) & 0x80000000u) ? kErrNegativeSymbol : 0u;
  }
  accumulate_symbol_stats(nz, mxs, err, stats);
}
void launch_predict_texcoord_records(const uint4* records, uint32_t n, QuantDev q, QuantDev pos, uint32_t* symbols, uint8_t* orient, AttrStats* stats,
                                     cudaStream_t s) {
  if (n) predict_texcoord_rec_kernel<<<grid_for(n), kThreads, 0, s>>>(records, n, q, pos, symbols, orient, stats);
}

__device__ __forceinline__ void predict_texcoord_body(const uint32_t* __restrict__ seq, const TableDev& t, const QuantDev& q, const QuantDev& pos,
                                                      uint32_t pos_num_points, const uint32_t* __restrict__ rank, uint32_t* __restrict__ symbols,
                                                      uint8_t* __restrict__ orient, AttrStats* stats, uint32_t i0, uint32_t i1, uint32_t istep) {
  const WrapParams w = wrap_params(stats);
  uint32_t nz = 0, mxs = 0, err = 0;
  for (uint32_t i = i0; i < i1; i += istep) {
    const uint32_t c = ld_stream(seq + i);
    const Tri pts = load_tri(t.corner_point4, c), vts = load_tri(t.corner_vertex4, c);
    const uint32_t next_pt = pts.next, prev_pt = pts.prev, curr_pt = pts.self;
    const int2 cur = __ldg(reinterpret_cast<const int2*>(q.values) + value_index(q, curr_pt));
    int32_t pred[2];
    int r = -1;
    const bool next_seen = __ldg(rank + vts.next) < i;
    if (next_seen && __ldg(rank + vts.prev) < i) {
      const int2 nu = __ldg(reinterpret_cast<const int2*>(q.values) + value_index(q, next_pt));
      const int2 pu = __ldg(reinterpret_cast<const int2*>(q.values) + value_index(q, prev_pt));
      int32_t c3[3] = {0, 0, 0}, n3[3] = {0, 0, 0}, p3[3] = {0, 0, 0};
      if (!(nu.x == pu.x && nu.y == pu.y)) {
        if (curr_pt < pos_num_points) load_q<3>(pos, value_index(pos, curr_pt), c3);
        if (next_pt < pos_num_points) load_q<3>(pos, value_index(pos, next_pt), n3);
        if (prev_pt < pos_num_points) load_q<3>(pos, value_index(pos, prev_pt), p3);
      }
      r = texcoord_math(cur, nu, pu, c3, n3, p3, pred);
    }
    if (r < 0) texcoord_fallback(seq, i, next_seen, next_pt, t, q, pred);
    orient[i] = r > 0 ? (uint8_t)r : 0;
    const uint32_t s0 = wrapped_symbol(cur.x, pred[0], w), s1 = wrapped_symbol(cur.y, pred[1], w);
    reinterpret_cast<uint2*>(symbols)[i] = make_uint2(s0, s1);
    nz += (s0 != 0) + (s1 != 0);
    mxs = max(mxs, max(s0, s1));
    err |= ((s0 | s1) & 0x80000000u) ? kErrNegativeSymbol : 0u;
  }
  accumulate_symbol_stats(nz, mxs, err, stats);
}
__global__ void __launch_bounds__(kThreads, 4) predict_texcoord_kernel(const uint32_t* __restrict__ seq, uint32_t n, TableDev t, QuantDev q, QuantDev pos,
                                                                    uint32_t pos_num_points, const uint32_t* __restrict__ rank,
                                                                    uint32_t* __restrict__ symbols, uint8_t* __restrict__ orient, AttrStats* stats) {
  predict_texcoord_body(seq, t, q, pos, pos_num_points, rank, symbols, orient, stats, blockIdx.x * blockDim.x + threadIdx.x, n, gridDim.x * blockDim.x);
}

void launch_predict_texcoord(const uint32_t* seq, uint32_t n, TableDev t, QuantDev q, QuantDev pos, uint32_t pos_num_points, const uint32_t* rank,
                             uint32_t* symbols, uint8_t* orient, AttrStats* stats, cudaStream_t s) {
  predict_texcoord_kernel<<<grid_for(n), kThreads, 0, s>>>(seq, n, t, q, pos, pos_num_points, rank, symbols, orient, stats);
}

// ---------------------------------------------------------------------------------------
// K8 — symbol histogram (symbol_coding.rs:149-157). Shared-memory bins when the
// alphabet fits (always for the default 11/8/10 bits), warp-aggregated through
// match_any so equal symbols inside a warp cost one atomic; global atomics otherwise.
constexpr uint32_t kSmemBins = kSmemHistBins;

__device__ __forceinline__ void histogram_smem_body(const uint32_t* __restrict__ symbols, uint32_t* __restrict__ hist, uint32_t capacity, AttrStats* stats,
                                                    uint64_t i0, uint64_t i1, uint64_t istep) {
  __shared__ uint32_t bins[kSmemBins];
  const uint32_t nb = min(stats->max_symbol + 1u, capacity);
  for (uint32_t b = threadIdx.x; b < nb; b += blockDim.x) bins[b] = 0;
  __syncthreads();
  for (uint64_t i = i0; i < i1; i += istep) {
    const uint32_t s = ld_stream(symbols + i);
    if (s < nb) atomicAdd(&bins[s], 1u);
  }
  __syncthreads();
  for (uint32_t b = threadIdx.x; b < nb; b += blockDim.x) {
    const uint32_t v = bins[b];
    if (v) atomicAdd(hist + b, v);
  }
}
__global__ void __launch_bounds__(kThreads) histogram_smem_kernel(const uint32_t* __restrict__ symbols, uint64_t n, uint32_t* __restrict__ hist,
                                                                  uint32_t capacity, AttrStats* stats) {
  histogram_smem_body(symbols, hist, capacity, stats, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, n, (uint64_t)gridDim.x * blockDim.x);
}

__device__ __forceinline__ void histogram_global_body(const uint32_t* __restrict__ symbols, uint32_t* __restrict__ hist, uint32_t capacity, AttrStats* stats,
                                                      uint64_t i0, uint64_t i1, uint64_t istep) {
  bool over = false;
  for (uint64_t i = i0; i < i1; i += istep) {
    const uint32_t s = ld_stream(symbols + i);
    if (s < capacity) atomicAdd(hist + s, 1u); else over = true;
  }
  if (over) atomicOr(&stats->error_flags, kErrAlphabet);
}
__global__ void __launch_bounds__(kThreads) histogram_global_kernel(const uint32_t* __restrict__ symbols, uint64_t n, uint32_t* __restrict__ hist,
                                                                    uint32_t capacity, AttrStats* stats) {
  histogram_global_body(symbols, hist, capacity, stats, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, n, (uint64_t)gridDim.x * blockDim.x);
}

// ---- K8 with bulk-asynchronous staging (TMA, 1-D `cp.async.bulk` completing on an mbarrier) ---------------------------
// The plain kernel above walks its symbols with one dependent global load per loop iteration: 16 round trips to HBM per
// thread are what its 14 us consist of, not bandwidth and not the shared-memory atomics. Here the copy engine brings whole
// tiles of the symbol stream into shared memory — one elected thread arms the stage's mbarrier with the byte count and
// issues the bulk copy; the next tile is in flight while the CTA bins the current one out of shared memory with 128-bit
// LDS — so no register or issue slot is spent on the loads and their latency is paid once per tile, not per symbol.
namespace tma {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }  // init visible to the async proxy
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 16-byte aligned source / destination, size a multiple of 16
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)), "l"(gmem_src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
}  // namespace tma

constexpr uint32_t kHistStageBytes = 32u << 10;  // staging area: 2 stages of 4096 symbols (plain launches) or 1 tile of 8192 (segmented)
constexpr size_t kHistTmaSmem = kSmemBins * 4 + kHistStageBytes + 64;

// Tiles first_tile, first_tile + tile_step, ... of `tile_syms` symbols each (tile_syms * 4 * stages <= kHistStageBytes); symbols is
// 16-byte aligned. Symbols beyond the last full 16 bytes of the stream are binned by plain loads.
__device__ __forceinline__ void histogram_tma_body(const uint32_t* __restrict__ symbols, uint64_t n, uint32_t* __restrict__ hist, uint32_t capacity,
                                                   AttrStats* stats, uint64_t first_tile, uint64_t tile_step, uint64_t num_tiles, uint32_t tile_syms,
                                                   uint32_t stages, bool takes_tail) {
  extern __shared__ __align__(128) uint8_t hist_smem[];
  uint32_t* bins = reinterpret_cast<uint32_t*>(hist_smem);
  uint32_t* stage_base = reinterpret_cast<uint32_t*>(hist_smem + kSmemBins * 4);
  uint64_t* full = reinterpret_cast<uint64_t*>(hist_smem + kSmemBins * 4 + kHistStageBytes);
  const uint64_t bulk_syms = n & ~3ull;  // what the bulk copies may touch
  auto tile_count = [&](uint64_t t) -> uint32_t {  // symbols of tile t that travel by bulk copy
    const uint64_t b = t * tile_syms;
    return b >= bulk_syms ? 0u : (uint32_t)min((uint64_t)tile_syms, bulk_syms - b);
  };
  auto issue = [&](uint64_t k) {  // k-th tile of this CTA into stage k % stages (thread 0)
    const uint64_t t = first_tile + k * tile_step;
    const uint32_t cnt = t < num_tiles ? tile_count(t) : 0u;
    if (!cnt) return;
    uint64_t* bar = full + (k % stages);
    tma::mbar_expect_tx(bar, cnt * 4u);
    tma::bulk_load(stage_base + (size_t)(k % stages) * tile_syms, symbols + t * tile_syms, cnt * 4u, bar);
  };
  if (threadIdx.x == 0) {
    for (uint32_t s = 0; s < stages; ++s) tma::mbar_init(full + s, 1);
    tma::fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) for (uint32_t s = 0; s < stages; ++s) issue(s);  // the first tiles travel while the bins are cleared
  const uint32_t nb = min(stats->max_symbol + 1u, capacity);
  for (uint32_t b = threadIdx.x; b < nb; b += blockDim.x) bins[b] = 0;
  __syncthreads();
  for (uint64_t k = 0;; ++k) {
    const uint64_t t = first_tile + k * tile_step;
    if (t >= num_tiles) break;
    const uint32_t cnt = tile_count(t);
    if (cnt) {
      tma::mbar_wait(full + (k % stages), (uint32_t)((k / stages) & 1));
      const uint4* src = reinterpret_cast<const uint4*>(stage_base + (size_t)(k % stages) * tile_syms);
      for (uint32_t q = threadIdx.x; q < cnt / 4; q += blockDim.x) {
        const uint4 v = src[q];
        if (v.x < nb) atomicAdd(&bins[v.x], 1u);
        if (v.y < nb) atomicAdd(&bins[v.y], 1u);
        if (v.z < nb) atomicAdd(&bins[v.z], 1u);
        if (v.w < nb) atomicAdd(&bins[v.w], 1u);
      }
    }
    __syncthreads();  // everyone is done with the stage before it is refilled
    if (threadIdx.x == 0) issue(k + stages);
  }
  if (takes_tail && threadIdx.x < (uint32_t)(n - bulk_syms)) {  // at most 3 symbols
    const uint32_t s = symbols[bulk_syms + threadIdx.x];
    if (s < nb) atomicAdd(&bins[s], 1u);
  }
  __syncthreads();
  for (uint32_t b = threadIdx.x; b < nb; b += blockDim.x) {
    const uint32_t v = bins[b];
    if (v) atomicAdd(hist + b, v);
  }
}
constexpr uint32_t kHistTileSyms = 4096;
__global__ void __launch_bounds__(kThreads) histogram_tma_kernel(const uint32_t* __restrict__ symbols, uint64_t n, uint32_t* __restrict__ hist,
                                                                 uint32_t capacity, AttrStats* stats) {
  const uint64_t num_tiles = (n + kHistTileSyms - 1) / kHistTileSyms;
  histogram_tma_body(symbols, n, hist, capacity, stats, blockIdx.x, gridDim.x, num_tiles, kHistTileSyms, 2, blockIdx.x == 0);
}
static void allow_hist_smem(const void* kernel) {
  static std::mutex mu;
  static std::vector<std::pair<const void*, int>> done;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  for (const auto& d : done) if (d.first == kernel && d.second == dev) return;
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHistTmaSmem);
  done.push_back({kernel, dev});
}
static bool hist_use_tma() { static const bool on = getenv("DXO_NO_TMA") == nullptr; return on; }

void launch_histogram(const uint32_t* symbols, uint64_t num_symbols, uint32_t* hist, uint32_t hist_capacity, AttrStats* stats, cudaStream_t s) {
  if (hist_capacity <= kSmemBins && hist_use_tma() && ((uintptr_t)symbols & 15) == 0) {
    // persistent CTAs, two per SM (64 KB of shared memory each), a few tiles each so that the flush of the bins is amortised
    const uint64_t tiles = (num_symbols + kHistTileSyms - 1) / kHistTileSyms;
    const int g = (int)std::max<uint64_t>(1, std::min<uint64_t>(tiles, 2ull * 148));
    allow_hist_smem((const void*)histogram_tma_kernel);
    histogram_tma_kernel<<<g, kThreads, kHistTmaSmem, s>>>(symbols, num_symbols, hist, hist_capacity, stats);
  } else if (hist_capacity <= kSmemBins) {
    int g = grid_for(num_symbols, kThreads * 16);
    histogram_smem_kernel<<<g, kThreads, 0, s>>>(symbols, num_symbols, hist, hist_capacity, stats);
  } else {
    histogram_global_kernel<<<grid_for(num_symbols, kThreads * 4), kThreads, 0, s>>>(symbols, num_symbols, hist, hist_capacity, stats);
  }
}

// ---------------------------------------------------------------------------------------
// K9 — RansSymbolEncoder::new (rans.rs:146-230) as a single CTA:
//   1. f64 normalisation of the histogram to 2^P with the min-1 rule,
//   2. the sum correction with the reference's stable-sort semantics (deficit -> the
//      largest entry, highest index among ties; excess -> minus one on the `err`
//      largest entries walking down the stably sorted order),
//   3. serialisation (leb128 count, 1-3 byte frequencies, zero-run tokens incl. the
//      64-wrap quirk of :202-211),
//   4. the per-symbol {freq, cumulative, reciprocal} table for K10.
// 256 threads, not 1024: a CTA that needs half an SM at once can starve for milliseconds when other sessions keep every
// SM topped up with small CTAs (their rANS kernels queue thousands of them); a small CTA always finds room.
constexpr int kTableThreads = 256;
constexpr uint32_t kTableWarps = kTableThreads / 32;

struct BlockScan {  // exclusive scans over a block of kTableThreads threads
  uint32_t* warp_tmp;  // 32 entries of shared memory
  __device__ uint32_t exclusive_sum(uint32_t v, uint32_t& total) {
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (lane >= (uint32_t)d) inc += o; }
    __syncthreads();
    if (lane == 31) warp_tmp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
      uint32_t w = lane < kTableWarps ? warp_tmp[lane] : 0u, winc = w;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, winc, d); if (lane >= (uint32_t)d) winc += o; }
      warp_tmp[lane] = winc - w;
      if (lane == 31) warp_tmp[32] = winc;
    }
    __syncthreads();
    total = warp_tmp[32];
    return warp_tmp[wid] + inc - v;
  }
  __device__ uint32_t reduce_max(uint32_t v) {
    v = __reduce_max_sync(0xFFFFFFFFu, v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) warp_tmp[threadIdx.x >> 5] = v;
    __syncthreads();
    uint32_t r = (threadIdx.x & 31) < kTableWarps ? warp_tmp[threadIdx.x & 31] : 0u;
    r = __reduce_max_sync(0xFFFFFFFFu, r);
    return r;
  }
  __device__ unsigned long long reduce_sum64(unsigned long long v, unsigned long long* tmp64) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) tmp64[threadIdx.x >> 5] = v;
    __syncthreads();
    unsigned long long r = (threadIdx.x & 31) < kTableWarps ? tmp64[threadIdx.x & 31] : 0ull;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) r += __shfl_xor_sync(0xFFFFFFFFu, r, d);
    return r;
  }
};

__device__ __forceinline__ uint32_t freq_token_bytes(uint32_t f) { return 1u + (f >= (1u << 6)) + (f >= (1u << 14)); }

__device__ __forceinline__ void build_table_body(const uint32_t* __restrict__ hist, uint32_t capacity, unsigned long long total_symbols,
                                                 uint32_t* __restrict__ work, uint4* __restrict__ rans_table,
                                                 uint8_t* __restrict__ table_bytes, uint32_t table_bytes_capacity, AttrStats* stats) {
  __shared__ uint32_t s_warp[33];
  __shared__ unsigned long long s_tmp64[32];
  __shared__ uint32_t s_scalar[4];
  BlockScan scan{s_warp};
  const uint32_t tid = threadIdx.x;

  // bit_length from the count of non-zero symbols (symbol_coding.rs:46-52,118) and the precision table (:120-140)
  const uint32_t nonzero = stats->nonzero_symbols;
  uint32_t bit_length = (nonzero == 0 ? 0u : (32u - (uint32_t)__clz(nonzero))) + 1u;
  bit_length = min(max(bit_length, 1u), 18u);
  const uint32_t P = bit_length <= 8 ? 12u : bit_length == 9 ? 13u : bit_length == 10 ? 15u : bit_length == 11 ? 16u
                   : bit_length == 12 ? 18u : bit_length == 13 ? 19u : 20u;
  const uint32_t target = 1u << P;
  const uint32_t K = stats->max_symbol + 1u;  // histogram is sized max_symbol + 1 and its last bin is non-zero
  if (K > capacity || total_symbols == 0) {
    if (tid == 0) { atomicOr(&stats->error_flags, K > capacity ? kErrAlphabet : kErrRansFreq); stats->bit_length = bit_length; stats->precision = P; }
    return;
  }
  uint32_t* dist = work;                 // normalised frequencies
  uint32_t* prev_nz = work + capacity;   // index+1 of the last non-zero entry at or before i (0 = none)
  uint32_t* next_nz = work + 2 * (size_t)capacity;  // index of the first non-zero entry after i
  const uint32_t per = (K + kTableThreads - 1) / kTableThreads;  // contiguous entries per thread
  const uint32_t lo = min(tid * per, K), hi = min(lo + per, K);

  // 1. normalisation (rans.rs:156-168): prob = freq / total (f64); (prob * 2^P + 0.5) as usize; 0 -> 1 for present symbols
  const double total_f = (double)total_symbols;
  const double scale = (double)target;
  unsigned long long local_sum = 0;
  for (uint32_t i = lo; i < hi; ++i) {
    const uint32_t f = hist[i];
    const double prob = (double)f / total_f;
    const double x = prob * scale;
    const double y = x + 0.5;
    uint32_t nf = (uint32_t)(unsigned long long)y;
    if (nf == 0 && f > 0) nf = 1;
    dist[i] = nf;
    local_sum += nf;
  }
  const unsigned long long sum = scan.reduce_sum64(local_sum, s_tmp64);
  __syncthreads();

  // 2. correction
  if (sum < target) {
    // deficit goes to sorted.last(): the largest value, highest index among ties
    uint32_t best = 0;
    for (uint32_t i = lo; i < hi; ++i) best = max(best, dist[i]);
    const uint32_t gmax = scan.reduce_max(best);
    uint32_t idx = 0;
    for (uint32_t i = lo; i < hi; ++i) if (dist[i] == gmax) idx = i + 1;
    const uint32_t gidx = scan.reduce_max(idx);
    __syncthreads();
    if (tid == 0) dist[gidx - 1] += target - (uint32_t)sum;
    __syncthreads();
  } else if (sum > target) {
    const unsigned long long err64 = sum - target;
    if (err64 > K) {  // the reference walks past index 0 and panics
      if (tid == 0) atomicOr(&stats->error_flags, kErrRansFreq);
      return;
    }
    const uint32_t err = (uint32_t)err64;
    // Find the value T such that all entries > T are decremented plus the highest-index
    // entries == T: smallest T with count(dist > T) <= err.
    uint32_t lo_t = 0, hi_t = target + K;  // count(dist > hi_t) == 0 <= err
    while (lo_t < hi_t) {
      const uint32_t mid = lo_t + (hi_t - lo_t) / 2;
      uint32_t cnt = 0;
      for (uint32_t i = lo; i < hi; ++i) cnt += dist[i] > mid;
      uint32_t tot;
      scan.exclusive_sum(cnt, tot);
      if (tot <= err) hi_t = mid; else lo_t = mid + 1;
      __syncthreads();
    }
    const uint32_t T = lo_t;
    uint32_t above = 0, equal = 0;
    for (uint32_t i = lo; i < hi; ++i) { above += dist[i] > T; equal += dist[i] == T; }
    uint32_t tot_above, tot_equal;
    scan.exclusive_sum(above, tot_above);
    const uint32_t eq_before = scan.exclusive_sum(equal, tot_equal);
    const uint32_t need_equal = err - tot_above;  // taken from the highest indices among dist == T
    if (need_equal > tot_equal || (T == 0 && need_equal > 0)) {  // would decrement a zero entry: reference underflows
      if (tid == 0) atomicOr(&stats->error_flags, kErrRansFreq);
      return;
    }
    uint32_t eq_rank = eq_before;  // rank among equal entries in index order
    for (uint32_t i = lo; i < hi; ++i) {
      const uint32_t d = dist[i];
      if (d > T) dist[i] = d - 1;
      else if (d == T) { if (tot_equal - eq_rank <= need_equal) dist[i] = d - 1; ++eq_rank; }
    }
    __syncthreads();
  }

  // 3+4. cumulative frequencies, zero-run structure, serialisation
  // prev_nz / next_nz by per-thread chunk + block-wide propagation
  uint32_t chunk_sum = 0, last_nz_local = 0, first_nz_local = kNoneDev;
  for (uint32_t i = lo; i < hi; ++i) {
    const uint32_t d = dist[i];
    chunk_sum += d;
    if (d) { last_nz_local = i + 1; if (first_nz_local == kNoneDev) first_nz_local = i; }
  }
  uint32_t tot_freq;
  uint32_t cum = scan.exclusive_sum(chunk_sum, tot_freq);
  if (tot_freq != target) { if (tid == 0) atomicOr(&stats->error_flags, kErrRansFreq); return; }
  // exclusive max-scan of last_nz over threads (values are increasing with tid when non-zero)
  __shared__ uint32_t s_last[kTableThreads];
  __shared__ uint32_t s_first[kTableThreads];
  s_last[tid] = last_nz_local;
  s_first[tid] = first_nz_local;
  __syncthreads();
  if (tid == 0) {
    uint32_t run = 0;
    for (int k = 0; k < kTableThreads; ++k) { const uint32_t v = s_last[k]; s_last[k] = run; if (v) run = v; }
    uint32_t nxt = kNoneDev;
    for (int k = kTableThreads - 1; k >= 0; --k) { const uint32_t v = s_first[k]; s_first[k] = nxt; if (v != kNoneDev) nxt = v; }
  }
  __syncthreads();
  {
    uint32_t run = s_last[tid];
    for (uint32_t i = lo; i < hi; ++i) { if (dist[i]) run = i + 1; prev_nz[i] = run; }
    uint32_t nxt = s_first[tid];
    for (uint32_t i = hi; i-- > lo;) { next_nz[i] = nxt; if (dist[i]) nxt = i; }
  }
  __syncthreads();
  // bytes emitted at each entry
  auto entry_bytes = [&](uint32_t i, uint8_t* token) -> uint32_t {
    const uint32_t d = dist[i];
    if (d) return freq_token_bytes(d);
    const uint32_t run_start = prev_nz[i];         // first index of this zero run
    const uint32_t run_end = next_nz[i];           // index of the terminating non-zero entry (exists: last bin is non-zero)
    const uint32_t R = run_end - run_start, j = i - run_start;
    const uint32_t singles = R > 64 ? R - 64 : 0;  // entries emitted as lone zeros: offset counter wrapped to 0 (:202-211)
    if (j < singles) { *token = 3; return 1; }
    if (j == singles) { *token = (uint8_t)((((R - singles) - 1u) << 2) | 3u); return 1; }
    return 0;
  };
  uint32_t nbytes = 0;
  for (uint32_t i = lo; i < hi; ++i) { uint8_t tk; nbytes += entry_bytes(i, &tk); }
  uint32_t tot_bytes;
  uint32_t off = scan.exclusive_sum(nbytes, tot_bytes);
  // leb128(num_symbols) header
  uint32_t hdr = 0;
  { uint32_t v = K; do { ++hdr; v >>= 7; } while (v); }
  if (hdr + tot_bytes > table_bytes_capacity) { if (tid == 0) atomicOr(&stats->error_flags, kErrAlphabet); return; }
  if (tid == 0) {
    uint32_t v = K, p = 0;
    do { uint8_t b = v & 0x7F; v >>= 7; table_bytes[p++] = v ? (b | 0x80) : b; } while (v);
    stats->bit_length = bit_length; stats->precision = P; stats->num_table_symbols = K; stats->table_bytes = hdr + tot_bytes;
    s_scalar[0] = hdr;
  }
  __syncthreads();
  off += s_scalar[0];
  for (uint32_t i = lo; i < hi; ++i) {
    const uint32_t d = dist[i];
    uint8_t tk = 0;
    const uint32_t nb = entry_bytes(i, &tk);
    if (d) {
      const uint32_t extra = nb - 1;
      table_bytes[off] = (uint8_t)((d << 2) | extra);
      for (uint32_t b = 0; b < extra; ++b) table_bytes[off + 1 + b] = (uint8_t)(d >> (8 * (b + 1) - 2));
      // K10 lookup: floor(x / d) for x < 2^30 as umulhi(x, M) >> lp with lp = ceil(log2 d) - 1,
      // M = ceil(2^(32+lp) / d) < 2^32 (exact, DESIGN.md "rANS division"). d = 1 uses
      // umulhi(x, 2^32 - 1) + 1 = x (x >= 1), signalled by lp = 0, M = 0xFFFFFFFF.
      uint32_t lp = 0, magic = 0xFFFFFFFFu;
      if (d >= 2) {
        lp = 31u - (uint32_t)__clz(d - 1);  // ceil(log2 d) - 1
        magic = (uint32_t)(((1ull << (32u + lp)) + d - 1) / d);
      }
      rans_table[i] = make_uint4(d, cum, magic, lp);
      cum += d;
    } else {
      if (nb) table_bytes[off] = tk;
      rans_table[i] = make_uint4(0, cum, 0, 0);
    }
    off += nb;
  }
}
__global__ void __launch_bounds__(kTableThreads) build_table_kernel(const uint32_t* __restrict__ hist, uint32_t capacity, unsigned long long total_symbols,
                                                                    uint32_t* __restrict__ work, uint4* __restrict__ rans_table,
                                                                    uint8_t* __restrict__ table_bytes, uint32_t table_bytes_capacity, AttrStats* stats) {
  build_table_body(hist, capacity, total_symbols, work, rans_table, table_bytes, table_bytes_capacity, stats);
}

void launch_build_table(const uint32_t* hist, uint32_t hist_capacity, uint64_t total_symbols, uint32_t* work, uint4* rans_table,
                        uint8_t* table_bytes, uint32_t table_bytes_capacity, AttrStats* stats, cudaStream_t s) {
  build_table_kernel<<<1, kTableThreads, 0, s>>>(hist, hist_capacity, total_symbols, work, rans_table, table_bytes, table_bytes_capacity, stats);
}

// ---------------------------------------------------------------------------------------
// K10 — RansCoder::write / flush (rans.rs:33-68), symbols fed last-to-first
// (symbol_coding.rs:161).
//
// The state recurrence x -> x' is serial, but it FORGETS: two encoders that see the same
// symbols from different states end up in exactly the same state — slowly from an arbitrary
// state, within a few hundred steps from a NEARBY one (every renormalisation shifts the low
// state bits, where small differences live, out into the byte stream). That makes an exact
// speculative-parallel coder possible (DESIGN.md "Parallel rANS"; tools/rans_merge_sim.py):
//   explore   the stream is cut into chunks of C steps. For every chunk a warp runs 32
//             trajectories (one per lane, states spread geometrically over the whole state
//             interval) through the same symbols, starting W steps before the chunk; each lane
//             records its state at the chunk start, at up to 15 evenly spaced checkpoints
//             inside the chunk and at the chunk end. No bytes.
//   chain     one warp walks the chunks carrying the TRUE state (chunk 0: l_base): the lane
//             whose recorded entering state equals it hands over its exit state; a chunk with
//             no such lane (rare) is run from the true state on the spot.
//   encode    every piece (a sixteenth of a chunk; its true entering state is the chain's state
//             or a checkpoint of the matched lane) is encoded once, one thread per piece.
//   fix-up    checks in parallel that every piece was encoded from its predecessor's exit
//             state and repairs sequentially otherwise, so the result never depends on the
//             speculation succeeding; leaves the prefix sum of the pieces' byte counts.
//   gather    the pieces' byte strings are concatenated and the 2-bit-tagged final state is
//             appended.
// The bytes are those of the sequential coder by construction: chunk 0 starts from
// l_base and every other chunk is encoded from its predecessor's true exit state.
//
// Inside a chunk, one CTA of two warps works as producer / consumer:
//   * PRODUCER warp: prefetches symbols several groups ahead (coalesced), gathers their
//     table rows, derives the three renormalisation thresholds and fills a ring of
//     32-row stages; it also turns the consumer's per-step record (x before the step,
//     byte count) into output bytes with a warp scan, 32 steps at a time.
//     (A CTA carries two producer/consumer pairs; which warp of a pair takes which role is
//     drawn when the pair starts so that the consumers of an SM spread over all four
//     schedulers, see rans_role.)
//   * CONSUMER warp: the serial chain only. Per symbol: two broadcast LDS.128,
//     q = ((umulhi(x, M) + c) >> lp) >> 8k with M = ceil(2^(32+lp)/f), lp = ceil(log2 f) - 1
//     (exact for x < 2^30, DESIGN.md "rANS division"; the multiply does not wait for k),
//     k from three independent compares and a two-level select,
//     x' = (x >> 8k) + cum + q * (2^P - f), and one STS of (x, k) for the producer.
// Stages are handed over with named barriers (bar.arrive / bar.sync), one pair per stage.
constexpr int kRansSubMax = 16;       // the encode pass works on up to 16 sub-chunks of an exploration chunk (see phase C)
constexpr int kRansStages = 3;        // 2 pairs x (3 FULL + 3 EMPTY + 1 END) named barriers = 14 of the 15 available
constexpr int kRansLookahead = 3;      // groups of symbols in flight in the producer's registers
// steps per chunk / warm-up steps of the exploration (multiples of 32). Defaults tuned on B200 (profiles/);
// DXO_RANS_CHUNK and DXO_RANS_WARMUP override them for experiments. Correctness never depends on these values;
// DXO_RANS_FAULT=1 makes the chain kernel deliberately record a wrong entering state for every fifth chunk
// (tests of the fix-up path).
struct RansPlan { uint32_t chunk, warmup; int fault; int lanes; uint32_t sub; bool fixed_chunk; };
static RansPlan rans_plan() {
  static RansPlan plan = [] {
    RansPlan p{4096, 1024, 0, 1, (uint32_t)kRansSubMax, false};
    if (const char* e = getenv("DXO_RANS_LANES")) p.lanes = atoi(e);  // phase C: 1 = one thread per chunk (default), 0 = one warp pair per chunk
    if (const char* e = getenv("DXO_RANS_FAULT")) p.fault = atoi(e);  // tests: the chain hands out wrong states, the fix-up must repair
    if (const char* e = getenv("DXO_RANS_CHUNK")) { p.chunk = (uint32_t)atoi(e); p.fixed_chunk = true; }
    if (const char* e = getenv("DXO_RANS_WARMUP")) p.warmup = (uint32_t)atoi(e);
    if (const char* e = getenv("DXO_RANS_SUB")) p.sub = (uint32_t)atoi(e);
    p.chunk = (p.chunk < 32u ? 32u : p.chunk) / 32u * 32u;
    p.warmup = p.warmup / 32 * 32;
    // sub-chunks are whole groups of 32 steps and a warp of the encode pass covers whole chunks: a power of two <= 16
    uint32_t sub = 1;
    while (sub * 2 <= p.sub && sub * 2 <= (uint32_t)kRansSubMax && p.chunk % (64u * sub) == 0) sub *= 2;
    p.sub = sub;
    return p;
  }();
  return plan;
}

// stage row: a = {thr1, thr2, thr3, cum}, b = {M, lp, g = 2^P - f, c = (f == 1)}

__device__ __forceinline__ void named_bar_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }

struct RansShared {
  uint4 rows_a[kRansStages][32];
  uint4 rows_b[kRansStages][32];
  uint32_t xk[kRansStages][32];
  uint32_t x_main[32], x_exit[32];  // per consumer lane: state at e_main / after the last step
  uint32_t x_mid[kRansSubMax - 1][32]; // ... and at the inner sub-chunk boundaries (e_main + k * sub_groups * 32 steps)
  uint32_t nbytes, err;
};

// Encodes steps [e_begin, e_end) of the stream (step e codes symbols[n - 1 - e]). Every consumer lane
// carries its own state, starting from its x_in (lanes given the same x_in stay identical); with `emit`
// the bytes of LANE 0's trajectory are produced for the steps >= e_main (e_begin..e_main is the
// warm-up). All three bounds except e_end are multiples of 32. Called by a 64-thread CTA.
// Results in sh.x_main[lane] (state at e_main), sh.x_exit[lane], sh.nbytes after the final barrier.
__device__ __forceinline__ void rans_encode_range(RansShared& sh, const uint32_t* __restrict__ symbols, unsigned long long n,
                                                  const uint4* __restrict__ table, uint32_t K, uint32_t P, unsigned long long e_begin,
                                                  unsigned long long e_main, unsigned long long e_end, uint32_t x_in, uint8_t* __restrict__ out,
                                                  int bar_base, bool is_consumer, bool emit = true, uint32_t sub_groups = 0) {
  const uint32_t lane = threadIdx.x & 31;
  const unsigned long long steps = e_end - e_begin;
  const unsigned long long ngroups = (steps + 31) / 32;
  const unsigned long long g_main = (e_main - e_begin) / 32;  // first group that produces bytes
  // barrier ids: bar_base + s = FULL[s], bar_base + kRansStages + s = EMPTY[s], bar_base + 2 * kRansStages = END
  if (is_consumer) {
    // ------------------------------- consumer: the serial chain -------------------------------
    uint32_t x = x_in;
    for (unsigned long long g = 0; g < ngroups; ++g) {
      const int s = (int)(g % kRansStages);
      const uint32_t cnt = (g + 1 == ngroups) ? (uint32_t)(steps - 32 * g) : 32u;
      if (g == g_main) sh.x_main[lane] = x;
      if (sub_groups && g > g_main && (g - g_main) % sub_groups == 0 && (g - g_main) / sub_groups < (unsigned long long)kRansSubMax)
        sh.x_mid[(g - g_main) / sub_groups - 1][lane] = x;  // state at an inner sub-chunk boundary
      named_bar_sync(bar_base + s);
      const uint32_t ra = (uint32_t)__cvta_generic_to_shared(&sh.rows_a[s][0]);
      const uint32_t rb = (uint32_t)__cvta_generic_to_shared(&sh.rows_b[s][0]);
      uint32_t* xk = sh.xk[s];
      // Rows are fetched three steps ahead with volatile shared loads so that the ~30-cycle
      // LDS latency never sits on the x -> x' chain.
      auto load_row = [&](uint32_t j, uint4& a, uint4& b) {
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "r"(ra + j * 16u));
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "r"(rb + j * 16u));
      };
      auto step = [&](uint32_t j, const uint4& a, const uint4& b) {
        const uint32_t hi = __umulhi(x, b.x) + b.w;                     // floor(x * M / 2^32) (+1 for f = 1), independent of k
        const uint32_t q0 = hi >> b.y;                                  // floor(x / f): still independent of k
        const bool p1 = x >= a.x, p2 = x >= a.y, p3 = x >= a.z;
        const uint32_t k8 = p2 ? (p3 ? 24u : 16u) : (p1 ? 8u : 0u);
        if (emit && lane == 0) xk[j] = x + (k8 << 27);                  // x < 2^30; k8/8 in bits 30..31 (only the byte writer reads it)
        x = (q0 >> k8) * b.z + ((x >> k8) + a.w);
      };
      if (cnt == 32) {
        uint4 a0, a1, a2, a3; uint4 b0, b1, b2, b3;
        load_row(0, a0, b0); load_row(1, a1, b1); load_row(2, a2, b2);
#pragma unroll
        for (uint32_t j = 0; j < 32; ++j) {
          if (j + 3 < 32) load_row(j + 3, a3, b3);
          step(j, a0, b0);
          a0 = a1; b0 = b1; a1 = a2; b1 = b2; a2 = a3; b2 = b3;
        }
      } else {
        for (uint32_t j = 0; j < cnt; ++j) { uint4 a, b; load_row(j, a, b); step(j, a, b); }
      }
      named_bar_arrive(bar_base + kRansStages + s);
    }
    sh.x_exit[lane] = x;
    if (g_main >= ngroups) sh.x_main[lane] = x;
  } else {
    // ------------------------------- producer / byte writer -----------------------------------
    uint32_t err = 0;
    uint32_t pos = 0;
    uint32_t pre[kRansLookahead];
    auto load_syms = [&](unsigned long long g) -> uint32_t {
      if (g >= ngroups) return 0;
      const unsigned long long e = e_begin + 32 * g + lane;  // this lane's step
      return e < e_end ? __ldcs(symbols + (n - 1 - e)) : 0u;
    };
#pragma unroll
    for (int d = 0; d < kRansLookahead; ++d) pre[d] = load_syms(d);
    auto emit_bytes = [&](int s, uint32_t cnt) {
      const uint32_t v = lane < cnt ? sh.xk[s][lane] : 0u;
      const uint32_t k = v >> 30, xv = v & 0x3FFFFFFFu;
      uint32_t inc = k;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (lane >= (uint32_t)d) inc += o; }
      uint8_t* dst = out + pos + (inc - k);
      if (k > 0) dst[0] = (uint8_t)xv;
      if (k > 1) dst[1] = (uint8_t)(xv >> 8);
      if (k > 2) dst[2] = (uint8_t)(xv >> 16);
      pos += __shfl_sync(0xFFFFFFFFu, inc, 31);
    };
    for (unsigned long long g = 0; g < ngroups; ++g) {
      const int s = (int)(g % kRansStages);
      if (g >= kRansStages) {  // the stage is being reused: wait until the consumer is done with it, then write its bytes
        named_bar_sync(bar_base + kRansStages + s);
        if (emit && g - kRansStages >= g_main) emit_bytes(s, 32);
      }
      const uint32_t sym = pre[0];
#pragma unroll
      for (int d = 0; d + 1 < kRansLookahead; ++d) pre[d] = pre[d + 1];
      pre[kRansLookahead - 1] = load_syms(g + kRansLookahead);
      const uint32_t cnt = (g + 1 == ngroups) ? (uint32_t)(steps - 32 * g) : 32u;
      uint4 e = make_uint4(1, 0, 0xFFFFFFFFu, 0);
      if (lane < cnt) {
        if (sym < K) e = __ldg(table + sym); else err |= kErrRansFreq;
        if (e.x == 0) { err |= kErrRansFreq; e = make_uint4(1, 0, 0xFFFFFFFFu, 0); }
      }
      const uint32_t thr1 = e.x << 10;  // f <= 2^20: fits; x < 2^30 never reaches the threshold of f = 2^20
      uint4 a;
      a.x = thr1;
      a.y = thr1 >= (1u << 24) ? 0xFFFFFFFFu : (thr1 << 8);   // x >> 8 >= thr1  <=>  x >= thr1 << 8
      a.z = thr1 >= (1u << 16) ? 0xFFFFFFFFu : (thr1 << 16);
      a.w = e.y;
      const uint4 b = make_uint4(e.z, e.w, (1u << P) - e.x, e.x == 1u ? 1u : 0u);  // {M, lp, g, c}
      sh.rows_a[s][lane] = a;
      sh.rows_b[s][lane] = b;
      named_bar_arrive(bar_base + s);
    }
    // drain: bytes of the last min(ngroups, kRansStages) groups, in order
    const unsigned long long first_pending = ngroups > (unsigned long long)kRansStages ? ngroups - kRansStages : 0;
    for (unsigned long long g = first_pending; g < ngroups; ++g) {
      const int s = (int)(g % kRansStages);
      const uint32_t cnt = (g + 1 == ngroups) ? (uint32_t)(steps - 32 * g) : 32u;
      named_bar_sync(bar_base + kRansStages + s);
      if (emit && g >= g_main) emit_bytes(s, cnt);
    }
    err = __reduce_or_sync(0xFFFFFFFFu, err);
    if (lane == 0) { sh.nbytes = pos; sh.err = err; }
  }
  named_bar_sync(bar_base + 2 * kRansStages);  // both warps of the pair: results in sh are visible
}

// role / pair of the calling warp inside a 128-thread CTA (two pairs)
// Which warp of a pair runs the serial chain is decided when the pair starts: the consumer's ~16 instructions per step keep
// the integer pipe of its scheduler busy for a good part of a step, so consumers sharing a scheduler slow each other down
// (ncu: issue rate 0.54 on the busiest scheduler against 0.04 on the idlest with fixed roles). A warp's scheduler is its
// hardware warp slot modulo 4 (tools/probe/smsp_map.cu); the pair puts its consumer on the one of its two schedulers that
// currently runs fewer consumers of any K10 kernel on this SM (live counters, updated under a per-SM lock, released when
// the consumer is done).
__device__ uint32_t g_rans_consumers[1024 * 4];
__device__ uint32_t g_rans_lock[1024];
struct RansRole { int pair; int bar_base; bool is_consumer; uint32_t counter; };
__device__ __forceinline__ RansRole rans_role(int* role_slot /* shared, one per pair */, uint32_t* sched_slot /* shared, one per warp */) {
  const int warp = threadIdx.x >> 5;
  RansRole r;
  r.pair = warp >> 1;
  r.bar_base = 1 + r.pair * (2 * kRansStages + 1);
  uint32_t smid, wid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
  smid &= 1023u;
  if ((threadIdx.x & 31) == 0) sched_slot[warp] = wid & 3u;
  named_bar_sync(r.bar_base + 2 * kRansStages);
  if ((warp & 1) == 0 && (threadIdx.x & 31) == 0) {
    uint32_t* cnt = g_rans_consumers + smid * 4u;
    const uint32_t a = sched_slot[warp], b = sched_slot[warp + 1];
    while (atomicCAS(&g_rans_lock[smid], 0u, 1u) != 0u) {}
    const int pick = atomicAdd(&cnt[a], 0u) <= atomicAdd(&cnt[b], 0u) ? 0 : 1;  // 0: the even warp is the consumer
    atomicAdd(&cnt[pick ? b : a], 1u);
    __threadfence();
    atomicExch(&g_rans_lock[smid], 0u);
    role_slot[r.pair] = pick;
  }
  named_bar_sync(r.bar_base + 2 * kRansStages);
  r.is_consumer = (warp & 1) == role_slot[r.pair];
  r.counter = smid * 4u + sched_slot[warp];
  return r;
}
__device__ __forceinline__ void rans_role_release(const RansRole& r) {
  if (r.is_consumer && (threadIdx.x & 31) == 0) atomicSub(&g_rans_consumers[r.counter], 1u);
}

// chunk state arrays (device scratch): start[] = entering state each piece was last encoded from, exit[] = its exit
// state, nbytes[]; cand_start / cand_exit[32 J] = the candidate entering / exit states found by the exploration,
// cand_mid = the candidates' states at the inner sub-chunk boundaries.
struct RansChunkState {
  uint32_t* start; uint32_t* exit; uint32_t* nbytes;  // per encoded piece (a chunk, or a sub-chunk in the lane-parallel encode pass)
  uint32_t* chain_start;                               // per exploration chunk: its true entering state (phase B)
  uint32_t* cand_start; uint32_t* cand_exit;           // per chunk x 32 lanes
  uint32_t* cand_mid;                                  // per chunk x (sub - 1) inner boundaries x 32 lanes
  uint32_t* offset;                                    // per piece: where its bytes go in the payload (phase D)
};

__host__ __device__ __forceinline__ uint64_t rans_chunk_capacity(uint32_t chunk) { return 3ull * chunk + 8; }

// 32 start states spread geometrically over the state interval [l_base, 256 l_base): lane -> l_base * 2^(lane / 4)
__device__ __forceinline__ uint32_t rans_guess_state(uint32_t lane, uint32_t l_base) {
  const uint32_t frac = (lane & 3u) == 0 ? 65536u : (lane & 3u) == 1 ? 77936u : (lane & 3u) == 2 ? 92682u : 110218u;  // 2^(k/4) in Q16
  return (uint32_t)(((unsigned long long)l_base * frac) >> 16) << (lane >> 2);
}

// phase A — exploration (two chunks per CTA): no bytes, 32 candidate (entering state -> exit state) pairs per chunk
__device__ __forceinline__ void rans_explore_body(const uint32_t* __restrict__ symbols, unsigned long long n, const uint4* __restrict__ table,
                                                  const RansChunkState& cs, uint32_t num_chunks, uint32_t chunk_steps, uint32_t warmup_steps, uint32_t sub,
                                                  AttrStats* stats, uint32_t blk) {
  __shared__ RansShared sh2[2];
  __shared__ int role_slot[2];
  __shared__ uint32_t sched_slot[4];
  __shared__ uint32_t s_abort;  // one read for the whole CTA: warps that saw different values would part ways before a named barrier
  if (threadIdx.x == 0) s_abort = *(volatile uint32_t*)&stats->error_flags;
  __syncthreads();
  if (s_abort) return;
  if (2ull * blk + (threadIdx.x >> 6) >= num_chunks) return;  // both warps of the pair leave together
  const RansRole role = rans_role(role_slot, sched_slot);
  const unsigned long long j = 2ull * blk + role.pair;
  RansShared& sh = sh2[role.pair];
  const uint32_t P = stats->precision, K = stats->num_table_symbols;
  const uint32_t lane = threadIdx.x & 31;
  const unsigned long long e_main = j * chunk_steps;
  const unsigned long long e_end = min(e_main + (unsigned long long)chunk_steps, n);
  const unsigned long long e_begin = e_main > warmup_steps ? e_main - warmup_steps : 0;
  const uint32_t l_base = 4u << P;
  const uint32_t x_in = e_begin == 0 ? l_base : rans_guess_state(lane, l_base);  // a warm-up from step 0 is the true trajectory
  rans_encode_range(sh, symbols, n, table, K, P, e_begin, e_main, e_end, x_in, nullptr, role.bar_base, role.is_consumer, false,
                    chunk_steps / sub / 32);
  if (role.is_consumer) {
    cs.cand_start[j * 32 + lane] = sh.x_main[lane];
    cs.cand_exit[j * 32 + lane] = sh.x_exit[lane];
    for (uint32_t k = 0; k + 1 < sub; ++k) cs.cand_mid[(j * (sub - 1) + k) * 32 + lane] = sh.x_mid[k][lane];  // unreached boundaries lie past the end
    if (lane == 0 && sh.err) atomicOr(&stats->error_flags, sh.err);
  }
  rans_role_release(role);
}
__global__ void __launch_bounds__(128) rans_explore_kernel(const uint32_t* __restrict__ symbols, unsigned long long n, const uint4* __restrict__ table,
                                                           RansChunkState cs, uint32_t num_chunks, uint32_t chunk_steps, uint32_t warmup_steps, uint32_t sub,
                                                           AttrStats* stats) {
  rans_explore_body(symbols, n, table, cs, num_chunks, chunk_steps, warmup_steps, sub, stats, blockIdx.x);
}

// phase B — the chain (one warp pair): walks the chunks in order carrying the TRUE state. A chunk whose candidates contain
// it (one ballot) hands over the matching exit state; otherwise (rare: every trajectory of the exploration missed) the
// chunk is run from the true state right here. The candidate rows go through shared memory in tiles of 32 chunks; the
// next tile's global loads are in flight while the current tile is walked.
constexpr int kChainTile = 32;
__device__ __forceinline__ void rans_chain_body(const uint32_t* __restrict__ symbols, unsigned long long n, const uint4* __restrict__ table,
                                                const RansChunkState& cs, uint32_t num_chunks, uint32_t chunk_steps, uint32_t nsub, AttrStats* stats) {
  __shared__ RansShared sh;
  __shared__ uint32_t tile_s[2][kChainTile][32], tile_e[2][kChainTile][32];
  __shared__ uint32_t s_abort;  // one read for the whole CTA: warps that saw different values would part ways before a named barrier
  if (threadIdx.x == 0) s_abort = *(volatile uint32_t*)&stats->error_flags;
  __syncthreads();
  if (s_abort) return;

  const uint32_t P = stats->precision, K = stats->num_table_symbols;
  const uint32_t lane = threadIdx.x & 31, half = threadIdx.x >> 5;
  const bool is_consumer = threadIdx.x < 32;
  uint32_t s = 4u << P;
  uint32_t misses = 0;
  constexpr int kRowsPerThread = kChainTile / 2;  // the two warps split the rows of a tile
  uint32_t reg_s[kRowsPerThread], reg_e[kRowsPerThread];
  auto fetch = [&](uint32_t tile) {
#pragma unroll
    for (int r = 0; r < kRowsPerThread; ++r) {
      const size_t j = (size_t)tile * kChainTile + 2 * r + half;
      const bool in = j < num_chunks;
      reg_s[r] = in ? __ldcs(cs.cand_start + j * 32 + lane) : 0u;
      reg_e[r] = in ? __ldcs(cs.cand_exit + j * 32 + lane) : 0u;
    }
  };
  auto park = [&](int buf) {
#pragma unroll
    for (int r = 0; r < kRowsPerThread; ++r) { tile_s[buf][2 * r + half][lane] = reg_s[r]; tile_e[buf][2 * r + half][lane] = reg_e[r]; }
  };
  const uint32_t num_tiles = (num_chunks + kChainTile - 1) / kChainTile;
  fetch(0);
  park(0);
  __syncthreads();
  for (uint32_t t = 0; t < num_tiles; ++t) {
    const int buf = (int)(t & 1);
    if (t + 1 < num_tiles) fetch(t + 1);
    const uint32_t j_end = min(num_chunks - t * kChainTile, (uint32_t)kChainTile);
    uint32_t cand_s = tile_s[buf][0][lane], cand_e = tile_e[buf][0][lane];
    for (uint32_t r = 0; r < j_end; ++r) {
      const uint32_t j = t * kChainTile + r;
      const uint32_t next_s = r + 1 < j_end ? tile_s[buf][r + 1][lane] : 0u, next_e = r + 1 < j_end ? tile_e[buf][r + 1][lane] : 0u;
      if (threadIdx.x == 0) cs.chain_start[j] = s;
      // lanes whose candidate matches hold the same trajectory, hence the same exit state: one OR-reduction hands it
      // over (states are >= l_base > 0, so 0 means that no lane matched)
      const uint32_t hit = __reduce_or_sync(0xFFFFFFFFu, cand_s == s ? cand_e : 0u);
      if (hit) {
        s = hit;
      } else {  // both warps see the same values and take this branch together
        // The chunk is run from the true state here, one sub-chunk at a time, and lane 0's candidate slots are overwritten
        // with the true states so that the encode pass finds them like any other match.
        const unsigned long long e_main = (unsigned long long)j * chunk_steps;
        const unsigned long long e_end = min(e_main + (unsigned long long)chunk_steps, n);
        const uint32_t sub = chunk_steps / nsub;
        if (threadIdx.x == 0) cs.cand_start[(size_t)j * 32] = s;
        for (uint32_t k = 0; k < nsub; ++k) {
          const unsigned long long e0 = e_main + (unsigned long long)k * sub;
          if (e0 >= e_end) break;
          rans_encode_range(sh, symbols, n, table, K, P, e0, e0, min(e0 + sub, e_end), s, nullptr, 1, is_consumer, false);
          s = sh.x_exit[0];
          if (threadIdx.x == 0 && k + 1 < nsub) cs.cand_mid[((size_t)j * (nsub - 1) + k) * 32] = s;
          __syncthreads();  // sh is rewritten by the next piece
        }
        ++misses;
      }
      cand_s = next_s; cand_e = next_e;
    }
    __syncthreads();  // everyone is done with buf^1's previous contents (tile t-1) ... and with this tile before it is overwritten
    if (t + 1 < num_tiles) park(buf ^ 1);
    __syncthreads();
  }
  if (threadIdx.x == 0) stats->pad[0] = misses;  // chunks whose true state matched no candidate
}
__global__ void __launch_bounds__(64) rans_chain_kernel(const uint32_t* __restrict__ symbols, unsigned long long n, const uint4* __restrict__ table,
                                                        RansChunkState cs, uint32_t num_chunks, uint32_t chunk_steps, uint32_t nsub, AttrStats* stats) {
  rans_chain_body(symbols, n, table, cs, num_chunks, chunk_steps, nsub, stats);
}

// tests only (DXO_RANS_FAULT): every fifth chunk gets a wrong entering state, which the fix-up has to repair
__global__ void rans_fault_kernel(RansChunkState cs, uint32_t num_chunks, AttrStats* stats) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < num_chunks && j % 5 == 2) cs.chain_start[j] = rans_guess_state(j & 31, 4u << stats->precision);
}

// phase C — every chunk is encoded once from its true entering state (two chunks per CTA)
__device__ __forceinline__ void rans_encode_body(const uint32_t* __restrict__ symbols, unsigned long long n, const uint4* __restrict__ table,
                                                 uint8_t* __restrict__ scratch, const RansChunkState& cs, uint32_t num_chunks, uint32_t chunk_steps,
                                                 AttrStats* stats, uint32_t blk) {
  __shared__ RansShared sh2[2];
  __shared__ int role_slot[2];
  __shared__ uint32_t sched_slot[4];
  __shared__ uint32_t s_abort;  // one read for the whole CTA: warps that saw different values would part ways before a named barrier
  if (threadIdx.x == 0) s_abort = *(volatile uint32_t*)&stats->error_flags;
  __syncthreads();
  if (s_abort) return;
  if (2ull * blk + (threadIdx.x >> 6) >= num_chunks) return;
  const RansRole role = rans_role(role_slot, sched_slot);
  const unsigned long long j = 2ull * blk + role.pair;
  RansShared& sh = sh2[role.pair];
  const uint32_t P = stats->precision, K = stats->num_table_symbols;
  const uint32_t in = j == 0 ? (4u << P) : cs.chain_start[j];
  const unsigned long long e_main = j * chunk_steps;
  const unsigned long long e_end = min(e_main + (unsigned long long)chunk_steps, n);
  rans_encode_range(sh, symbols, n, table, K, P, e_main, e_main, e_end, in, scratch + j * rans_chunk_capacity(chunk_steps), role.bar_base, role.is_consumer);
  if (role.is_consumer && (threadIdx.x & 31) == 0) {
    cs.start[j] = in;
    cs.exit[j] = sh.x_exit[0];
    cs.nbytes[j] = sh.nbytes;
    if (sh.err) atomicOr(&stats->error_flags, sh.err);
  }
  rans_role_release(role);
}
__global__ void __launch_bounds__(128) rans_encode_kernel(const uint32_t* __restrict__ symbols, unsigned long long n, const uint4* __restrict__ table,
                                                          uint8_t* __restrict__ scratch, RansChunkState cs, uint32_t num_chunks, uint32_t chunk_steps, AttrStats* stats) {
  rans_encode_body(symbols, n, table, scratch, cs, num_chunks, chunk_steps, stats, blockIdx.x);
}

// phase C, lane-parallel variant — one THREAD per piece (a chunk's sub-chunk; 32 pieces per warp) instead of one warp pair
// per chunk. All that the serial chain needs is one lane, so this spends ~1/20 of the instruction slots of the warp-pair
// kernel; a lone warp issues its ~40 instructions per step more slowly than the specialised consumer (≈ 85 ns against
// 28 ns per step), which is why the pieces are short: the exploration's checkpoints give the true entering state of up to 16
// pieces per chunk, so a stream makes 16x more pieces than chunks and the pass takes ~30 us instead of ~200. To keep the
// inner loop's memory accesses coalesced or in shared memory: per 32 steps the warp loads, for each of its 32 pieces, one
// coalesced 128-byte row of symbols and parks it in shared memory with a 33-word pitch (conflict-free both ways), the next
// group's loads are in flight meanwhile; table rows live in shared memory (alphabets up to kLaneSmemRows, else read through
// L1); bytes are collected in a 64-bit register and written as aligned 32-bit words. A row with f = 2^21, M = 0, cum = 0 is
// the identity step and stands in for every step outside a piece, so the loop has no per-lane control flow.
constexpr int kLaneThreads = 128;          // pieces per CTA (4 warps)
constexpr uint32_t kLaneSmemRows = 4096;   // 64 KB of table rows in shared memory
constexpr uint32_t kLanePitch = 33;        // words per staged symbol row
constexpr uint32_t kLaneDeadSymbol = 0xFFFFFFFFu;

struct RansLane {
  uint32_t x;               // coder state
  unsigned long long acc;   // pending output bytes (little end first)
  uint32_t fill8;           // bits in acc (< 32 between steps)
  uint8_t* out;             // next aligned word of this chunk's byte string
};

__device__ __forceinline__ void rans_lane_step(RansLane& L, const uint4 r, uint32_t two_p) {
  const uint32_t x = L.x;
  const uint32_t thr = r.x << 10;
  const uint32_t q0 = (__umulhi(x, r.z) + (r.x == 1u ? 1u : 0u)) >> r.w;  // floor(x / f), see "rANS division"
  const bool p1 = x >= thr, p2 = (x >> 8) >= thr, p3 = (x >> 16) >= thr;
  const uint32_t k8 = p2 ? (p3 ? 24u : 16u) : (p1 ? 8u : 0u);
  L.acc |= (unsigned long long)(x & ((1u << k8) - 1u)) << L.fill8;
  L.fill8 += k8;
  if (L.fill8 >= 32) {
    *reinterpret_cast<uint32_t*>(L.out) = (uint32_t)L.acc;
    L.out += 4; L.acc >>= 32; L.fill8 -= 32;
  }
  L.x = (q0 >> k8) * (two_p - r.x) + ((x >> k8) + r.y);
}

// 32 steps of every lane; `mine` = this lane's staged symbols, rows[K] = identity row
template <bool SMEM>
__device__ __forceinline__ void rans_lane_group(RansLane& L, const uint32_t* mine, const uint4* __restrict__ rows, uint32_t K, uint32_t two_p) {
  uint32_t s[32];
  uint4 r[32];
  const uint4 dead = make_uint4(1u << 21, 0u, 0u, 0u);
  auto row = [&](uint32_t sym) -> uint4 {
    if (SMEM) return rows[min(sym, K)];
    return sym < K ? __ldg(rows + sym) : dead;
  };
#pragma unroll
  for (int t = 0; t < 4; ++t) s[t] = mine[t];
  r[0] = row(s[0]); r[1] = row(s[1]);
#pragma unroll
  for (int t = 0; t < 32; ++t) {
    if (t + 4 < 32) s[t + 4] = mine[t + 4];
    if (t + 2 < 32) r[t + 2] = row(s[t + 2]);
    rans_lane_step(L, r[t], two_p);
  }
}

template <bool SMEM>
__device__ __forceinline__ void rans_encode_lanes_body(const uint32_t* __restrict__ symbols, unsigned long long n, const uint4* __restrict__ table,
                                                       uint8_t* __restrict__ scratch, const RansChunkState& cs, uint32_t num_chunks, uint32_t num_pieces,
                                                       uint32_t Cs, uint32_t sub, uint32_t P, uint32_t K, AttrStats* stats, uint32_t blk) {
  extern __shared__ uint4 lane_smem[];
  const uint4* rows = table;
  if (SMEM) {
    for (uint32_t i = threadIdx.x; i < K; i += blockDim.x) {
      uint4 e = __ldg(table + i);
      if (e.x == 0) e = make_uint4(1u << 21, 0u, 0u, 0u);  // zero-frequency symbols cannot occur in a stream the table was built from
      lane_smem[i] = e;
    }
    if (threadIdx.x == 0) lane_smem[K] = make_uint4(1u << 21, 0u, 0u, 0u);
    __syncthreads();
    rows = lane_smem;
  }
  uint32_t* stage = reinterpret_cast<uint32_t*>(lane_smem + (SMEM ? K + 1 : 0)) + (threadIdx.x >> 5) * (32 * kLanePitch);
  const uint32_t lane = threadIdx.x & 31;
  // piece q = sub-chunk (q % sub) of exploration chunk (q / sub); Cs steps each
  const unsigned long long q = (unsigned long long)blk * kLaneThreads + threadIdx.x;
  const unsigned long long q0 = q - lane;
  if (q0 >= num_pieces) return;
  const uint32_t two_p = 1u << P;
  // Entering states: a chunk's first piece enters with the chain's state; the inner pieces with the checkpoints of the
  // exploration lane that the chain matched (it IS the true trajectory from the chunk start on). The warp looks that lane
  // up for the 32 / sub chunks it covers with one ballot each. No match (never observed) leaves state 0, whose bytes
  // the verification of phase D rejects and repairs.
  uint32_t in = 0;
  {
    const unsigned long long my_chunk = q / sub;
    const uint32_t my_sub = (uint32_t)(q % sub);
    for (uint32_t c = 0; c < 32u / sub; ++c) {
      const unsigned long long jj = q0 / sub + c;
      if (jj >= num_chunks) break;
      const uint32_t s_j = jj == 0 ? (4u << P) : __ldg(cs.chain_start + jj);
      const uint32_t cand = num_chunks > 1 ? cs.cand_start[jj * 32 + lane] : s_j;
      const uint32_t m = __ballot_sync(0xFFFFFFFFu, cand == s_j);
      if (jj == my_chunk && q < num_pieces) {
        if (my_sub == 0) in = s_j;
        else if (m) in = cs.cand_mid[(jj * (sub - 1) + (my_sub - 1)) * 32 + (__ffs(m) - 1)];
      }
    }
  }
  uint8_t* out = scratch + q * rans_chunk_capacity(Cs);
  RansLane L{in, 0ull, 0u, out};
  const uint32_t groups = Cs / 32;
  const unsigned long long e_first = q0 * Cs + lane;  // lane's step inside piece row 0, group 0
  uint32_t nxt[32];
  auto load_group = [&](uint32_t g) {
#pragma unroll
    for (int l = 0; l < 32; ++l) {
      const unsigned long long e = e_first + (unsigned long long)l * Cs + 32ull * g;
      nxt[l] = e < n ? __ldcs(symbols + (n - 1 - e)) : kLaneDeadSymbol;  // steps past the end are identity steps
    }
  };
  auto park_group = [&]() {
#pragma unroll
    for (int l = 0; l < 32; ++l) stage[l * kLanePitch + lane] = nxt[l];
  };
  load_group(0);
  park_group();
  __syncwarp();
  const uint32_t* mine = stage + lane * kLanePitch;
  for (uint32_t g = 0; g < groups; ++g) {
    if (g + 1 < groups) load_group(g + 1);
    rans_lane_group<SMEM>(L, mine, rows, K, two_p);
    __syncwarp();
    if (g + 1 < groups) { park_group(); __syncwarp(); }
  }
  if (q < num_pieces) {
    uint32_t nb = (uint32_t)(L.out - out);
    for (uint32_t b = 0; b < L.fill8; b += 8) { L.out[b >> 3] = (uint8_t)(L.acc >> b); ++nb; }
    cs.start[q] = in;
    cs.exit[q] = L.x;
    cs.nbytes[q] = nb;
  }
}
__global__ void __launch_bounds__(kLaneThreads) rans_encode_lanes_kernel(const uint32_t* __restrict__ symbols, unsigned long long n,
                                                                         const uint4* __restrict__ table, uint8_t* __restrict__ scratch,
                                                                         RansChunkState cs, uint32_t num_chunks, uint32_t num_pieces, uint32_t Cs,
                                                                         uint32_t sub, uint32_t smem_rows, AttrStats* stats) {
  __shared__ uint32_t s_abort;  // one read for the whole CTA (the body synchronises the CTA; other CTAs may raise the flag meanwhile)
  if (threadIdx.x == 0) s_abort = *(volatile uint32_t*)&stats->error_flags;
  __syncthreads();
  if (s_abort) return;
  const uint32_t P = stats->precision, K = stats->num_table_symbols;
  if (K < smem_rows) rans_encode_lanes_body<true>(symbols, n, table, scratch, cs, num_chunks, num_pieces, Cs, sub, P, K, stats, blockIdx.x);
  else rans_encode_lanes_body<false>(symbols, n, table, scratch, cs, num_chunks, num_pieces, Cs, sub, P, K, stats, blockIdx.x);
}

// phase D — verification, fix-up and payload offsets (one CTA). The stream is exact iff every piece was encoded from the exit
// state of its predecessor; that is checked in parallel by the whole CTA, and only a violated link (which the chain makes
// impossible unless something upstream went wrong) starts the sequential repair by the CTA's first warp pair, so exactness
// never rests on the speculation. Then the exclusive prefix sum of the pieces' byte counts is left in cs.offset.
constexpr int kFixupThreads = 1024, kFixupItems = 8;
constexpr int kFixupPairBarrier = 2 * kRansStages + 2;  // named barrier of the repairing pair (rans_encode_range with bar_base 1 uses 1..7)
__device__ __forceinline__ void rans_fixup_body(const uint32_t* __restrict__ symbols, unsigned long long n,
                                                const uint4* __restrict__ table, uint8_t* __restrict__ scratch, const RansChunkState& cs,
                                                uint32_t num_chunks, uint32_t chunk_steps, AttrStats* stats) {
  __shared__ RansShared sh;
  __shared__ uint32_t s_in, s_bad, s_warp[kFixupThreads / 32];
  __shared__ uint32_t s_abort;  // one read for the whole CTA: warps that saw different values would part ways before a named barrier
  if (threadIdx.x == 0) s_abort = *(volatile uint32_t*)&stats->error_flags;
  __syncthreads();
  if (s_abort) return;

  const uint32_t P = stats->precision, K = stats->num_table_symbols;
  const uint32_t l_base = 4u << P;
  if (threadIdx.x == 0) s_bad = 0;
  __syncthreads();
  for (uint32_t base = 0; base < num_chunks; base += kFixupThreads * kFixupItems) {  // all loads of a tile are in flight together
    uint32_t e[kFixupItems], st[kFixupItems];
#pragma unroll
    for (int i = 0; i < kFixupItems; ++i) {
      const uint32_t j = base + i * kFixupThreads + threadIdx.x;
      e[i] = j < num_chunks && j > 0 ? __ldcg(cs.exit + j - 1) : l_base;
      st[i] = j < num_chunks ? __ldcg(cs.start + j) : l_base;
    }
    uint32_t bad = 0;
#pragma unroll
    for (int i = 0; i < kFixupItems; ++i) bad |= e[i] != st[i];
    if (bad) s_bad = 1;
  }
  __syncthreads();
  if (s_bad && threadIdx.x < 64) {
    const bool is_consumer = threadIdx.x < 32;
    for (uint32_t j = 0; j < num_chunks; ++j) {
      if (threadIdx.x == 0) s_in = j == 0 ? l_base : cs.exit[j - 1];
      named_bar_sync(kFixupPairBarrier);
      const uint32_t in = s_in;
      named_bar_sync(kFixupPairBarrier);  // s_in is rewritten by thread 0 in the next iteration
      if (in == cs.start[j]) continue;    // uniform: both warps read the same values
      const unsigned long long e_main = (unsigned long long)j * chunk_steps;
      const unsigned long long e_end = min(e_main + (unsigned long long)chunk_steps, n);
      rans_encode_range(sh, symbols, n, table, K, P, e_main, e_main, e_end, in, scratch + (unsigned long long)j * rans_chunk_capacity(chunk_steps), 1, is_consumer);
      if (threadIdx.x == 0) {
        cs.start[j] = in;
        cs.exit[j] = sh.x_exit[0];
        cs.nbytes[j] = sh.nbytes;
        stats->pad[1] += 1;  // pieces re-encoded by the sequential fix-up
        if (sh.err) atomicOr(&stats->error_flags, sh.err);
      }
      named_bar_sync(kFixupPairBarrier);
    }
  }
  __syncthreads();
  // exclusive scan of nbytes: kFixupItems rows of kFixupThreads pieces are loaded together, then scanned row by row
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t carry = 0;
  for (uint32_t base = 0; base < num_chunks; base += kFixupThreads * kFixupItems) {
    uint32_t v[kFixupItems];
#pragma unroll
    for (int i = 0; i < kFixupItems; ++i) {
      const uint32_t j = base + i * kFixupThreads + threadIdx.x;
      v[i] = j < num_chunks ? cs.nbytes[j] : 0u;
    }
#pragma unroll
    for (int i = 0; i < kFixupItems; ++i) {
      if (base + i * kFixupThreads >= num_chunks) break;  // uniform
      const uint32_t j = base + i * kFixupThreads + threadIdx.x;
      uint32_t inc = v[i];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (lane >= (uint32_t)d) inc += o; }
      if (lane == 31) s_warp[warp] = inc;
      __syncthreads();
      const uint32_t w = s_warp[lane];  // kFixupThreads / 32 == 32 partial sums
      uint32_t winc = w;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, winc, d); if (lane >= (uint32_t)d) winc += o; }
      const uint32_t before = __shfl_sync(0xFFFFFFFFu, winc - w, warp);
      const uint32_t total = __shfl_sync(0xFFFFFFFFu, winc, 31);
      if (j < num_chunks) cs.offset[j] = carry + before + (inc - v[i]);
      carry += total;
      __syncthreads();
    }
  }
}
__global__ void __launch_bounds__(kFixupThreads) rans_fixup_kernel(const uint32_t* __restrict__ symbols, unsigned long long n,
                                                                   const uint4* __restrict__ table, uint8_t* __restrict__ scratch, RansChunkState cs,
                                                                   uint32_t num_chunks, uint32_t chunk_steps, AttrStats* stats) {
  rans_fixup_body(symbols, n, table, scratch, cs, num_chunks, chunk_steps, stats);
}

// gather: piece j's bytes go to payload[offset[j]]; the last piece's threads append the flush bytes. A CTA copies
// `pieces_per_cta` pieces (a power of two <= 8), 256 / pieces_per_cta threads each.
__device__ __forceinline__ void rans_gather_body(const uint8_t* __restrict__ scratch, const RansChunkState& cs, uint32_t num_chunks, uint32_t chunk_steps,
                                                 uint32_t pieces_per_cta, uint8_t* __restrict__ out, AttrStats* stats, uint32_t blk) {
  if (stats->error_flags) { if (blk == 0 && threadIdx.x == 0) stats->payload_bytes = 0; return; }
  const uint32_t tpp = 256u / pieces_per_cta;  // threads per piece
  const uint32_t j = blk * pieces_per_cta + threadIdx.x / tpp, t = threadIdx.x % tpp;
  if (j >= num_chunks) return;
  const uint32_t off = num_chunks > 1 ? cs.offset[j] : 0u, nb = cs.nbytes[j];
  const uint8_t* src = scratch + (unsigned long long)j * rans_chunk_capacity(chunk_steps);
  for (uint32_t i = t; i < nb; i += tpp) out[off + i] = src[i];
  if (j + 1 == num_chunks && t == 0) {
    uint32_t pos = off + nb, err = 0;
    const uint32_t l_base = 4u << stats->precision;
    const uint32_t t = cs.exit[j] - l_base;  // flush (:48-68)
    if (t < (1u << 6)) { out[pos++] = (uint8_t)t; }
    else if (t < (1u << 14)) { const uint32_t v = 0x4000u + t; out[pos++] = (uint8_t)v; out[pos++] = (uint8_t)(v >> 8); }
    else if (t < (1u << 22)) { const uint32_t v = 0x800000u + t; out[pos++] = (uint8_t)v; out[pos++] = (uint8_t)(v >> 8); out[pos++] = (uint8_t)(v >> 16); }
    else if (t < (1u << 30)) { const uint32_t v = 0xC0000000u + t; out[pos++] = (uint8_t)v; out[pos++] = (uint8_t)(v >> 8); out[pos++] = (uint8_t)(v >> 16); out[pos++] = (uint8_t)(v >> 24); }
    else err |= kErrRansState;
    stats->payload_bytes = pos;
    if (err) atomicOr(&stats->error_flags, err);
  }
}
__global__ void __launch_bounds__(256) rans_gather_kernel(const uint8_t* __restrict__ scratch, RansChunkState cs, uint32_t num_chunks, uint32_t chunk_steps,
                                                          uint32_t pieces_per_cta, uint8_t* __restrict__ out, AttrStats* stats) {
  rans_gather_body(scratch, cs, num_chunks, chunk_steps, pieces_per_cta, out, stats, blockIdx.x);
}

// The lane-parallel encode kernels use more than 48 KB of dynamic shared memory: opted in once per (kernel, device).
static void allow_lane_smem(const void* kernel) {
  static std::mutex mu;
  static std::vector<std::pair<const void*, int>> done;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  for (const auto& d : done) if (d.first == kernel && d.second == dev) return;
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       (int)((size_t)(kLaneSmemRows + 1) * 16 + (size_t)(kLaneThreads / 32) * 32 * kLanePitch * 4));
  done.push_back({kernel, dev});
}

// Chunk size for a stream of n symbols. The exploration is fastest with at most one consumer warp per scheduler, i.e. up to
// 4 pairs per SM, so longer streams get longer chunks (up to 16384 steps; beyond that the SMs are full either way and
// the chain kernel would gain nothing). DXO_RANS_CHUNK pins the size.
constexpr uint32_t kRansChunkMin = 4096, kRansChunkMax = 16384;
static uint32_t rans_sm_count() {
  static const uint32_t num_sms = [] {
    int dev = 0, v = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    return (uint32_t)(v > 0 ? v : 148);
  }();
  return num_sms;
}
static RansPlan rans_plan_for(uint64_t n) {
  RansPlan p = rans_plan();
  if (!p.fixed_chunk) {
    const uint64_t unit = 32ull * p.sub, pairs = 4ull * rans_sm_count();
    const uint64_t c = ((n + pairs - 1) / pairs + unit - 1) / unit * unit;
    p.chunk = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(c, kRansChunkMin), kRansChunkMax);
  }
  return p;
}
uint32_t rans_num_chunks(uint64_t num_symbols) { const uint32_t c = rans_plan_for(num_symbols).chunk; return (uint32_t)((num_symbols + c - 1) / c); }
size_t rans_scratch_bytes(uint64_t num_symbols) {
  // Valid for every stream of at most num_symbols symbols and for both layouts of launch_rans_encode (pieces = sub-chunks
  // for the lane-parallel encode, whole chunks for the warp pairs): Q pieces of c steps hold Q (3 c + 8) <= 3 (n + c) + 8 Q bytes.
  const RansPlan p = rans_plan();
  const uint64_t c_min = p.fixed_chunk ? p.chunk : kRansChunkMin, c_max = p.fixed_chunk ? p.chunk : kRansChunkMax;
  const uint64_t piece_min = p.lanes ? c_min / p.sub : c_min;
  const size_t J = (size_t)(num_symbols / c_min + 1), Q = (size_t)(num_symbols / piece_min + 1);
  const size_t area = 3 * ((size_t)num_symbols + c_max) + 8 * Q;
  return area + 512 + (4 * Q + J + 64 * J + 32 * (kRansSubMax - 1) * J) * sizeof(uint32_t) + 64;
}

void launch_rans_encode(const uint32_t* symbols, uint64_t num_symbols, const uint4* rans_table, uint32_t table_capacity, void* scratch,
                        uint8_t* payload, AttrStats* stats, cudaStream_t s) {
  const RansPlan plan = rans_plan_for(num_symbols);
  const uint32_t J = (uint32_t)((num_symbols + plan.chunk - 1) / plan.chunk);
  const bool lanes = plan.lanes && J > 1;  // a single chunk has no exploration, hence no checkpoints: one warp pair codes it
  const uint32_t piece = lanes ? plan.chunk / plan.sub : plan.chunk;
  const uint32_t Q = (uint32_t)((num_symbols + piece - 1) / piece);
  uint8_t* bytes = (uint8_t*)scratch;
  size_t off = ((size_t)Q * rans_chunk_capacity(piece) + 255) / 256 * 256;
  uint32_t* u = (uint32_t*)(bytes + off);
  RansChunkState cs;
  cs.start = u; cs.exit = u + Q; cs.nbytes = u + 2 * (size_t)Q;
  cs.offset = u + 3 * (size_t)Q;
  cs.chain_start = u + 4 * (size_t)Q;
  cs.cand_start = cs.chain_start + J; cs.cand_exit = cs.cand_start + 32 * (size_t)J; cs.cand_mid = cs.cand_exit + 32 * (size_t)J;
  if (J > 1) {
    rans_explore_kernel<<<(J + 1) / 2, 128, 0, s>>>(symbols, num_symbols, rans_table, cs, J, plan.chunk, plan.warmup, plan.sub, stats);
    rans_chain_kernel<<<1, 64, 0, s>>>(symbols, num_symbols, rans_table, cs, J, plan.chunk, plan.sub, stats);
    if (plan.fault) rans_fault_kernel<<<(J + 255) / 256, 256, 0, s>>>(cs, J, stats);
  }
  if (lanes) {
    // shared memory is reserved for what the alphabet bound allows; the kernel reads rows through L1 when K does not fit
    const uint32_t smem_rows = std::min(table_capacity + 1u, kLaneSmemRows + 1u);
    const size_t sm = (size_t)smem_rows * 16 + (size_t)(kLaneThreads / 32) * 32 * kLanePitch * 4;
    allow_lane_smem((const void*)rans_encode_lanes_kernel);
    rans_encode_lanes_kernel<<<(Q + kLaneThreads - 1) / kLaneThreads, kLaneThreads, sm, s>>>(symbols, num_symbols, rans_table, bytes, cs, J, Q, piece,
                                                                                               plan.sub, smem_rows, stats);
  } else {
    rans_encode_kernel<<<(J + 1) / 2, 128, 0, s>>>(symbols, num_symbols, rans_table, bytes, cs, J, plan.chunk, stats);
  }
  if (Q > 1) rans_fixup_kernel<<<1, kFixupThreads, 0, s>>>(symbols, num_symbols, rans_table, bytes, cs, Q, piece, stats);
  const uint32_t ppc = piece <= 256 ? 8u : piece <= 512 ? 4u : piece <= 1024 ? 2u : 1u;
  rans_gather_kernel<<<(Q + ppc - 1) / ppc, 256, 0, s>>>(bytes, cs, Q, piece, ppc, payload, stats);
}
int rans_launch_count(uint64_t num_symbols) { return rans_num_chunks(num_symbols) > 1 ? 5 : 2; }

// ---------------------------------------------------------------------------------------
// Side-stream preparation for the host coders (mesh_normal_prediction.rs:147-163,
// mesh_prediction_for_texture_coordinates.rs:221-260). The coding itself is a serial binary rANS on the host; what is
// data-parallel moves here: K5's flips are counted; K6's per-element flags (0 none / 1 false / 2 true) are compacted in
// order (cub::DeviceSelect) and the forward transitions (first compared with `true`) are counted.
// scalars[0] = number of entries, scalars[1] = ones (normals) / transitions (texcoords).
struct NonZeroFlag {
  __host__ __device__ bool operator()(uint8_t v) const { return v != 0; }
};
__global__ void __launch_bounds__(kThreads) count_ones_kernel(const uint8_t* __restrict__ flags, uint32_t n, uint32_t* scalars) {
  uint32_t ones = 0;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) ones += __ldcs(flags + i) != 0;
  ones = __reduce_add_sync(0xFFFFFFFFu, ones);
  if ((threadIdx.x & 31) == 0 && ones) atomicAdd(scalars + 1, ones);
  if (blockIdx.x == 0 && threadIdx.x == 0) scalars[0] = n;
}
__global__ void __launch_bounds__(kThreads) count_transitions_kernel(const uint8_t* __restrict__ compact, const uint32_t* __restrict__ scalars_in,
                                                                     uint32_t* scalars) {
  const uint32_t m = scalars_in[0];
  uint32_t t = 0;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < m; k += stride) {
    const uint8_t cur = compact[k], last = k ? compact[k - 1] : (uint8_t)2;  // the scan starts from `true`
    t += cur != last;
  }
  t = __reduce_add_sync(0xFFFFFFFFu, t);
  if ((threadIdx.x & 31) == 0 && t) atomicAdd(scalars + 1, t);
}
size_t side_prepare_scratch_bytes(uint32_t n) {
  size_t b = 0;
  cub::DeviceSelect::If(nullptr, b, (const uint8_t*)nullptr, (uint8_t*)nullptr, (uint32_t*)nullptr, (int)n, NonZeroFlag{});
  return b + 256;
}
void launch_count_flips(const uint8_t* flips, uint32_t n, uint32_t* scalars, cudaStream_t s) {
  cudaMemsetAsync(scalars, 0, 8, s);
  count_ones_kernel<<<grid_for(n, kThreads * 8), kThreads, 0, s>>>(flips, n, scalars);
}
void launch_compact_orientations(const uint8_t* flags, uint32_t n, uint8_t* compact, uint32_t* scalars, void* scratch, size_t scratch_bytes,
                                 cudaStream_t s) {
  cudaMemsetAsync(scalars, 0, 8, s);
  cub::DeviceSelect::If(scratch, scratch_bytes, flags, compact, scalars, (int)n, NonZeroFlag{}, s);
  count_transitions_kernel<<<grid_for(n, kThreads * 8), kThreads, 0, s>>>(compact, scalars, scalars);
}

// ---------------------------------------------------------------------------------------
// K12 — CornerTable::compute_table (corner_table/mod.rs:252-340) for the manifold,
// consistently oriented case: sort half edges by
// (min(a,b), max(a,b)); an undirected edge with exactly two half edges of opposite
// direction and different tips pairs them. Anything else (3+ half edges, equal
// direction, equal tips, degenerate faces) raises not_exact: those results depend on
// corner order and the caller must use the sequential matcher.
__global__ void __launch_bounds__(kThreads) halfedge_keys_kernel(const uint32_t* __restrict__ cv, unsigned long long num_corners, uint32_t vertex_bits,
                                                                 unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals, uint32_t* not_exact) {
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long c = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; c < num_corners; c += stride) {
    const uint32_t cc = (uint32_t)c;
    const uint32_t tip = cv[cc], src = cv[cnext(cc)], snk = cv[cprev(cc)];
    if (tip == src || tip == snk || src == snk) *not_exact = 1;  // degenerate face
    const uint32_t a = min(src, snk), b = max(src, snk);
    keys[c] = ((unsigned long long)a << vertex_bits) | b;  // only 2 * vertex_bits key bits are sorted
    vals[c] = cc;
  }
}
__global__ void __launch_bounds__(kThreads) halfedge_pair_kernel(const uint32_t* __restrict__ cv, unsigned long long num_corners,
                                                                 const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ vals,
                                                                 uint32_t* __restrict__ opposite, uint32_t* not_exact) {
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < num_corners; i += stride) {
    const unsigned long long k = keys[i];
    const bool same_prev = i > 0 && keys[i - 1] == k;
    const bool same_next = i + 1 < num_corners && keys[i + 1] == k;
    const uint32_t c = vals[i];
    if (same_prev && same_next) { *not_exact = 1; continue; }  // 3+ half edges on one edge
    if (!same_prev && !same_next) { opposite[c] = kNoneDev; continue; }
    const uint32_t other = same_next ? vals[i + 1] : vals[i - 1];
    if (same_next && i + 2 < num_corners && keys[i + 2] == k) { *not_exact = 1; continue; }
    if (same_prev && i >= 2 && keys[i - 2] == k) { *not_exact = 1; continue; }
    const uint32_t src = cv[cnext(c)], osrc = cv[cnext(other)];
    if (src == osrc) { *not_exact = 1; continue; }               // same direction: inconsistent orientation
    if (cv[c] == cv[other]) { *not_exact = 1; continue; }        // equal tips are skipped by the reference (:308-310)
    opposite[c] = other;
  }
}

size_t corner_table_scratch_bytes(uint64_t num_corners) {
  size_t cub_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                  (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)num_corners);
  const size_t a = ((num_corners * 8 + 255) / 256) * 256, b = ((num_corners * 4 + 255) / 256) * 256;
  return 2 * a + 2 * b + cub_bytes + 256;
}

void launch_corner_table_opposites(const uint32_t* corner_vertex, uint64_t num_corners, uint32_t num_vertices, uint32_t* opposite, uint32_t* not_exact_flag,
                                   void* scratch, size_t scratch_bytes, cudaStream_t s) {
  // vertex ids below num_vertices (0 = unknown: all 32 bits): the radix sort runs over 2 * vertex_bits key bits only
  uint32_t vertex_bits = 32;
  if (num_vertices) { vertex_bits = 1; while (vertex_bits < 32 && ((uint64_t)(num_vertices - 1) >> vertex_bits) != 0) ++vertex_bits; }
  const size_t a = ((num_corners * 8 + 255) / 256) * 256, b = ((num_corners * 4 + 255) / 256) * 256;
  uint8_t* p = (uint8_t*)scratch;
  unsigned long long* keys_in = (unsigned long long*)p; p += a;
  unsigned long long* keys_out = (unsigned long long*)p; p += a;
  uint32_t* vals_in = (uint32_t*)p; p += b;
  uint32_t* vals_out = (uint32_t*)p; p += b;
  size_t cub_bytes = scratch_bytes - (2 * a + 2 * b);
  const int g = grid_for(num_corners);
  halfedge_keys_kernel<<<g, kThreads, 0, s>>>(corner_vertex, num_corners, vertex_bits, keys_in, vals_in, not_exact_flag);
  cub::DeviceRadixSort::SortPairs(p, cub_bytes, keys_in, keys_out, vals_in, vals_out, (int)num_corners, 0, (int)(2 * vertex_bits), s);
  halfedge_pair_kernel<<<g, kThreads, 0, s>>>(corner_vertex, num_corners, keys_out, vals_out, opposite, not_exact_flag);
}

// ---------------------------------------------------------------------------------------
// K13 — CornerTable::compute_left_most_corners (corner_table/mod.rs:342-416) for meshes whose vertices
// all have a single fan. The reference walks the corners in order; the first corner of a vertex starts
// its fan, and left_most[v] is the last corner reached by swinging left before the walk hits a boundary
// or returns to the start. With one fan per vertex that first corner is the vertex's smallest corner
// index, so the vertices are independent. A vertex whose fan does not contain all of its corners has a
// second fan (the reference then splits it): flagged, and the caller runs the sequential pass instead.
constexpr uint32_t kFanUnusedVertex = 1u, kFanSplitVertex = 2u;

__global__ void __launch_bounds__(kThreads) vertex_first_corner_kernel(const uint32_t* __restrict__ cv, unsigned long long num_corners,
                                                                       uint32_t* __restrict__ first_corner, uint32_t* __restrict__ valence) {
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long c = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; c < num_corners; c += stride) {
    const uint32_t v = __ldcs(cv + c);
    atomicMin(first_corner + v, (uint32_t)c);
    atomicAdd(valence + v, 1u);
  }
}
__global__ void __launch_bounds__(kThreads) left_most_kernel(const uint32_t* __restrict__ opposite, const uint32_t* __restrict__ first_corner,
                                                             const uint32_t* __restrict__ valence, uint32_t num_vertices,
                                                             uint32_t* __restrict__ left_most, uint8_t* __restrict__ interior, uint32_t* flags) {
  uint32_t bad = 0;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < num_vertices; v += stride) {
    const uint32_t n = valence[v], c = first_corner[v];
    if (n == 0) { bad |= kFanUnusedVertex; left_most[v] = kNoneDev; if (interior) interior[v] = 0; continue; }
    uint32_t last = c, count = 1, a = c;
    bool open = false;
    for (;;) {  // swing left: opposite(next(a)) -> next
      const uint32_t o = __ldg(opposite + cnext(a));
      if (o == kNoneDev) { open = true; break; }
      a = cnext(o);
      if (a == c || count > n) break;
      last = a;
      ++count;
    }
    if (open) {  // the corners to the right of the start belong to the fan as well
      a = c;
      for (;;) {
        const uint32_t o = __ldg(opposite + cprev(a));
        if (o == kNoneDev || count > n) break;
        a = cprev(o);
        ++count;
      }
    }
    if (count != n) bad |= kFanSplitVertex;
    left_most[v] = last;
    if (interior) interior[v] = __ldg(opposite + cnext(last)) != kNoneDev ? 1 : 0;  // !is_on_boundary(v): swing_left(left_most[v]) exists (corner_table/mod.rs:36-38)
  }
  if (bad) atomicOr(flags, bad);
}

// corners without an opposite, ascending (ordered compaction): the outer loop of Edgebreaker::compute_boundaries
// (edgebreaker.rs:195-224) visits exactly these
struct IsBoundaryCorner {
  const uint32_t* opposite;
  __host__ __device__ bool operator()(uint32_t c) const { return opposite[c] == kNoneDev; }
};
size_t boundary_list_scratch_bytes(uint64_t num_corners) {
  size_t b = 0;
  thrust::counting_iterator<uint32_t> it(0);
  cub::DeviceSelect::If(nullptr, b, it, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)num_corners, IsBoundaryCorner{nullptr});
  return b + 256;
}
void launch_boundary_list(const uint32_t* opposite, uint64_t num_corners, void* scratch, size_t scratch_bytes, uint32_t* list, uint32_t* count,
                          cudaStream_t s) {
  thrust::counting_iterator<uint32_t> it(0);
  cub::DeviceSelect::If(scratch, scratch_bytes, it, list, count, (int)num_corners, IsBoundaryCorner{opposite}, s);
}

size_t left_most_scratch_bytes(uint32_t num_vertices) { return 2 * (size_t)num_vertices * sizeof(uint32_t) + 256; }
void launch_left_most(const uint32_t* corner_vertex, const uint32_t* opposite, uint64_t num_corners, uint32_t num_vertices, void* scratch,
                      uint32_t* left_most, uint32_t* flags, cudaStream_t s, uint8_t* interior) {
  uint32_t* first_corner = (uint32_t*)scratch;
  uint32_t* valence = first_corner + num_vertices;
  cudaMemsetAsync(first_corner, 0xFF, sizeof(uint32_t) * num_vertices, s);
  cudaMemsetAsync(valence, 0, sizeof(uint32_t) * num_vertices, s);
  vertex_first_corner_kernel<<<grid_for(num_corners), kThreads, 0, s>>>(corner_vertex, num_corners, first_corner, valence);
  left_most_kernel<<<grid_for(num_vertices), kThreads, 0, s>>>(opposite, first_corner, valence, num_vertices, left_most, interior, flags);
}

// ---------------------------------------------------------------------------------------
// K14 — AttributeCornerTable::new + recompute_vertices (attribute_corner_table.rs:16-137) on the device.
// (a) per corner: the edge opposite to it is a seam when it is a mesh boundary or when the attribute values
//     at its two end points differ between the two faces (symmetric, so each corner decides for itself);
// (b) per universal vertex, in vertex order: the fan is walked to the right starting at the first corner after a
//     seam, and every seam crossed opens a new attribute vertex. The ids are consecutive in (vertex, walk) order,
//     i.e. an exclusive prefix sum of the per-vertex counts — so pass (b) runs twice around a scan.
constexpr uint32_t kSeamBadPoint = 1u, kSeamClosedFan = 2u, kSeamInterior = 4u;  // kSeamInterior is informational: a seam that is not a mesh boundary exists

__global__ void __launch_bounds__(kThreads) seam_flags_kernel(const uint32_t* __restrict__ corner_point, const uint32_t* __restrict__ map,
                                                              uint32_t num_points, const uint32_t* __restrict__ cv,
                                                              const uint32_t* __restrict__ opposite, unsigned long long num_corners,
                                                              uint8_t* __restrict__ seam, uint8_t* __restrict__ vertex_on_seam, uint32_t* flags) {
  uint32_t bad = 0;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  auto val = [&](uint32_t corner) -> uint32_t {
    const uint32_t p = __ldg(corner_point + corner);
    if (p >= num_points) { bad |= kSeamBadPoint; return 0u; }
    return map ? __ldg(map + p) : p;
  };
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < num_corners; i += stride) {
    const uint32_t c = (uint32_t)i;
    const uint32_t o = __ldg(opposite + c);
    const uint32_t cn = cnext(c), cp = cprev(c);
    bool is_seam;
    if (o == kNoneDev) is_seam = true;  // mesh boundary counts as a seam
    else {
      const uint32_t on = cnext(o), op = cprev(o);
      is_seam = val(cn) != val(op) || val(cp) != val(on);
    }
    seam[c] = is_seam ? 1 : 0;
    if (is_seam) { vertex_on_seam[__ldg(cv + cn)] = 1; vertex_on_seam[__ldg(cv + cp)] = 1; }
    if (is_seam && o != kNoneDev) bad |= kSeamInterior;
  }
  if (bad) atomicOr(flags, bad);
}

// first corner of the attribute fan walk of universal vertex v (rotated to just after a seam)
__device__ __forceinline__ uint32_t seam_walk_start(uint32_t v, const uint32_t* __restrict__ left_most_u, const uint32_t* __restrict__ opposite,
                                                    const uint8_t* __restrict__ seam, const uint8_t* __restrict__ vertex_on_seam,
                                                    uint32_t num_corners, uint32_t& bad) {
  const uint32_t c = __ldg(left_most_u + v);
  uint32_t first = c;
  if (vertex_on_seam[v]) {
    uint32_t guard = num_corners;
    for (;;) {  // attribute swing left: stops at a seam
      const uint32_t e = cnext(first);
      const uint32_t o = seam[e] ? kNoneDev : __ldg(opposite + e);
      if (o == kNoneDev) break;
      first = cnext(o);
      if (first == c || --guard == 0) { bad |= kSeamClosedFan; break; }
    }
  }
  return first;
}

template <bool ASSIGN>
__global__ void __launch_bounds__(kThreads) seam_vertices_kernel(const uint32_t* __restrict__ left_most_u, const uint32_t* __restrict__ opposite,
                                                                 const uint8_t* __restrict__ seam, const uint8_t* __restrict__ vertex_on_seam,
                                                                 uint32_t num_vertices, uint32_t num_corners, uint32_t* __restrict__ count_or_base,
                                                                 uint32_t* __restrict__ corner_vertex, uint32_t* __restrict__ left_most_a,
                                                                 uint32_t* total, uint32_t* flags) {
  uint32_t bad = 0;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < num_vertices; v += stride) {
    const uint32_t first = seam_walk_start(v, left_most_u, opposite, seam, vertex_on_seam, num_corners, bad);
    uint32_t id = ASSIGN ? count_or_base[v] : 0u, n = 1;
    if (ASSIGN) { corner_vertex[first] = id; left_most_a[id] = first; }
    uint32_t s = first, guard = num_corners;
    for (;;) {  // universal swing right
      const uint32_t o = __ldg(opposite + cprev(s));
      if (o == kNoneDev) break;
      s = cprev(o);
      if (s == first || --guard == 0) break;
      if (seam[cnext(s)]) {  // crossing a seam starts a new attribute vertex
        ++n;
        if (ASSIGN) { ++id; left_most_a[id] = s; }
      }
      if (ASSIGN) corner_vertex[s] = id;
    }
    if (!ASSIGN) count_or_base[v] = n;
    else if (v + 1 == num_vertices) *total = count_or_base[v] + n;
  }
  if (bad) atomicOr(flags, bad);
}

size_t seam_table_scratch_bytes(uint32_t num_vertices) {
  size_t cub_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)num_vertices);
  return (((size_t)num_vertices + 255) / 256) * 256 + 2 * (((size_t)num_vertices * 4 + 255) / 256) * 256 + cub_bytes + 256;
}
void launch_seam_table(const uint32_t* corner_point, const uint32_t* map, uint32_t num_points, const uint32_t* cv, const uint32_t* opposite,
                       const uint32_t* left_most_u, uint64_t num_corners, uint32_t num_vertices, void* scratch, size_t scratch_bytes,
                       uint8_t* seam, uint32_t* corner_vertex, uint32_t* left_most_a, uint32_t* total, uint32_t* flags, cudaStream_t s) {
  uint8_t* p = (uint8_t*)scratch;
  const size_t a = (((size_t)num_vertices + 255) / 256) * 256, b = (((size_t)num_vertices * 4 + 255) / 256) * 256;
  uint8_t* vertex_on_seam = p; p += a;
  uint32_t* count = (uint32_t*)p; p += b;
  uint32_t* base = (uint32_t*)p; p += b;
  size_t cub_bytes = scratch_bytes - (a + 2 * b);
  cudaMemsetAsync(vertex_on_seam, 0, num_vertices, s);
  seam_flags_kernel<<<grid_for(num_corners), kThreads, 0, s>>>(corner_point, map, num_points, cv, opposite, num_corners, seam, vertex_on_seam, flags);
  const int g = grid_for(num_vertices);
  seam_vertices_kernel<false><<<g, kThreads, 0, s>>>(left_most_u, opposite, seam, vertex_on_seam, num_vertices, (uint32_t)num_corners, count, nullptr, nullptr, total, flags);
  cub::DeviceScan::ExclusiveSum(p, cub_bytes, count, base, (int)num_vertices, s);
  seam_vertices_kernel<true><<<g, kThreads, 0, s>>>(left_most_u, opposite, seam, vertex_on_seam, num_vertices, (uint32_t)num_corners, base, corner_vertex, left_most_a, total, flags);
}

#include "segmented.inl"

}  // namespace gpu
}  // namespace dxo
