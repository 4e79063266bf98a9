// Mesh build — the step immediately before the encode() boundary:
//   MeshBuilder::add_attribute -> Attribute::from / remove_duplicate_values   core/attribute/mod.rs:87-103, 394-452
//   MeshBuilder::build: dependency check, position first, point merge,         core/mesh/builder.rs:62-125
//     degenerate-face removal, unused-vertex removal                           core/mesh/builder.rs:129-373
// Paths relative to /root/reference/draco-oxide/src/.
//
// The reference finds duplicates by all-pairs comparison (values) and by hashing (points); both define the
// result as "first occurrence wins, survivors keep first-occurrence order". That result is order-independent
// given the input, so the device computes it by sorting:
//   1. every item gets a key of 32-bit words (value bits with -0.0 -> +0.0; an item containing a NaN gets a unique
//      salt word, since NaN != NaN makes it a value of its own);
//   2. a stable LSD radix sort of the item indices by those words (one cub::DeviceRadixSort pass per word) makes
//      equal keys contiguous with ascending original index inside a group, so a group's head is its first occurrence;
//   3. a scan over "is a first occurrence" in ORIGINAL index order numbers the survivors in first-occurrence order.
// The linear bookkeeping after that (compacting maps and value buffers, re-indexing faces) runs on the host.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <future>
#include <limits>
#include <memory>
#include <string>
#include <vector>

#include "common.hpp"

namespace dxo {
void cuda_check(cudaError_t e, const char* what);  // encoder.cpp

namespace {

constexpr int kThreads = 256;
inline int grid_for(uint64_t n) { return (int)std::min<uint64_t>((n + kThreads - 1) / kThreads, 148ull * 16); }

size_t component_bytes(uint32_t ct) {
  switch (ct) {
    case DXO_U8: case DXO_I8: return 1;
    case DXO_U16: case DXO_I16: return 2;
    case DXO_U32: case DXO_I32: case DXO_F32: return 4;
    case DXO_U64: case DXO_I64: case DXO_F64: return 8;
    default: return 0;
  }
}
uint32_t words_per_component(uint32_t ct) { return component_bytes(ct) == 8 ? 2u : 1u; }

// ---- kernels ---------------------------------------------------------------------------------------------------
// Key words of the values (SoA: words[w * n + i]); word index ncomp * wpc is the NaN salt.
__global__ void __launch_bounds__(kThreads) value_keys_kernel(const uint8_t* __restrict__ values, uint32_t n, uint32_t ct, uint32_t ncomp,
                                                              uint32_t comp_bytes, uint32_t wpc, uint32_t* __restrict__ words, uint32_t* has_nan) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    bool nan = false;
    const uint8_t* p = values + (size_t)i * ncomp * comp_bytes;
    for (uint32_t k = 0; k < ncomp; ++k) {
      uint32_t lo = 0, hi = 0;
      if (comp_bytes == 1) lo = p[k];
      else if (comp_bytes == 2) lo = (uint32_t)p[2 * k] | ((uint32_t)p[2 * k + 1] << 8);
      else if (comp_bytes == 4) { memcpy(&lo, p + 4 * k, 4); }
      else { memcpy(&lo, p + 8 * k, 4); memcpy(&hi, p + 8 * k + 4, 4); }
      if (ct == DXO_F32) {
        if ((lo & 0x7FFFFFFFu) == 0) lo = 0;                 // -0.0 == +0.0
        if ((lo & 0x7FFFFFFFu) > 0x7F800000u) nan = true;    // NaN != NaN
      } else if (ct == DXO_F64) {
        if (((hi & 0x7FFFFFFFu) | lo) == 0) hi = 0;
        if ((hi & 0x7FFFFFFFu) > 0x7FF00000u || ((hi & 0x7FFFFFFFu) == 0x7FF00000u && lo != 0)) nan = true;
      }
      words[(size_t)(k * wpc) * n + i] = lo;
      if (wpc == 2) words[(size_t)(k * wpc + 1) * n + i] = hi;
    }
    words[(size_t)(ncomp * wpc) * n + i] = nan ? i + 1u : 0u;
    if (nan) *has_nan = 1;
  }
}
__global__ void __launch_bounds__(kThreads) iota_kernel(uint32_t* p, uint32_t n) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = i;
}
__global__ void __launch_bounds__(kThreads) gather_kernel(const uint32_t* __restrict__ src, const uint32_t* __restrict__ perm, uint32_t n,
                                                          uint32_t* __restrict__ dst) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = src[perm[i]];
}
// head[i] = 1 when the item at sorted position i differs from its predecessor in any key word
__global__ void __launch_bounds__(kThreads) heads_kernel(const uint32_t* __restrict__ words, uint32_t num_words, const uint32_t* __restrict__ perm,
                                                         uint32_t n, uint32_t* __restrict__ head) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    uint32_t h = i == 0;
    if (i) {
      const uint32_t a = perm[i], b = perm[i - 1];
      for (uint32_t w = 0; w < num_words && !h; ++w) h = words[(size_t)w * n + a] != words[(size_t)w * n + b];
    }
    head[i] = h;
  }
}
// group_first[g] = original index of the head of group g; is_first[that index] = 1
__global__ void __launch_bounds__(kThreads) group_first_kernel(const uint32_t* __restrict__ head, const uint32_t* __restrict__ group_incl,
                                                               const uint32_t* __restrict__ perm, uint32_t n, uint32_t* __restrict__ group_first,
                                                               uint32_t* __restrict__ is_first) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    if (head[i]) { group_first[group_incl[i] - 1] = perm[i]; is_first[perm[i]] = 1; }
}
// uid[item] = rank (first-occurrence order) of the first occurrence of the item's group; first_index[rank] = that occurrence
__global__ void __launch_bounds__(kThreads) assign_ids_kernel(const uint32_t* __restrict__ group_incl, const uint32_t* __restrict__ perm,
                                                              const uint32_t* __restrict__ group_first, const uint32_t* __restrict__ rank,
                                                              const uint32_t* __restrict__ is_first, uint32_t n, uint32_t* __restrict__ uid,
                                                              uint32_t* __restrict__ first_index) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t first = group_first[group_incl[i] - 1];
    uid[perm[i]] = rank[first];
    if (is_first[i]) first_index[rank[i]] = i;
  }
}

// ---- device dedup ----------------------------------------------------------------------------------------------
struct DeviceBuffer {
  void* p = nullptr;
  cudaStream_t s = nullptr;
  DeviceBuffer(size_t bytes, cudaStream_t stream) : s(stream) { cuda_check(cudaMallocAsync(&p, std::max<size_t>(bytes, 16), s), "cudaMallocAsync"); }
  ~DeviceBuffer() { if (p) cudaFreeAsync(p, s); }
  DeviceBuffer(const DeviceBuffer&) = delete;
  DeviceBuffer& operator=(const DeviceBuffer&) = delete;
  template <class T> T* as() const { return (T*)p; }
};

struct DedupResult {
  std::vector<uint32_t> uid;          // item -> survivor id, first-occurrence order
  std::vector<uint32_t> first_index;  // survivor id -> index of its first occurrence
};

// d_words: num_words x n key words on the device (SoA). Sorts, groups, numbers; copies the result back.
DedupResult dedup_on_device(const uint32_t* d_words, uint32_t num_words, uint32_t n, cudaStream_t s) {
  DedupResult r;
  if (n == 0) return r;
  DeviceBuffer perm_a(4ull * n, s), perm_b(4ull * n, s), key_a(4ull * n, s), key_b(4ull * n, s);
  DeviceBuffer head(4ull * n, s), group_incl(4ull * n, s), group_first(4ull * n, s), is_first(4ull * n, s), rank(4ull * n, s);
  DeviceBuffer uid(4ull * n, s), first_index(4ull * n, s);
  size_t sort_bytes = 0, scan_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)n);
  cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)n);
  DeviceBuffer tmp(std::max(sort_bytes, scan_bytes) + 256, s);
  size_t tmp_bytes = std::max(sort_bytes, scan_bytes) + 256;
  const int g = grid_for(n);
  uint32_t *pa = perm_a.as<uint32_t>(), *pb = perm_b.as<uint32_t>();
  iota_kernel<<<g, kThreads, 0, s>>>(pa, n);
  for (uint32_t w = num_words; w-- > 0;) {  // least significant word first; each pass is stable
    gather_kernel<<<g, kThreads, 0, s>>>(d_words + (size_t)w * n, pa, n, key_a.as<uint32_t>());
    cuda_check(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, key_a.as<uint32_t>(), key_b.as<uint32_t>(), pa, pb, (int)n, 0, 32, s), "cub SortPairs");
    std::swap(pa, pb);
  }
  heads_kernel<<<g, kThreads, 0, s>>>(d_words, num_words, pa, n, head.as<uint32_t>());
  cuda_check(cub::DeviceScan::InclusiveSum(tmp.p, tmp_bytes, head.as<uint32_t>(), group_incl.as<uint32_t>(), (int)n, s), "cub InclusiveSum");
  cuda_check(cudaMemsetAsync(is_first.p, 0, 4ull * n, s), "cudaMemsetAsync");
  group_first_kernel<<<g, kThreads, 0, s>>>(head.as<uint32_t>(), group_incl.as<uint32_t>(), pa, n, group_first.as<uint32_t>(), is_first.as<uint32_t>());
  cuda_check(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, is_first.as<uint32_t>(), rank.as<uint32_t>(), (int)n, s), "cub ExclusiveSum");
  assign_ids_kernel<<<g, kThreads, 0, s>>>(group_incl.as<uint32_t>(), pa, group_first.as<uint32_t>(), rank.as<uint32_t>(), is_first.as<uint32_t>(), n,
                                           uid.as<uint32_t>(), first_index.as<uint32_t>());
  cuda_check(cudaGetLastError(), "dedup kernels");
  uint32_t tail[2] = {0, 0};  // rank[n-1], is_first[n-1]
  cuda_check(cudaMemcpyAsync(&tail[0], rank.as<uint32_t>() + (n - 1), 4, cudaMemcpyDeviceToHost, s), "cudaMemcpyAsync D2H");
  cuda_check(cudaMemcpyAsync(&tail[1], is_first.as<uint32_t>() + (n - 1), 4, cudaMemcpyDeviceToHost, s), "cudaMemcpyAsync D2H");
  r.uid.resize(n);
  cuda_check(cudaMemcpyAsync(r.uid.data(), uid.p, 4ull * n, cudaMemcpyDeviceToHost, s), "cudaMemcpyAsync D2H");
  cuda_check(cudaStreamSynchronize(s), "cudaStreamSynchronize");
  const uint32_t num_unique = tail[0] + tail[1];
  r.first_index.resize(num_unique);
  cuda_check(cudaMemcpyAsync(r.first_index.data(), first_index.p, 4ull * num_unique, cudaMemcpyDeviceToHost, s), "cudaMemcpyAsync D2H");
  cuda_check(cudaStreamSynchronize(s), "cudaStreamSynchronize");
  return r;
}

// Attribute::remove_duplicate_values for n values of one attribute; has_nan reports whether any value holds a NaN.
DedupResult dedup_values(const void* values, uint64_t n64, uint32_t ct, uint32_t ncomp, cudaStream_t s, bool* has_nan) {
  const size_t cb = component_bytes(ct);
  if (cb == 0) throw Error(DXO_ERR_UNSUPPORTED_DATA_TYPE, "unknown component type");
  if (ncomp == 0 || ncomp > 16) throw Error(DXO_ERR_UNSUPPORTED_NUM_COMPONENTS, "component count must be in [1, 16]");
  if (n64 > 0x7FFFFFF0ull) throw Error(DXO_ERR_INVALID_ARGUMENT, "too many values");
  const uint32_t n = (uint32_t)n64, wpc = words_per_component(ct), num_words = ncomp * wpc + 1;
  if (has_nan) *has_nan = false;
  if (n == 0) return {};
  DeviceBuffer d_values((size_t)n * ncomp * cb, s), d_words(4ull * num_words * n, s), d_flag(4, s);
  cuda_check(cudaMemcpyAsync(d_values.p, values, (size_t)n * ncomp * cb, cudaMemcpyHostToDevice, s), "cudaMemcpyAsync H2D");
  cuda_check(cudaMemsetAsync(d_flag.p, 0, 4, s), "cudaMemsetAsync");
  value_keys_kernel<<<grid_for(n), kThreads, 0, s>>>(d_values.as<uint8_t>(), n, ct, ncomp, (uint32_t)cb, wpc, d_words.as<uint32_t>(), d_flag.as<uint32_t>());
  DedupResult r = dedup_on_device(d_words.as<uint32_t>(), num_words, n, s);
  uint32_t flag = 0;
  cuda_check(cudaMemcpyAsync(&flag, d_flag.p, 4, cudaMemcpyDeviceToHost, s), "cudaMemcpyAsync D2H");
  cuda_check(cudaStreamSynchronize(s), "cudaStreamSynchronize");
  if (has_nan) *has_nan = flag != 0;
  return r;
}

// ---- host side of MeshBuilder ------------------------------------------------------------------------------------
struct BuiltAttribute {
  uint32_t att_type, component_type, num_components, domain, unique_id;
  std::vector<uint32_t> parents;
  std::vector<uint8_t> values;   // unique values
  std::vector<uint32_t> map;     // point -> value (when has_map)
  bool has_map = false;
  size_t value_bytes() const { return component_bytes(component_type) * num_components; }
  size_t num_unique() const { return values.size() / value_bytes(); }
  size_t len() const { return has_map ? map.size() : num_unique(); }  // Attribute::len — core/attribute/mod.rs:195-202
  uint32_t value_of(uint32_t point) const { return has_map ? map[point] : point; }
};

// Attribute::remove for a set of points (core/attribute/mod.rs:454-482), batched: `gone[p]` marks removed points.
void remove_points(BuiltAttribute& a, const std::vector<uint8_t>& gone) {
  const size_t vs = a.value_bytes(), L = a.len();
  auto removed = [&](size_t p) { return p < gone.size() && gone[p]; };
  if (a.has_map) {
    std::vector<uint32_t> kept_map(L);
    std::vector<uint8_t> referenced(a.num_unique(), 0);
    size_t kept = 0;
    for (size_t p = 0; p < L; ++p) if (!removed(p)) { kept_map[kept++] = a.map[p]; referenced[a.map[p]] = 1; }
    kept_map.resize(kept);
    std::vector<uint32_t> new_index(referenced.size(), kNone);
    uint32_t next = 0;
    for (size_t v = 0; v < referenced.size(); ++v) if (referenced[v]) new_index[v] = next++;
    if (next != referenced.size()) {  // values no remaining point refers to are dropped, larger indices shift down
      std::vector<uint8_t> kept_values((size_t)next * vs);
      for (size_t v = 0; v < referenced.size(); ++v) if (referenced[v]) memcpy(kept_values.data() + (size_t)new_index[v] * vs, a.values.data() + v * vs, vs);
      a.values.swap(kept_values);
      for (uint32_t& m : kept_map) m = new_index[m];
    }
    a.map.swap(kept_map);
  } else {
    std::vector<uint8_t> kept_values(a.values.size());
    size_t kept = 0;
    for (size_t v = 0; v < L; ++v) if (!removed(v)) { memcpy(kept_values.data() + kept * vs, a.values.data() + v * vs, vs); ++kept; }
    kept_values.resize(kept * vs);
    a.values.swap(kept_values);
  }
}

// ---------------------------------------------------------------------------------------
// Accessor bounds — compute_vec3_bounds / compute_vec4_bounds (io/gltf/encode.rs:815-899): running f32::min / f32::max over
// the values of all points, starting from point 0's value. f32::min / max return the other operand when one is NaN, so NaNs
// are skipped unless every value is NaN. The order of the points matters only for that and for a mix of -0.0 and +0.0,
// where the reference's result depends on how LLVM lowers minnum (unpinned): here -0.0 orders below +0.0.
// Floats are compared through an order-preserving unsigned key; key 0 (below every real key) marks "no value yet".
__device__ __forceinline__ uint32_t bounds_key(float v) {
  const uint32_t b = __float_as_uint(v);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__global__ void __launch_bounds__(256) bounds_kernel(const float* __restrict__ values, const uint32_t* __restrict__ point_to_value, uint64_t num_points,
                                                     uint32_t ncomp, uint32_t* __restrict__ keys /* [0..4) min, [4..8) max, [8..12) non-NaN seen */) {
  uint32_t mn[4], mx[4], any[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) { mn[k] = 0xFFFFFFFFu; mx[k] = 0u; any[k] = 0u; }
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < num_points; p += stride) {
    const uint64_t v = point_to_value ? (uint64_t)__ldg(point_to_value + p) : p;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if ((uint32_t)k >= ncomp) break;
      const float c = __ldg(values + v * ncomp + k);
      if (c == c) { const uint32_t key = bounds_key(c); mn[k] = min(mn[k], key); mx[k] = max(mx[k], key); any[k] = 1u; }
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if ((uint32_t)k >= ncomp) break;
    const uint32_t a = __reduce_min_sync(0xFFFFFFFFu, mn[k]), b = __reduce_max_sync(0xFFFFFFFFu, mx[k]), c = __reduce_or_sync(0xFFFFFFFFu, any[k]);
    if ((threadIdx.x & 31) == 0 && c) { atomicMin(&keys[k], a); atomicMax(&keys[4 + k], b); atomicOr(&keys[8 + k], 1u); }
  }
}

}  // namespace
}  // namespace dxo

using namespace dxo;

struct dxo_built_mesh {
  std::vector<uint32_t> faces;
  std::vector<BuiltAttribute> attributes;
  std::vector<dxo_attribute> views;
};

namespace {
thread_local std::string g_build_error;
template <class F> int guarded_build(F&& f) {
  try { f(); return DXO_OK; }
  catch (const Error& e) { g_build_error = e.what(); return e.status; }
  catch (const std::bad_alloc&) { return DXO_ERR_OUT_OF_MEMORY; }
  catch (const std::exception& e) { g_build_error = e.what(); return DXO_ERR_INTERNAL; }
}

struct BuildClock {  // DXO_TIMING=1: stage times of dxo_mesh_build on stderr
  bool on = getenv("DXO_TIMING") != nullptr;
  std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
  void lap(const char* what) {
    if (!on) return;
    const auto n = std::chrono::steady_clock::now();
    fprintf(stderr, "[dxo] mesh build: %-34s %9.3f ms\n", what, std::chrono::duration<double, std::milli>(n - t).count());
    t = n;
  }
};

struct BuildStream {  // one stream per call; the default memory pool keeps its blocks between calls
  cudaStream_t s = nullptr;
  explicit BuildStream(int device) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) { cudaGetLastError(); throw Error(DXO_ERR_NO_DEVICE, "no CUDA device available (this path has no CPU fallback)"); }
    if (device < 0) cuda_check(cudaGetDevice(&device), "cudaGetDevice");
    if (device >= count) throw Error(DXO_ERR_NO_DEVICE, "CUDA device ordinal out of range");
    cuda_check(cudaSetDevice(device), "cudaSetDevice");
    // keep freed blocks in the stream-ordered pool (as DeviceContext does): a build allocates ~a dozen arrays per dedup
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      uint64_t keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cuda_check(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking), "cudaStreamCreate");
  }
  ~BuildStream() { if (s) { cudaStreamSynchronize(s); cudaStreamDestroy(s); } }
};
}  // namespace

extern "C" {

int dxo_dedup_values(const void* values, uint64_t n, uint32_t component_type, uint32_t num_components, int device, uint32_t* out_map,
                     uint32_t* out_first_index, uint64_t* out_num_unique) {
  if ((!values && n) || !out_map || !out_first_index || !out_num_unique) return DXO_ERR_INVALID_ARGUMENT;
  return guarded_build([&] {
    BuildStream bs(device);
    DedupResult r = dedup_values(values, n, component_type, num_components, bs.s, nullptr);
    if (n) memcpy(out_map, r.uid.data(), 4 * (size_t)n);
    if (!r.first_index.empty()) memcpy(out_first_index, r.first_index.data(), 4 * r.first_index.size());
    *out_num_unique = r.first_index.size();
  });
}

int dxo_attribute_bounds(const float* values, uint64_t num_values, uint32_t num_components, const uint32_t* point_to_value, uint64_t num_points,
                         int device, float* out_min, float* out_max) {
  if (!out_min || !out_max || num_components == 0 || num_components > 4 || (!values && num_values)) return DXO_ERR_INVALID_ARGUMENT;
  if (!point_to_value) num_points = num_values;
  if (num_points == 0) return DXO_OK;  // the reference returns empty vectors: outputs untouched
  if (num_values == 0) return DXO_ERR_INVALID_ARGUMENT;
  return guarded_build([&] {
    if (point_to_value)
      for (uint64_t p = 0; p < num_points; ++p)
        if (point_to_value[p] >= num_values) throw Error(DXO_ERR_INVALID_ARGUMENT, "point_to_value entry out of range");
    BuildStream bs(device);
    float* d_values = nullptr; uint32_t *d_map = nullptr, *d_keys = nullptr;
    const size_t vb = (size_t)num_values * num_components * sizeof(float);
    cuda_check(cudaMallocAsync((void**)&d_values, vb, bs.s), "cudaMallocAsync");
    cuda_check(cudaMemcpyAsync(d_values, values, vb, cudaMemcpyHostToDevice, bs.s), "cudaMemcpyAsync H2D");
    if (point_to_value) {
      cuda_check(cudaMallocAsync((void**)&d_map, (size_t)num_points * 4, bs.s), "cudaMallocAsync");
      cuda_check(cudaMemcpyAsync(d_map, point_to_value, (size_t)num_points * 4, cudaMemcpyHostToDevice, bs.s), "cudaMemcpyAsync H2D");
    }
    uint32_t init[12];
    for (int k = 0; k < 4; ++k) { init[k] = 0xFFFFFFFFu; init[4 + k] = 0u; init[8 + k] = 0u; }
    cuda_check(cudaMallocAsync((void**)&d_keys, sizeof(init), bs.s), "cudaMallocAsync");
    cuda_check(cudaMemcpyAsync(d_keys, init, sizeof(init), cudaMemcpyHostToDevice, bs.s), "cudaMemcpyAsync H2D");
    const uint64_t want = (num_points + 256ull * 4 - 1) / (256ull * 4);
    const int grid = (int)std::min<uint64_t>(std::max<uint64_t>(want, 1), 148ull * 8);
    bounds_kernel<<<grid, 256, 0, bs.s>>>(d_values, d_map, num_points, num_components, d_keys);
    cuda_check(cudaGetLastError(), "bounds_kernel");
    uint32_t keys[12];
    cuda_check(cudaMemcpyAsync(keys, d_keys, sizeof(keys), cudaMemcpyDeviceToHost, bs.s), "cudaMemcpyAsync D2H");
    cuda_check(cudaStreamSynchronize(bs.s), "cudaStreamSynchronize");
    cudaFreeAsync(d_values, bs.s); if (d_map) cudaFreeAsync(d_map, bs.s); cudaFreeAsync(d_keys, bs.s);
    auto unkey = [](uint32_t k) { const uint32_t b = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k; float f; memcpy(&f, &b, 4); return f; };
    for (uint32_t k = 0; k < num_components; ++k) {
      if (keys[8 + k]) { out_min[k] = unkey(keys[k]); out_max[k] = unkey(keys[4 + k]); }
      else out_min[k] = out_max[k] = std::numeric_limits<float>::quiet_NaN();  // every value of the component is NaN
    }
  });
}

int dxo_mesh_build(const uint32_t* faces, uint64_t num_faces, const dxo_attribute* per_point_attributes, uint32_t num_attributes, int device,
                   dxo_built_mesh** out) {
  if (!out) return DXO_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  if ((!faces && num_faces) || (!per_point_attributes && num_attributes)) return DXO_ERR_INVALID_ARGUMENT;
  std::unique_ptr<dxo_built_mesh> mesh(new dxo_built_mesh);
  const int st = guarded_build([&] {
    if (num_faces > 0x55555554ull) throw Error(DXO_ERR_INVALID_ARGUMENT, "too many faces");
    BuildStream bs(device);
    BuildClock clk;
    std::vector<BuiltAttribute>& atts = mesh->attributes;
    // ---- add_attribute: Attribute::from with value dedup (builder.rs:30-39, attribute/mod.rs:87-103, 394-452) ----
    uint64_t L = 0;
    std::vector<uint8_t> att_has_nan(num_attributes, 0);
    for (uint32_t i = 0; i < num_attributes; ++i) {
      const dxo_attribute& in = per_point_attributes[i];
      if (!in.values && in.num_unique_values) throw Error(DXO_ERR_INVALID_ARGUMENT, "attribute without values");
      if (i == 0) L = in.num_unique_values;
      else if (in.num_unique_values != L)  // ragged attributes take reference paths this builder does not restate
        throw Error(DXO_ERR_UNSUPPORTED_INPUT, "all attributes must hold one value per point (equal counts)");
      BuiltAttribute a;
      a.att_type = in.att_type; a.component_type = in.component_type; a.num_components = in.num_components; a.domain = in.domain;
      a.unique_id = i;  // MeshBuilder numbers attributes in insertion order (builder.rs:30-39)
      a.parents.assign(in.parent_ids, in.parent_ids + in.num_parents);
      bool has_nan = false;
      DedupResult r = dedup_values(in.values, in.num_unique_values, in.component_type, in.num_components, bs.s, &has_nan);
      clk.lap("attribute value dedup (device)");
      att_has_nan[i] = has_nan ? 1 : 0;
      const size_t vs = a.value_bytes();
      const uint8_t* src = (const uint8_t*)in.values;
      if (r.first_index.size() == in.num_unique_values) {
        a.values.assign(src, src + (size_t)in.num_unique_values * vs);  // no duplicates: the attribute keeps its buffer, no map
      } else {
        a.has_map = true;
        a.map.swap(r.uid);
        a.values.resize(r.first_index.size() * vs);
        for (size_t u = 0; u < r.first_index.size(); ++u) memcpy(a.values.data() + u * vs, src + (size_t)r.first_index[u] * vs, vs);
      }
      atts.push_back(std::move(a));
      clk.lap("attribute unique values (host)");
    }
    // ---- build(): dependency_check (builder.rs:94-111) ----
    for (const BuiltAttribute& a : atts) {
      if (a.att_type != DXO_ATT_TEXCOORD) continue;
      bool ok = false;
      for (uint32_t pid : a.parents) for (const BuiltAttribute& b : atts) if (b.unique_id == pid && b.att_type == DXO_ATT_POSITION) ok = true;
      if (!ok) throw Error(DXO_ERR_INVALID_ARGUMENT, "MinimumDependencyError: TextureCoordinate must depend on Position");
    }
    // ---- get_sorted_attributes (:115-125): the first position attribute moves to the front ----
    for (size_t i = 0; i < atts.size(); ++i) if (atts[i].att_type == DXO_ATT_POSITION) { std::swap(atts[0], atts[i]); std::swap(att_has_nan[0], att_has_nan[i]); break; }
    std::vector<uint32_t>& f = mesh->faces;
    f.assign(faces, faces + 3 * num_faces);
    // ---- deduplicate points (:194-373): points whose values agree in every attribute merge, first occurrence wins ----
    if (!atts.empty() && !f.empty()) {
      uint32_t max_idx = 0;
      for (uint32_t p : f) max_idx = std::max(max_idx, p);
      if ((uint64_t)max_idx >= L) throw Error(DXO_ERR_INVALID_ARGUMENT, "face references a point outside the attributes");
      const uint32_t num_vertices = max_idx + 1;
      // key of a point: per attribute the id of its unique value — or, for an attribute holding NaNs, the raw bytes of that
      // value (the reference hashes value bytes here, so NaNs with equal bit patterns merge although they are distinct values)
      uint32_t num_words = 0;
      for (size_t a = 0; a < atts.size(); ++a) num_words += att_has_nan[a] ? (uint32_t)((atts[a].value_bytes() + 3) / 4) : 1u;
      DeviceBuffer d_words(4ull * num_words * num_vertices, bs.s);
      uint32_t w0 = 0;
      std::vector<uint32_t> tmp_words;
      for (size_t a = 0; a < atts.size(); ++a) {
        const BuiltAttribute& A = atts[a];
        uint32_t* dst = d_words.as<uint32_t>() + (size_t)w0 * num_vertices;
        if (!att_has_nan[a]) {
          if (A.has_map) cuda_check(cudaMemcpyAsync(dst, A.map.data(), 4ull * num_vertices, cudaMemcpyHostToDevice, bs.s), "cudaMemcpyAsync H2D");
          else iota_kernel<<<grid_for(num_vertices), kThreads, 0, bs.s>>>(dst, num_vertices);
          w0 += 1;
        } else {
          const size_t vs = A.value_bytes(), nw = (vs + 3) / 4;
          tmp_words.assign(nw * num_vertices, 0);
          for (uint32_t p = 0; p < num_vertices; ++p) {
            uint32_t tmp[32] = {0};
            memcpy(tmp, A.values.data() + (size_t)A.value_of(p) * vs, vs);
            for (size_t k = 0; k < nw; ++k) tmp_words[k * num_vertices + p] = tmp[k];
          }
          cuda_check(cudaMemcpyAsync(dst, tmp_words.data(), 4ull * nw * num_vertices, cudaMemcpyHostToDevice, bs.s), "cudaMemcpyAsync H2D");
          cuda_check(cudaStreamSynchronize(bs.s), "cudaStreamSynchronize");  // tmp_words is reused
          w0 += (uint32_t)nw;
        }
      }
      clk.lap("point keys (host)");
      DedupResult pm = dedup_on_device(d_words.as<uint32_t>(), num_words, num_vertices, bs.s);
      clk.lap("point dedup (device)");
      const uint32_t unique_count = (uint32_t)pm.first_index.size();
      if (unique_count != num_vertices) {
        std::vector<uint8_t> gone(num_vertices, 0);
        for (uint32_t v = 0; v < num_vertices; ++v) gone[v] = pm.first_index[pm.uid[v]] != v;  // later duplicates
        std::vector<std::future<void>> jobs;  // the attributes are independent
        for (BuiltAttribute& A : atts) {
          if (unique_count == A.len()) continue;  // remap_attribute's early return (:274-276)
          jobs.push_back(std::async(std::launch::async, [&A, &gone] { remove_points(A, gone); }));
        }
        for (auto& j : jobs) j.get();
        for (uint32_t& p : f) p = pm.uid[p];
      }
      clk.lap("point removal + face remap (host)");
    }
    // ---- degenerate faces (:76-79) ----
    {
      size_t k = 0;
      for (size_t i = 0; i + 2 < f.size(); i += 3)
        if (f[i] != f[i + 1] && f[i + 1] != f[i + 2] && f[i + 2] != f[i]) { f[k] = f[i]; f[k + 1] = f[i + 1]; f[k + 2] = f[i + 2]; k += 3; }
      f.resize(k);
    }
    // ---- remove_unused_vertices (:129-189) ----
    if (!f.empty() && !atts.empty()) {
      uint32_t max_idx = 0;
      for (uint32_t p : f) max_idx = std::max(max_idx, p);
      std::vector<uint8_t> used((size_t)max_idx + 1, 0);
      for (uint32_t p : f) used[p] = 1;
      bool all_used = true;
      for (uint8_t u : used) all_used &= u != 0;
      for (BuiltAttribute& A : atts) {
        if (all_used && A.len() == used.size()) continue;  // nothing to remove from this attribute
        std::vector<uint8_t> gone(A.len(), 0);
        for (size_t p = 0; p < gone.size(); ++p) gone[p] = p > max_idx || !used[p];
        remove_points(A, gone);
      }
      std::vector<uint32_t> offset(used.size());
      uint32_t removed = 0;
      for (size_t v = 0; v < used.size(); ++v) { offset[v] = removed; removed += !used[v]; }
      for (uint32_t& p : f) p -= offset[p];
    }
    clk.lap("degenerate faces, unused points");
    // ---- views ----
    mesh->views.resize(atts.size());
    for (size_t i = 0; i < atts.size(); ++i) {
      BuiltAttribute& A = atts[i];
      dxo_attribute& v = mesh->views[i];
      v.att_type = A.att_type; v.component_type = A.component_type; v.num_components = A.num_components; v.domain = A.domain;
      v.unique_id = A.unique_id; v.num_parents = (uint32_t)A.parents.size(); v.parent_ids = A.parents.data();
      v.num_unique_values = A.num_unique(); v.values = A.values.data();
      v.num_points = A.len(); v.point_to_value = A.has_map ? A.map.data() : nullptr;
    }
  });
  if (st == DXO_OK) *out = mesh.release();
  return st;
}

int dxo_built_mesh_view(const dxo_built_mesh* mesh, dxo_mesh* out) {
  if (!mesh || !out) return DXO_ERR_INVALID_ARGUMENT;
  out->num_faces = mesh->faces.size() / 3;
  out->faces = mesh->faces.data();
  out->num_attributes = (uint32_t)mesh->views.size();
  out->attributes = mesh->views.data();
  return DXO_OK;
}

void dxo_built_mesh_free(dxo_built_mesh* mesh) { delete mesh; }

}  // extern "C"
