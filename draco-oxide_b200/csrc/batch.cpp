// Product host code — the batch entry: many meshes through ONE launch set per group.
//
// Replaces the transcoder's per-primitive loop (io/gltf/encode.rs:932-955, one encode() per primitive). The
// per-mesh path (encoder.cpp) costs ~80 launches, ~60 stream-ordered allocations and >= 4 synchronisations per
// mesh, which is what bounds a batch of small meshes. Here the meshes are cut into GROUPS (longest first); a group's
// arrays are laid out back to back in one device slab and one pinned host slab, and every stage is one segmented
// launch over the whole group (kernels.cuh "Segmented launches"):
//
//   stage 0  host workers   copy faces / values / maps into the pinned slab, range checks          (per mesh, parallel)
//   stage 1  device         H2D, K12 + K13 (one radix sort for the group), K14, D2H of the tables   (one launch set)
//   stage 2  host workers   Edgebreaker traversal, sequencer, connectivity bytes                    (per mesh, parallel)
//   stage 3  device         H2D of the sequences, K1-K10, side streams (device rABS), packing, D2H  (one launch set)
//   stage 4  host workers   stream assembly                                                          (per mesh, parallel)
//
// A group is driven by the thread of the slot it occupies (a slot = device slab + pinned slab + stream); several slots
// per GPU keep stages of different groups overlapped: while one group's meshes are walked on the host, another's
// kernels run. Meshes the device passes flag (non-manifold, inconsistent orientation, unused vertices, ...) take the
// per-mesh path, whose sequential host passes are the reference's algorithms. No CPU implementation of the
// attribute kernels exists here either: without a device the call fails.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <deque>
#include <functional>
#include <mutex>
#include <numeric>
#include <thread>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "encoder.hpp"

namespace dxo {

namespace {

using Clock = std::chrono::steady_clock;
inline double ms_since(Clock::time_point t0) { return std::chrono::duration<double, std::milli>(Clock::now() - t0).count(); }
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Fixed pool of host workers shared by every group of a batch call.
class Workers {
 public:
  explicit Workers(int n) { for (int i = 0; i < n; ++i) threads_.emplace_back([this] { loop(); }); }
  ~Workers() {
    { std::lock_guard<std::mutex> lock(mu_); stop_ = true; }
    cv_.notify_all();
    for (std::thread& t : threads_) t.join();
  }
  // runs fn(0) .. fn(count - 1) on the pool and returns when all are done (the caller sleeps meanwhile)
  void parallel_for(size_t count, const std::function<void(size_t)>& fn) {
    if (count == 0) return;
    struct Latch { std::mutex mu; std::condition_variable cv; size_t left; } latch;
    latch.left = count;
    {
      std::lock_guard<std::mutex> lock(mu_);
      for (size_t i = 0; i < count; ++i)
        queue_.push_back([&fn, &latch, i] {
          fn(i);
          std::lock_guard<std::mutex> l(latch.mu);
          if (--latch.left == 0) latch.cv.notify_one();
        });
    }
    cv_.notify_all();
    std::unique_lock<std::mutex> l(latch.mu);
    latch.cv.wait(l, [&] { return latch.left == 0; });
  }
 private:
  void loop() {
    for (;;) {
      std::function<void()> task;
      {
        std::unique_lock<std::mutex> lock(mu_);
        cv_.wait(lock, [this] { return stop_ || !queue_.empty(); });
        if (queue_.empty()) return;
        task = std::move(queue_.front());
        queue_.pop_front();
      }
      task();
    }
  }
  std::mutex mu_;
  std::condition_variable cv_;
  std::deque<std::function<void()>> queue_;
  std::vector<std::thread> threads_;
  bool stop_ = false;
};

// Bump allocator over one block. The "pair" slab exists twice with identical layout — in device memory and in
// pinned host memory — so that an array that crosses the bus has the same offset on both sides and a whole stage's
// traffic is one copy of a contiguous range.
struct Slab {
  uint8_t* base = nullptr;
  size_t cap = 0, used = 0;
  size_t take(size_t bytes, size_t align = 256) {
    used = align_up(used, align);
    const size_t off = used;
    used += bytes;
    if (used > cap) throw Error(DXO_ERR_INTERNAL, "group slab too small");
    return off;
  }
};

struct Slot {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev = nullptr;
  uint8_t* d_pair = nullptr; uint8_t* h_pair = nullptr; size_t pair_cap = 0, h_block_cap = 0;
  uint8_t* d_only = nullptr; size_t only_cap = 0;
  void ensure(size_t pair_bytes, size_t only_bytes) {
    cuda_check(cudaSetDevice(device), "cudaSetDevice");
    if (!stream) {
      cuda_check(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking), "cudaStreamCreate");
      cuda_check(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming | cudaEventBlockingSync), "cudaEventCreate");
    }
    if (pair_bytes > pair_cap) {
      if (d_pair) { cuda_check(cudaStreamSynchronize(stream), "sync"); cuda_check(cudaFree(d_pair), "cudaFree"); d_pair = nullptr; }
      if (h_pair) { pinned_block_give(h_pair, h_block_cap); h_pair = nullptr; }
      pair_cap = pair_bytes + pair_bytes / 8;
      cuda_check(cudaMalloc((void**)&d_pair, pair_cap), "cudaMalloc (group slab)");
      h_pair = (uint8_t*)pinned_block_take(pair_cap, &h_block_cap);
      if (!h_pair) throw Error(DXO_ERR_OUT_OF_MEMORY, "pinned host allocation failed");
    }
    if (only_bytes > only_cap) {
      if (d_only) { cuda_check(cudaStreamSynchronize(stream), "sync"); cuda_check(cudaFree(d_only), "cudaFree"); d_only = nullptr; }
      only_cap = only_bytes + only_bytes / 8;
      cuda_check(cudaMalloc((void**)&d_only, only_cap), "cudaMalloc (group scratch)");
    }
  }
  void wait() {
    cuda_check(cudaEventRecord(ev, stream), "cudaEventRecord");
    cuda_check(cudaEventSynchronize(ev), "cudaEventSynchronize");
  }
  ~Slot() {
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return; }
    if (stream) { cudaStreamSynchronize(stream); cudaStreamDestroy(stream); }
    if (ev) cudaEventDestroy(ev);
    if (d_pair) cudaFree(d_pair);
    if (d_only) cudaFree(d_only);
    if (h_pair) pinned_block_give(h_pair, h_block_cap);
    cudaGetLastError();
  }
};

// Slots outlive the call that created them: a slot's slabs are a gigabyte of device memory and hundreds of megabytes of
// pinned memory, and cudaMalloc / cudaFree of that size cost milliseconds and synchronise the device. Idle slots wait here
// (process lifetime, like the pinned pool).
struct SlotPool {
  std::mutex mu;
  std::vector<Slot*> idle;
  Slot* acquire(int device) {
    {
      std::lock_guard<std::mutex> lock(mu);
      for (size_t k = 0; k < idle.size(); ++k)
        if (idle[k]->device == device) { Slot* s = idle[k]; idle.erase(idle.begin() + k); return s; }
    }
    Slot* s = new Slot;
    s->device = device;
    return s;
  }
  void release(Slot* s) { std::lock_guard<std::mutex> lock(mu); idle.push_back(s); }
};
SlotPool& slot_pool() { static SlotPool* p = new SlotPool; return *p; }

// DXO_TIMING: CPU time of the host stages, summed over all workers of a batch call
struct HostClock {
  std::atomic<uint64_t> ns[8];
  const char* name[8] = {"stage 0 copy inputs", "stage 2 setup (tables, masks)", "stage 2 Edgebreaker traversal", "stage 2 connectivity bytes",
                         "stage 2 seam streams", "stage 2 sequencers", "stage 4 side streams (host)", "stage 4 assembly"};
  void reset() { for (auto& v : ns) v.store(0); }
  void report(uint64_t vertices) {
    uint64_t total = 0;
    for (auto& v : ns) total += v.load();
    for (int k = 0; k < 8; ++k) fprintf(stderr, "[dxo] host cpu  %-34s %9.3f ms  %6.1f ns/vertex\n", name[k], ns[k].load() * 1e-6, (double)ns[k].load() / (double)std::max<uint64_t>(vertices, 1));
    fprintf(stderr, "[dxo] host cpu  %-34s %9.3f ms  %6.1f ns/vertex\n", "total", total * 1e-6, (double)total / (double)std::max<uint64_t>(vertices, 1));
  }
};
HostClock g_host_clock;
struct HostLap {
  Clock::time_point t = Clock::now();
  void lap(int k) { const auto n = Clock::now(); g_host_clock.ns[k].fetch_add((uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(n - t).count(), std::memory_order_relaxed); t = n; }
};

// Copies into the pinned slab are read next by the copy engine, never by this core: with AVX2 they go out as
// non-temporal stores (no read-for-ownership, nothing evicted from the caches the traversal works in) and the index
// maximum comes from the same registers. Measured per thread: 2.5 GB/s (scalar copy + max) -> 7.7 GB/s. The library is
// built without -march flags (it travels between machines), so the AVX2 bodies are per-function targets picked at run time.
#if defined(__x86_64__)
__attribute__((target("avx2"))) uint32_t copy_max_u32_avx2(uint32_t* dst, const uint32_t* src, size_t n) {
  size_t i = 0;
  uint32_t mx = 0;
  while (i < n && ((uintptr_t)(dst + i) & 31)) { dst[i] = src[i]; mx = std::max(mx, src[i]); ++i; }
  __m256i m0 = _mm256_setzero_si256(), m1 = m0, m2 = m0, m3 = m0;
  for (; i + 32 <= n; i += 32) {
    const __m256i a = _mm256_loadu_si256((const __m256i*)(src + i)), b = _mm256_loadu_si256((const __m256i*)(src + i + 8));
    const __m256i c = _mm256_loadu_si256((const __m256i*)(src + i + 16)), d = _mm256_loadu_si256((const __m256i*)(src + i + 24));
    _mm256_stream_si256((__m256i*)(dst + i), a); _mm256_stream_si256((__m256i*)(dst + i + 8), b);
    _mm256_stream_si256((__m256i*)(dst + i + 16), c); _mm256_stream_si256((__m256i*)(dst + i + 24), d);
    m0 = _mm256_max_epu32(m0, a); m1 = _mm256_max_epu32(m1, b); m2 = _mm256_max_epu32(m2, c); m3 = _mm256_max_epu32(m3, d);
  }
  m0 = _mm256_max_epu32(_mm256_max_epu32(m0, m1), _mm256_max_epu32(m2, m3));
  alignas(32) uint32_t t[8];
  _mm256_store_si256((__m256i*)t, m0);
  for (int k = 0; k < 8; ++k) mx = std::max(mx, t[k]);
  for (; i < n; ++i) { dst[i] = src[i]; mx = std::max(mx, src[i]); }
  _mm_sfence();
  return mx;
}
__attribute__((target("avx2"))) void copy_stream_avx2(uint8_t* dst, const uint8_t* src, size_t n) {
  size_t i = 0;
  while (i < n && ((uintptr_t)(dst + i) & 31)) { dst[i] = src[i]; ++i; }
  for (; i + 128 <= n; i += 128) {
    const __m256i a = _mm256_loadu_si256((const __m256i*)(src + i)), b = _mm256_loadu_si256((const __m256i*)(src + i + 32));
    const __m256i c = _mm256_loadu_si256((const __m256i*)(src + i + 64)), d = _mm256_loadu_si256((const __m256i*)(src + i + 96));
    _mm256_stream_si256((__m256i*)(dst + i), a); _mm256_stream_si256((__m256i*)(dst + i + 32), b);
    _mm256_stream_si256((__m256i*)(dst + i + 64), c); _mm256_stream_si256((__m256i*)(dst + i + 96), d);
  }
  if (i < n) memcpy(dst + i, src + i, n - i);
  _mm_sfence();
}
bool have_avx2() { static const bool v = __builtin_cpu_supports("avx2") && !getenv("DXO_NO_AVX2"); return v; }
#else
bool have_avx2() { return false; }
#endif

// copy + maximum in one pass
uint32_t copy_max_u32(uint32_t* dst, const uint32_t* src, size_t n) {
#if defined(__x86_64__)
  if (have_avx2()) return copy_max_u32_avx2(dst, src, n);
#endif
  uint32_t m[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  size_t i = 0;
  for (; i + 8 <= n; i += 8)
    for (int k = 0; k < 8; ++k) { const uint32_t v = src[i + k]; dst[i + k] = v; m[k] = v > m[k] ? v : m[k]; }
  uint32_t mx = 0;
  for (; i < n; ++i) { dst[i] = src[i]; mx = std::max(mx, src[i]); }
  for (int k = 0; k < 8; ++k) mx = std::max(mx, m[k]);
  return mx;
}
// plain copy into the pinned slab
void copy_to_pinned(void* dst, const void* src, size_t bytes) {
#if defined(__x86_64__)
  if (have_avx2() && bytes >= 4096) { copy_stream_avx2((uint8_t*)dst, (const uint8_t*)src, bytes); return; }
#endif
  memcpy(dst, src, bytes);
}

void add_tiles(std::vector<gpu::Tile>& v, uint32_t seg, uint64_t count, uint32_t tile = gpu::kSegTile) {
  for (uint64_t f = 0; f < count; f += tile) v.push_back({seg, (uint32_t)f});
}

}  // namespace

// One mesh of a group: the per-mesh job (validation, plans, host connectivity) plus where its arrays live in the slabs.
struct GroupMesh {
  size_t index = 0;  // in the caller's array
  const dxo_mesh* mesh = nullptr;
  std::unique_ptr<MeshJob> job;
  int status = DXO_OK;
  std::string error;
  bool alive = true;      // still on the group path
  bool fallback = false;  // flagged by the device passes: encoded by the per-mesh path
  uint32_t C = 0, F = 0, V = 0;
  // offsets into the pair slab
  size_t faces = 0, cv = 0, opposite = 0, left_most = 0, interior = 0, flags = 0;
  struct Att {
    size_t values = 0, map = 0;  // inputs (map == 0 && !has_map: identity)
    bool has_map = false;
    size_t value_bytes = 0;
    // seam table (non-position attributes)
    size_t seam = 0, cv_a = 0, left_most_a = 0, interior_a = 0, scalars = 0;
    uint32_t capacity = 0, count_base = 0;
    size_t vertex_on_seam = 0;  // device-only slab
    int seam_seg = -1;
    // stage 3
    int stream = -1;        // index of this attribute's AttrSeg / RansJob
    size_t seq = 0;         // pair slab (when the attribute has its own sequence)
    size_t host_flags = 0;  // pair slab: K5 / K6 flags of a side stream that the host codes (long streams)
    bool host_side = false;
  };
  std::vector<Att> atts;
  // device-only offsets of stage 1
  size_t first_corner = 0, valence = 0;
  uint32_t corner_base = 0;   // first slot in the group-wide corner order (sort slots, opposite array, boundary list)
};

struct GroupSizes { uint64_t sumC = 0, sum_seamV = 0; uint32_t vertex_bits = 1; size_t sort_bytes = 0, scan_bytes = 0, pair_bytes = 0, only_bytes = 0; };

class GroupRunner {
 public:
  GroupRunner(Slot& slot, Workers& workers, const dxo_config& cfg, dxo_bytes* outs, size_t min_pair = 0, size_t min_only = 0)
      : slot_(slot), workers_(workers), cfg_(cfg), outs_(outs), min_pair_(min_pair), min_only_(min_only) {}
  void run(std::vector<GroupMesh*>& meshes);
  // also fills the per-mesh sizes (C, F, V, attribute capacities)
  static GroupSizes measure(std::vector<GroupMesh*>& meshes);

 private:
  Slot& slot_;
  Workers& workers_;
  dxo_config cfg_;
  dxo_bytes* outs_;
  size_t min_pair_, min_only_;  // the largest group of the batch: slabs are sized once, not regrown group by group
  Slab pair_, only_;
  template <class T> T* dp(size_t off) const { return (T*)(slot_.d_pair + off); }
  template <class T> T* hp(size_t off) const { return (T*)(slot_.h_pair + off); }
  template <class T> T* dn(size_t off) const { return (T*)(slot_.d_only + off); }
  void h2d(size_t begin, size_t end) {
    if (end > begin) cuda_check(cudaMemcpyAsync(slot_.d_pair + begin, slot_.h_pair + begin, end - begin, cudaMemcpyHostToDevice, slot_.stream), "cudaMemcpyAsync H2D");
  }
  void d2h(size_t begin, size_t end) {
    if (end > begin) cuda_check(cudaMemcpyAsync(slot_.h_pair + begin, slot_.d_pair + begin, end - begin, cudaMemcpyDeviceToHost, slot_.stream), "cudaMemcpyAsync D2H");
  }
  template <class T> size_t put_vector(const std::vector<T>& v) {  // host vector -> pair slab (uploaded with its region)
    const size_t off = pair_.take(std::max<size_t>(v.size(), 1) * sizeof(T));
    if (!v.empty()) memcpy(hp<uint8_t>(off), v.data(), v.size() * sizeof(T));
    return off;
  }
  static void fail(GroupMesh& m, int status, const std::string& what) { m.alive = false; m.status = status; m.error = what; }
  template <class F> static void guarded(GroupMesh& m, F&& f) {
    try { f(); }
    catch (const Error& e) { fail(m, e.status, e.what()); }
    catch (const std::bad_alloc&) { fail(m, DXO_ERR_OUT_OF_MEMORY, "out of host memory"); }
    catch (const std::exception& e) { fail(m, DXO_ERR_INTERNAL, e.what()); }
  }
  void stage0_copy_inputs(GroupMesh& m);
  void stage2_host_connectivity(GroupMesh& m);
  void encode_fallback(GroupMesh& m);
  void give_bytes(GroupMesh& m, std::vector<uint8_t>& bytes);
};

void GroupRunner::give_bytes(GroupMesh& m, std::vector<uint8_t>& bytes) {
  dxo_bytes& out = outs_[m.index];
  out.data = (uint8_t*)malloc(bytes.size() ? bytes.size() : 1);
  if (!out.data) throw Error(DXO_ERR_OUT_OF_MEMORY, "out of memory");
  memcpy(out.data, bytes.data(), bytes.size());
  out.len = bytes.size();
}

void GroupRunner::encode_fallback(GroupMesh& m) {
  guarded(m, [&] {
    dxo_config c = cfg_;
    c.device = slot_.device;
    std::vector<uint8_t> bytes;
    dxo_timing tm;
    encode_one_mesh(m.mesh, c, bytes, tm, false);
    give_bytes(m, bytes);
    m.status = DXO_OK;
  });
  m.alive = false;
}

// faces, values and maps into the pinned slab, with the range checks the kernels rely on
void GroupRunner::stage0_copy_inputs(GroupMesh& m) {
  HostLap hl;
  guarded(m, [&] {
    const uint32_t* faces = m.mesh->faces;
    const uint32_t max_p = copy_max_u32(hp<uint32_t>(m.faces), faces, m.C);
    MeshJob& job = *m.job;
    for (size_t i = 0; i < job.plans_.size(); ++i) {
      const AttrView& v = job.plans_[i].view;
      if (v.num_points <= max_p) throw Error(DXO_ERR_INVALID_ARGUMENT, i == 0 ? "face references a point outside the position attribute" : "face references a point outside an attribute");
      copy_to_pinned(hp<uint8_t>(m.atts[i].values), job.plans_[i].values32, m.atts[i].value_bytes);
      if (v.map) {
        const uint32_t mx = copy_max_u32(hp<uint32_t>(m.atts[i].map), v.map, v.num_points);
        if (v.num_points && mx >= v.num_unique) throw Error(DXO_ERR_INVALID_ARGUMENT, "point_to_value entry out of range");
      }
    }
  });
  hl.lap(0);
}

// The tables of K12-K14 have landed in the pinned slab: wrap them as the job's universal / seam tables and run what
// north_star keeps on the host — the Edgebreaker traversal, the sequencers, the connectivity bytes.
void GroupRunner::stage2_host_connectivity(GroupMesh& m) {
  const uint32_t flags = *hp<uint32_t>(m.flags);
  MeshJob& job = *m.job;
  const size_t natt = job.plans_.size();
  bool flagged = flags != 0;
  for (size_t i = 1; i < natt && !flagged; ++i) {
    const uint32_t* sc = hp<uint32_t>(m.atts[i].scalars);
    if ((sc[1] & (1u | 2u | 8u)) || sc[0] > m.atts[i].capacity || sc[0] < m.V) flagged = true;
  }
  if (flagged) {
    if (getenv("DXO_DEBUG")) {
      fprintf(stderr, "[dxo] batch mesh %zu falls back to the per-mesh path: mesh flags %u", m.index, flags);
      for (size_t i = 1; i < natt; ++i) fprintf(stderr, ", att %zu: vertices %u (capacity %u, V %u) flags %u", i, hp<uint32_t>(m.atts[i].scalars)[0], m.atts[i].capacity, m.V, hp<uint32_t>(m.atts[i].scalars)[1]);
      fprintf(stderr, "\n");
    }
    m.fallback = true; encode_fallback(m); return;
  }
  HostLap hl;
  guarded(m, [&] {
    UniversalTable& ut = job.ut_;
    ut.num_faces = m.F; ut.num_corners = m.C; ut.num_vertices = m.V;
    ut.corner_point = m.mesh->faces;
    // corner -> vertex: the faces themselves without a position map (never written on this path), else map[faces] from stage 0
    if (!m.atts[0].has_map) ut.corner_vertex.adopt(const_cast<uint32_t*>(m.mesh->faces), m.C);
    else ut.corner_vertex.adopt(hp<uint32_t>(m.cv), m.C);
    ut.opposite.adopt(hp<uint32_t>(m.opposite), m.C);
    ut.left_most.adopt(hp<uint32_t>(m.left_most), m.V);
    ut.matched_on_device = true;
    job.write_stream_header();
    job.seams_.resize(natt - 1);
    job.table_refs_.assign(natt, TableRef{});
    job.interior_.assign(natt, {});
    job.masked_opposite_.clear();
    job.masked_opposite_.resize(natt);
    job.table_refs_[0] = table_ref(ut);
    job.table_refs_[0].interior = hp<uint8_t>(m.interior);
    for (size_t i = 1; i < natt; ++i) {
      const GroupMesh::Att& a = m.atts[i];
      const uint32_t* sc = hp<uint32_t>(a.scalars);
      SeamTable& st = job.seams_[i - 1];
      st.num_vertices = sc[0];
      st.has_interior_seam = (sc[1] & 4u) != 0;
      if (!st.has_interior_seam && st.num_vertices == m.V) continue;  // same table as the universal one: nothing was fetched, nothing is needed
      st.corner_vertex.adopt(hp<uint32_t>(a.cv_a), m.C);
      st.seam.adopt(hp<uint8_t>(a.seam), m.C);
      st.left_most.adopt(hp<uint32_t>(a.left_most_a), st.num_vertices);
      job.table_refs_[i] = table_ref(ut, st);
      job.table_refs_[i].interior = hp<uint8_t>(a.interior_a);
      if (st.has_interior_seam) {  // the sequencer of this table runs: one array for opp()
        U32Array& mo = job.masked_opposite_[i];
        mo.resize(m.C);
        const uint32_t* opp = ut.opposite.data();
        const uint8_t* sm = st.seam.data();
        for (uint32_t k = 0; k < m.C; ++k) mo[k] = sm[k] ? kNone : opp[k];
        job.table_refs_[i].opposite_masked = mo.data();
      }
    }
    hl.lap(1);
    job.eb_.reset(new EdgebreakerEncoder(ut));
    EdgebreakerEncoder& eb = *job.eb_;
    eb.traverse();
    hl.lap(2);
    eb.write_head(job.head_, job.seams_.size());
    hl.lap(3);
    for (size_t i = 1; i < natt; ++i) {
      ByteSink sb;
      eb.write_seam_stream(job.seams_[i - 1], sb);
      job.head_.bytes(sb.data);
    }
    hl.lap(4);
    job.plans_[0].sequence = attribute_sequence(job.table_refs_[0], eb.corner_list());
    for (size_t i = 1; i < natt; ++i) {
      const SeamTable& st = job.seams_[i - 1];
      // an attribute whose only seams are mesh boundaries has the universal table, hence the position sequence
      if (!st.has_interior_seam && st.num_vertices == ut.num_vertices) job.plans_[i].shares_sequence_of = 0;
      else job.plans_[i].sequence = attribute_sequence(job.table_refs_[i], eb.corner_list());
    }
    job.write_attribute_section_headers();
    for (size_t i = 0; i < natt; ++i) job.plans_[i].table = &job.table_refs_[i];
    hl.lap(5);
  });
}

// Sizes of a group and upper bounds of its two slabs (every array is padded to 256 bytes: + 256 per array).
GroupSizes GroupRunner::measure(std::vector<GroupMesh*>& meshes) {
  const size_t G = meshes.size();
  uint64_t sumC = 0, sumV = 0, sumF = 0, sum_values = 0, sum_maps = 0, sum_cap = 0, sum_seamC = 0, sum_seamV = 0, sum_symbols = 0, sum_hist = 0;
  uint32_t maxV = 1;
  size_t num_streams = 0, num_seams = 0;
  for (GroupMesh* gm : meshes) {
    GroupMesh& m = *gm;
    const MeshJob& job = *m.job;
    m.F = (uint32_t)m.mesh->num_faces; m.C = m.F * 3u; m.V = job.plans_[0].view.num_unique;
    sumC += m.C; sumV += m.V; sumF += m.F;
    maxV = std::max(maxV, m.V);
    m.atts.assign(job.plans_.size(), GroupMesh::Att{});
    for (size_t i = 0; i < job.plans_.size(); ++i) {
      const AttrPlan& p = job.plans_[i];
      GroupMesh::Att& a = m.atts[i];
      a.value_bytes = (size_t)p.view.num_unique * p.ncomp_in * 4;
      a.has_map = p.view.map != nullptr;
      sum_values += align_up(a.value_bytes, 256);
      if (a.has_map) sum_maps += align_up((size_t)p.view.num_points * 4, 256);
      if (i > 0) {
        a.capacity = std::max(p.view.num_points, m.V) + 16u;
        sum_cap += a.capacity; sum_seamC += m.C; sum_seamV += m.V;
        ++num_seams;
      }
      const uint64_t max_elems = i == 0 ? m.V : a.capacity;
      sum_symbols += max_elems * p.ncomp_q;
      sum_hist += p.hist_capacity;
      ++num_streams;
    }
  }
  const uint32_t vertex_bits = [&] { uint32_t b = 1; while (((uint64_t)(maxV - 1) >> b) != 0) ++b; return b; }();
  const size_t sort_bytes = gpu::seg_corner_tables_scratch_bytes(sumC);
  const size_t scan_bytes = gpu::seg_seam_tables_scratch_bytes(sum_seamV);
  // upper bounds of the two slabs (every array is padded to 256 bytes: + 256 per array)
  const size_t per_array = 256;
  const size_t tiles_bound = ((sumC + sumV) / gpu::kSegTile + 2 * G) * sizeof(gpu::Tile) * (2 + 2 * 3) + (sum_symbols / gpu::kSegTile + num_streams) * sizeof(gpu::Tile) * 24;
  uint64_t rans_scratch = 0;
  for (GroupMesh* gm : meshes)
    for (size_t i = 0; i < gm->job->plans_.size(); ++i)
      rans_scratch += align_up(gpu::rans_scratch_bytes((uint64_t)(i == 0 ? gm->V : gm->atts[i].capacity) * gm->job->plans_[i].ncomp_q), 256);
  const size_t pair_bytes =
      (size_t)sumC * 4 + sum_values + sum_maps                                   // inputs
      + (size_t)sumC * 8 + (size_t)sumV * 5 + G * 4                               // cv, opposite, left_most, interior, flags
      + (size_t)sum_seamC * 5 + (size_t)sum_cap * 5 + num_seams * 8               // seam tables
      + (size_t)(sumV + sum_cap) * 4                                              // sequences
      + num_streams * (sizeof(gpu::AttrSeg) + sizeof(gpu::RansJob) + sizeof(gpu::AttrStats) + sizeof(gpu::SideStats) + 16 + 4)
      + G * sizeof(gpu::MeshSeg) + num_seams * sizeof(gpu::SeamSeg) + tiles_bound
      + (size_t)sum_symbols * 3 + (size_t)(sumV + sum_cap) + (size_t)sum_hist * 3 + num_streams * 64  // packed output
      + (G * (8 + 10 * 3) + 64) * per_array;
  const size_t only_stage1 = (size_t)sumV * 8 + (size_t)sum_seamV * 9 + sort_bytes + scan_bytes + (G * 2 + num_seams * 3 + 8) * per_array;
  const size_t only_stage3 =
      (size_t)sumF * 16 * 2 + (size_t)sum_seamC / 3 * 16 + (size_t)sum_seamC * 8  // face tuples, vertex tuples, fan links
      + (size_t)sum_values / 3 * 4 + (size_t)sum_values                           // quantized values (3 -> 4 components at most)
      + (size_t)(sumV + sum_cap) * 4 + (size_t)sum_symbols * 4 + (size_t)(sumV + sum_cap) * 2 + num_streams * 16
      + (size_t)sum_hist * (4 + 12 + 16 + 3) + num_streams * 64
      + (size_t)sum_symbols * 3 + num_streams * 16 + rans_scratch
      + (num_streams * 12 + G * 3 + 16) * per_array;
  GroupSizes gs;
  gs.sumC = sumC; gs.sum_seamV = sum_seamV; gs.vertex_bits = vertex_bits; gs.sort_bytes = sort_bytes; gs.scan_bytes = scan_bytes;
  gs.pair_bytes = pair_bytes + (size_t)sumC / 4 + 16384 + 64;                                                        // boundary list
  gs.only_bytes = std::max(only_stage1 + (size_t)sumC * 4 + gpu::boundary_list_scratch_bytes(sumC) + 512, only_stage3);
  return gs;
}

void GroupRunner::run(std::vector<GroupMesh*>& meshes) {
  const bool timing = getenv("DXO_TIMING") != nullptr;
  auto t0 = Clock::now();
  auto lap = [&](const char* what) {
    if (!timing) return;
    fprintf(stderr, "[dxo] group(%zu meshes) %-28s %9.3f ms\n", meshes.size(), what, ms_since(t0));
    t0 = Clock::now();
  };
  cuda_check(cudaSetDevice(slot_.device), "cudaSetDevice");
  const size_t G = meshes.size();
  if (G == 0) return;

  const GroupSizes gs = measure(meshes);
  const uint64_t sum_seamV = gs.sum_seamV;
  const uint32_t vertex_bits = gs.vertex_bits;
  const size_t sort_bytes = gs.sort_bytes, scan_bytes = gs.scan_bytes;
  slot_.ensure(std::max(gs.pair_bytes, min_pair_), std::max(gs.only_bytes, min_only_));
  cudaStream_t s = slot_.stream;  // exists from here on (created by the slot's first ensure())
  pair_ = Slab{slot_.h_pair, slot_.pair_cap, 0};
  only_ = Slab{slot_.d_only, slot_.only_cap, 0};

  // ---------------------------------------------------------------- stage 0: layout of the inputs, copies by the workers
  const size_t in1_begin = pair_.take(0);
  for (GroupMesh* gm : meshes) gm->faces = pair_.take((size_t)gm->C * 4);
  for (GroupMesh* gm : meshes)
    for (size_t i = 0; i < gm->atts.size(); ++i) {
      gm->atts[i].values = pair_.take(gm->atts[i].value_bytes);
      if (gm->atts[i].has_map) gm->atts[i].map = pair_.take((size_t)gm->job->plans_[i].view.num_points * 4);
    }
  workers_.parallel_for(G, [&](size_t k) { stage0_copy_inputs(*meshes[k]); });
  lap("stage 0 (copy inputs)");

  // ---------------------------------------------------------------- stage 1: K12-K14 over the group
  // tables that go back to the host (one contiguous range), by kind so that each kind is initialised by one memset
  std::vector<GroupMesh*> live;
  for (GroupMesh* gm : meshes) if (gm->alive) live.push_back(gm);
  // What every mesh's host walks need goes back in one contiguous range: opposite (one gap-free array in the group-wide
  // corner order, so that the boundary list can be selected from it in one pass), left-most corners, interior flags,
  // the flag words and the ordered list of boundary corners. corner -> vertex stays on the device (the host has the
  // faces / computes map[faces] itself), and the seam tables are fetched afterwards only for the attributes that turn
  // out to have interior seams — for the others the attribute's table IS the universal one.
  uint64_t liveC = 0;
  for (GroupMesh* gm : live) { gm->corner_base = (uint32_t)liveC; liveC += gm->C; }
  const size_t out1_begin = pair_.take(0);
  const size_t opp_begin = pair_.take((size_t)liveC * 4);
  for (GroupMesh* gm : live) gm->opposite = opp_begin + (size_t)gm->corner_base * 4;
  const size_t opp_end = pair_.used;
  for (GroupMesh* gm : live) gm->left_most = pair_.take((size_t)gm->V * 4, 4);
  for (GroupMesh* gm : live) gm->interior = pair_.take(gm->V, 1);
  const size_t scalars_begin = pair_.take(0, 4);
  for (GroupMesh* gm : live) gm->flags = pair_.take(4, 4);
  for (GroupMesh* gm : live) for (size_t i = 1; i < gm->atts.size(); ++i) gm->atts[i].scalars = pair_.take(8, 4);
  const size_t boundary_count = pair_.take(4, 4);
  const size_t scalars_end = pair_.used;
  const uint32_t boundary_cap = (uint32_t)std::min<uint64_t>(liveC, liveC / 16 + 4096);  // list entries copied back (more: the host scans instead)
  const size_t boundary_list = pair_.take((size_t)boundary_cap * 4, 4);
  // corner -> vertex of the meshes whose position attribute has a point map follows directly (same copy); the others' is their faces
  for (GroupMesh* gm : live) if (gm->atts[0].has_map) gm->cv = pair_.take((size_t)gm->C * 4, 4);
  const size_t out1_end = pair_.used;
  for (GroupMesh* gm : live) if (!gm->atts[0].has_map) gm->cv = pair_.take((size_t)gm->C * 4);
  for (GroupMesh* gm : live)
    for (size_t i = 1; i < gm->atts.size(); ++i) {
      GroupMesh::Att& a = gm->atts[i];
      a.cv_a = pair_.take((size_t)gm->C * 4);       // the four arrays of an attribute are one contiguous range (one copy when fetched)
      a.left_most_a = pair_.take((size_t)a.capacity * 4, 4);
      a.seam = pair_.take(gm->C, 1);
      a.interior_a = pair_.take(a.capacity, 1);
    }

  // device-only scratch of stage 1
  const size_t fc_begin = only_.take(0);
  for (GroupMesh* gm : live) gm->first_corner = only_.take((size_t)gm->V * 4);
  const size_t fc_end = only_.used;
  const size_t zero_begin = only_.take(0);
  for (GroupMesh* gm : live) gm->valence = only_.take((size_t)gm->V * 4);
  for (GroupMesh* gm : live) for (size_t i = 1; i < gm->atts.size(); ++i) gm->atts[i].vertex_on_seam = only_.take(gm->V);
  const size_t counts_off = only_.take((size_t)std::max<uint64_t>(sum_seamV, 1) * 4);
  const size_t zero_end = only_.used;
  const size_t bases_off = only_.take((size_t)std::max<uint64_t>(sum_seamV, 1) * 4);
  const size_t sort_off = only_.take(sort_bytes);
  const size_t scan_off = only_.take(scan_bytes);
  const size_t blist_bytes = gpu::boundary_list_scratch_bytes(liveC);
  const size_t blist_scratch = only_.take(blist_bytes);
  const size_t blist_full = only_.take((size_t)std::max<uint64_t>(liveC, 1) * 4);  // the select writes every boundary corner; a prefix goes back

  // descriptors and tiles (part of the input range)
  std::vector<gpu::MeshSeg> mesh_segs(live.size());
  std::vector<gpu::SeamSeg> seam_segs;
  std::vector<gpu::Tile> t_corner, t_vertex, t_seam_corner, t_seam_vertex, t_seam_attr;
  uint32_t count_base = 0;
  for (size_t k = 0; k < live.size(); ++k) {
    GroupMesh& m = *live[k];
    const MeshJob& job = *m.job;
    gpu::MeshSeg& g = mesh_segs[k];
    g.faces = dp<uint32_t>(m.faces);
    g.pos_map = m.atts[0].has_map ? dp<uint32_t>(m.atts[0].map) : nullptr;
    g.num_corners = m.C; g.num_points = job.plans_[0].view.num_points; g.num_vertices = m.V;
    g.corner_base = m.corner_base;
    g.cv = dp<uint32_t>(m.cv); g.opposite = dp<uint32_t>(m.opposite); g.left_most = dp<uint32_t>(m.left_most); g.interior = dp<uint8_t>(m.interior);
    g.first_corner = dn<uint32_t>(m.first_corner); g.valence = dn<uint32_t>(m.valence);
    g.flags = dp<uint32_t>(m.flags);
    add_tiles(t_corner, (uint32_t)k, m.C);
    add_tiles(t_vertex, (uint32_t)k, m.V);
    for (size_t i = 1; i < m.atts.size(); ++i) {
      GroupMesh::Att& a = m.atts[i];
      gpu::SeamSeg sg{};
      sg.faces = g.faces;
      sg.map = a.has_map ? dp<uint32_t>(a.map) : nullptr;
      sg.num_points = job.plans_[i].view.num_points;
      sg.cv_u = g.cv; sg.opposite = g.opposite; sg.left_most_u = g.left_most;
      sg.num_corners = m.C; sg.num_vertices_u = m.V;
      sg.count_base = count_base; a.count_base = count_base;
      count_base += m.V;
      sg.capacity = a.capacity;
      sg.seam = dp<uint8_t>(a.seam); sg.vertex_on_seam = dn<uint8_t>(a.vertex_on_seam);
      sg.cv_a = dp<uint32_t>(a.cv_a); sg.left_most_a = dp<uint32_t>(a.left_most_a); sg.interior_a = dp<uint8_t>(a.interior_a);
      sg.scalars = dp<uint32_t>(a.scalars);
      sg.mesh_flags = g.flags;
      a.seam_seg = (int)seam_segs.size();
      add_tiles(t_seam_corner, (uint32_t)seam_segs.size(), m.C);
      add_tiles(t_seam_vertex, (uint32_t)seam_segs.size(), m.V);
      add_tiles(t_seam_attr, (uint32_t)seam_segs.size(), a.capacity);
      seam_segs.push_back(sg);
    }
  }
  // descriptors travel in a second small range right behind the tables (the layout above had to exist first)
  const size_t desc1_begin = pair_.take(0);
  const size_t o_mesh_segs = put_vector(mesh_segs), o_seam_segs = put_vector(seam_segs);
  const size_t o_t_corner = put_vector(t_corner), o_t_vertex = put_vector(t_vertex);
  const size_t o_t_sc = put_vector(t_seam_corner), o_t_sv = put_vector(t_seam_vertex), o_t_sa = put_vector(t_seam_attr);
  const size_t desc1_end = pair_.used;

  if (!live.empty()) {
    h2d(in1_begin, out1_begin);
    h2d(desc1_begin, desc1_end);
    cuda_check(cudaMemsetAsync(slot_.d_pair + opp_begin, 0xFF, opp_end - opp_begin, s), "cudaMemsetAsync");
    cuda_check(cudaMemsetAsync(slot_.d_pair + scalars_begin, 0, scalars_end - scalars_begin, s), "cudaMemsetAsync");
    cuda_check(cudaMemsetAsync(slot_.d_only + fc_begin, 0xFF, fc_end - fc_begin, s), "cudaMemsetAsync");
    cuda_check(cudaMemsetAsync(slot_.d_only + zero_begin, 0, zero_end - zero_begin, s), "cudaMemsetAsync");
    gpu::launch_seg_corner_tables(dp<gpu::MeshSeg>(o_mesh_segs), (uint32_t)live.size(), dp<gpu::Tile>(o_t_corner), (uint32_t)t_corner.size(),
                                  dp<gpu::Tile>(o_t_vertex), (uint32_t)t_vertex.size(), liveC, vertex_bits, dn<uint8_t>(sort_off), sort_bytes, s);
    // corners without an opposite, ascending in the group-wide corner order (flagged meshes contribute garbage that nobody reads)
    gpu::launch_boundary_list(dp<uint32_t>(opp_begin), liveC, dn<uint8_t>(blist_scratch), blist_bytes, dn<uint32_t>(blist_full), dp<uint32_t>(boundary_count), s);
    cuda_check(cudaMemcpyAsync(slot_.d_pair + boundary_list, slot_.d_only + blist_full, (size_t)boundary_cap * 4, cudaMemcpyDeviceToDevice, s), "cudaMemcpyAsync D2D");
    gpu::launch_seg_seam_tables(dp<gpu::SeamSeg>(o_seam_segs), (uint32_t)seam_segs.size(), dp<gpu::Tile>(o_t_sc), (uint32_t)t_seam_corner.size(),
                                dp<gpu::Tile>(o_t_sv), (uint32_t)t_seam_vertex.size(), dp<gpu::Tile>(o_t_sa), (uint32_t)t_seam_attr.size(),
                                dn<uint32_t>(counts_off), dn<uint32_t>(bases_off), sum_seamV, dn<uint8_t>(scan_off), scan_bytes, s);
    cuda_check(cudaGetLastError(), "kernel launch (group stage 1)");
    d2h(out1_begin, out1_end);
    slot_.wait();
    // second step: the seam tables of the attributes that have interior seams (or more vertices than the universal table)
    bool any = false;
    for (GroupMesh* gm : live) {
      if (*hp<uint32_t>(gm->flags)) continue;
      for (size_t i = 1; i < gm->atts.size(); ++i) {
        const GroupMesh::Att& a = gm->atts[i];
        const uint32_t* sc = hp<uint32_t>(a.scalars);
        if ((sc[1] & (1u | 2u | 8u)) || sc[0] > a.capacity || sc[0] < gm->V) continue;  // the mesh falls back
        if (!(sc[1] & 4u) && sc[0] == gm->V) continue;                                    // the universal table serves
        d2h(a.cv_a, a.interior_a + a.capacity);
        any = true;
      }
    }
    if (any) slot_.wait();
  }
  lap("stage 1 (K12-K14)");
  // boundary corners per mesh: the list is ascending in the group-wide corner order
  const uint32_t num_boundary = live.empty() ? 0u : *hp<uint32_t>(boundary_count);
  const bool have_boundary_list = !live.empty() && num_boundary <= boundary_cap;
  const uint32_t* blist = hp<uint32_t>(boundary_list);

  // ---------------------------------------------------------------- stage 2: host connectivity per mesh
  workers_.parallel_for(live.size(), [&](size_t k) {
    GroupMesh& m = *live[k];
    if (have_boundary_list && *hp<uint32_t>(m.flags) == 0) {
      const uint32_t* lo = std::lower_bound(blist, blist + num_boundary, m.corner_base);
      const uint32_t* hi = std::lower_bound(lo, blist + num_boundary, m.corner_base + m.C);
      std::vector<uint32_t>& bc = m.job->ut_.boundary_corners;
      bc.resize((size_t)(hi - lo));
      for (size_t j = 0; j < bc.size(); ++j) bc[j] = lo[j] - m.corner_base;
      m.job->ut_.has_boundary_list = true;
    }
    stage2_host_connectivity(m);
  });
  lap("stage 2 (traversal, sequences)");

  // ---------------------------------------------------------------- stage 3: K1-K10 over the group
  std::vector<GroupMesh*> act;
  for (GroupMesh* gm : live) if (gm->alive) act.push_back(gm);
  if (act.empty()) return;
  only_.used = 0;  // stage 1's scratch is dead
  const size_t in3_begin = pair_.take(0);
  for (GroupMesh* gm : act)
    for (size_t i = 0; i < gm->atts.size(); ++i) {
      const AttrPlan& p = gm->job->plans_[i];
      if (p.shares_sequence_of >= 0) continue;
      gm->atts[i].seq = pair_.take(p.sequence.size() * 4);
      copy_to_pinned(hp<uint8_t>(gm->atts[i].seq), p.sequence.data(), p.sequence.size() * 4);
    }
  // per-stream device arrays
  struct StreamMem { size_t quant, rank, symbols, side_flags, hist, work, rans_table, table_bytes, payload, rans_scratch, side_payload, fan_link, cv4; };
  std::vector<gpu::AttrSeg> segs;
  std::vector<StreamMem> mem;
  using Pad3Job = gpu::Pad3Seg;
  std::vector<std::pair<size_t, size_t>> stream_of;  // (mesh index in act, attribute)
  // group-wide kinds first (single memsets): rank (0xFF), hist (0)
  size_t n_streams = 0;
  for (GroupMesh* gm : act) n_streams += gm->atts.size();
  mem.resize(n_streams);
  segs.resize(n_streams);
  {
    size_t k = 0;
    for (size_t mi = 0; mi < act.size(); ++mi)
      for (size_t i = 0; i < act[mi]->atts.size(); ++i) { act[mi]->atts[i].stream = (int)k; stream_of.push_back({mi, i}); ++k; }
  }
  auto table_vertices = [&](const GroupMesh& m, size_t i) { return i == 0 ? m.V : m.job->seams_[i - 1].num_vertices; };
  const size_t rank_begin = only_.take(0);
  for (size_t k = 0; k < n_streams; ++k) { const GroupMesh& m = *act[stream_of[k].first]; mem[k].rank = only_.take((size_t)table_vertices(m, stream_of[k].second) * 4); }
  const size_t rank_end = only_.used;
  const size_t hist_begin = only_.take(0);
  for (size_t k = 0; k < n_streams; ++k) mem[k].hist = only_.take((size_t)act[stream_of[k].first]->job->plans_[stream_of[k].second].hist_capacity * 4);
  const size_t hist_end = only_.used;
  // face tuples per mesh
  std::vector<size_t> faces4(act.size()), cv4_u(act.size(), 0);
  for (size_t mi = 0; mi < act.size(); ++mi) {
    faces4[mi] = only_.take((size_t)act[mi]->F * 16);
    if (act[mi]->atts[0].has_map) cv4_u[mi] = only_.take((size_t)act[mi]->F * 16);
  }
  // results that go back to the host: flags of the long side streams | stats | side stats | index, then the packed bytes.
  // A binary side stream is a serial chain of ~15 ns per bit on the device (one warp per stream): fine for the many short
  // streams of a group, which then never touch the host, but a 100k-vertex mesh's stream would take longer than all
  // other kernels of the group together — those are coded by the host workers during assembly (~1.5 ns per bit).
  static const uint32_t device_rabs_max = getenv("DXO_DEVICE_RABS_MAX") ? (uint32_t)atoi(getenv("DXO_DEVICE_RABS_MAX")) : 8192u;
  const size_t out3_begin = pair_.take(0);
  for (GroupMesh* gm : act)
    for (size_t i = 0; i < gm->atts.size(); ++i) {
      const AttrPlan& p = gm->job->plans_[i];
      GroupMesh::Att& a = gm->atts[i];
      const size_t M = gm->job->sequence_of(i).size();
      a.host_side = (p.scheme == Scheme::Normal || p.scheme == Scheme::TexCoord) && M > device_rabs_max;
      if (a.host_side) a.host_flags = pair_.take(M);
    }
  const size_t o_stats = pair_.take(n_streams * sizeof(gpu::AttrStats));
  const size_t o_side_stats = pair_.take(n_streams * sizeof(gpu::SideStats));
  const size_t o_index = pair_.take((n_streams + 1) * sizeof(uint4));
  const size_t out3_end = pair_.used;

  std::vector<gpu::RansJob> jobs(n_streams);
  std::vector<uint32_t> side_ids;
  uint64_t packed_capacity = 0;
  uint32_t max_table_capacity = 0;
  for (size_t k = 0; k < n_streams; ++k) {
    GroupMesh& m = *act[stream_of[k].first];
    const size_t i = stream_of[k].second;
    MeshJob& job = *m.job;
    const AttrPlan& p = job.plans_[i];
    const GroupMesh::Att& a = m.atts[i];
    const size_t mi = stream_of[k].first;
    StreamMem& sm = mem[k];
    const uint32_t U = p.view.num_unique;
    const uint32_t M = (uint32_t)job.sequence_of(i).size();
    const uint32_t S = M * p.ncomp_q;
    const size_t qstride = p.ncomp_q == 3 ? 4 : p.ncomp_q;
    const bool to_bits = p.port == Portabilization::ToBits;
    sm.quant = (to_bits && p.ncomp_q != 3) ? 0 : only_.take((size_t)U * qstride * 4);
    sm.symbols = only_.take((size_t)std::max<uint32_t>(S, 1) * 4);
    sm.side_flags = a.host_side ? 0 : only_.take(std::max<uint32_t>(M, 1));
    sm.work = only_.take((size_t)p.hist_capacity * 12);
    sm.rans_table = only_.take(((size_t)p.hist_capacity + 1) * 16);
    const uint32_t table_capacity = 3 * p.hist_capacity + 16;
    sm.table_bytes = only_.take(table_capacity);
    const uint64_t payload_capacity = 3ull * S + 16;
    sm.payload = only_.take(payload_capacity);
    sm.rans_scratch = only_.take(gpu::rans_scratch_bytes(S));
    const bool has_side = p.scheme == Scheme::Normal || p.scheme == Scheme::TexCoord;
    sm.side_payload = (has_side && !a.host_side) ? only_.take((size_t)M + 16) : 0;
    sm.fan_link = (i > 0 && p.scheme == Scheme::Normal) ? only_.take((size_t)m.C * 8) : 0;
    sm.cv4 = i > 0 ? only_.take((size_t)m.F * 16) : 0;
    max_table_capacity = std::max(max_table_capacity, p.hist_capacity);

    gpu::AttrSeg& g = segs[k];
    g = gpu::AttrSeg{};
    gpu::TableDev& t = g.t;
    t.corner_point = dp<uint32_t>(m.faces);
    t.corner_point4 = dn<uint4>(faces4[mi]);
    t.opposite = dp<uint32_t>(m.opposite);
    t.num_corners = m.C;
    const bool vertex_is_point = !m.atts[0].has_map;
    if (i == 0) {
      t.corner_vertex = dp<uint32_t>(m.cv);
      t.corner_vertex4 = vertex_is_point ? dn<uint4>(faces4[mi]) : dn<uint4>(cv4_u[mi]);
      t.vertex_is_point = vertex_is_point ? 1 : 0;
      t.seam = nullptr; t.left_most = dp<uint32_t>(m.left_most); t.num_vertices = m.V; t.fan_link = nullptr;
    } else {
      t.corner_vertex = dp<uint32_t>(a.cv_a);
      t.corner_vertex4 = dn<uint4>(sm.cv4);
      t.vertex_is_point = 0;
      t.seam = dp<uint8_t>(a.seam); t.left_most = dp<uint32_t>(a.left_most_a); t.num_vertices = job.seams_[i - 1].num_vertices;
      t.fan_link = sm.fan_link ? dn<uint2>(sm.fan_link) : nullptr;
    }
    g.values = dp<float>(a.values);
    const int32_t* quant = (to_bits && p.ncomp_q != 3) ? (const int32_t*)dp<float>(a.values) : dn<int32_t>(sm.quant);
    g.q = gpu::QuantDev{quant, a.has_map ? dp<uint32_t>(a.map) : nullptr, p.ncomp_q};
    if (p.parent >= 0) {
      const GroupMesh::Att& pa = m.atts[p.parent];
      const StreamMem& pm = mem[pa.stream];
      g.pos = gpu::QuantDev{dn<int32_t>(pm.quant), pa.has_map ? dp<uint32_t>(pa.map) : nullptr, 3};
      g.pos_num_points = job.plans_[p.parent].view.num_points;
    }
    g.num_unique = U; g.ncomp_in = p.ncomp_in; g.bits = p.bits;
    g.scheme = (uint32_t)p.scheme; g.wrapped = p.transform == Transform::Wrapped ? 1u : 0u;
    const size_t seq_att = p.shares_sequence_of >= 0 ? (size_t)p.shares_sequence_of : i;
    g.seq = dp<uint32_t>(m.atts[seq_att].seq);
    g.n = M; g.num_symbols = S;
    g.rank = dn<uint32_t>(sm.rank); g.symbols = dn<uint32_t>(sm.symbols);
    g.side_flags = a.host_side ? dp<uint8_t>(a.host_flags) : dn<uint8_t>(sm.side_flags);
    g.hist = dn<uint32_t>(sm.hist); g.hist_capacity = p.hist_capacity; g.work = dn<uint32_t>(sm.work);
    g.rans_table = dn<uint4>(sm.rans_table); g.table_bytes = dn<uint8_t>(sm.table_bytes); g.table_capacity = table_capacity;
    g.stats = dp<gpu::AttrStats>(o_stats) + k;
    g.fan_link_out = sm.fan_link ? dn<uint2>(sm.fan_link) : nullptr;
    g.seam_for_links = i > 0 ? dp<uint8_t>(a.seam) : nullptr;
    if (has_side && !a.host_side) {
      g.side_payload = dn<uint8_t>(sm.side_payload); g.side_capacity = M + 16; g.side_stats = dp<gpu::SideStats>(o_side_stats) + k;
      side_ids.push_back((uint32_t)k);
    }
    jobs[k] = gpu::rans_make_job(g.symbols, S, g.rans_table, dn<uint8_t>(sm.rans_scratch), dn<uint8_t>(sm.payload), g.stats);
    packed_capacity += align_up(table_capacity, 4) + align_up(payload_capacity, 4) + ((has_side && !a.host_side) ? align_up((size_t)M + 16, 4) : 0);
  }
  // tile lists per kernel class
  std::vector<gpu::Tile> t_fan, t_mm[5], t_oct, t_prep, t_par[5], t_delta[5], t_nrm, t_uv, t_hs, t_hg;
  std::vector<Pad3Job> pads;
  for (size_t mi = 0; mi < act.size(); ++mi) {
    GroupMesh& m = *act[mi];
    pads.push_back({dp<uint32_t>(m.faces), dn<uint4>(faces4[mi]), m.F});
    if (m.atts[0].has_map) pads.push_back({dp<uint32_t>(m.cv), dn<uint4>(cv4_u[mi]), m.F});
  }
  for (size_t k = 0; k < n_streams; ++k) {
    const GroupMesh& m = *act[stream_of[k].first];
    const size_t i = stream_of[k].second;
    const AttrPlan& p = m.job->plans_[i];
    const gpu::AttrSeg& g = segs[k];
    if (i > 0) pads.push_back({dp<uint32_t>(m.atts[i].cv_a), dn<uint4>(mem[k].cv4), m.F});
    if (g.fan_link_out) add_tiles(t_fan, (uint32_t)k, m.C);
    if (p.port == Portabilization::Quantize) add_tiles(t_mm[p.ncomp_in], (uint32_t)k, g.num_unique);
    else if (p.port == Portabilization::Octahedral) add_tiles(t_oct, (uint32_t)k, g.num_unique);
    else if (p.ncomp_q == 3) pads.push_back({(const uint32_t*)g.values, (uint4*)dn<int32_t>(mem[k].quant), g.num_unique});
    if (p.scheme == Scheme::Parallelogram || p.scheme == Scheme::TexCoord) add_tiles(t_prep, (uint32_t)k, g.n);
    switch (p.scheme) {
      case Scheme::Parallelogram: add_tiles(t_par[p.ncomp_q], (uint32_t)k, g.n); break;
      case Scheme::Delta: add_tiles(t_delta[p.ncomp_q], (uint32_t)k, g.n); break;
      case Scheme::Normal: add_tiles(t_nrm, (uint32_t)k, g.n); break;
      case Scheme::TexCoord: add_tiles(t_uv, (uint32_t)k, g.n); break;
    }
    add_tiles(p.hist_capacity <= gpu::kSmemHistBins ? t_hs : t_hg, (uint32_t)k, g.num_symbols, 8 * gpu::kSegTile);
  }
  // pad3 jobs run as tiles of a tiny descriptor list: reuse AttrSeg-free generic form (seg = job, first = tuple)
  std::vector<gpu::Tile> t_pad;
  for (size_t j = 0; j < pads.size(); ++j) add_tiles(t_pad, (uint32_t)j, pads[j].n);
  gpu::RansTiles rt;
  gpu::rans_plan_tiles(jobs.data(), (uint32_t)jobs.size(), rt);

  const size_t o_segs = put_vector(segs), o_jobs = put_vector(jobs), o_side_ids = put_vector(side_ids), o_pads = put_vector(pads), o_t_pad = put_vector(t_pad);
  const size_t o_t_fan = put_vector(t_fan), o_t_oct = put_vector(t_oct), o_t_prep = put_vector(t_prep), o_t_nrm = put_vector(t_nrm), o_t_uv = put_vector(t_uv);
  const size_t o_t_hs = put_vector(t_hs), o_t_hg = put_vector(t_hg);
  size_t o_t_mm[5], o_t_par[5], o_t_delta[5];
  for (int c = 1; c <= 4; ++c) { o_t_mm[c] = put_vector(t_mm[c]); o_t_par[c] = put_vector(t_par[c]); o_t_delta[c] = put_vector(t_delta[c]); }
  const size_t o_rt_explore = put_vector(rt.explore), o_rt_chain = put_vector(rt.chain), o_rt_lanes = put_vector(rt.lanes), o_rt_pairs = put_vector(rt.pairs);
  const size_t o_rt_fixup = put_vector(rt.fixup), o_rt_gather = put_vector(rt.gather);
  const size_t in3_end = pair_.used;
  const size_t o_packed = pair_.take(packed_capacity);
  lap("stage 3 (descriptors)");
  if (getenv("DXO_DEBUG"))
    fprintf(stderr, "[dxo] group: %zu meshes, %zu live, %zu active, %zu streams; tiles pad %zu explore %zu chain %zu lanes %zu pairs %zu fixup %zu gather %zu; packed capacity %llu\n", G,
            live.size(), act.size(), n_streams, t_pad.size(), rt.explore.size(), rt.chain.size(), rt.lanes.size(), rt.pairs.size(), rt.fixup.size(), rt.gather.size(),
            (unsigned long long)packed_capacity);

  const gpu::AttrSeg* d_segs = dp<gpu::AttrSeg>(o_segs);
  h2d(in3_begin, out3_begin);
  h2d(out3_end, in3_end);
  cuda_check(cudaMemsetAsync(slot_.d_only + rank_begin, 0xFF, rank_end - rank_begin, s), "cudaMemsetAsync");
  cuda_check(cudaMemsetAsync(slot_.d_only + hist_begin, 0, hist_end - hist_begin, s), "cudaMemsetAsync");
  gpu::launch_seg_pad3(dp<gpu::Pad3Seg>(o_pads), dp<gpu::Tile>(o_t_pad), (uint32_t)t_pad.size(), s);
  gpu::launch_seg_init_stats(d_segs, (uint32_t)n_streams, s);
  gpu::launch_seg_fan_links(d_segs, dp<gpu::Tile>(o_t_fan), (uint32_t)t_fan.size(), s);
  for (uint32_t c = 1; c <= 4; ++c) gpu::launch_seg_minmax(d_segs, dp<gpu::Tile>(o_t_mm[c]), (uint32_t)t_mm[c].size(), c, s);
  for (uint32_t c = 1; c <= 4; ++c) gpu::launch_seg_quantize(d_segs, dp<gpu::Tile>(o_t_mm[c]), (uint32_t)t_mm[c].size(), c, s);
  gpu::launch_seg_oct_quantize(d_segs, dp<gpu::Tile>(o_t_oct), (uint32_t)t_oct.size(), s);
  gpu::launch_seg_seq_prepare(d_segs, dp<gpu::Tile>(o_t_prep), (uint32_t)t_prep.size(), s);
  for (uint32_t c = 1; c <= 4; ++c) gpu::launch_seg_predict(d_segs, dp<gpu::Tile>(o_t_par[c]), (uint32_t)t_par[c].size(), (uint32_t)Scheme::Parallelogram, c, s);
  for (uint32_t c = 1; c <= 4; ++c) gpu::launch_seg_predict(d_segs, dp<gpu::Tile>(o_t_delta[c]), (uint32_t)t_delta[c].size(), (uint32_t)Scheme::Delta, c, s);
  gpu::launch_seg_predict(d_segs, dp<gpu::Tile>(o_t_nrm), (uint32_t)t_nrm.size(), (uint32_t)Scheme::Normal, 2, s);
  gpu::launch_seg_predict(d_segs, dp<gpu::Tile>(o_t_uv), (uint32_t)t_uv.size(), (uint32_t)Scheme::TexCoord, 2, s);
  gpu::launch_seg_histogram(d_segs, dp<gpu::Tile>(o_t_hs), (uint32_t)t_hs.size(), true, s);
  gpu::launch_seg_histogram(d_segs, dp<gpu::Tile>(o_t_hg), (uint32_t)t_hg.size(), false, s);
  gpu::launch_seg_build_tables(d_segs, (uint32_t)n_streams, s);
  gpu::launch_seg_side_streams(d_segs, dp<uint32_t>(o_side_ids), (uint32_t)side_ids.size(), s);
  gpu::RansTilesDev rtd{dp<gpu::Tile>(o_rt_explore), dp<gpu::Tile>(o_rt_chain), dp<gpu::Tile>(o_rt_lanes), dp<gpu::Tile>(o_rt_pairs), dp<gpu::Tile>(o_rt_fixup),
                        dp<gpu::Tile>(o_rt_gather), (uint32_t)rt.explore.size(), (uint32_t)rt.chain.size(), (uint32_t)rt.lanes.size(), (uint32_t)rt.pairs.size(),
                        (uint32_t)rt.fixup.size(), (uint32_t)rt.gather.size()};
  gpu::launch_seg_rans(dp<gpu::RansJob>(o_jobs), rtd, max_table_capacity, s);
  gpu::launch_seg_pack(d_segs, dp<gpu::RansJob>(o_jobs), (uint32_t)n_streams, dp<uint4>(o_index), dp<uint8_t>(o_packed), packed_capacity, nullptr, 0, s);
  cuda_check(cudaGetLastError(), "kernel launch (group stage 3)");
  d2h(out3_begin, out3_end);
  slot_.wait();
  const uint4* index = hp<uint4>(o_index);
  const uint64_t packed_total = index[n_streams].x;
  if (packed_total > packed_capacity) throw Error(DXO_ERR_INTERNAL, "packed output exceeds its buffer");
  d2h(o_packed, o_packed + packed_total);
  slot_.wait();
  lap("stage 3 (K1-K10, pack, D2H)");

  // ---------------------------------------------------------------- stage 4: assembly per mesh
  const gpu::AttrStats* stats = hp<gpu::AttrStats>(o_stats);
  const gpu::SideStats* side_stats = hp<gpu::SideStats>(o_side_stats);
  const uint8_t* packed = hp<uint8_t>(o_packed);
  workers_.parallel_for(act.size(), [&](size_t mi) {
    GroupMesh& m = *act[mi];
    HostLap hl;
    guarded(m, [&] {
      MeshJob& job = *m.job;
      job.results_.assign(job.plans_.size(), AttrResult{});
      for (size_t i = 0; i < job.plans_.size(); ++i) {
        const size_t k = (size_t)m.atts[i].stream;
        AttrResult& r = job.results_[i];
        r.stats = stats[k];
        if (r.stats.error_flags) {
          const uint32_t f = r.stats.error_flags;
          const int st = (f & gpu::kErrZeroNormal) ? DXO_ERR_ZERO_NORMAL : (f & gpu::kErrNegativeSymbol) ? DXO_ERR_RANS_INVALID_SYMBOL
                       : (f & gpu::kErrRansFreq) ? DXO_ERR_RANS_FREQ_TABLE : (f & gpu::kErrRansState) ? DXO_ERR_RANS_STATE_TOO_LARGE : DXO_ERR_UNSUPPORTED_INPUT;
          throw Error(st, "device reported an encoding error");
        }
        const uint4 ix = index[k];
        if (ix.y == 0xFFFFFFFFu || ix.y != r.stats.table_bytes || ix.z != r.stats.payload_bytes)
          throw Error(DXO_ERR_INTERNAL, "packed output index is inconsistent: stream " + std::to_string(k) + " of " + std::to_string(n_streams) + " index {" + std::to_string(ix.x) + "," +
                                            std::to_string(ix.y) + "," + std::to_string(ix.z) + "," + std::to_string(ix.w) + "} stats table " + std::to_string(r.stats.table_bytes) +
                                            " payload " + std::to_string(r.stats.payload_bytes) + " symbols " + std::to_string(job.sequence_of(i).size() * job.plans_[i].ncomp_q));
        const uint8_t* base = packed + ix.x;
        r.table_bytes = base;
        r.payload = base + align_up(ix.y, 4);
        if (m.atts[i].host_side) {
          HostLap side;
          job.encode_side_stream_from_flags(i, hp<uint8_t>(m.atts[i].host_flags), job.sequence_of(i).size());
          side.lap(6);
          hl.t = side.t;
        } else if (job.plans_[i].scheme == Scheme::Normal || job.plans_[i].scheme == Scheme::TexCoord) {
          r.side_bytes = base + align_up(ix.y, 4) + align_up(ix.z, 4);
          r.side_bytes_len = ix.w;
          r.side_count = side_stats[k].count;
          r.side_zero_prob = (uint8_t)side_stats[k].zero_prob;
        }
      }
      std::vector<uint8_t> bytes;
      job.assemble(bytes);
      give_bytes(m, bytes);
      m.status = DXO_OK;
    });
    m.job.reset();
    hl.lap(7);
  });
  lap("stage 4 (assembly)");
}

// -----------------------------------------------------------------------------------------
void encode_batch_grouped(const dxo_mesh* meshes, size_t n, const dxo_config& cfg, dxo_bytes* outs, int* statuses, int first_gpu, int num_gpus) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) { cudaGetLastError(); throw Error(DXO_ERR_NO_DEVICE, "no CUDA device available (this path has no CPU fallback)"); }
  if (num_gpus <= 0) num_gpus = 1;
  if (first_gpu < 0 || first_gpu + num_gpus > count) throw Error(DXO_ERR_NO_DEVICE, "GPU range out of bounds");
  if (n == 0) return;
  const int hw = std::max(1, (int)std::thread::hardware_concurrency());
  const char* env_w = getenv("DXO_BATCH_WORKERS");
  const int num_workers = env_w ? std::max(1, atoi(env_w)) : std::max(2, hw);
  const char* env_s = getenv("DXO_BATCH_SLOTS");
  const int slots_per_gpu = env_s ? std::max(1, atoi(env_s)) : 6;  // measured on config 4 (16 host threads): 3 / 4 / 6 slots -> 132 / 150 / 159 Mvertices/s
  const char* env_c = getenv("DXO_GROUP_CORNERS");
  uint64_t group_corners = env_c ? std::max<uint64_t>(3, strtoull(env_c, nullptr, 10)) : (12ull << 20);
  const size_t group_meshes = 512;

  std::vector<GroupMesh> all(n);
  Workers workers(num_workers);
  // validation and plans of every mesh (MeshJob's constructor), in parallel
  workers.parallel_for(n, [&](size_t i) {
    GroupMesh& m = all[i];
    m.index = i;
    m.mesh = &meshes[i];
    try { m.job.reset(new MeshJob(&meshes[i], cfg)); }
    catch (const Error& e) { m.alive = false; m.status = e.status; m.error = e.what(); }
    catch (const std::exception& e) { m.alive = false; m.status = DXO_ERR_INTERNAL; m.error = e.what(); }
  });
  // longest first, cut into groups
  std::vector<size_t> order;
  uint64_t total_corners = 0;
  for (size_t i = 0; i < n; ++i) if (all[i].alive) { order.push_back(i); total_corners += 3ull * meshes[i].num_faces; }
  // groups large enough to amortise a launch set, small enough that every slot of every GPU gets a few of them
  if (!env_c) group_corners = std::max<uint64_t>(1ull << 20, std::min<uint64_t>(group_corners, total_corners / ((uint64_t)num_gpus * slots_per_gpu * 2) + 1));
  std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return meshes[a].num_faces > meshes[b].num_faces; });
  std::vector<std::vector<GroupMesh*>> groups;
  {
    // A group's host stage runs one mesh per worker, so groups of very large meshes are also kept wide enough that the
    // groups in flight can occupy the workers (up to a hard cap of corners per group: slab memory).
    const size_t min_meshes = std::max<size_t>(1, ((size_t)num_workers + slots_per_gpu - 2) / std::max(1, slots_per_gpu - 1));
    const uint64_t hard_cap = 72ull << 20;
    uint64_t corners = 0;
    for (size_t i : order) {
      const uint64_t c = 3ull * meshes[i].num_faces;
      const bool full = !groups.empty() && corners + c > group_corners && groups.back().size() >= min_meshes;
      if (groups.empty() || (!groups.back().empty() && (full || corners + c > hard_cap)) || groups.back().size() >= group_meshes) { groups.emplace_back(); corners = 0; }
      groups.back().push_back(&all[i]);
      corners += c;
    }
  }
  size_t max_pair = 0, max_only = 0;
  for (auto& g : groups) { const GroupSizes gs = GroupRunner::measure(g); max_pair = std::max(max_pair, gs.pair_bytes); max_only = std::max(max_only, gs.only_bytes); }
  std::atomic<size_t> next{0};
  std::mutex err_mu;
  std::exception_ptr first_error;
  const int num_slots = (int)std::min<size_t>((size_t)num_gpus * slots_per_gpu, std::max<size_t>(groups.size(), 1));
  struct SlotLease {
    std::vector<Slot*> v;
    ~SlotLease() { for (Slot* s : v) slot_pool().release(s); }
    Slot* operator[](size_t k) const { return v[k]; }
  } slots;
  for (int k = 0; k < num_slots; ++k) slots.v.push_back(slot_pool().acquire(first_gpu + k % num_gpus));
  std::vector<std::thread> drivers;
  for (int k = 0; k < num_slots; ++k)
    drivers.emplace_back([&, k] {
      for (;;) {
        const size_t gi = next.fetch_add(1);
        if (gi >= groups.size()) break;
        try {
          GroupRunner runner(*slots[k], workers, cfg, outs, max_pair, max_only);
          runner.run(groups[gi]);
        } catch (...) {
          // a failure of the group as a whole (CUDA error, slab overflow): its unfinished meshes report it
          int st = DXO_ERR_INTERNAL;
          std::string what = "group failed";
          try { throw; } catch (const Error& e) { st = e.status; what = e.what(); } catch (const std::exception& e) { what = e.what(); } catch (...) {}
          for (GroupMesh* gm : groups[gi]) if (gm->alive) { gm->alive = false; gm->status = st; gm->error = what; }
          std::lock_guard<std::mutex> lock(err_mu);
          if (!first_error) first_error = std::current_exception();
          cudaGetLastError();
        }
      }
    });
  if (getenv("DXO_TIMING")) g_host_clock.reset();
  for (std::thread& t : drivers) t.join();
  if (getenv("DXO_TIMING")) {
    uint64_t verts = 0;
    for (size_t i = 0; i < n; ++i) if (all[i].status == DXO_OK && meshes[i].num_attributes) verts += meshes[i].attributes[0].num_unique_values;
    g_host_clock.report(verts);
  }
  for (size_t i = 0; i < n; ++i) {
    // a mesh that is still marked alive was never finished (its group died before stage 4)
    if (all[i].alive && all[i].status == DXO_OK && outs[i].data == nullptr) all[i].status = DXO_ERR_INTERNAL;
    if (statuses) statuses[i] = all[i].status;
    if (all[i].status != DXO_OK && getenv("DXO_DEBUG")) fprintf(stderr, "[dxo] batch mesh %zu: status %d (%s)\n", i, all[i].status, all[i].error.c_str());
  }
}

}  // namespace dxo
