// Product host code — see encoder.hpp. Mirrors the order of operations of
// encode::encode (encode/mod.rs:59-97), encode_attributes (encode/attribute/mod.rs:13-93)
// and AttributeEncoder (encode/attribute/attribute_encoder.rs:158-389); the per-element
// loops of the latter run as CUDA kernels (kernels.cu). No CPU implementation of those
// loops exists in this library: without a device the job fails.
#include "encoder.hpp"
#include <sys/mman.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <future>

namespace dxo {

void cuda_check(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return;
  const int st = (e == cudaErrorMemoryAllocation) ? DXO_ERR_OUT_OF_MEMORY
               : (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorInvalidDevice) ? DXO_ERR_NO_DEVICE
               : DXO_ERR_CUDA;
  throw Error(st, std::string(what) + ": " + cudaGetErrorString(e));
}

DeviceContext& DeviceContext::get(int device) {
  static thread_local std::map<int, std::unique_ptr<DeviceContext>> ctxs;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0) {
    cudaGetLastError();
    throw Error(DXO_ERR_NO_DEVICE, "no CUDA device available (this path has no CPU fallback)");
  }
  if (device < 0) cuda_check(cudaGetDevice(&device), "cudaGetDevice");
  if (device >= count) throw Error(DXO_ERR_NO_DEVICE, "CUDA device ordinal out of range");
  cuda_check(cudaSetDevice(device), "cudaSetDevice");
  auto it = ctxs.find(device);
  if (it != ctxs.end()) return *it->second;
  auto c = std::make_unique<DeviceContext>();
  c->device = device;
  for (auto& s : c->stream) cuda_check(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking), "cudaStreamCreate");
  cuda_check(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking), "cudaStreamCreate");
  cuda_check(cudaStreamCreateWithFlags(&c->upload_stream, cudaStreamNonBlocking), "cudaStreamCreate");
  cuda_check(cudaEventCreateWithFlags(&c->ev_uploaded, cudaEventDisableTiming), "cudaEventCreate");
  cuda_check(cudaEventCreateWithFlags(&c->ev_inputs, cudaEventDisableTiming), "cudaEventCreate");
  cuda_check(cudaEventCreateWithFlags(&c->ev_serial, cudaEventDisableTiming), "cudaEventCreate");
  cuda_check(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming), "cudaEventCreate");
  cuda_check(cudaEventCreateWithFlags(&c->ev_copy_join, cudaEventDisableTiming), "cudaEventCreate");
  const unsigned wait_flags = cudaEventDisableTiming | (getenv("DXO_BLOCKING_WAIT") ? (unsigned)cudaEventBlockingSync : 0u);
  for (auto& ev : c->ev_block) cuda_check(cudaEventCreateWithFlags(&ev, wait_flags), "cudaEventCreate");
  c->helpers.reset(new HelperThreads(2));
  cuda_check(cudaEventCreate(&c->ev_begin), "cudaEventCreate");
  cuda_check(cudaEventCreate(&c->ev_end), "cudaEventCreate");
  cuda_check(cudaEventCreateWithFlags(&c->ev_pos_ready, cudaEventDisableTiming), "cudaEventCreate");
  cuda_check(cudaEventCreateWithFlags(&c->ev_layout, cudaEventDisableTiming), "cudaEventCreate");
  for (auto& ev : c->ev_join) cuda_check(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming), "cudaEventCreate");
  // keep freed blocks in the stream-ordered pool: repeated encodes reuse them without cudaMalloc
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    uint64_t keep = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  DeviceContext& ref = *c;
  ctxs[device] = std::move(c);
  return ref;
}

// Process-wide pool of pinned host blocks (cudaHostAlloc is far too slow to call per mesh). Blocks are handed
// out best-fit and kept for the life of the process.
namespace {
struct PinnedPool {
  std::mutex mu;
  std::vector<std::pair<void*, size_t>> free_blocks;
  void* take(size_t bytes, size_t* cap) {
    {
      std::lock_guard<std::mutex> lock(mu);
      size_t best = free_blocks.size();
      for (size_t i = 0; i < free_blocks.size(); ++i)
        if (free_blocks[i].second >= bytes && free_blocks[i].second <= std::max<size_t>(2 * bytes + 4096, 2u << 20) &&  // blocks are >= 2 MB
            (best == free_blocks.size() || free_blocks[i].second < free_blocks[best].second)) best = i;
      if (best != free_blocks.size()) {
        auto b = free_blocks[best];
        free_blocks.erase(free_blocks.begin() + best);
        *cap = b.second;
        return b.first;
      }
    }
    // New block: 2 MB-aligned memory advised as transparent huge pages, touched, then pinned with cudaHostRegister — the
    // host walks over the tables in these blocks (CLERS traversal, sequencer) are random accesses over ~100 MB and miss the
    // TLB at every step with 4 KB pages. cudaHostAlloc is the fallback (and the choice with DXO_NO_HUGEPAGES=1).
    void* p = nullptr;
    static const bool huge = getenv("DXO_NO_HUGEPAGES") == nullptr;
    if (huge) {
      constexpr size_t kHuge = 2u << 20;
      const size_t want = (bytes + bytes / 8 + kHuge - 1) / kHuge * kHuge;
      p = aligned_alloc(kHuge, want);
      if (p) {
        madvise(p, want, MADV_HUGEPAGE);
        for (size_t o = 0; o < want; o += 4096) static_cast<volatile uint8_t*>(p)[o] = 0;
        if (cudaHostRegister(p, want, cudaHostRegisterPortable) == cudaSuccess) { *cap = want; return p; }
        cudaGetLastError();
        free(p);
        p = nullptr;
      }
    }
    const size_t want = bytes + bytes / 8 + 4096;
    if (cudaHostAlloc(&p, want, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    *cap = want;
    return p;
  }
  void give(void* p, size_t cap) {
    std::lock_guard<std::mutex> lock(mu);
    free_blocks.push_back({p, cap});
  }
};
PinnedPool& pinned_pool() { static PinnedPool* pool = new PinnedPool; return *pool; }  // never destroyed: outlives every job
}  // namespace
void* pinned_block_take(size_t bytes, size_t* capacity) { return pinned_pool().take(bytes, capacity); }
void pinned_block_give(void* p, size_t capacity) { pinned_pool().give(p, capacity); }

HelperThreads::HelperThreads(int n) { for (int i = 0; i < n; ++i) threads_.emplace_back([this] { loop(); }); }
HelperThreads::~HelperThreads() {
  { std::lock_guard<std::mutex> lock(mu_); stop_ = true; }
  cv_.notify_all();
  for (std::thread& t : threads_) t.join();
}
std::future<void> HelperThreads::run(std::function<void()> fn) {
  std::packaged_task<void()> task(std::move(fn));
  std::future<void> f = task.get_future();
  { std::lock_guard<std::mutex> lock(mu_); queue_.push_back(std::move(task)); }
  cv_.notify_one();
  return f;
}
void HelperThreads::loop() {
  for (;;) {
    std::packaged_task<void()> task;
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

void DeviceContext::wait_stream(int k) {
  cuda_check(cudaEventRecord(ev_block[k], stream[k]), "cudaEventRecord");
  cuda_check(cudaEventSynchronize(ev_block[k]), "cudaEventSynchronize");
}

DeviceContext::~DeviceContext() {
  helpers.reset();
  // Errors are ignored: at process exit the CUDA runtime may already be gone when the main thread's context is destroyed.
  if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return; }
  for (cudaStream_t s : {stream[0], stream[1], stream[2], copy_stream, upload_stream}) if (s) { cudaStreamSynchronize(s); cudaStreamDestroy(s); }
  for (cudaEvent_t e : {ev_begin, ev_end, ev_pos_ready, ev_layout, ev_join[0], ev_join[1], ev_join[2], ev_uploaded, ev_inputs, ev_serial, ev_fork, ev_copy_join, ev_block[0], ev_block[1], ev_block[2]}) if (e) cudaEventDestroy(e);
  // staging buffers go back to the process-wide pool: cudaFreeHost synchronises the whole device and stalled every other
  // thread's launches for hundreds of milliseconds when a worker thread ended while others were still encoding
  for (auto& b : pinned) if (b.first) pinned_pool().give(b.first, b.second);
  cudaGetLastError();
}

uint8_t* DeviceContext::pinned_buffer(size_t slot, size_t bytes) {
  if (pinned.size() <= slot) pinned.resize(slot + 1, {nullptr, 0});
  auto& b = pinned[slot];
  if (b.second < bytes) {
    if (b.first) pinned_pool().give(b.first, b.second);
    b = {nullptr, 0};
    size_t cap = 0;
    void* p = pinned_pool().take(bytes + bytes / 4 + 4096, &cap);
    if (!p) throw Error(DXO_ERR_OUT_OF_MEMORY, "pinned host allocation failed");
    b = {(uint8_t*)p, cap};
  }
  return b.first;
}

void* MeshJob::pinned_source(void* user, size_t bytes) {
  MeshJob* job = (MeshJob*)user;
  if (bytes < (1u << 16)) return nullptr;  // small arrays stay in ordinary memory
  size_t cap = 0;
  void* p = pinned_pool().take(bytes, &cap);
  if (!p) return nullptr;
  std::lock_guard<std::mutex> lock(job->alloc_mu_);
  job->pinned_blocks_.push_back({p, cap});
  return p;
}

cudaEvent_t Profile::take() {
  if (pool_used == pool.size()) { cudaEvent_t e; cuda_check(cudaEventCreate(&e), "cudaEventCreate"); pool.push_back(e); }
  return pool[pool_used++];
}
void Profile::begin(const char* name, uint64_t bytes, cudaStream_t s) {
  ++launches;
  if (!enabled) return;
  KernelRecord r{name, bytes, take(), take()};
  cuda_check(cudaEventRecord(r.a, s), "cudaEventRecord");
  records.push_back(r);
}
void Profile::end(cudaStream_t s) {
  if (!enabled) return;
  cuda_check(cudaEventRecord(records.back().b, s), "cudaEventRecord");
}
Profile::~Profile() { for (cudaEvent_t e : pool) cudaEventDestroy(e); }

// ---------------------------------------------------------------------------------------
namespace {

size_t component_size(uint32_t ct) {
  switch (ct) {
    case DXO_U8: case DXO_I8: return 1;
    case DXO_U16: case DXO_I16: return 2;
    case DXO_U32: case DXO_I32: case DXO_F32: return 4;
    case DXO_U64: case DXO_I64: case DXO_F64: return 8;
    default: return 0;
  }
}

int status_from_flags(uint32_t f) {
  if (f & gpu::kErrZeroNormal) return DXO_ERR_ZERO_NORMAL;
  if (f & gpu::kErrNegativeSymbol) return DXO_ERR_RANS_INVALID_SYMBOL;
  if (f & gpu::kErrRansFreq) return DXO_ERR_RANS_FREQ_TABLE;
  if (f & gpu::kErrRansState) return DXO_ERR_RANS_STATE_TOO_LARGE;
  if (f & (gpu::kErrAlphabet | gpu::kErrFanWalk)) return DXO_ERR_UNSUPPORTED_INPUT;
  return DXO_OK;
}

}  // namespace

MeshJob::MeshJob(const dxo_mesh* mesh, const dxo_config& cfg) : mesh_(mesh), cfg_(cfg) {
  graph_replay = (cfg.flags & DXO_FLAG_GRAPH_REPLAY) != 0;
  if (!mesh || (mesh->num_faces && !mesh->faces) || (mesh->num_attributes && !mesh->attributes))
    throw Error(DXO_ERR_INVALID_ARGUMENT, "null mesh / faces / attributes");
  if (mesh->num_faces == 0 || mesh->num_faces > 0x55555554ull) throw Error(DXO_ERR_INVALID_ARGUMENT, "face count must be in [1, 2^32/3)");
  if (mesh->num_attributes == 0 || mesh->attributes[0].att_type != DXO_ATT_POSITION)
    throw Error(DXO_ERR_UNSUPPORTED_INPUT, "attributes[0] must be the position attribute");  // edgebreaker.rs:132-134 unwrap
  if (mesh->num_attributes > 255) throw Error(DXO_ERR_TOO_MANY_ATTRIBUTES, "too many attributes");
  auto check_bits = [](uint32_t b) { if (b < 1 || b > 20) throw Error(DXO_ERR_INVALID_ARGUMENT, "quantization bits must be in [1, 20]"); };
  check_bits(cfg.position_bits); check_bits(cfg.texcoord_bits); check_bits(cfg.generic_bits);

  plans_.resize(mesh->num_attributes);
  for (uint32_t i = 0; i < mesh->num_attributes; ++i) {
    const dxo_attribute& a = mesh->attributes[i];
    AttrPlan& p = plans_[i];
    if (component_size(a.component_type) == 0) throw Error(DXO_ERR_UNSUPPORTED_DATA_TYPE, "unsupported component type");
    if (a.num_components < 1 || a.num_components > 4) throw Error(DXO_ERR_UNSUPPORTED_NUM_COMPONENTS, "attribute must have 1..4 components");
    if (!a.values && a.num_unique_values) throw Error(DXO_ERR_INVALID_ARGUMENT, "attribute without values");
    if (a.num_unique_values == 0 || a.num_unique_values > 0xFFFFFFFEull || a.num_points > 0xFFFFFFFEull)
      throw Error(DXO_ERR_INVALID_ARGUMENT, "attribute value / point count out of range");
    if (i > 0 && a.att_type == DXO_ATT_POSITION) throw Error(DXO_ERR_UNSUPPORTED_INPUT, "more than one position attribute");
    p.view.raw = &a;
    p.view.num_unique = (uint32_t)a.num_unique_values;
    p.view.map = a.point_to_value;
    p.view.num_points = a.point_to_value ? (uint32_t)a.num_points : (uint32_t)a.num_unique_values;  // Attribute::len()
    p.ncomp_in = a.num_components;
    // GroupConfig::default_for (attribute_encoder.rs:59-108) + portabilization::Config::default_for (portabilization/mod.rs:116-142)
    switch (a.att_type) {
      case DXO_ATT_POSITION: p.scheme = Scheme::Parallelogram; p.transform = Transform::Wrapped; p.port = Portabilization::Quantize; p.bits = cfg.position_bits; break;
      case DXO_ATT_NORMAL: p.scheme = Scheme::Normal; p.transform = Transform::OctOrthogonal; p.port = Portabilization::Octahedral; p.bits = 8; break;
      case DXO_ATT_TEXCOORD: p.scheme = Scheme::TexCoord; p.transform = Transform::Wrapped; p.port = Portabilization::Quantize; p.bits = cfg.texcoord_bits; break;
      case DXO_ATT_CUSTOM: p.scheme = Scheme::Parallelogram; p.transform = Transform::Wrapped; p.port = Portabilization::ToBits; break;
      default:
        if (a.att_type > DXO_ATT_WEIGHT) throw Error(DXO_ERR_INVALID_ARGUMENT, "unknown attribute type");
        p.scheme = Scheme::Delta; p.transform = Transform::Difference; p.port = Portabilization::Quantize; p.bits = cfg.generic_bits; break;
    }
    p.values32 = a.values;
    if (p.port == Portabilization::ToBits) {
      if (component_size(a.component_type) != 4) throw Error(DXO_ERR_UNSUPPORTED_DATA_TYPE, "ToBits attributes must have 4-byte components");
    } else if (a.component_type != DXO_F32) {
      // The reference's quantisers take every component type through DataValue::to_f64 and `as f32`
      // (quantization_coordinate_wise.rs:30-90; the glTF path creates u8 / u16 / u32 attributes such as JOINTS_0 and COLOR_0).
      // f64 normals are the exception: their octahedral map runs in f64 (geom.rs:84-88), which this path does not do.
      if (p.port == Portabilization::Octahedral && a.component_type == DXO_F64)
        throw Error(DXO_ERR_UNSUPPORTED_DATA_TYPE, "f64 normals are not supported (the reference maps them to the octahedron in f64)");
      const size_t cnt = (size_t)a.num_unique_values * a.num_components;
      p.converted.resize(cnt);
      const uint8_t* src = (const uint8_t*)a.values;
      for (size_t k = 0; k < cnt; ++k) {
        double d = 0;
        switch (a.component_type) {
          case DXO_U8: d = (double)src[k]; break;
          case DXO_I8: d = (double)((const int8_t*)src)[k]; break;
          case DXO_U16: { uint16_t v; memcpy(&v, src + 2 * k, 2); d = (double)v; break; }
          case DXO_I16: { int16_t v; memcpy(&v, src + 2 * k, 2); d = (double)v; break; }
          case DXO_U32: { uint32_t v; memcpy(&v, src + 4 * k, 4); d = (double)v; break; }
          case DXO_I32: { int32_t v; memcpy(&v, src + 4 * k, 4); d = (double)v; break; }
          case DXO_U64: { uint64_t v; memcpy(&v, src + 8 * k, 8); d = (double)v; break; }
          case DXO_I64: { int64_t v; memcpy(&v, src + 8 * k, 8); d = (double)v; break; }
          case DXO_F64: memcpy(&d, src + 8 * k, 8); break;
          default: break;
        }
        p.converted[k] = (float)d;
      }
      p.values32 = p.converted.data();
    }
    p.ncomp_q = p.port == Portabilization::Octahedral ? 2 : a.num_components;
    if (p.port == Portabilization::Octahedral && a.num_components != 3) throw Error(DXO_ERR_UNSUPPORTED_INPUT, "normals must have 3 components");
    if (p.scheme == Scheme::TexCoord && a.num_components != 2) throw Error(DXO_ERR_UNSUPPORTED_INPUT, "texture coordinates must have 2 components");
    // parents are looked up among already encoded attributes by id (encode/attribute/mod.rs:63-66)
    if (p.scheme == Scheme::Normal || p.scheme == Scheme::TexCoord) {
      if (!a.parent_ids || a.num_parents < 1 || (p.scheme == Scheme::Normal && a.num_parents != 1))
        throw Error(DXO_ERR_UNSUPPORTED_INPUT, "normal / texcoord attributes need a position parent");
      for (uint32_t j = 0; j < i; ++j) if (mesh->attributes[j].unique_id == a.parent_ids[0]) { p.parent = (int)j; break; }
      if (p.parent < 0) throw Error(DXO_ERR_UNSUPPORTED_INPUT, "parent attribute must precede its child");
      const dxo_attribute& par = mesh->attributes[p.parent];
      if (p.scheme == Scheme::Normal && par.att_type != DXO_ATT_POSITION) throw Error(DXO_ERR_UNSUPPORTED_INPUT, "normal parent must be the position attribute");
      if (par.num_components != 3 || plans_[p.parent].port != Portabilization::Quantize) throw Error(DXO_ERR_UNSUPPORTED_INPUT, "parent must be a quantized 3-component attribute");
    }
    // alphabet bound (DESIGN.md "Histogram capacity")
    if (p.transform == Transform::OctOrthogonal) p.hist_capacity = 512;
    else if (p.port == Portabilization::Quantize) p.hist_capacity = (2u << p.bits) + 4u;
    else {  // ToBits: bound from the raw value range
      const int32_t* v = (const int32_t*)a.values;
      int32_t mn = v[0], mx = v[0];
      const uint64_t cnt = a.num_unique_values * a.num_components;
      for (uint64_t k = 1; k < cnt; ++k) { mn = std::min(mn, v[k]); mx = std::max(mx, v[k]); }
      const uint64_t span = (uint64_t)((int64_t)mx - (int64_t)mn) + 1;
      if (span > (1ull << 22)) throw Error(DXO_ERR_UNSUPPORTED_INPUT, "custom attribute value range too large for the symbol histogram");
      p.hist_capacity = (uint32_t)(2 * span + 4);
    }
  }
}

MeshJob::~MeshJob() {
  if (graph_exec_) cudaGraphExecDestroy(graph_exec_);
  if (ev_graph_done_) cudaEventDestroy(ev_graph_done_);
  if (inputs_upload_.valid()) { try { inputs_upload_.wait(); } catch (...) {} }  // error paths: let the helper finish before freeing
  // normally released by release(); this covers error paths. cudaFree (synchronising) rather than cudaFreeAsync: the
  // stream the blocks were allocated on may belong to a thread that has exited
  for (void* p : allocations_) cudaFree(p);
  if (!pinned_blocks_.empty()) {
    // Every copy into the blocks is synchronised by the pass that issued it; only an exception between the enqueue and
    // that synchronisation leaves one in flight. The stream it ran on may belong to a thread that is gone by now, so
    // the whole device is drained in that case.
    if (pinned_copy_in_flight_.load()) cudaDeviceSynchronize();
    for (auto& b : pinned_blocks_) pinned_pool().give(b.first, b.second);
  }
  for (size_t i = 0; i < side_host_.size(); ++i) if (side_host_[i]) pinned_pool().give(side_host_[i], side_host_cap_[i]);
  for (cudaEvent_t e : side_ready_) if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : side_copied_) if (e) cudaEventDestroy(e);
}

// ---------------------------------------------------------------------------------------
namespace {
struct StageClock {
  bool on = getenv("DXO_TIMING") != nullptr;
  std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
  void lap(const char* what) {
    if (!on) return;
    auto n = std::chrono::steady_clock::now();
    fprintf(stderr, "[dxo] %-28s %9.3f ms\n", what, std::chrono::duration<double, std::milli>(n - t).count());
    t = n;
  }
};
}  // namespace

// K12 + K13: half-edge matching and left-most corners on the device, one synchronisation. Keeps the device
// copies of corner_vertex / opposite / left_most for the attribute kernels when the results are exact.
uint32_t MeshJob::device_matcher(void* user, const uint32_t* corner_vertex, uint32_t num_faces, uint32_t num_vertices, uint32_t* opposite_out,
                                 uint32_t* left_most_out, uint8_t* interior_out, std::vector<uint32_t>* boundary_corners) {
  MeshJob* job = (MeshJob*)user;
  DeviceContext& ctx = *job->match_ctx_;
  cudaStream_t s = ctx.stream[0];
  const size_t C = (size_t)num_faces * 3;
  StageClock clk;
  // vertex ids = the faces themselves (no position map): they are on their way to the device already
  uint32_t* d_cv = nullptr;
  if (corner_vertex == job->mesh_->faces && job->inputs_upload_.valid()) {
    job->inputs_upload_.wait();
    if (job->d_faces_) { cuda_check(cudaStreamWaitEvent(s, ctx.ev_inputs, 0), "cudaStreamWaitEvent"); d_cv = job->d_faces_; }
  }
  if (!d_cv) d_cv = job->dupload(corner_vertex, C, s);
  clk.lap("    K12 H2D corner_vertex");
  uint32_t* d_opp = job->dalloc<uint32_t>(C, s);
  uint32_t* d_lm = job->dalloc<uint32_t>(num_vertices, s);
  uint8_t* d_int = job->dalloc<uint8_t>(num_vertices, s);
  uint32_t* d_flag = job->dalloc<uint32_t>(3, s);  // [0] K12 not exact, [1] K13 fan flags, [2] number of boundary corners
  constexpr uint32_t kBoundaryPrefix = 1u << 18;   // boundary corners copied back together with the flags (1 MB)
  const size_t bb = gpu::boundary_list_scratch_bytes(C);
  uint32_t* d_blist = nullptr;
  void* bscratch = nullptr;
  cuda_check(cudaMallocAsync((void**)&d_blist, C * 4, s), "cudaMallocAsync");
  cuda_check(cudaMallocAsync(&bscratch, bb, s), "cudaMallocAsync");
  const size_t sb = gpu::corner_table_scratch_bytes(C), lb = gpu::left_most_scratch_bytes(num_vertices);
  void *scratch = nullptr, *lscratch = nullptr;
  cuda_check(cudaMallocAsync(&scratch, sb, s), "cudaMallocAsync");
  cuda_check(cudaMallocAsync(&lscratch, lb, s), "cudaMallocAsync");
  cuda_check(cudaMemsetAsync(d_flag, 0, 12, s), "cudaMemsetAsync");
  cuda_check(cudaMemsetAsync(d_opp, 0xFF, C * 4, s), "cudaMemsetAsync");
  gpu::launch_corner_table_opposites(d_cv, C, num_vertices, d_opp, d_flag, scratch, sb, s);
  gpu::launch_left_most(d_cv, d_opp, C, num_vertices, lscratch, d_lm, d_flag + 1, s, d_int);
  gpu::launch_boundary_list(d_opp, C, bscratch, bb, d_blist, d_flag + 2, s);
  uint32_t flag[3] = {1, 0, 0};
  if (clk.on) { cudaStreamSynchronize(s); clk.lap("    K12 + K13 kernels"); }
  job->pinned_copy_in_flight_.fetch_add(1);
  cuda_check(cudaMemcpyAsync(flag, d_flag, 12, cudaMemcpyDeviceToHost, s), "cudaMemcpyAsync D2H");
  const uint32_t prefix = (uint32_t)std::min<size_t>(kBoundaryPrefix, C);
  boundary_corners->resize(prefix);
  cuda_check(cudaMemcpyAsync(boundary_corners->data(), d_blist, (size_t)prefix * 4, cudaMemcpyDeviceToHost, s), "cudaMemcpyAsync D2H");
  // the results travel with the flags (one synchronisation); they are ignored when a flag is raised
  cuda_check(cudaMemcpyAsync(opposite_out, d_opp, C * 4, cudaMemcpyDeviceToHost, s), "cudaMemcpyAsync D2H");
  cuda_check(cudaMemcpyAsync(left_most_out, d_lm, (size_t)num_vertices * 4, cudaMemcpyDeviceToHost, s), "cudaMemcpyAsync D2H");
  cuda_check(cudaMemcpyAsync(interior_out, d_int, (size_t)num_vertices, cudaMemcpyDeviceToHost, s), "cudaMemcpyAsync D2H");
  cuda_check(cudaFreeAsync(scratch, s), "cudaFreeAsync");
  cuda_check(cudaFreeAsync(lscratch, s), "cudaFreeAsync");
  cuda_check(cudaFreeAsync(bscratch, s), "cudaFreeAsync");
  cuda_check(cudaStreamSynchronize(s), "cudaStreamSynchronize");
  job->pinned_copy_in_flight_.fetch_sub(1);
  clk.lap("    K12 D2H opposite, left_most");
  uint32_t done = 0;
  if (flag[1] & 1u) done |= UniversalTable::kUnusedVertices;
  if (flag[0] == 0 && flag[2] <= C) {  // boundary list: the part beyond the prefix needs a second copy (rare)
    boundary_corners->resize(flag[2]);
    if (flag[2] > prefix) {
      cuda_check(cudaMemcpyAsync(boundary_corners->data() + prefix, d_blist + prefix, (size_t)(flag[2] - prefix) * 4, cudaMemcpyDeviceToHost, s), "cudaMemcpyAsync D2H");
      cuda_check(cudaStreamSynchronize(s), "cudaStreamSynchronize");
    }
    done |= UniversalTable::kBoundaryListDone;
    job->d2h_bytes += (size_t)flag[2] * 4;
  }
  cuda_check(cudaFreeAsync(d_blist, s), "cudaFreeAsync");
  if (flag[0]) return done;  // order-dependent case: the sequential matcher decides
  done |= UniversalTable::kMatchExact;
  job->d2h_bytes += C * 4;
  job->d_corner_vertex_ = d_cv;
  job->d_opposite_ = d_opp;
  if (flag[1] == 0) {
    done |= UniversalTable::kLeftMostDone;
    job->d2h_bytes += (size_t)num_vertices * 5;
    job->d_left_most_ = d_lm;
  }
  return done;
}

void MeshJob::build_connectivity(DeviceContext* ctx) {
  const uint32_t nfaces = (uint32_t)mesh_->num_faces;
  StageClock clk;
  write_stream_header();

  match_ctx_ = ctx;
  dev_.assign(plans_.size(), AttrDevice{});
  const bool early_uploads = ctx != nullptr && parallel_host && !getenv("DXO_NO_EARLY_UPLOAD");
  const bool threads = !inline_host;
  if (early_uploads && threads)
    inputs_upload_ = std::async(std::launch::async, [this, ctx] { cuda_check(cudaSetDevice(ctx->device), "cudaSetDevice"); upload_inputs(*ctx); }).share();
  else if (early_uploads) {
    upload_inputs(*ctx);  // asynchronous copies on the upload stream
    std::promise<void> done;
    done.set_value();
    inputs_upload_ = done.get_future().share();
  }
  if (ctx != nullptr && !getenv("DXO_NO_PINNED_TABLES")) ut_.set_memory_source(&MeshJob::pinned_source, this);
  const bool use_k12 = ctx != nullptr && nfaces >= 4096 && !getenv("DXO_NO_K12");  // tiny meshes: the launch + sync costs more than it saves
  ut_.single_thread = inline_host;
  ut_.build(mesh_->faces, nfaces, plans_[0].view, use_k12 ? &MeshJob::device_matcher : nullptr, this);
  validate_attribute_indices(ut_.max_point);
  clk.lap("universal corner table");
  const size_t natt = plans_.size();
  seams_.resize(natt - 1);
  if (ctx != nullptr && !getenv("DXO_NO_PINNED_TABLES")) for (SeamTable& st : seams_) st.set_memory_source(&MeshJob::pinned_source, this);
  table_refs_.assign(natt, TableRef{});
  interior_.assign(natt, {});
  std::vector<ByteSink> seam_bytes(natt - 1);
  eb_.reset(new EdgebreakerEncoder(ut_));
  EdgebreakerEncoder& eb = *eb_;
  masked_opposite_.clear();
  masked_opposite_.resize(natt);
  if (!parallel_host || natt == 1) {
    for (size_t i = 1; i < natt; ++i) seams_[i - 1].build(ut_, plans_[i].view);
    clk.lap("seam tables");
    eb.traverse();
    eb.write_head(head_, seams_.size());
    for (size_t i = 1; i < natt; ++i) eb.write_seam_stream(seams_[i - 1], seam_bytes[i - 1]);
    clk.lap("edgebreaker");
    table_refs_[0] = table_ref(ut_);
    for (size_t i = 1; i < natt; ++i) table_refs_[i] = table_ref(ut_, seams_[i - 1]);
    for (size_t i = 0; i < natt; ++i) plans_[i].sequence = attribute_sequence(table_refs_[i], eb.corner_list());
    clk.lap("attribute sequences");
  } else {
    // Independent host passes on their own threads: seam tables (one per attribute) next to the
    // CLERS traversal; then the per-attribute sequencers and seam streams next to the symbol packing.
    std::vector<std::future<void>> tasks;
    auto spawn = [&](std::function<void()> fn) { if (threads) tasks.push_back(std::async(std::launch::async, std::move(fn))); else fn(); };
    auto wait_all = [&] { std::exception_ptr first; for (auto& t : tasks) { try { if (t.valid()) t.get(); } catch (...) { if (!first) first = std::current_exception(); } } tasks.clear(); if (first) std::rethrow_exception(first); };
    try {
      // One thread for everything (many encodes in flight): K14 of every attribute is started now and collected after
      // the traversal, which hides the device's turnaround behind 20+ ms of host work.
      std::vector<std::function<void()>> after_traversal;
      const bool defer_k14 = !threads && early_uploads && !getenv("DXO_NO_K14");
      if (defer_k14) for (size_t i = 1; i < natt; ++i) device_seam_table_begin(*ctx, i);
      auto spawn_att = [&](std::function<void()> fn) { if (defer_k14) after_traversal.push_back(std::move(fn)); else spawn(std::move(fn)); };
      for (size_t i = 1; i < natt; ++i)
        spawn_att([this, i, ctx, early_uploads] {
          StageClock c;
          if (early_uploads && !getenv("DXO_NO_K14") && device_seam_table(*ctx, i)) c.lap("  (thread) seam table (K14)");
          else {
            seams_[i - 1].build(ut_, plans_[i].view);
            c.lap("  (thread) seam table");
            if (early_uploads) { cuda_check(cudaSetDevice(ctx->device), "cudaSetDevice"); upload_seam_table(*ctx, i); c.lap("  (thread) seam table upload"); }
          }
          table_refs_[i] = table_ref(ut_, seams_[i - 1]);
          if (seams_[i - 1].has_interior_seam) {  // the sequencer of this table runs: give it one array for opp()
            const uint32_t C = ut_.num_corners;
            U32Array& mo = masked_opposite_[i];
            mo.resize(C);
            const uint32_t* opp = ut_.opposite.data();
            const uint8_t* sm = seams_[i - 1].seam.data();
            for (uint32_t k = 0; k < C; ++k) mo[k] = sm[k] ? kNone : opp[k];
            table_refs_[i].opposite_masked = mo.data();
          }
          // the attribute's own sequencer runs only when its table differs from the universal one (see shares_position_sequence)
          if (seams_[i - 1].has_interior_seam || seams_[i - 1].num_vertices != ut_.num_vertices || getenv("DXO_NO_SHARED_SEQUENCE")) {
            interior_[i] = vertex_interior_flags(table_refs_[i]);
            table_refs_[i].interior = interior_[i].data();
          }
          c.lap("  (thread) masked opposites, interior flags");
        });
      table_refs_[0] = table_ref(ut_);
      if (ut_.has_interior) table_refs_[0].interior = ut_.interior.data();  // came back with K13's left-most corners
      else spawn([this] { interior_[0] = vertex_interior_flags(table_refs_[0]); table_refs_[0].interior = interior_[0].data(); });
      { StageClock c; eb.traverse(); c.lap("  (main) CLERS traversal"); }
      for (auto& fn : after_traversal) fn();
      wait_all();  // seam tables, interior flags
      clk.lap("seam tables + traversal");
      auto seq0_fn = [this] { StageClock c; plans_[0].sequence = attribute_sequence(table_refs_[0], eb_->corner_list()); c.lap("  (thread) position sequence"); };
      std::future<void> seq0;
      if (threads) seq0 = std::async(std::launch::async, seq0_fn); else seq0_fn();
      // an attribute whose only seams are mesh boundaries has the universal table (same ids, opposites, left-most
      // corners), hence the position sequence: it is copied instead of recomputed
      auto shares_position_sequence = [this](size_t i) {
        const SeamTable& st = seams_[i - 1];
        return !st.has_interior_seam && st.num_vertices == ut_.num_vertices && !getenv("DXO_NO_SHARED_SEQUENCE");
      };
      for (size_t i = 1; i < natt; ++i) {
        if (!shares_position_sequence(i))
          spawn([this, i] { StageClock c; plans_[i].sequence = attribute_sequence(table_refs_[i], eb_->corner_list()); c.lap("  (thread) attribute sequence"); });
        spawn([this, i, &eb, &seam_bytes] { StageClock c; eb.write_seam_stream(seams_[i - 1], seam_bytes[i - 1]); c.lap("  (thread) seam stream"); });
      }
      if (seq0.valid()) tasks.push_back(std::move(seq0));
      { StageClock c; eb.write_head(head_, seams_.size()); c.lap("  (main) connectivity head"); }
      wait_all();
      for (size_t i = 1; i < natt; ++i) if (shares_position_sequence(i)) plans_[i].shares_sequence_of = 0;
      clk.lap("sequences + streams");
    } catch (...) {
      try { wait_all(); } catch (...) {}
      throw;
    }
  }
  for (const ByteSink& b : seam_bytes) head_.bytes(b.data);
  write_attribute_section_headers();
  for (size_t i = 0; i < natt; ++i) plans_[i].table = &table_refs_[i];
}

// header (encode/header/mod.rs:26-54)
void MeshJob::write_stream_header() {
  for (char ch : std::string("DRACO")) head_.u8((uint8_t)ch);
  head_.u8(2); head_.u8(2);
  head_.u8(1);    // EncodedGeometryType::TrianglarMesh
  head_.u8(1);    // EncoderMethod::Edgebreaker
  head_.u16(0);   // flags: no metadata
}

// Every attribute must cover the points the faces use, and every point -> value entry must name an existing value:
// the predict kernels gather quant[map[point]] without further checks (the reference would panic on the slice index).
void MeshJob::validate_attribute_indices(uint32_t max_face_point) const {
  for (size_t i = 0; i < plans_.size(); ++i) {
    const AttrView& v = plans_[i].view;
    if (v.num_points <= max_face_point) throw Error(DXO_ERR_INVALID_ARGUMENT, "face references a point outside an attribute");
    if (i == 0 || !v.map) continue;  // the position map is checked by the corner-table build
    const uint32_t mx = max_u32(v.map, v.num_points);
    if (v.num_points && mx >= v.num_unique) throw Error(DXO_ERR_INVALID_ARGUMENT, "point_to_value entry out of range");
  }
}

// attribute section headers (encode/attribute/mod.rs:26-57)
void MeshJob::write_attribute_section_headers() {
  head_.u8((uint8_t)plans_.size());
  for (size_t i = 0; i < plans_.size(); ++i) {
    head_.u8((uint8_t)((uint8_t)i - 1u));
    head_.u8((uint8_t)plans_[i].view.raw->domain);
    head_.u8(0);  // TraversalType::DepthFirst
  }
  for (const AttrPlan& p : plans_) {
    const dxo_attribute& a = *p.view.raw;
    head_.u8(1);
    head_.u8((uint8_t)a.att_type);
    head_.u8((uint8_t)a.component_type);
    head_.u8((uint8_t)a.num_components);
    head_.u8(0);
    head_.u8((uint8_t)a.unique_id);
    head_.u8((uint8_t)p.port);
  }
}

// ---------------------------------------------------------------------------------------
template <class T> T* MeshJob::dalloc(size_t count, cudaStream_t s) {
  void* p = nullptr;
  cuda_check(cudaMallocAsync(&p, std::max<size_t>(count, 1) * sizeof(T), s), "cudaMallocAsync");
  std::lock_guard<std::mutex> lock(alloc_mu_);
  allocations_.push_back(p);
  alloc_stream_ = s;
  return (T*)p;
}
template <class T> T* MeshJob::dupload(const T* host, size_t count, cudaStream_t s) {
  T* d = dalloc<T>(count, s);
  if (count) cuda_check(cudaMemcpyAsync(d, host, count * sizeof(T), cudaMemcpyHostToDevice, s), "cudaMemcpyAsync H2D");
  std::lock_guard<std::mutex> lock(alloc_mu_);
  h2d_bytes += count * sizeof(T);
  return d;
}

// Inputs that do not depend on the connectivity (called on a helper thread at the start of build_connectivity,
// or from upload() when nothing was started early).
void MeshJob::upload_inputs(DeviceContext& ctx) {
  cudaStream_t s = ctx.upload_stream;
  d_faces_ = dupload(mesh_->faces, (size_t)mesh_->num_faces * 3, s);
  for (size_t i = 0; i < plans_.size(); ++i) {
    const AttrPlan& p = plans_[i];
    const dxo_attribute& a = *p.view.raw;
    (void)a;
    dev_[i].values = (float*)dupload((const uint32_t*)p.values32, (size_t)p.view.num_unique * p.ncomp_in, s);
    if (p.view.map) dev_[i].map = dupload(p.view.map, p.view.num_points, s);
  }
  cuda_check(cudaEventRecord(ctx.ev_inputs, s), "cudaEventRecord");
}

// K14: the seam table of attribute `att` from the device-resident universal table (K12 + K13 results).
// Runs on the helper thread of that attribute; the host copies are needed by the sequencer and the seam stream.
// First half: the kernels only (asynchronous) — a caller that keeps its host passes on one thread starts every attribute's
// table before the CLERS traversal and collects the results after it.
bool MeshJob::device_seam_table_begin(DeviceContext& ctx, size_t att) {
  if (seam_pending_.size() != plans_.size()) seam_pending_.assign(plans_.size(), SeamPending{});
  if (!d_opposite_ || !d_corner_vertex_ || !d_left_most_ || !inputs_upload_.valid()) return false;
  cuda_check(cudaSetDevice(ctx.device), "cudaSetDevice");
  inputs_upload_.wait();
  cudaStream_t s = ctx.stream[std::min<size_t>(att, 2)];
  cuda_check(cudaStreamWaitEvent(s, ctx.ev_inputs, 0), "cudaStreamWaitEvent");
  const size_t C = ut_.num_corners;
  const uint32_t V = ut_.num_vertices;
  AttrDevice& d = dev_[att];
  if (!d_faces_) return false;  // the inputs upload failed; upload() reports it
  SeamPending& sp = seam_pending_[att];
  sp.d_seam = dalloc<uint8_t>(C, s);
  sp.d_cv = dalloc<uint32_t>(C, s);
  sp.d_lm = dalloc<uint32_t>(C, s);  // at most one attribute vertex per corner
  sp.d_scalars = dalloc<uint32_t>(2, s);  // [0] number of attribute vertices, [1] flags
  const size_t sb = gpu::seam_table_scratch_bytes(V);
  void* scratch = nullptr;
  cuda_check(cudaMallocAsync(&scratch, sb, s), "cudaMallocAsync");
  cuda_check(cudaMemsetAsync(sp.d_scalars, 0, 8, s), "cudaMemsetAsync");
  gpu::launch_seam_table(d_faces_, d.map, plans_[att].view.num_points, d_corner_vertex_, d_opposite_, d_left_most_, C, V, scratch, sb, sp.d_seam, sp.d_cv,
                         sp.d_lm, sp.d_scalars, sp.d_scalars + 1, s);
  cuda_check(cudaFreeAsync(scratch, s), "cudaFreeAsync");
  sp.begun = true;
  return true;
}

bool MeshJob::device_seam_table(DeviceContext& ctx, size_t att) {
  if (seam_pending_.size() != plans_.size() || !seam_pending_[att].begun) {
    if (!device_seam_table_begin(ctx, att)) return false;
  }
  cuda_check(cudaSetDevice(ctx.device), "cudaSetDevice");
  cudaStream_t s = ctx.stream[std::min<size_t>(att, 2)];
  const size_t C = ut_.num_corners;
  const uint32_t V = ut_.num_vertices;
  AttrDevice& d = dev_[att];
  const SeamPending sp = seam_pending_[att];
  uint8_t* const d_seam = sp.d_seam;
  uint32_t *const d_cv = sp.d_cv, *const d_lm = sp.d_lm, *const d_scalars = sp.d_scalars;
  uint32_t scalars[2] = {0, 1};
  cuda_check(cudaMemcpyAsync(scalars, d_scalars, 8, cudaMemcpyDeviceToHost, s), "cudaMemcpyAsync D2H");
  cuda_check(cudaStreamSynchronize(s), "cudaStreamSynchronize");
  if ((scalars[1] & 3u) != 0 || scalars[0] > C) return false;  // the sequential pass reports the error
  SeamTable& st = seams_[att - 1];
  st.num_vertices = scalars[0];
  st.has_interior_seam = (scalars[1] & 4u) != 0;
  d.corner_vertex = d_cv; d.seam = d_seam; d.left_most = d_lm;
  // Only mesh boundaries are seams and no vertex was split: the table IS the universal one. The host then needs none of
  // it — the attribute shares the position sequence and its seam stream is a run of zeros — so the 5 bytes per corner
  // stay on the device (traces and DXO_NO_SHARED_SEQUENCE still fetch them).
  if (!st.has_interior_seam && st.num_vertices == V && !trace && !getenv("DXO_NO_SHARED_SEQUENCE")) return true;
  st.corner_vertex.resize(C);
  st.seam.resize(C);
  st.left_most.resize(st.num_vertices);
  pinned_copy_in_flight_.fetch_add(1);
  cuda_check(cudaMemcpyAsync(st.corner_vertex.data(), d_cv, C * 4, cudaMemcpyDeviceToHost, s), "cudaMemcpyAsync D2H");
  cuda_check(cudaMemcpyAsync(st.seam.data(), d_seam, C, cudaMemcpyDeviceToHost, s), "cudaMemcpyAsync D2H");
  cuda_check(cudaMemcpyAsync(st.left_most.data(), d_lm, (size_t)st.num_vertices * 4, cudaMemcpyDeviceToHost, s), "cudaMemcpyAsync D2H");
  cuda_check(cudaStreamSynchronize(s), "cudaStreamSynchronize");
  pinned_copy_in_flight_.fetch_sub(1);
  { std::lock_guard<std::mutex> lock(alloc_mu_); d2h_bytes += C * 5 + (size_t)st.num_vertices * 4; }
  return true;
}
// Seam table of attribute `att` (called by the thread that has just built it).
void MeshJob::upload_seam_table(DeviceContext& ctx, size_t att) {
  cudaStream_t s = ctx.upload_stream;
  const SeamTable& st = seams_[att - 1];
  AttrDevice& d = dev_[att];
  d.corner_vertex = dupload(st.corner_vertex.data(), st.corner_vertex.size(), s);
  d.seam = dupload(st.seam.data(), st.seam.size(), s);
  d.left_most = dupload(st.left_most.data(), st.left_most.size(), s);
}

void MeshJob::upload(DeviceContext& ctx) {
  cudaStream_t s = ctx.stream[0];
  const size_t C = ut_.num_corners;
  if (dev_.size() != plans_.size()) dev_.assign(plans_.size(), AttrDevice{});
  if (inputs_upload_.valid()) inputs_upload_.get();
  else if (!d_faces_) upload_inputs(ctx);
  for (size_t i = 1; i < plans_.size(); ++i) if (!dev_[i].corner_vertex) upload_seam_table(ctx, i);
  cuda_check(cudaEventRecord(ctx.ev_uploaded, ctx.upload_stream), "cudaEventRecord");
  cuda_check(cudaStreamWaitEvent(s, ctx.ev_uploaded, 0), "cudaStreamWaitEvent");
  if (!d_opposite_) d_opposite_ = dupload(ut_.opposite.data(), C, s);
  if (!d_corner_vertex_) d_corner_vertex_ = dupload(ut_.corner_vertex.data(), C, s);
  else if (ut_.num_vertices != plans_[0].view.num_unique) {  // non-manifold vertices were split after K12 ran: refresh the labels
    if (d_corner_vertex_ == d_faces_) d_corner_vertex_ = dupload(ut_.corner_vertex.data(), C, s);  // K12 read the faces in place
    else cuda_check(cudaMemcpyAsync(d_corner_vertex_, ut_.corner_vertex.data(), C * 4, cudaMemcpyHostToDevice, s), "cudaMemcpyAsync H2D");
  }
  if (!d_left_most_) d_left_most_ = dupload(ut_.left_most.data(), ut_.left_most.size(), s);
  // padded per-face layouts are derived on the device at the start of every launch() (part of the timed step)
  d_faces4_ = dalloc<uint4>(ut_.num_faces, s);
  vertex_is_point_ = plans_[0].view.map == nullptr && ut_.num_vertices == plans_[0].view.num_unique;
  if (!vertex_is_point_) d_corner_vertex4_ = dalloc<uint4>(ut_.num_faces, s);
  for (size_t i = 0; i < plans_.size(); ++i) {
    const AttrPlan& p = plans_[i];
    AttrDevice& d = dev_[i];
    const size_t U = p.view.num_unique, M = sequence_of(i).size();
    d.seq = p.shares_sequence_of >= 0 ? dev_[p.shares_sequence_of].seq : dupload(p.sequence.data(), M, s);
    const uint32_t V = p.table->num_vertices;
    const size_t qstride = p.ncomp_q == 3 ? 4 : p.ncomp_q;  // one value = one vector load (kernels.cu load_q)
    if (p.port == Portabilization::ToBits && p.ncomp_q != 3) d.quant = (int32_t*)d.values;
    else d.quant = dalloc<int32_t>(U * qstride, s);
    if (i > 0) d.corner_vertex4 = dalloc<uint4>(ut_.num_faces, s);
    // Resident sessions resolve the predictors' operand indices once per upload (rings / records, kernels.cu) instead of
    // chasing them through the tables in every step; DXO_NO_RINGS keeps the per-step forms.
    const bool records = resident && !getenv("DXO_NO_RINGS");
    if (i > 0 && p.scheme == Scheme::Normal) d.fan_link = dalloc<uint2>(C, s);  // fan walks read one link per swing
    if (i > 0 && p.scheme == Scheme::Normal && records) {
      d.ring = dalloc<uint4>(2 * M, s); d.ring_head = dalloc<uint2>(M, s); d.ring_count = dalloc<uint8_t>(M, s);
    }
    if (i > 0 && p.scheme == Scheme::TexCoord && records) d.tex_records = dalloc<uint4>(2 * M, s);
    if (p.scheme == Scheme::Parallelogram && records) d.para_records = dalloc<uint4>(M, s);
    // K4's fast path: {opposite, its point} links and ranks carried in the values' padding component
    if (i == 0) rank_in_w_ = false;
    if (i == 0 && !records && vertex_is_point_ && p.scheme == Scheme::Parallelogram && p.port == Portabilization::Quantize && p.ncomp_q == 3 && !getenv("DXO_NO_K4_FAST")) {
      d.fan_link = dalloc<uint2>(C, s);
      rank_in_w_ = true;
    }
    d.rank = dalloc<uint32_t>(V, s);
    if (p.transform == Transform::Wrapped) d.used = dalloc<uint8_t>(U, s);
    d.symbols = dalloc<uint32_t>(M * p.ncomp_q, s);
    d.side = dalloc<uint8_t>(M, s);
    if (p.scheme == Scheme::Normal || p.scheme == Scheme::TexCoord) {
      d.side_out = dalloc<uint8_t>(M + 16, s);
      if (p.scheme == Scheme::TexCoord) { d.side_scratch_bytes = gpu::side_prepare_scratch_bytes((uint32_t)M); d.side_scratch = dalloc<uint8_t>(d.side_scratch_bytes, s); }
    }
    d.hist = dalloc<uint32_t>(p.hist_capacity, s);
    d.work = dalloc<uint32_t>(3 * (size_t)p.hist_capacity, s);
    d.rans_table = dalloc<uint4>(p.hist_capacity + 1, s);
    d.table_capacity = 3 * p.hist_capacity + 16;
    d.table_bytes = dalloc<uint8_t>(d.table_capacity, s);
    d.payload_capacity = 3 * (uint64_t)M * p.ncomp_q + 16;  // <= P/8 <= 2.5 bytes per symbol + tail
    d.payload = dalloc<uint8_t>(d.payload_capacity, s);
    d.rans_scratch = dalloc<uint8_t>(gpu::rans_scratch_bytes((uint64_t)M * p.ncomp_q), s);
    d.stats = dalloc<gpu::AttrStats>(1, s);
  }
  // rank / used tables of the sequences: functions of the connectivity, computed here once (see launch_sequence_tables)
  for (size_t i = 0; i < plans_.size(); ++i) {
    const AttrPlan& p = plans_[i];
    AttrDevice& d = dev_[i];
    const bool needs_rank = p.scheme == Scheme::Parallelogram || p.scheme == Scheme::TexCoord;
    if (!needs_rank && !d.used) continue;
    if (needs_rank) cuda_check(cudaMemsetAsync(d.rank, 0xFF, sizeof(uint32_t) * p.table->num_vertices, s), "cudaMemsetAsync");
    if (d.used) cuda_check(cudaMemsetAsync(d.used, 0, p.view.num_unique, s), "cudaMemsetAsync");
    gpu::launch_sequence_tables(d.seq, (uint32_t)sequence_of(i).size(), table_dev(i), d.map, needs_rank ? d.rank : nullptr, d.used, s);
  }
  // Layouts derived from the tables alone — per-face corner tuples (one 128-bit load per face in the predictors), the fan
  // links of K5, 3-component ToBits values padded to 4 — are built here, once per mesh, not in every step.
  gpu::launch_pad3(d_faces_, ut_.num_faces, d_faces4_, s);
  if (!vertex_is_point_) gpu::launch_pad3(d_corner_vertex_, ut_.num_faces, d_corner_vertex4_, s);
  for (size_t i = 0; i < plans_.size(); ++i) {
    const AttrPlan& p = plans_[i];
    AttrDevice& d = dev_[i];
    if (i > 0) gpu::launch_pad3(d.corner_vertex, ut_.num_faces, d.corner_vertex4, s);
    if (d.fan_link) gpu::launch_fan_link(d_opposite_, i == 0 ? nullptr : d.seam, d_faces_, C, d.fan_link, s);
    if (p.port == Portabilization::ToBits && p.ncomp_q == 3) gpu::launch_pad3((const uint32_t*)d.values, p.view.num_unique, (uint4*)d.quant, s);
  }
  uint32_t ring_launches = 0;
  for (size_t i = 1; i < plans_.size(); ++i) {
    AttrDevice& d = dev_[i];
    if (!d.ring) continue;
    const AttrDevice& pd = dev_[plans_[i].parent];
    gpu::TableDev t = table_dev(i);
    t.ring = nullptr;
    gpu::launch_normal_rings(d.seq, (uint32_t)sequence_of(i).size(), t, gpu::QuantDev{nullptr, d.map, (uint32_t)plans_[i].ncomp_q}, gpu::QuantDev{nullptr, pd.map, 3},
                             d.ring, d.ring_head, d.ring_count, s);
    ++ring_launches;
  }
  for (size_t i = 1; i < plans_.size(); ++i) {
    AttrDevice& d = dev_[i];
    if (!d.tex_records) continue;
    const AttrDevice& pd = dev_[plans_[i].parent];
    gpu::launch_texcoord_records(d.seq, (uint32_t)sequence_of(i).size(), table_dev(i), gpu::QuantDev{nullptr, d.map, (uint32_t)plans_[i].ncomp_q},
                                 gpu::QuantDev{nullptr, pd.map, 3}, plans_[plans_[i].parent].view.num_points, d.rank, d.tex_records, s);
    ++ring_launches;
  }
  for (size_t i = 0; i < plans_.size(); ++i) {
    AttrDevice& d = dev_[i];
    if (!d.para_records) continue;
    gpu::launch_parallelogram_records(d.seq, (uint32_t)sequence_of(i).size(), table_dev(i), gpu::QuantDev{nullptr, d.map, (uint32_t)plans_[i].ncomp_q}, d.rank,
                                      d.para_records, s);
    ++ring_launches;
  }
  layout_launches_ = 1 + (vertex_is_point_ ? 0 : 1) + ring_launches;
  for (size_t i = 0; i < plans_.size(); ++i) layout_launches_ += (plans_[i].scheme == Scheme::Parallelogram || plans_[i].scheme == Scheme::TexCoord || dev_[i].used) ? 1 : 0;
  for (size_t i = 0; i < plans_.size(); ++i)
    layout_launches_ += (i > 0 ? 1 : 0) + (dev_[i].fan_link ? 1 : 0) + ((plans_[i].port == Portabilization::ToBits && plans_[i].ncomp_q == 3) ? 1 : 0);
  cuda_check(cudaEventRecord(ctx.ev_join[0], s), "cudaEventRecord");
  for (int k = 1; k < 3; ++k) cuda_check(cudaStreamWaitEvent(ctx.stream[k], ctx.ev_join[0], 0), "cudaStreamWaitEvent");
  cuda_check(cudaStreamWaitEvent(ctx.copy_stream, ctx.ev_join[0], 0), "cudaStreamWaitEvent");
  side_ready_.assign(plans_.size(), nullptr);
  side_copied_.assign(plans_.size(), nullptr);
  side_host_.assign(plans_.size(), nullptr);
  side_host_cap_.assign(plans_.size(), 0);
  for (size_t i = 0; i < plans_.size(); ++i) {
    if (plans_[i].scheme != Scheme::Normal && plans_[i].scheme != Scheme::TexCoord) continue;
    side_host_[i] = (uint8_t*)pinned_pool().take(sequence_of(i).size() + 16, &side_host_cap_[i]);
    if (!side_host_[i]) throw Error(DXO_ERR_OUT_OF_MEMORY, "pinned host allocation failed");
    cuda_check(cudaEventCreateWithFlags(&side_ready_[i], cudaEventDisableTiming), "cudaEventCreate");
    cuda_check(cudaEventCreateWithFlags(&side_copied_[i], cudaEventDisableTiming | (getenv("DXO_BLOCKING_WAIT") ? (unsigned)cudaEventBlockingSync : 0u)), "cudaEventCreate");
  }
  uploaded_ = true;
}

gpu::TableDev MeshJob::table_dev(size_t att) const {
  gpu::TableDev t;
  t.corner_point = d_faces_;
  t.corner_point4 = d_faces4_;
  t.opposite = d_opposite_;
  t.fan_link = dev_[att].fan_link;
  t.ring = dev_[att].ring; t.ring_head = dev_[att].ring_head; t.ring_count = dev_[att].ring_count;
  t.num_corners = ut_.num_corners;
  if (att == 0) {
    t.corner_vertex = d_corner_vertex_; t.corner_vertex4 = vertex_is_point_ ? d_faces4_ : d_corner_vertex4_; t.vertex_is_point = vertex_is_point_ ? 1 : 0;
    t.seam = nullptr; t.left_most = d_left_most_; t.num_vertices = ut_.num_vertices;
  } else {
    t.corner_vertex = dev_[att].corner_vertex; t.corner_vertex4 = dev_[att].corner_vertex4; t.vertex_is_point = 0;
    t.seam = dev_[att].seam; t.left_most = dev_[att].left_most; t.num_vertices = seams_[att - 1].num_vertices;
  }
  return t;
}

// Algorithmic bytes per launch follow SURVEY.md §8(d): each distinct input array read
// once, each output written once.
void MeshJob::launch_all(DeviceContext& ctx, Profile& prof, bool capturing) {
  cudaStream_t s0 = ctx.stream[0];
  cuda_check(cudaEventRecord(ctx.ev_fork, s0), "cudaEventRecord");
  for (int k = 1; k < 3; ++k) cuda_check(cudaStreamWaitEvent(ctx.stream[k], ctx.ev_fork, 0), "cudaStreamWaitEvent");
  capturing_ = capturing;
  launch(ctx, prof);
  capturing_ = false;
  for (int k = 1; k < 3; ++k) {
    cuda_check(cudaEventRecord(ctx.ev_join[k], ctx.stream[k]), "cudaEventRecord");
    cuda_check(cudaStreamWaitEvent(s0, ctx.ev_join[k], 0), "cudaStreamWaitEvent");
  }
  bool any_side = false;
  for (cudaEvent_t e : side_copied_) any_side |= e != nullptr;
  if (any_side) {  // the copy stream took part (flag copies): it joins the origin stream as well
    cuda_check(cudaEventRecord(ctx.ev_copy_join, ctx.copy_stream), "cudaEventRecord");
    cuda_check(cudaStreamWaitEvent(s0, ctx.ev_copy_join, 0), "cudaStreamWaitEvent");
  }
}

void MeshJob::launch_graph(DeviceContext& ctx, Profile& prof) {
  cudaStream_t s0 = ctx.stream[0];
  if (!graph_exec_ || graph_ctx_ != &ctx) {
    if (graph_exec_) { cudaGraphExecDestroy(graph_exec_); graph_exec_ = nullptr; }
    if (!ev_graph_done_) cuda_check(cudaEventCreateWithFlags(&ev_graph_done_, cudaEventDisableTiming), "cudaEventCreate");
    for (int k = 0; k < 3; ++k) cuda_check(cudaStreamSynchronize(ctx.stream[k]), "cudaStreamSynchronize");
    cuda_check(cudaStreamSynchronize(ctx.copy_stream), "cudaStreamSynchronize");
    cudaGraph_t graph = nullptr;
    cuda_check(cudaStreamBeginCapture(s0, cudaStreamCaptureModeThreadLocal), "cudaStreamBeginCapture");
    try {
      const uint64_t d2h_before = d2h_bytes;
      launch_all(ctx, prof, true);
      graph_d2h_bytes_ = d2h_bytes - d2h_before;
    } catch (...) {
      cudaStreamEndCapture(s0, &graph);
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      throw;
    }
    cuda_check(cudaStreamEndCapture(s0, &graph), "cudaStreamEndCapture");
    graph_launches_ = prof.launches;
    const cudaError_t e = cudaGraphInstantiate(&graph_exec_, graph, 0);
    cudaGraphDestroy(graph);
    cuda_check(e, "cudaGraphInstantiate");
    graph_ctx_ = &ctx;
  } else {
    d2h_bytes += graph_d2h_bytes_;
  }
  prof.launches = graph_launches_;
  cuda_check(cudaGraphLaunch(graph_exec_, s0), "cudaGraphLaunch");
  // later work on the attribute streams (result copies) is ordered after the graph
  cuda_check(cudaEventRecord(ev_graph_done_, s0), "cudaEventRecord");
  for (int k = 1; k < 3; ++k) cuda_check(cudaStreamWaitEvent(ctx.stream[k], ev_graph_done_, 0), "cudaStreamWaitEvent");
}

void MeshJob::launch(DeviceContext& ctx, Profile& prof) {
  const uint64_t C = ut_.num_corners;
  const uint64_t Upos = plans_[0].view.num_unique;
  if (device_runs <= 1) prof.launches += layout_launches_;  // issued by upload(): counted with the first step
  for (size_t i = 0; i < plans_.size(); ++i) {
    const AttrPlan& p = plans_[i];
    AttrDevice& d = dev_[i];
    cudaStream_t s = ctx.stream[std::min<size_t>(i, 2)];
    const uint64_t U = p.view.num_unique, V = p.table->num_vertices;
    const uint32_t M = (uint32_t)sequence_of(i).size();
    const uint64_t S = (uint64_t)M * p.ncomp_q;
    const gpu::TableDev t = table_dev(i);
    gpu::QuantDev q{d.quant, d.map, p.ncomp_q};
    q.rank_in_w = (i == 0 && rank_in_w_) ? 1u : 0u;

    if (prof.serial && i > 0) cuda_check(cudaStreamWaitEvent(s, ctx.ev_serial, 0), "cudaStreamWaitEvent");
    gpu::init_stats(d.stats, s);
    ++prof.launches;
    cuda_check(cudaMemsetAsync(d.hist, 0, sizeof(uint32_t) * p.hist_capacity, s), "cudaMemsetAsync");
    const bool wrapped = p.transform == Transform::Wrapped;

    if (p.port == Portabilization::Quantize) {
      prof.begin("K1_minmax", 4 * p.ncomp_in * U, s);
      gpu::launch_minmax(d.values, U, p.ncomp_in, d.stats, s);
      prof.end(s);
      prof.begin("K2_quantize", 8 * p.ncomp_in * U, s);
      gpu::launch_quantize(d.values, U, p.ncomp_in, p.bits, d.quant, d.stats, s, q.rank_in_w ? -1 : 0, wrapped ? d.used : nullptr, q.rank_in_w ? d.rank : nullptr);
      prof.end(s);
    } else if (p.port == Portabilization::Octahedral) {
      prof.begin("K3_oct_quantize", 20 * U, s);
      gpu::launch_oct_quantize(d.values, U, d.quant, d.stats, s);
      prof.end(s);
    }
    if (i == 0) cuda_check(cudaEventRecord(ctx.ev_pos_ready, s), "cudaEventRecord");
    if (p.parent >= 0) cuda_check(cudaStreamWaitEvent(s, ctx.ev_pos_ready, 0), "cudaStreamWaitEvent");

    if (wrapped && p.port != Portabilization::Quantize) {  // ToBits: the values are what they are, only their bounds are needed
      prof.begin("wrap_minmax", 4ull * p.ncomp_q * U + U, s);
      gpu::launch_wrap_minmax(d.quant, U, p.ncomp_q, p.ncomp_q == 3 ? 4 : p.ncomp_q, d.used, d.stats, s);
      prof.end(s);
    }
    switch (p.scheme) {
      case Scheme::Parallelogram:
        prof.begin("K4_predict_parallelogram", 4ull * M + 4 * C + 4 * C + 4 * V + 4 * V + 4ull * p.ncomp_q * U + 4 * S, s);
        if (d.para_records) gpu::launch_predict_parallelogram_records(d.para_records, M, q, d.symbols, d.stats, s);
        else gpu::launch_predict_parallelogram(d.seq, M, t, q, d.rank, d.symbols, d.stats, s);
        prof.end(s);
        break;
      case Scheme::Normal: {
        const AttrDevice& pd = dev_[p.parent];
        const gpu::QuantDev pos{pd.quant, pd.map, 3};
        prof.begin("K5_predict_normal", 4ull * M + 4 * C + C + 4 * C + 4 * C + 12 * Upos + 8 * U + 4 * S + M, s);
        gpu::launch_predict_normal(d.seq, M, t, q, pos, d.symbols, d.side_out + 8, d.stats, s);  // flips go straight behind the scalars
        prof.end(s);
        break;
      }
      case Scheme::TexCoord: {
        const AttrDevice& pd = dev_[p.parent];
        const gpu::QuantDev pos{pd.quant, pd.map, 3};
        prof.begin("K6_predict_texcoord", 4ull * M + 4 * C + 4 * C + 4 * V + 4 * V + 12 * Upos + 8 * U + 4 * S + M, s);
        if (d.tex_records) gpu::launch_predict_texcoord_records(d.tex_records, M, q, pos, d.symbols, d.side, d.stats, s);
        else gpu::launch_predict_texcoord(d.seq, M, t, q, pos, plans_[p.parent].view.num_points, d.rank, d.symbols, d.side, d.stats, s);
        prof.end(s);
        break;
      }
      case Scheme::Delta:
        prof.begin("K7_predict_delta", 4ull * M + 4 * V + 4ull * p.ncomp_q * U + 4 * S, s);
        gpu::launch_predict_delta(d.seq, M, t, q, d.symbols, d.stats, s);
        prof.end(s);
        break;
    }
    prof.begin("K8_histogram", 4 * S, s);
    gpu::launch_histogram(d.symbols, S, d.hist, p.hist_capacity, d.stats, s);
    prof.end(s);
    prof.begin("K9_build_table", 4ull * p.hist_capacity, s);
    gpu::launch_build_table(d.hist, p.hist_capacity, S, d.work, d.rans_table, d.table_bytes, d.table_capacity, d.stats, s);
    prof.end(s);
    if (side_ready_[i]) {
      // The side stream input leaves for the host as soon as it is ready; the host codes it during K9-K10. What is
      // data-parallel about it is done here: flips are counted, orientation flags compacted and their transitions counted.
      uint8_t* host = side_host_[i];
      uint32_t* scalars = (uint32_t*)d.side_out;
      if (p.scheme == Scheme::Normal) {
        gpu::launch_count_flips(d.side_out + 8, M, scalars, s);
        prof.launches += 1;
      } else {
        gpu::launch_compact_orientations(d.side, M, d.side_out + 8, scalars, d.side_scratch, d.side_scratch_bytes, s);
        prof.launches += 2;  // select + transition count (cub's own init kernel not counted)
      }
      cuda_check(cudaEventRecord(side_ready_[i], s), "cudaEventRecord");
      cuda_check(cudaStreamWaitEvent(ctx.copy_stream, side_ready_[i], 0), "cudaStreamWaitEvent");
      cuda_check(cudaMemcpyAsync(host, d.side_out, M + 8, cudaMemcpyDeviceToHost, ctx.copy_stream), "cudaMemcpyAsync D2H");
      // inside a graph the event the host coder waits for has to be an external event-record node
      cuda_check(cudaEventRecordWithFlags(side_copied_[i], ctx.copy_stream, capturing_ ? cudaEventRecordExternal : cudaEventRecordDefault), "cudaEventRecord");
      d2h_bytes += M + 8;
    }
    prof.begin("K10_rans_encode", 4 * S, s);
    gpu::launch_rans_encode(d.symbols, S, d.rans_table, p.hist_capacity, d.rans_scratch, d.payload, d.stats, s);
    prof.launches += gpu::rans_launch_count(S) - 1;  // explore + chain + encode + fix-up + gather
    prof.end(s);
    if (prof.serial) cuda_check(cudaEventRecord(ctx.ev_serial, s), "cudaEventRecord");
  }
  cuda_check(cudaGetLastError(), "kernel launch");
}

// Binary side stream of attribute `att` (runs on a host worker thread):
// normals — flips in sequence order, zero_prob from the zero count (mesh_normal_prediction.rs:147-163);
// texcoords — order-preserving compaction of the orientation flags, zero_prob from the forward
// transitions over len + 0.001, backward-delta bits written forward (mesh_prediction_for_texture_coordinates.rs:221-260).
// bits[k] = (o[k] == o[k + 1]) with o[n] = true (2), coded in order (mesh_prediction_for_texture_coordinates.rs:247-260). A stream skewed
// enough for the table-driven coder gets its delta bits materialised first (a vectorisable pass) so that the rare ones can
// be found eight at a time; otherwise they are formed inside the coder's loop.
static void encode_orientation_deltas(const uint8_t* o, size_t n, uint8_t zero_prob, std::vector<uint8_t>& out) {
  if (n && rabs_sparse_applies(n, zero_prob, 0)) {
    U8Array d(n);
    uint8_t* const dp = d.data();
    for (size_t k = 0; k + 1 < n; ++k) dp[k] = o[k] == o[k + 1];
    dp[n - 1] = o[n - 1] == 2;
    rabs_encode_forward(dp, n, zero_prob, out);
    return;
  }
  rabs_encode_forward_fn(n, zero_prob, out, [o, n](size_t k) { return (uint8_t)(o[k] == (k + 1 < n ? o[k + 1] : (uint8_t)2)); });
}

void MeshJob::encode_side_stream(size_t att) {
  AttrResult& r = results_[att];
  uint32_t scalars[2];
  memcpy(scalars, r.side, 8);
  const uint8_t* flags = r.side + 8;
  const size_t n = scalars[0];
  if (n > r.side_len) throw Error(DXO_ERR_INTERNAL, "side stream longer than its attribute");
  r.side_count = (uint32_t)n;
  if (plans_[att].scheme == Scheme::Normal) {
    r.side_zero_prob = side_stream_zero_prob(n - scalars[1], (float)n);  // scalars[1] = flips set
    if (scalars[1] == 0) rabs_encode_zero_run(n, r.side_zero_prob, r.side_payload);  // no flip at all (smooth normals): cycle jumps
    else rabs_encode_forward(flags, n, r.side_zero_prob, r.side_payload);
  } else {
    // flags[0..n) = orientation values (1 = false, 2 = true) in order, scalars[1] = forward transitions; the delta bits
    // bits[k] = (o[k] == o[k+1]), o[len] = true, are formed inside the coder's loop
    r.side_zero_prob = side_stream_zero_prob(scalars[1], (float)n + 0.001f);
    encode_orientation_deltas(flags, n, r.side_zero_prob, r.side_payload);
  }
}

// Two side streams (a normal attribute's flips and a texcoord attribute's orientations, or two of a kind) in one loop.
void MeshJob::encode_side_stream_pair(size_t ia, size_t ib) {
  struct View { const uint8_t* flags; size_t n; uint8_t p0; bool normal; };
  auto view = [&](size_t att) {
    AttrResult& r = results_[att];
    uint32_t scalars[2];
    memcpy(scalars, r.side, 8);
    const size_t n = scalars[0];
    if (n > r.side_len) throw Error(DXO_ERR_INTERNAL, "side stream longer than its attribute");
    r.side_count = (uint32_t)n;
    const bool normal = plans_[att].scheme == Scheme::Normal;
    r.side_zero_prob = normal ? side_stream_zero_prob(n - scalars[1], (float)n) : side_stream_zero_prob(scalars[1], (float)n + 0.001f);
    return View{r.side + 8, n, r.side_zero_prob, normal};
  };
  const View a = view(ia), b = view(ib);
  {  // a flip stream without a single flip is coded by cycle jumps; the other stream then has the loop to itself
    uint32_t sa[2], sb[2];
    memcpy(sa, results_[ia].side, 8);
    memcpy(sb, results_[ib].side, 8);
    const bool za = a.normal && sa[1] == 0, zb = b.normal && sb[1] == 0;
    // ... and a stream skewed enough for the table-driven coder (rabs.cpp) gains nothing from sharing a loop either
    const bool sparse = rabs_sparse_applies(a.n, a.p0, 0) || rabs_sparse_applies(b.n, b.p0, 0);
    if (za || zb || sparse) {
      if (za) rabs_encode_zero_run(a.n, a.p0, results_[ia].side_payload); else encode_side_stream(ia);
      if (zb) rabs_encode_zero_run(b.n, b.p0, results_[ib].side_payload); else encode_side_stream(ib);
      return;
    }
  }
  auto bit = [](const View& v) {
    const uint8_t* f = v.flags;
    const size_t n = v.n;
    const bool normal = v.normal;
    return [f, n, normal](size_t k) -> uint8_t { return normal ? f[k] : (uint8_t)(f[k] == (k + 1 < n ? f[k + 1] : (uint8_t)2)); };
  };
  rabs_encode_forward_pair(a.n, a.p0, results_[ia].side_payload, bit(a), b.n, b.p0, results_[ib].side_payload, bit(b));
}

void MeshJob::encode_side_stream_from_flags(size_t att, const uint8_t* flags, size_t n) {
  AttrResult& r = results_[att];
  if (plans_[att].scheme == Scheme::Normal) {
    size_t ones = 0;
    for (size_t k = 0; k < n; ++k) ones += flags[k] != 0;
    r.side_count = (uint32_t)n;
    r.side_zero_prob = side_stream_zero_prob(n - ones, (float)n);
    if (ones == 0) rabs_encode_zero_run(n, r.side_zero_prob, r.side_payload);
    else rabs_encode_forward(flags, n, r.side_zero_prob, r.side_payload);
    return;
  }
  U8Array o(n);  // the orientation values that exist, in order
  size_t m = 0, transitions = 0;
  uint8_t last = 2;  // the transition scan starts from `true`
  for (size_t k = 0; k < n; ++k) {
    const uint8_t f = flags[k];
    if (!f) continue;
    o[m++] = f;
    transitions += f != last;
    last = f;
  }
  r.side_count = (uint32_t)m;
  r.side_zero_prob = side_stream_zero_prob(transitions, (float)m + 0.001f);
  encode_orientation_deltas(o.data(), m, r.side_zero_prob, r.side_payload);
}

constexpr size_t kStatsSlot = 1024;  // pinned_buffer slot of the per-attribute scalars (attribute slots are 2i, 2i+1)

void MeshJob::download(DeviceContext& ctx) {
  results_.resize(plans_.size());
  // side streams: start a host worker per stream as soon as its flags have landed — or, with DXO_SIDE_INLINE=1 (hosts with
  // few threads per GPU), code them on this thread, two at a time in one interleaved loop, while the device runs K8-K10
  static const bool side_inline_env = getenv("DXO_SIDE_INLINE") != nullptr && getenv("DXO_SIDE_INLINE")[0] != '0';
  const bool side_inline = side_inline_env || inline_host;  // many encodes in flight: no helper threads for this either
  std::vector<std::future<void>> workers;
  std::vector<size_t> inline_streams;
  for (size_t i = 0; i < plans_.size(); ++i) {
    if (!side_copied_[i]) continue;
    results_[i].side = side_host_[i];
    results_[i].side_len = sequence_of(i).size();
    if (side_inline) { inline_streams.push_back(i); continue; }
    const int device = ctx.device;
    cudaEvent_t ev = side_copied_[i];
    workers.push_back(ctx.helpers->run([this, i, device, ev] {
      cuda_check(cudaSetDevice(device), "cudaSetDevice");
      const auto t0 = std::chrono::steady_clock::now();
      cuda_check(cudaEventSynchronize(ev), "cudaEventSynchronize");
      const auto t1 = std::chrono::steady_clock::now();
      encode_side_stream(i);
      if (getenv("DXO_TIMING"))
        fprintf(stderr, "[dxo] side stream %zu: waited %.3f ms for flags, coded in %.3f ms\n", i, std::chrono::duration<double, std::milli>(t1 - t0).count(),
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count());
    }));
  }
  auto join_workers = [&] { for (auto& w : workers) w.get(); };
  try {
    for (size_t k = 0; k < inline_streams.size(); k += 2) {
      cuda_check(cudaEventSynchronize(side_copied_[inline_streams[k]]), "cudaEventSynchronize");
      if (k + 1 < inline_streams.size()) {
        cuda_check(cudaEventSynchronize(side_copied_[inline_streams[k + 1]]), "cudaEventSynchronize");
        encode_side_stream_pair(inline_streams[k], inline_streams[k + 1]);
      } else {
        encode_side_stream(inline_streams[k]);
      }
    }
    // the scalars land in pinned memory: a device-to-host copy into pageable memory blocks the calling thread inside the
    // driver until the stream gets there, which stalls the launches of every other thread of the process
    gpu::AttrStats* stats_host = (gpu::AttrStats*)ctx.pinned_buffer(kStatsSlot, plans_.size() * sizeof(gpu::AttrStats));
    for (size_t i = 0; i < plans_.size(); ++i) {
      cudaStream_t s = ctx.stream[std::min<size_t>(i, 2)];
      cuda_check(cudaMemcpyAsync(stats_host + i, dev_[i].stats, sizeof(gpu::AttrStats), cudaMemcpyDeviceToHost, s), "cudaMemcpyAsync D2H");
      d2h_bytes += sizeof(gpu::AttrStats);
    }
    for (size_t i = 0; i < plans_.size(); ++i) {
      cudaStream_t s = ctx.stream[std::min<size_t>(i, 2)];
      if (i < 3) ctx.wait_stream((int)i);  // attributes 3.. share stream 2, already drained
      AttrResult& r = results_[i];
      r.stats = stats_host[i];
      if (int st = status_from_flags(r.stats.error_flags)) throw Error(st, "device reported an encoding error");
      static const bool rans_debug = getenv("DXO_RANS_DEBUG") != nullptr;
      if (rans_debug) fprintf(stderr, "[dxo] att %zu: rANS symbols=%llu P=%u K=%u chunks=%u chain misses=%u fixup=%u\n", i,
                              (unsigned long long)sequence_of(i).size() * plans_[i].ncomp_q, r.stats.precision, r.stats.num_table_symbols,
                              gpu::rans_num_chunks((uint64_t)sequence_of(i).size() * plans_[i].ncomp_q), r.stats.pad[0], r.stats.pad[1]);
      if (r.stats.table_bytes > dev_[i].table_capacity || r.stats.payload_bytes > dev_[i].payload_capacity)
        throw Error(DXO_ERR_INTERNAL, "device output exceeds its buffer");
      uint8_t* host = ctx.pinned_buffer(2 * i, (size_t)r.stats.table_bytes + r.stats.payload_bytes + 16);
      r.table_bytes = host;
      r.payload = host + r.stats.table_bytes;
      cuda_check(cudaMemcpyAsync(host, dev_[i].table_bytes, r.stats.table_bytes, cudaMemcpyDeviceToHost, s), "cudaMemcpyAsync D2H");
      cuda_check(cudaMemcpyAsync(host + r.stats.table_bytes, dev_[i].payload, r.stats.payload_bytes, cudaMemcpyDeviceToHost, s), "cudaMemcpyAsync D2H");
      d2h_bytes += (size_t)r.stats.table_bytes + r.stats.payload_bytes;
    }
    for (int k = 0; k < 3; ++k) ctx.wait_stream(k);
  } catch (...) {
    try { join_workers(); } catch (...) {}
    throw;
  }
  join_workers();
  if (trace) capture_trace(ctx);
}

// ---------------------------------------------------------------------------------------
void MeshJob::assemble(std::vector<uint8_t>& out) {
  ByteSink w;
  w.data.swap(out);  // a caller that keeps its vector across steps (resident sessions) keeps the storage too
  w.data.clear();
  size_t total = head_.size() + 64;
  for (const AttrResult& r : results_) total += (size_t)r.stats.table_bytes + r.stats.payload_bytes + r.side_payload.size() + r.side_bytes_len + 64;
  w.data.reserve(total);
  w.bytes(head_.data);
  for (size_t i = 0; i < plans_.size(); ++i) {
    const AttrPlan& p = plans_[i];
    const AttrResult& r = results_[i];
    w.u8((uint8_t)p.scheme);     // attribute_encoder.rs:159
    w.u8((uint8_t)p.transform);  // :160
    w.u8(1);                     // rans_encoding (:344)
    w.u8(1);                     // SymbolEncodingMethod::DirectCoded (symbol_coding.rs:22)
    w.u8((uint8_t)r.stats.bit_length);
    w.bytes(r.table_bytes, r.stats.table_bytes);  // leb128 #symbols + frequency table (rans.rs:193-230)
    w.varint(r.stats.payload_bytes);              // RansSymbolEncoder::flush (rans.rs:248-255)
    w.bytes(r.payload, r.stats.payload_bytes);
    // metadata: order depends on the scheme (:362-386)
    auto transform_info = [&] {
      if (p.transform == Transform::Wrapped) { w.i32(r.stats.wrap_min); w.i32(r.stats.wrap_max); }   // wrapped_difference.rs:95-96
      else if (p.transform == Transform::OctOrthogonal) { w.u32(255); w.u32(127); }                  // oct_orthogonal.rs:80-82
    };
    auto side_stream = [&] {  // zero_prob byte, leb128 size, rABS bytes (coded by encode_side_stream, or on the device in the group path)
      w.u8(r.side_zero_prob);
      if (r.side_bytes) { w.varint(r.side_bytes_len); w.bytes(r.side_bytes, r.side_bytes_len); }
      else { w.varint(r.side_payload.size()); w.bytes(r.side_payload); }
    };
    if (p.scheme == Scheme::Normal) {
      transform_info();
      side_stream();
    } else if (p.scheme == Scheme::TexCoord) {
      w.u32(r.side_count);
      side_stream();
      transform_info();
    } else {
      transform_info();
    }
    // portabilization metadata (:384-386)
    if (p.port == Portabilization::Quantize) {
      for (uint32_t k = 0; k < p.ncomp_in; ++k) w.u32(r.stats.vmin_bits[k]);  // min values (quantization_coordinate_wise.rs:57)
      w.f32(r.stats.range);
      w.u8((uint8_t)p.bits);
    } else if (p.port == Portabilization::Octahedral) {
      w.u8(8);  // octahedral_quantization.rs:40
    }
  }
  out = std::move(w.data);
}

void MeshJob::release(DeviceContext& ctx) {
  for (void* p : allocations_) cudaFreeAsync(p, ctx.stream[0]);
  allocations_.clear();
  uploaded_ = false;
}

// Copies intermediate device results back under the same keys / layouts the oracle's
// trace uses, for stage-by-stage parity tests.
void MeshJob::capture_host_trace() {
  auto put = [&](const std::string& key, const void* p, size_t bytes) { auto& v = trace_items[key]; v.assign((const uint8_t*)p, (const uint8_t*)p + bytes); };
  put("head_bytes", head_.data.data(), head_.data.size());
  put("opposite", ut_.opposite.data(), ut_.opposite.size() * 4);
  put("corner_to_vertex", ut_.corner_vertex.data(), ut_.corner_vertex.size() * 4);
  put("left_most", ut_.left_most.data(), ut_.left_most.size() * 4);
  { uint64_t nv = ut_.num_vertices; put("num_vertices", &nv, 8); }
  { const std::vector<uint32_t> all = eb_->corners_of_edgebreaker(); put("corners_of_edgebreaker", all.data(), all.size() * 4); }
  for (size_t i = 0; i < plans_.size(); ++i) {
    const std::string k = "att" + std::to_string(i) + ".";
    if (i > 0) {
      put(k + "c2v", seams_[i - 1].corner_vertex.data(), seams_[i - 1].corner_vertex.size() * 4);
      put(k + "left_most", seams_[i - 1].left_most.data(), seams_[i - 1].left_most.size() * 4);
      put(k + "seam", seams_[i - 1].seam.data(), seams_[i - 1].seam.size());
    }
    put(k + "sequence", sequence_of(i).data(), sequence_of(i).size() * 4);
  }
}

void MeshJob::capture_trace(DeviceContext& ctx) {
  capture_host_trace();

  auto put = [&](const std::string& key, const void* p, size_t bytes) { auto& v = trace_items[key]; v.assign((const uint8_t*)p, (const uint8_t*)p + bytes); };
  auto fetch = [&](const std::string& key, const void* dptr, size_t bytes) {
    auto& v = trace_items[key];
    v.resize(bytes);
    if (bytes) cuda_check(cudaMemcpy(v.data(), dptr, bytes, cudaMemcpyDeviceToHost), "cudaMemcpy trace");
  };
  (void)ctx;
  for (size_t i = 1; i < plans_.size(); ++i) {  // seam tables that stayed on the device (device_seam_table: same as the universal table)
    const SeamTable& st = seams_[i - 1];
    if (!st.corner_vertex.empty() || !dev_[i].corner_vertex) continue;
    const std::string k = "att" + std::to_string(i) + ".";
    fetch(k + "c2v", dev_[i].corner_vertex, (size_t)ut_.num_corners * 4);
    fetch(k + "left_most", dev_[i].left_most, (size_t)st.num_vertices * 4);
    fetch(k + "seam", dev_[i].seam, (size_t)ut_.num_corners);
  }
  for (size_t i = 0; i < plans_.size(); ++i) {
    const AttrPlan& p = plans_[i];
    const AttrResult& r = results_[i];
    const std::string k = "att" + std::to_string(i) + ".";
    if (p.ncomp_q == 3) {  // stored with stride 4 on the device
      std::vector<int32_t> padded((size_t)p.view.num_unique * 4), packed((size_t)p.view.num_unique * 3);
      if (!padded.empty()) cuda_check(cudaMemcpy(padded.data(), dev_[i].quant, padded.size() * 4, cudaMemcpyDeviceToHost), "cudaMemcpy trace");
      for (size_t v = 0; v < p.view.num_unique; ++v) for (int c = 0; c < 3; ++c) packed[v * 3 + c] = padded[v * 4 + c];
      put(k + "quantized", packed.data(), packed.size() * 4);
    } else {
      fetch(k + "quantized", dev_[i].quant, (size_t)p.view.num_unique * p.ncomp_q * 4);
    }
    fetch(k + "symbols", dev_[i].symbols, sequence_of(i).size() * p.ncomp_q * 4);
    {
      std::vector<uint32_t> h(r.stats.num_table_symbols);
      if (!h.empty()) cuda_check(cudaMemcpy(h.data(), dev_[i].hist, h.size() * 4, cudaMemcpyDeviceToHost), "cudaMemcpy trace");
      std::vector<uint64_t> h64(h.begin(), h.end());
      put(k + "histogram", h64.data(), h64.size() * 8);
      std::vector<uint32_t> d(r.stats.num_table_symbols);
      if (!d.empty()) cuda_check(cudaMemcpy(d.data(), dev_[i].work, d.size() * 4, cudaMemcpyDeviceToHost), "cudaMemcpy trace");
      std::vector<uint64_t> d64(d.begin(), d.end());
      put(k + "distribution", d64.data(), d64.size() * 8);
    }
    put(k + "table_bytes", r.table_bytes, r.stats.table_bytes);
    put(k + "payload", r.payload, r.stats.payload_bytes);
    { int32_t mm[2] = {r.stats.wrap_min, r.stats.wrap_max}; put(k + "wrap_minmax", mm, 8); }
    { uint32_t bl[2] = {r.stats.bit_length, r.stats.precision}; put(k + "bit_length", bl, 8); }
    std::vector<uint8_t> side;
    if (p.scheme == Scheme::Normal) side.assign(r.side + 8, r.side + 8 + r.side_count);
    else if (p.scheme == Scheme::TexCoord) for (size_t e = 0; e < r.side_count; ++e) side.push_back(r.side[8 + e] == 2 ? 1 : 0);
    put(k + "side_bits", side.data(), side.size());
  }
}

}  // namespace dxo
