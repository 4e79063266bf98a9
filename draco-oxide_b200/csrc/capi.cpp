// extern "C" surface declared in include/dxo.h.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <thread>

#include "encoder.hpp"

using namespace dxo;

namespace {

using Clock = std::chrono::steady_clock;
double ms_since(Clock::time_point t0) { return std::chrono::duration<double, std::milli>(Clock::now() - t0).count(); }

thread_local dxo_timing g_timing{};
thread_local std::string g_error;
std::atomic<int> g_profiling{0};

template <class F> int guarded(F&& f) {
  try { f(); return DXO_OK; }
  catch (const Error& e) { g_error = e.what(); return e.status; }
  catch (const std::bad_alloc&) { g_error = "out of host memory"; return DXO_ERR_OUT_OF_MEMORY; }
  catch (const std::exception& e) { g_error = e.what(); return DXO_ERR_INTERNAL; }
}

int give(std::vector<uint8_t>& bytes, dxo_bytes* out) {
  out->data = (uint8_t*)malloc(bytes.size() ? bytes.size() : 1);
  if (!out->data) return DXO_ERR_OUT_OF_MEMORY;
  memcpy(out->data, bytes.data(), bytes.size());
  out->len = bytes.size();
  return DXO_OK;
}

dxo_config effective_config(const dxo_config* cfg) {
  dxo_config c;
  dxo_config_default(&c);
  if (cfg) {
    if (cfg->abi_version != DXO_ABI_VERSION) throw Error(DXO_ERR_INVALID_ARGUMENT, "dxo_config.abi_version mismatch");
    c = *cfg;
  }
  return c;
}

// Runs launch + download + assemble for a job whose inputs are already on the device,
// filling the thread's timing record.
void run_device_phase(MeshJob& job, DeviceContext& ctx, Profile& prof, std::vector<uint8_t>& bytes, dxo_timing& tm) {
  prof.reset();
  prof.enabled = g_profiling.load() != 0;
  prof.serial = g_profiling.load() == 2;
  cudaStream_t s0 = ctx.stream[0];
  // opt-in (DXO_FLAG_GRAPH_REPLAY); one-shot encodes, traced and profiled runs launch directly
  const bool replay = job.graph_replay && !prof.enabled && !job.trace && job.device_runs >= 1;
  ++job.device_runs;
  const auto t_launch = Clock::now();
  cuda_check(cudaEventRecord(ctx.ev_begin, s0), "cudaEventRecord");
  if (replay) job.launch_graph(ctx, prof);
  else job.launch_all(ctx, prof);
  cuda_check(cudaEventRecord(ctx.ev_end, s0), "cudaEventRecord");
  const double launch_cpu_ms = ms_since(t_launch);
  if (prof.enabled) cuda_check(cudaEventSynchronize(ctx.ev_end), "cudaEventSynchronize");
  const auto t_d2h = Clock::now();
  job.download(ctx);
  tm.d2h_ms = (float)ms_since(t_d2h);
  cuda_check(cudaEventSynchronize(ctx.ev_end), "cudaEventSynchronize");
  cuda_check(cudaEventElapsedTime(&tm.device_ms, ctx.ev_begin, ctx.ev_end), "cudaEventElapsedTime");
  const auto t_asm = Clock::now();
  job.assemble(bytes);
  if (getenv("DXO_TIMING"))
    fprintf(stderr, "[dxo] thread %zx: launch (cpu) %.3f ms | device %.3f ms | download+side streams (wall) %.3f ms | assemble %.3f ms\n",
            (size_t)std::hash<std::thread::id>()(std::this_thread::get_id()) & 0xFFFF, launch_cpu_ms, tm.device_ms, tm.d2h_ms, ms_since(t_asm));
  tm.num_launches = prof.launches;
  tm.num_kernels = 0;
  for (const KernelRecord& r : prof.records) {
    if (tm.num_kernels >= 64) break;
    dxo_kernel_time& k = tm.kernels[tm.num_kernels++];
    k.name = r.name;
    k.algorithmic_bytes = r.bytes;
    cuda_check(cudaEventElapsedTime(&k.ms, r.a, r.b), "cudaEventElapsedTime");
  }
  tm.d2h_bytes = job.d2h_bytes;
}

thread_local Profile g_profile;

void encode_one(const dxo_mesh* mesh, const dxo_config& cfg, std::vector<uint8_t>& bytes, dxo_timing& tm, bool parallel_host = true) {
  dxo::encode_one_mesh(mesh, cfg, bytes, tm, parallel_host);
}

}  // namespace

void dxo::encode_one_mesh(const dxo_mesh* mesh, const dxo_config& cfg, std::vector<uint8_t>& bytes, dxo_timing& tm, bool parallel_host) {
  tm = dxo_timing{};
  const auto t0 = Clock::now();
  MeshJob job(mesh, cfg);
  job.parallel_host = parallel_host;
  // encodes in flight in this process: from a few on, every call keeps its host passes on its own thread
  static std::atomic<int> in_flight{0};
  struct InFlight { std::atomic<int>& n; int before; explicit InFlight(std::atomic<int>& a) : n(a), before(a.fetch_add(1)) {} ~InFlight() { n.fetch_sub(1); } } guard(in_flight);
  static const int crowd = getenv("DXO_INLINE_HOST_FROM") ? atoi(getenv("DXO_INLINE_HOST_FROM")) : std::max(2, (int)std::thread::hardware_concurrency() / 4);
  job.inline_host = parallel_host && guard.before + 1 >= crowd;
  DeviceContext& ctx = DeviceContext::get(cfg.device);
  job.build_connectivity(&ctx);
  tm.host_connectivity_ms = (float)ms_since(t0);
  const auto t1 = Clock::now();
  job.upload(ctx);
  if (g_profiling.load()) cuda_check(cudaStreamSynchronize(ctx.stream[0]), "cudaStreamSynchronize");
  tm.h2d_ms = (float)ms_since(t1);
  tm.h2d_bytes = job.h2d_bytes;
  try { run_device_phase(job, ctx, g_profile, bytes, tm); }
  catch (...) { job.release(ctx); cudaStreamSynchronize(ctx.stream[0]); throw; }
  job.release(ctx);
  tm.total_ms = (float)ms_since(t0);
}

struct dxo_session {
  std::unique_ptr<MeshJob> job;
  dxo_config cfg;
  int device = 0;
};

extern "C" {

void dxo_config_default(dxo_config* cfg) {
  if (!cfg) return;
  cfg->abi_version = DXO_ABI_VERSION;
  cfg->position_bits = 11;   // portabilization/mod.rs:107-112
  cfg->texcoord_bits = 10;   // :127-130
  cfg->generic_bits = 11;
  cfg->device = -1;
  cfg->flags = 0;
}

int dxo_encode(const dxo_mesh* mesh, const dxo_config* cfg, dxo_bytes* out) {
  if (!out) return DXO_ERR_INVALID_ARGUMENT;
  out->data = nullptr; out->len = 0;
  return guarded([&] {
    const dxo_config c = effective_config(cfg);
    std::vector<uint8_t> bytes;
    encode_one(mesh, c, bytes, g_timing);
    if (int st = give(bytes, out)) throw Error(st, "out of memory");
  });
}

int dxo_encode_batch(const dxo_mesh* meshes, size_t n, const dxo_config* cfg, dxo_bytes* outs, int* statuses, int first_gpu, int num_gpus) {
  if ((!meshes || !outs) && n) return DXO_ERR_INVALID_ARGUMENT;
  for (size_t i = 0; i < n; ++i) { outs[i].data = nullptr; outs[i].len = 0; if (statuses) statuses[i] = DXO_OK; }
  int first_error = DXO_OK;
  int st = guarded([&] {
    const dxo_config base = effective_config(cfg);
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) { cudaGetLastError(); throw Error(DXO_ERR_NO_DEVICE, "no CUDA device available (this path has no CPU fallback)"); }
    if (num_gpus <= 0) num_gpus = 1;
    if (first_gpu < 0 || first_gpu + num_gpus > count) throw Error(DXO_ERR_NO_DEVICE, "GPU range out of bounds");
    if (!getenv("DXO_BATCH_PER_MESH")) {  // groups of meshes through segmented launches (batch.cpp)
      std::vector<int> sts(n, DXO_OK);
      encode_batch_grouped(meshes, n, base, outs, sts.data(), first_gpu, num_gpus);
      for (size_t i = 0; i < n; ++i) { if (statuses) statuses[i] = sts[i]; if (sts[i] != DXO_OK && first_error == DXO_OK) first_error = sts[i]; }
      return;
    }
    // DXO_BATCH_PER_MESH=1: the per-mesh path for every mesh, longest-processing-time-first order over a shared queue
    std::vector<size_t> order(n);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return meshes[a].num_faces > meshes[b].num_faces; });
    std::atomic<size_t> next{0};
    std::vector<int> sts(n, DXO_OK);
    const char* env = getenv("DXO_WORKERS_PER_GPU");
    // Inside a worker the host passes of a mesh run on one thread, so the batch is host-bound: one worker per host
    // thread (at least 3, at most 16 per GPU); their device work overlaps on the GPU.
    const int hw = (int)std::thread::hardware_concurrency();
    int per_gpu = env ? atoi(env) : std::min(16, std::max(3, hw / std::max(1, num_gpus)));
    if (per_gpu < 1) per_gpu = 1;
    const int workers = (int)std::min<size_t>((size_t)num_gpus * per_gpu, std::max<size_t>(n, 1));
    std::vector<std::thread> pool;
    for (int w = 0; w < workers; ++w) {
      pool.emplace_back([&, w] {
        dxo_config c = base;
        c.device = first_gpu + w % num_gpus;
        dxo_timing tm;
        for (;;) {
          const size_t k = next.fetch_add(1);
          if (k >= n) break;
          const size_t i = order[k];
          sts[i] = guarded([&] {
            std::vector<uint8_t> bytes;
            encode_one(&meshes[i], c, bytes, tm, workers == 1);
            if (int s2 = give(bytes, &outs[i])) throw Error(s2, "out of memory");
          });
        }
      });
    }
    for (auto& t : pool) t.join();
    for (size_t i = 0; i < n; ++i) { if (statuses) statuses[i] = sts[i]; if (sts[i] != DXO_OK && first_error == DXO_OK) first_error = sts[i]; }
  });
  return st != DXO_OK ? st : first_error;
}

int dxo_encode_glb(const dxo_mesh* meshes, size_t n, const dxo_config* cfg, const dxo_bytes* streams, dxo_bytes* glb_out, int first_gpu, int num_gpus) {
  if (!glb_out || (!meshes && n)) return DXO_ERR_INVALID_ARGUMENT;
  glb_out->data = nullptr; glb_out->len = 0;
  std::vector<dxo_bytes> own;
  int st = DXO_OK;
  if (!streams) {  // encode here: one batch call
    own.assign(n, dxo_bytes{nullptr, 0});
    st = dxo_encode_batch(meshes, n, cfg, own.data(), nullptr, first_gpu, num_gpus);
    streams = own.data();
  }
  if (st == DXO_OK) st = guarded([&] {
    std::vector<uint8_t> glb;
    assemble_glb(meshes, streams, n, glb);
    if (int s2 = give(glb, glb_out)) throw Error(s2, "out of memory");
  });
  for (dxo_bytes& b : own) dxo_free_bytes(&b);
  return st;
}

void dxo_free_bytes(dxo_bytes* b) {
  if (b && b->data) { free(b->data); b->data = nullptr; b->len = 0; }
}

const char* dxo_strerror(int status) {
  switch (status) {
    case DXO_OK: return "ok";
    case DXO_ERR_INVALID_ARGUMENT: return "invalid argument";
    case DXO_ERR_UNSUPPORTED_INPUT: return "unsupported input (the reference panics or is unimplemented here)";
    case DXO_ERR_UNSUPPORTED_DATA_TYPE: return "unsupported data type";
    case DXO_ERR_UNSUPPORTED_NUM_COMPONENTS: return "attribute data has too many components";
    case DXO_ERR_TOO_MANY_ATTRIBUTES: return "too many connectivity attributes";
    case DXO_ERR_RANS_INVALID_SYMBOL: return "rANS: invalid symbol index";
    case DXO_ERR_RANS_STATE_TOO_LARGE: return "rANS: state too large";
    case DXO_ERR_RANS_FREQ_TABLE: return "rANS: frequency table not compatible with the precision";
    case DXO_ERR_ZERO_NORMAL: return "zero vector cannot be transformed to octahedron space";
    case DXO_ERR_UNUSED_VERTICES: return "mesh contains unused vertices";
    case DXO_ERR_NO_DEVICE: return "no usable CUDA device (no CPU fallback on this path)";
    case DXO_ERR_CUDA: return "CUDA error";
    case DXO_ERR_OUT_OF_MEMORY: return "out of memory";
    default: return "internal error";
  }
}

int dxo_device_count(void) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess) { cudaGetLastError(); return 0; }
  return count;
}

int dxo_session_create(const dxo_mesh* mesh, const dxo_config* cfg, dxo_session** out) {
  if (!out) return DXO_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  return guarded([&] {
    auto s = std::make_unique<dxo_session>();
    s->cfg = effective_config(cfg);
    g_timing = dxo_timing{};
    const auto t0 = Clock::now();
    s->job = std::make_unique<MeshJob>(mesh, s->cfg);
    s->job->resident = true;
    DeviceContext& ctx = DeviceContext::get(s->cfg.device);
    s->job->build_connectivity(&ctx);
    g_timing.host_connectivity_ms = (float)ms_since(t0);
    s->device = ctx.device;
    const auto t1 = Clock::now();
    s->job->upload(ctx);
    cuda_check(cudaStreamSynchronize(ctx.stream[0]), "cudaStreamSynchronize");
    g_timing.h2d_ms = (float)ms_since(t1);
    g_timing.h2d_bytes = s->job->h2d_bytes;
    *out = s.release();
  });
}

int dxo_connectivity_create(const dxo_mesh* mesh, const dxo_config* cfg, dxo_session** out) {
  if (!out) return DXO_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  return guarded([&] {
    auto s = std::make_unique<dxo_session>();
    s->cfg = effective_config(cfg);
    s->job = std::make_unique<MeshJob>(mesh, s->cfg);
    s->job->build_connectivity();
    s->job->capture_host_trace();
    s->device = -1;
    *out = s.release();
  });
}

int dxo_session_run(dxo_session* s, dxo_bytes* out) {
  if (!s) return DXO_ERR_INVALID_ARGUMENT;
  if (s->device < 0) return DXO_ERR_NO_DEVICE;  // host-only session (dxo_connectivity_create)
  if (out) { out->data = nullptr; out->len = 0; }
  return guarded([&] {
    DeviceContext& ctx = DeviceContext::get(s->device);
    const float conn = g_timing.host_connectivity_ms, h2d = g_timing.h2d_ms;
    const uint64_t h2db = g_timing.h2d_bytes;
    const auto t0 = Clock::now();
    std::vector<uint8_t> bytes;
    s->job->d2h_bytes = 0;
    run_device_phase(*s->job, ctx, g_profile, bytes, g_timing);
    g_timing.host_connectivity_ms = conn; g_timing.h2d_ms = h2d; g_timing.h2d_bytes = h2db;
    g_timing.total_ms = (float)ms_since(t0);
    if (out) { if (int st = give(bytes, out)) throw Error(st, "out of memory"); }
  });
}

int dxo_session_run_steps(dxo_session* s, uint32_t steps, float* ms_total, uint64_t* launches_total) {
  if (!s || !ms_total) return DXO_ERR_INVALID_ARGUMENT;
  if (s->device < 0) return DXO_ERR_NO_DEVICE;
  *ms_total = 0;
  if (launches_total) *launches_total = 0;
  return guarded([&] {
    DeviceContext& ctx = DeviceContext::get(s->device);
    static thread_local cudaEvent_t ev_a = nullptr, ev_b = nullptr;
    if (!ev_a) { cuda_check(cudaEventCreate(&ev_a), "cudaEventCreate"); cuda_check(cudaEventCreate(&ev_b), "cudaEventCreate"); }
    const auto t_enter = Clock::now();
    for (int k = 0; k < 3; ++k) cuda_check(cudaStreamSynchronize(ctx.stream[k]), "cudaStreamSynchronize");
    const double enter_ms = ms_since(t_enter);
    cuda_check(cudaEventRecord(ev_a, ctx.stream[0]), "cudaEventRecord");
    std::vector<uint8_t> bytes;
    uint64_t launches = 0;
    for (uint32_t k = 0; k < steps; ++k) {
      s->job->d2h_bytes = 0;
      run_device_phase(*s->job, ctx, g_profile, bytes, g_timing);
      launches += g_timing.num_launches;
    }
    const auto t_exit = Clock::now();
    cuda_check(cudaEventRecord(ev_b, ctx.stream[0]), "cudaEventRecord");
    cuda_check(cudaEventSynchronize(ev_b), "cudaEventSynchronize");
    cuda_check(cudaEventElapsedTime(ms_total, ev_a, ev_b), "cudaEventElapsedTime");
    if (getenv("DXO_TIMING")) fprintf(stderr, "[dxo] run_steps(%u): entry sync %.3f ms, steps %.3f ms, exit sync %.3f ms, events %.3f ms\n", steps, enter_ms,
                                      ms_since(t_enter) - enter_ms - ms_since(t_exit), ms_since(t_exit), *ms_total);
    if (launches_total) *launches_total = launches;
  });
}

void dxo_session_destroy(dxo_session* s) {
  if (!s) return;
  if (s->device >= 0) guarded([&] {
    DeviceContext& ctx = DeviceContext::get(s->device);
    s->job->release(ctx);
    cudaStreamSynchronize(ctx.stream[0]);
  });
  delete s;
}

void dxo_set_profiling(int mode) { g_profiling.store(mode == 2 ? 2 : (mode ? 1 : 0)); }

int dxo_last_timing(dxo_timing* out) {
  if (!out) return DXO_ERR_INVALID_ARGUMENT;
  *out = g_timing;
  return DXO_OK;
}

void dxo_session_set_trace(dxo_session* s, int enabled) { if (s) s->job->trace = enabled != 0; }

int dxo_session_trace_get(dxo_session* s, const char* key, const void** data, uint64_t* nbytes) {
  if (!s || !key || !data || !nbytes) return DXO_ERR_INVALID_ARGUMENT;
  auto it = s->job->trace_items.find(key);
  if (it == s->job->trace_items.end()) return DXO_ERR_INVALID_ARGUMENT;
  *data = it->second.data();
  *nbytes = it->second.size();
  return DXO_OK;
}

int dxo_encode_bits(const uint8_t* bits, uint64_t n, uint8_t zero_prob, int mode, dxo_bytes* out) {
  if (!out || (!bits && n) || zero_prob == 0) return DXO_ERR_INVALID_ARGUMENT;
  out->data = nullptr; out->len = 0;
  return guarded([&] {
    std::vector<uint8_t> bytes;
    if (mode == 0) { ByteSink sink; rabs_encode(bits, (size_t)n, false, zero_prob, sink); bytes = std::move(sink.data); }
    else rabs_encode_forward(bits, (size_t)n, zero_prob, bytes);
    if (int st = give(bytes, out)) throw Error(st, "out of memory");
  });
}

// encode_symbols(symbols, _, SymbolEncodingMethod::DirectCoded, writer) on the device:
// histogram (K8) -> table (K9) -> rANS (K10). Output = method byte, bit_length byte,
// leb128 #symbols + table, leb128 payload size, payload — what the reference writes.
int dxo_encode_symbols(const uint32_t* symbols, uint64_t n, int device, dxo_bytes* out, float* kernel_ms /*[3] or NULL*/) {
  if (!symbols || !out || n == 0 || n > 0xFFFFFFF0ull) return DXO_ERR_INVALID_ARGUMENT;
  out->data = nullptr; out->len = 0;
  return guarded([&] {
    DeviceContext& ctx = DeviceContext::get(device);
    cudaStream_t s = ctx.stream[0];
    uint32_t mx = 0, nz = 0;
    for (uint64_t i = 0; i < n; ++i) { mx = std::max(mx, symbols[i]); nz += symbols[i] != 0; }
    if (mx >= (1u << 22)) throw Error(DXO_ERR_UNSUPPORTED_INPUT, "alphabet too large");
    const uint32_t cap = mx + 2;
    uint32_t *d_sym, *d_hist, *d_work; uint4* d_tab; uint8_t *d_tb, *d_pay; gpu::AttrStats* d_st;
    // every block and event is released on all paths, the error paths included
    struct Scratch {
      cudaStream_t s; std::vector<void*> blocks; std::vector<cudaEvent_t> events;
      ~Scratch() { for (void* p : blocks) cudaFreeAsync(p, s); for (cudaEvent_t e : events) cudaEventDestroy(e); cudaStreamSynchronize(s); cudaGetLastError(); }
    } scratch{s, {}, {}};
    auto alloc = [&](void** p, size_t b) { cuda_check(cudaMallocAsync(p, b, s), "cudaMallocAsync"); scratch.blocks.push_back(*p); };
    alloc((void**)&d_sym, n * 4); alloc((void**)&d_hist, cap * 4ull); alloc((void**)&d_work, cap * 12ull); alloc((void**)&d_tab, (cap + 1) * 16ull);
    alloc((void**)&d_tb, cap * 3ull + 16); alloc((void**)&d_pay, n * 3 + 16); alloc((void**)&d_st, sizeof(gpu::AttrStats));
    void* d_scr; alloc(&d_scr, gpu::rans_scratch_bytes(n));
    cuda_check(cudaMemcpyAsync(d_sym, symbols, n * 4, cudaMemcpyHostToDevice, s), "H2D");
    gpu::init_stats(d_st, s);
    cuda_check(cudaMemsetAsync(d_hist, 0, cap * 4ull, s), "memset");
    // the predict kernels normally produce these two scalars
    cuda_check(cudaMemcpyAsync(&d_st->nonzero_symbols, &nz, 4, cudaMemcpyHostToDevice, s), "H2D");
    cuda_check(cudaMemcpyAsync(&d_st->max_symbol, &mx, 4, cudaMemcpyHostToDevice, s), "H2D");
    cudaEvent_t ev[4];
    for (auto& e : ev) { cuda_check(cudaEventCreate(&e), "cudaEventCreate"); scratch.events.push_back(e); }
    cuda_check(cudaEventRecord(ev[0], s), "rec");
    gpu::launch_histogram(d_sym, n, d_hist, cap, d_st, s);
    cuda_check(cudaEventRecord(ev[1], s), "rec");
    gpu::launch_build_table(d_hist, cap, n, d_work, d_tab, d_tb, cap * 3 + 16, d_st, s);
    cuda_check(cudaEventRecord(ev[2], s), "rec");
    gpu::launch_rans_encode(d_sym, n, d_tab, cap, d_scr, d_pay, d_st, s);
    cuda_check(cudaEventRecord(ev[3], s), "rec");
    gpu::AttrStats st;
    cuda_check(cudaMemcpyAsync(&st, d_st, sizeof st, cudaMemcpyDeviceToHost, s), "D2H");
    cuda_check(cudaStreamSynchronize(s), "sync");
    if (kernel_ms) for (int k = 0; k < 3; ++k) cudaEventElapsedTime(&kernel_ms[k], ev[k], ev[k + 1]);
    if (getenv("DXO_RANS_DEBUG")) fprintf(stderr, "[dxo] rANS chunks=%u chain misses=%u fixup=%u\n", gpu::rans_num_chunks(n), st.pad[0], st.pad[1]);
    std::vector<uint8_t> tb(st.table_bytes), pay(st.payload_bytes);
    int status = DXO_OK;
    if (st.error_flags) status = (st.error_flags & gpu::kErrRansFreq) ? DXO_ERR_RANS_FREQ_TABLE : DXO_ERR_UNSUPPORTED_INPUT;
    else {
      cuda_check(cudaMemcpyAsync(tb.data(), d_tb, tb.size(), cudaMemcpyDeviceToHost, s), "D2H");
      cuda_check(cudaMemcpyAsync(pay.data(), d_pay, pay.size(), cudaMemcpyDeviceToHost, s), "D2H");
    }
    cuda_check(cudaStreamSynchronize(s), "sync");
    if (status != DXO_OK) throw Error(status, "device reported an entropy coding error");
    ByteSink w;
    w.u8(1);
    w.u8((uint8_t)st.bit_length);
    w.bytes(tb);
    w.varint(pay.size());
    w.bytes(pay);
    if (int st2 = give(w.data, out)) throw Error(st2, "out of memory");
  });
}

int dxo_corner_table_opposites(const uint32_t* vertex_of_corner, uint64_t num_faces, uint32_t* opposite_out, int* exact_out, int device) {
  if (!vertex_of_corner || !opposite_out || !exact_out || num_faces == 0 || num_faces > 0x2AAAAAAAull) return DXO_ERR_INVALID_ARGUMENT;
  return guarded([&] {
    DeviceContext& ctx = DeviceContext::get(device);
    cudaStream_t s = ctx.stream[0];
    const uint64_t C = num_faces * 3;
    uint32_t *d_cv = nullptr, *d_opp = nullptr, *d_flag = nullptr;
    void* scratch = nullptr;
    const size_t sb = gpu::corner_table_scratch_bytes(C);
    cuda_check(cudaMallocAsync((void**)&d_cv, C * 4, s), "cudaMallocAsync");
    cuda_check(cudaMallocAsync((void**)&d_opp, C * 4, s), "cudaMallocAsync");
    cuda_check(cudaMallocAsync((void**)&d_flag, 4, s), "cudaMallocAsync");
    cuda_check(cudaMallocAsync(&scratch, sb, s), "cudaMallocAsync");
    cuda_check(cudaMemcpyAsync(d_cv, vertex_of_corner, C * 4, cudaMemcpyHostToDevice, s), "H2D");
    cuda_check(cudaMemsetAsync(d_flag, 0, 4, s), "memset");
    cuda_check(cudaMemsetAsync(d_opp, 0xFF, C * 4, s), "memset");
    gpu::launch_corner_table_opposites(d_cv, C, 0, d_opp, d_flag, scratch, sb, s);
    uint32_t flag = 0;
    cuda_check(cudaMemcpyAsync(opposite_out, d_opp, C * 4, cudaMemcpyDeviceToHost, s), "D2H");
    cuda_check(cudaMemcpyAsync(&flag, d_flag, 4, cudaMemcpyDeviceToHost, s), "D2H");
    cudaFreeAsync(d_cv, s); cudaFreeAsync(d_opp, s); cudaFreeAsync(d_flag, s); cudaFreeAsync(scratch, s);
    cuda_check(cudaStreamSynchronize(s), "cudaStreamSynchronize");
    *exact_out = flag ? 0 : 1;
  });
}

}  // extern "C"
