// Product host code — per-mesh encode job: host connectivity, device upload, the
// attribute kernels, download and stream assembly. See DESIGN.md for the data flow.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <condition_variable>
#include <deque>
#include <functional>
#include <thread>
#include <future>
#include <map>
#include <mutex>
#include <memory>
#include <string>

#include "connectivity.hpp"
#include "kernels.cuh"

namespace dxo {

void cuda_check(cudaError_t e, const char* what);
// process-wide pool of pinned host blocks (2 MB-aligned, huge-page advised, never freed)
void* pinned_block_take(size_t bytes, size_t* capacity);
void pinned_block_give(void* p, size_t capacity);
// one mesh through the per-mesh path (capi.cpp)
void encode_one_mesh(const dxo_mesh* mesh, const dxo_config& cfg, std::vector<uint8_t>& bytes, dxo_timing& tm, bool parallel_host);
// GLB assembly around the batch entry (glb.cpp)
void assemble_glb(const dxo_mesh* meshes, const dxo_bytes* streams, size_t n, std::vector<uint8_t>& out);
// the batch entry (batch.cpp): groups of meshes through segmented launches, sharded over GPUs
void encode_batch_grouped(const dxo_mesh* meshes, size_t n, const dxo_config& cfg, dxo_bytes* outs, int* statuses, int first_gpu, int num_gpus);

// Two persistent helper threads per DeviceContext (i.e. per calling thread): the side-stream coders of every step run
// on them, so a step neither creates threads nor leaves new threads at the mercy of the scheduler.
class HelperThreads {
 public:
  explicit HelperThreads(int n);
  ~HelperThreads();
  std::future<void> run(std::function<void()> fn);
 private:
  void loop();
  std::mutex mu_;
  std::condition_variable cv_;
  std::deque<std::packaged_task<void()>> queue_;
  std::vector<std::thread> threads_;
  bool stop_ = false;
};

// One per (host thread, device): streams, events and the launch / timing log.
struct DeviceContext {
  int device = 0;
  cudaStream_t stream[3] = {nullptr, nullptr, nullptr};  // attribute i runs on stream[min(i,2)]
  cudaStream_t copy_stream = nullptr;                    // early D2H of the side-stream flags
  cudaStream_t upload_stream = nullptr;                  // H2D issued by helper threads while the host builds the connectivity
  cudaEvent_t ev_uploaded = nullptr, ev_inputs = nullptr, ev_serial = nullptr, ev_fork = nullptr, ev_copy_join = nullptr;
  // pinned host staging, reused across calls (slot = attribute index * 2 + {0: results, 1: side flags})
  std::vector<std::pair<uint8_t*, size_t>> pinned;
  uint8_t* pinned_buffer(size_t slot, size_t bytes);
  // Host waits go through these events. With DXO_BLOCKING_WAIT=1 they are created with cudaEventBlockingSync: the waiting
  // thread sleeps instead of spinning, for hosts with fewer cores than encoder threads (spinning is ~10 % faster otherwise).
  cudaEvent_t ev_block[3] = {nullptr, nullptr, nullptr};
  void wait_stream(int k);
  std::unique_ptr<HelperThreads> helpers;
  cudaEvent_t ev_begin = nullptr, ev_end = nullptr, ev_pos_ready = nullptr, ev_layout = nullptr, ev_join[3] = {nullptr, nullptr, nullptr};
  static DeviceContext& get(int device);  // thread-local; throws DXO_ERR_NO_DEVICE when there is no usable GPU
  DeviceContext() = default;
  DeviceContext(const DeviceContext&) = delete;
  DeviceContext& operator=(const DeviceContext&) = delete;
  ~DeviceContext();  // runs when the owning thread exits (batch workers): streams, events and staging buffers are released
};

struct KernelRecord { const char* name; uint64_t bytes; cudaEvent_t a, b; };

struct Profile {
  bool enabled = false;
  bool serial = false;  // profiling mode 2: attributes run one after the other, so each kernel is timed alone
  std::vector<KernelRecord> records;
  std::vector<cudaEvent_t> pool;
  size_t pool_used = 0;
  uint32_t launches = 0;
  cudaEvent_t take();
  void begin(const char* name, uint64_t bytes, cudaStream_t s);
  void end(cudaStream_t s);
  void reset() { records.clear(); pool_used = 0; launches = 0; }
  ~Profile();
};

enum class Scheme : uint8_t { Delta = 0, Parallelogram = 1, TexCoord = 5, Normal = 6 };           // prediction_scheme/mod.rs:74-86
enum class Transform : uint8_t { Difference = 0, Wrapped = 1, OctOrthogonal = 3 };                // prediction_transform/mod.rs:92-102
enum class Portabilization : uint8_t { ToBits = 1, Quantize = 2, Octahedral = 3 };                // portabilization/mod.rs:85-92

struct AttrPlan {
  AttrView view;
  Scheme scheme;
  Transform transform;
  Portabilization port;
  uint32_t bits = 0;          // quantization bits
  uint32_t ncomp_in = 0;      // components of the original values
  uint32_t ncomp_q = 0;       // components after portabilization
  int parent = -1;            // index of the position attribute this one predicts from
  // 4-byte values as the kernels read them: the caller's buffer, or — for quantised attributes whose components are not
  // f32 — their `to_f64() as f32` conversion (quantization_coordinate_wise.rs:30-90 converts every component that way)
  const void* values32 = nullptr;
  std::vector<float> converted;
  uint32_t hist_capacity = 0; // upper bound of the alphabet
  const TableRef* table = nullptr;
  U32Array sequence;  // empty when shared (see MeshJob::sequence_of)
  int shares_sequence_of = -1;     // index of the attribute whose sequence (host and device copy) this one uses
};

struct AttrDevice {
  // inputs
  float* values = nullptr; uint32_t* map = nullptr;
  uint32_t *corner_vertex = nullptr, *left_most = nullptr, *seq = nullptr; uint8_t* seam = nullptr; uint4* corner_vertex4 = nullptr; uint2* fan_link = nullptr;
  uint4* ring = nullptr; uint2* ring_head = nullptr; uint8_t* ring_count = nullptr;  // K5's flattened fans (resident sessions)
  uint4* tex_records = nullptr;  // K6's per-element operand indices (resident sessions)
  uint4* para_records = nullptr;  // K4's per-element operand indices (resident sessions)
  uint8_t* side_out = nullptr; void* side_scratch = nullptr; size_t side_scratch_bytes = 0;  // [8-byte scalars][flags (+1)] for the host coder
  // intermediates / outputs
  int32_t* quant = nullptr; uint32_t *rank = nullptr, *symbols = nullptr, *hist = nullptr, *work = nullptr;
  uint8_t* used = nullptr;  // per value: a sequence element refers to it (WrappedDifference bounds); static per mesh, like rank
  uint8_t *side = nullptr, *table_bytes = nullptr, *payload = nullptr, *rans_scratch = nullptr; uint4* rans_table = nullptr;
  gpu::AttrStats* stats = nullptr;
  uint64_t payload_capacity = 0; uint32_t table_capacity = 0;
};

struct AttrResult {
  gpu::AttrStats stats;
  const uint8_t* table_bytes = nullptr;  // pinned staging (valid until the next run on this context)
  const uint8_t* payload = nullptr;
  // pinned staging of the side stream input: [8 bytes: entries, ones | transitions][flags]; normals: one flip flag per
  // element, texcoords: the orientation values (1 = false, 2 = true) of the elements that have one, in order
  const uint8_t* side = nullptr;
  size_t side_len = 0;                   // elements of the attribute (capacity of the flag area)
  // binary side stream, coded on a host worker while the device runs K8-K10
  uint32_t side_count = 0;
  uint8_t side_zero_prob = 0;
  std::vector<uint8_t> side_payload, side_scratch;
  // group path: the side stream was coded on the device, its bytes live in the group's output block
  const uint8_t* side_bytes = nullptr;
  size_t side_bytes_len = 0;
};

class GroupRunner;
class MeshJob {
  friend class GroupRunner;  // batch.cpp: the same job driven through segmented launches over a group of meshes
 public:
  MeshJob(const dxo_mesh* mesh, const dxo_config& cfg);
  ~MeshJob();
  // phase 1: host — corner tables, Edgebreaker bytes, attribute sequences. With a device
  // context the half-edge matching runs as K12 (radix sort) when that is provably exact.
  void build_connectivity(DeviceContext* ctx = nullptr);
  // phase 2: device
  void upload(DeviceContext& ctx);
  void launch(DeviceContext& ctx, Profile& prof);
  // launch() with the fork onto / join from the attribute streams; every operation it issues can be stream-captured
  void launch_all(DeviceContext& ctx, Profile& prof, bool capturing = false);
  // Resident sessions replay the whole step (all streams, ~45 launches, memsets, flag copies, events) as ONE CUDA graph
  // from their second run on: one driver call per step instead of ~80, which matters most when several sessions share
  // the driver. Captured per (job, calling thread's context); re-captured when the session moves to another thread.
  void launch_graph(DeviceContext& ctx, Profile& prof);
  void download(DeviceContext& ctx);  // D2H of stats, tables, payloads, side bits (synchronises)
  // phase 3: host — assemble the Draco stream
  void assemble(std::vector<uint8_t>& out);
  void release(DeviceContext& ctx);

  bool trace = false;
  bool parallel_host = true;  // run independent host passes on their own threads (off inside batch workers)
  // With a device but many encodes in flight (concurrent callers): the same passes, device tables included, in order on
  // the calling thread — a dozen helper threads per call only fight each other for the cores.
  bool inline_host = false;
  // The job will run its device phase many times (sessions): upload() also flattens the fans of normal attributes (K5).
  bool resident = false;
  std::map<std::string, std::vector<uint8_t>> trace_items;
  uint64_t h2d_bytes = 0, d2h_bytes = 0;
  uint64_t num_position_vertices() const { return plans_.empty() ? 0 : plans_[0].sequence.size(); }
  const U32Array& sequence_of(size_t att) const {
    return plans_[att].shares_sequence_of >= 0 ? plans_[plans_[att].shares_sequence_of].sequence : plans_[att].sequence;
  }

 private:
  const dxo_mesh* mesh_;
  dxo_config cfg_;
  std::vector<AttrPlan> plans_;
  UniversalTable ut_;
  std::vector<SeamTable> seams_;
  std::vector<TableRef> table_refs_;
  std::vector<U8Array> interior_;  // per table: vertex_interior_flags (helper threads, during the traversal)
  ByteSink head_;  // header + connectivity + attribute section headers
  std::unique_ptr<EdgebreakerEncoder> eb_;  // kept: owns the corner list the sequencers (and traces) read
  std::vector<U32Array> masked_opposite_;   // per attribute table: opposite with seam edges removed (host sequencer)
  // device
  uint32_t *d_faces_ = nullptr, *d_opposite_ = nullptr, *d_corner_vertex_ = nullptr, *d_left_most_ = nullptr;
  uint4 *d_faces4_ = nullptr, *d_corner_vertex4_ = nullptr;
  bool vertex_is_point_ = false;
  std::vector<AttrDevice> dev_;
  std::vector<AttrResult> results_;
  std::vector<void*> allocations_;
  std::mutex alloc_mu_;              // dalloc / dupload are also called from the helper threads of build_connectivity
  // pinned host blocks holding the connectivity tables that K12-K14 produce (borrowed from a process-wide pool,
  // returned by the destructor)
  std::vector<std::pair<void*, size_t>> pinned_blocks_;
  std::atomic<int> pinned_copy_in_flight_{0};
  static void* pinned_source(void* user, size_t bytes);
  std::shared_future<void> inputs_upload_;  // faces, values and point maps travel while the host builds the tables
  bool device_seam_table(DeviceContext& ctx, size_t att);  // K14; false = not applicable / flagged, use the host pass
  bool device_seam_table_begin(DeviceContext& ctx, size_t att);  // its launches alone (device_seam_table then only collects)
  struct SeamPending { uint8_t* d_seam = nullptr; uint32_t *d_cv = nullptr, *d_lm = nullptr, *d_scalars = nullptr; bool begun = false; };
  std::vector<SeamPending> seam_pending_;
  void upload_inputs(DeviceContext& ctx);
  void upload_seam_table(DeviceContext& ctx, size_t att);
  cudaStream_t alloc_stream_ = nullptr;
  std::vector<cudaEvent_t> side_ready_, side_copied_;
  // pinned staging of the side-stream flags, owned by the job (a captured graph bakes the destination address in, so it
  // must not belong to a thread's context, whose buffers move when another job on that thread needs larger ones)
  std::vector<uint8_t*> side_host_;
  std::vector<size_t> side_host_cap_;
  bool uploaded_ = false;
  cudaGraphExec_t graph_exec_ = nullptr;
  DeviceContext* graph_ctx_ = nullptr;
  uint32_t graph_launches_ = 0;
  uint32_t layout_launches_ = 0;  // pad3 / fan-link kernels issued by upload()
  bool rank_in_w_ = false;        // position attribute: K4's fast path (see gpu::QuantDev::rank_in_w)
  cudaEvent_t ev_graph_done_ = nullptr;
  uint64_t graph_d2h_bytes_ = 0;
  bool capturing_ = false;
 public:
  int device_runs = 0;  // run_device_phase calls so far (the graph is built on the second)
  bool graph_replay = false;  // DXO_FLAG_GRAPH_REPLAY
 private:
  void encode_side_stream(size_t att);
  void encode_side_stream_pair(size_t a, size_t b);  // both in one interleaved loop (DXO_SIDE_INLINE: no helper threads)
  // the same from K5 / K6's raw per-element flags (flips, or 0 = none / 1 = false / 2 = true orientations): group path, long streams
  void encode_side_stream_from_flags(size_t att, const uint8_t* flags, size_t n);

  static uint32_t device_matcher(void* user, const uint32_t* corner_vertex, uint32_t num_faces, uint32_t num_vertices, uint32_t* opposite_out,
                                 uint32_t* left_most_out, uint8_t* interior_out, std::vector<uint32_t>* boundary_corners);
  DeviceContext* match_ctx_ = nullptr;
  template <class T> T* dalloc(size_t count, cudaStream_t s);
  template <class T> T* dupload(const T* host, size_t count, cudaStream_t s);
  gpu::TableDev table_dev(size_t att) const;
  void capture_trace(DeviceContext& ctx);
 public:
  void capture_host_trace();  // host-side results only (no device needed)
 private:
  void write_stream_header();                // "DRACO", version, geometry type, method, flags
  void write_attribute_section_headers();    // encode_attributes' header part
  void validate_attribute_indices(uint32_t max_face_point) const;  // every attribute covers the faces' points, every map entry is a valid value
 public:
  bool has_device_buffers() const { return uploaded_; }
};

}  // namespace dxo
