// Device-side interface of the attribute hot path (sm_100a). Host code includes this
// header; kernels live in kernels.cu and are compiled with -fmad=false so that every
// f32/f64 operation is a separately rounded IEEE operation (bit-exact integer outputs).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <vector>

namespace dxo {
namespace gpu {

constexpr uint32_t kNoneDev = 0xFFFFFFFFu;
// alphabets up to this many symbols are counted in shared-memory bins (2 * 2^12 + 4: up to 12 quantisation bits), larger ones with global atomics
constexpr uint32_t kSmemHistBins = 8200;

// error bits accumulated on the device (AttrStats::error_flags)
enum : uint32_t {
  kErrZeroNormal = 1u << 0,      // geom.rs:45 assert
  kErrFanWalk = 1u << 1,         // fan walk did not terminate (malformed table)
  kErrNegativeSymbol = 1u << 2,  // residual that the reference would index out of range
  kErrRansFreq = 1u << 3,        // normalised table does not sum to 2^P / zero frequency of a used symbol
  kErrRansState = 1u << 4,       // rans.rs:64 StateTooLarge
  kErrAlphabet = 1u << 5,        // alphabet larger than the histogram capacity
};

// Per-attribute scalars produced on the device and read back once by the host.
struct AttrStats {
  uint32_t vmin_bits[4];   // f32 bit patterns: per-component min (starts at +0.0)
  uint32_t vmax_bits[4];   // per-component max (starts at +0.0)
  float range;             // max_i(max_i - min_i)
  int32_t wrap_min;        // WrappedDifference min/max over visited originals
  int32_t wrap_max;
  uint32_t nonzero_symbols;
  uint32_t max_symbol;
  uint32_t error_flags;
  uint32_t bit_length;     // symbol_coding.rs:118
  uint32_t precision;      // rANS precision bits
  uint32_t num_table_symbols;
  uint32_t table_bytes;    // serialized leb128 #symbols + frequency table
  uint32_t payload_bytes;  // rANS payload
  uint32_t pad[3];
};

// Immutable connectivity of one attribute as the predictors see it (GenericCornerTable).
struct TableDev {
  const uint32_t* corner_point;   // faces
  const uint32_t* corner_vertex;  // attribute (or universal) vertex of each corner
  const uint4* corner_point4;     // per face {p0, p1, p2, -}: one 128-bit load per corner triple
  const uint4* corner_vertex4;    // per face {v0, v1, v2, -}
  const uint2* fan_link;  // optional: per corner {opposite with seam edges set to none, point of that corner} (fan walks of K5, K4's fast path)
  int vertex_is_point;            // corner_vertex == corner_point (no point map, seams or splits): vertex tuples are skipped
  const uint32_t* opposite;       // universal opposite corners
  const uint8_t* seam;            // nullptr for the universal table
  const uint32_t* left_most;      // per vertex
  uint32_t num_corners;
  uint32_t num_vertices;
  // Optional (K5 of a resident session; launch_normal_rings): the fan of every sequence element flattened into the polyline of
  // position-value indices around its vertex — consecutive entries span one face — so that the step gathers the positions
  // with independent loads instead of chasing one link per swing.
  const uint4* ring = nullptr;         // [2 * n]: up to 8 position-value indices per element
  const uint2* ring_head = nullptr;    // [n]: {position-value index of the vertex, value index of the attribute at the vertex}
  const uint8_t* ring_count = nullptr; // [n]: entries in use (2..8), 0xFF = more than 8 or an inconsistent fan: walk it
};

// Quantized attribute: AoS int32 values with a power-of-two stride (1, 2, 4, 4 ints for 1..4
// components, so one value is one vector load) + optional point map.
struct QuantDev {
  const int32_t* values;
  const uint32_t* map;  // nullptr = identity
  uint32_t num_components;
  // 3-component values are stored as int4: when value index == vertex index (no point map, no seams, no splits) the unused
  // w component carries the vertex's rank in the sequence, so that a value gather delivers the rank test's operand as well
  uint32_t rank_in_w = 0;
};

// ---- K1/K2: coordinate-wise quantization (quantization_coordinate_wise.rs:24-117) ----
void launch_minmax(const float* values, uint64_t num_values, uint32_t ncomp, AttrStats* stats, cudaStream_t s);
// w_init: what the padding component of 3-component values is set to (0, or 0xFFFFFFFF = "not in the sequence" with rank_in_w)
// used (optional): per value, 1 when a sequence element refers to it — the quantised components of those values are folded
//   into stats->wrap_min / wrap_max (WrappedDifference's bounds over the visited originals, wrapped_difference.rs:41-49);
// rank_w (optional, 3 components): per value the rank that goes into the padding component (QuantDev::rank_in_w).
void launch_quantize(const float* values, uint64_t num_values, uint32_t ncomp, uint32_t bits, int32_t* out, AttrStats* stats, cudaStream_t s, int32_t w_init = 0,
                     const uint8_t* used = nullptr, const uint32_t* rank_w = nullptr);
// ---- tables derived from the sequence alone (once per mesh, not per step): rank[vertex] = position in the sequence
// (0xFFFFFFFF = absent; the caller pre-sets it) and used[value] = 1 for every value a sequence element refers to (pre-set to 0)
void launch_sequence_tables(const uint32_t* seq, uint32_t n, TableDev t, const uint32_t* map, uint32_t* rank, uint8_t* used, cudaStream_t s);
// WrappedDifference bounds for values that are not quantised on the device (ToBits): min / max over the used values
void launch_wrap_minmax(const int32_t* values, uint64_t num_values, uint32_t ncomp, uint32_t stride, const uint8_t* used, AttrStats* stats, cudaStream_t s);
void launch_fan_link(const uint32_t* opposite, const uint8_t* seam, const uint32_t* corner_point, uint64_t n, uint2* out, cudaStream_t s);
// Fills TableDev::ring / ring_head / ring_count for the sequence of a normal attribute (a function of the connectivity and
// the point maps alone: once per upload). q / pos: only their point maps are read.
void launch_normal_rings(const uint32_t* seq, uint32_t n, TableDev t, QuantDev q, QuantDev pos, uint4* ring, uint2* ring_head, uint8_t* ring_count,
                         cudaStream_t s);
void launch_pad3(const uint32_t* in, uint64_t n_tuples, uint4* out, cudaStream_t s);  // 3-wide -> 16-byte tuples
// ---- K3: octahedral normal quantization (octahedral_quantization.rs:49-64) ----
void launch_oct_quantize(const float* normals, uint64_t num_values, int32_t* out, AttrStats* stats, cudaStream_t s);
// ---- sequence preparation: rank[vertex] = position in the sequence, WrappedDifference min/max ----
void launch_seq_prepare(const uint32_t* seq, uint32_t n, TableDev t, QuantDev q, uint32_t* rank, bool want_minmax, AttrStats* stats, cudaStream_t s);
// ---- K4-K7: prediction + transform + symbolization, one thread per sequence element ----
// K4 from records (resident sessions, see kernels.cu): per sequence element {value index of the vertex, next, prev, opposite}
// when the parallelogram applies, {vertex, vertex sequenced just before or 0xFFFFFFFF, 0xFFFFFFFF, -} otherwise.
void launch_parallelogram_records(const uint32_t* seq, uint32_t n, TableDev t, QuantDev q, const uint32_t* rank, uint4* records, cudaStream_t s);
void launch_predict_parallelogram_records(const uint4* records, uint32_t n, QuantDev q, uint32_t* symbols, AttrStats* stats, cudaStream_t s);
void launch_predict_parallelogram(const uint32_t* seq, uint32_t n, TableDev t, QuantDev q, const uint32_t* rank, uint32_t* symbols, AttrStats* stats, cudaStream_t s);
void launch_predict_normal(const uint32_t* seq, uint32_t n, TableDev t, QuantDev q, QuantDev pos, uint32_t* symbols, uint8_t* flips, AttrStats* stats, cudaStream_t s);
// K6 from records (resident sessions): launch_texcoord_records resolves, once per upload, everything that depends on the
// connectivity and the point maps alone — per sequence element two uint4: {value index of the vertex, of next, of prev, of the
// vertex sequenced just before (0xFFFFFFFF = none)}, {position-value index of the three (0xFFFFFFFF = outside), -, flags}.
void launch_texcoord_records(const uint32_t* seq, uint32_t n, TableDev t, QuantDev q, QuantDev pos, uint32_t pos_num_points, const uint32_t* rank,
                             uint4* records, cudaStream_t s);
void launch_predict_texcoord_records(const uint4* records, uint32_t n, QuantDev q, QuantDev pos, uint32_t* symbols, uint8_t* orient, AttrStats* stats,
                                     cudaStream_t s);
void launch_predict_texcoord(const uint32_t* seq, uint32_t n, TableDev t, QuantDev q, QuantDev pos, uint32_t pos_num_points, const uint32_t* rank, uint32_t* symbols, uint8_t* orient, AttrStats* stats, cudaStream_t s);
void launch_predict_delta(const uint32_t* seq, uint32_t n, TableDev t, QuantDev q, uint32_t* symbols, AttrStats* stats, cudaStream_t s);
// ---- K8: symbol histogram (symbol_coding.rs:149-157) ----
void launch_histogram(const uint32_t* symbols, uint64_t num_symbols, uint32_t* hist, uint32_t hist_capacity, AttrStats* stats, cudaStream_t s);
// ---- K9: probability table normalisation + serialisation + rANS lookup table (rans.rs:146-230) ----
// rans_table entries: {freq, cumulative, magic multiplier, shift}
void launch_build_table(const uint32_t* hist, uint32_t hist_capacity, uint64_t total_symbols, uint32_t* work /*3*capacity*/, uint4* rans_table,
                        uint8_t* table_bytes, uint32_t table_bytes_capacity, AttrStats* stats, cudaStream_t s);
// ---- K10: rANS emission, serial within the stream (rans.rs:33-68) ----
// scratch: rans_scratch_bytes(num_symbols) bytes of device memory (chunk byte strings + chunk states)
size_t rans_scratch_bytes(uint64_t num_symbols);
uint32_t rans_num_chunks(uint64_t num_symbols);
int rans_launch_count(uint64_t num_symbols);  // kernels launch_rans_encode issues
void launch_rans_encode(const uint32_t* symbols, uint64_t num_symbols, const uint4* rans_table, uint32_t table_capacity, void* scratch,
                        uint8_t* payload, AttrStats* stats, cudaStream_t s);
// ---- side-stream preparation (counts / ordered compaction of the K5 / K6 flags; the binary coding stays on the host) ----
// scalars: two words, [0] = entries, [1] = ones (flips) or transitions (orientations).
size_t side_prepare_scratch_bytes(uint32_t n);
void launch_count_flips(const uint8_t* flips, uint32_t n, uint32_t* scalars, cudaStream_t s);
void launch_compact_orientations(const uint8_t* flags, uint32_t n, uint8_t* compact /*n + 1*/, uint32_t* scalars, void* scratch, size_t scratch_bytes,
                                 cudaStream_t s);
// ---- K12: half-edge matching by radix sort (corner_table/mod.rs:252-340, fast path) ----
// keys/vals/tmp are caller-provided scratch (see corner_table_scratch_bytes).
size_t corner_table_scratch_bytes(uint64_t num_corners);
// num_vertices (0 = unknown) bounds the vertex ids, so that only 2 * ceil(log2 num_vertices) key bits are sorted
void launch_corner_table_opposites(const uint32_t* corner_vertex, uint64_t num_corners, uint32_t num_vertices, uint32_t* opposite, uint32_t* not_exact_flag,
                                   void* scratch, size_t scratch_bytes, cudaStream_t s);

// ---- K13: left-most corners (corner_table/mod.rs:342-416, single-fan vertices) ----
// flags (one word, zeroed by the caller): bit 0 = a vertex id below num_vertices is unused, bit 1 = a vertex has a
// second fan (the reference splits it; the caller must run the sequential pass).
size_t left_most_scratch_bytes(uint32_t num_vertices);
// ordered list of the corners whose opposite is none (list: num_corners entries of capacity; *count = entries used)
size_t boundary_list_scratch_bytes(uint64_t num_corners);
void launch_boundary_list(const uint32_t* opposite, uint64_t num_corners, void* scratch, size_t scratch_bytes, uint32_t* list, uint32_t* count,
                          cudaStream_t s);
// interior (optional, [num_vertices]): 1 when swing_left(left_most[v]) exists, i.e. the vertex is not on a boundary
void launch_left_most(const uint32_t* corner_vertex, const uint32_t* opposite, uint64_t num_corners, uint32_t num_vertices, void* scratch,
                      uint32_t* left_most, uint32_t* flags, cudaStream_t s, uint8_t* interior = nullptr);

// ---- K14: per-attribute seam table (attribute_corner_table.rs:16-137) from the device-resident universal table ----
// Outputs: seam[C], corner_vertex[C], left_most_a[<= C] (attribute vertex -> corner), *total = number of attribute
// vertices. flags (zeroed by the caller): bit 0 = a face references a point outside the attribute, bit 1 = a seam
// vertex whose fan closes on itself; either sends the caller to the sequential pass (which reports the error).
// bit 2 is informational: the attribute has a seam that is not a mesh boundary.
size_t seam_table_scratch_bytes(uint32_t num_vertices);
void launch_seam_table(const uint32_t* corner_point, const uint32_t* map, uint32_t num_points, const uint32_t* cv, const uint32_t* opposite,
                       const uint32_t* left_most_u, uint64_t num_corners, uint32_t num_vertices, void* scratch, size_t scratch_bytes,
                       uint8_t* seam, uint32_t* corner_vertex, uint32_t* left_most_a, uint32_t* total, uint32_t* flags, cudaStream_t s);

void init_stats(AttrStats* stats, cudaStream_t s);

// =======================================================================================
// Segmented launches over a GROUP of meshes (the batch entry, SURVEY §7 step 8 / north_star "across every attribute
// stream of a mesh batch"): one launch set per group instead of one per mesh. Every kernel takes an array of segment
// descriptors in device memory plus a tile list: CTA b works on elements [tile.first, tile.first + kSegTile) of segment
// tile.seg. All indices stay local to their mesh; descriptors carry the base pointers.
struct Tile { uint32_t seg, first; };
constexpr uint32_t kSegTile = 1024;  // elements per CTA

// per-mesh flags raised by the connectivity kernels (any of them sends the mesh to the sequential host passes)
enum : uint32_t {
  kMeshNotExact = 1u,       // K12: non-manifold edge, inconsistent orientation, degenerate face, equal tips
  kMeshUnusedVertex = 2u,   // K13: a vertex id below num_vertices is used by no corner
  kMeshSplitVertex = 4u,    // K13: a vertex with a second fan
  kMeshBadIndex = 8u,       // a face references a point / value outside the position attribute
};
struct MeshSeg {  // K12 + K13 of one mesh (CornerTable, core/corner_table/mod.rs:84-460)
  const uint32_t* faces;     // corner -> point, 3 per face
  const uint32_t* pos_map;   // point -> position value (nullptr = identity)
  uint32_t num_corners, num_points, num_vertices;
  uint32_t corner_base;      // first slot of this mesh in the group-wide sort arrays
  uint32_t* cv;              // out: corner -> vertex
  uint32_t* opposite;        // out (pre-set to 0xFF by the caller)
  uint32_t* left_most;       // out [num_vertices]
  uint8_t* interior;         // out [num_vertices]: swing_left(left_most[v]) exists
  uint32_t* first_corner;    // scratch [num_vertices], pre-set to 0xFF
  uint32_t* valence;         // scratch [num_vertices], pre-set to 0
  uint32_t* flags;           // out: one word, pre-set to 0
};
size_t seg_corner_tables_scratch_bytes(uint64_t total_corners);
// vertex_bits: bits that hold the largest num_vertices of the group; keys are mesh << 2 vertex_bits | min << vertex_bits | max
void launch_seg_corner_tables(const MeshSeg* segs, uint32_t num_meshes, const Tile* corner_tiles, uint32_t num_corner_tiles, const Tile* vertex_tiles,
                              uint32_t num_vertex_tiles, uint64_t total_corners, uint32_t vertex_bits, void* scratch, size_t scratch_bytes, cudaStream_t s);

struct SeamSeg {  // K14 of one non-position attribute (AttributeCornerTable, attribute_corner_table.rs:16-137)
  const uint32_t* faces; const uint32_t* map; uint32_t num_points;
  const uint32_t* cv_u; const uint32_t* opposite; const uint32_t* left_most_u;
  uint32_t num_corners, num_vertices_u;
  uint32_t count_base;       // first slot of this segment in the group-wide count / base arrays (num_vertices_u slots)
  uint32_t capacity;         // slots of left_most_a / interior_a
  uint8_t* seam;             // out [num_corners]
  uint8_t* vertex_on_seam;   // scratch [num_vertices_u], pre-set to 0
  uint32_t* cv_a;            // out [num_corners]
  uint32_t* left_most_a;     // out [capacity]
  uint8_t* interior_a;       // out [capacity]
  uint32_t* scalars;         // out: [0] attribute vertices, [1] flags (kSeam* bits of launch_seam_table + 8 = capacity exceeded), pre-set to 0
  const uint32_t* mesh_flags;  // the mesh's MeshSeg::flags: a flagged mesh is skipped
};
size_t seg_seam_tables_scratch_bytes(uint64_t total_count_slots);
void launch_seg_seam_tables(const SeamSeg* segs, uint32_t num_segs, const Tile* corner_tiles, uint32_t num_corner_tiles, const Tile* vertex_tiles,
                            uint32_t num_vertex_tiles, const Tile* attr_vertex_tiles, uint32_t num_attr_vertex_tiles, uint32_t* counts, uint32_t* bases,
                            uint64_t total_count_slots, void* scratch, size_t scratch_bytes, cudaStream_t s);

// One attribute stream of one mesh for K1-K10.
struct SideStats { uint32_t count, zero_prob, nbytes, pad; };
struct AttrSeg {
  TableDev t; QuantDev q; QuantDev pos; uint32_t pos_num_points;
  const float* values; uint32_t num_unique, ncomp_in, bits;
  uint32_t scheme, wrapped;          // Scheme enum value; transform is WrappedDifference
  const uint32_t* seq; uint32_t n;   // sequence (corners), elements
  uint32_t num_symbols;              // n * q.num_components
  uint32_t* rank; uint32_t* symbols; uint8_t* side_flags /* K5 flips / K6 orientation flags, [n] */;
  uint32_t* hist; uint32_t hist_capacity; uint32_t* work; uint4* rans_table; uint8_t* table_bytes; uint32_t table_capacity;
  AttrStats* stats;
  // fan links of K5 (normals): out [num_corners]
  uint2* fan_link_out; const uint8_t* seam_for_links;
  // binary side stream (device rABS): out
  uint8_t* side_payload; uint32_t side_capacity; SideStats* side_stats;
};
void launch_seg_init_stats(const AttrSeg* segs, uint32_t num_segs, cudaStream_t s);
struct Pad3Seg { const uint32_t* in; uint4* out; uint32_t n; };  // 3-wide rows -> 16-byte tuples (faces, corner -> vertex tables, 3-component ToBits values)
void launch_seg_pad3(const Pad3Seg* segs, const Tile* tiles, uint32_t num_tiles, cudaStream_t s);
void launch_seg_fan_links(const AttrSeg* segs, const Tile* tiles, uint32_t num_tiles, cudaStream_t s);       // tiles over corners of normal streams
void launch_seg_minmax(const AttrSeg* segs, const Tile* tiles, uint32_t num_tiles, uint32_t ncomp, cudaStream_t s);      // tiles over unique values
void launch_seg_quantize(const AttrSeg* segs, const Tile* tiles, uint32_t num_tiles, uint32_t ncomp, cudaStream_t s);
void launch_seg_oct_quantize(const AttrSeg* segs, const Tile* tiles, uint32_t num_tiles, cudaStream_t s);
void launch_seg_seq_prepare(const AttrSeg* segs, const Tile* tiles, uint32_t num_tiles, cudaStream_t s);                 // tiles over sequence elements
void launch_seg_predict(const AttrSeg* segs, const Tile* tiles, uint32_t num_tiles, uint32_t scheme, uint32_t ncomp, cudaStream_t s);
void launch_seg_histogram(const AttrSeg* segs, const Tile* tiles, uint32_t num_tiles, bool smem, cudaStream_t s);        // tiles over symbols
void launch_seg_build_tables(const AttrSeg* segs, uint32_t num_segs, cudaStream_t s);
void launch_seg_side_streams(const AttrSeg* segs, const uint32_t* seg_ids, uint32_t num_side_streams, cudaStream_t s);   // one warp per stream

// K10 over every stream of the group
struct RansJob {
  const uint32_t* symbols; unsigned long long n; const uint4* table; uint8_t* scratch; uint8_t* payload; AttrStats* stats;
  uint32_t *start, *exit, *nbytes, *chain_start, *cand_start, *cand_exit, *cand_mid, *offset;
  uint32_t num_chunks, chunk_steps, warmup_steps, sub, num_pieces, piece_steps, pieces_per_cta, lanes;
};
RansJob rans_make_job(const uint32_t* symbols, uint64_t num_symbols, const uint4* rans_table, void* scratch, uint8_t* payload, AttrStats* stats);
struct RansTiles {  // built by the host from the jobs (rans_plan_tiles), uploaded with them
  std::vector<Tile> explore, chain, lanes, pairs, fixup, gather;
};
void rans_plan_tiles(const RansJob* jobs, uint32_t num_jobs, RansTiles& out);
struct RansTilesDev { const Tile *explore, *chain, *lanes, *pairs, *fixup, *gather; uint32_t n_explore, n_chain, n_lanes, n_pairs, n_fixup, n_gather; };
void launch_seg_rans(const RansJob* jobs, const RansTilesDev& tiles, uint32_t max_table_capacity, cudaStream_t s);

// Packs every stream's table bytes, rANS payload and side-stream bytes into one buffer (each part 4-byte aligned, in
// stream order) and leaves the layout in `index`: per stream {offset, table_bytes, payload_bytes, side_bytes}; index[num_segs].x = total.
void launch_seg_pack(const AttrSeg* segs, const RansJob* jobs, uint32_t num_segs, uint4* index, uint8_t* out, uint64_t out_capacity, const Tile* tiles,
                     uint32_t num_tiles, cudaStream_t s);

}  // namespace gpu
}  // namespace dxo
