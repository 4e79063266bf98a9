// Product host code — connectivity side of encode(): corner tables, per-attribute
// seam tables, the Edgebreaker CLERS traversal and the attribute sequencer.
// north_star keeps these pointer-chasing passes on the host; everything here is
// flat uint32 arrays (no maps, no per-face vectors) so the device uploads are
// plain memcpys of the same buffers.
//
// Reference paths are relative to /root/reference/draco-oxide/src/.
#pragma once
#include <algorithm>
#include "common.hpp"

namespace dxo {

// Validated, borrowed view of one dxo_attribute.
struct AttrView {
  const dxo_attribute* raw = nullptr;
  uint32_t num_points = 0;   // Attribute::len()
  uint32_t num_unique = 0;
  const uint32_t* map = nullptr;  // nullptr = identity
  uint32_t value_of(uint32_t point) const { return map ? map[point] : point; }
};

// Universal (position) corner table — CornerTable, core/corner_table/mod.rs:54-529.
struct UniversalTable {
  uint32_t num_faces = 0, num_corners = 0, num_vertices = 0;
  uint32_t max_point = 0;                  // largest point index used by a face (build)
  const uint32_t* corner_point = nullptr;  // faces, 3 per face (borrowed)
  HostArray<uint32_t> corner_vertex;     // vertex_idx(c), non-manifold splits applied
  HostArray<uint32_t> opposite;          // kNone = boundary
  HostArray<uint32_t> left_most;         // per vertex
  HostArray<uint8_t> interior;           // per vertex: not on a boundary; filled by the device pass (has_interior), else computed by the caller
  void set_memory_source(HostArray<uint32_t>::Source fn, void* user) {  // e.g. pinned blocks for the device passes
    corner_vertex.set_source(fn, user); opposite.set_source(fn, user); left_most.set_source(fn, user); interior.set_source(fn, user);
  }

  uint32_t swing_left(uint32_t c) const { uint32_t o = opposite[corner_next(c)]; return o == kNone ? kNone : corner_next(o); }
  uint32_t swing_right(uint32_t c) const { uint32_t o = opposite[corner_prev(c)]; return o == kNone ? kNone : corner_prev(o); }

  // Optional accelerator (K12 + K13 on the device). Fills `opposite_out` and returns kMatchExact only when
  // its result provably equals the sequential matcher's (manifold, consistently oriented input);
  // with kLeftMostDone, `left_most_out[num_vertices]` also holds the left-most corners (every vertex has
  // a single fan) and all vertex ids are used; kUnusedVertices reports the reference's unused-vertex
  // panic. Whatever is not reported as done runs sequentially on the host.
  // With kBoundaryListDone, `boundary_corners` holds the corners without an opposite, ascending.
  enum : uint32_t { kMatchExact = 1u, kLeftMostDone = 2u, kUnusedVertices = 4u, kBoundaryListDone = 8u };
  // With kLeftMostDone, `interior_out[num_vertices]` holds the per-vertex interior flags as well (see vertex_interior_flags).
  using DeviceMatcher = uint32_t (*)(void* user, const uint32_t* corner_vertex, uint32_t num_faces, uint32_t num_vertices,
                                     uint32_t* opposite_out, uint32_t* left_most_out, uint8_t* interior_out, std::vector<uint32_t>* boundary_corners);
  bool has_interior = false;
  std::vector<uint32_t> boundary_corners;  // valid only when has_boundary_list
  bool has_boundary_list = false;
  void build(const uint32_t* faces, uint32_t nfaces, const AttrView& pos, DeviceMatcher matcher = nullptr, void* matcher_user = nullptr);
  bool matched_on_device = false;
  bool single_thread = false;  // no helper threads inside build() (many encodes in flight)

 private:
  void match_half_edges();
  bool has_non_manifold_edge() const;
  void break_non_manifold_edges();
  void assign_left_most_corners();
};

// Per-attribute seam table — AttributeCornerTable, core/corner_table/attribute_corner_table.rs:4-192.
struct SeamTable {
  uint32_t num_vertices = 0;
  // false: the only seams are mesh boundaries, every universal vertex maps to one attribute vertex. The table then
  // equals the universal one (same vertex ids, opposites and left-most corners) and so does its sequence.
  bool has_interior_seam = true;
  HostArray<uint32_t> corner_vertex;  // attribute vertex of each corner
  HostArray<uint8_t> seam;            // edge opposite to the corner is a seam (or boundary)
  HostArray<uint32_t> left_most;      // per attribute vertex
  void set_memory_source(HostArray<uint32_t>::Source fn, void* user) {
    corner_vertex.set_source(fn, user); seam.set_source(fn, user); left_most.set_source(fn, user);
  }
  void build(const UniversalTable& ut, const AttrView& att);
};

// What the traversal-order consumers need from either kind of table
// (GenericCornerTable, core/corner_table/mod.rs:8-52).
struct TableRef {
  uint32_t num_faces, num_corners, num_vertices;
  const uint32_t* corner_vertex;
  const uint32_t* opposite;   // universal opposites
  const uint8_t* seam;        // nullptr for the universal table
  const uint32_t* left_most;
  const uint32_t* opposite_masked = nullptr;  // optional: opposite with seam edges set to kNone (one read per opp())
  const uint8_t* interior = nullptr;  // optional, per vertex: swing_left(left_most[v]) exists (see vertex_interior_flags)
  uint32_t opp(uint32_t c) const { return opposite_masked ? opposite_masked[c] : (seam && seam[c]) ? kNone : opposite[c]; }
  bool is_interior(uint32_t v) const { return interior ? interior[v] != 0 : opp(corner_next(left_most[v])) != kNone; }
};
// !is_on_boundary(v) for every vertex of a table (corner_table/mod.rs:36-38); lets the sequencer replace two dependent
// random reads per vertex by one byte. Independent of the traversal, so it is computed on a helper thread.
U8Array vertex_interior_flags(const TableRef& t);
inline TableRef table_ref(const UniversalTable& u) {
  return {u.num_faces, u.num_corners, u.num_vertices, u.corner_vertex.data(), u.opposite.data(), nullptr, u.left_most.data()};
}
inline TableRef table_ref(const UniversalTable& u, const SeamTable& s) {
  return {u.num_faces, u.num_corners, s.num_vertices, s.corner_vertex.data(), u.opposite.data(), s.seam.data(), s.left_most.data()};
}

// Edgebreaker<DefaultTraversal>::encode_connectivity — encode/connectivity/edgebreaker.rs:458-657,
// split into phases so independent parts can run on different host threads:
//   traverse()            CLERS traversal (needs only the universal table)
//   write_head()          traversal byte .. start-face stream
//   write_seam_stream()   one rABS stream per non-position attribute (const, thread-safe)
class EdgebreakerRun;
class EdgebreakerEncoder {
 public:
  explicit EdgebreakerEncoder(const UniversalTable& ut);
  ~EdgebreakerEncoder();
  EdgebreakerEncoder(const EdgebreakerEncoder&) = delete;
  EdgebreakerEncoder& operator=(const EdgebreakerEncoder&) = delete;
  void traverse();
  // corners_of_edgebreaker = the init-face corners of interior-start components, last component first, followed by
  // every visited corner in visiting order. Kept as its two parts (no concatenated copy).
  struct CornerList {
    const uint32_t* init_reversed; size_t num_init;
    const uint32_t* visited; size_t num_visited;
    size_t size() const { return num_init + num_visited; }
    uint32_t operator[](size_t i) const { return i < num_init ? init_reversed[i] : visited[i - num_init]; }
  };
  CornerList corner_list() const;
  std::vector<uint32_t> corners_of_edgebreaker() const;  // concatenated copy (tests, traces)
  void write_head(ByteSink& w, size_t num_seam_tables) const;
  void write_seam_stream(const SeamTable& st, ByteSink& w) const;
 private:
  EdgebreakerRun* run_;
};
// Serial convenience wrapper: appends the whole connectivity section, returns corners_of_edgebreaker.
std::vector<uint32_t> encode_edgebreaker(const UniversalTable& ut, const std::vector<SeamTable>& seams, ByteSink& w);

// Traverser::compute_seqeunce — shared/attribute/sequence.rs:48-151.
// One corner per attribute vertex, in the order the decoder will reconstruct them.
U32Array attribute_sequence(const TableRef& t, const EdgebreakerEncoder::CornerList& corners_of_edgebreaker);

}  // namespace dxo
