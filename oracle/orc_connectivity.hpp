// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_core.hpp header).
// Corner tables, Edgebreaker connectivity encoder and the attribute sequencer,
// restated from the reference. Paths relative to /root/reference/draco-oxide/src/.
#pragma once
#include <algorithm>
#include "orc_core.hpp"

namespace orc {

inline uint32_t c_next(uint32_t c) { return (c % 3 == 2) ? c - 2 : c + 1; }      // corner_table/mod.rs:514-523
inline uint32_t c_prev(uint32_t c) { return (c % 3 == 0) ? c + 2 : c - 1; }      // :503-512

// GenericCornerTable — core/corner_table/mod.rs:8-52
struct GenericCornerTable {
  virtual ~GenericCornerTable() {}
  virtual size_t num_faces() const = 0;
  virtual size_t num_corners() const = 0;
  virtual size_t num_vertices() const = 0;
  virtual uint32_t point_idx(uint32_t c) const = 0;
  virtual uint32_t vertex_idx(uint32_t c) const = 0;
  virtual uint32_t opposite(uint32_t c) const = 0;  // NONE when absent
  virtual uint32_t left_most_corner(uint32_t v) const = 0;
  uint32_t face_of(uint32_t c) const { return c / 3; }
  uint32_t swing_right(uint32_t c) const { uint32_t o = opposite(c_prev(c)); return o == NONE ? NONE : c_prev(o); }  // :20-26
  uint32_t swing_left(uint32_t c) const { uint32_t o = opposite(c_next(c)); return o == NONE ? NONE : c_next(o); }   // :28-34
  bool is_on_boundary(uint32_t v) const { return swing_left(left_most_corner(v)) == NONE; }                           // :36-38
  uint32_t get_left_corner(uint32_t c) const { return opposite(c_prev(c)); }                                          // :40-42
  uint32_t get_right_corner(uint32_t c) const { return opposite(c_next(c)); }                                         // :44-46
};

// CornerTable — core/corner_table/mod.rs:54-529
struct CornerTable : GenericCornerTable {
  std::vector<uint32_t> opposite_corners;
  const std::vector<std::array<uint32_t, 3>>* mesh_faces = nullptr;
  std::vector<std::array<uint32_t, 3>> conn_faces;
  size_t n_corners = 0, n_vertices = 0;
  std::vector<uint32_t> left_most_corners;
  std::vector<uint32_t> corner_to_vertex_override;  // dense stand-in for BTreeMap<CornerIdx,VertexIdx> (:77); NONE = absent
  std::vector<uint32_t> non_manifold_vertex_parents;

  size_t num_faces() const override { return mesh_faces->size(); }
  size_t num_corners() const override { return n_corners; }
  size_t num_vertices() const override { return n_vertices; }
  uint32_t point_idx(uint32_t c) const override { return (*mesh_faces)[c / 3][c % 3]; }   // :484-487
  uint32_t corner_to_vert(uint32_t c) const {                                              // :443-459
    if (!corner_to_vertex_override.empty() && corner_to_vertex_override[c] != NONE) return corner_to_vertex_override[c];
    return conn_faces[c / 3][c % 3];
  }
  uint32_t vertex_idx(uint32_t c) const override { return corner_to_vert(c); }
  uint32_t opposite(uint32_t c) const override { return opposite_corners[c]; }
  uint32_t left_most_corner(uint32_t v) const override { return left_most_corners[v]; }

  // CornerTable::new — :84-118
  CornerTable(const std::vector<std::array<uint32_t, 3>>& faces, const Attribute& pos_att) {
    mesh_faces = &faces;
    conn_faces.reserve(faces.size());
    for (auto& f : faces)
      conn_faces.push_back({pos_att.unique_val_idx(f[0]), pos_att.unique_val_idx(f[1]), pos_att.unique_val_idx(f[2])});
    n_corners = faces.size() * 3;
    // get_unused_vertices — :236-250
    {
      uint32_t mx = 0;
      for (auto& f : conn_faces) for (uint32_t v : f) mx = std::max(mx, v);
      std::vector<uint8_t> used((size_t)mx + 1, 0);
      for (auto& f : conn_faces) for (uint32_t v : f) used[v] = 1;
      for (size_t i = 0; i < used.size(); ++i)
        if (!used[i]) throw Panic(ST_UNUSED_VERTICES, "Mesh contains unused vertices. This is not supported by the corner table.");
    }
    compute_table();
    if (contains_non_manifold_edges(conn_faces)) handle_no_manifold_edges();
    compute_left_most_corners();
  }

  // contains_non_manifold_edges — :121-145
  static bool contains_non_manifold_edges(const std::vector<std::array<uint32_t, 3>>& faces) {
    std::vector<uint64_t> edges;
    edges.reserve(faces.size() * 3);
    for (auto& f : faces) {
      uint32_t v[3] = {f[0], f[1], f[2]};
      for (int k = 0; k < 3; ++k) {
        uint32_t a = v[k], b = v[(k + 1) % 3];
        if (a > b) std::swap(a, b);
        edges.push_back(((uint64_t)a << 32) | b);
      }
    }
    std::sort(edges.begin(), edges.end());
    int count = 1;
    for (size_t i = 1; i < edges.size(); ++i) {
      if (edges[i] == edges[i - 1]) { if (++count > 2) return true; }
      else count = 1;
    }
    return false;
  }

  // compute_table — :252-340 (Draco's per-vertex bucketed half-edge matcher)
  void compute_table() {
    const uint32_t DV = NONE, DC = NONE;
    opposite_corners.assign(n_corners, DC);
    std::vector<size_t> num_on;  // num_corners_on_vertices
    for (uint32_t c = 0; c < n_corners; ++c) {
      uint32_t v1 = vertex_idx(c);
      if (v1 >= num_on.size()) num_on.resize((size_t)v1 + 1, 0);
      num_on[v1] += 1;
    }
    struct HE { uint32_t v, c; };
    std::vector<HE> vertex_edges(n_corners, HE{DV, DC});
    std::vector<size_t> vertex_offset(num_on.size());
    { size_t off = 0; for (size_t i = 0; i < num_on.size(); ++i) { vertex_offset[i] = off; off += num_on[i]; } }
    for (uint32_t c = 0; c < n_corners; ++c) {
      uint32_t tip_v = vertex_idx(c);
      uint32_t source_v = vertex_idx(c_next(c));
      uint32_t sink_v = vertex_idx(c_prev(c));
      uint32_t f_idx = c / 3;
      if (c == f_idx * 3) {
        uint32_t v0 = vertex_idx(c);
        if (v0 == source_v || v0 == sink_v || source_v == sink_v) continue;  // skip degenerate corners (:290-295)
      }
      uint32_t opposite_c = DC;
      size_t n = num_on[sink_v];
      size_t offset = vertex_offset[sink_v];
      for (size_t i = 0; i < n; ++i) {
        uint32_t other_v = vertex_edges[offset].v;
        if (other_v == DV) break;
        if (other_v == source_v) {
          // :308-310 — `continue` without advancing `offset`: the same entry is
          // re-examined until `i` runs out, i.e. no match is found for this corner.
          if (tip_v == vertex_idx(vertex_edges[offset].c)) continue;
          opposite_c = vertex_edges[offset].c;
          for (size_t k = i + 1; k < n; ++k) {
            vertex_edges[offset] = vertex_edges[offset + 1];
            if (vertex_edges[offset].v == DV) break;
            offset += 1;
          }
          vertex_edges[offset].v = DV;
          break;
        }
        offset += 1;
      }
      if (opposite_c == DC) {
        size_t n2 = num_on[source_v];
        size_t first = vertex_offset[source_v];
        for (size_t k = first; k < first + n2; ++k) {
          if (vertex_edges[k].v == DV) { vertex_edges[k].v = sink_v; vertex_edges[k].c = c; break; }
        }
      } else {
        opposite_corners[c] = opposite_c;
        opposite_corners[opposite_c] = c;
      }
    }
    n_vertices = num_on.size();
  }

  // handle_no_manifold_edges — :149-234
  void handle_no_manifold_edges() {
    std::vector<uint8_t> visited_corners(n_corners, 0);
    std::vector<std::pair<uint32_t, uint32_t>> sink_vertices;
    bool connectivity_updated;
    for (;;) {
      connectivity_updated = false;
      for (uint32_t ci = 0; ci < n_corners; ++ci) {
        if (visited_corners[ci]) continue;
        uint32_t c = ci;
        sink_vertices.clear();
        uint32_t first_c = c, curr_c = c;
        for (;;) {
          uint32_t next_c = swing_left(curr_c);
          if (next_c == NONE) break;
          if (next_c == first_c || visited_corners[next_c]) break;
          curr_c = next_c;
        }
        first_c = curr_c;
        for (;;) {
          visited_corners[curr_c] = 1;
          uint32_t sink_c = c_next(curr_c);
          uint32_t sink_v = corner_to_vert(sink_c);
          uint32_t edge_c = c_prev(curr_c);
          bool vertex_connectivity_updated = false;
          for (auto& att : sink_vertices) {
            if (att.first == sink_v) {
              uint32_t other_edge_c = att.second;
              uint32_t opp_edge_c = opposite(edge_c);
              if (opp_edge_c != NONE && opp_edge_c == other_edge_c) continue;
              uint32_t opp_other_edge_c = opposite(other_edge_c);
              if (opp_edge_c != NONE) opposite_corners[opp_edge_c] = NONE;
              if (opp_other_edge_c != NONE) opposite_corners[opp_other_edge_c] = NONE;
              opposite_corners[edge_c] = NONE;
              opposite_corners[other_edge_c] = NONE;
              vertex_connectivity_updated = true;
              break;
            }
          }
          if (vertex_connectivity_updated) { connectivity_updated = true; break; }
          sink_vertices.push_back({corner_to_vert(c_prev(curr_c)), sink_c});
          uint32_t nx = swing_right(curr_c);
          if (nx == NONE) break;
          curr_c = nx;
          if (curr_c == first_c) break;
        }
      }
      if (!connectivity_updated) break;
    }
  }

  // compute_left_most_corners — :342-416
  void compute_left_most_corners() {
    left_most_corners.assign(n_vertices, NONE);
    std::vector<uint8_t> visited_vertices(n_vertices, 0);
    std::vector<uint8_t> visited_corners(n_corners, 0);
    const size_t nf = mesh_faces->size();
    for (size_t f = 0; f < nf; ++f) {
      for (int i = 0; i < 3; ++i) {
        uint32_t c = (uint32_t)(3 * f + i);
        if (visited_corners[c]) continue;
        uint32_t v = vertex_idx(c);
        bool is_non_manifold_vertex = false;
        if (visited_vertices[v]) {
          left_most_corners.push_back(NONE);
          non_manifold_vertex_parents.push_back(v);
          visited_vertices.push_back(0);
          v = (uint32_t)n_vertices;
          n_vertices += 1;
          is_non_manifold_vertex = true;
          if (corner_to_vertex_override.empty()) corner_to_vertex_override.assign(n_corners, NONE);
        }
        visited_vertices[v] = 1;
        visited_corners[c] = 1;
        left_most_corners[v] = c;
        if (is_non_manifold_vertex) corner_to_vertex_override[c] = v;
        uint32_t act_c = swing_left(c);
        while (act_c != NONE) {
          if (act_c == c) break;
          visited_corners[act_c] = 1;
          left_most_corners[v] = act_c;
          if (is_non_manifold_vertex) corner_to_vertex_override[act_c] = v;
          act_c = swing_left(act_c);
        }
        if (act_c == NONE) {
          act_c = c;
          while (act_c != NONE) {
            visited_corners[act_c] = 1;
            if (is_non_manifold_vertex) corner_to_vertex_override[act_c] = v;
            act_c = swing_right(act_c);
          }
        }
      }
    }
  }
};

// AttributeCornerTable — core/corner_table/attribute_corner_table.rs:4-192
struct AttributeCornerTable {
  std::vector<uint32_t> corner_to_vertex;
  std::vector<uint32_t> vertex_to_attribute_map;
  std::vector<uint8_t> is_edge_on_seam;
  std::vector<uint8_t> is_vertex_on_seam;
  std::vector<uint32_t> left_most_corners;
  size_t n_vertices = 0;

  AttributeCornerTable(const CornerTable& ct, const Attribute& att) {  // new — :16-77
    const size_t nc = ct.num_corners();
    is_edge_on_seam.assign(nc, 0);
    is_vertex_on_seam.assign(ct.num_vertices(), 0);
    for (uint32_t c = 0; c < nc; ++c) {
      uint32_t opp = ct.opposite(c);
      if (opp == NONE) {
        is_edge_on_seam[c] = 1;
        is_vertex_on_seam[ct.vertex_idx(c_next(c))] = 1;
        is_vertex_on_seam[ct.vertex_idx(c_prev(c))] = 1;
        continue;
      }
      if (opp < c) continue;
      uint32_t c1 = c, c2 = opp;
      for (int k = 0; k < 2; ++k) {
        c1 = c_next(c1);
        c2 = c_prev(c2);
        uint32_t i1 = ct.point_idx(c1), i2 = ct.point_idx(c2);
        if (att.unique_val_idx(i1) != att.unique_val_idx(i2)) {
          is_edge_on_seam[c] = 1;
          is_edge_on_seam[opp] = 1;
          is_vertex_on_seam[ct.vertex_idx(c_next(c))] = 1;
          is_vertex_on_seam[ct.vertex_idx(c_prev(c))] = 1;
          is_vertex_on_seam[ct.vertex_idx(c_next(opp))] = 1;
          is_vertex_on_seam[ct.vertex_idx(c_prev(opp))] = 1;
          break;
        }
      }
    }
    corner_to_vertex.assign(nc, 0);
    n_vertices = ct.num_vertices();
    recompute_vertices(att, ct);
  }

  bool seam(uint32_t c) const { return is_edge_on_seam[c] != 0; }                                // :184-186
  uint32_t opposite(uint32_t c, const CornerTable& ct) const { return seam(c) ? NONE : ct.opposite(c); }  // :160-166
  uint32_t swing_left(uint32_t c, const CornerTable& ct) const {                                 // :177-183
    uint32_t o = opposite(c_next(c), ct);
    return o == NONE ? NONE : c_next(o);
  }
  uint32_t swing_right(uint32_t c, const CornerTable& ct) const {                                // :169-175
    uint32_t o = opposite(c_prev(c), ct);
    return o == NONE ? NONE : c_prev(o);
  }

  void recompute_vertices(const Attribute& att, const CornerTable& ct) {  // :79-137
    vertex_to_attribute_map.clear();
    left_most_corners.clear();
    uint32_t num_new_vertices = 0;
    for (uint32_t v = 0; v < ct.num_vertices(); ++v) {
      uint32_t c = ct.left_most_corner(v);
      uint32_t first_vert_id = num_new_vertices++;
      vertex_to_attribute_map.push_back(att.unique_val_idx(ct.point_idx(c)));
      uint32_t first_c = c;
      if (is_vertex_on_seam[v]) {
        uint32_t curr = swing_left(first_c, ct);
        while (curr != NONE) {
          first_c = curr;
          if (curr == c) throw Panic(ST_INTERNAL, "Swinging left from the left most corner should never return the same corner.");
          curr = swing_left(curr, ct);
        }
      }
      corner_to_vertex[first_c] = first_vert_id;
      left_most_corners.push_back(first_c);
      uint32_t curr = ct.swing_right(first_c);
      while (curr != NONE) {
        if (curr == first_c) break;
        if (seam(c_next(curr))) {
          first_vert_id = num_new_vertices++;
          vertex_to_attribute_map.push_back(att.unique_val_idx(ct.point_idx(curr)));
          left_most_corners.push_back(curr);
        }
        corner_to_vertex[curr] = first_vert_id;
        curr = ct.swing_right(curr);
      }
    }
    n_vertices = num_new_vertices;
  }
};

// RefAttributeCornerTable — core/corner_table/all_inclusive_corner_table.rs:54-109
struct RefAttributeCornerTable : GenericCornerTable {
  const CornerTable& u;
  const AttributeCornerTable& a;
  RefAttributeCornerTable(const CornerTable& u_, const AttributeCornerTable& a_) : u(u_), a(a_) {}
  size_t num_faces() const override { return u.num_faces(); }
  size_t num_corners() const override { return u.num_corners(); }
  size_t num_vertices() const override { return a.n_vertices; }
  uint32_t point_idx(uint32_t c) const override { return u.point_idx(c); }
  uint32_t vertex_idx(uint32_t c) const override { return a.corner_to_vertex[c]; }
  uint32_t opposite(uint32_t c) const override { return a.opposite(c, u); }
  uint32_t left_most_corner(uint32_t v) const override { return a.left_most_corners[v]; }
};

// ---------------------------------------------------------------------------
// Edgebreaker<DefaultTraversal> — encode/connectivity/edgebreaker.rs:33-657
enum EbSymbol : uint8_t { SYM_C = 0, SYM_S = 1, SYM_L = 2, SYM_R = 3, SYM_E = 4 };  // symbol_encoder.rs:5-41

struct TopologySplit { uint64_t merging_symbol_idx, split_symbol_idx; int orientation; /*0 Left, 1 Right*/ };

struct EdgebreakerOutput {
  std::vector<uint32_t> corners_of_edgebreaker;
  std::vector<uint8_t> symbols;            // trace
  std::vector<TopologySplit> splits;       // trace
};

struct Edgebreaker {
  const CornerTable& ct;
  const std::vector<AttributeCornerTable>& att_data;
  std::vector<uint8_t> visited_vertices, visited_faces, visited_holes;
  std::vector<uint32_t> vertex_hole_id;  // NONE = not on a hole
  std::vector<uint32_t> corner_traversal_stack;
  uint64_t last_encoded_symbol_idx = ~(uint64_t)0;  // usize::MAX (:150)
  std::vector<uint32_t> processed_connectivity_corners;
  std::vector<uint64_t> face_to_split_symbol;  // BTreeMap<usize,usize> (:61) as dense array, ~0 = absent
  uint64_t num_split_symbols = 0;
  std::vector<uint32_t> init_face_connectivity_corners;
  std::vector<TopologySplit> topology_splits;
  // DefaultTraversal state (:544-548)
  std::vector<uint8_t> symbols;
  std::vector<uint8_t> interior_cfg;

  Edgebreaker(const CornerTable& ct_, const std::vector<AttributeCornerTable>& ad) : ct(ct_), att_data(ad) {
    visited_vertices.assign(ct.num_vertices(), 0);
    visited_faces.assign(ct.num_faces(), 0);
    face_to_split_symbol.assign(ct.num_faces(), ~(uint64_t)0);
  }

  void compute_boundaries() {  // :195-224
    vertex_hole_id.assign(ct.num_vertices(), NONE);
    for (uint32_t c0 = 0; c0 < ct.num_corners(); ++c0) {
      if (ct.opposite(c0) != NONE) continue;
      uint32_t v = ct.vertex_idx(c_next(c0));
      if (vertex_hole_id[v] != NONE) continue;
      uint32_t boundary_idx = (uint32_t)visited_holes.size();
      visited_holes.push_back(0);
      uint32_t c = c0;
      while (vertex_hole_id[v] == NONE) {
        vertex_hole_id[v] = boundary_idx;
        c = c_next(c);
        // :215-217 — the reference steps `c = next(c)` here (it stays inside the
        // face; Google Draco steps through `opposite`). Consequence: nearly every
        // boundary edge gets its own hole id. Reproduced literally.
        while (ct.opposite(c) != NONE) c = c_next(c);
        v = ct.vertex_idx(c_next(c));
      }
    }
  }

  size_t process_boundary(uint32_t start_corner, bool encode_first_vertex) {  // :226-256
    uint32_t corner = c_prev(start_corner);
    for (uint32_t opp; (opp = ct.opposite(corner)) != NONE;) corner = c_next(opp);
    uint32_t start_v = ct.vertex_idx(start_corner);
    size_t n = 0;
    if (encode_first_vertex) { visited_vertices[start_v] = 1; n++; }
    if (vertex_hole_id[start_v] == NONE) throw Panic(ST_INTERNAL, "unwrap on None: start_v is not on a hole");
    visited_holes[vertex_hole_id[start_v]] = 1;
    uint32_t curr_v = ct.vertex_idx(c_prev(corner));
    while (curr_v != start_v) {
      visited_vertices[curr_v] = 1;
      n++;
      corner = c_next(corner);
      for (uint32_t opp; (opp = ct.opposite(corner)) != NONE;) corner = c_next(opp);
      curr_v = ct.vertex_idx(c_prev(corner));
    }
    return n;
  }

  bool is_right_face_visited(uint32_t c) const {  // :355-361
    uint32_t r = ct.get_right_corner(c);
    return r == NONE ? true : visited_faces[r / 3] != 0;
  }
  bool is_left_face_visited(uint32_t c) const {  // :366-372
    uint32_t l = ct.get_left_corner(c);
    return l == NONE ? true : visited_faces[l / 3] != 0;
  }
  void check_and_store_topology_split_event(uint64_t merging_symbol_idx, int orientation, uint32_t split_face) {  // :434-448
    uint64_t s = face_to_split_symbol[split_face];
    if (s == ~(uint64_t)0) return;
    topology_splits.push_back({merging_symbol_idx, s, orientation});
  }

  void edgebreaker_from(uint32_t c) {  // :261-350
    corner_traversal_stack.clear();
    corner_traversal_stack.push_back(c);
    const size_t num_faces = ct.num_faces();
    while (!corner_traversal_stack.empty()) {
      c = corner_traversal_stack.back();
      if (visited_faces[c / 3]) { corner_traversal_stack.pop_back(); continue; }
      size_t num_visited_faces = 0;
      while (num_visited_faces < num_faces) {
        num_visited_faces++;
        last_encoded_symbol_idx += 1;  // wrapping_add(1)
        uint32_t face_idx = c / 3;
        visited_faces[face_idx] = 1;
        processed_connectivity_corners.push_back(c);
        uint32_t v = ct.vertex_idx(c);
        if (!visited_vertices[v]) {
          visited_vertices[v] = 1;
          if (vertex_hole_id[v] == NONE) {
            symbols.push_back(SYM_C);
            uint32_t r = ct.get_right_corner(c);
            if (r == NONE) throw Panic(ST_INTERNAL, "unwrap on None: right corner of a C face");
            c = r;
            continue;
          }
        }
        uint32_t right_c = ct.get_right_corner(c);
        uint32_t left_c = ct.get_left_corner(c);
        if (is_right_face_visited(c)) {
          if (right_c != NONE) check_and_store_topology_split_event(last_encoded_symbol_idx, 1, right_c / 3);
          if (is_left_face_visited(c)) {
            if (left_c != NONE) check_and_store_topology_split_event(last_encoded_symbol_idx, 0, left_c / 3);
            symbols.push_back(SYM_E);
            corner_traversal_stack.pop_back();
            break;
          } else {
            symbols.push_back(SYM_R);
            c = left_c;  // unwrap: left face unvisited => exists
          }
        } else {
          if (is_left_face_visited(c)) {
            if (left_c != NONE) check_and_store_topology_split_event(last_encoded_symbol_idx, 0, left_c / 3);
            symbols.push_back(SYM_L);
            c = right_c;
          } else {
            symbols.push_back(SYM_S);
            num_split_symbols++;
            uint32_t hole = vertex_hole_id[v];
            if (hole != NONE && !visited_holes[hole]) process_boundary(c, false);
            face_to_split_symbol[face_idx] = last_encoded_symbol_idx;
            corner_traversal_stack.back() = left_c;
            corner_traversal_stack.push_back(right_c);
            break;
          }
        }
      }
    }
  }

  std::pair<bool, uint32_t> begin_from(uint32_t face_idx) {  // :411-431
    uint32_t corner_index = 3 * face_idx;
    for (int k = 0; k < 3; ++k) {
      if (ct.opposite(corner_index) == NONE) return {false, corner_index};
      if (vertex_hole_id[ct.vertex_idx(corner_index)] != NONE) {
        uint32_t right = corner_index;
        while (right != NONE) { corner_index = right; right = ct.swing_right(right); }
        return {false, c_prev(corner_index)};
      }
      corner_index = c_next(corner_index);
    }
    return {true, corner_index};
  }

  void encode_topology_splits(Bytes& w) {  // :375-403
    uint64_t last_idx = 0;
    leb128_write(topology_splits.size(), w);
    for (auto& s : topology_splits) {
      leb128_write(s.merging_symbol_idx - last_idx, w);
      leb128_write(s.merging_symbol_idx - s.split_symbol_idx, w);
      last_idx = s.merging_symbol_idx;
    }
    BitWriterLsb bw(w);
    for (auto& s : topology_splits) bw.write_bits(1, s.orientation == 0 ? 0 : 1);
    bw.finish();
  }

  // DefaultTraversal::encode — :575-656
  void traversal_encode(Bytes& fw) {
    Bytes sym_bytes;
    {
      BitWriterLsb bw(sym_bytes);
      for (size_t i = symbols.size(); i-- > 0;) {
        switch (symbols[i]) {  // CrLight::encode_symbol — symbol_encoder.rs:50-58
          case SYM_C: bw.write_bits(1, 0); break;
          case SYM_S: bw.write_bits(3, 0b001); break;
          case SYM_L: bw.write_bits(3, 0b011); break;
          case SYM_R: bw.write_bits(3, 0b101); break;
          case SYM_E: bw.write_bits(3, 0b111); break;
        }
      }
      bw.finish();
    }
    leb128_write(sym_bytes.size(), fw);
    fw.insert(fw.end(), sym_bytes.begin(), sym_bytes.end());
    // start face configurations (:592-607)
    {
      size_t n0 = 0;
      for (uint8_t cfg : interior_cfg) if (!cfg) n0++;
      uint8_t zero_prob = zero_prob_f32(n0, (float)interior_cfg.size());
      w_u8(fw, zero_prob);
      RabsCoder rc(zero_prob);
      for (size_t i = interior_cfg.size(); i-- > 0;) rc.write(interior_cfg[i] ? 1 : 0);
      Bytes b = rc.flush();
      leb128_write(b.size(), fw);
      fw.insert(fw.end(), b.begin(), b.end());
    }
    // attribute seams (:610-653)
    std::vector<uint8_t> vf(ct.num_faces(), 0);
    std::vector<std::vector<uint8_t>> seams(att_data.size());
    for (size_t i = processed_connectivity_corners.size(); i-- > 0;) {
      uint32_t c = processed_connectivity_corners[i];
      uint32_t corners[3] = {c, c_next(c), c_prev(c)};
      vf[c / 3] = 1;
      for (int k = 0; k < 3; ++k) {
        uint32_t opp = ct.opposite(corners[k]);
        if (opp == NONE) continue;
        if (vf[opp / 3]) continue;
        for (size_t j = 0; j < att_data.size(); ++j) seams[j].push_back(att_data[j].opposite(corners[k], ct) == NONE ? 1 : 0);
      }
    }
    for (auto& sd : seams) {
      size_t n0 = 0;
      for (uint8_t s : sd) if (!s) n0++;
      uint8_t prob_zero = zero_prob_f32(n0, (float)sd.size());
      w_u8(fw, prob_zero);
      RabsCoder rc(prob_zero);
      for (size_t i = sd.size(); i-- > 0;) rc.write(sd[i] ? 1 : 0);
      Bytes b = rc.flush();
      leb128_write(b.size(), fw);
      fw.insert(fw.end(), b.begin(), b.end());
    }
  }

  // ConnectivityEncoder::encode_connectivity — :458-530
  EdgebreakerOutput encode_connectivity(size_t num_faces, Bytes& w) {
    w_u8(w, 0);  // EdgebreakerKind::Standard (:467)
    compute_boundaries();
    leb128_write(ct.num_vertices(), w);
    leb128_write(num_faces, w);
    w_u8(w, (uint8_t)att_data.size());
    for (uint32_t c = 0; c < ct.num_corners(); ++c) {
      uint32_t face_idx = c / 3;
      if (visited_faces[face_idx]) continue;
      auto bf = begin_from(face_idx);
      interior_cfg.push_back(bf.first ? 1 : 0);
      if (bf.first) {
        uint32_t ci = bf.second;
        visited_vertices[ct.vertex_idx(ci)] = 1;
        visited_vertices[ct.vertex_idx(c_next(ci))] = 1;
        visited_vertices[ct.vertex_idx(c_prev(ci))] = 1;
        visited_faces[face_idx] = 1;
        init_face_connectivity_corners.push_back(c_next(ci));
        uint32_t opp = ct.opposite(c_next(ci));
        if (opp == NONE) throw Panic(ST_INTERNAL, "unwrap on None: interior face without opposite");
        edgebreaker_from(opp);
      } else {
        process_boundary(c_next(bf.second), true);
        edgebreaker_from(bf.second);
      }
    }
    leb128_write(symbols.size(), w);
    leb128_write(num_split_symbols, w);
    encode_topology_splits(w);
    traversal_encode(w);
    EdgebreakerOutput out;
    out.symbols = symbols;
    out.splits = topology_splits;
    std::reverse(init_face_connectivity_corners.begin(), init_face_connectivity_corners.end());
    out.corners_of_edgebreaker = init_face_connectivity_corners;
    out.corners_of_edgebreaker.insert(out.corners_of_edgebreaker.end(), processed_connectivity_corners.begin(),
                                      processed_connectivity_corners.end());
    return out;
  }
};

// ---------------------------------------------------------------------------
// Traverser — shared/attribute/sequence.rs:4-151
// `literal` = true keeps the reference's per-face scan-and-remove of the stack
// (:98-131). With literal = false the removal is skipped: a removed entry's face
// is already marked visited, and every popped entry whose face is visited is
// discarded at :56-58 before any side effect, so the output is identical.
inline std::vector<uint32_t> compute_sequence(const GenericCornerTable& ct, std::vector<uint32_t> stack, bool literal) {
  std::vector<uint8_t> visited_vertices(ct.num_vertices(), 0), visited_faces(ct.num_faces(), 0);
  std::vector<uint32_t> out;
  out.reserve(ct.num_vertices());
  auto visit = [&](uint32_t v, uint32_t c) { if (!visited_vertices[v]) out.push_back(c); visited_vertices[v] = 1; };
  auto remove_face_entries = [&](uint32_t face_idx) {
    if (!literal) return;
    for (size_t i = stack.size(); i-- > 0;) if (stack[i] / 3 == face_idx) stack.erase(stack.begin() + (long)i);
  };
  while (!stack.empty()) {
    uint32_t curr = stack.back();
    stack.pop_back();
    uint32_t v = ct.vertex_idx(curr);
    if (visited_faces[curr / 3]) continue;
    uint32_t next_c = c_next(curr), prev_c = c_prev(curr);
    uint32_t next_v = ct.vertex_idx(next_c), prev_v = ct.vertex_idx(prev_c);
    if (!visited_vertices[next_v] || !visited_vertices[prev_v]) {
      visit(next_v, next_c);
      visit(prev_v, prev_c);
      stack.push_back(curr);
      continue;
    }
    uint32_t face_idx = curr / 3;
    visited_faces[face_idx] = 1;
    if (!visited_vertices[v]) {
      visit(v, curr);
      if (!ct.is_on_boundary(v)) {
        uint32_t r = ct.get_right_corner(curr);
        if (r == NONE) throw Panic(ST_INTERNAL, "unwrap on None: right corner in sequencer");
        stack.push_back(r);
        continue;
      }
    }
    visit(v, curr);
    uint32_t right_corner = ct.get_right_corner(curr), left_corner = ct.get_left_corner(curr);
    bool right_visited = right_corner != NONE && visited_faces[right_corner / 3];
    bool left_visited = left_corner != NONE && visited_faces[left_corner / 3];
    if (right_visited) {
      remove_face_entries(face_idx);
      if (!left_visited) { if (left_corner != NONE) stack.push_back(left_corner); }
    } else {
      if (left_visited) {
        remove_face_entries(face_idx);
        if (right_corner != NONE) stack.push_back(right_corner);
      } else {
        if (left_corner != NONE) stack.push_back(left_corner);
        if (right_corner != NONE) stack.push_back(right_corner);
      }
    }
  }
  return out;
}

}  // namespace orc
