// Product host code — see connectivity.hpp. Results must equal the reference's
// order-dependent algorithms exactly (the Edgebreaker symbols, split events and
// attribute sequences all depend on them), so each routine keeps the reference's
// visiting order while working on flat arrays.
// Reference paths are relative to /root/reference/draco-oxide/src/.
#include "connectivity.hpp"
#include <future>

#include <chrono>
#include <cstdio>
#include <cstdlib>

#include <mutex>
#include <unordered_map>
#include <sys/mman.h>

namespace dxo {

// ---------------------------------------------------------------------------------------
// Pool behind NoInitAllocator (common.hpp). Best fit among the cached blocks; a block is reused for requests down to
// half its size. Blocks of 2 MB and more are 2 MB-aligned and advised as transparent huge pages (DXO_NO_HUGEPAGES=1
// turns the advice off). At most kHostPoolCacheBytes stay cached, the rest goes back to the C library.
namespace {
constexpr size_t kHugePage = 2u << 20;
constexpr size_t kHostPoolCacheBytes = 6ull << 30;
struct HostBlockPool {
  std::mutex mu;
  std::vector<std::pair<void*, size_t>> free_blocks;
  std::unordered_map<void*, size_t> live;
  size_t cached = 0;
  void* take(size_t bytes) {
    {
      std::lock_guard<std::mutex> lock(mu);
      size_t best = free_blocks.size();
      for (size_t k = 0; k < free_blocks.size(); ++k)
        if (free_blocks[k].second >= bytes && free_blocks[k].second <= 2 * bytes + kHugePage &&
            (best == free_blocks.size() || free_blocks[k].second < free_blocks[best].second)) best = k;
      if (best != free_blocks.size()) {
        const auto b = free_blocks[best];
        free_blocks[best] = free_blocks.back();
        free_blocks.pop_back();
        cached -= b.second;
        live.emplace(b.first, b.second);
        return b.first;
      }
    }
    const bool huge = bytes >= kHugePage;
    const size_t align = huge ? kHugePage : 4096;
    const size_t cap = (bytes + align - 1) / align * align;
    void* p = aligned_alloc(align, cap);
    if (!p) throw std::bad_alloc();
    static const bool advise = getenv("DXO_NO_HUGEPAGES") == nullptr;
    if (huge && advise) madvise(p, cap, MADV_HUGEPAGE);  // best effort
    std::lock_guard<std::mutex> lock(mu);
    live.emplace(p, cap);
    return p;
  }
  void give(void* p) noexcept {
    std::lock_guard<std::mutex> lock(mu);
    auto it = live.find(p);
    if (it == live.end()) return;  // not ours (cannot happen)
    const size_t cap = it->second;
    live.erase(it);
    if (cached + cap > kHostPoolCacheBytes) { free(p); return; }
    cached += cap;
    free_blocks.push_back({p, cap});
  }
};
HostBlockPool& host_pool() { static HostBlockPool* pool = new HostBlockPool; return *pool; }  // never destroyed
}  // namespace
void* host_block_take(size_t bytes) { return host_pool().take(bytes); }
void host_block_give(void* p) noexcept { host_pool().give(p); }


// ---------------------------------------------------------------------------------------
// CornerTable::new — core/corner_table/mod.rs:84-118
void UniversalTable::build(const uint32_t* faces, uint32_t nfaces, const AttrView& pos, DeviceMatcher matcher, void* matcher_user) {
  num_faces = nfaces;
  num_corners = nfaces * 3u;
  corner_point = faces;
  const bool timing = getenv("DXO_TIMING") != nullptr;
  auto t0 = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!timing) return;
    auto n = std::chrono::steady_clock::now();
    fprintf(stderr, "[dxo]   %-26s %9.3f ms\n", what, std::chrono::duration<double, std::milli>(n - t0).count());
    t0 = n;
  };
  // vertex ids of the corners with the reference's range checks; large meshes split the passes over a few threads
  const uint32_t* map = pos.map;
  auto over_parts = [&](auto&& fn) -> uint32_t {  // max of fn(c0, c1) over the parts
    if (num_corners < (1u << 20) || single_thread) return fn(0u, num_corners);
    constexpr uint32_t kParts = 4;
    std::future<uint32_t> parts[kParts];
    for (uint32_t k = 0; k < kParts; ++k)
      parts[k] = std::async(std::launch::async, fn, (uint32_t)((uint64_t)num_corners * k / kParts), (uint32_t)((uint64_t)num_corners * (k + 1) / kParts));
    uint32_t mx = 0;
    for (auto& f : parts) mx = std::max(mx, f.get());
    return mx;
  };
  // Without a position map the vertex ids ARE the faces' point indices: with a device matcher (which uploads them and
  // leaves them alone) the array is the caller's own — no 12-byte-per-face copy; the sequential passes, which may
  // relabel split vertices, get their own copy further down.
  const bool alias_faces = !map && matcher != nullptr && num_corners > 0;
  uint32_t* cvw = nullptr;
  if (alias_faces) corner_vertex.adopt(const_cast<uint32_t*>(faces), num_corners);
  else { corner_vertex.resize(num_corners); cvw = corner_vertex.data(); }
  const uint32_t max_p = over_parts([&](uint32_t c0, uint32_t c1) {
    if (map || alias_faces) return max_u32(faces + c0, c1 - c0);
    uint32_t mx = 0;
    for (uint32_t c = c0; c < c1; ++c) { const uint32_t p = faces[c]; cvw[c] = p; mx = std::max(mx, p); }
    return mx;
  });
  if (num_corners && max_p >= pos.num_points) throw Error(DXO_ERR_INVALID_ARGUMENT, "face references a point outside the position attribute");
  max_point = max_p;
  uint32_t max_v = max_p;
  if (map) max_v = over_parts([&](uint32_t c0, uint32_t c1) {
    uint32_t mx = 0;
    for (uint32_t c = c0; c < c1; ++c) { const uint32_t v = map[faces[c]]; cvw[c] = v; mx = std::max(mx, v); }
    return mx;
  });
  if (num_corners && max_v >= pos.num_unique) throw Error(DXO_ERR_INVALID_ARGUMENT, "point_to_value entry out of range");
  num_vertices = num_corners ? max_v + 1u : 0u;
  // every vertex id up to the maximum must be used (get_unused_vertices, :236-250; panic :105-108);
  // the device pass (K13) reports this itself when it runs
  auto check_all_used = [&] {
    std::vector<uint8_t> used(num_vertices, 0);
    for (uint32_t c = 0; c < num_corners; ++c) used[corner_vertex[c]] = 1;
    for (uint32_t v = 0; v < num_vertices; ++v)
      if (!used[v]) throw Error(DXO_ERR_UNUSED_VERTICES, "mesh contains unused vertices");
  };
  lap("vertex ids + checks");
  bool nm = false;
  matched_on_device = false;
  has_boundary_list = false;
  has_interior = false;
  uint32_t device_done = 0;
  if (matcher) {
    opposite.resize(num_corners);
    left_most.resize(num_vertices);
    interior.resize(num_vertices);
    device_done = matcher(matcher_user, corner_vertex.data(), num_faces, num_vertices, opposite.data(), left_most.data(), interior.data(), &boundary_corners);
    if (device_done & kUnusedVertices) throw Error(DXO_ERR_UNUSED_VERTICES, "mesh contains unused vertices");
    matched_on_device = (device_done & kMatchExact) != 0;
    has_boundary_list = matched_on_device && (device_done & kBoundaryListDone);  // opposite[] is final: no non-manifold split follows
    has_interior = matched_on_device && (device_done & kLeftMostDone);
    lap(matched_on_device ? ((device_done & kLeftMostDone) ? "half edges + left-most (K12, K13)" : "half-edge matching (K12)") : "half-edge matching (K12, not exact)");
  }
  if (alias_faces && !(matched_on_device && (device_done & kLeftMostDone))) {  // the passes below may write vertex ids
    HostArray<uint32_t> own;
    own.resize(num_corners);
    memcpy(own.data(), faces, (size_t)num_corners * 4);
    corner_vertex = std::move(own);
  }
  if (!(matched_on_device && (device_done & kLeftMostDone))) { check_all_used(); lap("unused-vertex check"); }
  if (!matched_on_device) {
    match_half_edges();
    lap("half-edge matching");
    nm = has_non_manifold_edge();  // an exact K12 result implies no edge with 3+ sides
    lap("non-manifold edge test");
  }
  if (nm) break_non_manifold_edges();
  if (!(matched_on_device && (device_done & kLeftMostDone))) {
    assign_left_most_corners();
    lap("left-most corners");
  }
}

// compute_table — :252-340. Per-source-vertex buckets of open half edges; a corner
// looks in the bucket of its sink vertex for the first half edge coming back to its
// source vertex. Corner order and first-match semantics are the reference's.
void UniversalTable::match_half_edges() {
  const uint32_t C = num_corners, V = num_vertices;
  opposite.assign(C, kNone);
  std::vector<uint32_t> start(V + 1u, 0);
  for (uint32_t c = 0; c < C; ++c) start[corner_vertex[c] + 1u]++;
  for (uint32_t v = 0; v < V; ++v) start[v + 1u] += start[v];
  std::vector<uint32_t> he_far(C, kNone), he_corner(C, kNone);
  const uint32_t* cv = corner_vertex.data();
  for (uint32_t f = 0, c = 0; f < num_faces; ++f) {
    const uint32_t v3[3] = {cv[c], cv[c + 1], cv[c + 2]};
    for (uint32_t k = 0; k < 3; ++k, ++c) {
      const uint32_t tip = v3[k], src = v3[(k + 1) % 3], snk = v3[(k + 2) % 3];
      if (k == 0 && (tip == src || tip == snk || src == snk)) continue;  // :289-295 (first corner only)
      uint32_t found = kNone;
      const uint32_t n = start[snk + 1u] - start[snk];
      uint32_t off = start[snk];
      for (uint32_t i = 0; i < n; ++i, ++off) {
        const uint32_t far = he_far[off];
        if (far == kNone) break;
        if (far != src) continue;
        // :308-310 — a candidate with the same tip blocks the search for this corner.
        if (tip == cv[he_corner[off]]) break;
        found = he_corner[off];
        for (uint32_t j = i + 1; j < n; ++j) {  // close the gap (:312-319)
          he_far[off] = he_far[off + 1];
          he_corner[off] = he_corner[off + 1];
          if (he_far[off] == kNone) break;
          ++off;
        }
        he_far[off] = kNone;
        break;
      }
      if (found == kNone) {
        for (uint32_t s = start[src], e = start[src + 1u]; s < e; ++s)
          if (he_far[s] == kNone) { he_far[s] = snk; he_corner[s] = c; break; }
      } else {
        opposite[c] = found;
        opposite[found] = c;
      }
    }
  }
}

// contains_non_manifold_edges — :121-145: some undirected edge is used by 3+ face sides.
bool UniversalTable::has_non_manifold_edge() const {
  const uint32_t C = num_corners, V = num_vertices;
  std::vector<uint32_t> start(V + 1u, 0);
  const uint32_t* cv = corner_vertex.data();
  auto lo_hi = [&](uint32_t c, uint32_t& lo, uint32_t& hi) {
    const uint32_t a = cv[c], b = cv[corner_next(c)];
    lo = std::min(a, b); hi = std::max(a, b);
  };
  for (uint32_t c = 0; c < C; ++c) { uint32_t lo, hi; lo_hi(c, lo, hi); start[lo + 1u]++; }
  for (uint32_t v = 0; v < V; ++v) start[v + 1u] += start[v];
  std::vector<uint32_t> fill(start.begin(), start.end() - 1), far(C);
  for (uint32_t c = 0; c < C; ++c) { uint32_t lo, hi; lo_hi(c, lo, hi); far[fill[lo]++] = hi; }
  for (uint32_t v = 0; v < V; ++v) {
    uint32_t* b = far.data() + start[v];
    const uint32_t n = start[v + 1u] - start[v];
    if (n < 3) continue;
    std::sort(b, b + n);
    for (uint32_t i = 2; i < n; ++i) if (b[i] == b[i - 1] && b[i] == b[i - 2]) return true;
  }
  return false;
}

// handle_no_manifold_edges — :149-234.
void UniversalTable::break_non_manifold_edges() {
  const uint32_t C = num_corners;
  std::vector<uint8_t> seen(C, 0);
  std::vector<uint32_t> sink_v, sink_c;
  for (bool changed = true; changed;) {
    changed = false;
    for (uint32_t c0 = 0; c0 < C; ++c0) {
      if (seen[c0]) continue;
      sink_v.clear(); sink_c.clear();
      uint32_t first = c0, cur = c0;
      for (uint32_t nx; (nx = swing_left(cur)) != kNone;) {
        if (nx == first || seen[nx]) break;
        cur = nx;
      }
      first = cur;
      for (;;) {
        seen[cur] = 1;
        const uint32_t s_c = corner_next(cur), s_v = corner_vertex[s_c], edge_c = corner_prev(cur);
        bool updated = false;
        for (size_t k = 0; k < sink_v.size(); ++k) {
          if (sink_v[k] != s_v) continue;
          const uint32_t other_edge = sink_c[k];
          const uint32_t opp_edge = opposite[edge_c];
          if (opp_edge != kNone && opp_edge == other_edge) continue;
          const uint32_t opp_other = opposite[other_edge];
          if (opp_edge != kNone) opposite[opp_edge] = kNone;
          if (opp_other != kNone) opposite[opp_other] = kNone;
          opposite[edge_c] = kNone;
          opposite[other_edge] = kNone;
          updated = true;
          break;
        }
        if (updated) { changed = true; break; }
        sink_v.push_back(corner_vertex[corner_prev(cur)]);
        sink_c.push_back(s_c);
        const uint32_t nx = swing_right(cur);
        if (nx == kNone) break;
        cur = nx;
        if (cur == first) break;
      }
    }
  }
}

// compute_left_most_corners — :342-416. A second fan met at an already visited vertex
// becomes a new vertex (ids appended in face order); its corners are re-labelled in place.
void UniversalTable::assign_left_most_corners() {
  const uint32_t C = num_corners;
  left_most.assign(num_vertices, kNone);
  std::vector<uint8_t> vertex_seen(num_vertices, 0), corner_seen(C, 0);
  for (uint32_t c = 0; c < C; ++c) {
    if (corner_seen[c]) continue;
    uint32_t v = corner_vertex[c];
    bool split = false;
    if (vertex_seen[v]) {
      left_most.push_back(kNone);
      vertex_seen.push_back(0);
      v = num_vertices++;
      split = true;
    }
    vertex_seen[v] = 1;
    corner_seen[c] = 1;
    left_most[v] = c;
    if (split) corner_vertex[c] = v;
    uint32_t a = swing_left(c);
    while (a != kNone && a != c) {
      corner_seen[a] = 1;
      left_most[v] = a;
      if (split) corner_vertex[a] = v;
      a = swing_left(a);
    }
    if (a == kNone) {  // open fan: mark the corners to the right as well
      for (a = c; a != kNone; a = swing_right(a)) {
        corner_seen[a] = 1;
        if (split) corner_vertex[a] = v;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// AttributeCornerTable::new + recompute_vertices — attribute_corner_table.rs:16-137
void SeamTable::build(const UniversalTable& ut, const AttrView& att) {
  const uint32_t C = ut.num_corners;
  has_interior_seam = false;
  seam.assign(C, 0);
  std::vector<uint8_t> vertex_on_seam(ut.num_vertices, 0);
  const uint32_t* pt = ut.corner_point;
  const uint32_t* cv = ut.corner_vertex.data();
  auto val = [&](uint32_t corner) {
    const uint32_t p = pt[corner];
    if (p >= att.num_points) throw Error(DXO_ERR_INVALID_ARGUMENT, "face references a point outside an attribute");
    return att.value_of(p);
  };
  for (uint32_t c = 0; c < C; ++c) {
    const uint32_t o = ut.opposite[c];
    if (o == kNone) {  // mesh boundary counts as a seam
      seam[c] = 1;
      vertex_on_seam[cv[corner_next(c)]] = 1;
      vertex_on_seam[cv[corner_prev(c)]] = 1;
      continue;
    }
    if (o < c) continue;
    // the two end points of the shared edge must carry the same values on both faces
    const uint32_t cn = corner_next(c), cp = corner_prev(c), on = corner_next(o), op = corner_prev(o);
    if (val(cn) != val(op) || val(cp) != val(on)) {
      has_interior_seam = true;
      seam[c] = seam[o] = 1;
      vertex_on_seam[cv[cn]] = vertex_on_seam[cv[cp]] = 1;
      vertex_on_seam[cv[on]] = vertex_on_seam[cv[op]] = 1;
    }
  }
  auto a_opp = [&](uint32_t c) { return seam[c] ? kNone : ut.opposite[c]; };
  auto a_swing_left = [&](uint32_t c) { const uint32_t o = a_opp(corner_next(c)); return o == kNone ? kNone : corner_next(o); };

  corner_vertex.assign(C, 0);
  left_most.clear();
  left_most.reserve(ut.num_vertices);
  uint32_t next_id = 0;
  for (uint32_t v = 0; v < ut.num_vertices; ++v) {
    const uint32_t c = ut.left_most[v];
    uint32_t id = next_id++;
    uint32_t first = c;
    if (vertex_on_seam[v]) {  // rotate to the first corner after a seam, counter-clockwise
      for (uint32_t s = a_swing_left(first); s != kNone; s = a_swing_left(s)) {
        first = s;
        if (s == c) throw Error(DXO_ERR_UNSUPPORTED_INPUT, "seam vertex whose fan closes on itself");
      }
    }
    corner_vertex[first] = id;
    left_most.push_back(first);
    for (uint32_t s = ut.swing_right(first); s != kNone && s != first; s = ut.swing_right(s)) {
      if (seam[corner_next(s)]) {  // crossing a seam starts a new attribute vertex
        id = next_id++;
        left_most.push_back(s);
      }
      corner_vertex[s] = id;
    }
  }
  num_vertices = next_id;
}

// ---------------------------------------------------------------------------------------
// Edgebreaker — encode/connectivity/edgebreaker.rs
enum : uint8_t { kC = 0, kS = 1, kL = 2, kR = 3, kE = 4 };

struct SplitEvent { uint64_t merge_symbol, split_symbol; uint8_t right; };

class EdgebreakerRun {
 public:
  explicit EdgebreakerRun(const UniversalTable& ut) : ut_(ut) {
    vstate_.assign(ut.num_vertices, 0);
    face_done_.assign(ut.num_faces, 0);
    split_symbol_of_face_.assign(ut.num_faces, kNone);
  }

  // phase 1: the CLERS traversal of every connected component (needs only the universal table)
  void traverse_all() {
    const bool timing = getenv("DXO_TIMING") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
      if (!timing) return;
      auto n = std::chrono::steady_clock::now();
      fprintf(stderr, "[dxo]     %-24s %9.3f ms\n", what, std::chrono::duration<double, std::milli>(n - t0).count());
      t0 = n;
    };
    find_boundaries();
    lap("find boundaries");
    symbols_.resize(ut_.num_faces);      // every face gets exactly one symbol / one visit-order entry
    visit_order_.resize(ut_.num_faces);
    num_visited_ = 0;
    for (uint32_t c = 0; c < ut_.num_corners; ++c) {  // one traversal per connected component (:478-511)
      const uint32_t face = c / 3u;
      if (face_done_[face]) continue;
      bool interior;
      const uint32_t start = start_corner(face, interior);
      start_face_interior_.push_back(interior ? 1 : 0);
      if (interior) {
        vstate_[vtx(start)] |= 1; vstate_[vtx(corner_next(start))] |= 1; vstate_[vtx(corner_prev(start))] |= 1;
        face_done_[face] = 1;
        init_face_corners_.push_back(corner_next(start));
        const uint32_t o = ut_.opposite[corner_next(start)];
        if (o == kNone) throw Error(DXO_ERR_INTERNAL, "interior start face without neighbour");
        traverse(o);
      } else {
        walk_boundary(corner_next(start), true);
        traverse(start);
      }
    }
    lap("CLERS loop");
    symbols_.resize(num_visited_);
    visit_order_.resize(num_visited_);
    init_reversed_.assign(init_face_corners_.rbegin(), init_face_corners_.rend());
    lap("corner list");
  }
  EdgebreakerEncoder::CornerList corner_list() const { return {init_reversed_.data(), init_reversed_.size(), visit_order_.data(), visit_order_.size()}; }

  // phase 2a: everything of the connectivity section up to and including the start-face stream
  void write_head(ByteSink& w, size_t num_seam_tables) const {
    w.u8(0);  // EdgebreakerKind::Standard (:467)
    w.varint(ut_.num_vertices);
    w.varint(ut_.num_faces);
    w.u8((uint8_t)num_seam_tables);
    w.varint(symbols_.size());
    w.varint(num_splits_);
    write_split_events(w);
    write_symbols_and_start_faces(w);
  }

 private:
  const UniversalTable& ut_;
  // vstate_: bit 0 = vertex visited, bit 1 = vertex lies on a hole (hole_of_vertex_ != kNone);
  // face_done_: bit 0 = face visited, bit 1 = the face got an S symbol (split_symbol_of_face_ is set)
  U8Array vstate_, face_done_;  // pooled blocks (common.hpp)
  std::vector<uint8_t> hole_done_, start_face_interior_;
  U32Array hole_of_vertex_;
  std::vector<uint32_t> stack_, init_face_corners_, init_reversed_;
  U8Array symbols_;        // sized once, every used entry written by the traversal (no zero-fill)
  U32Array visit_order_;
  U32Array split_symbol_of_face_;  // symbol index of the S symbol a face got (< num_faces), kNone otherwise
  std::vector<SplitEvent> split_events_;
  uint64_t symbol_index_ = ~(uint64_t)0;  // usize::MAX, incremented with wrap before use (:150,:276)
  size_t num_visited_ = 0;                // entries of symbols_ / visit_order_ in use
  uint64_t num_splits_ = 0;

  uint32_t vtx(uint32_t c) const { return ut_.corner_vertex[c]; }
  uint32_t right_of(uint32_t c) const { return ut_.opposite[corner_next(c)]; }
  uint32_t left_of(uint32_t c) const { return ut_.opposite[corner_prev(c)]; }

  // compute_boundaries — :195-224. The inner walk advances with next(c) inside the face
  // (not through the opposite corner), so in practice each boundary edge opens a new hole id.
  void find_boundaries() {
    hole_of_vertex_.assign(ut_.num_vertices, kNone);
    auto open_hole = [&](uint32_t c0) {
      uint32_t v = vtx(corner_next(c0));
      if (hole_of_vertex_[v] != kNone) return;
      const uint32_t id = (uint32_t)hole_done_.size();
      hole_done_.push_back(0);
      uint32_t c = c0;
      while (hole_of_vertex_[v] == kNone) {
        hole_of_vertex_[v] = id;
        vstate_[v] |= 2;
        c = corner_next(c);
        while (ut_.opposite[c] != kNone) c = corner_next(c);
        v = vtx(corner_next(c));
      }
    };
    if (ut_.has_boundary_list) {  // the corners without an opposite, ascending (K12 by-product): same visiting order
      for (uint32_t c0 : ut_.boundary_corners) open_hole(c0);
    } else {
      for (uint32_t c0 = 0; c0 < ut_.num_corners; ++c0)
        if (ut_.opposite[c0] == kNone) open_hole(c0);
    }
  }

  // process_boundary — :226-256
  void walk_boundary(uint32_t start, bool mark_first) {
    uint32_t c = corner_prev(start);
    for (uint32_t o; (o = ut_.opposite[c]) != kNone;) c = corner_next(o);
    const uint32_t v0 = vtx(start);
    if (mark_first) vstate_[v0] |= 1;
    if (hole_of_vertex_[v0] == kNone) throw Error(DXO_ERR_INTERNAL, "boundary walk from a vertex that is not on a hole");
    hole_done_[hole_of_vertex_[v0]] = 1;
    for (uint32_t v = vtx(corner_prev(c)); v != v0; v = vtx(corner_prev(c))) {
      vstate_[v] |= 1;
      c = corner_next(c);
      for (uint32_t o; (o = ut_.opposite[c]) != kNone;) c = corner_next(o);
    }
  }

  // begin_from — :411-431
  uint32_t start_corner(uint32_t face, bool& interior) const {
    uint32_t c = 3u * face;
    for (int k = 0; k < 3; ++k, c = corner_next(c)) {
      if (ut_.opposite[c] == kNone) { interior = false; return c; }
      if (hole_of_vertex_[vtx(c)] != kNone) {
        for (uint32_t r = c; r != kNone; r = ut_.swing_right(r)) c = r;
        interior = false;
        return corner_prev(c);
      }
    }
    interior = true;
    return c;
  }

  // edgebreaker_from — :261-350
  void traverse(uint32_t c) {
    // Hot loop: every array is reached through a local pointer and the counters live in locals. The byte
    // stores (face / vertex state, symbols) could alias the vectors' own bookkeeping, which would otherwise
    // force the compiler to reload data pointers and sizes from `this` after each of them.
    const uint32_t* const opp = ut_.opposite.data();
    const uint32_t* const cv = ut_.corner_vertex.data();
    uint8_t* const fdone = face_done_.data();
    uint8_t* const vst = vstate_.data();
    uint32_t* vo = visit_order_.data();
    uint8_t* sym = symbols_.data();
    size_t cap = visit_order_.size(), nv = num_visited_;
    uint64_t si = symbol_index_;
    auto split_event = [&](uint8_t right, uint32_t neighbour_corner, uint8_t neighbour_state) {  // check_and_store_topology_split_event — :434-448
      if (neighbour_corner == kNone || !(neighbour_state & 2)) return;  // only faces that were given an S symbol carry a split symbol
      split_events_.push_back({si, (uint64_t)split_symbol_of_face_[neighbour_corner / 3u], right});
    };
    stack_.clear();
    stack_.push_back(c);
    const uint64_t face_budget = ut_.num_faces;
    // The walk is a pointer chase (the next corner comes out of opposite[]), so a cache miss costs its full
    // latency. On regularly indexed meshes the corner index advances by a constant amount every other step;
    // the lines two and three such strides ahead are prefetched. A wrong guess only costs the prefetch.
    const uint32_t num_corners = ut_.num_corners;
    static const uint32_t pf_a = getenv("DXO_PF_A") ? (uint32_t)atoi(getenv("DXO_PF_A")) : 6u, pf_b = getenv("DXO_PF_B") ? (uint32_t)atoi(getenv("DXO_PF_B")) : 12u;
    uint32_t c1 = c, c2 = c;  // corners of the previous two steps
    while (!stack_.empty()) {
      c = stack_.back();
      if (fdone[c / 3u]) { stack_.pop_back(); continue; }
      for (uint64_t n = 0; n < face_budget; ++n) {
        {
          const uint32_t d2 = c - c2;  // wrapping: negative strides work the same way
          const uint32_t a = c + pf_a * d2, b = c + pf_b * d2;
          if (a < num_corners) { __builtin_prefetch(opp + a); __builtin_prefetch(cv + a); }
          if (b < num_corners) { __builtin_prefetch(opp + b); __builtin_prefetch(cv + b); }
          c2 = c1; c1 = c;
        }
        ++si;
        const uint32_t face = c / 3u;
        if (nv >= cap) {  // cannot happen on a consistent table; keep the reference's unbounded growth
          visit_order_.resize(cap + cap / 2 + 16);
          symbols_.resize(visit_order_.size());
          vo = visit_order_.data(); sym = symbols_.data(); cap = visit_order_.size();
        }
        fdone[face] |= 1;
        vo[nv] = c;
        const uint32_t cn = corner_next(c);
        const uint32_t v = cv[c];
        const uint8_t vs = vst[v];
        if (!(vs & 1)) {
          vst[v] = vs | 1;
          if (!(vs & 2)) {
            sym[nv++] = kC;
            c = opp[cn];  // right neighbour
            if (c == kNone) throw Error(DXO_ERR_INTERNAL, "C symbol without a right neighbour");
            continue;
          }
        }
        const uint32_t rc = opp[cn], lc = opp[corner_prev(c)];
        const uint8_t rs = rc == kNone ? 1 : fdone[rc / 3u], ls = lc == kNone ? 1 : fdone[lc / 3u];
        const bool right_done = rs & 1, left_done = ls & 1;
        if (right_done) {
          split_event(1, rc, rs);
          if (left_done) {
            split_event(0, lc, ls);
            sym[nv++] = kE;
            stack_.pop_back();
            break;
          }
          sym[nv++] = kR;
          c = lc;
        } else if (left_done) {
          split_event(0, lc, ls);
          sym[nv++] = kL;
          c = rc;
        } else {
          sym[nv++] = kS;
          ++num_splits_;
          if ((vs & 2) && !hole_done_[hole_of_vertex_[v]]) walk_boundary(c, false);
          split_symbol_of_face_[face] = (uint32_t)si;
          fdone[face] |= 2;
          stack_.back() = lc;
          stack_.push_back(rc);
          break;
        }
      }
    }
    num_visited_ = nv;
    symbol_index_ = si;
  }

  void write_split_events(ByteSink& w) const {  // encode_topology_splits — :375-403
    w.varint(split_events_.size());
    uint64_t last = 0;
    for (const SplitEvent& e : split_events_) {
      w.varint(e.merge_symbol - last);
      w.varint(e.merge_symbol - e.split_symbol);
      last = e.merge_symbol;
    }
    BitPacker bits(w);
    for (const SplitEvent& e : split_events_) bits.put(1, e.right);
    bits.finish();
  }

  // DefaultTraversal::encode — :575-607
  void write_symbols_and_start_faces(ByteSink& w) const {
    {  // CLERS codes, last symbol first, LSB-first bit packing (symbol_encoder.rs:50-58)
      static const uint8_t kBits[5] = {1, 3, 3, 3, 3};
      static const uint8_t kCode[5] = {0b0, 0b001, 0b011, 0b101, 0b111};
      // the LSB-first packer of common.hpp, unrolled onto a raw buffer: a 64-bit accumulator flushed four bytes at a time
      std::vector<uint8_t, NoInitAllocator<uint8_t>> buf((symbols_.size() * 3 + 7) / 8 + 16);
      uint8_t* p = buf.data();
      uint64_t acc = 0;
      unsigned fill = 0;
      const uint8_t* const sym = symbols_.data();
      for (size_t i = symbols_.size(); i-- > 0;) {
        const uint8_t s = sym[i];
        acc |= (uint64_t)kCode[s] << fill;
        fill += kBits[s];
        if (fill >= 32) { const uint32_t lo = (uint32_t)acc; memcpy(p, &lo, 4); p += 4; acc >>= 32; fill -= 32; }  // little-endian host
      }
      while (fill >= 8) { *p++ = (uint8_t)acc; acc >>= 8; fill -= 8; }
      if (fill) *p++ = (uint8_t)acc;
      const size_t nbytes = (size_t)(p - buf.data());
      w.varint(nbytes);
      w.bytes(buf.data(), nbytes);
    }
    {  // start-face configurations (:592-607)
      uint64_t zeros = 0;
      for (uint8_t b : start_face_interior_) zeros += b ? 0 : 1;
      const uint8_t p0 = side_stream_zero_prob(zeros, (float)start_face_interior_.size());
      write_side_stream(start_face_interior_.data(), start_face_interior_.size(), true, p0, w);
    }
  }

 public:
  // phase 2b: the seam stream of one non-position attribute (:610-653); independent per attribute
  void write_seam_stream(const SeamTable& st, ByteSink& w) const {
    // One flag per interior edge. The reference walks the faces in reverse visiting order, lets the first of an edge's
    // two faces met report it (the init face of an interior-start component is never walked, so its neighbours report
    // its edges; corners in the order c, next, prev) and feeds the flags to the coder last-to-first (:610-653). In
    // coding order that is: faces in visiting order, corners prev, next, c, and a face reports an edge exactly when
    // the face across it was done BEFORE this face was visited — which the traversal has already decided:
    //   * the edge opposite to the tip c is the gate the traversal came through (done), or absent for a boundary start;
    //   * the right (next) / left (prev) neighbours were found done precisely for the symbols R, E / L, E (a C or S
    //     face has neither: a C tip was unvisited, so no done face contains it).
    // So the flags come out of one forward pass over (visit order, symbols) without a visited-face table.
    if (!st.has_interior_seam) {
      // only mesh boundaries are seams: every reported flag is 0, and there are (corners - boundary corners) / 2 of them
      size_t boundary = 0;
      if (ut_.has_boundary_list) boundary = ut_.boundary_corners.size();
      else for (uint32_t c = 0; c < ut_.num_corners; ++c) boundary += ut_.opposite[c] == kNone;
      const size_t n = (ut_.num_corners - boundary) / 2;
      const uint8_t p0 = side_stream_zero_prob(n, (float)n);
      std::vector<uint8_t> bytes;
      rabs_encode_zero_run(n, p0, bytes);  // n zero flags: coded by cycle jumps, same bytes as the bit-by-bit coder
      w.u8(p0);
      w.varint(bytes.size());
      w.bytes(bytes);
      return;
    }
    const size_t max_flags = ut_.num_corners / 2 + 1;
    U8Array flags(max_flags + 3);
    uint8_t* const head = flags.data();
    uint8_t* tail = head;
    uint8_t* const end = head + max_flags;
    uint64_t zeros = 0;
    {
      const uint32_t* const opp = ut_.opposite.data();
      const uint8_t* const seam = st.seam.data();
      const uint32_t* const visit = visit_order_.data();
      const uint8_t* const sym = symbols_.data();
      const size_t nv = visit_order_.size();
      uint64_t ones = 0;
      for (size_t i = 0; i < nv; ++i) {
        const uint32_t c = visit[i], cn = corner_next(c), cp = corner_prev(c);
        const uint8_t s = sym[i];
        if (tail > end) throw Error(DXO_ERR_INTERNAL, "seam stream: more flags than interior edges");
        if ((s == kL || s == kE) && opp[cp] != kNone) { const uint8_t f = seam[cp] != 0; *tail++ = f; ones += f; }
        if ((s == kR || s == kE) && opp[cn] != kNone) { const uint8_t f = seam[cn] != 0; *tail++ = f; ones += f; }
        if (opp[c] != kNone) { const uint8_t f = seam[c] != 0; *tail++ = f; ones += f; }
      }
      zeros = (uint64_t)(tail - head) - ones;
    }
    const size_t n = (size_t)(tail - head);
    const uint8_t p0 = side_stream_zero_prob(zeros, (float)n);
    std::vector<uint8_t> bytes;
    rabs_encode_forward(head, n, p0, bytes);
    w.u8(p0);
    w.varint(bytes.size());
    w.bytes(bytes);
  }
};

EdgebreakerEncoder::EdgebreakerEncoder(const UniversalTable& ut) : run_(new EdgebreakerRun(ut)) {}
EdgebreakerEncoder::~EdgebreakerEncoder() { delete run_; }
void EdgebreakerEncoder::traverse() { run_->traverse_all(); }
EdgebreakerEncoder::CornerList EdgebreakerEncoder::corner_list() const { return run_->corner_list(); }
std::vector<uint32_t> EdgebreakerEncoder::corners_of_edgebreaker() const {
  const CornerList l = corner_list();
  std::vector<uint32_t> all(l.init_reversed, l.init_reversed + l.num_init);
  all.insert(all.end(), l.visited, l.visited + l.num_visited);
  return all;
}
void EdgebreakerEncoder::write_head(ByteSink& w, size_t num_seam_tables) const {
  if (num_seam_tables > 255) throw Error(DXO_ERR_TOO_MANY_ATTRIBUTES, "too many connectivity attributes");
  run_->write_head(w, num_seam_tables);
}
void EdgebreakerEncoder::write_seam_stream(const SeamTable& st, ByteSink& w) const { run_->write_seam_stream(st, w); }

std::vector<uint32_t> encode_edgebreaker(const UniversalTable& ut, const std::vector<SeamTable>& seams, ByteSink& w) {
  EdgebreakerEncoder eb(ut);
  eb.traverse();
  eb.write_head(w, seams.size());
  for (const SeamTable& st : seams) eb.write_seam_stream(st, w);
  return eb.corners_of_edgebreaker();
}

// ---------------------------------------------------------------------------------------
// Traverser::compute_seqeunce — shared/attribute/sequence.rs:48-151.
// The reference also deletes stack entries of the face it has just finished
// (:98-131). Such entries can only be popped later, when the face is already marked
// visited, and are dropped at :56-58 before they have any effect — so they are simply
// left on the stack here (lazy deletion, identical output).
U8Array vertex_interior_flags(const TableRef& t) {
  U8Array f(t.num_vertices);
  for (uint32_t v = 0; v < t.num_vertices; ++v) f[v] = t.opp(corner_next(t.left_most[v])) != kNone ? 1 : 0;
  return f;
}

U32Array attribute_sequence(const TableRef& t, const EdgebreakerEncoder::CornerList& corners_of_edgebreaker) {
  U8Array vertex_seen_v(t.num_vertices), face_seen_v(t.num_faces);
  memset(vertex_seen_v.data(), 0, vertex_seen_v.size());
  memset(face_seen_v.data(), 0, face_seen_v.size());
  // The reference's stack starts as a copy of the corner list and is popped from the back. Here the list
  // itself is the (read-only) bottom of the stack, consumed from its end, and only pushed entries are stored.
  // As in the traversal, the hot loop works on local pointers and counters (see EdgebreakerRun::traverse).
  const EdgebreakerEncoder::CornerList bottom_list = corners_of_edgebreaker;
  size_t bottom = bottom_list.size();
  std::vector<uint32_t> stack_v(1024);
  U32Array out_v(t.num_vertices);
  uint32_t* stack = stack_v.data();
  size_t top = 0, stack_cap = stack_v.size();
  uint32_t* const out = out_v.data();
  size_t n_out = 0;
  uint8_t* const vertex_seen = vertex_seen_v.data();
  uint8_t* const face_seen = face_seen_v.data();
  const uint32_t* const cv = t.corner_vertex;
  const uint32_t* const opposite = t.opposite_masked ? t.opposite_masked : t.opposite;  // masked: seams already applied
  const uint8_t* const seam = t.opposite_masked ? nullptr : t.seam;
  const uint8_t* const interior = t.interior;
  const uint32_t* const left_most = t.left_most;
  auto opp = [&](uint32_t c) { return (seam && seam[c]) ? kNone : opposite[c]; };
  auto push = [&](uint32_t c) {
    if (top == stack_cap) { stack_v.resize(stack_cap * 2); stack = stack_v.data(); stack_cap = stack_v.size(); }
    stack[top++] = c;
  };
  auto emit = [&](uint32_t v, uint32_t c) { if (!vertex_seen[v]) { out[n_out++] = c; vertex_seen[v] = 1; } };  // at most one entry per vertex
  const uint32_t num_corners = t.num_corners;
  static const uint32_t pf_a = getenv("DXO_PF_A") ? (uint32_t)atoi(getenv("DXO_PF_A")) : 6u, pf_b = getenv("DXO_PF_B") ? (uint32_t)atoi(getenv("DXO_PF_B")) : 12u;
  uint32_t c1 = 0, c2 = 0;  // corners of the previous two visited faces (stride prefetch, see EdgebreakerRun::traverse)
  uint32_t faces_left = t.num_faces;
  while (top || bottom) {
    // Every face done: whatever is left of the list and the stack names visited faces and would be skipped one by one
    // (the whole visiting order of the traversal sits in the list: 2 M entries for a 1 M-vertex mesh).
    if (faces_left == 0) break;
    const uint32_t c = top ? stack[--top] : bottom_list[--bottom];
    const uint32_t face = c / 3u;
    if (face_seen[face]) continue;
    {
      const uint32_t d2 = c - c2;
      const uint32_t a = c + pf_a * d2, b = c + pf_b * d2;
      if (a < num_corners) { __builtin_prefetch(opposite + a); __builtin_prefetch(cv + a); }
      if (b < num_corners) { __builtin_prefetch(opposite + b); __builtin_prefetch(cv + b); }
      c2 = c1; c1 = c;
    }
    const uint32_t nc = corner_next(c), pc = corner_prev(c);
    const uint32_t v = cv[c], nv = cv[nc], pv = cv[pc];
    if (!vertex_seen[nv] || !vertex_seen[pv]) {  // first face of a component: next, prev, then the tip
      emit(nv, nc);
      emit(pv, pc);
      push(c);
      continue;
    }
    face_seen[face] = 1;
    --faces_left;
    if (!vertex_seen[v]) {
      emit(v, c);
      // is_on_boundary(v): swing_left(left_most_corner(v)) is None (corner_table/mod.rs:36-38)
      const bool is_interior = interior ? interior[v] != 0 : opp(corner_next(left_most[v])) != kNone;
      if (is_interior) {  // interior vertex: keep going to the right
        const uint32_t r = opp(nc);
        if (r == kNone) throw Error(DXO_ERR_INTERNAL, "sequencer: interior vertex without right neighbour");
        push(r);
        continue;
      }
    }
    const uint32_t rc = opp(nc), lc = opp(pc);
    const bool right_seen = rc != kNone && face_seen[rc / 3u];
    const bool left_seen = lc != kNone && face_seen[lc / 3u];
    if (right_seen) {
      if (!left_seen && lc != kNone) push(lc);
    } else if (left_seen) {
      if (rc != kNone) push(rc);
    } else {
      if (lc != kNone) push(lc);
      if (rc != kNone) push(rc);
    }
  }
  out_v.resize(n_out);
  return out_v;
}

}  // namespace dxo
