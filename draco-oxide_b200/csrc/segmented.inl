// Segmented (group-of-meshes) forms of the kernels in kernels.cu — included at the end of that file, inside
// namespace dxo::gpu, so that both forms share the same device bodies. A CTA looks up its tile {segment, first
// element}, loads the segment's descriptor and runs the body over [first, first + kSegTile) with the CTA as stride.
// Descriptors carry pointers, so every index stays local to its mesh and the bytes produced are those of the
// per-mesh launches. See kernels.cuh for the descriptor structs.

namespace {
__device__ __forceinline__ Tile load_tile(const Tile* __restrict__ tiles) {
  const uint2 v = __ldg(reinterpret_cast<const uint2*>(tiles) + blockIdx.x);
  return Tile{v.x, v.y};
}
inline uint32_t bits_for(uint64_t v) { uint32_t b = 1; while ((v >> b) != 0) ++b; return b; }
}  // namespace

// ---------------------------------------------------------------------------------------
// K12 / K13 over a group. Half-edge keys are (mesh, min vertex, max vertex) packed into the low
// mesh_bits + 2 * vertex_bits bits, values are group-wide corner slots; one radix sort over only those bits
// pairs the half edges of every mesh at once (CornerTable::compute_table, corner_table/mod.rs:252-340).
__global__ void __launch_bounds__(kThreads) seg_corner_vertex_kernel(const MeshSeg* __restrict__ segs, const Tile* __restrict__ tiles) {
  const Tile tl = load_tile(tiles);
  const MeshSeg& m = segs[tl.seg];
  const uint32_t end = min(tl.first + kSegTile, m.num_corners);
  uint32_t bad = 0;
  for (uint32_t c = tl.first + threadIdx.x; c < end; c += kThreads) {
    const uint32_t p = __ldcs(m.faces + c);
    uint32_t v = 0;
    if (p >= m.num_points) bad = kMeshBadIndex;
    else {
      v = m.pos_map ? __ldg(m.pos_map + p) : p;
      if (v >= m.num_vertices) { bad = kMeshBadIndex; v = 0; }
    }
    m.cv[c] = v;
  }
  if (bad) atomicOr(m.flags, bad);
}

__global__ void __launch_bounds__(kThreads) seg_halfedge_keys_kernel(const MeshSeg* __restrict__ segs, const Tile* __restrict__ tiles, uint32_t vertex_bits,
                                                                     unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals) {
  const Tile tl = load_tile(tiles);
  const MeshSeg& m = segs[tl.seg];
  const uint32_t end = min(tl.first + kSegTile, m.num_corners);
  const unsigned long long mesh_key = (unsigned long long)tl.seg << (2u * vertex_bits);
  uint32_t bad = 0;
  for (uint32_t c = tl.first + threadIdx.x; c < end; c += kThreads) {
    const uint32_t tip = m.cv[c], src = m.cv[cnext(c)], snk = m.cv[cprev(c)];
    if (tip == src || tip == snk || src == snk) bad = kMeshNotExact;  // degenerate face
    const uint32_t a = min(src, snk), b = max(src, snk);
    keys[(size_t)m.corner_base + c] = mesh_key | ((unsigned long long)a << vertex_bits) | b;
    vals[(size_t)m.corner_base + c] = m.corner_base + c;
  }
  if (bad) atomicOr(m.flags, bad);
}

__global__ void __launch_bounds__(kThreads) seg_halfedge_pair_kernel(const MeshSeg* __restrict__ segs, unsigned long long total, uint32_t vertex_bits,
                                                                     const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ vals) {
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const unsigned long long k = keys[i];
    const MeshSeg& m = segs[(uint32_t)(k >> (2u * vertex_bits))];
    const bool same_prev = i > 0 && keys[i - 1] == k;
    const bool same_next = i + 1 < total && keys[i + 1] == k;
    const uint32_t c = vals[i] - m.corner_base;
    if (same_prev && same_next) { atomicOr(m.flags, kMeshNotExact); continue; }  // 3+ half edges on one edge
    if (!same_prev && !same_next) { m.opposite[c] = kNoneDev; continue; }
    const uint32_t other = (same_next ? vals[i + 1] : vals[i - 1]) - m.corner_base;
    if (same_next && i + 2 < total && keys[i + 2] == k) { atomicOr(m.flags, kMeshNotExact); continue; }
    if (same_prev && i >= 2 && keys[i - 2] == k) { atomicOr(m.flags, kMeshNotExact); continue; }
    const uint32_t src = m.cv[cnext(c)], osrc = m.cv[cnext(other)];
    if (src == osrc) { atomicOr(m.flags, kMeshNotExact); continue; }         // same direction: inconsistent orientation
    if (m.cv[c] == m.cv[other]) { atomicOr(m.flags, kMeshNotExact); continue; }  // equal tips are skipped by the reference (:308-310)
    m.opposite[c] = other;
  }
}

__global__ void __launch_bounds__(kThreads) seg_first_corner_kernel(const MeshSeg* __restrict__ segs, const Tile* __restrict__ tiles) {
  const Tile tl = load_tile(tiles);
  const MeshSeg& m = segs[tl.seg];
  const uint32_t end = min(tl.first + kSegTile, m.num_corners);
  for (uint32_t c = tl.first + threadIdx.x; c < end; c += kThreads) {
    const uint32_t v = m.cv[c];
    atomicMin(m.first_corner + v, c);
    atomicAdd(m.valence + v, 1u);
  }
}

// left-most corners (compute_left_most_corners, corner_table/mod.rs:342-416, single-fan vertices) + interior flags
__global__ void __launch_bounds__(kThreads) seg_left_most_kernel(const MeshSeg* __restrict__ segs, const Tile* __restrict__ tiles) {
  const Tile tl = load_tile(tiles);
  const MeshSeg& m = segs[tl.seg];
  const uint32_t end = min(tl.first + kSegTile, m.num_vertices);
  const uint32_t* __restrict__ opposite = m.opposite;
  uint32_t bad = 0;
  for (uint32_t v = tl.first + threadIdx.x; v < end; v += kThreads) {
    const uint32_t n = m.valence[v], c = m.first_corner[v];
    if (n == 0) { bad |= kMeshUnusedVertex; m.left_most[v] = kNoneDev; m.interior[v] = 0; continue; }
    uint32_t last = c, count = 1, a = c;
    bool open = false;
    for (;;) {  // swing left: opposite(next(a)) -> next
      const uint32_t o = opposite[cnext(a)];
      if (o == kNoneDev) { open = true; break; }
      if (o >= m.num_corners) { bad |= kMeshNotExact; break; }  // only on a mesh K12 has flagged (entries left at 0xFF are kNoneDev)
      a = cnext(o);
      if (a == c || count > n) break;
      last = a;
      ++count;
    }
    if (open) {  // the corners to the right of the start belong to the fan as well
      a = c;
      for (;;) {
        const uint32_t o = opposite[cprev(a)];
        if (o == kNoneDev || count > n) break;
        a = cprev(o);
        ++count;
      }
    }
    if (count != n) bad |= kMeshSplitVertex;
    m.left_most[v] = last;
    m.interior[v] = opposite[cnext(last)] != kNoneDev ? 1 : 0;  // !is_on_boundary(v) (corner_table/mod.rs:36-38)
  }
  if (bad) atomicOr(m.flags, bad);
}

size_t seg_corner_tables_scratch_bytes(uint64_t total_corners) {
  size_t cub_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (const uint32_t*)nullptr,
                                  (uint32_t*)nullptr, (int)total_corners);
  const size_t a = ((total_corners * 8 + 255) / 256) * 256, b = ((total_corners * 4 + 255) / 256) * 256;
  return 2 * a + 2 * b + cub_bytes + 256;
}

void launch_seg_corner_tables(const MeshSeg* segs, uint32_t num_meshes, const Tile* corner_tiles, uint32_t num_corner_tiles, const Tile* vertex_tiles,
                              uint32_t num_vertex_tiles, uint64_t total_corners, uint32_t vertex_bits, void* scratch, size_t scratch_bytes, cudaStream_t s) {
  if (!num_corner_tiles) return;
  const size_t a = ((total_corners * 8 + 255) / 256) * 256, b = ((total_corners * 4 + 255) / 256) * 256;
  uint8_t* p = (uint8_t*)scratch;
  unsigned long long* keys_in = (unsigned long long*)p; p += a;
  unsigned long long* keys_out = (unsigned long long*)p; p += a;
  uint32_t* vals_in = (uint32_t*)p; p += b;
  uint32_t* vals_out = (uint32_t*)p; p += b;
  size_t cub_bytes = scratch_bytes - (2 * a + 2 * b);
  const int end_bit = (int)std::min<uint32_t>(64u, bits_for(num_meshes ? num_meshes - 1 : 0) + 2u * vertex_bits);
  seg_corner_vertex_kernel<<<num_corner_tiles, kThreads, 0, s>>>(segs, corner_tiles);
  seg_halfedge_keys_kernel<<<num_corner_tiles, kThreads, 0, s>>>(segs, corner_tiles, vertex_bits, keys_in, vals_in);
  cub::DeviceRadixSort::SortPairs(p, cub_bytes, keys_in, keys_out, vals_in, vals_out, (int)total_corners, 0, end_bit, s);
  seg_halfedge_pair_kernel<<<grid_for(total_corners), kThreads, 0, s>>>(segs, total_corners, vertex_bits, keys_out, vals_out);
  seg_first_corner_kernel<<<num_corner_tiles, kThreads, 0, s>>>(segs, corner_tiles);
  if (num_vertex_tiles) seg_left_most_kernel<<<num_vertex_tiles, kThreads, 0, s>>>(segs, vertex_tiles);
}

// ---------------------------------------------------------------------------------------
// K14 over a group (AttributeCornerTable::new + recompute_vertices, attribute_corner_table.rs:16-137).
constexpr uint32_t kSeamCapacity = 8u;  // more attribute vertices than reserved slots

__global__ void __launch_bounds__(kThreads) seg_seam_flags_kernel(const SeamSeg* __restrict__ segs, const Tile* __restrict__ tiles) {
  const Tile tl = load_tile(tiles);
  const SeamSeg& g = segs[tl.seg];
  if (*g.mesh_flags) return;
  const uint32_t end = min(tl.first + kSegTile, g.num_corners);
  uint32_t bad = 0;
  auto val = [&](uint32_t corner) -> uint32_t {
    const uint32_t p = __ldg(g.faces + corner);
    if (p >= g.num_points) { bad |= kSeamBadPoint; return 0u; }
    return g.map ? __ldg(g.map + p) : p;
  };
  for (uint32_t c = tl.first + threadIdx.x; c < end; c += kThreads) {
    const uint32_t o = g.opposite[c];
    const uint32_t cn = cnext(c), cp = cprev(c);
    bool is_seam;
    if (o == kNoneDev) is_seam = true;  // mesh boundary counts as a seam
    else {
      const uint32_t on = cnext(o), op = cprev(o);
      is_seam = val(cn) != val(op) || val(cp) != val(on);
    }
    g.seam[c] = is_seam ? 1 : 0;
    if (is_seam) { g.vertex_on_seam[g.cv_u[cn]] = 1; g.vertex_on_seam[g.cv_u[cp]] = 1; }
    if (is_seam && o != kNoneDev) bad |= kSeamInterior;
  }
  if (bad) atomicOr(g.scalars + 1, bad);
}

template <bool ASSIGN>
__global__ void __launch_bounds__(kThreads) seg_seam_vertices_kernel(const SeamSeg* __restrict__ segs, const Tile* __restrict__ tiles,
                                                                     uint32_t* __restrict__ counts, const uint32_t* __restrict__ bases) {
  const Tile tl = load_tile(tiles);
  const SeamSeg& g = segs[tl.seg];
  if (*g.mesh_flags) return;
  const uint32_t end = min(tl.first + kSegTile, g.num_vertices_u);
  uint32_t bad = 0;
  const uint32_t seg_base = ASSIGN ? bases[g.count_base] : 0u;
  for (uint32_t v = tl.first + threadIdx.x; v < end; v += kThreads) {
    const uint32_t first = seam_walk_start(v, g.left_most_u, g.opposite, g.seam, g.vertex_on_seam, g.num_corners, bad);
    uint32_t id = ASSIGN ? bases[g.count_base + v] - seg_base : 0u, n = 1;
    const uint32_t id0 = id;
    bool fits = true;
    if (ASSIGN) {
      fits = id < g.capacity;
      if (fits) { g.cv_a[first] = id; g.left_most_a[id] = first; }
    }
    uint32_t s = first, guard = g.num_corners;
    for (;;) {  // universal swing right
      const uint32_t o = g.opposite[cprev(s)];
      if (o == kNoneDev) break;
      s = cprev(o);
      if (s == first || --guard == 0) break;
      if (g.seam[cnext(s)]) {  // crossing a seam starts a new attribute vertex
        ++n;
        if (ASSIGN) { ++id; fits = id < g.capacity; if (fits) g.left_most_a[id] = s; }
      }
      if (ASSIGN && fits) g.cv_a[s] = id;
    }
    if (!ASSIGN) counts[g.count_base + v] = n;
    else {
      if (!fits || id0 >= g.capacity) bad |= kSeamCapacity;
      if (v + 1 == g.num_vertices_u) g.scalars[0] = id0 + n;
    }
  }
  if (bad) atomicOr(g.scalars + 1, bad);
}

// per attribute vertex: swing_left(left_most_a[v]) exists in the attribute's table (seam edges have no opposite)
__global__ void __launch_bounds__(kThreads) seg_seam_interior_kernel(const SeamSeg* __restrict__ segs, const Tile* __restrict__ tiles) {
  const Tile tl = load_tile(tiles);
  const SeamSeg& g = segs[tl.seg];
  if (*g.mesh_flags || (g.scalars[1] & (kSeamBadPoint | kSeamClosedFan | kSeamCapacity))) return;
  const uint32_t end = min(min(tl.first + kSegTile, g.scalars[0]), g.capacity);
  for (uint32_t v = tl.first + threadIdx.x; v < end; v += kThreads) {
    const uint32_t e = cnext(g.left_most_a[v]);
    g.interior_a[v] = (!g.seam[e] && g.opposite[e] != kNoneDev) ? 1 : 0;
  }
}

size_t seg_seam_tables_scratch_bytes(uint64_t total_count_slots) {
  size_t cub_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)total_count_slots);
  return cub_bytes + 256;
}

void launch_seg_seam_tables(const SeamSeg* segs, uint32_t num_segs, const Tile* corner_tiles, uint32_t num_corner_tiles, const Tile* vertex_tiles,
                            uint32_t num_vertex_tiles, const Tile* attr_vertex_tiles, uint32_t num_attr_vertex_tiles, uint32_t* counts, uint32_t* bases,
                            uint64_t total_count_slots, void* scratch, size_t scratch_bytes, cudaStream_t s) {
  if (!num_segs || !num_corner_tiles || !num_vertex_tiles) return;
  // counts of skipped (flagged) meshes stay as they were: the caller zeroes `counts`
  seg_seam_flags_kernel<<<num_corner_tiles, kThreads, 0, s>>>(segs, corner_tiles);
  seg_seam_vertices_kernel<false><<<num_vertex_tiles, kThreads, 0, s>>>(segs, vertex_tiles, counts, bases);
  cub::DeviceScan::ExclusiveSum(scratch, scratch_bytes, counts, bases, (int)total_count_slots, s);
  seg_seam_vertices_kernel<true><<<num_vertex_tiles, kThreads, 0, s>>>(segs, vertex_tiles, counts, bases);
  if (num_attr_vertex_tiles) seg_seam_interior_kernel<<<num_attr_vertex_tiles, kThreads, 0, s>>>(segs, attr_vertex_tiles);
}

// ---------------------------------------------------------------------------------------
// K1-K9 over a group
__global__ void __launch_bounds__(kThreads) seg_init_stats_kernel(const AttrSeg* __restrict__ segs, uint32_t num_segs) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= num_segs) return;
  AttrStats* st = segs[k].stats;
  for (int i = 0; i < 4; ++i) { st->vmin_bits[i] = 0; st->vmax_bits[i] = 0; }
  st->range = 0.0f;
  st->wrap_min = 0x7FFFFFFF;
  st->wrap_max = (int32_t)0x80000000;
  st->nonzero_symbols = 0; st->max_symbol = 0; st->error_flags = 0;
  st->bit_length = 0; st->precision = 0; st->num_table_symbols = 0; st->table_bytes = 0; st->payload_bytes = 0;
  st->pad[0] = st->pad[1] = st->pad[2] = 0;
  if (segs[k].side_stats) *segs[k].side_stats = SideStats{0, 0, 0, 0};
}
void launch_seg_init_stats(const AttrSeg* segs, uint32_t num_segs, cudaStream_t s) {
  if (num_segs) seg_init_stats_kernel<<<(num_segs + kThreads - 1) / kThreads, kThreads, 0, s>>>(segs, num_segs);
}

__global__ void __launch_bounds__(kThreads) seg_pad3_kernel(const Pad3Seg* __restrict__ segs, const Tile* __restrict__ tiles) {
  const Tile tl = load_tile(tiles);
  const Pad3Seg g = segs[tl.seg];
  const uint32_t end = min(tl.first + kSegTile, g.n);
  for (uint32_t i = tl.first + threadIdx.x; i < end; i += kThreads)
    g.out[i] = make_uint4(__ldcs(g.in + 3 * (size_t)i), __ldcs(g.in + 3 * (size_t)i + 1), __ldcs(g.in + 3 * (size_t)i + 2), 0u);
}
void launch_seg_pad3(const Pad3Seg* segs, const Tile* tiles, uint32_t num_tiles, cudaStream_t s) {
  if (num_tiles) seg_pad3_kernel<<<num_tiles, kThreads, 0, s>>>(segs, tiles);
}

__global__ void __launch_bounds__(kThreads) seg_fan_link_kernel(const AttrSeg* __restrict__ segs, const Tile* __restrict__ tiles) {
  const Tile tl = load_tile(tiles);
  const AttrSeg& g = segs[tl.seg];
  const uint32_t end = min(tl.first + kSegTile, g.t.num_corners);
  for (uint32_t c = tl.first + threadIdx.x; c < end; c += kThreads) {
    const uint32_t o = g.seam_for_links[c] ? kNoneDev : g.t.opposite[c];
    g.fan_link_out[c] = make_uint2(o, o == kNoneDev ? 0u : __ldg(g.t.corner_point + o));
  }
}
void launch_seg_fan_links(const AttrSeg* segs, const Tile* tiles, uint32_t num_tiles, cudaStream_t s) {
  if (num_tiles) seg_fan_link_kernel<<<num_tiles, kThreads, 0, s>>>(segs, tiles);
}

template <int N>
__global__ void __launch_bounds__(kThreads) seg_minmax_kernel(const AttrSeg* __restrict__ segs, const Tile* __restrict__ tiles) {
  const Tile tl = load_tile(tiles);
  const AttrSeg& g = segs[tl.seg];
  minmax_body<N>(g.values, g.stats, tl.first + threadIdx.x, min(tl.first + kSegTile, g.num_unique), kThreads);
}
void launch_seg_minmax(const AttrSeg* segs, const Tile* tiles, uint32_t num_tiles, uint32_t ncomp, cudaStream_t s) {
  if (!num_tiles) return;
  switch (ncomp) {
    case 1: seg_minmax_kernel<1><<<num_tiles, kThreads, 0, s>>>(segs, tiles); break;
    case 2: seg_minmax_kernel<2><<<num_tiles, kThreads, 0, s>>>(segs, tiles); break;
    case 3: seg_minmax_kernel<3><<<num_tiles, kThreads, 0, s>>>(segs, tiles); break;
    default: seg_minmax_kernel<4><<<num_tiles, kThreads, 0, s>>>(segs, tiles); break;
  }
}

template <int N>
__global__ void __launch_bounds__(kThreads) seg_quantize_kernel(const AttrSeg* __restrict__ segs, const Tile* __restrict__ tiles) {
  const Tile tl = load_tile(tiles);
  const AttrSeg& g = segs[tl.seg];
  quantize_body<N>(g.values, g.bits, const_cast<int32_t*>(g.q.values), g.stats, tl.first == 0, tl.first + threadIdx.x,
                   min(tl.first + kSegTile, g.num_unique), kThreads);
}
void launch_seg_quantize(const AttrSeg* segs, const Tile* tiles, uint32_t num_tiles, uint32_t ncomp, cudaStream_t s) {
  if (!num_tiles) return;
  switch (ncomp) {
    case 1: seg_quantize_kernel<1><<<num_tiles, kThreads, 0, s>>>(segs, tiles); break;
    case 2: seg_quantize_kernel<2><<<num_tiles, kThreads, 0, s>>>(segs, tiles); break;
    case 3: seg_quantize_kernel<3><<<num_tiles, kThreads, 0, s>>>(segs, tiles); break;
    default: seg_quantize_kernel<4><<<num_tiles, kThreads, 0, s>>>(segs, tiles); break;
  }
}

__global__ void __launch_bounds__(kThreads) seg_oct_quantize_kernel(const AttrSeg* __restrict__ segs, const Tile* __restrict__ tiles) {
  const Tile tl = load_tile(tiles);
  const AttrSeg& g = segs[tl.seg];
  oct_quantize_body(g.values, const_cast<int32_t*>(g.q.values), g.stats, tl.first + threadIdx.x, min(tl.first + kSegTile, g.num_unique), kThreads);
}
void launch_seg_oct_quantize(const AttrSeg* segs, const Tile* tiles, uint32_t num_tiles, cudaStream_t s) {
  if (num_tiles) seg_oct_quantize_kernel<<<num_tiles, kThreads, 0, s>>>(segs, tiles);
}

__global__ void __launch_bounds__(kThreads) seg_seq_prepare_kernel(const AttrSeg* __restrict__ segs, const Tile* __restrict__ tiles) {
  const Tile tl = load_tile(tiles);
  const AttrSeg& g = segs[tl.seg];
  seq_prepare_body(g.seq, g.t, g.q, g.rank, (int)g.wrapped, g.stats, tl.first + threadIdx.x, min(tl.first + kSegTile, g.n), kThreads);
}
void launch_seg_seq_prepare(const AttrSeg* segs, const Tile* tiles, uint32_t num_tiles, cudaStream_t s) {
  if (num_tiles) seg_seq_prepare_kernel<<<num_tiles, kThreads, 0, s>>>(segs, tiles);
}

template <int N>
__global__ void __launch_bounds__(kThreads) seg_predict_parallelogram_kernel(const AttrSeg* __restrict__ segs, const Tile* __restrict__ tiles) {
  const Tile tl = load_tile(tiles);
  const AttrSeg& g = segs[tl.seg];
  predict_parallelogram_body<N>(g.seq, g.t, g.q, g.rank, g.symbols, g.stats, tl.first + threadIdx.x, min(tl.first + kSegTile, g.n), kThreads);
}
template <int N>
__global__ void __launch_bounds__(kThreads) seg_predict_delta_kernel(const AttrSeg* __restrict__ segs, const Tile* __restrict__ tiles) {
  const Tile tl = load_tile(tiles);
  const AttrSeg& g = segs[tl.seg];
  predict_delta_body<N>(g.seq, g.t, g.q, g.symbols, g.stats, tl.first + threadIdx.x, min(tl.first + kSegTile, g.n), kThreads);
}
__global__ void __launch_bounds__(kThreads, 8) seg_predict_normal_kernel(const AttrSeg* __restrict__ segs, const Tile* __restrict__ tiles) {
  const Tile tl = load_tile(tiles);
  const AttrSeg& g = segs[tl.seg];
  predict_normal_body(g.seq, g.t, g.q, g.pos, g.symbols, g.side_flags, g.stats, tl.first + threadIdx.x, min(tl.first + kSegTile, g.n), kThreads);
}
__global__ void __launch_bounds__(kThreads, 4) seg_predict_texcoord_kernel(const AttrSeg* __restrict__ segs, const Tile* __restrict__ tiles) {
  const Tile tl = load_tile(tiles);
  const AttrSeg& g = segs[tl.seg];
  predict_texcoord_body(g.seq, g.t, g.q, g.pos, g.pos_num_points, g.rank, g.symbols, g.side_flags, g.stats, tl.first + threadIdx.x,
                        min(tl.first + kSegTile, g.n), kThreads);
}
// scheme: the values of encoder.hpp's Scheme enum (prediction_scheme/mod.rs:74-86)
void launch_seg_predict(const AttrSeg* segs, const Tile* tiles, uint32_t num_tiles, uint32_t scheme, uint32_t ncomp, cudaStream_t s) {
  if (!num_tiles) return;
  if (scheme == 6) { seg_predict_normal_kernel<<<num_tiles, kThreads, 0, s>>>(segs, tiles); return; }
  if (scheme == 5) { seg_predict_texcoord_kernel<<<num_tiles, kThreads, 0, s>>>(segs, tiles); return; }
  if (scheme == 1) {
    switch (ncomp) {
      case 1: seg_predict_parallelogram_kernel<1><<<num_tiles, kThreads, 0, s>>>(segs, tiles); break;
      case 2: seg_predict_parallelogram_kernel<2><<<num_tiles, kThreads, 0, s>>>(segs, tiles); break;
      case 3: seg_predict_parallelogram_kernel<3><<<num_tiles, kThreads, 0, s>>>(segs, tiles); break;
      default: seg_predict_parallelogram_kernel<4><<<num_tiles, kThreads, 0, s>>>(segs, tiles); break;
    }
    return;
  }
  switch (ncomp) {
    case 1: seg_predict_delta_kernel<1><<<num_tiles, kThreads, 0, s>>>(segs, tiles); break;
    case 2: seg_predict_delta_kernel<2><<<num_tiles, kThreads, 0, s>>>(segs, tiles); break;
    case 3: seg_predict_delta_kernel<3><<<num_tiles, kThreads, 0, s>>>(segs, tiles); break;
    default: seg_predict_delta_kernel<4><<<num_tiles, kThreads, 0, s>>>(segs, tiles); break;
  }
}

constexpr uint32_t kSegHistTile = 8 * kSegTile;  // symbols per CTA: amortises zeroing / flushing the shared-memory bins
__global__ void __launch_bounds__(kThreads) seg_histogram_smem_kernel(const AttrSeg* __restrict__ segs, const Tile* __restrict__ tiles) {
  const Tile tl = load_tile(tiles);
  const AttrSeg& g = segs[tl.seg];
  histogram_smem_body(g.symbols, g.hist, g.hist_capacity, g.stats, (uint64_t)tl.first + threadIdx.x, min(tl.first + kSegHistTile, g.num_symbols), kThreads);
}
__global__ void __launch_bounds__(kThreads) seg_histogram_global_kernel(const AttrSeg* __restrict__ segs, const Tile* __restrict__ tiles) {
  const Tile tl = load_tile(tiles);
  const AttrSeg& g = segs[tl.seg];
  histogram_global_body(g.symbols, g.hist, g.hist_capacity, g.stats, (uint64_t)tl.first + threadIdx.x, min(tl.first + kSegHistTile, g.num_symbols), kThreads);
}
// one tile of kSegHistTile symbols per CTA, brought in by a single bulk copy (see histogram_tma_body)
__global__ void __launch_bounds__(kThreads) seg_histogram_tma_kernel(const AttrSeg* __restrict__ segs, const Tile* __restrict__ tiles) {
  const Tile tl = load_tile(tiles);
  const AttrSeg& g = segs[tl.seg];
  const uint32_t t = tl.first / kSegHistTile;
  histogram_tma_body(g.symbols, g.num_symbols, g.hist, g.hist_capacity, g.stats, t, 1u << 30, ((uint64_t)g.num_symbols + kSegHistTile - 1) / kSegHistTile,
                     kSegHistTile, 1, (uint64_t)tl.first + kSegHistTile >= g.num_symbols);
}
void launch_seg_histogram(const AttrSeg* segs, const Tile* tiles, uint32_t num_tiles, bool smem, cudaStream_t s) {
  if (!num_tiles) return;
  static_assert(kSegHistTile * 4 <= kHistStageBytes, "a segmented histogram tile must fit the staging area");
  if (smem && hist_use_tma()) {  // the arena slots of the symbol streams are 256-byte aligned
    allow_hist_smem((const void*)seg_histogram_tma_kernel);
    seg_histogram_tma_kernel<<<num_tiles, kThreads, kHistTmaSmem, s>>>(segs, tiles);
  } else if (smem) seg_histogram_smem_kernel<<<num_tiles, kThreads, 0, s>>>(segs, tiles);
  else seg_histogram_global_kernel<<<num_tiles, kThreads, 0, s>>>(segs, tiles);
}

__global__ void __launch_bounds__(kTableThreads) seg_build_table_kernel(const AttrSeg* __restrict__ segs) {
  const AttrSeg& g = segs[blockIdx.x];
  build_table_body(g.hist, g.hist_capacity, g.num_symbols, g.work, g.rans_table, g.table_bytes, g.table_capacity, g.stats);
}
void launch_seg_build_tables(const AttrSeg* segs, uint32_t num_segs, cudaStream_t s) {
  if (num_segs) seg_build_table_kernel<<<num_segs, kTableThreads, 0, s>>>(segs);
}

// ---------------------------------------------------------------------------------------
// Binary side streams on the device (RabsCoder, encode/entropy/rans.rs:71-127), one warp per stream:
//   normals   — the flips in sequence order, zero_prob from the zero count (mesh_normal_prediction.rs:147-163);
//   texcoords — the orientation values that exist, in order; zero_prob from the forward transitions (the scan starts from
//               `true`) over len + 0.001; coded bits are the backward deltas o[k] == o[k+1], o[len] = true
//               (mesh_prediction_for_texture_coordinates.rs:221-260).
// The coder is a serial chain of ~30 cycles per bit (it does not forget its state the way the multi-symbol coder does,
// DESIGN.md), which is why the single-mesh path codes its one or two long streams on host threads; across a group there
// are hundreds of short streams and a warp each keeps them off the host altogether. Lanes load 32 flags at a time and
// every lane runs the same chain on the ballot masks; lane 0 writes the bytes.
__device__ __forceinline__ uint32_t device_zero_prob(uint32_t zeros, float len) {  // side_stream_zero_prob (common.hpp), f32 steps kept apart
  const float ratio = (float)zeros / len;
  const float scaled = ratio * 256.0f;
  const float p = scaled + 0.5f;
  uint32_t v;
  if (!(p == p) || p <= 0.0f) v = 0;
  else if (p >= 65535.0f) v = 65535;
  else v = (uint32_t)p;
  return min(max(v, 1u), 255u);
}

struct RabsDev {
  uint32_t x, pos;
  // per bit value, in registers (selected, never indexed): renormalisation threshold f << 12, g = 256 - f, cum,
  // m = ceil(2^32 / f) (f = 1 has m = 2^32: its quotient is x itself, flagged by one[])
  uint32_t thr0, thr1, g0, g1, cum0, m0, m1, one0, one1;
  uint8_t* out; uint32_t cap; bool lane0; bool overflow;
  __device__ __forceinline__ void init(uint32_t zero_prob, uint8_t* o, uint32_t capacity, bool l0) {
    const uint32_t f0 = zero_prob, f1 = 256u - zero_prob;
    thr0 = f0 << 12; thr1 = f1 << 12;
    g0 = 256u - f0; g1 = 256u - f1;
    cum0 = f1;
    // q = floor(x / f) = umulhi(x, ceil(2^32 / f)) for x < 2^20, 2 <= f <= 255 (common.hpp rabs_encode_forward_fn)
    one0 = f0 == 1u; one1 = f1 == 1u;
    m0 = one0 ? 0u : (uint32_t)(((1ull << 32) + f0 - 1) / f0); m1 = one1 ? 0u : (uint32_t)(((1ull << 32) + f1 - 1) / f1);
    x = 4096u; pos = 0; out = o; cap = capacity; lane0 = l0; overflow = false;
  }
  __device__ __forceinline__ void put(uint32_t b) {
    const uint32_t thr = b ? thr1 : thr0, g = b ? g1 : g0, cum = b ? 0u : cum0, m = b ? m1 : m0, one = b ? one1 : one0;
    if (x >= thr) {
      if (pos < cap) { if (lane0) out[pos] = (uint8_t)x; } else overflow = true;
      ++pos; x >>= 8;
    }
    const uint32_t q = one ? x : __umulhi(x, m);
    x = x + q * g + cum;
  }
  __device__ __forceinline__ void finish() {  // ans_write_tail
    const uint32_t t = x - 4096u;
    const uint32_t nb = t < (1u << 6) ? 1u : t < (1u << 14) ? 2u : t < (1u << 22) ? 3u : 4u;
    const uint32_t v = t + (nb == 1 ? 0u : nb == 2 ? 0x4000u : nb == 3 ? 0x800000u : 0xC0000000u);
    if (pos + nb > cap) overflow = true;
    else if (lane0) for (uint32_t k = 0; k < nb; ++k) out[pos + k] = (uint8_t)(v >> (8 * k));
    pos += nb;
  }
};

__global__ void __launch_bounds__(32) seg_side_stream_kernel(const AttrSeg* __restrict__ segs, const uint32_t* __restrict__ seg_ids) {
  const AttrSeg& g = segs[seg_ids[blockIdx.x]];
  if (g.stats->error_flags) return;
  const uint32_t lane = threadIdx.x, n = g.n;
  const uint8_t* __restrict__ flags = g.side_flags;
  const bool normal = g.scheme == 6;
  // pass 1: entries and ones (normals) / transitions (texcoords)
  uint32_t entries = 0, tally = 0, last_true = 1;  // the transition scan starts from `true`
  for (uint32_t base = 0; base < n; base += 32) {
    const uint32_t f = base + lane < n ? flags[base + lane] : 0u;
    const uint32_t nz = __ballot_sync(0xFFFFFFFFu, f != 0), tr = __ballot_sync(0xFFFFFFFFu, f == 2);
    if (normal) { tally += __popc(nz); continue; }
    entries += __popc(nz);
    // value of the previous existing entry, for every lane that holds one
    const uint32_t below = nz & ((1u << lane) - 1u);
    const uint32_t prev_true = below ? ((tr >> (31 - __clz(below))) & 1u) : last_true;
    const uint32_t mine = (tr >> lane) & 1u;
    tally += __popc(__ballot_sync(0xFFFFFFFFu, (f != 0) && mine != prev_true));
    if (nz) last_true = (tr >> (31 - __clz(nz))) & 1u;
  }
  uint32_t zero_prob;
  if (normal) { entries = n; zero_prob = device_zero_prob(n - tally, (float)n); }
  else zero_prob = device_zero_prob(tally, (float)entries + 0.001f);
  RabsDev coder;
  coder.init(zero_prob, g.side_payload, g.side_capacity, lane == 0);
  // pass 2: the chain. Texcoords: every existing entry after the first emits (previous == this); the end emits (last == true).
  bool have_prev = false;
  uint32_t prev = 0;
  uint32_t f_next = lane < n ? flags[lane] : 0u;
  for (uint32_t base = 0; base < n; base += 32) {
    const uint32_t f = f_next;
    f_next = base + 32 + lane < n ? flags[base + 32 + lane] : 0u;
    uint32_t nz = __ballot_sync(0xFFFFFFFFu, f != 0);
    const uint32_t tr = __ballot_sync(0xFFFFFFFFu, f == 2);
    if (normal) {
      const uint32_t cnt = min(32u, n - base);
      for (uint32_t l = 0; l < cnt; ++l) coder.put((nz >> l) & 1u);
    } else {
      while (nz) {
        const uint32_t l = __ffs(nz) - 1;
        nz &= nz - 1;
        const uint32_t cur = (tr >> l) & 1u;
        if (have_prev) coder.put(prev == cur ? 1u : 0u);
        prev = cur; have_prev = true;
      }
    }
  }
  if (!normal && have_prev) coder.put(prev == 1u ? 1u : 0u);
  coder.finish();
  if (lane == 0) {
    if (coder.overflow) atomicOr(&g.stats->error_flags, kErrAlphabet);
    *g.side_stats = SideStats{entries, zero_prob, coder.pos, 0u};
  }
}
void launch_seg_side_streams(const AttrSeg* segs, const uint32_t* seg_ids, uint32_t num_side_streams, cudaStream_t s) {
  if (num_side_streams) seg_side_stream_kernel<<<num_side_streams, 32, 0, s>>>(segs, seg_ids);
}

// ---------------------------------------------------------------------------------------
// K10 over a group: the five passes of launch_rans_encode, each as one launch over the chunks / pieces of every stream.
namespace {
__device__ __forceinline__ RansChunkState chunk_state_of(const RansJob& j) {
  RansChunkState cs;
  cs.start = j.start; cs.exit = j.exit; cs.nbytes = j.nbytes; cs.chain_start = j.chain_start;
  cs.cand_start = j.cand_start; cs.cand_exit = j.cand_exit; cs.cand_mid = j.cand_mid; cs.offset = j.offset;
  return cs;
}
}  // namespace

__global__ void __launch_bounds__(128) seg_rans_explore_kernel(const RansJob* __restrict__ jobs, const Tile* __restrict__ tiles) {
  const Tile tl = load_tile(tiles);
  const RansJob& j = jobs[tl.seg];
  rans_explore_body(j.symbols, j.n, j.table, chunk_state_of(j), j.num_chunks, j.chunk_steps, j.warmup_steps, j.sub, j.stats, tl.first);
}
__global__ void __launch_bounds__(64) seg_rans_chain_kernel(const RansJob* __restrict__ jobs, const Tile* __restrict__ tiles) {
  const Tile tl = load_tile(tiles);
  const RansJob& j = jobs[tl.seg];
  rans_chain_body(j.symbols, j.n, j.table, chunk_state_of(j), j.num_chunks, j.chunk_steps, j.sub, j.stats);
}
__global__ void __launch_bounds__(128) seg_rans_encode_kernel(const RansJob* __restrict__ jobs, const Tile* __restrict__ tiles) {
  const Tile tl = load_tile(tiles);
  const RansJob& j = jobs[tl.seg];
  rans_encode_body(j.symbols, j.n, j.table, j.scratch, chunk_state_of(j), j.num_chunks, j.chunk_steps, j.stats, tl.first);
}
__global__ void __launch_bounds__(kLaneThreads) seg_rans_lanes_kernel(const RansJob* __restrict__ jobs, const Tile* __restrict__ tiles, uint32_t smem_rows) {
  const Tile tl = load_tile(tiles);
  const RansJob& j = jobs[tl.seg];
  __shared__ uint32_t s_abort;
  if (threadIdx.x == 0) s_abort = *(volatile uint32_t*)&j.stats->error_flags;
  __syncthreads();
  if (s_abort) return;
  const uint32_t P = j.stats->precision, K = j.stats->num_table_symbols;
  if (K < smem_rows) rans_encode_lanes_body<true>(j.symbols, j.n, j.table, j.scratch, chunk_state_of(j), j.num_chunks, j.num_pieces, j.piece_steps, j.sub, P, K, j.stats, tl.first);
  else rans_encode_lanes_body<false>(j.symbols, j.n, j.table, j.scratch, chunk_state_of(j), j.num_chunks, j.num_pieces, j.piece_steps, j.sub, P, K, j.stats, tl.first);
}
__global__ void __launch_bounds__(kFixupThreads) seg_rans_fixup_kernel(const RansJob* __restrict__ jobs, const Tile* __restrict__ tiles) {
  const Tile tl = load_tile(tiles);
  const RansJob& j = jobs[tl.seg];
  rans_fixup_body(j.symbols, j.n, j.table, j.scratch, chunk_state_of(j), j.num_pieces, j.piece_steps, j.stats);
}
__global__ void __launch_bounds__(256) seg_rans_gather_kernel(const RansJob* __restrict__ jobs, const Tile* __restrict__ tiles) {
  const Tile tl = load_tile(tiles);
  const RansJob& j = jobs[tl.seg];
  rans_gather_body(j.scratch, chunk_state_of(j), j.num_pieces, j.piece_steps, j.pieces_per_cta, j.payload, j.stats, tl.first);
}

// The layout of launch_rans_encode's scratch area, as a job descriptor.
RansJob rans_make_job(const uint32_t* symbols, uint64_t num_symbols, const uint4* rans_table, void* scratch, uint8_t* payload, AttrStats* stats) {
  const RansPlan plan = rans_plan_for(num_symbols);
  RansJob j{};
  j.symbols = symbols; j.n = num_symbols; j.table = rans_table; j.scratch = (uint8_t*)scratch; j.payload = payload; j.stats = stats;
  const uint32_t J = (uint32_t)((num_symbols + plan.chunk - 1) / plan.chunk);
  const bool lanes = plan.lanes && J > 1;
  const uint32_t piece = lanes ? plan.chunk / plan.sub : plan.chunk;
  const uint32_t Q = (uint32_t)((num_symbols + piece - 1) / piece);
  const size_t off = ((size_t)Q * rans_chunk_capacity(piece) + 255) / 256 * 256;
  uint32_t* u = (uint32_t*)(j.scratch + off);
  j.start = u; j.exit = u + Q; j.nbytes = u + 2 * (size_t)Q; j.offset = u + 3 * (size_t)Q;
  j.chain_start = u + 4 * (size_t)Q;
  j.cand_start = j.chain_start + J; j.cand_exit = j.cand_start + 32 * (size_t)J; j.cand_mid = j.cand_exit + 32 * (size_t)J;
  j.num_chunks = J; j.chunk_steps = plan.chunk; j.warmup_steps = plan.warmup; j.sub = plan.sub;
  j.num_pieces = Q; j.piece_steps = piece; j.lanes = lanes ? 1u : 0u;
  j.pieces_per_cta = piece <= 256 ? 8u : piece <= 512 ? 4u : piece <= 1024 ? 2u : 1u;
  return j;
}

void rans_plan_tiles(const RansJob* jobs, uint32_t num_jobs, RansTiles& out) {
  out = RansTiles{};
  for (uint32_t k = 0; k < num_jobs; ++k) {
    const RansJob& j = jobs[k];
    if (j.n == 0) continue;
    if (j.num_chunks > 1) {
      for (uint32_t b = 0; b < (j.num_chunks + 1) / 2; ++b) out.explore.push_back({k, b});
      out.chain.push_back({k, 0});
    }
    if (j.lanes) for (uint32_t b = 0; b < (j.num_pieces + kLaneThreads - 1) / kLaneThreads; ++b) out.lanes.push_back({k, b});
    else for (uint32_t b = 0; b < (j.num_chunks + 1) / 2; ++b) out.pairs.push_back({k, b});
    if (j.num_pieces > 1) out.fixup.push_back({k, 0});
    for (uint32_t b = 0; b < (j.num_pieces + j.pieces_per_cta - 1) / j.pieces_per_cta; ++b) out.gather.push_back({k, b});
  }
}

void launch_seg_rans(const RansJob* jobs, const RansTilesDev& t, uint32_t max_table_capacity, cudaStream_t s) {
  if (t.n_explore) seg_rans_explore_kernel<<<t.n_explore, 128, 0, s>>>(jobs, t.explore);
  if (t.n_chain) seg_rans_chain_kernel<<<t.n_chain, 64, 0, s>>>(jobs, t.chain);
  if (t.n_lanes) {
    const uint32_t smem_rows = std::min(max_table_capacity + 1u, kLaneSmemRows + 1u);
    const size_t sm = (size_t)smem_rows * 16 + (size_t)(kLaneThreads / 32) * 32 * kLanePitch * 4;
    allow_lane_smem((const void*)seg_rans_lanes_kernel);
    seg_rans_lanes_kernel<<<t.n_lanes, kLaneThreads, sm, s>>>(jobs, t.lanes, smem_rows);
  }
  if (t.n_pairs) seg_rans_encode_kernel<<<t.n_pairs, 128, 0, s>>>(jobs, t.pairs);
  if (t.n_fixup) seg_rans_fixup_kernel<<<t.n_fixup, kFixupThreads, 0, s>>>(jobs, t.fixup);
  if (t.n_gather) seg_rans_gather_kernel<<<t.n_gather, 256, 0, s>>>(jobs, t.gather);
}

// ---------------------------------------------------------------------------------------
// Output packing: one scan over the streams' sizes, then one CTA per stream copies its three parts.
__global__ void __launch_bounds__(1024) seg_pack_index_kernel(const AttrSeg* __restrict__ segs, uint32_t num_segs, uint4* __restrict__ index,
                                                              unsigned long long out_capacity) {
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint32_t base = 0; base < num_segs; base += 1024) {
    const uint32_t k = base + threadIdx.x;
    uint32_t tb = 0, pb = 0, sb = 0;
    if (k < num_segs) {
      const AttrSeg& g = segs[k];
      if (!g.stats->error_flags) { tb = g.stats->table_bytes; pb = g.stats->payload_bytes; sb = g.side_stats ? g.side_stats->nbytes : 0u; }
      if (tb > g.table_capacity) tb = 0;  // cannot happen: K9 checks its capacity
    }
    const uint32_t sz = ((tb + 3u) & ~3u) + ((pb + 3u) & ~3u) + ((sb + 3u) & ~3u);
    uint32_t inc = sz;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (lane >= (uint32_t)d) inc += o; }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    const uint32_t w = s_warp[lane];
    uint32_t winc = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, winc, d); if (lane >= (uint32_t)d) winc += o; }
    const uint32_t before = __shfl_sync(0xFFFFFFFFu, winc - w, warp);
    const uint32_t total = __shfl_sync(0xFFFFFFFFu, winc, 31);
    const uint32_t carry = s_carry;
    if (k < num_segs) {
      uint32_t off = carry + before + (inc - sz);
      if ((unsigned long long)off + sz > out_capacity) { off = 0; tb = pb = sb = 0xFFFFFFFFu; }  // reported to the host as an overflow
      index[k] = make_uint4(off, tb, pb, sb);
    }
    __syncthreads();
    if (threadIdx.x == 0) s_carry = carry + total;
    __syncthreads();
  }
  if (threadIdx.x == 0) index[num_segs] = make_uint4(s_carry, 0, 0, 0);
}

__global__ void __launch_bounds__(kThreads) seg_pack_copy_kernel(const AttrSeg* __restrict__ segs, const RansJob* __restrict__ jobs,
                                                                 const uint4* __restrict__ index, uint8_t* __restrict__ out) {
  const AttrSeg& g = segs[blockIdx.x];
  const uint4 ix = index[blockIdx.x];
  if (ix.y == 0xFFFFFFFFu) return;
  uint8_t* dst = out + ix.x;
  for (uint32_t i = threadIdx.x; i < ix.y; i += kThreads) dst[i] = g.table_bytes[i];
  dst += (ix.y + 3u) & ~3u;
  const uint8_t* __restrict__ pay = jobs[blockIdx.x].payload;  // 256-byte aligned (arena slots), dst is 4-byte aligned
  const uint32_t words = ix.z / 4;
  for (uint32_t i = threadIdx.x; i < words; i += kThreads) reinterpret_cast<uint32_t*>(dst)[i] = reinterpret_cast<const uint32_t*>(pay)[i];
  for (uint32_t i = words * 4 + threadIdx.x; i < ix.z; i += kThreads) dst[i] = pay[i];
  dst += (ix.z + 3u) & ~3u;
  for (uint32_t i = threadIdx.x; i < ix.w; i += kThreads) dst[i] = g.side_payload[i];
}

void launch_seg_pack(const AttrSeg* segs, const RansJob* jobs, uint32_t num_segs, uint4* index, uint8_t* out, uint64_t out_capacity, const Tile*, uint32_t,
                     cudaStream_t s) {
  if (!num_segs) return;
  seg_pack_index_kernel<<<1, 1024, 0, s>>>(segs, num_segs, index, out_capacity);
  seg_pack_copy_kernel<<<num_segs, kThreads, 0, s>>>(segs, jobs, index, out);
}
