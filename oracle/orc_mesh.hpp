// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_core.hpp header).
// Caller side of the boundary: Attribute::from value dedup, MeshBuilder::build
// and an OBJ reader with tobj's `single_index + triangulate` semantics, so that
// the reference's fixtures produce the same Mesh the reference would hand to
// encode(). Paths relative to /root/reference/draco-oxide/src/.
//
// Third-party behaviour restated here (not in /root/reference): tobj "4.0.3"
// (draco-oxide/Cargo.toml) — single_index unifies (v,vt,vn) triples in first-use
// order, polygons are fan-triangulated. Pinned only by the reference's OBJ
// loader test (io/obj/mod.rs:73-89), which tests/test_oracle_known_answers.py repeats.
// MeshBuilder identifies equal points by SipHash-1-3 of the value bytes
// (core/mesh/builder.rs:242-279); this restatement compares the full key bytes.
#pragma once
#include <cstdio>
#include <fstream>
#include <sstream>
#include <unordered_map>
#include "orc_core.hpp"

namespace orc {

// Attribute::remove_duplicate_values — core/attribute/mod.rs:394-452.
// First occurrence wins, unique values keep first-occurrence order, equality is
// component-wise `==` on the typed values (-0.0 == 0.0, NaN != NaN; Appendix B.16).
// The all-pairs scan is replaced by a hash on canonicalised bytes (same result).
inline void remove_duplicate_values(Attribute& a) {
  const size_t vs = a.value_size(), n = a.num_unique(), cs = comp_size(a.comp_type);
  if (n == 0) return;
  std::unordered_map<std::string, uint32_t> first;
  first.reserve(n * 2);
  std::vector<uint32_t> map(n);
  std::vector<uint8_t> out;
  out.reserve(a.buffer.size());
  uint32_t next = 0;
  bool any_dup = false;
  for (size_t i = 0; i < n; ++i) {
    const uint8_t* p = a.buffer.data() + i * vs;
    std::string key((const char*)p, vs);
    bool has_nan = false;
    if (a.comp_type == CT_F32) {
      for (size_t k = 0; k < a.num_components; ++k) {
        float f; memcpy(&f, p + k * cs, 4);
        if (f != f) has_nan = true;
        if (f == 0.0f) { float z = 0.0f; memcpy(&key[k * cs], &z, 4); }
      }
    } else if (a.comp_type == CT_F64) {
      for (size_t k = 0; k < a.num_components; ++k) {
        double f; memcpy(&f, p + k * cs, 8);
        if (f != f) has_nan = true;
        if (f == 0.0) { double z = 0.0; memcpy(&key[k * cs], &z, 8); }
      }
    }
    if (!has_nan) {
      auto it = first.find(key);
      if (it != first.end()) { map[i] = it->second; any_dup = true; continue; }
      first.emplace(std::move(key), next);
    }
    map[i] = next++;
    out.insert(out.end(), p, p + vs);
  }
  if (any_dup) { a.has_map = true; a.map = std::move(map); a.buffer = std::move(out); }
}

// Attribute::remove applied to a set of points (core/attribute/mod.rs:454-482),
// batched: mapping entries of removed points disappear; values no remaining
// point refers to are dropped and larger indices shift down.
inline void remove_points(Attribute& a, const std::vector<uint32_t>& removed_sorted) {
  if (removed_sorted.empty()) return;
  const size_t vs = a.value_size();
  std::vector<uint8_t> gone(a.len(), 0);
  for (uint32_t p : removed_sorted) { if (p >= a.len()) throw Panic(ST_INVALID_ARGUMENT, "Point index out of bounds"); gone[p] = 1; }
  if (a.has_map) {
    std::vector<uint32_t> nm;
    nm.reserve(a.map.size());
    std::vector<uint32_t> ref(a.num_unique(), 0);
    for (size_t p = 0; p < a.map.size(); ++p) if (!gone[p]) { nm.push_back(a.map[p]); ref[a.map[p]]++; }
    std::vector<uint32_t> newidx(ref.size(), NONE);
    std::vector<uint8_t> nb;
    uint32_t k = 0;
    for (size_t v = 0; v < ref.size(); ++v) if (ref[v]) { newidx[v] = k++; nb.insert(nb.end(), a.buffer.begin() + v * vs, a.buffer.begin() + (v + 1) * vs); }
    for (auto& m : nm) m = newidx[m];
    a.map = std::move(nm);
    a.buffer = std::move(nb);
  } else {
    std::vector<uint8_t> nb;
    for (size_t v = 0; v < gone.size(); ++v) if (!gone[v]) nb.insert(nb.end(), a.buffer.begin() + v * vs, a.buffer.begin() + (v + 1) * vs);
    a.buffer = std::move(nb);
  }
}

struct MeshBuilder {  // core/mesh/builder.rs:15-90
  std::vector<Attribute> attributes;
  std::vector<std::array<uint32_t, 3>> faces;
  uint32_t current_id = 0;

  uint32_t add_attribute(const void* data, size_t count, uint32_t comp_type, uint32_t ncomp, uint32_t att_type, uint32_t domain,
                         std::vector<uint32_t> parents) {  // :30-39 -> Attribute::from (core/attribute/mod.rs:87-103)
    Attribute a;
    a.id = current_id++;
    a.att_type = att_type; a.comp_type = comp_type; a.num_components = ncomp; a.domain = domain; a.parents = std::move(parents);
    a.buffer.assign((const uint8_t*)data, (const uint8_t*)data + count * ncomp * comp_size(comp_type));
    remove_duplicate_values(a);
    attributes.push_back(std::move(a));
    return attributes.back().id;
  }

  Mesh build() {  // :62-90
    // dependency_check — :94-111 (TextureCoordinate needs a Position parent)
    for (auto& a : attributes) {
      if (a.att_type == AT_TEXCOORD) {
        bool ok = false;
        for (uint32_t pid : a.parents) for (auto& b : attributes) if (b.id == pid && b.att_type == AT_POSITION) ok = true;
        if (!ok) throw EncodeError(ST_INVALID_ARGUMENT, "MinimumDependencyError: TextureCoordinate must depend on Position");
      }
    }
    // get_sorted_attributes — :115-125
    for (size_t i = 0; i < attributes.size(); ++i) if (attributes[i].att_type == AT_POSITION) { std::swap(attributes[0], attributes[i]); break; }
    deduplicate_points();
    // drop degenerate faces — :76-79
    std::vector<std::array<uint32_t, 3>> nf;
    for (auto& f : faces) if (f[0] != f[1] && f[1] != f[2] && f[2] != f[0]) nf.push_back(f);
    faces = std::move(nf);
    remove_unused_vertices();
    Mesh m;
    m.faces = std::move(faces);
    m.atts = std::move(attributes);
    return m;
  }

  void deduplicate_points() {  // deduplicate_vertices_based_on_positions — :194-239 (+hash_vertex :242-269, remap_attribute :272-373)
    if (attributes.empty()) return;
    size_t num_vertices = 0;
    for (auto& f : faces) for (uint32_t p : f) num_vertices = std::max<size_t>(num_vertices, p);
    num_vertices += 1;
    std::unordered_map<std::string, uint32_t> unique_points;
    std::vector<uint32_t> point_mapping;
    point_mapping.reserve(num_vertices);
    uint32_t unique_count = 0;
    for (size_t p = 0; p < num_vertices; ++p) {
      std::string key;
      for (auto& a : attributes) {
        if (p < a.len()) {
          uint32_t hdr[3] = {a.att_type, a.comp_type, a.num_components};
          key.append((const char*)hdr, sizeof hdr);
          uint32_t v = a.unique_val_idx((uint32_t)p);
          key.append((const char*)a.buffer.data() + (size_t)v * a.value_size(), a.value_size());
        }
      }
      auto it = unique_points.find(key);
      if (it != unique_points.end()) point_mapping.push_back(it->second);
      else { unique_points.emplace(std::move(key), unique_count); point_mapping.push_back(unique_count++); }
    }
    if (unique_count == num_vertices) return;
    for (auto& a : attributes) {
      if (unique_count == a.len()) continue;  // remap_attribute early return (:274-276)
      std::vector<uint8_t> met(unique_count, 0);
      std::vector<uint32_t> removed;
      for (size_t v = 0; v < point_mapping.size(); ++v) { if (met[point_mapping[v]]) removed.push_back((uint32_t)v); else met[point_mapping[v]] = 1; }
      remove_points(a, removed);
    }
    for (auto& f : faces) for (auto& p : f) p = point_mapping[p];
  }

  void remove_unused_vertices() {  // :129-189
    if (faces.empty() || attributes.empty()) return;
    uint32_t max_idx = 0;
    for (auto& f : faces) for (uint32_t p : f) max_idx = std::max(max_idx, p);
    std::vector<uint8_t> used((size_t)max_idx + 1, 0);
    for (auto& f : faces) for (uint32_t p : f) used[p] = 1;
    std::vector<uint32_t> unused;
    for (size_t i = 0; i < used.size(); ++i) if (!used[i]) unused.push_back((uint32_t)i);
    for (auto& a : attributes) {
      std::vector<uint32_t> rem = unused;
      for (size_t p = (size_t)max_idx + 1; p < a.len(); ++p) rem.push_back((uint32_t)p);
      remove_points(a, rem);
    }
    std::vector<uint32_t> offsets(used.size());
    uint32_t removed_count = 0;
    for (size_t v = 0; v < used.size(); ++v) { offsets[v] = removed_count; if (!used[v]) removed_count++; }
    for (auto& f : faces) for (auto& p : f) p -= offsets[p];
  }
};

// load_obj — io/obj/mod.rs:14-44 with tobj::LoadOptions{triangulate, single_index}
inline Mesh load_obj(const std::string& path) {
  std::ifstream in(path);
  if (!in) throw EncodeError(ST_INVALID_ARGUMENT, "Failed to load OBJ file: " + path);
  std::vector<float> v, vt, vn;
  struct Key { long a, b, c; bool operator==(const Key& o) const { return a == o.a && b == o.b && c == o.c; } };
  struct KeyHash { size_t operator()(const Key& k) const { return std::hash<long>()(k.a * 73856093L ^ k.b * 19349663L ^ k.c * 83492791L); } };
  std::unordered_map<Key, uint32_t, KeyHash> index;
  std::vector<float> positions, texcoords, normals;
  std::vector<uint32_t> indices;
  std::string line;
  while (std::getline(in, line)) {
    size_t hash_pos = line.find('#');
    if (hash_pos != std::string::npos) line = line.substr(0, hash_pos);
    std::istringstream ss(line);
    std::string tag;
    if (!(ss >> tag)) continue;
    if (tag == "v") { float x, y, z; ss >> x >> y >> z; v.push_back(x); v.push_back(y); v.push_back(z); }
    else if (tag == "vt") { float x = 0, y = 0; ss >> x >> y; vt.push_back(x); vt.push_back(y); }
    else if (tag == "vn") { float x, y, z; ss >> x >> y >> z; vn.push_back(x); vn.push_back(y); vn.push_back(z); }
    else if (tag == "f") {
      std::vector<uint32_t> poly;
      std::string tok;
      while (ss >> tok) {
        long iv = 0, it = 0, in_ = 0;  // 0 = missing
        {
          size_t s1 = tok.find('/');
          if (s1 == std::string::npos) iv = std::stol(tok);
          else {
            iv = std::stol(tok.substr(0, s1));
            size_t s2 = tok.find('/', s1 + 1);
            std::string t = s2 == std::string::npos ? tok.substr(s1 + 1) : tok.substr(s1 + 1, s2 - s1 - 1);
            if (!t.empty()) it = std::stol(t);
            if (s2 != std::string::npos) { std::string nn = tok.substr(s2 + 1); if (!nn.empty()) in_ = std::stol(nn); }
          }
        }
        auto fix = [](long i, size_t count) -> long { return i < 0 ? (long)count + i + 1 : i; };
        iv = fix(iv, v.size() / 3); it = fix(it, vt.size() / 2); in_ = fix(in_, vn.size() / 3);
        Key key{iv, it, in_};
        auto f = index.find(key);
        uint32_t id;
        if (f != index.end()) id = f->second;
        else {
          id = (uint32_t)index.size();
          index.emplace(key, id);
          positions.insert(positions.end(), v.begin() + (iv - 1) * 3, v.begin() + (iv - 1) * 3 + 3);
          if (it > 0) texcoords.insert(texcoords.end(), vt.begin() + (it - 1) * 2, vt.begin() + (it - 1) * 2 + 2);
          if (in_ > 0) normals.insert(normals.end(), vn.begin() + (in_ - 1) * 3, vn.begin() + (in_ - 1) * 3 + 3);
        }
        poly.push_back(id);
      }
      for (size_t k = 1; k + 1 < poly.size(); ++k) { indices.push_back(poly[0]); indices.push_back(poly[k]); indices.push_back(poly[k + 1]); }
    }
  }
  MeshBuilder b;
  for (size_t i = 0; i + 2 < indices.size(); i += 3) b.faces.push_back({indices[i], indices[i + 1], indices[i + 2]});
  uint32_t pos_id = b.add_attribute(positions.data(), positions.size() / 3, CT_F32, 3, AT_POSITION, 0, {});
  if (!normals.empty()) b.add_attribute(normals.data(), normals.size() / 3, CT_F32, 3, AT_NORMAL, 1, {pos_id});
  if (!texcoords.empty()) b.add_attribute(texcoords.data(), texcoords.size() / 2, CT_F32, 2, AT_TEXCOORD, 1, {pos_id});
  return b.build();
}

}  // namespace orc
