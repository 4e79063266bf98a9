// Product CLI over the C ABI (include/dxo.h) — the two callers of the path (SURVEY §8b "Callers"):
//   dxo_cli -i in.obj -o out.drc                 cli/src/main.rs:33-73 (convert_obj_to_drc): load OBJ -> Config::default()
//                                                -> encode -> write file
//   dxo_cli --glb -o out.glb a.obj b.obj ...     the transcoder's output side for a set of primitives: one batch encode
//                                                (dxo_encode_batch inside dxo_encode_glb) and one GLB
//                                                (io/gltf/encode.rs:932-1097, :362-415). glTF INPUT parsing is out of scope.
// OBJ reading follows the reference's loader (io/obj/mod.rs:14-44: tobj with triangulate + single_index — (v, vt, vn)
// triples unified in first-use order, polygons fan-triangulated; positions, then normals, then texture coordinates, the
// latter two children of the position attribute); the mesh is then built on the device by dxo_mesh_build
// (MeshBuilder::build: value dedup, point merge, degenerate faces and unused points removed).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/dxo.h"

namespace {

struct PointArrays {
  std::vector<float> pos, nrm, uv;
  std::vector<uint32_t> faces;
};

bool read_file(const std::string& path, std::string& out) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  char buf[1 << 16];
  for (size_t n; (n = fread(buf, 1, sizeof buf, f)) > 0;) out.append(buf, n);
  fclose(f);
  return true;
}

bool load_obj_points(const std::string& path, PointArrays& m, std::string& err) {
  std::string text;
  if (!read_file(path, text)) { err = "cannot read " + path; return false; }
  std::vector<float> v, vt, vn;
  std::unordered_map<std::string, uint32_t> point_of;  // "iv/it/in" (resolved, 0 = absent) -> point
  std::vector<uint32_t> poly;
  const char* p = text.c_str();
  const char* end = p + text.size();
  auto skip_ws = [&](const char*& q) { while (q < end && (*q == ' ' || *q == '\t' || *q == '\r')) ++q; };
  while (p < end) {
    const char* eol = (const char*)memchr(p, '\n', (size_t)(end - p));
    if (!eol) eol = end;
    std::string line(p, eol);
    p = eol + (eol < end ? 1 : 0);
    const size_t hash = line.find('#');
    if (hash != std::string::npos) line.resize(hash);
    const char* q = line.c_str();
    const char* lend = q + line.size();
    while (q < lend && (*q == ' ' || *q == '\t' || *q == '\r')) ++q;
    auto floats = [&](const char* s, std::vector<float>& dst, int want) {
      for (int k = 0; k < want; ++k) { char* e = nullptr; const float f = strtof(s, &e); dst.push_back(e == s ? 0.0f : f); s = e; }
    };
    if (!strncmp(q, "v ", 2) || !strncmp(q, "v\t", 2)) floats(q + 2, v, 3);
    else if (!strncmp(q, "vt", 2) && (q[2] == ' ' || q[2] == '\t')) floats(q + 3, vt, 2);
    else if (!strncmp(q, "vn", 2) && (q[2] == ' ' || q[2] == '\t')) floats(q + 3, vn, 3);
    else if (!strncmp(q, "f ", 2) || !strncmp(q, "f\t", 2)) {
      poly.clear();
      const char* s = q + 2;
      (void)skip_ws;
      while (s < lend) {
        while (s < lend && (*s == ' ' || *s == '\t' || *s == '\r')) ++s;
        if (s >= lend) break;
        long idx[3] = {0, 0, 0};
        for (int k = 0; k < 3; ++k) {
          char* e = nullptr;
          const long x = strtol(s, &e, 10);
          if (e != s) idx[k] = x;
          s = e;
          if (s < lend && *s == '/') ++s; else break;
        }
        while (s < lend && *s != ' ' && *s != '\t' && *s != '\r') ++s;
        const long counts[3] = {(long)(v.size() / 3), (long)(vt.size() / 2), (long)(vn.size() / 3)};
        for (int k = 0; k < 3; ++k) {
          if (idx[k] < 0) idx[k] = counts[k] + idx[k] + 1;  // relative indices
          if (idx[k] < 0 || idx[k] > counts[k]) { err = "index out of range in " + path; return false; }
        }
        if (idx[0] == 0) { err = "face without a position index in " + path; return false; }
        const std::string key = std::to_string(idx[0]) + "/" + std::to_string(idx[1]) + "/" + std::to_string(idx[2]);
        auto it = point_of.find(key);
        uint32_t pt;
        if (it != point_of.end()) pt = it->second;
        else {
          pt = (uint32_t)point_of.size();
          point_of.emplace(key, pt);
          m.pos.insert(m.pos.end(), v.begin() + (idx[0] - 1) * 3, v.begin() + idx[0] * 3);
          if (idx[1] > 0) m.uv.insert(m.uv.end(), vt.begin() + (idx[1] - 1) * 2, vt.begin() + idx[1] * 2);
          if (idx[2] > 0) m.nrm.insert(m.nrm.end(), vn.begin() + (idx[2] - 1) * 3, vn.begin() + idx[2] * 3);
        }
        poly.push_back(pt);
      }
      for (size_t k = 1; k + 1 < poly.size(); ++k) { m.faces.push_back(poly[0]); m.faces.push_back(poly[k]); m.faces.push_back(poly[k + 1]); }
    }
  }
  if (m.faces.empty()) { err = "no faces in " + path; return false; }
  return true;
}

// per-point arrays -> the Mesh encode() takes (device dedup / point merge)
int build_mesh(const PointArrays& m, dxo_built_mesh** built, dxo_mesh* view) {
  const uint64_t points = m.pos.size() / 3;
  static const uint32_t parent0[1] = {0};
  dxo_attribute atts[3];
  uint32_t n = 0;
  auto add = [&](const std::vector<float>& vals, uint32_t type, uint32_t comps, uint32_t domain, bool child) {
    dxo_attribute a{};
    a.att_type = type; a.component_type = DXO_F32; a.num_components = comps; a.domain = domain;
    a.unique_id = n; a.num_parents = child ? 1 : 0; a.parent_ids = child ? parent0 : nullptr;
    a.num_unique_values = points; a.values = vals.data(); a.num_points = points; a.point_to_value = nullptr;
    atts[n++] = a;
  };
  add(m.pos, DXO_ATT_POSITION, 3, 0, false);
  if (!m.nrm.empty() && m.nrm.size() / 3 == points) add(m.nrm, DXO_ATT_NORMAL, 3, 1, true);
  if (!m.uv.empty() && m.uv.size() / 2 == points) add(m.uv, DXO_ATT_TEXCOORD, 2, 1, true);
  int st = dxo_mesh_build(m.faces.data(), m.faces.size() / 3, atts, n, -1, built);
  if (st == DXO_OK) st = dxo_built_mesh_view(*built, view);
  return st;
}

bool ends_with(const std::string& s, const char* suffix) {
  const size_t n = strlen(suffix);
  return s.size() >= n && s.compare(s.size() - n, n, suffix) == 0;
}

bool write_file(const std::string& path, const dxo_bytes& b) {
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) return false;
  const bool ok = fwrite(b.data, 1, b.len, f) == b.len;
  fclose(f);
  return ok;
}

}  // namespace

int main(int argc, char** argv) {
  std::string input, output;
  std::vector<std::string> inputs;
  bool glb = false;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    if ((a == "-i" || a == "--input") && i + 1 < argc) input = argv[++i];
    else if ((a == "-o" || a == "--output") && i + 1 < argc) output = argv[++i];
    else if (a == "--glb") glb = true;
    else if (a == "--transcode") { fprintf(stderr, "error: glTF input parsing is outside this library's scope (use --glb with OBJ primitives)\n"); return 2; }
    else inputs.push_back(a);
  }
  if (!input.empty()) inputs.insert(inputs.begin(), input);
  if (output.empty() || inputs.empty()) {
    fprintf(stderr, "usage: dxo_cli -i in.obj -o out.drc\n       dxo_cli --glb -o out.glb a.obj [b.obj ...]\n");
    return 2;
  }
  for (const std::string& in : inputs)
    if (!ends_with(in, ".obj")) { fprintf(stderr, "error: Input file must be a .obj file for conversion mode\n"); return 1; }  // cli/src/main.rs:42-44
  if (!glb && !ends_with(output, ".drc")) { fprintf(stderr, "error: Output file must be a .drc file for conversion mode\n"); return 1; }  // :52-54
  if (glb && !ends_with(output, ".glb")) { fprintf(stderr, "error: Output file must be a .glb file\n"); return 1; }

  std::vector<PointArrays> arrays(inputs.size());
  std::vector<dxo_built_mesh*> built(inputs.size(), nullptr);
  std::vector<dxo_mesh> meshes(inputs.size());
  int rc = 0;
  for (size_t k = 0; k < inputs.size() && rc == 0; ++k) {
    std::string err;
    if (!load_obj_points(inputs[k], arrays[k], err)) { fprintf(stderr, "error: Failed to load OBJ file: %s\n", err.c_str()); rc = 1; break; }
    const int st = build_mesh(arrays[k], &built[k], &meshes[k]);
    if (st != DXO_OK) { fprintf(stderr, "error: Failed to build mesh from %s: %s\n", inputs[k].c_str(), dxo_strerror(st)); rc = 1; }
  }
  if (rc == 0) {
    dxo_config cfg;
    dxo_config_default(&cfg);  // encode::Config::default()
    dxo_bytes out{nullptr, 0};
    const int st = glb ? dxo_encode_glb(meshes.data(), meshes.size(), &cfg, nullptr, &out, 0, 1) : dxo_encode(&meshes[0], &cfg, &out);
    if (st != DXO_OK) { fprintf(stderr, "error: Failed to encode mesh: %s\n", dxo_strerror(st)); rc = 1; }
    else if (!write_file(output, out)) { fprintf(stderr, "error: Failed to write output file: %s\n", output.c_str()); rc = 1; }
    else fprintf(stderr, "%zu primitive(s), %zu bytes -> %s\n", meshes.size(), out.len, output.c_str());
    dxo_free_bytes(&out);
  }
  for (dxo_built_mesh* b : built) if (b) dxo_built_mesh_free(b);
  return rc;
}
