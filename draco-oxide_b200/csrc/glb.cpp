// Product host code — GLB assembly around the batch entry: what the reference's glTF writer does
// with each encoded primitive — io/gltf/encode.rs:932-1097 (add_draco_mesh_internal: Draco bytes appended to the GLB BIN
// buffer and padded to 4, one bufferView per primitive, placeholder accessors without bufferViews for the indices and
// each attribute, POSITION bounds from compute_vec3_bounds :815-856, KHR_draco_mesh_compression ids) and :362-415
// (write_glb_format: 12-byte header, JSON chunk padded with spaces, BIN chunk padded with zeros).
// The scene around the primitives is the minimal one: one mesh + node per primitive, one scene. OBJ / glTF parsing and
// materials are out of scope.
#include <cmath>
#include <cstdio>
#include <string>

#include "encoder.hpp"

namespace dxo {
namespace {

void put_u32(std::vector<uint8_t>& b, uint32_t v) { for (int k = 0; k < 4; ++k) b.push_back((uint8_t)(v >> (8 * k))); }

std::string fmt_float(float f) {  // shortest decimal that reads back as the same f32
  if (f == 0.0f) return std::signbit(f) ? "-0.0" : "0.0";
  char buf[64];
  for (int prec = 1; prec <= 9; ++prec) {
    snprintf(buf, sizeof buf, "%.*g", prec, (double)f);
    if (strtof(buf, nullptr) == f) break;
  }
  std::string s(buf);
  if (s.find_first_of(".eEn") == std::string::npos) s += ".0";
  return s;
}

// compute_vec3_bounds (io/gltf/encode.rs:815-856): true per-point bounds starting from point 0, f32::min / max (NaNs skipped)
void position_bounds(const dxo_attribute& a, float* mn, float* mx, bool* any) {
  const float* v = (const float*)a.values;
  const uint32_t n = a.num_components;
  const uint64_t points = a.point_to_value ? a.num_points : a.num_unique_values;
  *any = points > 0;
  for (uint64_t p = 0; p < points; ++p) {
    const float* x = v + (size_t)(a.point_to_value ? a.point_to_value[p] : p) * n;
    for (uint32_t k = 0; k < n && k < 3; ++k) {
      if (p == 0) { mn[k] = mx[k] = x[k]; continue; }
      mn[k] = std::fmin(mn[k], x[k]);
      mx[k] = std::fmax(mx[k], x[k]);
      if (x[k] == 0.0f && mn[k] == 0.0f && std::signbit(x[k])) mn[k] = x[k];   // -0.0 orders below +0.0 (as dxo_attribute_bounds)
      if (x[k] == 0.0f && mx[k] == 0.0f && !std::signbit(x[k])) mx[k] = x[k];
    }
  }
}

}  // namespace

// One GLB holding `n` Draco-compressed primitives. streams[i] = the .drc bytes of meshes[i] (dxo_encode / dxo_encode_batch).
void assemble_glb(const dxo_mesh* meshes, const dxo_bytes* streams, size_t n, std::vector<uint8_t>& out) {
  std::vector<uint8_t> bin;
  std::string views, accessors, gl_meshes, nodes, scene_nodes;
  size_t num_accessors = 0, num_views = 0, num_meshes = 0;
  auto comma = [](std::string& s) { if (!s.empty()) s += ","; };
  for (size_t i = 0; i < n; ++i) {
    const dxo_mesh& m = meshes[i];
    if (m.num_faces == 0) continue;  // :934-937 — empty meshes are skipped
    if (!streams[i].data || !streams[i].len) throw Error(DXO_ERR_INVALID_ARGUMENT, "primitive without an encoded stream");
    const size_t start = bin.size();
    bin.insert(bin.end(), streams[i].data, streams[i].data + streams[i].len);
    while (bin.size() % 4) bin.push_back(0);  // pad_buffer
    comma(views);
    views += "{\"buffer\":0,\"byteOffset\":" + std::to_string(start) + ",\"byteLength\":" + std::to_string(bin.size() - start) + "}";
    const size_t view = num_views++;
    // indices placeholder (:984-996)
    comma(accessors);
    accessors += "{\"componentType\":5121,\"count\":" + std::to_string(m.num_faces * 3) + ",\"type\":\"SCALAR\"}";
    const size_t indices = num_accessors++;
    std::string attrs, draco_attrs;
    // POSITION, NORMAL, TEXCOORD_0 accessors in that order; Draco ids as the reference assigns them (:1009-1016)
    const uint32_t order[3] = {DXO_ATT_POSITION, DXO_ATT_NORMAL, DXO_ATT_TEXCOORD};
    const char* names[3] = {"POSITION", "NORMAL", "TEXCOORD_0"};
    for (int t = 0; t < 3; ++t) {
      for (uint32_t k = 0; k < m.num_attributes; ++k) {
        const dxo_attribute& a = m.attributes[k];
        if (a.att_type != order[t]) continue;
        const uint64_t count = a.point_to_value ? a.num_points : a.num_unique_values;  // Attribute::len
        const int draco_id = t == 0 ? 1 : t == 1 ? 0 : (int)k;
        comma(accessors);
        accessors += "{\"componentType\":5126,\"count\":" + std::to_string(count) + ",\"type\":\"" + (t == 2 ? "VEC2" : "VEC3") + "\"";
        if (t == 0 && a.component_type == DXO_F32) {
          float mn[3] = {0, 0, 0}, mx[3] = {0, 0, 0};
          bool any = false;
          position_bounds(a, mn, mx, &any);
          if (any) {
            accessors += ",\"max\":[" + fmt_float(mx[0]) + "," + fmt_float(mx[1]) + "," + fmt_float(mx[2]) + "]";
            accessors += ",\"min\":[" + fmt_float(mn[0]) + "," + fmt_float(mn[1]) + "," + fmt_float(mn[2]) + "]";
          }
        }
        accessors += "}";
        comma(attrs); comma(draco_attrs);
        attrs += std::string("\"") + names[t] + "\":" + std::to_string(num_accessors++);
        draco_attrs += std::string("\"") + names[t] + "\":" + std::to_string(draco_id);
        break;  // the first attribute of a type is the one the writer records
      }
    }
    comma(gl_meshes);
    gl_meshes += "{\"primitives\":[{\"attributes\":{" + attrs + "},\"indices\":" + std::to_string(indices) +
                 ",\"mode\":4,\"extensions\":{\"KHR_draco_mesh_compression\":{\"bufferView\":" + std::to_string(view) + ",\"attributes\":{" + draco_attrs + "}}}}]}";
    comma(nodes); comma(scene_nodes);
    nodes += "{\"mesh\":" + std::to_string(num_meshes) + "}";
    scene_nodes += std::to_string(num_meshes);
    ++num_meshes;
  }
  std::string json = "{\"asset\":{\"version\":\"2.0\",\"generator\":\"draco-oxide_b200\"}";
  if (num_meshes) {
    json += ",\"extensionsUsed\":[\"KHR_draco_mesh_compression\"],\"extensionsRequired\":[\"KHR_draco_mesh_compression\"]";
    json += ",\"buffers\":[{\"byteLength\":" + std::to_string(bin.size()) + "}]";
    json += ",\"bufferViews\":[" + views + "],\"accessors\":[" + accessors + "],\"meshes\":[" + gl_meshes + "],\"nodes\":[" + nodes + "]";
  }
  json += ",\"scenes\":[{\"nodes\":[" + scene_nodes + "]}],\"scene\":0}";
  // write_glb_format (:362-415)
  const size_t json_padded = (json.size() + 3) & ~(size_t)3, bin_padded = bin.empty() ? 0 : ((bin.size() + 3) & ~(size_t)3);
  const size_t total = 12 + 8 + json_padded + (bin_padded ? 8 + bin_padded : 0);
  if (total > 0xFFFFFFFFull) throw Error(DXO_ERR_UNSUPPORTED_INPUT, "GLB larger than 4 GiB");
  out.clear();
  out.reserve(total);
  out.insert(out.end(), {'g', 'l', 'T', 'F'});
  put_u32(out, 2);
  put_u32(out, (uint32_t)total);
  put_u32(out, (uint32_t)json_padded);
  out.insert(out.end(), {'J', 'S', 'O', 'N'});
  out.insert(out.end(), json.begin(), json.end());
  out.insert(out.end(), json_padded - json.size(), (uint8_t)' ');
  if (bin_padded) {
    put_u32(out, (uint32_t)bin_padded);
    out.insert(out.end(), {'B', 'I', 'N', 0});
    out.insert(out.end(), bin.begin(), bin.end());
    out.insert(out.end(), bin_padded - bin.size(), (uint8_t)0);
  }
}

}  // namespace dxo
