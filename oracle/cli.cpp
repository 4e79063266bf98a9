// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_core.hpp header).
// CPU stand-in for `cli -i X.obj -o Y.drc` (reference: cli/src/main.rs:33-73):
// load_obj -> Config::default() -> encode -> write file.
#include <chrono>
#include <cstdio>
#include "orc_attribute.hpp"
#include "orc_mesh.hpp"

int main(int argc, char** argv) {
  std::string in, out;
  bool literal = false;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    if (a == "-i" && i + 1 < argc) in = argv[++i];
    else if (a == "-o" && i + 1 < argc) out = argv[++i];
    else if (a == "--literal") literal = true;
  }
  if (in.empty() || out.empty()) { fprintf(stderr, "usage: orc_cli -i in.obj -o out.drc [--literal]\n"); return 2; }
  try {
    orc::Mesh m = orc::load_obj(in);
    orc::OracleConfig cfg;
    cfg.literal = literal;
    auto t0 = std::chrono::steady_clock::now();
    orc::Bytes b = orc::encode_mesh(m, cfg);
    double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    FILE* f = fopen(out.c_str(), "wb");
    if (!f) { perror("fopen"); return 1; }
    fwrite(b.data(), 1, b.size(), f);
    fclose(f);
    fprintf(stderr, "faces=%zu points=%zu bytes=%zu encode_ms=%.3f\n", m.faces.size(), m.atts.empty() ? 0 : m.atts[0].len(), b.size(), ms);
  } catch (const std::exception& e) { fprintf(stderr, "error: %s\n", e.what()); return 1; }
  return 0;
}
