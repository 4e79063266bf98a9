"""GPU parity tests (run with -m gpu on the B200 box). Every test goes through the
C ABI (include/dxo.h) and compares against the CPU oracle on the same seeded inputs:
bit-exact for every integer / byte result (quantized values, symbols, histograms,
normalised tables, table bytes, rANS payloads, side bits) and byte-identical streams."""
import ctypes as C
import hashlib

import numpy as np
import pytest

import draco_oxide_b200 as dxo
import drc_parse
import meshes
from draco_oxide_b200 import _capi, synth

pytestmark = pytest.mark.gpu

STAGES = [("quantized", np.int32), ("symbols", np.uint32), ("histogram", np.uint64), ("distribution", np.uint64),
          ("table_bytes", np.uint8), ("payload", np.uint8), ("side_bits", np.uint8), ("wrap_minmax", np.int32),
          ("bit_length", np.uint32), ("sequence", np.uint32)]


@pytest.fixture(scope="module", autouse=True)
def _built():
    import __graft_entry__ as g
    g.build()
    assert dxo.device_count() >= 1, "GPU tests need a CUDA device"


def gpu_encode(mesh, cfg=None):
    out = bytearray()
    dxo.encode(mesh, out, cfg)
    return bytes(out)


def assert_stage_parity(orc, mesh, cfg=None):
    drc, tr = orc.encode(mesh, cfg, trace=True)
    s = dxo.Session(mesh, cfg)
    s.set_trace(True)
    got = s.run()
    for i, a in enumerate(mesh.attributes):
        for key, dt in STAGES:
            if key == "wrap_minmax" and a.att_type == dxo.AttributeType.Normal:
                continue
            ref, dev = tr.get(f"att{i}.{key}", dt), s.trace(f"att{i}.{key}", dt)
            assert np.array_equal(ref, dev), f"att{i}.{key}: first diff at {np.flatnonzero(ref[:min(ref.size, dev.size)] != dev[:min(ref.size, dev.size)])[:5]} sizes {ref.size}/{dev.size}"
    assert got == drc
    again = s.run()
    assert again == drc, "second run of a resident session differs"
    s.close()
    return drc


@pytest.mark.parametrize("name", sorted(meshes.zoo().keys()))
def test_zoo_streams_byte_identical(orc, name):
    m = meshes.drop_unused_points(meshes.zoo()[name])
    assert gpu_encode(m) == orc.encode(m)


@pytest.mark.parametrize("name", meshes.golden_names())
def test_reference_fixtures_match_golden(name):
    mesh, drc = meshes.load_golden(name)
    assert gpu_encode(mesh) == drc


@pytest.mark.parametrize("name", ["grid_medium", "torus_medium", "custom_attribute", "color_attribute", "grid_with_hole", "single_triangle"])
def test_stage_by_stage_parity(orc, name):
    assert_stage_parity(orc, meshes.drop_unused_points(meshes.zoo()[name]))


def test_config1_50k_triangles(orc):
    m = synth.config1_mesh()
    drc = assert_stage_parity(orc, m)
    assert hashlib.sha256(drc).hexdigest() == open(meshes.GOLDEN + "/config1.sha256").read().strip()


def test_config2_1m_vertex_grid_full_size(orc):
    """BASELINE config 2 at full size: 1 000 000 vertices, 1 996 002 triangles."""
    m = synth.config2_mesh()
    assert m.num_points() == 1_000_000 and m.faces.shape[0] == 1_996_002
    assert gpu_encode(m) == orc.encode(m)


def test_config3_quarter_size_torus_with_seams(orc):
    m = synth.torus_mesh(1000, 1250, 3)  # 2.5M triangles, uv seams on both cuts, position point map
    assert gpu_encode(m) == orc.encode(m)


def test_config3_full_size_byte_parity(orc):
    """BASELINE config 3 at full size (2000 x 2500-quad torus, 10 000 000 triangles, uv seams on both cuts):
    the stream is byte-identical to the oracle's."""
    m = synth.config3_mesh()
    assert m.faces.shape[0] == 10_000_000 and m.attributes[0].num_unique_values == 5_000_000
    assert gpu_encode(m) == orc.encode(m)


def test_config3_full_size_properties(orc):
    """BASELINE config 3 (10M triangles): size-independent properties — the stream parses,
    every rANS / rABS section decodes back to the symbols and bits the GPU produced, and
    a second encode is bit-identical."""
    m = synth.config3_mesh()
    assert m.faces.shape[0] == 10_000_000
    s = dxo.Session(m)
    s.set_trace(True)
    drc = s.run()
    head = s.trace("head_bytes", np.uint8)
    n = len(m.attributes)
    seq_lens = [s.trace(f"att{i}.sequence", np.uint32).size for i in range(n)]
    parsed = drc_parse.parse_attributes(drc, head.size - 1 - 10 * n, seq_lens)
    for i, d in enumerate(parsed):
        assert np.array_equal(d["symbols"], s.trace(f"att{i}.symbols", np.uint32))
        if "side_bits" in d:
            assert np.array_equal(d["side_bits"], s.trace(f"att{i}.side_bits", np.uint8))
    # every position vertex is sequenced exactly once
    seq0 = s.trace("att0.sequence", np.uint32)
    c2v = s.trace("corner_to_vertex", np.uint32)
    assert np.unique(c2v[seq0]).size == seq0.size == 5_000_000
    assert s.run() == drc
    s.close()


@pytest.mark.parametrize("bits", [8, 10, 12, 14, 16])
def test_quantization_bit_sweep(orc, bits):
    """BASELINE config 5 (qp sweep); bits != 11 is oracle-defined, the reference hard-codes 11."""
    m = synth.grid_mesh(120, 90, 5)
    cfg = dxo.Config(position_bits=bits, texcoord_bits=min(bits, 12))
    assert_stage_parity(orc, m, cfg)


@pytest.fixture(scope="module")
def config2_mesh():
    return synth.config2_mesh()


@pytest.mark.parametrize("bits", [8, 10, 11, 12, 14, 16])
def test_config5_qp_sweep_on_the_1m_vertex_mesh(orc, config2_mesh, bits):
    """BASELINE config 5 at its stated size: qp in {8,10,11,12,14,16} on config 2's 1M-vertex mesh, Edgebreaker
    (the reference has no sequential-connectivity attribute path). Byte parity against the oracle."""
    cfg = dxo.Config(position_bits=bits)
    assert gpu_encode(config2_mesh, cfg) == orc.encode(config2_mesh, cfg)


def _config4_full_batch(num_gpus):
    """All 4096 primitives of BASELINE config 4 (87.9M vertices) through dxo_encode_batch, in slabs of 512 to bound
    host memory; every stream's sha256 against tests/golden/config4_hashes.txt (oracle bytes,
    tests/golden/make_config4_hashes.py) and the checksum of checksums."""
    lines = open(meshes.GOLDEN + "/config4_hashes.txt").read().split()
    counts = synth.batch_vertex_counts()
    assert counts.size == 4096 and len(lines) == 4097
    digests = []
    for lo in range(0, counts.size, 512):
        ms = synth.batch_meshes(counts[lo:lo + 512], first=lo)
        got = dxo.encode_batch(ms, first_gpu=0, num_gpus=num_gpus)
        assert len(got) == len(ms)
        for k, g in enumerate(got):
            d = hashlib.sha256(g).hexdigest()
            assert d[:16] == lines[lo + k], f"primitive {lo + k} ({ms[k].num_points()} points) differs from the oracle's stream"
            digests.append(d)
    assert hashlib.sha256("".join(digests).encode()).hexdigest() == lines[4096]


def test_config4_full_batch_one_gpu():
    _config4_full_batch(1)


def test_config4_full_batch_over_all_gpus():
    """Runs when the lease has more than one GPU (`gpurun --gpus N`): the batch is sharded over all of them."""
    n = dxo.device_count()
    if n < 2:
        pytest.skip("single-GPU lease: covered by test_config4_full_batch_one_gpu")
    _config4_full_batch(n)


def test_batch_entry_matches_per_mesh_streams(orc):
    counts = synth.batch_vertex_counts(24, 200, 4000, seed=5)
    ms = [synth.batch_mesh(k, int(c)) for k, c in enumerate(counts)]
    got = dxo.encode_batch(ms, first_gpu=0, num_gpus=1)
    for m, g in zip(ms, got):
        assert g == orc.encode(m)


def test_batch_entry_mixed_zoo_with_fallbacks_and_errors(orc):
    """One batch holding every zoo mesh (non-manifold edges and vertices, holes, several components, custom and colour
    attributes, a single triangle), a mesh with a zero normal and one whose face points outside its attributes: the
    meshes the group kernels flag take the per-mesh path, failures are reported per mesh, everything else is unaffected."""
    zoo = meshes.zoo()
    names = sorted(zoo.keys())
    ms = [meshes.drop_unused_points(zoo[k]) for k in names]
    g = synth.grid_mesh(12, 12, 3)
    n = g.attributes[1]
    vals = n.values.copy()
    vals[5] = 0
    ms.append(dxo.Mesh(g.faces, [g.attributes[0], dxo.Attribute(vals, n.att_type, n.domain, n.parents, n.point_to_value, n.unique_id), g.attributes[2]]))
    bad_faces = g.faces.copy()
    bad_faces[3, 1] = g.num_points() + 7
    ms.append(dxo.Mesh(bad_faces, g.attributes))
    ms.append(synth.torus_mesh(30, 20, 8))
    got, sts = dxo.encode_batch(ms, first_gpu=0, num_gpus=1, return_statuses=True)
    for k, name in enumerate(names):
        assert sts[k] == 0, name
        assert got[k] == orc.encode(ms[k]), name
    assert sts[len(names)] == -9 and got[len(names)] == b""       # DXO_ERR_ZERO_NORMAL
    assert sts[len(names) + 1] == -1 and got[len(names) + 1] == b""  # DXO_ERR_INVALID_ARGUMENT
    assert sts[-1] == 0 and got[-1] == orc.encode(ms[-1])


def test_batch_entry_over_all_visible_gpus(orc):
    """The batch entry shards independent meshes over every visible GPU (one on the default test box, N under
    `gpurun --gpus N`); results come back in input order and equal the per-mesh streams, large meshes included."""
    n_gpus = dxo.device_count()
    counts = list(synth.batch_vertex_counts(20, 500, 30000, seed=9)) + [60000, 90000]
    ms = [synth.batch_mesh(k, int(c)) for k, c in enumerate(counts)]
    got = dxo.encode_batch(ms, first_gpu=0, num_gpus=n_gpus)
    assert len(got) == len(ms)
    for m, g in zip(ms, got):
        out = bytearray(); dxo.encode(m, out)
        assert g == bytes(out)
    for k in (0, 7, len(ms) - 1):
        assert got[k] == orc.encode(ms[k])


def test_non_f32_component_types(orc):
    """Quantised attributes whose components are not f32 (the glTF path creates u8 colours, u16 joints / texcoords): converted
    with `to_f64() as f32` like the reference's quantiser (quantization_coordinate_wise.rs:30-90), header keeps the original type."""
    g = synth.grid_mesh(21, 17, 12)
    n = g.num_points()
    rng = np.random.default_rng(8)
    color = dxo.Attribute.from_points(rng.integers(0, 256, (n, 4)).astype(np.uint8), dxo.AttributeType.Color, dxo.AttributeDomain.Corner, (), 3)
    joints = dxo.Attribute.from_points((np.arange(n)[:, None] // np.array([3, 5, 7, 11])).astype(np.uint16), dxo.AttributeType.Joint, dxo.AttributeDomain.Corner, (), 4)
    weights = dxo.Attribute.from_points(rng.random((n, 2)), dxo.AttributeType.Weight, dxo.AttributeDomain.Corner, (), 5)  # f64
    uv16 = dxo.Attribute.from_points((g.attributes[2].values * 1000).astype(np.int16), dxo.AttributeType.TextureCoordinate, dxo.AttributeDomain.Corner, (0,), 2)
    m = dxo.Mesh(g.faces, [g.attributes[0], g.attributes[1], uv16, color, joints, weights])
    drc = gpu_encode(m)
    assert drc == orc.encode(m)
    orc.assert_decodes(m, drc)
    got = dxo.encode_batch([m, g])
    assert got[0] == drc and got[1] == orc.encode(g)


def test_zero_normal_error_code():
    m = synth.grid_mesh(6, 6, 3)
    n = m.attributes[1]
    vals = n.values.copy()
    vals[3] = 0
    bad = dxo.Mesh(m.faces, [m.attributes[0], dxo.Attribute(vals, n.att_type, n.domain, n.parents, n.point_to_value, n.unique_id), m.attributes[2]])
    with pytest.raises(dxo.Err) as e:
        gpu_encode(bad)
    assert e.value.status == -9  # DXO_ERR_ZERO_NORMAL (geom.rs:45)
    assert gpu_encode(m)  # the library stays usable afterwards


def test_ragged_and_tiny_inputs(orc):
    for nx, ny in [(2, 2), (2, 3), (3, 2), (2, 50), (33, 2), (5, 5)]:
        m = synth.grid_mesh(nx, ny, nx * 100 + ny)
        assert gpu_encode(m) == orc.encode(m), (nx, ny)


def test_large_alphabet_global_histogram_path(orc):
    """qp=16 on a rough surface: alphabet > the shared-memory histogram (global atomics path)."""
    m = synth.grid_mesh(300, 300, 9, with_normals=False)
    pos = m.attributes[0].values.copy()
    pos[:, 2] = np.random.default_rng(3).random(pos.shape[0]).astype(np.float32)
    m = dxo.Mesh(m.faces, [dxo.Attribute.from_points(pos, 0, 0)] + m.attributes[1:])
    cfg = dxo.Config(position_bits=16, texcoord_bits=10)
    assert_stage_parity(orc, m, cfg)


def test_corner_table_kernel(orc):
    """K12 half-edge matching by radix sort against the sequential matcher."""
    L = _capi.lib()
    for name, mesh, want_exact in [("grid", synth.grid_mesh(50, 40, 1), 1), ("torus", synth.torus_mesh(30, 20, 2), 1),
                                   ("fin", meshes.zoo()["fin_nonmanifold_edge"], 0)]:
        tr = orc.corner_tables(mesh)
        pos = mesh.attributes[0]
        cv = mesh.faces.ravel() if pos.point_to_value is None else pos.point_to_value[mesh.faces.ravel()]
        cv = np.ascontiguousarray(cv, np.uint32)
        opp = np.zeros(cv.size, np.uint32)
        exact = C.c_int()
        st = L.dxo_corner_table_opposites(cv.ctypes.data_as(C.POINTER(C.c_uint32)), mesh.faces.shape[0],
                                          opp.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(exact), -1)
        assert st == 0 and exact.value == want_exact, name
        if want_exact:
            assert np.array_equal(opp, tr.get("opposite", np.uint32)), name


def test_timing_record_counts_our_kernels():
    dxo.set_profiling(True)
    gpu_encode(synth.grid_mesh(64, 64, 2))
    t = dxo.last_timing()
    dxo.set_profiling(False)
    assert t["num_launches"] >= 20 and t["device_ms"] > 0
    names = {k["name"] for k in t["kernels"]}
    for k in ("K1_minmax", "K2_quantize", "K3_oct_quantize", "K4_predict_parallelogram", "K5_predict_normal",
              "K6_predict_texcoord", "K8_histogram", "K9_build_table", "K10_rans_encode"):
        assert k in names


@pytest.mark.parametrize("case", ["uniform_small", "geometric", "sparse_large_alphabet", "single_symbol", "two_symbols", "all_zero_but_one"])
def test_encode_symbols_entry(orc, case):
    """Entropy stage alone (encode_symbols, symbol_coding.rs:17-55): bytes equal the oracle's,
    covering the table-normalisation corner cases (deficit, excess, zero runs > 64)."""
    rng = np.random.default_rng(11)
    if case == "uniform_small":
        sym = rng.integers(0, 23, 5000)
    elif case == "geometric":
        sym = np.minimum(rng.geometric(0.02, 200000) - 1, 4000)
    elif case == "sparse_large_alphabet":
        sym = rng.choice(np.array([0, 1, 70, 200, 201, 1000, 5000, 70000]), 30000, p=[.5, .2, .1, .05, .05, .05, .04, .01])
    elif case == "single_symbol":
        sym = np.full(1000, 7)
    elif case == "two_symbols":
        sym = np.array([0, 300] * 50 + [300])
    else:
        sym = np.zeros(100000, np.int64)
        sym[777] = 9
    sym = sym.astype(np.uint32)
    got = dxo.encode_symbols(sym)
    assert got == orc.encode_symbols(sym)
    dec, used = orc.decode_symbols(got, sym.size)
    assert used == len(got) and np.array_equal(dec, sym)


# ---- device connectivity passes (K12 half edges, K13 left-most corners, K14 seam tables): meshes above the
# 4096-face threshold on which they must hand over to the sequential host passes ---------------------------
def _points_of(mesh):
    return [a.values if a.point_to_value is None else a.values[a.point_to_value] for a in mesh.attributes]


def _rebuild(mesh, faces, points):
    atts = [dxo.Attribute.from_points(p, a.att_type, a.domain, a.parents, a.unique_id) for a, p in zip(mesh.attributes, points)]
    return dxo.Mesh(np.asarray(faces, np.uint32), atts)


def test_large_bowtie_vertex_falls_back_to_host_left_most(orc):
    """Two 50x50 grids welded at one point: that vertex has two fans, the reference splits it
    (corner_table/mod.rs:342-416). K13 must flag it; left-most corners and seam tables then run on the host."""
    g = synth.grid_mesh(50, 50, 31)
    n = g.num_points()
    pts = _points_of(g)
    shifted = [p.copy() for p in pts]
    shifted[0][:, 0] += 3.0
    faces_b = g.faces.astype(np.int64) + n
    faces_b[faces_b == n] = n - 1  # B's first point becomes A's last point
    faces = np.concatenate([g.faces.astype(np.int64), faces_b])
    m = meshes.drop_unused_points(_rebuild(g, faces, [np.concatenate([a, b]) for a, b in zip(pts, shifted)]))
    assert m.faces.shape[0] >= 4096
    assert_stage_parity(orc, m)


def test_large_fin_edge_falls_back_to_host_matching(orc):
    """A third triangle on an interior edge of a 60x60 grid: K12 reports 'not exact' and the sequential matcher,
    the non-manifold edge split and the host left-most pass decide (corner_table/mod.rs:149-340)."""
    g = synth.grid_mesh(60, 60, 32)
    n = g.num_points()
    pts = _points_of(g)
    a, b = int(g.faces[2000, 0]), int(g.faces[2000, 1])
    extra = [np.concatenate([p, p[a:a + 1] * 0.5 + p[b:b + 1] * 0.5 + (0.3 if i == 0 else 0.0)]).astype(p.dtype) for i, p in enumerate(pts)]
    extra[1][-1] = extra[1][-1] / np.linalg.norm(extra[1][-1])
    faces = np.concatenate([g.faces.astype(np.int64), [[b, a, n]]])
    m = _rebuild(g, faces, extra)
    assert m.faces.shape[0] >= 4096
    assert gpu_encode(m) == orc.encode(m)


def test_large_mesh_with_unused_vertex_is_reported():
    """get_unused_vertices panics in the reference (corner_table/mod.rs:105-108); K13 reports it for large meshes."""
    g = synth.grid_mesh(50, 50, 33, with_normals=False, with_uvs=False)
    v = g.attributes[0].values
    pos = np.concatenate([v[:100], [[9.0, 9.0, 9.0]], v[100:]]).astype(np.float32)  # point 100 is referenced by no face
    faces = (g.faces + (g.faces >= 100)).astype(np.uint32)
    bad = dxo.Mesh(faces, [dxo.Attribute(pos, 0, 0)])
    with pytest.raises(dxo.Err) as e:
        gpu_encode(bad)
    assert e.value.status == -10  # DXO_ERR_UNUSED_VERTICES
    assert gpu_encode(g)


def test_large_grid_with_uv_seam_line(orc):
    """80x80 grid whose texture coordinates jump along a column (an interior seam that ends inside the mesh on one
    side): attribute vertices are split along the seam by K14 exactly as recompute_vertices does."""
    g = synth.grid_mesh(80, 80, 34)
    pts = _points_of(g)
    ny = 80
    corner_uv = pts[2][g.faces.ravel()].copy()            # per-corner uv
    quad_col = (np.arange(g.faces.shape[0]) // 2) % (ny - 1)
    quad_row = (np.arange(g.faces.shape[0]) // 2) // (ny - 1)
    island = np.repeat((quad_col >= 40) & (quad_row < 60), 3)  # faces right of the cut, but only for the first 60 rows
    corner_uv[island] += np.float32(0.25)
    # de-index: one point per corner, then weld equal (position, normal, uv) tuples back through from_points
    faces = np.arange(g.faces.size, dtype=np.uint32).reshape(-1, 3)
    atts = [dxo.Attribute.from_points(pts[0][g.faces.ravel()], 0, 0),
            dxo.Attribute.from_points(pts[1][g.faces.ravel()], g.attributes[1].att_type, g.attributes[1].domain, g.attributes[1].parents, g.attributes[1].unique_id),
            dxo.Attribute.from_points(corner_uv, g.attributes[2].att_type, g.attributes[2].domain, g.attributes[2].parents, g.attributes[2].unique_id)]
    m = dxo.Mesh(faces, atts)
    assert m.faces.shape[0] >= 4096
    assert_stage_parity(orc, m)


# ---- K10 off its tuned operating point: the paths that the default parameters almost never take -------------------
_K10_CHILD = r"""
import sys, numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
import draco_oxide_b200 as dxo, orc
from draco_oxide_b200 import synth
orc.build()
m = synth.grid_mesh(120, 130, 77)
out = bytearray(); dxo.encode(m, out)
assert bytes(out) == orc.encode(m), "mesh stream differs"
rng = np.random.default_rng(5)
for n, gen in [(200_000, lambda: np.minimum(rng.geometric(0.08, 200_000) - 1, 4000)), (50_001, lambda: rng.integers(0, 3, 50_001)),
               (33, lambda: rng.integers(0, 900, 33)), (100_000, lambda: rng.integers(0, 70_000, 100_000)),
               (2_700_001, lambda: np.minimum(rng.geometric(0.05, 2_700_001) - 1, 2000))]:  # long enough for the chunk size to adapt
    sym = gen().astype(np.uint32)
    assert dxo.encode_symbols(sym) == orc.encode_symbols(sym), n
print("ok")
"""


@pytest.mark.parametrize("env", [
    {"DXO_RANS_CHUNK": "64", "DXO_RANS_WARMUP": "0"},       # no warm-up: almost every chunk misses in the chain and is run there
    {"DXO_RANS_CHUNK": "32", "DXO_RANS_WARMUP": "32"},      # smallest chunks, thousands of them
    {"DXO_RANS_CHUNK": "256", "DXO_RANS_WARMUP": "64", "DXO_RANS_FAULT": "1"},  # wrong entering states: the fix-up must repair
    {"DXO_RANS_CHUNK": "1048576", "DXO_RANS_WARMUP": "1024"},  # a single chunk: the sequential coder
    {"DXO_RANS_CHUNK": "512", "DXO_RANS_WARMUP": "0"},      # misses with 16 pieces per chunk: the chain publishes the checkpoints itself
    {"DXO_RANS_CHUNK": "512", "DXO_RANS_WARMUP": "64", "DXO_RANS_FAULT": "1"},  # wrong entering states with 16 pieces per chunk
    {"DXO_RANS_SUB": "1"},                                   # one piece per chunk (no checkpoints), adaptive chunk size
    {"DXO_RANS_CHUNK": "4096", "DXO_RANS_WARMUP": "512", "DXO_RANS_SUB": "4"},
    {"DXO_NO_HUGEPAGES": "1"},                               # plain pinned / pageable blocks
    {"DXO_RANS_LANES": "0"},                                 # encode pass by warp pairs instead of one thread per chunk
    {"DXO_RANS_LANES": "0", "DXO_RANS_CHUNK": "256", "DXO_RANS_WARMUP": "64", "DXO_RANS_FAULT": "1"},
])
def test_rans_paths_off_the_operating_point(env):
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    e = dict(os.environ); e.update(env); e["DXO_RANS_DEBUG"] = "1"
    r = subprocess.run([sys.executable, "-c", _K10_CHILD.format(root=root)], env=e, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout + r.stderr
    import re
    misses = sum(int(x) for x in re.findall(r"chain misses=(\d+)", r.stderr))
    fixups = sum(int(x) for x in re.findall(r"fixup=(\d+)", r.stderr))
    if env.get("DXO_RANS_WARMUP") == "0":
        assert misses > 100, "the chain's run-from-the-true-state path was not exercised"
    if "DXO_RANS_FAULT" in env:
        assert fixups > 100, "the sequential fix-up was not exercised"


def test_rans_stream_of_config3_size(orc):
    """12M symbols (a 10M-triangle mesh's position stream is 15M): more chunks than the GPU holds warp pairs, chunk size at
    its upper bound, ~47k encode pieces. Bit-exact against the sequential oracle coder, and deterministic."""
    rng = np.random.default_rng(11)
    sym = np.minimum(rng.geometric(0.03, 12_000_001) - 1, 3000).astype(np.uint32)
    ref = orc.encode_symbols(sym)
    assert dxo.encode_symbols(sym) == ref
    assert dxo.encode_symbols(sym) == ref


def test_resident_session_graph_replay(orc):
    """With DXO_FLAG_GRAPH_REPLAY a resident session replays the step as one CUDA graph from its second run on (all streams, flag copies, external
    events for the host coders). Streams must stay byte-identical, also after run_steps and from another thread."""
    import threading
    for m in (synth.grid_mesh(90, 80, 41), synth.torus_mesh(50, 40, 42)):
        ref = orc.encode(m)
        s = dxo.Session(m, dxo.Config(flags=dxo.Config.GRAPH_REPLAY))
        assert s.run() == ref          # direct launches
        assert s.run() == ref          # capture + first replay
        assert s.run() == ref          # replay
        s.run_steps(3)
        assert s.run() == ref
        got = []
        t = threading.Thread(target=lambda: got.append(s.run()))  # another thread: its own streams, the graph is re-captured
        t.start(); t.join()
        assert got == [ref]
        assert s.run() == ref
        s.close()


@pytest.mark.parametrize("seed", range(8))
def test_fuzz_large_irregular_meshes(orc, seed):
    """Meshes above the device-connectivity threshold that are far from the manifold fast path: random triangle soups
    (non-manifold edges and vertices everywhere), grids with randomly flipped faces (inconsistent orientation), welded
    points, duplicated attribute values. Whatever K12-K14 flag must come out byte-identical through the host passes."""
    rng = np.random.default_rng(1000 + seed)
    if seed % 2 == 0:  # random soup
        n_pts, n_faces = 2500, 5200
        faces = rng.integers(0, n_pts, (n_faces, 3))
        faces = faces[(faces[:, 0] != faces[:, 1]) & (faces[:, 1] != faces[:, 2]) & (faces[:, 0] != faces[:, 2])].astype(np.uint32)
        pos = rng.integers(0, 40, (n_pts, 3)).astype(np.float32) * 0.25       # many coinciding positions
        nrm = rng.normal(size=(n_pts, 3)).astype(np.float32)
        nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
        uv = rng.integers(0, 30, (n_pts, 2)).astype(np.float32) / 30
    else:  # grid with flipped faces and welded columns
        g = synth.grid_mesh(60, 50, 200 + seed)
        faces = g.faces.copy()
        flip = rng.random(faces.shape[0]) < 0.03
        faces[flip] = faces[flip][:, ::-1]
        pts = [a.values if a.point_to_value is None else a.values[a.point_to_value] for a in g.attributes]
        pos, nrm, uv = (p.copy() for p in pts)
        weld = rng.integers(0, pos.shape[0], 40)
        pos[weld] = pos[(weld + 1) % pos.shape[0]]                              # pairs of points share a position
    m = dxo.Mesh(faces, [dxo.Attribute.from_points(pos, 0, 0), dxo.Attribute.from_points(nrm, 1, 1, (0,), 1), dxo.Attribute.from_points(uv, 3, 1, (0,), 2)])
    m = meshes.drop_unused_points(m)
    assert m.faces.shape[0] >= 4096
    try:
        ref = orc.encode(m)
    except Exception as e:  # inputs the reference rejects must be rejected with the same status
        with pytest.raises(dxo.Err) as ge:
            gpu_encode(m)
        assert ge.value.status == getattr(e, "status", ge.value.status)
        return
    assert gpu_encode(m) == ref
