"""Pins the oracle (oracle/) against every known answer the REFERENCE's own tests hold
for the hot path. Each test cites the reference test it repeats.
Paths are relative to /root/reference/draco-oxide/src/. CPU only."""
import numpy as np
import pytest

import draco_oxide_b200 as dxo
from draco_oxide_b200 import AttributeDomain as Dom
from draco_oxide_b200 import AttributeType as Ty

# Tetrahedron fixture with the same geometry / index structure as the reference's
# tests/data/tetrahedron.obj (4 positions, 6 uvs, 4 normals, faces v/vt/vn).
TETRA_OBJ = """
v 1 1 1
v -1 -1 1
v -1 1 -1
v 1 -1 -1
vt 0.25 1.00
vt 0.00 0.50
vt 0.50 0.50
vt 0.75 1.00
vt 1.00 0.50
vt 0.25 0.50
vn 0.57735 0.57735 0.57735
vn -0.57735 -0.57735 0.57735
vn -0.57735 0.57735 -0.57735
vn 0.57735 -0.57735 -0.57735
f 1/1/1 2/2/2 3/3/3
f 1/1/1 4/4/4 2/2/2
f 1/1/1 3/3/3 4/5/4
f 2/2/2 4/6/4 3/3/3
"""


@pytest.fixture(scope="module")
def tetra(orc, tmp_path_factory):
    p = tmp_path_factory.mktemp("obj") / "tetrahedron.obj"
    p.write_text(TETRA_OBJ)
    return orc.load_obj(str(p))


def pos_mesh(faces, positions):
    pos = np.asarray(positions, np.float32)
    if pos.shape[1] == 2:
        pos = np.concatenate([pos, np.zeros((pos.shape[0], 1), np.float32)], axis=1)
    return dxo.Mesh(np.asarray(faces, np.uint32), [dxo.Attribute.from_points(pos, Ty.Position, Dom.Position)])


# ---- utils/bit_coder.rs:41-65 -------------------------------------------------------------
def test_leb128_known_bytes(orc):
    assert orc.leb128(300) == bytes([172, 2])


def test_leb128_roundtrip(orc):
    data = [0, 1, 127, 128, 255, 256, 1234567890, 0xFFFFFFFFFFFFFFFF]
    buf = b"".join(orc.leb128(v) for v in data)
    out, pos = [], 0
    while pos < len(buf):
        v, shift = 0, 0
        while True:
            b = buf[pos]
            pos += 1
            v |= (b & 0x7F) << shift
            if not b & 0x80:
                break
            shift += 7
        out.append(v)
    assert out == data


# ---- core/bit_coder.rs:514-627 ------------------------------------------------------------
def test_bitwriter_msb_first_known_bytes(orc):
    assert len(orc.bitwriter(True, [(2, 0b10), (3, 0b011)])) == 1
    assert len(orc.bitwriter(True, [(7, 0b0111010)])) == 1
    assert orc.bitwriter(True, [(8, 0b10111010)]) == bytes([0b10111010])
    assert orc.bitwriter(True, [(9, 0b110111011)]) == bytes([0b11011101, 0b10000000])
    b = orc.bitwriter(True, [(9, 0b101010100), (8, 0b10101110), (7, 0b0101010), (6, 0b111100), (5, 0b00001), (4, 0b1100)])
    assert len(b) == (9 + 8 + 7 + 6 + 5 + 4) // 8 + 1
    assert list(b[:5]) == [0b10101010, 0b01010111, 0b00101010, 0b11110000, 0b00111000]
    assert len(orc.bitwriter(True, [(11, 0b10111010110)])) == 2


def _read_lsb(buf, sizes):
    bits = "".join(f"{b:08b}"[::-1] for b in buf)  # LSB-first bit string
    out, pos = [], 0
    for s in sizes:
        out.append(int(bits[pos:pos + s][::-1], 2))
        pos += s
    return out


def test_bitwriter_lsb_first_roundtrip(orc):
    items = [(9, 0b101010100), (8, 0b10101010), (7, 0b0101010), (6, 0b111100), (5, 0b00001), (4, 0b1100)]
    b = orc.bitwriter(False, items)
    assert len(b) == (9 + 8 + 7 + 6 + 5 + 4) // 8 + 1
    assert _read_lsb(b, [s for s, _ in items]) == [v for _, v in items]
    b = orc.bitwriter(False, [(10, 0b1010101010)])
    assert len(b) == 2 and _read_lsb(b, [2] * 5) == [0b10] * 5


# ---- core/attribute/mod.rs:760-839 --------------------------------------------------------
def test_attribute_dedup_map(orc):
    pts = [[0, 0, 0], [1, 0, 0], [0.5, 1, 0], [0, 0, 0], [1, 0, 0], [2, 0, 0]]
    m, u, has = orc.dedup_and_remove(pts)
    assert (m, u, has) == ([0, 1, 2, 0, 1, 3], 4, True)
    # the numpy mirror used by the synthetic generators must agree
    a = dxo.Attribute.from_points(np.asarray(pts, np.float32), Ty.Position, Dom.Position)
    assert a.point_to_value.tolist() == m and a.num_unique_values == 4


def test_attribute_remove(orc):
    pts = [[0, 0, 0], [1, 0, 0], [2, 0, 0], [3, 0, 0], [2, 0, 0], [5, 0, 0]]
    assert orc.dedup_and_remove(pts) == ([0, 1, 2, 3, 2, 4], 5, True)
    assert orc.dedup_and_remove(pts, [2]) == ([0, 1, 3, 2, 4], 5, True)
    assert orc.dedup_and_remove(pts, [2, 1]) == ([0, 2, 1, 3], 4, True)


def test_dedup_float_equality_semantics(orc):
    # -0.0 == 0.0 merges (first occurrence's bytes win); Appendix B.16
    m, u, _ = orc.dedup_and_remove([[0.0, 1, 2], [-0.0, 1, 2], [3, 4, 5]])
    assert m == [0, 0, 1] and u == 2
    a = dxo.Attribute.from_points(np.asarray([[0.0, 1, 2], [-0.0, 1, 2], [3, 4, 5]], np.float32), Ty.Position, Dom.Position)
    assert a.point_to_value.tolist() == [0, 0, 1]


# ---- core/corner_table/mod.rs:539-671 -----------------------------------------------------
def test_corner_table_two_triangles(orc):
    m = pos_mesh([[0, 1, 2], [2, 1, 3]], [[0, 0], [1, 0], [0, 1], [1, 1]])
    t = orc.corner_tables(m)
    N = 0xFFFFFFFF
    assert t.get("opposite", np.uint32).tolist() == [5, N, N, N, N, 0]
    assert t.get("num_vertices", np.uint64)[0] == 4
    assert t.get("corner_to_vertex", np.uint32).tolist() == [0, 1, 2, 2, 1, 3]  # corner_to_vertex map empty


def test_corner_table_no_nonmanifold_vertices(orc):
    m = pos_mesh([[0, 1, 2], [1, 3, 2], [2, 3, 4], [2, 4, 5]],
                 [[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0], [0, .5, 0], [1, .5, 0]])
    t = orc.corner_tables(m)
    assert t.get("num_vertices", np.uint64)[0] == 6
    assert t.get("corner_to_vertex", np.uint32).tolist() == m.faces.ravel().tolist()


def test_corner_table_single_triangle_left_most(orc):
    t = orc.corner_tables(pos_mesh([[0, 1, 2]], [[0, 0], [1, 0], [0, 1]]))
    assert t.get("left_most", np.uint32).tolist() == [0, 1, 2]
    assert t.get("num_vertices", np.uint64)[0] == 3


def test_corner_table_non_manifold_vertex_split(orc):
    t = orc.corner_tables(pos_mesh([[0, 1, 2], [0, 3, 4]], [[0, 0], [1, 0], [0, 1], [-1, 1], [0, -1]]))
    assert t.get("num_vertices", np.uint64)[0] == 6  # vertex 0 is duplicated
    assert t.get("left_most", np.uint32).tolist() == [0, 1, 2, 4, 5, 3]


def test_contains_non_manifold_edges(orc):
    pts = [[0, 0], [1, 0], [0, 1], [1, 1], [2, 2]]
    t = orc.corner_tables(pos_mesh([[0, 1, 2], [1, 3, 2], [2, 1, 4]], pts))
    assert t.get("non_manifold_edges", np.uint32)[0] == 1
    t = orc.corner_tables(pos_mesh([[0, 1, 2], [1, 3, 2]], pts[:4]))
    assert t.get("non_manifold_edges", np.uint32)[0] == 0


# ---- io/obj/mod.rs:73-89, core/mesh/builder.rs:406-437 ------------------------------------
def test_obj_loader_tetrahedron(tetra):
    assert tetra.faces.tolist() == [[0, 1, 2], [0, 3, 1], [0, 2, 4], [1, 5, 2]]
    assert len(tetra.attributes) == 3
    p = tetra.attributes[0]
    assert p.att_type == Ty.Position and p.domain == Dom.Position and p.get_num_components() == 3
    assert p.num_unique_values == 4 and len(p) == 6


def test_mesh_builder_soup_to_tetrahedron(orc):
    faces = [[0, 1, 2], [3, 4, 5], [6, 7, 8], [9, 10, 11]]
    xs = [0, 1, 2, 0, 3, 1, 1, 3, 2, 0, 2, 3]
    pos = np.asarray([[x, 0, 0] for x in xs], np.float32)
    m = orc.build_mesh(faces, [(pos, Ty.Position, Dom.Position, [])])
    assert m.faces.shape[0] == 4 and len(m.attributes) == 1 and len(m.attributes[0]) == 4


# ---- core/corner_table/attribute_corner_table.rs:200-292 ----------------------------------
def test_attribute_corner_table_tetrahedron_uv_seams(orc, tetra):
    t = orc.corner_tables(tetra)
    nv = int(t.get("num_vertices", np.uint64)[0])
    assert t.get("att2.num_vertices", np.uint64)[0] == nv + 2
    assert t.get("att2.c2v", np.uint32)[0] == 0
    N = 0xFFFFFFFF
    sl, sr = t.get("att2.swing_left", np.uint32), t.get("att2.swing_right", np.uint32)
    for c in (4, 8, 10):
        assert sl[c] == N and sr[c] == N
    seam = t.get("att2.seam", np.uint8)
    for c in (3, 5, 6, 7, 9, 11):
        assert seam[c]
    lm = t.get("att2.left_most", np.uint32)
    assert lm.tolist() == [6, 5, 11, 10, 8, 4]
    for c in lm:
        assert sl[c] == N


def test_attribute_corner_table_closed_surface_without_seams(orc):
    # reference uses sphere.obj with per-vertex normals (no seams); same property on a
    # closed synthetic surface whose normals are a function of the position vertex
    from draco_oxide_b200 import synth
    m = synth.torus_mesh(8, 6, 5, with_uvs=False)
    t = orc.corner_tables(m)
    assert t.get("att1.num_vertices", np.uint64)[0] == t.get("num_vertices", np.uint64)[0]
    assert not t.get("att1.seam", np.uint8).any()
    assert not t.get("att1.vertex_on_seam", np.uint8).any()
    assert (t.get("att1.c2v", np.uint32) == t.get("corner_to_vertex", np.uint32)).all()


# ---- shared/attribute/sequence.rs:163-207 -------------------------------------------------
def test_sequencer_tetrahedron_orders(orc, tetra):
    _, tr = orc.encode(tetra, trace=True)
    pts = tetra.faces.ravel()
    assert [int(pts[c]) for c in tr.get("att0.sequence", np.uint32)] == [3, 1, 0, 2]
    assert [int(pts[c]) for c in tr.get("att1.sequence", np.uint32)] == [3, 1, 0, 2]
    assert [int(pts[c]) for c in tr.get("att2.sequence", np.uint32)] == [3, 1, 0, 2, 5, 4]


# ---- decode/entropy/rans.rs:219-279, decode/entropy/symbol_coding.rs:125-210 ---------------
def test_rans_coder_roundtrip(orc):
    n = 43
    data, freq, x = [], [0] * n, 3
    for _ in range(1 << 12):
        x = (x + 37) % n
        data.append(x)
        freq[x] += 1
    buf = orc.rans_encode_raw(freq, 12, data)
    assert orc.rans_decode_raw(freq, 12, buf, len(data)).tolist() == data[::-1]


def test_rabs_coder_roundtrip(orc):
    zeros, size = 100, 1 << 8
    srt = [0] * zeros + [1] * (size - zeros)
    data = [0] * size
    for i in range(size):
        data[(67 * i) % size] = srt[i]
    buf = orc.rabs_encode(zeros, data)
    assert orc.rabs_decode(zeros, buf, size).tolist() == data[::-1]


@pytest.mark.parametrize("length", [100, 300])
def test_encode_decode_symbols_direct_coded(orc, length):
    sym = [(x * x * x) % 23 for x in range(length)]
    buf = orc.encode_symbols(sym)
    out, used = orc.decode_symbols(buf, length)
    assert used == len(buf) and out.tolist() == sym


# ---- encode/attribute/prediction_transform/geom.rs:167-196 --------------------------------
def test_octahedral_transform_inverse(orc):
    import ctypes as C
    dirs = [[1, 0, 0], [0, 1, 0], [0, 0, 1], [-1, 0, 0], [0, -1, 0], [0, 0, -1], [1, 1, 1], [-1, -1, -1],
            [1, -1, 1], [-1, 1, -1], [1, 1, -1], [-1, -1, 1], [1, -1, -1]]
    for d in dirs:
        n = np.asarray(d, np.float64)
        n /= np.sqrt((n * n).sum())
        out = (C.c_float * 2)()
        orc.lib().orc_oct_transform(float(n[0]), float(n[1]), float(n[2]), out)
        u, v = float(out[0]), float(out[1])
        x, y, z = 1.0 - abs(u) - abs(v), u, v
        if abs(u) + abs(v) > 1.0:
            y = (1.0 - abs(v)) * (1.0 if u > 0 else -1.0)
            z = (1.0 - abs(u)) * (1.0 if v > 0 else -1.0)
        r = np.asarray([x, y, z])
        r /= np.sqrt((r * r).sum())
        assert ((n - r) ** 2).sum() < 1e-10


def test_accessor_bounds_restatement(orc):
    """compute_vec3_bounds (io/gltf/encode.rs:815-856): true per-point bounds starting from point 0 — unlike the
    quantizer's zero-initialised bounds (Appendix B.3) an all-positive attribute keeps a positive minimum; NaNs are
    skipped by f32::min / max; the point map selects which values count; no points -> empty."""
    v = np.array([[1, 2, 3], [4, 5, 6], [2, 9, 3.5]], np.float32)
    mn, mx = orc.attribute_bounds(v)
    assert mn.tolist() == [1, 2, 3] and mx.tolist() == [4, 9, 6]
    mn, mx = orc.attribute_bounds(v, [2, 2, 0])          # value 1 is referenced by no point
    assert mn.tolist() == [1, 2, 3] and mx.tolist() == [2, 9, 3.5]
    w = np.array([[np.nan, 1], [3, np.nan], [-2, np.nan]], np.float32)
    mn, mx = orc.attribute_bounds(w)
    assert mn.tolist() == [-2, 1] and mx.tolist() == [3, 1]
    mn, mx = orc.attribute_bounds(np.full((4, 1), np.nan, np.float32))
    assert np.isnan(mn[0]) and np.isnan(mx[0])
    assert orc.attribute_bounds(np.zeros((0, 3), np.float32)) == (None, None)
    rng = np.random.default_rng(3)
    r = rng.normal(size=(1000, 4)).astype(np.float32)
    mn, mx = orc.attribute_bounds(r)
    assert np.array_equal(mn, r.min(axis=0)) and np.array_equal(mx, r.max(axis=0))
