"""GPU mesh build: dxo_mesh_build / dxo_dedup_values against the oracle's restatement of
MeshBuilder (core/mesh/builder.rs) and Attribute::remove_duplicate_values (core/attribute/mod.rs:394-452)."""
import numpy as np
import pytest

import draco_oxide_b200 as dxo
from draco_oxide_b200 import AttributeDomain as Dom
from draco_oxide_b200 import AttributeType as Ty
from draco_oxide_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _built():
    import __graft_entry__ as g
    g.build()
    assert dxo.device_count() >= 1


def same_mesh(a, b):
    assert np.array_equal(a.faces, b.faces)
    assert len(a.attributes) == len(b.attributes)
    for x, y in zip(a.attributes, b.attributes):
        assert (int(x.att_type), int(x.domain), x.unique_id, tuple(x.parents)) == (int(y.att_type), int(y.domain), y.unique_id, tuple(y.parents))
        assert x.values.dtype == y.values.dtype and x.values.tobytes() == y.values.tobytes(), "unique values differ"
        assert (x.point_to_value is None) == (y.point_to_value is None)
        if x.point_to_value is not None:
            assert np.array_equal(x.point_to_value, y.point_to_value)


def soup_from(mesh, rng, extra_degenerate=0, extra_unused=0):
    """De-indexes a mesh into one point per corner (an OBJ-like soup), optionally with degenerate faces and
    points no face uses, so that every stage of MeshBuilder::build has work to do."""
    pts = [a.values if a.point_to_value is None else a.values[a.point_to_value] for a in mesh.attributes]
    corners = mesh.faces.ravel()
    per_point = [p[corners] for p in pts]
    faces = np.arange(corners.size, dtype=np.uint32).reshape(-1, 3)
    if extra_unused:
        per_point = [np.concatenate([p, p[:extra_unused] + (1 if p.dtype.kind == "f" else 0)]) for p in per_point]
    if extra_degenerate:
        pick = rng.integers(0, faces.shape[0], extra_degenerate)
        deg = faces[pick].copy()
        deg[:, 2] = deg[:, 1]
        faces = np.concatenate([faces[: faces.shape[0] // 2], deg, faces[faces.shape[0] // 2:]])
    atts = [(p, a.att_type, a.domain, tuple(a.parents)) for p, a in zip(per_point, mesh.attributes)]
    return faces, atts


def test_reference_builder_known_answer():
    """core/mesh/builder.rs test: a 2-triangle soup with repeated positions becomes 4 points."""
    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0]], np.float32)
    faces = np.array([[0, 1, 2], [3, 4, 5]], np.uint32)
    m = dxo.build_mesh(faces, [(pos, Ty.Position, Dom.Position, ())])
    assert m.num_points() == 4 and m.faces.tolist() == [[0, 1, 2], [1, 3, 2]]


@pytest.mark.parametrize("case", ["grid", "torus", "grid_degenerate_unused", "position_only"])
def test_mesh_build_matches_oracle(orc, case):
    rng = np.random.default_rng(11)
    base = {"grid": synth.grid_mesh(40, 33, 51), "torus": synth.torus_mesh(24, 17, 52), "grid_degenerate_unused": synth.grid_mesh(25, 25, 53),
            "position_only": synth.grid_mesh(30, 20, 54, with_normals=False, with_uvs=False)}[case]
    faces, atts = soup_from(base, rng, extra_degenerate=40 if "degenerate" in case else 0, extra_unused=25 if "unused" in case else 0)
    got, ref = dxo.build_mesh(faces, atts), orc.build_mesh(faces, atts)
    same_mesh(got, ref)
    out = bytearray(); dxo.encode(got, out)
    assert bytes(out) == orc.encode(ref)


def test_dedup_semantics_zero_nan_and_types(orc):
    """-0.0 == +0.0 (the first occurrence's bytes survive), a NaN equals nothing, integer and 64-bit types."""
    v = np.array([[0.0, 1.0], [-0.0, 1.0], [np.nan, 2.0], [np.nan, 2.0], [0.0, 1.0], [3.0, -0.0], [3.0, 0.0]], np.float32)
    m, first = dxo.dedup_values(v)
    assert m.tolist() == [0, 0, 1, 2, 0, 3, 3] and first.tolist() == [0, 2, 3, 5]
    ref_map, ref_unique, _ = orc.dedup_and_remove(v)
    assert m.tolist() == ref_map and first.size == ref_unique
    rng = np.random.default_rng(3)
    for dt in (np.uint8, np.int16, np.uint32, np.int64, np.float64):
        x = rng.integers(0, 7, (5000, 3)).astype(dt)
        m, first = dxo.dedup_values(x)
        _, idx, inv = np.unique(x, axis=0, return_index=True, return_inverse=True)
        order = np.argsort(idx)
        rank = np.empty_like(order); rank[order] = np.arange(order.size)
        assert np.array_equal(m, rank[inv.ravel()]) and np.array_equal(first, np.sort(idx))


def test_large_soup_config1(orc):
    """config 1 de-indexed (150 000 corner points -> 25 351 points), then encoded: same stream as the indexed mesh's."""
    base = synth.config1_mesh()
    faces, atts = soup_from(base, np.random.default_rng(1))
    got = dxo.build_mesh(faces, atts)
    same_mesh(got, orc.build_mesh(faces, atts))
    assert got.num_points() == base.num_points()
    out = bytearray(); dxo.encode(got, out)
    assert bytes(out) == orc.encode(got)


def test_build_errors():
    pos = np.zeros((3, 3), np.float32); pos[1, 0] = 1; pos[2, 1] = 1
    uv = np.zeros((3, 2), np.float32)
    faces = np.array([[0, 1, 2]], np.uint32)
    with pytest.raises(dxo.Err) as e:  # TextureCoordinate must depend on Position (builder.rs:94-111)
        dxo.build_mesh(faces, [(pos, Ty.Position, Dom.Position, ()), (uv, Ty.TextureCoordinate, Dom.Corner, ())])
    assert e.value.status == -1
    with pytest.raises(dxo.Err) as e:  # ragged attributes are not restated on this path
        dxo.build_mesh(faces, [(pos, Ty.Position, Dom.Position, ()), (np.zeros((5, 3), np.float32), Ty.Normal, Dom.Corner, ())])
    assert e.value.status == -2
    with pytest.raises(dxo.Err) as e:
        dxo.build_mesh(np.array([[0, 1, 7]], np.uint32), [(pos, Ty.Position, Dom.Position, ())])
    assert e.value.status == -1


def test_accessor_bounds_match_oracle(orc):
    """dxo_attribute_bounds (io/gltf/encode.rs:815-899) against the oracle: bit-exact floats, with and
    without a point map, NaNs, signed zeros, one point, no points, 1-4 components, config-2-sized input."""
    rng = np.random.default_rng(9)
    def same(values, pm=None):
        a, b = dxo.attribute_bounds(values, pm), orc.attribute_bounds(values, pm)
        if b[0] is None:
            assert a == (None, None)
            return
        assert a[0].tobytes() == b[0].tobytes() and a[1].tobytes() == b[1].tobytes(), (a, b)
    for ncomp in (1, 2, 3, 4):
        v = (rng.normal(size=(5000, ncomp)) * 100).astype(np.float32)
        same(v)
        same(v, rng.integers(0, 5000, 12345).astype(np.uint32))
        same(v, rng.integers(10, 20, 7).astype(np.uint32))
        v[rng.integers(0, 5000, 300), rng.integers(0, ncomp, 300)] = np.nan
        same(v)
    same(np.array([[3.5, -1.25, 7.0]], np.float32))
    same(np.zeros((0, 3), np.float32))
    same(np.array([[0.0], [-0.0], [0.0]], np.float32))
    same(np.array([[-0.0], [0.0]], np.float32))
    same(np.full((9, 2), np.nan, np.float32))
    same(np.array([[5.0, 6.0], [1.0, 2.0]], np.float32), np.array([0, 0, 0], np.uint32))
    m = synth.config2_mesh(600)
    pos = m.attributes[0]
    same(pos.values, pos.point_to_value)
    with pytest.raises(dxo.Err):
        dxo.attribute_bounds(np.zeros((3, 3), np.float32), np.array([0, 3], np.uint32))  # map entry out of range
