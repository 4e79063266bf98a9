"""CPU: the product's host-side connectivity (flat-array corner tables, seam tables,
Edgebreaker, sequencer — draco-oxide_b200/csrc/connectivity.cpp) against the oracle's
independent restatement: same bytes, same tables, same traversal sequences."""
import numpy as np
import pytest

import draco_oxide_b200 as dxo
import meshes
from draco_oxide_b200 import synth


def _compare(orc, mesh):
    drc, tr = orc.encode(mesh, trace=True)
    s = dxo.Session(mesh, host_only=True)
    n = len(mesh.attributes)
    end = int(tr.get("connectivity_end", np.uint64)[0])
    head = s.trace("head_bytes", np.uint8).tobytes()
    assert head == drc[: end + 1 + 10 * n]  # header + connectivity + attribute section headers
    for key in ("opposite", "corner_to_vertex", "left_most", "corners_of_edgebreaker"):
        assert np.array_equal(s.trace(key, np.uint32), tr.get(key, np.uint32)), key
    assert s.trace("num_vertices", np.uint64)[0] == tr.get("num_vertices", np.uint64)[0]
    for i in range(n):
        assert np.array_equal(s.trace(f"att{i}.sequence", np.uint32), tr.get(f"att{i}.sequence", np.uint32)), i
        if i > 0:
            assert np.array_equal(s.trace(f"att{i}.c2v", np.uint32), tr.get(f"att{i}.c2v", np.uint32))
            assert np.array_equal(s.trace(f"att{i}.left_most", np.uint32), tr.get(f"att{i}.left_most", np.uint32))
            assert np.array_equal(s.trace(f"att{i}.seam", np.uint8), tr.get(f"att{i}.seam", np.uint8))
    s.close()


@pytest.mark.parametrize("name", sorted(meshes.zoo().keys()))
def test_zoo(orc, name):
    _compare(orc, meshes.drop_unused_points(meshes.zoo()[name]))


@pytest.mark.parametrize("name", meshes.golden_names())
def test_reference_fixtures(orc, name):
    _compare(orc, meshes.load_golden(name)[0])


def test_config1(orc):
    _compare(orc, synth.config1_mesh())


def test_large_mesh_twice_through_the_host_block_pool(orc):
    """Tables of a 400x300 grid (239k faces) are large enough to come from the pooled host blocks (common.hpp); the
    second pass reuses the blocks of the first and must not see stale contents. Two threads run at once as well."""
    import threading
    m = synth.grid_mesh(400, 300, 5)
    _compare(orc, m)
    _compare(orc, m)
    errs = []
    def run():
        try:
            _compare(orc, synth.torus_mesh(300, 200, 6))
        except Exception as e:  # noqa: BLE001
            errs.append(e)
    ts = [threading.Thread(target=run) for _ in range(2)]
    for t in ts: t.start()
    for t in ts: t.join()
    assert not errs, errs


def test_random_face_soups(orc):
    """Random index soups: non-manifold edges, flipped faces, shared vertices — the
    order-dependent paths of compute_table / handle_no_manifold_edges."""
    rng = np.random.default_rng(7)
    done = 0
    for trial in range(200):
        nv = int(rng.integers(4, 12))
        nf = int(rng.integers(2, 16))
        faces = rng.integers(0, nv, (nf, 3)).astype(np.uint32)
        faces = faces[(faces[:, 0] != faces[:, 1]) & (faces[:, 1] != faces[:, 2]) & (faces[:, 2] != faces[:, 0])]
        if faces.shape[0] == 0:
            continue
        pos = rng.random((nv, 3)).astype(np.float32)
        m = meshes.drop_unused_points(dxo.Mesh(faces, [dxo.Attribute.from_points(pos, 0, 0)]))
        try:
            orc.encode(m)
        except orc.OracleError:
            # inputs on which the reference panics: the product must refuse them too
            with pytest.raises(dxo.Err):
                dxo.Session(m, host_only=True)
            continue
        _compare(orc, m)
        done += 1
    assert done > 50


@pytest.mark.parametrize("seed", range(6))
def test_fuzz_irregular_meshes_host_passes(orc, seed):
    """Random triangle soups (non-manifold edges and vertices, many components, interior-start components) and grids with
    flipped faces and welded points through the host passes alone: corner tables, traversal, the seam streams derived from
    the traversal's symbols, the sequencers (which stop once every face is visited) — same bytes and tables as the oracle."""
    rng = np.random.default_rng(500 + seed)
    if seed % 2 == 0:
        n_pts, n_faces = 400 + 150 * seed, 900 + 300 * seed
        faces = rng.integers(0, n_pts, (n_faces, 3))
        faces = faces[(faces[:, 0] != faces[:, 1]) & (faces[:, 1] != faces[:, 2]) & (faces[:, 0] != faces[:, 2])].astype(np.uint32)
        pos = rng.integers(0, 25, (n_pts, 3)).astype(np.float32) * 0.25
        nrm = rng.normal(size=(n_pts, 3)).astype(np.float32)
        nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
        uv = rng.integers(0, 12, (n_pts, 2)).astype(np.float32) / 12
    else:
        g = synth.grid_mesh(30 + seed, 25, 300 + seed)
        faces = g.faces.copy()
        flip = rng.random(faces.shape[0]) < 0.04
        faces[flip] = faces[flip][:, ::-1]
        pts = [a.values if a.point_to_value is None else a.values[a.point_to_value] for a in g.attributes]
        pos, nrm, uv = (p.copy() for p in pts)
        weld = rng.integers(0, pos.shape[0], 25)
        pos[weld] = pos[(weld + 1) % pos.shape[0]]
        uv[rng.integers(0, uv.shape[0], 60)] += 0.37  # interior uv seams
    m = dxo.Mesh(faces, [dxo.Attribute.from_points(pos, 0, 0), dxo.Attribute.from_points(nrm, 1, 1, (0,), 1), dxo.Attribute.from_points(uv, 3, 1, (0,), 2)])
    m = meshes.drop_unused_points(m)
    try:
        orc.encode(m)
    except Exception:
        pytest.skip("the reference rejects this input (covered by the error-code tests)")
    _compare(orc, m)
