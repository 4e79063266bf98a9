"""Oracle self-consistency (CPU): golden regression, literal-vs-fast equivalence of the
Appendix C reformulations, and an independent decode of every attribute section."""
import hashlib

import numpy as np
import pytest

import drc_parse
import meshes
from draco_oxide_b200 import synth


@pytest.mark.parametrize("name", meshes.golden_names())
def test_oracle_matches_golden(orc, name):
    mesh, drc = meshes.load_golden(name)
    assert orc.encode(mesh) == drc


@pytest.mark.parametrize("name", sorted(meshes.zoo().keys()))
def test_literal_quadratic_loops_equal_rank_reformulation(orc, name):
    m = meshes.drop_unused_points(meshes.zoo()[name])
    assert orc.encode(m) == orc.encode(m, literal=True)


def _roundtrip(orc, mesh):
    drc, tr = orc.encode(mesh, trace=True)
    end = int(tr.get("connectivity_end", np.uint64)[0])
    n = len(mesh.attributes)
    seq_lens = [tr.get(f"att{i}.sequence", np.uint32).size for i in range(n)]
    parsed = drc_parse.parse_attributes(drc, end, seq_lens)
    for i, d in enumerate(parsed):
        assert np.array_equal(d["symbols"], tr.get(f"att{i}.symbols", np.uint32)), i
        if "side_bits" in d:
            assert np.array_equal(d["side_bits"], tr.get(f"att{i}.side_bits", np.uint8)), i
        if "wrap_min" in d:
            assert [d["wrap_min"], d["wrap_max"]] == tr.get(f"att{i}.wrap_minmax", np.int32).tolist()
    return drc


@pytest.mark.parametrize("name", sorted(meshes.zoo().keys()))
def test_attribute_sections_decode(orc, name):
    _roundtrip(orc, meshes.drop_unused_points(meshes.zoo()[name]))


def test_config1_stream_is_stable_and_decodes(orc):
    # BASELINE config 1: 50 000 triangles, 25 351 vertices
    m = synth.config1_mesh()
    assert m.faces.shape[0] == 50000 and m.num_points() == 25351
    drc = _roundtrip(orc, m)
    assert drc[:11] == b"DRACO\x02\x02\x01\x01\x00\x00"
    assert hashlib.sha256(drc).hexdigest() == open(meshes.GOLDEN + "/config1.sha256").read().strip()


def test_zero_normal_is_rejected(orc):
    import draco_oxide_b200 as dxo
    m = synth.grid_mesh(4, 4, 3)
    n = m.attributes[1]
    vals = n.values.copy()
    vals[0] = 0
    bad = dxo.Mesh(m.faces, [m.attributes[0], dxo.Attribute(vals, n.att_type, n.domain, n.parents, n.point_to_value, n.unique_id), m.attributes[2]])
    with pytest.raises(orc.OracleError) as e:
        orc.encode(bad)
    assert e.value.status == -9


def test_quantization_bit_sweep(orc):
    import draco_oxide_b200 as dxo
    m = synth.grid_mesh(20, 20, 4)
    sizes = [len(orc.encode(m, dxo.Config(position_bits=b))) for b in (8, 10, 11, 12, 14, 16)]
    assert sizes == sorted(sizes)  # more bits never shrinks the stream on this mesh


def test_config4_sample_matches_golden_hashes(orc):
    """A sample of BASELINE config 4's 4096 primitives (the smallest 24 of the first 256 plus two larger ones)
    against tests/golden/config4_hashes.txt; the GPU suite checks all 4096 at full size."""
    lines = open(meshes.GOLDEN + "/config4_hashes.txt").read().split()
    counts = synth.batch_vertex_counts()
    assert len(lines) == counts.size + 1
    pick = sorted(np.argsort(counts[:256])[:24].tolist() + [2, 255])
    for k in pick:
        m = synth.batch_mesh(k, int(counts[k]))
        assert hashlib.sha256(orc.encode(m)).hexdigest()[:16] == lines[k], k
