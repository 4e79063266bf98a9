"""Round trip through the INVERSE of the attribute path (oracle/orc_inverse.hpp, CPU): the oracle's streams are decoded
causally — element i predicted only from values decoded before it, decoded positions feeding the normal / texture-coordinate
predictors, flips and orientations taken from the side streams — and must reproduce every quantised attribute value; the
dequantised floats must lie within half a quantisation step of the originals. This pins the whole-stream semantics of the
attribute sections independently of the encoder restatement (the reference ships no working decoder, see the header of
orc_inverse.hpp). The GPU suite runs the same check on the GPU's streams (tests/test_gpu_roundtrip.py)."""
import numpy as np
import pytest

import draco_oxide_b200 as dxo
import meshes
from draco_oxide_b200 import synth


@pytest.mark.parametrize("name", sorted(meshes.zoo().keys()))
def test_zoo_round_trip(orc, name):
    m = meshes.drop_unused_points(meshes.zoo()[name])
    rep = orc.assert_decodes(m, orc.encode(m))
    assert sum(a["values_checked"] for a in rep["attributes"]) > 0


@pytest.mark.parametrize("name", meshes.golden_names())
def test_reference_fixtures_round_trip(orc, name):
    mesh, drc = meshes.load_golden(name)
    orc.assert_decodes(mesh, drc)


def test_config1_round_trip(orc):
    m = synth.config1_mesh()
    rep = orc.assert_decodes(m, orc.encode(m))
    assert [a["values_checked"] for a in rep["attributes"]] == [a.num_unique_values for a in m.attributes]
    # the reference's octahedral transform is not injective (orc_inverse.hpp): a small share of the normals has more than
    # one preimage; everything else decodes exactly
    assert rep["attributes"][1]["not_invertible"] < 0.03 * rep["attributes"][1]["values_checked"]
    assert rep["attributes"][0]["not_invertible"] == rep["attributes"][2]["not_invertible"] == 0


@pytest.mark.parametrize("bits", [8, 12, 16])
def test_round_trip_other_quantisations(orc, bits):
    m = synth.torus_mesh(40, 30, 9)
    cfg = dxo.Config(position_bits=bits, texcoord_bits=min(bits, 12))
    orc.assert_decodes(m, orc.encode(m, cfg), cfg)


def test_tampered_streams_are_caught(orc):
    """The checker is not vacuous: a flipped payload bit, a wrong symbol order or a stream of another mesh fail."""
    m = synth.grid_mesh(30, 20, 5)
    drc = orc.encode(m)
    orc.assert_decodes(m, drc)
    hits = 0
    for pos in range(len(drc) // 3, len(drc) - 8, 97):
        bad = bytearray(drc)
        bad[pos] ^= 0x04
        try:
            rep = orc.decode_check(m, bytes(bad))
            ok = rep["prefix_mismatch_at"] is None and rep["consumed"] == rep["length"] and all(a["mismatches"] == 0 for a in rep["attributes"])
        except orc.OracleError:
            ok = False
        hits += not ok
    assert hits >= 0.9 * len(range(len(drc) // 3, len(drc) - 8, 97)), hits  # a flip inside float metadata can go unnoticed by the integer checks
    other = synth.grid_mesh(30, 20, 6)
    rep = orc.decode_check(other, drc)
    assert rep["prefix_mismatch_at"] is not None or any(a["mismatches"] for a in rep["attributes"])


def test_octahedral_inverse_is_exact_where_the_transform_is_injective(orc):
    """Normals that avoid a zero centred component (where the reference's diamond inversion collapses) decode exactly."""
    rng = np.random.default_rng(3)
    g = synth.grid_mesh(25, 25, 2)
    n = rng.normal(size=(g.num_points(), 3)).astype(np.float32)
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    att = dxo.Attribute.from_points(n, dxo.AttributeType.Normal, dxo.AttributeDomain.Corner, (0,), 1)
    m = dxo.Mesh(g.faces, [g.attributes[0], att, g.attributes[2]])
    rep = orc.assert_decodes(m, orc.encode(m))
    assert rep["attributes"][1]["not_invertible"] <= 0.02 * rep["attributes"][1]["values_checked"]
