"""GLB assembly around the batch entry (dxo_encode_glb; io/gltf/encode.rs:932-1097, :362-415). The container bookkeeping
needs no device when the streams are given (here: the oracle's); the GPU variant encodes the batch itself."""
import numpy as np
import pytest

import draco_oxide_b200 as dxo
import glb_reader
import meshes
from draco_oxide_b200 import synth


def _check_glb(glb, ms, streams):
    doc, bin_chunk = glb_reader.read_glb(glb)
    assert doc["asset"]["version"] == "2.0"
    assert doc["extensionsRequired"] == ["KHR_draco_mesh_compression"] and doc["extensionsUsed"] == ["KHR_draco_mesh_compression"]
    assert doc["buffers"] == [{"byteLength": len(bin_chunk)}] or doc["buffers"][0]["byteLength"] <= len(bin_chunk)
    prims = glb_reader.draco_primitives(doc, bin_chunk)
    assert len(prims) == len(ms) == len(doc["nodes"]) == len(doc["scenes"][0]["nodes"])
    for (prim, blob), m, s in zip(prims, ms, streams):
        assert len(blob) % 4 == 0 and len(blob) - len(s) < 4
        assert blob[: len(s)] == s and set(blob[len(s):]) <= {0}
        acc = doc["accessors"]
        idx = acc[prim["indices"]]
        assert idx == {"componentType": 5121, "count": 3 * m.faces.shape[0], "type": "SCALAR"} and "bufferView" not in idx
        ext = prim["extensions"]["KHR_draco_mesh_compression"]["attributes"]
        for k, a in enumerate(m.attributes):
            name = {dxo.AttributeType.Position: "POSITION", dxo.AttributeType.Normal: "NORMAL", dxo.AttributeType.TextureCoordinate: "TEXCOORD_0"}.get(a.att_type)
            if name is None:
                continue
            ac = acc[prim["attributes"][name]]
            assert ac["componentType"] == 5126 and ac["count"] == len(a) and ac["type"] == ("VEC2" if name == "TEXCOORD_0" else "VEC3")
            assert ext[name] == {"POSITION": 1, "NORMAL": 0}.get(name, k)   # the ids the reference's writer assigns (:1009-1016)
            if name == "POSITION":
                pts = a.values if a.point_to_value is None else a.values[a.point_to_value]
                assert np.array_equal(np.asarray(ac["min"], np.float32), pts.min(axis=0)) and np.array_equal(np.asarray(ac["max"], np.float32), pts.max(axis=0))
            else:
                assert "min" not in ac and "max" not in ac


def test_glb_container_from_given_streams(orc):
    ms = [synth.grid_mesh(9, 7, 3), synth.torus_mesh(8, 6, 4), meshes.zoo()["grid_pos_only"], synth.grid_mesh(5, 5, 6, with_normals=False)]
    streams = [orc.encode(m) for m in ms]
    glb = dxo.encode_glb(ms, streams=streams)
    _check_glb(glb, ms, streams)
    for (prim, blob), m in zip(glb_reader.draco_primitives(*glb_reader.read_glb(glb)), ms):
        orc.assert_decodes(m, blob[: len(orc.encode(m))])


def test_empty_batch_is_a_valid_glb():
    doc, bin_chunk = glb_reader.read_glb(dxo.encode_glb([], streams=[]))
    assert bin_chunk == b"" and "meshes" not in doc and doc["scenes"] == [{"nodes": []}]


@pytest.mark.gpu
def test_glb_from_the_batch_entry(orc):
    """File-to-file shape of config 4: primitives in, one GLB out (encoded by dxo_encode_batch inside the call)."""
    counts = synth.batch_vertex_counts(40, 300, 20000, seed=11)
    ms = [synth.batch_mesh(k, int(c)) for k, c in enumerate(counts)]
    glb = dxo.encode_glb(ms)
    _check_glb(glb, ms, [orc.encode(m) for m in ms])
