"""Minimal GLB reader for the tests (container per glTF 2.0: 12-byte header, JSON chunk, BIN chunk)."""
import json
import struct


def read_glb(data):
    magic, version, total = struct.unpack_from("<4sII", data, 0)
    assert magic == b"glTF" and version == 2 and total == len(data), (magic, version, total, len(data))
    pos, chunks = 12, []
    while pos < len(data):
        length, kind = struct.unpack_from("<I4s", data, pos)
        assert length % 4 == 0
        chunks.append((kind, data[pos + 8: pos + 8 + length]))
        pos += 8 + length
    assert pos == len(data)
    assert chunks[0][0] == b"JSON"
    doc = json.loads(chunks[0][1].decode())
    assert chunks[0][1].decode().rstrip(" ").endswith("}")  # padded with spaces only
    bin_chunk = chunks[1][1] if len(chunks) > 1 else b""
    if len(chunks) > 1:
        assert chunks[1][0] == b"BIN\0"
    return doc, bin_chunk


def draco_primitives(doc, bin_chunk):
    """[(primitive dict, compressed bytes incl. padding)] for every KHR_draco_mesh_compression primitive, in order."""
    out = []
    for mesh in doc.get("meshes", []):
        for prim in mesh["primitives"]:
            ext = prim["extensions"]["KHR_draco_mesh_compression"]
            view = doc["bufferViews"][ext["bufferView"]]
            assert view["buffer"] == 0
            out.append((prim, bin_chunk[view["byteOffset"]: view["byteOffset"] + view["byteLength"]]))
    return out
