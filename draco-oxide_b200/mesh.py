"""Host-side mirror of the reference's data model for the encode() boundary.

Mirrors (names, argument meaning, error behaviour):
  AttributeType / AttributeDomain / ComponentDataType  core/attribute/mod.rs:527-716
  Attribute                                           core/attribute/mod.rs:26-49
  Mesh                                                core/mesh/mod.rs:13-23
Only the fields that cross the encode() boundary exist here; parsing, dedup and
mesh building are the caller's side — `Attribute.from_points`
provides the reference's first-occurrence value dedup for synthetic inputs.
"""
import ctypes as C
from enum import IntEnum

import numpy as np

from . import _capi


class AttributeType(IntEnum):
    Position = 0
    Normal = 1
    Color = 2
    TextureCoordinate = 3
    Custom = 4
    Tangent = 5
    Material = 6
    Joint = 7
    Weight = 8


class AttributeDomain(IntEnum):
    Position = 0
    Corner = 1


class ComponentDataType(IntEnum):
    U8 = 1
    I8 = 2
    U16 = 3
    I16 = 4
    U32 = 5
    I32 = 6
    U64 = 7
    I64 = 8
    F32 = 9
    F64 = 10


_NP_TO_CT = {
    np.dtype("uint8"): ComponentDataType.U8, np.dtype("int8"): ComponentDataType.I8,
    np.dtype("uint16"): ComponentDataType.U16, np.dtype("int16"): ComponentDataType.I16,
    np.dtype("uint32"): ComponentDataType.U32, np.dtype("int32"): ComponentDataType.I32,
    np.dtype("uint64"): ComponentDataType.U64, np.dtype("int64"): ComponentDataType.I64,
    np.dtype("float32"): ComponentDataType.F32, np.dtype("float64"): ComponentDataType.F64,
}


class Attribute:
    """Unique values (AoS, shape [U, N]) plus the optional point -> value map."""

    def __init__(self, values, att_type, domain, parents=(), point_to_value=None, unique_id=0):
        v = np.ascontiguousarray(values)
        if v.ndim == 1:
            v = v.reshape(-1, 1)
        if v.dtype not in _NP_TO_CT:
            raise TypeError(f"unsupported component dtype {v.dtype}")
        self.values = v
        self.att_type = AttributeType(att_type)
        self.domain = AttributeDomain(domain)
        self.parents = [int(p) for p in parents]
        self.point_to_value = None if point_to_value is None else np.ascontiguousarray(point_to_value, dtype=np.uint32)
        self.unique_id = int(unique_id)

    # Attribute::from — core/attribute/mod.rs:87-103 (+ remove_duplicate_values :394-452):
    # first occurrence wins, unique values keep first-occurrence order, == on components.
    @classmethod
    def from_points(cls, per_point_values, att_type, domain, parents=(), unique_id=0):
        v = np.ascontiguousarray(per_point_values)
        if v.ndim == 1:
            v = v.reshape(-1, 1)
        key = v + 0 if v.dtype.kind == "f" else v  # -0.0 + 0 == +0.0: float == semantics
        if v.dtype.kind == "f" and np.isnan(key).any():
            raise ValueError("NaN attribute values are not supported by from_points")
        key = np.ascontiguousarray(key)
        # Group rows by a 64-bit mix of their bytes (integer sort), then check that every row equals the first row of
        # its group: if so the hash groups ARE the value groups. Otherwise (a collision) compare the raw bytes.
        words = key.view(np.uint8).reshape(key.shape[0], -1).astype(np.uint64) if key.dtype.itemsize < 4 else \
            key.view(np.uint32).reshape(key.shape[0], -1).astype(np.uint64)
        h = np.full(key.shape[0], 0x9E3779B97F4A7C15, np.uint64)
        with np.errstate(over="ignore"):
            for c in range(words.shape[1]):
                h = (h ^ words[:, c]) * np.uint64(0xFF51AFD7ED558CCD)
                h ^= h >> np.uint64(29)
        hs = np.sort(h)
        if not (hs[1:] == hs[:-1]).any():  # distinct hashes: distinct rows
            return cls(v, att_type, domain, parents, None, unique_id)
        _, first, inv = np.unique(h, return_index=True, return_inverse=True)
        if not np.array_equal(key[first[inv]], key):
            kb = key.view(np.dtype((np.void, key.dtype.itemsize * key.shape[1]))).ravel()
            _, first, inv = np.unique(kb, return_index=True, return_inverse=True)
        order = np.argsort(first, kind="stable")            # unique ids sorted by first occurrence
        rank = np.empty_like(order)
        rank[order] = np.arange(order.size)
        mapping = rank[inv].astype(np.uint32)
        uniq = v[np.sort(first)]
        if uniq.shape[0] == v.shape[0]:
            return cls(v, att_type, domain, parents, None, unique_id)
        return cls(uniq, att_type, domain, parents, mapping, unique_id)

    @property
    def num_unique_values(self):
        return self.values.shape[0]

    def __len__(self):  # Attribute::len — core/attribute/mod.rs:195-202
        return self.values.shape[0] if self.point_to_value is None else self.point_to_value.shape[0]

    def get_component_type(self):
        return _NP_TO_CT[self.values.dtype]

    def get_num_components(self):
        return self.values.shape[1]


class Mesh:
    """faces: [F,3] point indices; attributes[0] must be the position attribute."""

    def __init__(self, faces, attributes):
        self.faces = np.ascontiguousarray(faces, dtype=np.uint32).reshape(-1, 3)
        self.attributes = list(attributes)

    def get_faces(self):
        return self.faces

    def get_attributes(self):
        return self.attributes

    def num_points(self):
        return len(self.attributes[0]) if self.attributes else 0

    # ---- ctypes view (keeps the numpy buffers alive through `self`) ----
    def as_c(self):
        n = len(self.attributes)
        arr = (_capi.dxo_attribute * max(n, 1))()
        keep = []
        for i, a in enumerate(self.attributes):
            d = arr[i]
            d.att_type = int(a.att_type)
            d.component_type = int(a.get_component_type())
            d.num_components = a.get_num_components()
            d.domain = int(a.domain)
            d.unique_id = a.unique_id
            par = np.asarray(a.parents, dtype=np.uint32)
            keep.append(par)
            d.num_parents = par.size
            d.parent_ids = par.ctypes.data_as(C.POINTER(C.c_uint32))
            d.num_unique_values = a.values.shape[0]
            d.values = a.values.ctypes.data
            d.num_points = len(a)
            d.point_to_value = (a.point_to_value.ctypes.data_as(C.POINTER(C.c_uint32))
                                if a.point_to_value is not None else C.POINTER(C.c_uint32)())
        m = _capi.dxo_mesh()
        m.num_faces = self.faces.shape[0]
        m.faces = self.faces.ctypes.data_as(C.POINTER(C.c_uint32))
        m.num_attributes = n
        m.attributes = arr
        m._keep = (arr, keep, self)
        return m
