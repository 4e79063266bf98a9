"""ctypes binding of oracle/liborc.so — the CPU oracle (TEST INFRASTRUCTURE).
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this."""
import ctypes as C
import os
import subprocess

import numpy as np

import draco_oxide_b200 as dxo
from draco_oxide_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "liborc.so")
_lib = None


def build():
    """Compiles the oracle's C++ restatement (g++); no-op when up to date."""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        L = C.CDLL(LIB)
        L.orc_last_error.restype = C.c_char_p
        L.orc_encode.argtypes = [C.POINTER(_capi.dxo_mesh), C.POINTER(_capi.dxo_config), C.c_int,
                                 C.POINTER(_capi.dxo_bytes), C.POINTER(C.c_void_p)]
        L.orc_encode_timed.argtypes = [C.POINTER(_capi.dxo_mesh), C.POINTER(_capi.dxo_config), C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.orc_encode_timed.restype = C.c_double
        L.orc_corner_tables.argtypes = [C.POINTER(_capi.dxo_mesh), C.POINTER(C.c_void_p)]
        L.orc_trace_get.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
        L.orc_trace_free.argtypes = [C.c_void_p]
        L.orc_free_bytes.argtypes = [C.POINTER(_capi.dxo_bytes)]
        L.orc_mesh_from_obj.argtypes = [C.c_char_p, C.POINTER(C.c_int)]
        L.orc_mesh_from_obj.restype = C.c_void_p
        L.orc_mesh_build.argtypes = [C.POINTER(C.c_uint32), C.c_uint64, C.POINTER(_capi.dxo_attribute), C.c_uint32, C.POINTER(C.c_int)]
        L.orc_mesh_build.restype = C.c_void_p
        L.orc_mesh_view.argtypes = [C.c_void_p, C.POINTER(_capi.dxo_mesh)]
        L.orc_mesh_free.argtypes = [C.c_void_p]
        L.orc_dedup_and_remove.argtypes = [C.POINTER(C.c_float), C.c_uint64, C.c_uint32, C.POINTER(C.c_uint32), C.c_uint64,
                                           C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_int)]
        L.orc_attribute_bounds.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_uint64, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.orc_attribute_bounds.restype = C.c_uint64
        L.orc_leb128.argtypes = [C.c_uint64, C.POINTER(C.c_uint8), C.POINTER(C.c_uint64)]
        L.orc_bitwriter.argtypes = [C.c_int, C.POINTER(C.c_uint8), C.POINTER(C.c_uint64), C.c_uint64, C.POINTER(C.c_uint8), C.POINTER(C.c_uint64)]
        L.orc_rans_encode_raw.argtypes = [C.POINTER(C.c_uint64), C.c_uint64, C.c_uint32, C.POINTER(C.c_uint32), C.c_uint64, C.POINTER(_capi.dxo_bytes)]
        L.orc_rans_decode_raw.argtypes = [C.POINTER(C.c_uint64), C.c_uint64, C.c_uint32, C.POINTER(C.c_uint8), C.c_uint64, C.c_uint64, C.POINTER(C.c_uint32)]
        L.orc_rabs_encode.argtypes = [C.c_uint32, C.POINTER(C.c_uint8), C.c_uint64, C.POINTER(_capi.dxo_bytes)]
        L.orc_rabs_decode.argtypes = [C.c_uint32, C.POINTER(C.c_uint8), C.c_uint64, C.c_uint64, C.POINTER(C.c_uint8)]
        L.orc_encode_symbols.argtypes = [C.POINTER(C.c_uint32), C.c_uint64, C.POINTER(_capi.dxo_bytes)]
        L.orc_decode_symbols.argtypes = [C.POINTER(C.c_uint8), C.c_uint64, C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]
        L.orc_zero_prob.argtypes = [C.c_uint64, C.c_uint64, C.c_int]
        L.orc_decode_check.argtypes = [C.POINTER(_capi.dxo_mesh), C.POINTER(_capi.dxo_config), C.POINTER(C.c_uint8), C.c_uint64, C.POINTER(C.c_uint64),
                                       C.POINTER(C.c_double), C.c_uint32]
        L.orc_to_positive_i32.argtypes = [C.c_int32]
        L.orc_to_positive_i32.restype = C.c_int32
        L.orc_oct_quantize.argtypes = [C.c_float, C.c_float, C.c_float, C.POINTER(C.c_int32), C.POINTER(C.c_int)]
        L.orc_oct_transform.argtypes = [C.c_float, C.c_float, C.c_float, C.POINTER(C.c_float)]
        _lib = L
    return _lib


class OracleError(Exception):
    def __init__(self, status):
        super().__init__(f"oracle status {status}: {lib().orc_last_error().decode()}")
        self.status = status


def _take(b):
    out = C.string_at(b.data, b.len) if b.len else b""
    lib().orc_free_bytes(C.byref(b))
    return out


class Trace:
    def __init__(self, handle):
        self.h = handle

    def get(self, key, dtype):
        p, n = C.c_void_p(), C.c_uint64()
        if lib().orc_trace_get(self.h, key.encode(), C.byref(p), C.byref(n)) != 0:
            raise KeyError(key)
        return np.frombuffer(C.string_at(p, n.value), dtype=dtype).copy()

    def __del__(self):
        if self.h:
            lib().orc_trace_free(self.h)
            self.h = None


def _cfg(cfg):
    c = _capi.dxo_config()
    cfg = cfg or dxo.Config.default()
    c.abi_version, c.position_bits, c.texcoord_bits, c.generic_bits, c.device, c.flags = 1, cfg.position_bits, cfg.texcoord_bits, cfg.generic_bits, -1, 0
    return c


def encode(mesh, cfg=None, literal=False, trace=False):
    """Oracle encode(): returns bytes, or (bytes, Trace) with trace=True."""
    cm, out, th = mesh.as_c(), _capi.dxo_bytes(), C.c_void_p()
    cc = _cfg(cfg)
    st = lib().orc_encode(C.byref(cm), C.byref(cc), int(literal), C.byref(out), C.byref(th) if trace else None)
    if st != 0:
        raise OracleError(st)
    data = _take(out)
    return (data, Trace(th)) if trace else data


def encode_timed(mesh, reps=1, threads=1, cfg=None):
    cm, st = mesh.as_c(), C.c_int()
    cc = _cfg(cfg)
    secs = lib().orc_encode_timed(C.byref(cm), C.byref(cc), reps, threads, C.byref(st))
    if st.value != 0:
        raise OracleError(st.value)
    return secs


def corner_tables(mesh):
    cm, th = mesh.as_c(), C.c_void_p()
    st = lib().orc_corner_tables(C.byref(cm), C.byref(th))
    if st != 0:
        raise OracleError(st)
    return Trace(th)


def _mesh_from_handle(h):
    """Copies an oracle-built mesh into a draco_oxide_b200.Mesh."""
    v = _capi.dxo_mesh()
    lib().orc_mesh_view(h, C.byref(v))
    faces = np.ctypeslib.as_array(v.faces, shape=(v.num_faces * 3,)).copy().reshape(-1, 3) if v.num_faces else np.zeros((0, 3), np.uint32)
    atts = []
    np_of = {int(k): dt for dt, k in dxo.mesh._NP_TO_CT.items()}
    for i in range(v.num_attributes):
        a = v.attributes[i]
        dt = np.dtype(np_of[a.component_type])
        nbytes = a.num_unique_values * a.num_components * dt.itemsize
        vals = np.frombuffer(C.string_at(a.values, nbytes), dtype=dt).reshape(-1, a.num_components).copy()
        pmap = np.ctypeslib.as_array(a.point_to_value, shape=(a.num_points,)).copy() if a.point_to_value else None
        parents = [a.parent_ids[k] for k in range(a.num_parents)]
        atts.append(dxo.Attribute(vals, a.att_type, a.domain, parents, pmap, a.unique_id))
    lib().orc_mesh_free(h)
    return dxo.Mesh(faces, atts)


def load_obj(path):
    st = C.c_int()
    h = lib().orc_mesh_from_obj(path.encode(), C.byref(st))
    if st.value != 0:
        raise OracleError(st.value)
    return _mesh_from_handle(h)


def build_mesh(faces, atts):
    """atts: list of (per_point_values ndarray, AttributeType, AttributeDomain, parents).
    Runs the oracle's MeshBuilder (value dedup, point merge, degenerate/unused removal)."""
    faces = np.ascontiguousarray(faces, dtype=np.uint32).reshape(-1, 3)
    arr = (_capi.dxo_attribute * max(len(atts), 1))()
    keep = []
    for i, (vals, ty, dom, parents) in enumerate(atts):
        v = np.ascontiguousarray(vals)
        if v.ndim == 1:
            v = v.reshape(-1, 1)
        par = np.asarray(parents, dtype=np.uint32)
        keep += [v, par]
        d = arr[i]
        d.att_type, d.component_type, d.num_components, d.domain = int(ty), int(dxo.mesh._NP_TO_CT[v.dtype]), v.shape[1], int(dom)
        d.unique_id, d.num_parents, d.parent_ids = i, par.size, par.ctypes.data_as(C.POINTER(C.c_uint32))
        d.num_unique_values, d.values, d.num_points = v.shape[0], v.ctypes.data, v.shape[0]
    st = C.c_int()
    h = lib().orc_mesh_build(faces.ctypes.data_as(C.POINTER(C.c_uint32)), faces.shape[0], arr, len(atts), C.byref(st))
    if st.value != 0:
        raise OracleError(st.value)
    return _mesh_from_handle(h)


def dedup_and_remove(values, removed=()):
    """Attribute::from + Attribute::remove; returns (map list, num_unique, has_map)."""
    v = np.ascontiguousarray(values, dtype=np.float32)
    r = np.ascontiguousarray(removed, dtype=np.uint32)
    out = np.zeros(v.shape[0], np.uint32)
    n, u, hm = C.c_uint64(), C.c_uint64(), C.c_int()
    st = lib().orc_dedup_and_remove(v.ctypes.data_as(C.POINTER(C.c_float)), v.shape[0], v.shape[1],
                                    r.ctypes.data_as(C.POINTER(C.c_uint32)), r.size,
                                    out.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(n), C.byref(u), C.byref(hm))
    if st != 0:
        raise OracleError(st)
    return out[: n.value].tolist(), u.value, bool(hm.value)


def leb128(v):
    buf, n = (C.c_uint8 * 16)(), C.c_uint64()
    lib().orc_leb128(v, buf, C.byref(n))
    return bytes(buf[: n.value])


def bitwriter(msb_first, items):
    sizes = (C.c_uint8 * len(items))(*[s for s, _ in items])
    vals = (C.c_uint64 * len(items))(*[v for _, v in items])
    buf, n = (C.c_uint8 * (8 * len(items) + 8))(), C.c_uint64()
    lib().orc_bitwriter(int(msb_first), sizes, vals, len(items), buf, C.byref(n))
    return bytes(buf[: n.value])


def _u32(a):
    a = np.ascontiguousarray(a, dtype=np.uint32)
    return a, a.ctypes.data_as(C.POINTER(C.c_uint32))


def rans_encode_raw(freqs, precision, symbols):
    f = np.ascontiguousarray(freqs, dtype=np.uint64)
    s, sp = _u32(symbols)
    out = _capi.dxo_bytes()
    st = lib().orc_rans_encode_raw(f.ctypes.data_as(C.POINTER(C.c_uint64)), f.size, precision, sp, s.size, C.byref(out))
    if st != 0:
        raise OracleError(st)
    return _take(out)


def rans_decode_raw(freqs, precision, buf, n):
    f = np.ascontiguousarray(freqs, dtype=np.uint64)
    b = np.frombuffer(buf, dtype=np.uint8).copy()
    out = np.zeros(n, np.uint32)
    st = lib().orc_rans_decode_raw(f.ctypes.data_as(C.POINTER(C.c_uint64)), f.size, precision,
                                   b.ctypes.data_as(C.POINTER(C.c_uint8)), b.size, n, out.ctypes.data_as(C.POINTER(C.c_uint32)))
    if st != 0:
        raise OracleError(st)
    return out


def rabs_encode(zero_prob, bits):
    b = np.ascontiguousarray(bits, dtype=np.uint8)
    out = _capi.dxo_bytes()
    st = lib().orc_rabs_encode(zero_prob, b.ctypes.data_as(C.POINTER(C.c_uint8)), b.size, C.byref(out))
    if st != 0:
        raise OracleError(st)
    return _take(out)


def rabs_decode(zero_prob, buf, n):
    b = np.frombuffer(buf, dtype=np.uint8).copy()
    out = np.zeros(n, np.uint8)
    st = lib().orc_rabs_decode(zero_prob, b.ctypes.data_as(C.POINTER(C.c_uint8)), b.size, n, out.ctypes.data_as(C.POINTER(C.c_uint8)))
    if st != 0:
        raise OracleError(st)
    return out


def encode_symbols(symbols):
    s, sp = _u32(symbols)
    out = _capi.dxo_bytes()
    st = lib().orc_encode_symbols(sp, s.size, C.byref(out))
    if st != 0:
        raise OracleError(st)
    return _take(out)


def decode_symbols(buf, n):
    b = np.frombuffer(buf, dtype=np.uint8).copy()
    out, used = np.zeros(n, np.uint32), C.c_uint64()
    st = lib().orc_decode_symbols(b.ctypes.data_as(C.POINTER(C.c_uint8)), b.size, n, out.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(used))
    if st != 0:
        raise OracleError(st)
    return out, used.value


def attribute_bounds(values, point_to_value=None):
    """compute_vec3_bounds / compute_vec4_bounds of the reference's glTF writer: (min, max) or (None, None) without points."""
    v = np.ascontiguousarray(values, dtype=np.float32)
    if v.ndim == 1:
        v = v.reshape(-1, 1)
    pm = None if point_to_value is None else np.ascontiguousarray(point_to_value, dtype=np.uint32)
    mn, mx = (C.c_float * 4)(), (C.c_float * 4)()
    n = lib().orc_attribute_bounds(v.ctypes.data, v.shape[0], v.shape[1], None if pm is None else pm.ctypes.data,
                                   0 if pm is None else pm.shape[0], mn, mx)
    if n == 0:
        return None, None
    return np.array(mn[: v.shape[1]], np.float32), np.array(mx[: v.shape[1]], np.float32)


def decode_check(mesh, drc, cfg=None):
    """The inverse of the attribute path (oracle/orc_inverse.hpp): decodes `drc` causally against `mesh` and reports, per
    attribute, how many decoded values were compared with the quantised attribute and how many differ."""
    cm, cc = mesh.as_c(), _cfg(cfg)
    buf = np.frombuffer(drc, dtype=np.uint8).copy()
    counts = (C.c_uint64 * (3 + 5 * 16))()
    errors = (C.c_double * 32)()
    st = lib().orc_decode_check(C.byref(cm), C.byref(cc), buf.ctypes.data_as(C.POINTER(C.c_uint8)), buf.size, counts, errors, 16)
    if st != 0:
        raise OracleError(st)
    n = int(counts[0])
    rep = {"num_attributes": n, "prefix_mismatch_at": None if counts[1] == 2**64 - 1 else int(counts[1]), "consumed": int(counts[2]), "length": buf.size, "attributes": []}
    for i in range(min(n, 16)):
        c = counts[3 + 5 * i: 8 + 5 * i]
        rep["attributes"].append({"values_checked": int(c[0]), "mismatches": int(c[1]), "inconsistent_writes": int(c[2]), "not_invertible": int(c[3]),
                                  "unreferenced": int(c[4]), "max_abs_error": errors[2 * i], "error_bound": errors[2 * i + 1]})
    return rep


def assert_decodes(mesh, drc, cfg=None):
    """decode(drc) reproduces every quantised attribute value of `mesh` (and the dequantised floats within half a step)."""
    rep = decode_check(mesh, drc, cfg)
    assert rep["prefix_mismatch_at"] is None, f"header / connectivity bytes differ at {rep['prefix_mismatch_at']}"
    assert rep["consumed"] == rep["length"], (rep["consumed"], rep["length"])
    assert rep["num_attributes"] == len(mesh.attributes)
    for i, a in enumerate(rep["attributes"]):
        assert a["values_checked"] + a["unreferenced"] == mesh.attributes[i].num_unique_values, (i, a)
        assert a["mismatches"] == 0 and a["inconsistent_writes"] == 0, (i, a)
        assert a["max_abs_error"] <= a["error_bound"] or a["error_bound"] == 0, (i, a)
    return rep
