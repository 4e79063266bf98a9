"""draco-oxide_b200 — B200-native drop-in for draco-oxide's attribute-encoding hot path.

Host-side mirror of the reference's public interface for this path:

    encode(mesh, writer, Config.default())        encode/mod.rs:59-97
    Config                                         encode/mod.rs:22-43
    Err                                            encode/mod.rs:44-56

Everything below the boundary runs in csrc/libdxo_b200.so (C++ host + CUDA
sm_100a kernels) through the C ABI in include/dxo.h. There is no CPU fallback:
calls raise when the library or a CUDA device is missing.
"""
import ctypes as C

import numpy as np

from . import _capi
from .mesh import Attribute, AttributeDomain, AttributeType, ComponentDataType, Mesh  # noqa: F401


class Err(Exception):
    """encode::Err — carries the dxo_status code of the leaf error."""

    def __init__(self, status, message):
        super().__init__(f"{message} (dxo status {status})")
        self.status = status


class Config:
    """encode::Config: only `default()` exists in the reference. Quantization bit
    counts are exposed for the qp sweep; only 11/10 is reference behaviour."""

    GRAPH_REPLAY = 1  # DXO_FLAG_GRAPH_REPLAY

    def __init__(self, position_bits=11, texcoord_bits=10, generic_bits=11, device=-1, flags=0):
        self.position_bits = position_bits
        self.texcoord_bits = texcoord_bits
        self.generic_bits = generic_bits
        self.device = device
        self.flags = flags

    @classmethod
    def default(cls):
        return cls()

    def as_c(self):
        c = _capi.dxo_config()
        _capi.lib().dxo_config_default(C.byref(c))
        c.position_bits = self.position_bits
        c.texcoord_bits = self.texcoord_bits
        c.generic_bits = self.generic_bits
        c.device = self.device
        c.flags = self.flags
        return c


def _check(status):
    if status != 0:
        raise Err(status, _capi.lib().dxo_strerror(status).decode())


def _take(b):
    out = C.string_at(b.data, b.len) if b.len else b""
    _capi.lib().dxo_free_bytes(C.byref(b))
    return out


def encode(mesh, writer, cfg=None):
    """encode::encode(mesh, &mut writer, cfg). `writer` is any object with
    `extend(bytes)` (bytearray, list) or `write(bytes)`; bytes are appended."""
    cfg = cfg or Config.default()
    L = _capi.lib()
    cm, cc, out = mesh.as_c(), cfg.as_c(), _capi.dxo_bytes()
    _check(L.dxo_encode(C.byref(cm), C.byref(cc), C.byref(out)))
    _give(out, writer)


def _give(b, writer):
    """Appends a dxo_bytes result to the writer with ONE copy (straight out of the library's buffer), then frees it."""
    try:
        if b.len:
            view = (C.c_char * b.len).from_address(C.addressof(b.data.contents))
            if isinstance(writer, bytearray):
                writer += view
            elif hasattr(writer, "extend"):
                writer.extend(bytes(view))
            else:
                writer.write(memoryview(view))
    finally:
        _capi.lib().dxo_free_bytes(C.byref(b))


class Batch:
    """The `dxo_mesh[]` array of a list of meshes, marshalled once (ctypes structs over the meshes' numpy buffers, which it
    keeps alive). Building it costs ~70 us of Python per mesh — as much as encoding a small mesh — so callers that time
    or repeat `encode_batch` build it up front; the C ABI itself takes plain pointers."""

    def __init__(self, meshes):
        self.meshes = list(meshes)
        self._c = [m.as_c() for m in self.meshes]
        self.array = (_capi.dxo_mesh * max(len(self._c), 1))(*self._c)

    def __len__(self):
        return len(self.meshes)


def encode_batch(meshes, cfg=None, first_gpu=0, num_gpus=1, return_statuses=False):
    """The transcoder loop (io/gltf/encode.rs:941-953): one stream per mesh. `meshes`: a list of Mesh or a Batch. With
    return_statuses the call does not raise for per-mesh failures and returns (streams, statuses) instead (a failed mesh has
    an empty stream)."""
    cfg = cfg or Config.default()
    L = _capi.lib()
    batch = meshes if isinstance(meshes, Batch) else Batch(meshes)
    n = len(batch)
    arr = batch.array
    outs = (_capi.dxo_bytes * max(n, 1))()
    sts = (C.c_int * max(n, 1))()
    cc = cfg.as_c()
    st = L.dxo_encode_batch(arr, n, C.byref(cc), outs, sts, first_gpu, num_gpus)
    res = [_take(outs[i]) for i in range(n)]
    if return_statuses:
        return res, [int(sts[i]) for i in range(n)]
    _check(st)
    return res


def encode_glb(meshes, cfg=None, streams=None, first_gpu=0, num_gpus=1):
    """The transcoder's output side (io/gltf/encode.rs:932-1097, :362-415): one GLB holding every mesh as a
    KHR_draco_mesh_compression primitive. streams: already encoded .drc bytes per mesh, or None to encode here
    (one dxo_encode_batch call)."""
    cfg = cfg or Config.default()
    L = _capi.lib()
    n = len(meshes)
    cms = [m.as_c() for m in meshes]
    arr = (_capi.dxo_mesh * max(n, 1))(*cms)
    cc = cfg.as_c()
    out = _capi.dxo_bytes()
    sarr, keep = None, []
    if streams is not None:
        sarr = (_capi.dxo_bytes * max(n, 1))()
        for i, b in enumerate(streams):
            buf = (C.c_uint8 * max(len(b), 1)).from_buffer_copy(b if b else b"\0")
            keep.append(buf)
            sarr[i].data = C.cast(buf, C.POINTER(C.c_uint8))
            sarr[i].len = len(b)
    st = L.dxo_encode_glb(arr, n, C.byref(cc), sarr, C.byref(out), first_gpu, num_gpus)
    data = _take(out)
    _check(st)
    return data


class Session:
    """Resident session: connectivity done on the host once, every array the
    attribute kernels read kept in HBM. run() executes the device hot path."""

    def __init__(self, mesh, cfg=None, host_only=False):
        cfg = cfg or Config.default()
        self._mesh_c = mesh.as_c()
        self._h = C.c_void_p()
        cc = cfg.as_c()
        make = _capi.lib().dxo_connectivity_create if host_only else _capi.lib().dxo_session_create
        _check(make(C.byref(self._mesh_c), C.byref(cc), C.byref(self._h)))

    def run(self, want_bytes=True):
        out = _capi.dxo_bytes()
        _check(_capi.lib().dxo_session_run(self._h, C.byref(out) if want_bytes else None))
        return _take(out) if want_bytes else None

    def run_steps(self, steps):
        """Bench loop: returns (CUDA-event milliseconds for all steps, kernels launched)."""
        ms, n = C.c_float(), C.c_uint64()
        _check(_capi.lib().dxo_session_run_steps(self._h, steps, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def set_trace(self, on=True):
        _capi.lib().dxo_session_set_trace(self._h, int(on))

    def trace(self, key, dtype):
        import numpy as np
        p, n = C.c_void_p(), C.c_uint64()
        st = _capi.lib().dxo_session_trace_get(self._h, key.encode(), C.byref(p), C.byref(n))
        if st != 0:
            raise KeyError(key)
        return np.frombuffer(C.string_at(p, n.value), dtype=dtype).copy()

    def close(self):
        if self._h:
            _capi.lib().dxo_session_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def encode_symbols(symbols, device=-1, timing=False):
    """entropy::symbol_coding::encode_symbols(.., DirectCoded, ..) on the device."""
    import numpy as np
    sym = np.ascontiguousarray(symbols, dtype=np.uint32)
    out, ms = _capi.dxo_bytes(), (C.c_float * 3)()
    _check(_capi.lib().dxo_encode_symbols(sym.ctypes.data_as(C.POINTER(C.c_uint32)), sym.size, device, C.byref(out), ms))
    data = _take(out)
    return (data, {"histogram_ms": ms[0], "table_ms": ms[1], "rans_ms": ms[2]}) if timing else data


def dedup_values(values, device=-1):
    """Attribute::remove_duplicate_values on the device: returns (point -> unique value map, first index of each
    unique value), unique values numbered in first-occurrence order (core/attribute/mod.rs:394-452)."""
    import numpy as np
    from .mesh import _NP_TO_CT
    v = np.ascontiguousarray(values)
    if v.ndim == 1:
        v = v.reshape(-1, 1)
    n = v.shape[0]
    out_map, first, nu = np.zeros(max(n, 1), np.uint32), np.zeros(max(n, 1), np.uint32), C.c_uint64()
    _check(_capi.lib().dxo_dedup_values(v.ctypes.data, n, int(_NP_TO_CT[v.dtype]), v.shape[1], device,
                                        out_map.ctypes.data_as(C.POINTER(C.c_uint32)), first.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(nu)))
    return out_map[:n], first[: nu.value]


def attribute_bounds(values, point_to_value=None, device=-1):
    """Accessor min / max of the glTF writer (compute_vec3_bounds / compute_vec4_bounds, io/gltf/encode.rs:815-899) on the
    device: per-component bounds over the values of all points, NaNs skipped. Returns (min, max) float32 arrays, or
    (None, None) for an attribute without points (the reference returns empty vectors)."""
    import numpy as np
    v = np.ascontiguousarray(values, dtype=np.float32)
    if v.ndim == 1:
        v = v.reshape(-1, 1)
    pm = None if point_to_value is None else np.ascontiguousarray(point_to_value, dtype=np.uint32)
    npoints = v.shape[0] if pm is None else pm.shape[0]
    mn, mx = (C.c_float * 4)(), (C.c_float * 4)()
    _check(_capi.lib().dxo_attribute_bounds(v.ctypes.data, v.shape[0], v.shape[1], None if pm is None else pm.ctypes.data, npoints, device, mn, mx))
    if npoints == 0:
        return None, None
    return np.array(mn[: v.shape[1]], np.float32), np.array(mx[: v.shape[1]], np.float32)


def build_mesh(faces, atts, device=-1):
    """MeshBuilder: atts = list of (per_point_values ndarray, AttributeType, AttributeDomain, parents). Value dedup per
    attribute, position first, point merge, degenerate-face and unused-point removal (core/mesh/builder.rs:30-125);
    the duplicate searches run on the device. Returns a Mesh ready for encode()."""
    import numpy as np
    from .mesh import _NP_TO_CT, Attribute, Mesh
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
        d.att_type, d.component_type, d.num_components, d.domain = int(ty), int(_NP_TO_CT[v.dtype]), v.shape[1], int(dom)
        d.unique_id, d.num_parents, d.parent_ids = i, par.size, par.ctypes.data_as(C.POINTER(C.c_uint32))
        d.num_unique_values, d.values, d.num_points = v.shape[0], v.ctypes.data, v.shape[0]
    h = C.c_void_p()
    _check(_capi.lib().dxo_mesh_build(faces.ctypes.data_as(C.POINTER(C.c_uint32)), faces.shape[0], arr, len(atts), device, C.byref(h)))
    try:
        view = _capi.dxo_mesh()
        _check(_capi.lib().dxo_built_mesh_view(h, C.byref(view)))
        out_faces = np.ctypeslib.as_array(view.faces, shape=(view.num_faces * 3,)).copy().reshape(-1, 3) if view.num_faces else np.zeros((0, 3), np.uint32)
        out_atts = []
        ct_to_np = {int(ct): dt for dt, ct in _NP_TO_CT.items()}
        for i in range(view.num_attributes):
            a = view.attributes[i]
            dt = np.dtype(ct_to_np[a.component_type])
            nbytes = a.num_unique_values * a.num_components * dt.itemsize
            vals = np.frombuffer(C.string_at(a.values, nbytes), dtype=dt).reshape(-1, a.num_components).copy() if nbytes else np.zeros((0, a.num_components), dt)
            pmap = np.ctypeslib.as_array(a.point_to_value, shape=(a.num_points,)).copy() if a.point_to_value and a.num_points else None
            parents = tuple(a.parent_ids[k] for k in range(a.num_parents))
            out_atts.append(Attribute(vals, a.att_type, a.domain, parents, pmap, a.unique_id))
        return Mesh(out_faces, out_atts)
    finally:
        _capi.lib().dxo_built_mesh_free(h)


def set_profiling(on):
    _capi.lib().dxo_set_profiling(int(on))


def last_timing():
    t = _capi.dxo_timing()
    _check(_capi.lib().dxo_last_timing(C.byref(t)))
    return {
        "device_ms": t.device_ms, "host_connectivity_ms": t.host_connectivity_ms, "h2d_ms": t.h2d_ms,
        "d2h_ms": t.d2h_ms, "total_ms": t.total_ms, "h2d_bytes": t.h2d_bytes, "d2h_bytes": t.d2h_bytes,
        "num_launches": t.num_launches,
        "kernels": [{"name": t.kernels[i].name.decode(), "ms": t.kernels[i].ms,
                     "algorithmic_bytes": t.kernels[i].algorithmic_bytes} for i in range(t.num_kernels)],
    }


def encode_bits(bits, zero_prob, mode=1):
    """The binary rANS coder behind the side streams (host only): mode 0 = bit by bit, 1 = as the encoder runs it."""
    b = np.ascontiguousarray(bits, dtype=np.uint8)
    out = _capi.dxo_bytes()
    _check(_capi.lib().dxo_encode_bits(b.ctypes.data_as(C.POINTER(C.c_uint8)), b.size, int(zero_prob), int(mode), C.byref(out)))
    return _take(out)


def device_count():
    return _capi.lib().dxo_device_count()
