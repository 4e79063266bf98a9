"""ctypes view of include/dxo.h. The shared library is built in-tree by
`__graft_entry__.build()` (nvcc, sm_100a). Loading fails loudly when the library
is missing: there is no CPU fallback for the attribute path."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libdxo_b200.so")


class dxo_attribute(C.Structure):
    _fields_ = [
        ("att_type", C.c_uint32),
        ("component_type", C.c_uint32),
        ("num_components", C.c_uint32),
        ("domain", C.c_uint32),
        ("unique_id", C.c_uint32),
        ("num_parents", C.c_uint32),
        ("parent_ids", C.POINTER(C.c_uint32)),
        ("num_unique_values", C.c_uint64),
        ("values", C.c_void_p),
        ("num_points", C.c_uint64),
        ("point_to_value", C.POINTER(C.c_uint32)),
    ]


class dxo_mesh(C.Structure):
    _fields_ = [
        ("num_faces", C.c_uint64),
        ("faces", C.POINTER(C.c_uint32)),
        ("num_attributes", C.c_uint32),
        ("attributes", C.POINTER(dxo_attribute)),
    ]


class dxo_config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_uint32),
        ("position_bits", C.c_uint32),
        ("texcoord_bits", C.c_uint32),
        ("generic_bits", C.c_uint32),
        ("device", C.c_int32),
        ("flags", C.c_uint32),
    ]


class dxo_bytes(C.Structure):
    _fields_ = [("data", C.POINTER(C.c_uint8)), ("len", C.c_size_t)]


class dxo_kernel_time(C.Structure):
    _fields_ = [("name", C.c_char_p), ("ms", C.c_float), ("algorithmic_bytes", C.c_uint64)]


class dxo_timing(C.Structure):
    _fields_ = [
        ("device_ms", C.c_float),
        ("host_connectivity_ms", C.c_float),
        ("h2d_ms", C.c_float),
        ("d2h_ms", C.c_float),
        ("total_ms", C.c_float),
        ("h2d_bytes", C.c_uint64),
        ("d2h_bytes", C.c_uint64),
        ("num_launches", C.c_uint32),
        ("num_kernels", C.c_uint32),
        ("kernels", dxo_kernel_time * 64),
    ]


# every symbol include/dxo.h declares (checked by tests/test_capi_symbols.py)
EXPORTED = [
    "dxo_config_default", "dxo_encode", "dxo_encode_batch", "dxo_free_bytes", "dxo_strerror",
    "dxo_device_count", "dxo_session_create", "dxo_connectivity_create", "dxo_session_run_steps", "dxo_session_run", "dxo_session_destroy",
    "dxo_set_profiling", "dxo_last_timing", "dxo_session_set_trace", "dxo_session_trace_get",
    "dxo_corner_table_opposites", "dxo_encode_symbols", "dxo_encode_bits",
    "dxo_mesh_build", "dxo_built_mesh_view", "dxo_built_mesh_free", "dxo_dedup_values", "dxo_attribute_bounds", "dxo_encode_glb",
]

_lib = None


def lib():
    """Returns the loaded C-ABI library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback for this path."
        )
    L = C.CDLL(LIB_PATH)
    L.dxo_config_default.argtypes = [C.POINTER(dxo_config)]
    L.dxo_config_default.restype = None
    L.dxo_encode.argtypes = [C.POINTER(dxo_mesh), C.POINTER(dxo_config), C.POINTER(dxo_bytes)]
    L.dxo_encode.restype = C.c_int
    L.dxo_encode_batch.argtypes = [C.POINTER(dxo_mesh), C.c_size_t, C.POINTER(dxo_config), C.POINTER(dxo_bytes),
                                   C.POINTER(C.c_int), C.c_int, C.c_int]
    L.dxo_encode_batch.restype = C.c_int
    L.dxo_free_bytes.argtypes = [C.POINTER(dxo_bytes)]
    L.dxo_free_bytes.restype = None
    L.dxo_strerror.argtypes = [C.c_int]
    L.dxo_strerror.restype = C.c_char_p
    L.dxo_device_count.argtypes = []
    L.dxo_device_count.restype = C.c_int
    L.dxo_session_create.argtypes = [C.POINTER(dxo_mesh), C.POINTER(dxo_config), C.POINTER(C.c_void_p)]
    L.dxo_session_create.restype = C.c_int
    L.dxo_connectivity_create.argtypes = [C.POINTER(dxo_mesh), C.POINTER(dxo_config), C.POINTER(C.c_void_p)]
    L.dxo_connectivity_create.restype = C.c_int
    L.dxo_session_run.argtypes = [C.c_void_p, C.POINTER(dxo_bytes)]
    L.dxo_session_run.restype = C.c_int
    L.dxo_session_run_steps.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_float), C.POINTER(C.c_uint64)]
    L.dxo_session_run_steps.restype = C.c_int
    L.dxo_session_destroy.argtypes = [C.c_void_p]
    L.dxo_session_destroy.restype = None
    L.dxo_set_profiling.argtypes = [C.c_int]
    L.dxo_set_profiling.restype = None
    L.dxo_last_timing.argtypes = [C.POINTER(dxo_timing)]
    L.dxo_last_timing.restype = C.c_int
    L.dxo_session_set_trace.argtypes = [C.c_void_p, C.c_int]
    L.dxo_session_set_trace.restype = None
    L.dxo_session_trace_get.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    L.dxo_session_trace_get.restype = C.c_int
    L.dxo_corner_table_opposites.argtypes = [C.POINTER(C.c_uint32), C.c_uint64, C.POINTER(C.c_uint32),
                                             C.POINTER(C.c_int), C.c_int]
    L.dxo_corner_table_opposites.restype = C.c_int
    L.dxo_encode_symbols.argtypes = [C.POINTER(C.c_uint32), C.c_uint64, C.c_int, C.POINTER(dxo_bytes), C.POINTER(C.c_float)]
    L.dxo_encode_symbols.restype = C.c_int
    L.dxo_encode_bits.argtypes = [C.POINTER(C.c_uint8), C.c_uint64, C.c_uint8, C.c_int, C.POINTER(dxo_bytes)]
    L.dxo_encode_bits.restype = C.c_int
    L.dxo_mesh_build.argtypes = [C.POINTER(C.c_uint32), C.c_uint64, C.POINTER(dxo_attribute), C.c_uint32, C.c_int, C.POINTER(C.c_void_p)]
    L.dxo_mesh_build.restype = C.c_int
    L.dxo_built_mesh_view.argtypes = [C.c_void_p, C.POINTER(dxo_mesh)]
    L.dxo_built_mesh_view.restype = C.c_int
    L.dxo_built_mesh_free.argtypes = [C.c_void_p]
    L.dxo_built_mesh_free.restype = None
    L.dxo_dedup_values.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                   C.POINTER(C.c_uint64)]
    L.dxo_dedup_values.restype = C.c_int
    L.dxo_attribute_bounds.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_uint64, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.dxo_attribute_bounds.restype = C.c_int
    L.dxo_encode_glb.argtypes = [C.POINTER(dxo_mesh), C.c_size_t, C.POINTER(dxo_config), C.POINTER(dxo_bytes), C.POINTER(dxo_bytes), C.c_int, C.c_int]
    L.dxo_encode_glb.restype = C.c_int
    _lib = L
    return L
