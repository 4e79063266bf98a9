"""CPU: the C-ABI library builds, loads and exports every symbol include/dxo.h declares;
without a CUDA device the product path fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

import draco_oxide_b200 as dxo
from draco_oxide_b200 import _capi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    return _capi.lib()


def test_every_declared_symbol_is_exported(built):
    header = open(os.path.join(ROOT, "include", "dxo.h")).read()
    declared = set(re.findall(r"\b(dxo_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations found"
    assert declared == set(_capi.EXPORTED), declared ^ set(_capi.EXPORTED)
    for name in declared:
        assert hasattr(built, name), name


def test_no_torch_or_cuda_types_in_signatures():
    header = open(os.path.join(ROOT, "include", "dxo.h")).read()
    code = re.sub(r"/\*.*?\*/", "", header, flags=re.S)  # strip comments
    for bad in ("torch", "at::", "cudaStream_t", "std::", "#include <cuda", "Tensor"):
        assert bad not in code


def test_default_config_matches_reference_constants(built):
    c = _capi.dxo_config()
    built.dxo_config_default(C.byref(c))
    # portabilization/mod.rs:116-142: positions 11 bits, texcoords 10 bits
    assert (c.abi_version, c.position_bits, c.texcoord_bits, c.generic_bits, c.device) == (1, 11, 10, 11, -1)


def test_strerror_covers_all_codes(built):
    for code in (0, -1, -2, -3, -4, -5, -6, -7, -8, -9, -10, -20, -21, -22, -99):
        assert built.dxo_strerror(code)


def test_argument_validation_without_device(built):
    out = _capi.dxo_bytes()
    assert built.dxo_encode(None, None, C.byref(out)) == -1          # null mesh
    m = synth.grid_mesh(3, 3, 1)
    bad = dxo.Mesh(m.faces, m.attributes[1:])                          # no position attribute first
    with pytest.raises(dxo.Err) as e:
        dxo.Session(bad, host_only=True)
    assert e.value.status == -2
    unused = dxo.Mesh(m.faces[:2], m.attributes)                       # unused vertices -> reference panics
    with pytest.raises(dxo.Err) as e:
        dxo.Session(unused, host_only=True)
    assert e.value.status == -10
    with pytest.raises(dxo.Err) as e:
        dxo.Session(m, dxo.Config(position_bits=31), host_only=True)
    assert e.value.status == -1


def test_product_fails_loudly_without_gpu(built):
    """No silent CPU path: encode() must raise DXO_ERR_NO_DEVICE when no GPU is present."""
    if dxo.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(dxo.Err) as e:
        dxo.encode(synth.grid_mesh(4, 4, 1), bytearray())
    assert e.value.status == -20
    s = dxo.Session(synth.grid_mesh(4, 4, 1), host_only=True)
    with pytest.raises(dxo.Err) as e:
        s.run()
    assert e.value.status == -20
    import numpy as np
    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    with pytest.raises(dxo.Err) as e:  # the mesh build has no host path either
        dxo.build_mesh(np.array([[0, 1, 2]], np.uint32), [(pos, 0, 0, ())])
    assert e.value.status == -20
    with pytest.raises(dxo.Err) as e:
        dxo.dedup_values(pos)
    assert e.value.status == -20
    with pytest.raises(dxo.Err) as e:
        dxo.attribute_bounds(pos)
    assert e.value.status == -20
