"""GPU streams through the inverse of the attribute path (oracle/orc_inverse.hpp): decode(encode_gpu(m)) reproduces the
quantised attributes of m. Independent of the byte comparison with the oracle's encoder: a stream that decodes to the
right values is semantically right even where no golden bytes exist."""
import pytest

import draco_oxide_b200 as dxo
import meshes
from draco_oxide_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _built():
    import __graft_entry__ as g
    g.build()
    assert dxo.device_count() >= 1, "GPU tests need a CUDA device"


def gpu_encode(mesh, cfg=None):
    out = bytearray()
    dxo.encode(mesh, out, cfg)
    return bytes(out)


@pytest.mark.parametrize("name", sorted(meshes.zoo().keys()))
def test_zoo_round_trip(orc, name):
    m = meshes.drop_unused_points(meshes.zoo()[name])
    orc.assert_decodes(m, gpu_encode(m))


def test_config2_round_trip(orc):
    m = synth.config2_mesh()
    rep = orc.assert_decodes(m, gpu_encode(m))
    assert rep["attributes"][0]["values_checked"] == 1_000_000


@pytest.mark.parametrize("bits", [8, 16])
def test_config5_round_trip(orc, bits):
    m = synth.config2_mesh(500)
    cfg = dxo.Config(position_bits=bits)
    orc.assert_decodes(m, gpu_encode(m, cfg), cfg)


def test_config3_round_trip(orc):
    m = synth.config3_mesh()
    rep = orc.assert_decodes(m, gpu_encode(m))
    assert rep["attributes"][0]["values_checked"] == 5_000_000


def test_config4_batch_round_trip(orc):
    """256 primitives of config 4 (every 16th of the size-sorted list) through the batch entry, each decoded."""
    import numpy as np
    counts = synth.batch_vertex_counts()
    pick = np.argsort(-counts, kind="stable")[::16]
    ms = [synth.batch_mesh(int(k), int(counts[k])) for k in pick]
    got = dxo.encode_batch(ms, first_gpu=0, num_gpus=1)
    for m, g in zip(ms, got):
        orc.assert_decodes(m, g)
