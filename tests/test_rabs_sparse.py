"""CPU: the table-driven binary rANS coder for long, heavily skewed side streams (draco-oxide_b200/csrc/rabs.cpp) against the
bit-by-bit coder (RabsCoder, encode/entropy/rans.rs:71-127 as restated in common.hpp): same bytes for every density, either
rare value, every probability it accepts, and for rare bits at the ends / next to each other. The oracle's streams pin the
bit-by-bit coder (test_host_connectivity, test_oracle_streams); the seam stream of the large torus there takes this path."""
import numpy as np
import pytest

import draco_oxide_b200 as dxo


def _prob(bits):
    z = np.float32((bits == 0).sum())
    p = int(np.float32(z / np.float32(bits.size)) * np.float32(256.0) + np.float32(0.5))
    return min(255, max(1, p))


@pytest.mark.parametrize("n", [16384, 20011, 300000])
@pytest.mark.parametrize("rare_is_one", [True, False])
def test_sparse_coder_matches_bit_by_bit(n, rare_is_one):
    rng = np.random.default_rng(n + int(rare_is_one))
    for density in (0.0, 1e-5, 1e-4, 1e-3, 3e-3, 8e-3, 0.015, 0.03):
        bits = (rng.random(n) < density).astype(np.uint8)
        if not rare_is_one:
            bits = (1 - bits).astype(np.uint8)
        forced = (255, 252, 248) if rare_is_one else (1, 4, 8)
        for p0 in {_prob(bits), *forced}:
            assert dxo.encode_bits(bits, p0, 0) == dxo.encode_bits(bits, p0, 1), (density, p0)


def test_sparse_coder_edge_positions():
    for n in (16384, 70001):
        bits = np.zeros(n, np.uint8)
        bits[[0, 1, n - 1]] = 1
        bits[n // 2: n // 2 + 40] = 1
        for p0 in (255, 254, 251, 249):
            assert dxo.encode_bits(bits, p0, 0) == dxo.encode_bits(bits, p0, 1)
        assert dxo.encode_bits(1 - bits, 2, 0) == dxo.encode_bits(1 - bits, 2, 1)


def test_sparse_coder_from_many_threads():
    """The trajectory tables are filled on demand under a lock and read without one."""
    import threading
    rng = np.random.default_rng(3)
    streams = [(rng.random(50000) < 0.002).astype(np.uint8) for _ in range(8)]
    want = [dxo.encode_bits(b, 253, 0) for b in streams]
    bad = []

    def run(k):
        for _ in range(5):
            if dxo.encode_bits(streams[k], 253, 1) != want[k]:
                bad.append(k)
    ts = [threading.Thread(target=run, args=(k,)) for k in range(8)]
    for t in ts: t.start()
    for t in ts: t.join()
    assert not bad


@pytest.mark.parametrize("p0,density", [(255, 0.002), (253, 0.012), (249, 0.02), (3, 0.99), (1, 0.9985)])
def test_sparse_coder_against_the_oracle_and_its_decoder(orc, p0, density):
    """Pins both host coders to the oracle's RabsCoder restatement (oracle/orc_core.hpp) and to the oracle's decoder, which
    returns the bits last-coded-first."""
    rng = np.random.default_rng(p0)
    bits = (rng.random(120000) < density).astype(np.uint8)
    want = orc.rabs_encode(p0, bits)
    assert dxo.encode_bits(bits, p0, 0) == want
    got = dxo.encode_bits(bits, p0, 1)
    assert got == want
    assert np.array_equal(orc.rabs_decode(p0, got, bits.size), bits[::-1])
