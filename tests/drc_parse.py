"""Independent reader for the attribute section of a .drc stream (test helper).
Walks the layout of SURVEY.md Appendix A.3-A.4 and decodes every entropy-coded part with
the oracle's DEcoders (oracle/orc_decode.hpp), so a stream is proven decodable and fully
consumed. Returns per-attribute symbols, side bits and metadata."""
import struct

import numpy as np

import orc


def _leb(buf, pos):
    v, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        v |= (b & 0x7F) << shift
        if not b & 0x80:
            return v, pos
        shift += 7


def parse_attributes(drc, connectivity_end, seq_lens):
    """seq_lens[i] = number of sequence elements of attribute i. Returns list of dicts."""
    buf = bytes(drc)
    pos = connectivity_end
    natt = buf[pos]
    pos += 1
    for i in range(natt):
        assert buf[pos] == (i - 1) % 256 and buf[pos + 2] == 0
        pos += 3
    heads = []
    for i in range(natt):
        one, att_type, comp_type, ncomp, zero, uid, dec = buf[pos:pos + 7]
        assert one == 1 and zero == 0
        heads.append(dict(att_type=att_type, comp_type=comp_type, ncomp=ncomp, unique_id=uid, decoder=dec))
        pos += 7
    out = []
    for i in range(natt):
        h = heads[i]
        scheme, transform, rans = buf[pos:pos + 3]
        pos += 3
        assert rans == 1
        nq = 2 if h["decoder"] == 3 else h["ncomp"]
        nsym = seq_lens[i] * nq
        symbols, used = orc.decode_symbols(buf[pos:], nsym)
        pos += used
        d = dict(h, scheme=scheme, transform=transform, symbols=symbols)

        def side_stream(pos, nbits, reversed_):
            p0 = buf[pos]
            size, pos = _leb(buf, pos + 1)
            bits = orc.rabs_decode(p0, buf[pos:pos + size], nbits)  # comes out in reverse write order
            return (bits if reversed_ else bits[::-1]), p0, pos + size

        def transform_info(pos):
            if transform == 1:
                d["wrap_min"], d["wrap_max"] = struct.unpack_from("<ii", buf, pos)
                return pos + 8
            if transform == 3:
                assert struct.unpack_from("<II", buf, pos) == (255, 127)
                return pos + 8
            return pos
        if scheme == 6:      # normal: transform info, flips
            pos = transform_info(pos)
            d["side_bits"], d["zero_prob"], pos = side_stream(pos, seq_lens[i], False)
        elif scheme == 5:    # texcoord: orientation stream, then transform info
            (n_or,) = struct.unpack_from("<I", buf, pos)
            bits, d["zero_prob"], pos = side_stream(pos + 4, n_or, False)
            # bits[k] = (o[k] == o[k+1]) with o[n] = True -> recover o from the end
            o, nxt = np.zeros(n_or, np.uint8), 1
            for k in range(n_or - 1, -1, -1):
                o[k] = nxt if bits[k] else 1 - nxt
                nxt = o[k]
            d["side_bits"] = o
            pos = transform_info(pos)
        else:
            pos = transform_info(pos)
        if h["decoder"] == 2:
            d["qmin"] = struct.unpack_from("<%df" % h["ncomp"], buf, pos)
            pos += 4 * h["ncomp"]
            (d["qrange"],) = struct.unpack_from("<f", buf, pos)
            d["bits"] = buf[pos + 4]
            pos += 5
        elif h["decoder"] == 3:
            assert buf[pos] == 8
            pos += 1
        out.append(d)
    assert pos == len(buf), f"stream not fully consumed: {pos} of {len(buf)}"
    return out
