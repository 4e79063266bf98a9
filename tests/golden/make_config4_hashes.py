"""Generates tests/golden/config4_hashes.txt: one line per primitive of BASELINE config 4 (4096 synthetic glTF
primitives, 1k-100k vertices) = the first 16 hex digits of sha256(oracle .drc stream), plus a last line
with the sha256 over all the full digests (a checksum of checksums).

    python tests/golden/make_config4_hashes.py        # ~1 min of mesh generation + ~2 min of oracle on 8 threads

The bytes come from the ORACLE (no Rust toolchain here): they pin the GPU path and the oracle against regressions
at full config-4 size; reference parity of the oracle itself is what tests/test_oracle_*.py pin.
"""
import hashlib
import os
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import orc  # noqa: E402
from draco_oxide_b200 import synth  # noqa: E402


def main():
    orc.build()
    counts = synth.batch_vertex_counts()
    digests = []
    with ThreadPoolExecutor(os.cpu_count() or 1) as ex:
        for lo in range(0, counts.size, 512):
            ms = synth.batch_meshes(counts[lo:lo + 512], first=lo)
            digests += list(ex.map(lambda m: hashlib.sha256(orc.encode(m)).hexdigest(), ms))
            print(lo + len(ms), "primitives", flush=True)
    total = hashlib.sha256("".join(digests).encode()).hexdigest()
    with open(os.path.join(HERE, "config4_hashes.txt"), "w") as f:
        f.write("".join(d[:16] + "\n" for d in digests))
        f.write(total + "\n")
    print("checksum of checksums", total)


if __name__ == "__main__":
    main()
