"""Generates tests/golden/*.npz from the reference's own fixtures.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py

For each OBJ fixture under /root/reference/draco-oxide/tests/data the oracle's OBJ
reader + MeshBuilder restatement produces the Mesh the reference would hand to
encode(); the arrays are stored (not the OBJ text) together with the oracle's .drc
bytes. NOTE: the .drc bytes come from the oracle, not from the reference binary
(no Rust toolchain here): they pin regressions, not reference parity.
"""
import glob
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import orc  # noqa: E402

REF = "/root/reference/draco-oxide/tests/data"


def main():
    for path in sorted(glob.glob(os.path.join(REF, "*.obj"))):
        name = os.path.splitext(os.path.basename(path))[0]
        m = orc.load_obj(path)
        drc = orc.encode(m)
        assert drc == orc.encode(m, literal=True), name  # O(V^2) literal loops agree
        d = {"faces": m.faces, "drc": np.frombuffer(drc, np.uint8), "num_attributes": np.int64(len(m.attributes))}
        for i, a in enumerate(m.attributes):
            d[f"a{i}_values"] = a.values
            d[f"a{i}_meta"] = np.asarray([int(a.att_type), int(a.domain), a.unique_id] + a.parents, np.int64)
            if a.point_to_value is not None:
                d[f"a{i}_map"] = a.point_to_value
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
        print(name, m.faces.shape[0], "faces", len(drc), "bytes", hashlib.sha256(drc).hexdigest()[:16])


if __name__ == "__main__":
    main()
