"""Times dxo.build_mesh (device dedup) against the oracle's MeshBuilder on a de-indexed grid:
python tools/mesh_build_bench.py [grid side, default 1000 = config 2 as a 6M-point soup]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import draco_oxide_b200 as dxo
from draco_oxide_b200 import synth
import orc
orc.build()
side = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
base = synth.config2_mesh(side)
corners = base.faces.ravel()
atts = [((a.values if a.point_to_value is None else a.values[a.point_to_value])[corners], a.att_type, a.domain, tuple(a.parents)) for a in base.attributes]
faces = np.arange(corners.size, dtype=np.uint32).reshape(-1, 3)
print(f"soup: {corners.size} points, {faces.shape[0]} faces")
for r in range(3):
    t0 = time.perf_counter(); got = dxo.build_mesh(faces, atts); t1 = time.perf_counter()
    print(f"device build {1e3 * (t1 - t0):8.1f} ms -> {got.num_points()} points  ({corners.size / (t1 - t0) / 1e6:.1f} Mpoints/s in)")
t0 = time.perf_counter(); ref = orc.build_mesh(faces, atts); t1 = time.perf_counter()
print(f"oracle build {1e3 * (t1 - t0):8.1f} ms -> {ref.num_points()} points")
assert np.array_equal(got.faces, ref.faces) and all(x.values.tobytes() == y.values.tobytes() for x, y in zip(got.attributes, ref.attributes))
for dt, n in ((np.float32, 6_000_000),):
    v = atts[0][0]
    for r in range(3):
        t0 = time.perf_counter(); m, first = dxo.dedup_values(v); t1 = time.perf_counter()
        print(f"dedup_values {v.shape} f32: {1e3 * (t1 - t0):7.1f} ms -> {first.size} unique")
