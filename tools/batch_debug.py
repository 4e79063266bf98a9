"""Debug helper: small batches through dxo_encode_batch with DXO_DEBUG=1, compared with the per-mesh path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import draco_oxide_b200 as dxo
from draco_oxide_b200 import synth
import meshes

zoo = meshes.zoo()
full = synth.grid_mesh(7, 9, 11)
g = synth.grid_mesh(12, 12, 3)
bad_faces = g.faces.copy(); bad_faces[3, 1] = g.num_points() + 7
cases = {
    "A full+pos_uv": [full, zoo["grid_pos_uv"]],
    "B full+pos_only": [full, zoo["grid_pos_only"]],
    "C full+fin": [full, zoo["fin_nonmanifold_edge"]],
    "D full+bad_faces": [full, dxo.Mesh(bad_faces, g.attributes)],
    "E full+custom": [full, zoo["custom_attribute"]],
    "F full+torus": [full, zoo["torus_small"]],
    "G full alone": [full],
}
for name, ms in cases.items():
    print("=====", name, flush=True)
    got, sts = dxo.encode_batch(ms, return_statuses=True)
    for m, b, st in zip(ms, got, sts):
        out = bytearray()
        try:
            dxo.encode(m, out)
        except dxo.Err as e:
            print("   per-mesh error", e.status, "batch status", st)
            continue
        print("   status", st, "equal" if bytes(out) == b else "DIFFERENT", len(b), flush=True)
