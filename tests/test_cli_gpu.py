"""The product CLI (draco-oxide_b200/csrc/dxo_cli; cli/src/main.rs:33-73) against the oracle's CLI on OBJ files written
here: byte-identical .drc, and the --glb mode's primitives equal the per-file streams."""
import os
import subprocess

import numpy as np
import pytest

import draco_oxide_b200 as dxo
import glb_reader
from draco_oxide_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "draco-oxide_b200", "csrc", "dxo_cli")
ORC_CLI = os.path.join(ROOT, "oracle", "orc_cli")


def write_obj(path, mesh, shared_indices=True, quads=False):
    """OBJ text of a synthetic mesh: `f i/i/i` over the point arrays, or separate v / vt / vn pools with duplicates."""
    pts = [a.values if a.point_to_value is None else a.values[a.point_to_value] for a in mesh.attributes]
    with open(path, "w") as f:
        f.write("# written by tests/test_cli_gpu.py\n")
        for p in pts[0]:
            f.write("v %r %r %r\n" % tuple(float(x) for x in p))
        for n in pts[1]:
            f.write("vn %r %r %r\n" % tuple(float(x) for x in n))
        for t in pts[2]:
            f.write("vt %r %r\n" % tuple(float(x) for x in t))
        faces = mesh.faces.astype(np.int64) + 1
        if quads:  # pairs of triangles of a grid share a diagonal: emit them as quads (fan triangulation restores them)
            for a, b in zip(faces[0::2], faces[1::2]):
                q = [a[0], a[1], a[2], b[2]]
                f.write("f " + " ".join(f"{i}/{i}/{i}" for i in q) + "\n")
        else:
            for tri in faces:
                if shared_indices:
                    f.write("f " + " ".join(f"{i}/{i}/{i}" for i in tri) + "\n")
                else:  # negative (relative) indices
                    n = pts[0].shape[0]
                    f.write("f " + " ".join(f"{i - n - 1}/{i - n - 1}/{i - n - 1}" for i in tri) + "\n")


@pytest.fixture(scope="module", autouse=True)
def _built():
    import __graft_entry__ as g
    g.build()
    assert os.path.exists(CLI) and os.path.exists(ORC_CLI)


@pytest.mark.parametrize("case", ["grid", "torus_with_duplicate_points", "quads", "relative_indices"])
def test_obj_to_drc_matches_the_oracle_cli(tmp_path, case):
    mesh = {"grid": synth.grid_mesh(40, 30, 3), "torus_with_duplicate_points": synth.torus_mesh(24, 18, 4), "quads": synth.grid_mesh(12, 9, 5),
            "relative_indices": synth.grid_mesh(8, 8, 6)}[case]
    obj = str(tmp_path / "m.obj")
    write_obj(obj, mesh, shared_indices=case != "relative_indices", quads=case == "quads")
    got, want = str(tmp_path / "got.drc"), str(tmp_path / "want.drc")
    r = subprocess.run([CLI, "-i", obj, "-o", got], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([ORC_CLI, "-i", obj, "-o", want], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert open(got, "rb").read() == open(want, "rb").read()


def test_cli_rejects_wrong_extensions(tmp_path):
    obj = str(tmp_path / "m.obj")
    write_obj(obj, synth.grid_mesh(4, 4, 1))
    assert subprocess.run([CLI, "-i", obj, "-o", str(tmp_path / "x.bin")], capture_output=True).returncode == 1
    assert subprocess.run([CLI, "-i", str(tmp_path / "m.ply"), "-o", str(tmp_path / "x.drc")], capture_output=True).returncode == 1


def test_objs_to_glb(tmp_path):
    ms = [synth.grid_mesh(20, 15, 7), synth.torus_mesh(12, 10, 8), synth.grid_mesh(6, 31, 9)]
    paths, streams = [], []
    for k, m in enumerate(ms):
        p = str(tmp_path / f"p{k}.obj")
        write_obj(p, m)
        paths.append(p)
        drc = str(tmp_path / f"p{k}.drc")
        assert subprocess.run([CLI, "-i", p, "-o", drc], capture_output=True).returncode == 0
        streams.append(open(drc, "rb").read())
    glb = str(tmp_path / "scene.glb")
    r = subprocess.run([CLI, "--glb", "-o", glb] + paths, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    doc, bin_chunk = glb_reader.read_glb(open(glb, "rb").read())
    prims = glb_reader.draco_primitives(doc, bin_chunk)
    assert len(prims) == 3
    for (prim, blob), s in zip(prims, streams):
        assert blob[: len(s)] == s and len(blob) - len(s) < 4
