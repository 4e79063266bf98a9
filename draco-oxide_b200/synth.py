"""Deterministic synthetic meshes for BASELINE.json's configs.

Generators emit Mesh-level arrays in the canonical form the reference's
MeshBuilder would hand to encode(): unique values in first-occurrence order,
point -> value maps only where duplicates exist, position attribute first,
normals / uvs Corner-domain children of the position attribute (OBJ entry:
io/obj/mod.rs:27-41; ids POSITION=0, NORMAL=1, TEXCOORD=2).
"""
import numpy as np

from .mesh import Attribute, AttributeDomain, AttributeType, Mesh


def _hash01(seed, idx):
    """Integer hash -> float32 in [0,1): murmur3 finaliser over (seed, i)."""
    h = (idx.astype(np.uint64) * np.uint64(0x9E3779B1) + np.uint64(seed) * np.uint64(0x85EBCA6B)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(16)
    h = (h * np.uint64(0x85EBCA6B)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(13)
    h = (h * np.uint64(0xC2B2AE35)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(16)
    return (h >> np.uint64(8)).astype(np.float32) / np.float32(1 << 24)


def _quads_to_faces(nx, ny, stride):
    """(nx x ny) quads over a point lattice with row stride `stride` -> CCW triangles."""
    i, j = np.meshgrid(np.arange(nx, dtype=np.uint32), np.arange(ny, dtype=np.uint32), indexing="ij")
    a = (i * stride + j).ravel()
    b = ((i + 1) * stride + j).ravel()
    c = ((i + 1) * stride + j + 1).ravel()
    d = (i * stride + j + 1).ravel()
    f = np.empty((a.size * 2, 3), dtype=np.uint32)
    f[0::2] = np.stack([a, b, c], axis=1)
    f[1::2] = np.stack([a, c, d], axis=1)
    return f


def _assemble(faces, pos, nrm, uv, with_normals=True, with_uvs=True):
    atts = [Attribute.from_points(pos, AttributeType.Position, AttributeDomain.Position, (), unique_id=0)]
    nid = 1
    if with_normals:
        atts.append(Attribute.from_points(nrm, AttributeType.Normal, AttributeDomain.Corner, (0,), unique_id=nid))
        nid += 1
    if with_uvs:
        atts.append(Attribute.from_points(uv, AttributeType.TextureCoordinate, AttributeDomain.Corner, (0,), unique_id=nid))
    return Mesh(faces, atts)


def grid_mesh(nx_pts, ny_pts, seed, with_normals=True, with_uvs=True):
    """Height-field grid: z = 0.15 sin(7x) cos(5y) + 0.05 hash(seed,i), x,y in [0,1];
    analytic unit normals of the smooth part; uv = (x, y). nx_pts*ny_pts vertices,
    2*(nx_pts-1)*(ny_pts-1) triangles. Open boundary, no seams."""
    xs = np.linspace(0.0, 1.0, nx_pts, dtype=np.float64)
    ys = np.linspace(0.0, 1.0, ny_pts, dtype=np.float64)
    X, Y = np.meshgrid(xs, ys, indexing="ij")
    idx = np.arange(nx_pts * ny_pts, dtype=np.uint64)
    Z = 0.15 * np.sin(7 * X) * np.cos(5 * Y)
    z = (Z.ravel() + 0.05 * _hash01(seed, idx).astype(np.float64))
    pos = np.stack([X.ravel(), Y.ravel(), z], axis=1).astype(np.float32)
    fx = 0.15 * 7 * np.cos(7 * X) * np.cos(5 * Y)
    fy = -0.15 * 5 * np.sin(7 * X) * np.sin(5 * Y)
    n = np.stack([-fx.ravel(), -fy.ravel(), np.ones(fx.size)], axis=1)
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    uv = np.stack([X.ravel(), Y.ravel()], axis=1).astype(np.float32)
    faces = _quads_to_faces(nx_pts - 1, ny_pts - 1, ny_pts)
    return _assemble(faces, pos, n.astype(np.float32), uv, with_normals, with_uvs)


def torus_mesh(nu, nv, seed, with_normals=True, with_uvs=True, R=1.0, r=0.35):
    """Closed torus, nu x nv quads. Points form a (nu+1) x (nv+1) lattice whose last
    row/column repeat the first (same position and normal, different uv), so the
    position attribute carries a point map and both parametric cuts are uv seams."""
    I, J = np.meshgrid(np.arange(nu + 1), np.arange(nv + 1), indexing="ij")
    Im, Jm = I % nu, J % nv
    th = 2 * np.pi * Im / nu
    ph = 2 * np.pi * Jm / nv
    node = (Im * nv + Jm).ravel().astype(np.uint64)
    rr = r * (1.0 + 0.04 * (_hash01(seed, node).astype(np.float64) - 0.5)).reshape(I.shape)
    cx, cy, cz = np.cos(th) * np.cos(ph), np.sin(th) * np.cos(ph), np.sin(ph)
    pos = np.stack([(R * np.cos(th) + rr * cx).ravel(), (R * np.sin(th) + rr * cy).ravel(), (rr * cz).ravel()], axis=1).astype(np.float32)
    nrm = np.stack([cx.ravel(), cy.ravel(), cz.ravel()], axis=1).astype(np.float32)
    uv = np.stack([(I / nu).ravel(), (J / nv).ravel()], axis=1).astype(np.float32)
    faces = _quads_to_faces(nu, nv, nv + 1)
    return _assemble(faces, pos, nrm, uv, with_normals, with_uvs)


def config1_mesh():
    """BASELINE config 1: 101 x 251 grid => 25 351 vertices, 50 000 triangles, seed 1."""
    return grid_mesh(101, 251, 1)


def config2_mesh(n=1000):
    """BASELINE config 2: 1000 x 1000 grid => 1 000 000 vertices, 1 996 002 triangles, seed 2."""
    return grid_mesh(n, n, 2)


def config3_mesh(nu=2000, nv=2500):
    """BASELINE config 3: 2000 x 2500-quad torus => 10 000 000 triangles, seed 3."""
    return torus_mesh(nu, nv, 3)


def batch_vertex_counts(n=4096, lo=1_000, hi=100_000, seed=0xD1AC0):
    """BASELINE config 4: vertex counts log-uniform in [lo, hi]."""
    rng = np.random.default_rng(seed)
    return np.exp(rng.uniform(np.log(lo), np.log(hi), size=n)).astype(np.int64)


def batch_mesh(k, target_vertices):
    """k-th primitive of config 4: alternating grid patch / torus with ~target vertices."""
    if k % 2 == 0:
        nx = max(2, int(round(np.sqrt(target_vertices))))
        ny = max(2, int(target_vertices) // nx)
        return grid_mesh(nx, ny, 1000 + k)
    nu = max(3, int(round(np.sqrt(target_vertices / 1.25))))
    nv = max(3, int(target_vertices) // nu)
    return torus_mesh(nu, nv, 1000 + k)


def _batch_mesh_job(kc):
    return batch_mesh(kc[0], int(kc[1]))


def batch_meshes(counts=None, procs=None, first=0):
    """Config 4's primitives (or the first len(counts) of them) generated by a pool of forked worker processes
    (numpy only; ~45 s of single-thread work for all 4096). `first` offsets the primitive index."""
    import multiprocessing as mp
    import os
    if counts is None:
        counts = batch_vertex_counts()
    jobs = [(first + k, int(c)) for k, c in enumerate(counts)]
    procs = procs or min(32, os.cpu_count() or 1)
    if procs <= 1 or len(jobs) < 8:
        return [_batch_mesh_job(j) for j in jobs]
    with mp.get_context("fork").Pool(procs) as pool:
        return pool.map(_batch_mesh_job, jobs, chunksize=4)
