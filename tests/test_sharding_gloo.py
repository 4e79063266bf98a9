"""CPU, world_size 2 over gloo: the N>1 host logic — deterministic cost-balanced sharding of
independent meshes, per-rank encoding, and reassembly in input order. The per-rank encoder
here is the CPU oracle (the device path needs a GPU); the partitioning is draco_oxide_b200.sharding,
the function bench.py deals config 4's primitives to its ranks with (bench.config4_shard)."""
import hashlib
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_meshes, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import orc
    from draco_oxide_b200 import sharding, synth
    counts = synth.batch_vertex_counts(n_meshes, 100, 3000, seed=7)
    meshes = [synth.batch_mesh(k, int(c)) for k, c in enumerate(counts)]
    costs = [sharding.mesh_cost(m) for m in meshes]
    mine = sharding.my_shard(costs, rank, world)
    local = {i: hashlib.sha256(orc.encode(meshes[i])).hexdigest() for i in mine}
    gathered = [None] * world
    dist.all_gather_object(gathered, local)
    dist.barrier()
    if rank == 0:
        merged = {}
        for g in gathered:
            assert not (set(g) & set(merged)), "a mesh was encoded by two ranks"
            merged.update(g)
        with open(os.path.join(out_dir, "result.txt"), "w") as f:
            for i in range(n_meshes):
                f.write(merged[i] + "\n")
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process(orc, tmp_path):
    from draco_oxide_b200 import sharding, synth
    n = 12
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n, str(tmp_path)), nprocs=2, join=True)
    got = open(tmp_path / "result.txt").read().split()
    counts = synth.batch_vertex_counts(n, 100, 3000, seed=7)
    want = [hashlib.sha256(orc.encode(synth.batch_mesh(k, int(c)))).hexdigest() for k, c in enumerate(counts)]
    assert got == want


def test_shard_by_cost_properties():
    from draco_oxide_b200 import sharding
    rng = np.random.default_rng(3)
    costs = rng.integers(1_000, 100_000, 4096)
    for world in (1, 2, 4, 8):
        shards = sharding.shard_by_cost(costs, world)
        flat = sorted(i for s in shards for i in s)
        assert flat == list(range(costs.size))                     # a partition: nothing lost, nothing twice
        loads = [int(costs[s].sum()) for s in shards]
        assert max(loads) - min(loads) <= int(costs.max())          # LPT bound
        assert shards == sharding.shard_by_cost(costs, world)       # deterministic


def test_bench_deals_config4_with_the_sharding_module():
    """bench.py's per-rank shard of config 4 is sharding.my_shard: the shards of all ranks partition the primitives."""
    import importlib.util
    from draco_oxide_b200 import sharding, synth
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    n = 64
    counts = synth.batch_vertex_counts()[:n]
    seen = []
    for rank in range(4):
        meshes, total_vertices, total = bench.config4_shard(n, rank, 4)
        assert total == n and total_vertices == int(counts.sum())
        mine = sharding.my_shard(counts, rank, 4)
        assert sorted(m.num_points() for m in meshes) == sorted(synth.batch_mesh(k, int(counts[k])).num_points() for k in mine)
        seen += mine
    assert sorted(seen) == list(range(n))
