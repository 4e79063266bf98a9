"""Partitioning of independent meshes / glTF primitives over ranks (one process per GPU).

encode() is a pure function of one mesh (no shared tables, no cross-mesh statistics:
SURVEY.md §8e), so ranks exchange nothing on the data path. Work is balanced by a
deterministic longest-processing-time-first assignment on a per-mesh cost estimate
(corners + points, or the vertex count when the meshes do not exist yet); every rank computes
the same assignment from the same costs. Used by bench.py to deal config 4's primitives to the
ranks of a multi-GPU job (one process per GPU, each calling dxo_encode_batch on its shard); inside
a process the batch entry orders its groups the same way (batch.cpp: longest first, shared queue)."""
import heapq

import numpy as np


def mesh_cost(mesh):
    return int(mesh.faces.shape[0]) * 3 + int(mesh.num_points())


def shard_by_cost(costs, world_size):
    """Returns a list of index lists, one per rank; deterministic for equal inputs."""
    costs = np.asarray(costs, dtype=np.int64)
    order = np.lexsort((np.arange(costs.size), -costs))  # cost descending, index ascending
    heap = [(0, r) for r in range(world_size)]
    heapq.heapify(heap)
    shards = [[] for _ in range(world_size)]
    for i in order:
        load, r = heapq.heappop(heap)
        shards[r].append(int(i))
        heapq.heappush(heap, (load + int(costs[i]), r))
    for s in shards:
        s.sort()
    return shards


def my_shard(costs, rank, world_size):
    return shard_by_cost(costs, world_size)[rank]
