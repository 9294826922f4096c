"""Topology generators: ring, fully-connected, Erdős–Rényi, k-regular ring lattice.

Behavioural parity with reference ``murmura/topology/generators.py:11-140`` (aliases,
RNG stream of the Erdős–Rényi scan, isolated-node repair, odd-``k`` bump, ``k>=n`` →
complete graph).  Implemented over a boolean adjacency matrix instead of list surgery.
"""
from __future__ import annotations

import random
from typing import List

import numpy as np

from murmura_b200.topology.base import Topology

_ALIASES = {
    "ring": "ring",
    "fully": "fully", "full": "fully",
    "erdos": "erdos", "er": "erdos", "erdos-renyi": "erdos",
    "k-regular": "k-regular", "kregular": "k-regular",
}


def _from_adjacency(adj: np.ndarray) -> Topology:
    n = adj.shape[0]
    neighbors: List[List[int]] = [np.flatnonzero(adj[i]).tolist() for i in range(n)]
    iu, ju = np.nonzero(np.triu(adj, k=1))
    edges = sorted(zip(iu.tolist(), ju.tolist()))
    return Topology(num_nodes=n, neighbors=neighbors, edges=edges)


def _link(adj: np.ndarray, i: int, j: int) -> None:
    if i != j:
        adj[i, j] = adj[j, i] = True


def _ring(n: int) -> Topology:
    adj = np.zeros((n, n), dtype=bool)
    for i in range(n):
        _link(adj, i, (i + 1) % n)
    return _from_adjacency(adj)


def _fully(n: int) -> Topology:
    adj = ~np.eye(n, dtype=bool)
    return _from_adjacency(adj)


def _erdos_renyi(n: int, p: float, seed: int) -> Topology:
    if not 0 <= p <= 1:
        raise ValueError(f"Edge probability p must be in [0, 1], got {p}")
    rng = random.Random(seed)
    adj = np.zeros((n, n), dtype=bool)
    # one uniform draw per unordered pair, scanned row-major (i < j): keeps the
    # reference's RNG stream so seeds reproduce the same graphs.
    for i in range(n):
        for j in range(i + 1, n):
            if rng.random() < p:
                _link(adj, i, j)
    for i in range(n):          # repair isolated nodes (sequential, sees earlier repairs)
        if not adj[i].any():
            _link(adj, i, (i + 1) % n)
    return _from_adjacency(adj)


def _k_regular(n: int, k: int) -> Topology:
    if k % 2:
        print(f"Warning: k={k} is odd, using k={k+1} for regular ring lattice")
        k += 1
    if k >= n:
        print(f"Warning: k={k} >= n={n}, creating fully connected graph")
        return _fully(n)
    adj = np.zeros((n, n), dtype=bool)
    for i in range(n):
        for off in range(1, k // 2 + 1):
            _link(adj, i, (i + off) % n)
    return _from_adjacency(adj)


def create_topology(topology_type: str, num_nodes: int, **kwargs) -> Topology:
    """Build a :class:`Topology`.

    kwargs: ``p`` (erdos, default 0.3), ``k`` (k-regular, default 4), ``seed`` (12345).
    Passing ``p=None``/``k=None`` explicitly is an error for the topologies that need
    them, exactly as in the reference.
    """
    kind = _ALIASES.get(topology_type.lower())
    if kind is None:
        raise ValueError(f"Unknown topology type: {topology_type.lower()}")
    if kind == "ring":
        return _ring(num_nodes)
    if kind == "fully":
        return _fully(num_nodes)
    if kind == "erdos":
        p = kwargs.get("p", 0.3)
        if p is None:
            raise TypeError("erdos topology requires an edge probability 'p'")
        return _erdos_renyi(num_nodes, p, kwargs.get("seed", 12345))
    k = kwargs.get("k", 4)
    if k is None:
        raise TypeError("k-regular topology requires a degree 'k'")
    return _k_regular(num_nodes, k)
