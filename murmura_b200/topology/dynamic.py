"""Random-walk mobility on a 2-D torus → time-varying graph G^t.

Parity: reference ``murmura/topology/dynamic.py:16-105``.  Positions come from
``np.random.default_rng(seed)`` (PCG64) and are generated strictly sequentially, so a
given seed reproduces the reference's trajectories bit-for-bit (golden test in
``tests/test_topology.py``).  Adjacency is vectorised here; the B200 engine uploads all
``R`` rounds of positions once (``positions_tensor``) and builds the per-round CSR on the
device (``ops.mobility_adjacency``).
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np


class MobilityModel:
    def __init__(self, num_nodes: int, area_size: float = 100.0, comm_range: float = 30.0,
                 max_speed: float = 5.0, seed: int = 42, ensure_connected: bool = True):
        self.num_nodes = num_nodes
        self.area_size = area_size
        self.comm_range = comm_range
        self.max_speed = max_speed
        self.ensure_connected = ensure_connected
        self._rng = np.random.default_rng(seed)
        self._trace: List[np.ndarray] = [self._rng.uniform(0.0, area_size, size=(num_nodes, 2))]

    # ---- positions ------------------------------------------------------------------
    def positions_at(self, round_idx: int) -> np.ndarray:
        while len(self._trace) <= round_idx:
            step = self._rng.uniform(-self.max_speed, self.max_speed, size=(self.num_nodes, 2))
            self._trace.append((self._trace[-1] + step) % self.area_size)
        return self._trace[round_idx]

    def positions_tensor(self, rounds: int) -> np.ndarray:
        """``[rounds, N, 2]`` float64 — what the device adjacency kernel consumes."""
        self.positions_at(max(rounds - 1, 0))
        return np.stack(self._trace[:rounds], axis=0)

    # ---- distances / adjacency ------------------------------------------------------
    def _dist_matrix(self, pos: np.ndarray) -> np.ndarray:
        d = np.abs(pos[:, None, :] - pos[None, :, :])
        d = np.minimum(d, self.area_size - d)
        return np.sqrt(d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1])

    def torus_dist(self, i: int, j: int, round_idx: int) -> float:
        pos = self.positions_at(round_idx)
        return float(self._dist_matrix(pos[[i, j]])[0, 1])

    def adjacency_at(self, round_idx: int) -> np.ndarray:
        pos = self.positions_at(round_idx)
        dist = self._dist_matrix(pos)
        adj = dist < self.comm_range
        np.fill_diagonal(adj, False)
        if self.ensure_connected:
            masked = dist + np.where(np.eye(self.num_nodes, dtype=bool), np.inf, 0.0)
            for i in range(self.num_nodes):      # sequential: later nodes see earlier repairs
                if not adj[i].any():
                    j = int(np.argmin(masked[i]))
                    adj[i, j] = adj[j, i] = True
        return adj

    def neighbors_at(self, round_idx: int) -> Dict[int, List[int]]:
        """Adjacency list at ``round_idx``.

        For non-isolated nodes lists are ascending.  A repaired (isolated) node is
        appended to its nearest peer's list *after* that peer's in-range neighbours, as in
        the reference, so list order (not just membership) matches.
        """
        pos = self.positions_at(round_idx)
        dist = self._dist_matrix(pos)
        base = dist < self.comm_range
        np.fill_diagonal(base, False)
        adj: Dict[int, List[int]] = {i: np.flatnonzero(base[i]).tolist() for i in range(self.num_nodes)}
        if self.ensure_connected:
            masked = dist + np.where(np.eye(self.num_nodes, dtype=bool), np.inf, 0.0)
            for i in range(self.num_nodes):
                if not adj[i]:
                    j = int(np.argmin(masked[i]))
                    adj[i].append(j)
                    adj[j].append(i)
        return adj
