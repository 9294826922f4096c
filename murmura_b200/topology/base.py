"""Static topology container (parity: reference ``murmura/topology/base.py:7-60``)."""
from __future__ import annotations

from collections import deque
from dataclasses import dataclass
from typing import List, Tuple

import numpy as np


@dataclass
class Topology:
    """Undirected communication graph.

    ``neighbors[i]`` is the sorted adjacency list of node ``i``; ``edges`` holds each
    undirected edge once as ``(lo, hi)``.
    """

    num_nodes: int
    neighbors: List[List[int]]
    edges: List[Tuple[int, int]]

    def __post_init__(self) -> None:
        assert len(self.neighbors) == self.num_nodes, (
            f"neighbors list length ({len(self.neighbors)}) != num_nodes ({self.num_nodes})"
        )

    def degree(self, node_id: int) -> int:
        return len(self.neighbors[node_id])

    def avg_degree(self) -> float:
        return sum(map(len, self.neighbors)) / self.num_nodes

    def is_connected(self) -> bool:
        if self.num_nodes == 0:
            return True
        seen = {0}
        frontier = deque([0])
        while frontier:
            for nb in self.neighbors[frontier.popleft()]:
                if nb not in seen:
                    seen.add(nb)
                    frontier.append(nb)
        return len(seen) == self.num_nodes

    # ---- B200 engine helpers (not in the reference) -------------------------------
    def to_csr(self, include_self: bool = True) -> Tuple[np.ndarray, np.ndarray]:
        """CSR edge list consumed by the fused exchange+aggregate kernels.

        Row ``i`` lists the sources node ``i`` reads; with ``include_self`` the node's own
        id is always the first entry of its row (the kernels rely on that).
        """
        row_ptr = np.zeros(self.num_nodes + 1, dtype=np.int32)
        cols: List[int] = []
        for i, nbrs in enumerate(self.neighbors):
            row = ([i] if include_self else []) + list(nbrs)
            cols.extend(row)
            row_ptr[i + 1] = len(cols)
        return row_ptr, np.asarray(cols, dtype=np.int32)

    def adjacency(self) -> np.ndarray:
        adj = np.zeros((self.num_nodes, self.num_nodes), dtype=bool)
        for a, b in self.edges:
            adj[a, b] = adj[b, a] = True
        return adj
