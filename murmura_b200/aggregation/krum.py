"""Krum selection (single winner, un-squared distance sums).

Parity: reference ``murmura/aggregation/krum.py:8-75``: ``m = d+1`` candidates, fallback to
own state when ``c >= (m-2)/2``, score = sum of the ``max(1, m-c-2)`` smallest L2
distances (float tensors only), first arg-min wins and its *whole* state is returned.
B200 path: Gram matrix on tcgen05 (``ops.gram``) → ``ops.krum_select`` → winner copy.
"""
from __future__ import annotations

from typing import Dict, List

from murmura_b200.aggregation.base import Aggregator, compute_model_distance
from murmura_b200.core.types import ModelState


def krum_scores(dist: List[List[float]], num_compromised: int) -> List[float]:
    """Scores from a full ``m×m`` distance table (diagonal ignored)."""
    m = len(dist)
    keep = max(1, m - num_compromised - 2)
    scores = []
    for i in range(m):
        others = sorted(dist[i][j] for j in range(m) if j != i)
        scores.append(sum(others[:keep]))
    return scores


class KrumAggregator(Aggregator):
    kernel_family = "krum"

    def __init__(self, num_compromised: int = 0, **kwargs):
        super().__init__(**kwargs)
        self.num_compromised = num_compromised

    def aggregate(self, node_id: int, own_state: ModelState, neighbor_states: Dict[int, ModelState],
                  round_num: int, **kwargs) -> ModelState:
        candidates = [own_state, *neighbor_states.values()]
        m = len(candidates)
        if self.num_compromised >= (m - 2) / 2:
            return own_state
        dist = [[0.0] * m for _ in range(m)]
        for i in range(m):
            for j in range(i + 1, m):
                dist[i][j] = dist[j][i] = compute_model_distance(candidates[i], candidates[j])
        scores = krum_scores(dist, self.num_compromised)
        return candidates[scores.index(min(scores))]
