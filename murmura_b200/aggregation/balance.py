"""BALANCE: distance filter with an exponentially tightening threshold.

Parity: reference ``murmura/aggregation/balance.py:13-185``.  Accept neighbour ``j`` iff
``‖own-θ_j‖ ≤ γ·exp(-κ·t/T)·‖own‖`` (norms over *all* tensors), closest-neighbour fallback
when fewer than ``min_neighbors`` pass, output ``α·own + (1-α)·mean(accepted)`` on every key.
B200 path: ``ops.edge_distances`` (phase A) → ``ops.balance_filter`` → ``ops.weighted_gather``.
"""
from __future__ import annotations

import math
import time
from collections import defaultdict
from typing import Dict, List

import numpy as np
import torch

from murmura_b200.aggregation.base import (Aggregator, blend_states, squared_distance_all_keys,
                                           squared_norm_all_keys)
from murmura_b200.core.types import ModelState


def decayed_factor(gamma: float, kappa: float, round_num: int, total_rounds: int) -> float:
    return gamma * math.exp(-kappa * (round_num / max(1, total_rounds)))


class BALANCEAggregator(Aggregator):
    kernel_family = "balance"

    def __init__(self, gamma: float = 2.0, kappa: float = 1.0, alpha: float = 0.5,
                 min_neighbors: int = 1, total_rounds: int = 20, **kwargs):
        super().__init__(**kwargs)
        self.gamma, self.kappa, self.alpha = gamma, kappa, alpha
        self.min_neighbors, self.total_rounds = min_neighbors, total_rounds
        self.acceptance_history: List[float] = []
        self.threshold_history: List[float] = []
        self.neighbor_distances = defaultdict(list)
        self.filtering_computation_time = 0.0
        self.aggregation_computation_time = 0.0

    # -- filter -------------------------------------------------------------------
    def threshold(self, own_norm: float, round_num: int) -> float:
        thr = decayed_factor(self.gamma, self.kappa, round_num, self.total_rounds) * own_norm
        self.threshold_history.append(thr)
        return thr

    def select(self, distances: Dict[int, float], thr: float) -> List[int]:
        """Accepted neighbour ids given their distances (records statistics)."""
        accepted = [nid for nid, d in distances.items() if d <= thr]
        self.acceptance_history.append(len(accepted) / max(1, len(distances)))
        if len(accepted) < self.min_neighbors and distances:
            closest = min(distances.items(), key=lambda kv: kv[1])[0]
            if closest not in accepted:
                accepted.append(closest)
        return accepted

    def aggregate(self, node_id: int, own_state: ModelState, neighbor_states: Dict[int, ModelState],
                  round_num: int, **kwargs) -> ModelState:
        t0 = time.time()
        thr = self.threshold(math.sqrt(squared_norm_all_keys(own_state)), round_num)
        distances = {}
        for nid, st in neighbor_states.items():
            distances[nid] = math.sqrt(squared_distance_all_keys(own_state, st))
            self.neighbor_distances[nid].append(distances[nid])
        accepted = self.select(distances, thr)
        self.filtering_computation_time += time.time() - t0

        t1 = time.time()
        if not accepted:
            self.aggregation_computation_time += time.time() - t1
            return own_state
        mean: ModelState = {}
        for key, own_t in own_state.items():
            acc = torch.zeros_like(own_t)
            for nid in accepted:
                acc += neighbor_states[nid][key].to(acc.dtype)
            mean[key] = acc / len(accepted)
        out = blend_states(own_state, mean, self.alpha)
        self.aggregation_computation_time += time.time() - t1
        return out

    def get_statistics(self) -> Dict:
        return {
            "mean_acceptance_rate": float(np.mean(self.acceptance_history)) if self.acceptance_history else 0.0,
            "current_threshold": self.threshold_history[-1] if self.threshold_history else 0.0,
            "total_rounds_processed": len(self.acceptance_history),
            "filtering_computation_time": self.filtering_computation_time,
            "aggregation_computation_time": self.aggregation_computation_time,
        }
