"""Sketchguard: Count-Sketch filtering, full-state aggregation.

Parity: reference ``murmura/aggregation/sketchguard.py:13-274``.  Bucket/sign tables are
the ``RandomState(network_seed)`` draws (``randint`` then ``choice([-1,1])``) so every node
— and the device kernel ``ops.count_sketch`` which receives them packed as
``uint16 = bucket | sign<<15`` — agrees with the reference bit-for-bit on table contents.
Filter: ``‖s_own-s_j‖ ≤ γ·exp(-κt/T)·attack_factor·‖s_own‖`` with ``attack_factor=1.5`` when
the mean of the last three acceptance rates is below 0.3; closest fallback; then
``α·own + (1-α)·average_states(accepted)``.
"""
from __future__ import annotations

import time
from collections import defaultdict, deque
from typing import Dict, List, Optional

import numpy as np

from murmura_b200.aggregation.balance import decayed_factor
from murmura_b200.aggregation.base import Aggregator, average_states, blend_states
from murmura_b200.core.types import ModelState


def count_sketch_tables(model_dim: int, sketch_size: int, network_seed: int):
    rng = np.random.RandomState(network_seed)
    buckets = rng.randint(0, sketch_size, size=model_dim)
    signs = rng.choice([-1, 1], size=model_dim)
    return buckets, signs


def pack_sketch_tables(buckets: np.ndarray, signs: np.ndarray) -> np.ndarray:
    """``uint16`` per element: low 15 bits bucket, top bit set ⇔ sign == -1."""
    if buckets.max(initial=0) >= (1 << 15):
        raise ValueError("packed sketch tables support sketch_size <= 32768")
    return (buckets.astype(np.uint16) | ((signs < 0).astype(np.uint16) << 15)).astype(np.uint16)


class SketchguardAggregator(Aggregator):
    kernel_family = "sketchguard"

    def __init__(self, model_dim: int, sketch_size: int = 1000, gamma: float = 2.0, kappa: float = 1.0,
                 alpha: float = 0.5, min_neighbors: int = 1, network_seed: int = 42,
                 attack_detection_window: int = 5, total_rounds: int = 20, **kwargs):
        super().__init__(**kwargs)
        self.model_dim, self.sketch_size = model_dim, sketch_size
        self.gamma, self.kappa, self.alpha = gamma, kappa, alpha
        self.min_neighbors, self.network_seed, self.total_rounds = min_neighbors, network_seed, total_rounds
        self.hash_table, self.sign_table = count_sketch_tables(model_dim, sketch_size, network_seed)
        self.acceptance_history: List[float] = []
        self.threshold_history: List[float] = []
        self.neighbor_scores = defaultdict(list)
        self.attack_history: deque = deque(maxlen=attack_detection_window)
        self.sketch_computation_time = 0.0
        self.filtering_computation_time = 0.0
        self.aggregation_computation_time = 0.0

    # -- sketching ----------------------------------------------------------------
    def _flatten(self, state: ModelState) -> np.ndarray:
        return np.concatenate([t.detach().cpu().numpy().ravel() for t in state.values()
                               if t.is_floating_point()])

    def get_sketch(self, model_state: ModelState) -> np.ndarray:
        t0 = time.time()
        vec = self._flatten(model_state)
        n = len(vec)
        sketch = np.bincount(self.hash_table[:n], weights=self.sign_table[:n] * vec,
                             minlength=self.sketch_size)
        self.sketch_computation_time += time.time() - t0
        return sketch

    # -- filter -------------------------------------------------------------------
    def attack_factor(self) -> float:
        if len(self.attack_history) >= 3 and np.mean(list(self.attack_history)[-3:]) < 0.3:
            return 1.5
        return 1.0

    def threshold(self, own_sketch_norm: float, round_num: int) -> float:
        thr = (decayed_factor(self.gamma, self.kappa, round_num, self.total_rounds)
               * self.attack_factor() * own_sketch_norm)
        self.threshold_history.append(thr)
        return thr

    def select(self, distances: Dict[int, float], thr: float) -> List[int]:
        accepted = [nid for nid, d in distances.items() if d <= thr]
        rate = len(accepted) / max(1, len(distances))
        self.acceptance_history.append(rate)
        self.attack_history.append(rate)
        if len(accepted) < self.min_neighbors and distances:
            closest = min(distances.items(), key=lambda kv: kv[1])[0]
            if closest not in accepted:
                accepted.append(closest)
        return accepted

    def aggregate(self, node_id: int, own_state: ModelState, neighbor_states: Dict[int, ModelState],
                  round_num: int, neighbor_sketches: Optional[Dict[int, np.ndarray]] = None,
                  **kwargs) -> ModelState:
        own_sketch = self.get_sketch(own_state)
        if neighbor_sketches is None:
            neighbor_sketches = {nid: self.get_sketch(st) for nid, st in neighbor_states.items()}
        t0 = time.time()
        distances = {}
        for nid, sk in neighbor_sketches.items():
            distances[nid] = float(np.linalg.norm(own_sketch - sk))
            self.neighbor_scores[nid].append(distances[nid])
        self.filtering_computation_time += time.time() - t0
        accepted = self.select(distances, self.threshold(float(np.linalg.norm(own_sketch)), round_num))

        t1 = time.time()
        chosen = [neighbor_states[nid] for nid in accepted if nid in neighbor_states]
        if not chosen:
            self.aggregation_computation_time += time.time() - t1
            return own_state
        out = blend_states(own_state, average_states(chosen), self.alpha)
        self.aggregation_computation_time += time.time() - t1
        return out

    def get_statistics(self) -> Dict:
        return {
            "algorithm": "Sketchguard",
            "mean_acceptance_rate": float(np.mean(self.acceptance_history)) if self.acceptance_history else 0.0,
            "current_threshold": self.threshold_history[-1] if self.threshold_history else 0.0,
            "total_rounds_processed": len(self.acceptance_history),
            "sketch_computation_time": self.sketch_computation_time,
            "filtering_computation_time": self.filtering_computation_time,
            "aggregation_computation_time": self.aggregation_computation_time,
            "compression_ratio": self.model_dim / self.sketch_size,
        }
