"""Topology-liar attack for DMTT: Byzantine nodes falsify their TOPO_CLAIM.

Parity: reference ``murmura/attacks/topology_liar.py:21-102``: private RNG selection with
at least one compromised node, optional wrapped model attack, and
``get_false_claims = sorted(true ∪ other Byzantine ids)``.  ``claim_bitmask`` is the
device-side form used by ``ops.dmtt``.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Set

import numpy as np

from murmura_b200.attacks.base import select_compromised
from murmura_b200.core.types import ModelState


class TopologyLiarAttack:
    def __init__(self, num_nodes: int, attack_percentage: float, seed: int = 42, model_attack=None):
        self.num_nodes = num_nodes
        self.seed = seed
        self._model_attack = model_attack
        self._compromised: Set[int] = select_compromised(
            num_nodes, attack_percentage, seed, at_least_one=True, reseed_global=False)

    def is_compromised(self, node_id: int) -> bool:
        return node_id in self._compromised

    def get_compromised_nodes(self) -> Set[int]:
        return set(self._compromised)

    def apply_attack(self, node_id: int, model_state: ModelState, round_num: int, **kwargs) -> ModelState:
        if self._model_attack is None:
            return model_state
        return self._model_attack.apply_attack(node_id=node_id, model_state=model_state,
                                               round_num=round_num, **kwargs)

    def get_false_claims(self, node_id: int, true_neighbors: List[int], round_num: int) -> List[int]:
        return sorted(set(true_neighbors) | (self._compromised - {node_id}))

    # ---- B200 engine helpers ----------------------------------------------------
    def claim_bitmask(self, adjacency: np.ndarray) -> np.ndarray:
        """``[N, N]`` bool claim matrix: honest rows = truth, liar rows = truth ∪ liars."""
        claims = adjacency.copy()
        liars = sorted(self._compromised)
        for i in liars:
            claims[i, liars] = True
            claims[i, i] = False
        return claims

    def device_spec(self) -> Optional[Dict[str, float]]:
        inner = getattr(self._model_attack, "device_spec", None)
        return inner() if inner else None
