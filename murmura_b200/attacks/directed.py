"""Directed-deviation attack: broadcast ``lambda * theta`` (default ``lambda=-5``).

Parity: reference ``murmura/attacks/directed.py:10-89``.  On the B200 engine the scale is
folded into the publish kernel so it costs no extra pass.
"""
from __future__ import annotations

from typing import Dict, Set

import torch

from murmura_b200.attacks.base import select_compromised
from murmura_b200.core.types import ModelState

_SCALED_DTYPES = (torch.float32, torch.float64, torch.float16)


class DirectedDeviationAttack:
    def __init__(self, num_nodes: int, attack_percentage: float, lambda_param: float = -5.0, seed: int = 42):
        self.num_nodes = num_nodes
        self.attack_percentage = attack_percentage
        self.lambda_param = lambda_param
        self.seed = seed
        self.compromised_nodes: Set[int] = select_compromised(
            num_nodes, attack_percentage, seed, at_least_one=False, reseed_global=True)
        print(f"Directed Deviation Attack: Compromised {len(self.compromised_nodes)}/{num_nodes} nodes")
        print(f"  Compromised nodes: {sorted(self.compromised_nodes)}")
        print(f"  Lambda (scaling factor): {lambda_param}")

    def is_compromised(self, node_id: int) -> bool:
        return node_id in self.compromised_nodes

    def get_compromised_nodes(self) -> Set[int]:
        return self.compromised_nodes

    def apply_attack(self, node_id: int, model_state: ModelState, round_num: int, **kwargs) -> ModelState:
        if node_id not in self.compromised_nodes:
            return model_state
        return {name: (t * self.lambda_param if t.dtype in _SCALED_DTYPES else t.clone())
                for name, t in model_state.items()}

    def device_spec(self) -> Dict[str, float]:
        return {"kind": "directed_deviation", "scale": float(self.lambda_param), "noise_std": 0.0,
                "seed": int(self.seed)}
