"""Gaussian model-poisoning attack: ``theta + sigma * N(0, 1)`` on float tensors.

Parity: reference ``murmura/attacks/gaussian.py:10-90`` (selection, banner, float-only
noise on the *state*, ints cloned).  ``device_spec()`` tells the B200 engine how to fuse
the attack into the publish kernel (``ops.publish`` with Philox noise).
"""
from __future__ import annotations

from typing import Dict, Set

import torch

from murmura_b200.attacks.base import select_compromised
from murmura_b200.core.types import ModelState

_NOISY_DTYPES = (torch.float32, torch.float64, torch.float16)


class GaussianAttack:
    def __init__(self, num_nodes: int, attack_percentage: float, noise_std: float = 10.0, seed: int = 42):
        self.num_nodes = num_nodes
        self.attack_percentage = attack_percentage
        self.noise_std = noise_std
        self.seed = seed
        self.compromised_nodes: Set[int] = select_compromised(
            num_nodes, attack_percentage, seed, at_least_one=False, reseed_global=True)
        print(f"Gaussian Attack: Compromised {len(self.compromised_nodes)}/{num_nodes} nodes")
        print(f"  Compromised nodes: {sorted(self.compromised_nodes)}")
        print(f"  Noise std: {noise_std}")

    def is_compromised(self, node_id: int) -> bool:
        return node_id in self.compromised_nodes

    def get_compromised_nodes(self) -> Set[int]:
        return self.compromised_nodes

    def apply_attack(self, node_id: int, model_state: ModelState, round_num: int, **kwargs) -> ModelState:
        if node_id not in self.compromised_nodes:
            return model_state
        out: ModelState = {}
        for name, t in model_state.items():
            if t.dtype in _NOISY_DTYPES:
                out[name] = t + torch.randn_like(t) * self.noise_std
            else:
                out[name] = t.clone()
        return out

    def device_spec(self) -> Dict[str, float]:
        """Fused-publish description: ``pub = scale * theta + noise_std * philox_normal``."""
        return {"kind": "gaussian", "scale": 1.0, "noise_std": float(self.noise_std), "seed": int(self.seed)}
