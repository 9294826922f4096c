"""Per-node DMTT trust state (link reliability, Beta trust, collaboration score, Top-B).

Parity: reference ``murmura/dmtt/state.py:22-159``.  Stored as dense length-N vectors rather than
dicts so the same layout is what the device kernel (``ops.dmtt_update``) works on; the
scalar accessors keep the reference's method names.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import numpy as np

from murmura_b200.config.schema import DMTTConfig


class DMTTNodeState:
    def __init__(self, node_id: int, cfg: DMTTConfig, num_nodes: Optional[int] = None):
        self.node_id, self.cfg = node_id, cfg
        n = num_nodes or 0
        self._c_hat = np.full(n, 0.5)
        self._alpha = np.ones(n)
        self._beta = np.ones(n)
        self._seen = np.zeros(n, dtype=bool)

    def _init(self, j: int) -> None:
        if j >= len(self._c_hat):
            grow = j + 1 - len(self._c_hat)
            self._c_hat = np.concatenate([self._c_hat, np.full(grow, 0.5)])
            self._alpha = np.concatenate([self._alpha, np.ones(grow)])
            self._beta = np.concatenate([self._beta, np.ones(grow)])
            self._seen = np.concatenate([self._seen, np.zeros(grow, dtype=bool)])
        self._seen[j] = True

    # Algorithm 1 — link reliability EMA
    def update_link_reliability(self, j: int, received: bool) -> None:
        self._init(j)
        rho = self.cfg.rho
        self._c_hat[j] = (1.0 - rho) * self._c_hat[j] + rho * (1.0 if received else 0.0)

    # Algorithm 4 — Beta evidence with forgetting
    def update_trust(self, j: int, d: float, x: float, c: float = 0.0) -> None:
        self._init(j)
        cfg = self.cfg
        self._alpha[j] = max(0.01, cfg.lambda_forget * self._alpha[j] + cfg.w_d * d + cfg.w_c * c)
        self._beta[j] = max(0.01, cfg.lambda_forget * self._beta[j] + cfg.w_x * x)

    def topo_trust(self, j: int) -> float:
        self._init(j)
        a, b = float(self._alpha[j]), float(self._beta[j])
        s = a + b
        mean = a / s
        spread = math.sqrt(max(0.0, a * b / (s * s * (s + 1.0))))
        return mean * math.exp(-self.cfg.eta * max(0.0, spread - self.cfg.tau_U))

    def link_reliability(self, j: int) -> float:
        self._init(j)
        return float(self._c_hat[j])

    def model_score(self, a_ij: float, u_bar_ij: float = 0.0) -> float:
        cfg = self.cfg
        s = (1.0 - u_bar_ij) * (cfg.w_a * a_ij + (1.0 - cfg.w_a))
        if u_bar_ij > cfg.tau_u:
            s *= math.exp(-(u_bar_ij - cfg.tau_u))
        return max(0.0, s)

    def collab_score(self, j: int, s_model: float, c_comm: float = 0.0) -> float:
        cfg = self.cfg
        return (cfg.lambda1 * s_model + cfg.lambda2 * self.topo_trust(j)
                + cfg.lambda3 * self.link_reliability(j) - cfg.lambda4 * c_comm)

    def top_b(self, candidates: List[int], model_scores: Dict[int, float], B: int) -> List[int]:
        if not candidates:
            return []
        ranked = sorted(((j, self.collab_score(j, model_scores.get(j, 0.5))) for j in candidates),
                        key=lambda kv: kv[1], reverse=True)     # stable: ties keep candidate order
        return [j for j, _ in ranked[:B]]

    def state_summary(self, peers: Optional[List[int]] = None) -> Dict[int, dict]:
        targets = peers if peers is not None else np.flatnonzero(self._seen).tolist()
        return {j: {"c_hat": self.link_reliability(j), "T_topo": self.topo_trust(j),
                    "alpha": float(self._alpha[j]) if j < len(self._alpha) else 1.0,
                    "beta": float(self._beta[j]) if j < len(self._beta) else 1.0} for j in targets}
