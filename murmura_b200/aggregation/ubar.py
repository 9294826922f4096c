"""UBAR: two-stage (distance shortlist, then loss test) robust aggregation.

Parity: reference ``murmura/aggregation/ubar.py:15-271``.  Stage 1 keeps the
``max(min_neighbors, int(ρ·d))`` closest neighbours (L2 over all tensors).  Stage 2 draws one
batch from the train loader, evaluates ``CrossEntropyLoss`` of own and each shortlisted
state through the node's live model (left in ``eval()`` holding the last candidate's
weights, as in the reference) and keeps ``loss_j ≤ loss_own`` (best-loss fallback).
Output ``α·own + (1-α)·average_states(kept)``.
"""
from __future__ import annotations

import math
import time
from collections import defaultdict
from typing import Dict, List

import numpy as np
import torch
import torch.nn as nn

from murmura_b200.aggregation.base import (Aggregator, average_states, blend_states,
                                           squared_distance_all_keys)
from murmura_b200.core.types import ModelState


class UBARAggregator(Aggregator):
    kernel_family = "ubar"

    def __init__(self, rho: float = 0.4, alpha: float = 0.5, min_neighbors: int = 1, **kwargs):
        super().__init__(**kwargs)
        self.rho, self.alpha, self.min_neighbors = rho, alpha, min_neighbors
        self.stage1_acceptance_history: List[float] = []
        self.stage2_acceptance_history: List[float] = []
        self.neighbor_distances = defaultdict(list)
        self.neighbor_losses = defaultdict(list)
        self.distance_computation_time = 0.0
        self.loss_computation_time = 0.0
        self.aggregation_computation_time = 0.0
        self.criterion = nn.CrossEntropyLoss()

    def num_shortlisted(self, degree: int) -> int:
        return max(self.min_neighbors, int(self.rho * degree))

    def shortlist(self, distances: Dict[int, float]) -> List[int]:
        ranked = sorted(distances.items(), key=lambda kv: kv[1])
        chosen = [nid for nid, _ in ranked[: self.num_shortlisted(len(distances))]]
        self.stage1_acceptance_history.append(len(chosen) / max(1, len(distances)))
        return chosen

    def loss_filter(self, own_loss: float, losses: Dict[int, float]) -> List[int]:
        kept = [nid for nid, l in losses.items() if l <= own_loss]
        if not kept and losses:
            kept = [min(losses.items(), key=lambda kv: kv[1])[0]]
        self.stage2_acceptance_history.append(len(kept) / max(1, len(losses)))
        return kept

    def _batch_loss(self, model: nn.Module, batch, device) -> float:
        t0 = time.time()
        model.eval()
        xb, yb = batch
        with torch.no_grad():
            loss = float(self.criterion(model(xb.to(device)), yb.to(device)))
        self.loss_computation_time += time.time() - t0
        return loss

    def aggregate(self, node_id: int, own_state: ModelState, neighbor_states: Dict[int, ModelState],
                  round_num: int, train_loader=None, model_template: nn.Module = None,
                  device: torch.device = None, **kwargs) -> ModelState:
        if not neighbor_states:
            return own_state
        t0 = time.time()
        distances = {}
        for nid, st in neighbor_states.items():
            distances[nid] = math.sqrt(squared_distance_all_keys(own_state, st))
            self.neighbor_distances[nid].append(distances[nid])
        self.distance_computation_time += time.time() - t0
        kept = self.shortlist(distances)

        if kept and train_loader is not None and model_template is not None and device is not None:
            try:
                batch = next(iter(train_loader))
            except StopIteration:
                batch = None
            if batch is not None:
                model_template.load_state_dict(own_state, strict=False)
                own_loss = self._batch_loss(model_template, batch, device)
                losses = {}
                for nid in kept:
                    model_template.load_state_dict(neighbor_states[nid], strict=False)
                    losses[nid] = self._batch_loss(model_template, batch, device)
                    self.neighbor_losses[nid].append(losses[nid])
                kept = self.loss_filter(own_loss, losses)

        t1 = time.time()
        if not kept:
            self.aggregation_computation_time += time.time() - t1
            return own_state
        out = blend_states(own_state, average_states([neighbor_states[n] for n in kept]), self.alpha)
        self.aggregation_computation_time += time.time() - t1
        return out

    def get_statistics(self) -> Dict:
        s1 = float(np.mean(self.stage1_acceptance_history)) if self.stage1_acceptance_history else 0.0
        s2 = float(np.mean(self.stage2_acceptance_history)) if self.stage2_acceptance_history else 0.0
        return {
            "algorithm": "UBAR",
            "total_rounds_processed": len(self.stage1_acceptance_history),
            "stage1_mean_acceptance_rate": s1,
            "stage2_mean_acceptance_rate": s2,
            "overall_acceptance_rate": s1 * s2 if self.stage1_acceptance_history and self.stage2_acceptance_history else 0.0,
            "distance_computation_time": self.distance_computation_time,
            "loss_computation_time": self.loss_computation_time,
            "aggregation_computation_time": self.aggregation_computation_time,
        }
