"""Evidential trust-aware aggregation (cross-evaluating neighbours' Dirichlet outputs).

Parity: reference ``murmura/aggregation/evidential_trust.py:25-469``.  For every neighbour:
run its weights on ≥``max_eval_samples`` local samples, measure vacuity ``K/S``, entropy and
accuracy → ``trust=(1-vac)(w_a·acc+1-w_a)``, exponentially penalised when ``vac>τ_u``, clamped
to [0,1], EMA-smoothed per neighbour, compared with the tightening threshold
``τ(t)=clamp(τ_base(1-γe^{-κt/T}), 0.05, τ_base)``.  Output
``self_weight·own + (1-self_weight)·Σ_j (trust_j/Σtrust)·θ_j``; own state when nobody passes;
plain mean when no evaluation context is supplied.  ``min_neighbors`` is accepted but unused,
as in the reference.  B200 path: ``ops.dirichlet_stats`` fused epilogue + ``ops.trust_filter``.
"""
from __future__ import annotations

import copy
import math
from collections import defaultdict
from typing import Any, Dict, List, Tuple

import torch
import torch.nn as nn

from murmura_b200.aggregation.base import Aggregator, average_states, set_model_state
from murmura_b200.core.types import ModelState


def trust_from_metrics(vacuity: float, accuracy: float, accuracy_weight: float,
                       vacuity_threshold: float) -> Tuple[float, float]:
    """(base_trust, final_trust) — shared by the CPU path and the device filter test."""
    base = (1.0 - vacuity) * (accuracy_weight * accuracy + (1.0 - accuracy_weight))
    trust = base * math.exp(-(vacuity - vacuity_threshold)) if vacuity > vacuity_threshold else base
    return base, max(0.0, min(1.0, trust))


class EvidentialTrustAggregator(Aggregator):
    kernel_family = "evidential_trust"

    def __init__(self, vacuity_threshold: float = 0.5, accuracy_weight: float = 0.5,
                 trust_threshold: float = 0.3, self_weight: float = 0.5,
                 use_adaptive_trust: bool = True, trust_momentum: float = 0.7,
                 use_tightening_threshold: bool = True, gamma: float = 0.5, kappa: float = 1.0,
                 total_rounds: int = 50, min_neighbors: int = 1, max_eval_samples: int = 100,
                 track_statistics: bool = True, **kwargs):
        super().__init__(**kwargs)
        self.vacuity_threshold = vacuity_threshold
        self.accuracy_weight = accuracy_weight
        self.trust_threshold = trust_threshold
        self.base_trust_threshold = trust_threshold
        self.self_weight = self_weight
        self.use_adaptive_trust = use_adaptive_trust
        self.trust_momentum = trust_momentum
        self.use_tightening_threshold = use_tightening_threshold
        self.gamma, self.kappa, self.total_rounds = gamma, kappa, total_rounds
        self.min_neighbors = min_neighbors
        self.max_eval_samples = max_eval_samples
        self.track_statistics = track_statistics
        self._trust_history: Dict[int, List[float]] = defaultdict(list)
        self._smoothed_trust: Dict[int, float] = {}
        self._statistics = self._fresh_stats()

    @staticmethod
    def _fresh_stats() -> Dict[str, Any]:
        return {"rounds_processed": 0, "neighbors_evaluated": 0, "neighbors_accepted": 0,
                "neighbors_rejected": 0, "avg_trust_scores": [], "avg_vacuity": [], "avg_entropy": [],
                "threshold_history": []}

    # -- pieces reused by the device engine -----------------------------------------
    def current_threshold(self, round_num: int) -> float:
        if not self.use_tightening_threshold:
            return self.trust_threshold
        decay = math.exp(-self.kappa * (round_num / max(1, self.total_rounds)))
        thr = self.base_trust_threshold * (1.0 - self.gamma * decay)
        return max(0.05, min(thr, self.base_trust_threshold))

    def smooth(self, neighbor_id: int, new_trust: float) -> float:
        prev = self._smoothed_trust.get(neighbor_id)
        value = new_trust if prev is None else self.trust_momentum * new_trust + (1 - self.trust_momentum) * prev
        self._smoothed_trust[neighbor_id] = value
        self._trust_history[neighbor_id].append(value)
        return value

    def score_from_metrics(self, metrics: Dict[str, float]) -> float:
        base, trust = trust_from_metrics(metrics["vacuity"], metrics["accuracy"], self.accuracy_weight,
                                         self.vacuity_threshold)
        metrics["base_trust"], metrics["final_trust"] = base, trust
        return trust

    def _cross_evaluate(self, state: ModelState, template: nn.Module, loader, device) -> Dict[str, float]:
        probe = copy.deepcopy(template)
        set_model_state(probe, state)
        probe.to(device).eval()
        vac = ent = strength = 0.0
        correct = seen = 0
        with torch.no_grad():
            for xb, yb in loader:
                if seen >= self.max_eval_samples:
                    break
                xb, yb = xb.to(device), yb.to(device)
                alpha = probe(xb)
                S = alpha.sum(dim=-1)
                probs = alpha / S.unsqueeze(-1)
                vac += float((alpha.shape[-1] / S).sum())
                ent += float(-(probs * torch.log(probs + 1e-10)).sum(dim=-1).sum())
                strength += float(S.sum())
                correct += int((alpha.argmax(dim=-1) == yb).sum())
                seen += xb.size(0)
        if seen == 0:
            return {"vacuity": 1.0, "entropy": 0.0, "strength": 0.0, "accuracy": 0.0}
        return {"vacuity": vac / seen, "entropy": ent / seen, "strength": strength / seen,
                "accuracy": correct / seen}

    def record(self, trust: Dict[int, float], metrics: Dict[int, Dict[str, float]],
               accepted: Dict[int, float], threshold: float) -> None:
        if not self.track_statistics:
            return
        st = self._statistics
        st["rounds_processed"] += 1
        st["neighbors_evaluated"] += len(trust)
        st["neighbors_accepted"] += len(accepted)
        st["neighbors_rejected"] += len(trust) - len(accepted)
        if trust:
            st["avg_trust_scores"].append(sum(trust.values()) / len(trust))
        if metrics:
            st["avg_vacuity"].append(sum(m["vacuity"] for m in metrics.values()) / len(metrics))
            st["avg_entropy"].append(sum(m["entropy"] for m in metrics.values()) / len(metrics))
        st["threshold_history"].append(threshold)

    # -- reference-compatible entry point ------------------------------------------
    def aggregate(self, node_id: int, own_state: ModelState, neighbor_states: Dict[int, ModelState],
                  round_num: int, **kwargs) -> ModelState:
        loader = kwargs.get("train_loader")
        template = kwargs.get("model_template")
        device = kwargs.get("device", torch.device("cpu"))
        if loader is None or template is None:
            return average_states([own_state, *neighbor_states.values()])
        threshold = self.current_threshold(round_num)
        trust: Dict[int, float] = {}
        metrics: Dict[int, Dict[str, float]] = {}
        for nid, st in neighbor_states.items():
            metrics[nid] = self._cross_evaluate(st, template, loader, device)
            score = self.score_from_metrics(metrics[nid])
            trust[nid] = self.smooth(nid, score) if self.use_adaptive_trust else score
        accepted = {nid: s for nid, s in trust.items() if s >= threshold}
        self.record(trust, metrics, accepted, threshold)
        if not accepted:
            return own_state
        total = sum(accepted.values())
        peers = average_states([neighbor_states[n] for n in accepted], [s / total for s in accepted.values()])
        return average_states([own_state, peers], [self.self_weight, 1.0 - self.self_weight])

    def get_statistics(self) -> Dict[str, Any]:
        stats = dict(self._statistics)
        ev = stats["neighbors_evaluated"]
        stats["acceptance_rate"] = stats["neighbors_accepted"] / ev if ev else 0.0
        for src, dst in (("avg_trust_scores", "mean_trust"), ("avg_vacuity", "mean_vacuity"),
                         ("avg_entropy", "mean_entropy")):
            stats[dst] = sum(stats[src]) / len(stats[src]) if stats[src] else 0.0
        stats["trust_history_per_neighbor"] = dict(self._trust_history)
        stats["current_smoothed_trust"] = dict(self._smoothed_trust)
        return stats

    def reset_statistics(self) -> None:
        self._trust_history.clear()
        self._smoothed_trust.clear()
        self._statistics = self._fresh_stats()
