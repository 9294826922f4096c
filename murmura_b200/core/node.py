"""One federated participant (simulation / distributed backends).

Parity: reference ``murmura/core/node.py:14-270``: fresh plain-SGD optimiser every round,
batches with <2 samples skipped, evidential criterion receives ``epoch=round_num``,
evaluation returns CE metrics or the evidential set (accuracy, MSE-style loss, vacuity,
entropy, strength), aggregation injects ``train_loader/model_template/device``.
The B200 engine does not instantiate ``Node`` per participant; it keeps all virtual nodes
of a GPU in one flat arena (``murmura_b200.parallel.engine``) but exposes the same metrics.
"""
from __future__ import annotations

from typing import Any, Dict, Optional

import torch
import torch.nn as nn
from torch.optim import SGD

from murmura_b200.aggregation.base import Aggregator, get_model_state, set_model_state
from murmura_b200.core.types import ModelState
from murmura_b200.utils.metrics import evaluate_model


def evidential_batch_stats(alpha: torch.Tensor, targets: torch.Tensor) -> Dict[str, torch.Tensor]:
    """Per-batch *sums* of the evidential evaluation metrics (device tensors, no sync)."""
    S = alpha.sum(dim=-1)
    probs = alpha / S.unsqueeze(-1)
    onehot = torch.zeros_like(alpha).scatter_(1, targets.unsqueeze(1), 1.0)
    return {
        "correct": (alpha.argmax(dim=-1) == targets).sum(),
        "vacuity": (alpha.shape[-1] / S).sum(),
        "entropy": -(probs * torch.log(probs + 1e-10)).sum(dim=-1).sum(),
        "strength": S.sum(),
        "sq_err": ((onehot - probs) ** 2).sum(),
    }


class Node:
    def __init__(self, node_id: int, model: nn.Module, train_loader, test_loader=None,
                 aggregator: Optional[Aggregator] = None, device: Optional[torch.device] = None,
                 criterion: Optional[nn.Module] = None, evidential: bool = False):
        self.node_id = node_id
        self.model = model
        self.train_loader = train_loader
        self.test_loader = test_loader
        self.aggregator = aggregator
        self.device = device or torch.device("cpu")
        self.evidential = evidential
        self.current_round = 0
        self.model.to(self.device)
        self.criterion = criterion if criterion is not None else nn.CrossEntropyLoss()

    # ---- training -----------------------------------------------------------------
    def local_train(self, epochs: int, lr: float = 0.01, round_num: int = 0) -> Dict[str, Any]:
        self.model.train()
        self.model.to(self.device)
        self.current_round = round_num
        opt = SGD(self.model.parameters(), lr=lr)
        running = torch.zeros((), device=self.device)   # accumulate on device: one sync per call
        steps = 0
        for _ in range(epochs):
            for xb, yb in self.train_loader:
                if xb.size(0) < 2:
                    continue
                xb, yb = xb.to(self.device), yb.to(self.device)
                opt.zero_grad()
                out = self.model(xb)
                if self.evidential and hasattr(self.criterion, "forward"):
                    loss = self.criterion(out, yb, epoch=self.current_round)
                else:
                    loss = self.criterion(out, yb)
                loss.backward()
                opt.step()
                running += loss.detach()
                steps += 1
        return {"avg_loss": float(running) / steps if steps else 0.0, "num_batches": steps, "epochs": epochs}

    # ---- evaluation ---------------------------------------------------------------
    def evaluate(self) -> Dict[str, Any]:
        if self.test_loader is None:
            return {"accuracy": 0.0, "loss": 0.0, "note": "No test data available"}
        if self.evidential:
            return self._evaluate_evidential()
        acc, loss, correct, total = evaluate_model(self.model, self.test_loader, self.device)
        return {"accuracy": acc, "loss": loss, "correct": correct, "total": total}

    def _evaluate_evidential(self) -> Dict[str, Any]:
        self.model.eval()
        sums: Dict[str, torch.Tensor] = {}
        total = 0
        with torch.no_grad():
            for xb, yb in self.test_loader:
                xb, yb = xb.to(self.device), yb.to(self.device)
                for k, v in evidential_batch_stats(self.model(xb), yb).items():
                    sums[k] = sums.get(k, 0) + v.double()
                total += yb.size(0)
        if total == 0:
            return {"accuracy": 0.0, "loss": 0.0, "correct": 0, "total": 0,
                    "vacuity": 0.0, "entropy": 0.0, "strength": 0.0}
        host = {k: float(v) for k, v in sums.items()}
        return {"accuracy": host["correct"] / total, "loss": host["sq_err"] / total,
                "correct": int(host["correct"]), "total": total,
                "vacuity": host["vacuity"] / total, "entropy": host["entropy"] / total,
                "strength": host["strength"] / total}

    # ---- state / aggregation ----------------------------------------------------------
    def get_state(self) -> ModelState:
        return get_model_state(self.model)

    def set_state(self, state: ModelState) -> None:
        set_model_state(self.model, state)

    def aggregate_with_neighbors(self, neighbor_states: Dict[int, ModelState], round_num: int,
                                 **kwargs) -> ModelState:
        if self.aggregator is None:
            return self.get_state()
        ctx = dict(kwargs)
        ctx.update(train_loader=self.train_loader, model_template=self.model, device=self.device)
        return self.aggregator.aggregate(node_id=self.node_id, own_state=self.get_state(),
                                         neighbor_states=neighbor_states, round_num=round_num, **ctx)

    def apply_aggregated_state(self, aggregated_state: ModelState) -> None:
        self.set_state(aggregated_state)

    def get_aggregator_statistics(self) -> Dict[str, Any]:
        return self.aggregator.get_statistics() if self.aggregator is not None else {}
