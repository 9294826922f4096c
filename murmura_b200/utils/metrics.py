"""Evaluation helpers (parity: reference ``murmura/utils/metrics.py:9-66``)."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.nn as nn


def evaluate_model(model: nn.Module, loader, device: torch.device) -> Tuple[float, float, int, int]:
    """Returns ``(accuracy, mean CE loss, correct, total)`` over ``loader``."""
    model.eval()
    model.to(device)
    loss_sum = torch.zeros((), dtype=torch.float64)
    correct = total = 0
    with torch.no_grad():
        for xb, yb in loader:
            xb, yb = xb.to(device), yb.to(device)
            out = model(xb)
            loss_sum += float(nn.functional.cross_entropy(out, yb)) * xb.size(0)
            correct += int((out.argmax(dim=1) == yb).sum())
            total += yb.size(0)
    if total == 0:
        return 0.0, 0.0, 0, 0
    return correct / total, float(loss_sum) / total, correct, total


def compute_accuracy(predictions: torch.Tensor, targets: torch.Tensor) -> float:
    if predictions.dim() > 1:
        predictions = predictions.argmax(dim=1)
    n = targets.size(0)
    return int((predictions == targets).sum()) / n if n else 0.0
