"""Aggregator base class and model-state helpers.

Parity: reference ``murmura/aggregation/base.py:9-170``.  These run on whatever device
the states live on (the reference forces CPU clones; the ``simulation`` backend here
keeps that behaviour through :func:`get_model_state`).  The B200 engine never calls
these per-key helpers on its hot path — it uses the flat-arena kernels in
``murmura_b200.ops`` — but each aggregator's ``plan()`` (see subclasses) and these
helpers share the same semantics and are cross-checked in ``tests/``.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Any, Dict, Optional, Sequence

import torch

from murmura_b200.core.types import ModelState


class Aggregator(ABC):
    """Combine a node's own state with the states received from its neighbours.

    Unknown keyword arguments are swallowed into ``self.config`` (reference
    ``aggregation/base.py:16-18``): many shipped YAMLs carry inert knobs.
    """

    #: name used by the B200 engine to pick the fused kernel plan
    kernel_family: str = "generic"

    def __init__(self, **kwargs: Any):
        self.config = kwargs

    @abstractmethod
    def aggregate(self, node_id: int, own_state: ModelState, neighbor_states: Dict[int, ModelState],
                  round_num: int, **kwargs: Any) -> ModelState:
        ...

    def get_statistics(self) -> Dict[str, Any]:
        return {}


# ---- helpers ------------------------------------------------------------------------

def get_model_state(model: torch.nn.Module) -> ModelState:
    """CPU snapshot of ``model.state_dict()`` (reference ``aggregation/base.py:54-63``)."""
    return {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}


def set_model_state(model: torch.nn.Module, state: ModelState) -> None:
    model.load_state_dict(state)


def average_states(states: Sequence[ModelState], weights: Optional[Sequence[float]] = None) -> ModelState:
    """Weighted mean of float tensors; non-float tensors are taken from ``states[0]``."""
    if not states:
        raise ValueError("Cannot average empty list of states")
    if weights is None:
        weights = [1.0 / len(states)] * len(states)
    elif len(weights) != len(states):
        raise ValueError(f"weights length ({len(weights)}) != states length ({len(states)})")
    elif abs(sum(weights) - 1.0) >= 1e-6:
        raise ValueError(f"weights must sum to 1.0, got {sum(weights)}")
    out: ModelState = {}
    for key, first in states[0].items():
        if not first.is_floating_point():
            out[key] = first.clone()
            continue
        acc = torch.zeros_like(first)
        for st, w in zip(states, weights):
            acc.add_(st[key], alpha=float(w))
        out[key] = acc
    return out


def blend_states(own: ModelState, others: ModelState, alpha: float) -> ModelState:
    """``alpha*own + (1-alpha)*others`` on *every* key (int buffers become float)."""
    return {k: alpha * own[k] + (1 - alpha) * others[k] for k in own}


def compute_model_distance(state1: ModelState, state2: ModelState) -> float:
    """L2 distance over float tensors only (Krum's metric, reference ``:118-135``)."""
    total = 0.0
    for key, a in state1.items():
        if a.is_floating_point():
            total += float(torch.sum((a.float() - state2[key].float()) ** 2))
    return total ** 0.5


def squared_distance_all_keys(state1: ModelState, state2: ModelState) -> float:
    """Σ (a-b)² over **all** common tensors, ints included (BALANCE/UBAR metric)."""
    total = 0.0
    for key in state1.keys() & state2.keys():
        diff = state1[key] - state2[key]
        total += float(torch.sum(diff * diff))
    return total


def squared_norm_all_keys(state: ModelState) -> float:
    return sum(float(torch.sum(t * t)) for t in state.values() if t.numel() > 0)


def flatten_model_state(state: ModelState) -> torch.Tensor:
    parts = [t.flatten().float() for t in state.values() if t.is_floating_point()]
    if not parts:
        raise ValueError("No floating-point parameters found in model state")
    return torch.cat(parts)


def calculate_model_dimension(model: torch.nn.Module) -> int:
    return sum(t.numel() for t in model.state_dict().values() if t.is_floating_point())
