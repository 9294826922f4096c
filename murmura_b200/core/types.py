"""Shared type aliases and the structural model protocol.

Parity: reference ``murmura/core/types.py:8-45`` (``ModelState``, ``DataPartition``,
``ModelProtocol``).  In the B200 engine a ``ModelState`` is usually a *view* dict
into one slot of the flat parameter arena (see ``murmura_b200.parallel.arena``).
"""
from __future__ import annotations

from typing import Any, Dict, Iterator, List, Protocol, runtime_checkable

import torch

ModelState = Dict[str, torch.Tensor]
DataPartition = List[int]


@runtime_checkable
class ModelProtocol(Protocol):
    """Anything that behaves like an ``nn.Module`` as far as the engine cares."""

    def forward(self, x: torch.Tensor) -> torch.Tensor: ...

    def state_dict(self) -> Dict[str, Any]: ...

    def load_state_dict(self, state_dict: Dict[str, Any]) -> Any: ...

    def parameters(self) -> Iterator[torch.nn.Parameter]: ...

    def train(self, mode: bool = True) -> Any: ...

    def eval(self) -> Any: ...

    def to(self, device: Any) -> Any: ...
