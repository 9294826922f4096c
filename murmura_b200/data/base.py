"""Structural protocol for federated datasets (parity: reference ``data/base.py:8-39``)."""
from __future__ import annotations

from typing import List, Protocol, runtime_checkable

from torch.utils.data import Dataset


@runtime_checkable
class DatasetProtocol(Protocol):
    def get_client_data(self, client_id: int) -> Dataset: ...

    def get_num_clients(self) -> int: ...

    def get_client_partitions(self) -> List[List[int]]: ...
