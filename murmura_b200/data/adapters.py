"""Dataset adapter: a torch ``Dataset`` plus per-client index lists.

Parity: reference ``murmura/data/adapters.py:7-57``.  ``client_tensors`` is the B200
engine's fast path: it materialises one client's shard as a pair of dense tensors so the
shard can live on the GPU for the whole run (no DataLoader worker, no per-batch H2D).
"""
from __future__ import annotations

from typing import List, Tuple

import torch
from torch.utils.data import Dataset, Subset


class DatasetAdapter:
    def __init__(self, dataset: Dataset, client_partitions: List[List[int]]):
        self.dataset = dataset
        self.client_partitions = client_partitions
        self.num_clients = len(client_partitions)

    def get_client_data(self, client_id: int) -> Dataset:
        if not 0 <= client_id < self.num_clients:
            raise ValueError(f"client_id {client_id} out of range [0, {self.num_clients})")
        return Subset(self.dataset, self.client_partitions[client_id])

    def get_num_clients(self) -> int:
        return self.num_clients

    def get_client_partitions(self) -> List[List[int]]:
        return self.client_partitions

    # ---- B200 engine fast path ---------------------------------------------------
    def client_tensors(self, client_id: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """Dense ``(X, y)`` for one client.  Uses ``dataset.tensors`` when the wrapped
        dataset exposes them (all bundled datasets do), else falls back to indexing."""
        idx = self.client_partitions[client_id]
        tensors = getattr(self.dataset, "tensors", None)
        if tensors is not None and len(tensors) == 2:
            sel = torch.as_tensor(idx, dtype=torch.long)
            return tensors[0][sel], tensors[1][sel]
        xs, ys = zip(*(self.dataset[i] for i in idx)) if idx else ((), ())
        return torch.stack([torch.as_tensor(x) for x in xs]), torch.as_tensor(ys, dtype=torch.long)


class TorchDatasetAdapter(DatasetAdapter):
    """Alias kept for API parity."""
