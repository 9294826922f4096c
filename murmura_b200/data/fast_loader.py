"""Dense drop-in for ``DataLoader(TensorDataset-like shard, batch_size, shuffle, drop_last)``.

The reference builds one ``DataLoader`` over a ``Subset`` per node (``murmura/core/network.py:281-294``); every batch is then
assembled sample by sample in Python (index → tuple → ``default_collate``), which is ≈ 60 % of a CPU round for the small models
of the shipped configs.  :class:`FastTensorLoader` slices the node's dense ``(X, y)`` instead and — so that the simulation
backend stays *bit-identical* to the reference (``tests/test_reference_crosscheck.py``) — consumes the global RNG exactly like
the stock loader does:

* ``iter(loader)`` draws one int64 (``_BaseDataLoaderIter._base_seed``),
* with ``shuffle=True`` the first ``next()`` draws one int64 seed for a private generator (``RandomSampler.__iter__``) and the
  epoch order is ``torch.randperm(n, generator=private)``.
"""
from __future__ import annotations

from typing import Iterator, Optional, Tuple

import torch
from torch.utils.data import Dataset, TensorDataset


class FastTensorLoader:
    def __init__(self, x: torch.Tensor, y: torch.Tensor, batch_size: int, shuffle: bool = False, drop_last: bool = False,
                 dataset: Optional[Dataset] = None):
        assert len(x) == len(y)
        self.x, self.y = x, y
        self.batch_size, self.shuffle, self.drop_last = int(batch_size), bool(shuffle), bool(drop_last)
        self.dataset = dataset if dataset is not None else TensorDataset(x, y)

    def __len__(self) -> int:
        n = len(self.x)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        torch.empty((), dtype=torch.int64).random_()                      # the stock iterator's _base_seed draw
        return self._batches()

    def _batches(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        n, bs = len(self.x), self.batch_size
        order = None
        if self.shuffle:
            seed = int(torch.empty((), dtype=torch.int64).random_().item())      # RandomSampler's private generator
            g = torch.Generator()
            g.manual_seed(seed)
            order = torch.randperm(n, generator=g)
        for start in range(0, n, bs):
            stop = start + bs
            if stop > n:
                if self.drop_last:
                    break
                stop = n
            if order is None:
                yield self.x[start:stop], self.y[start:stop]
            else:
                idx = order[start:stop]
                yield self.x.index_select(0, idx), self.y.index_select(0, idx)


def make_loaders(dataset_adapter, client_id: int, batch_size: int):
    """(train_loader, test_loader, n) for one client: dense fast loaders when the adapter wraps a two-tensor dataset, the stock
    ``DataLoader`` otherwise (datasets with per-sample transforms, custom adapters).  Same batch-size / drop-last rule as the
    reference: ``bs = min(batch_size, max(2, n))``, ``drop_last = n > bs``; the test loader walks the *training* shard."""
    from torch.utils.data import DataLoader
    shard = dataset_adapter.get_client_data(client_id)
    n = len(shard)
    bs = min(batch_size, max(2, n))
    tensors = getattr(getattr(dataset_adapter, "dataset", None), "tensors", None)
    if tensors is not None and len(tensors) == 2 and hasattr(dataset_adapter, "client_tensors") and n > 0:
        x, y = dataset_adapter.client_tensors(client_id)
        return (FastTensorLoader(x, y, bs, shuffle=True, drop_last=n > bs, dataset=shard),
                FastTensorLoader(x, y, bs, shuffle=False, dataset=shard), n)
    return DataLoader(shard, batch_size=bs, shuffle=True, drop_last=n > bs), DataLoader(shard, batch_size=bs, shuffle=False), n


def replay_round_orders(sizes, batch_size: int, epochs: int, skip=()) -> dict:
    """Sample orders of ONE round of local training for every client, drawn from the global torch CPU stream exactly as the simulation
    backend's shuffling loaders draw them (clients in id order, ``skip`` = compromised clients, which do not train; per epoch one
    int64 for the iterator's base seed, one int64 seeding a private generator, ``randperm(n)`` from it, truncated to whole batches like
    ``drop_last``).  Used by the B200 engine's seed-parity mode; returns ``{client_id: LongTensor[epochs, nb·eb]}``."""
    out = {}
    for cid, n in enumerate(sizes):
        if cid in skip:
            continue
        n = int(n)
        eb = min(int(batch_size), max(2, n))
        nb = (n // eb) if n > eb else (1 if n >= 2 else 0)
        rows = []
        for _ in range(epochs):
            torch.empty((), dtype=torch.int64).random_()
            if n == 0:
                continue
            g = torch.Generator()
            g.manual_seed(int(torch.empty((), dtype=torch.int64).random_().item()))
            rows.append(torch.randperm(n, generator=g)[: nb * min(eb, n)])
        if rows and nb > 0:
            out[cid] = torch.stack(rows)
    return out
