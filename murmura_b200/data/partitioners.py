"""Client partitioners: Dirichlet non-IID, IID, natural ids, natural→Dirichlet.

Parity: reference ``murmura/data/partitioners.py:7-223`` including the use of the *global*
NumPy RNG (``np.random.seed(seed)``) and the draw order (per class: Dirichlet proportions →
remainder clients → class shuffle; then min-sample rebalancing; then per-client shuffles),
so a given ``(labels, seed)`` yields the same shards as the reference.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np


def _rebalance(parts: List[List[int]], floor: int) -> None:
    """Move tail samples from the richest-first donors to clients below ``floor``."""
    if floor <= 0:
        return
    need = {i: floor - len(p) for i, p in enumerate(parts) if len(p) < floor}
    if not need:
        return
    spare = {i: len(p) - floor for i, p in enumerate(parts) if len(p) > floor}
    for poor, missing in need.items():
        for donor in spare:
            if missing <= 0:
                break
            take = min(missing, spare[donor])
            if take <= 0:
                continue
            parts[poor].extend(parts[donor][-take:])
            del parts[donor][-take:]
            spare[donor] -= take
            missing -= take


def dirichlet_partition(labels: np.ndarray, num_clients: int, alpha: float = 0.5,
                        min_samples_per_client: int = 1, seed: Optional[int] = None) -> List[List[int]]:
    """Label-skewed split: for each class draw client proportions from ``Dir(alpha)``."""
    if seed is not None:
        np.random.seed(seed)
    labels = np.asarray(labels)
    parts: List[List[int]] = [[] for _ in range(num_clients)]
    for cls in np.unique(labels):
        members = np.where(labels == cls)[0]
        share = np.random.dirichlet(np.repeat(alpha, num_clients))
        share = share / share.sum()
        counts = (share * len(members)).astype(int)
        leftover = len(members) - counts.sum()
        if leftover > 0:
            counts[np.random.choice(num_clients, leftover, replace=False)] += 1
        np.random.shuffle(members)
        bounds = np.concatenate([[0], np.cumsum(counts)])
        for cid in range(num_clients):
            parts[cid].extend(members[bounds[cid]:bounds[cid + 1]].tolist())
    _rebalance(parts, min_samples_per_client)
    for p in parts:
        np.random.shuffle(p)
    return parts


def iid_partition(num_samples: int, num_clients: int, seed: Optional[int] = None) -> List[List[int]]:
    if seed is not None:
        np.random.seed(seed)
    order = np.arange(num_samples)
    np.random.shuffle(order)
    return [chunk.tolist() for chunk in np.array_split(order, num_clients)]


def natural_partition(client_ids: np.ndarray, num_clients: Optional[int] = None) -> Tuple[List[List[int]], int]:
    """One shard per distinct id (first ``num_clients`` ids in sorted order)."""
    ids = np.unique(client_ids)
    if num_clients is not None and num_clients < len(ids):
        ids = ids[:num_clients]
    return [np.where(client_ids == cid)[0].tolist() for cid in ids], len(ids)


def combine_partitions_with_dirichlet(natural_partitions: List[List[int]], labels: np.ndarray,
                                      num_clients: int, alpha: float = 0.5,
                                      seed: Optional[int] = None) -> List[List[int]]:
    pool = [i for part in natural_partitions for i in part]
    local = dirichlet_partition(labels=np.asarray(labels)[pool], num_clients=num_clients, alpha=alpha, seed=seed)
    return [[pool[i] for i in part] for part in local]
