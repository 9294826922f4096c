"""Plain-PyTorch (fp32/fp64) oracles of every CUDA kernel — used by ``tests/`` for numerics checks."""
from __future__ import annotations

from typing import List, Tuple

import numpy as np
import torch


def weighted_gather(live: torch.Tensor, pub: torch.Tensor, rows: List[List[Tuple[int, float]]], self_w: List[float],
                    length: int) -> torch.Tensor:
    """``out[v,:length] = self_w[v]*live[v] + Σ (slot,w) w*pub[slot]``; rest of the row untouched."""
    out = live.clone()
    for v, srcs in enumerate(rows):
        acc = self_w[v] * live[v, :length].double()
        for slot, w in srcs:
            acc = acc + w * pub[slot, :length].double()
        out[v, :length] = acc.float()
    return out


def edge_sq_distances(live: torch.Tensor, pub: torch.Tensor, v: int, slots: List[int], length: int) -> torch.Tensor:
    own = live[v, :length].double()
    return torch.stack([((own - pub[s, :length].double()) ** 2).sum() for s in slots]).float()


def pairwise_sq(cands: torch.Tensor) -> torch.Tensor:
    c = cands.double()
    return torch.cdist(c, c).pow(2).float()


def count_sketch(vec: torch.Tensor, buckets: np.ndarray, signs: np.ndarray, K: int) -> torch.Tensor:
    v = vec.detach().cpu().double().numpy()
    return torch.from_numpy(np.bincount(buckets[: len(v)], weights=signs[: len(v)] * v, minlength=K)).float()


def mxfp8_roundtrip(x: torch.Tensor) -> torch.Tensor:
    """Block-scaled (32-element, power-of-two scale) e4m3 quantise→dequantise."""
    K = x.shape[-1]
    pad = (-K) % 32
    xp = torch.nn.functional.pad(x.float(), (0, pad)).view(*x.shape[:-1], -1, 32)
    amax = xp.abs().amax(dim=-1, keepdim=True)
    e = torch.where(amax > 0, torch.frexp(amax / 448.0).exponent.float(), torch.zeros_like(amax))
    scale = torch.exp2(e.clamp(-127, 127))
    q = (xp / scale).to(torch.float8_e4m3fn).float() * scale
    return q.view(*x.shape[:-1], -1)[..., :K]


def ce_stats(logits: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
    loss = torch.nn.functional.cross_entropy(logits.double(), targets, reduction="sum")
    correct = (logits.argmax(dim=1) == targets).sum()
    return torch.tensor([float(loss), float(correct), float(len(targets))])


def dirichlet_stats(alpha: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
    a = alpha.double()
    S = a.sum(-1)
    p = a / S.unsqueeze(-1)
    onehot = torch.nn.functional.one_hot(targets, a.shape[-1]).double()
    return torch.tensor([float(((onehot - p) ** 2).sum()), float((a.argmax(-1) == targets).sum()), float(len(targets)),
                         float((a.shape[-1] / S).sum()), float(-(p * torch.log(p + 1e-10)).sum()), float(S.sum())])
