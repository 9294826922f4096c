"""Synthetic, learnable, non-IID workloads with the shapes named in BASELINE.json.

The GPU boxes have no network and the reference's datasets are not shipped, so
every benchmark config uses these generators: class-conditional Gaussian prototypes pushed
through a fixed random nonlinearity (so a network must actually learn), partitioned with the
same Dirichlet/IID partitioners as real data.  Shapes: ``mnist`` 784-d/10, ``cifar10``
3×32×32/10, ``femnist`` 1×28×28/62, ``celeba`` 3×84×84/2, ``uci_har`` 561/6, ``pamap2``
4000/12, ``ppg_dalia`` 192/7.

Used via the dotted adapter path (the reference's extension point, ``utils/factories.py:40-42``)
``data.adapter: murmura_b200.data.synthetic.SyntheticAdapter`` or the short alias
``synthetic.<name>``.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np
import torch
from torch.utils.data import TensorDataset

from murmura_b200.data.adapters import DatasetAdapter
from murmura_b200.data.partitioners import dirichlet_partition, iid_partition

SHAPES: Dict[str, Tuple[Tuple[int, ...], int]] = {
    "mnist": ((784,), 10),
    "mnist_image": ((1, 28, 28), 10),
    "cifar10": ((3, 32, 32), 10),
    "femnist": ((1, 28, 28), 62),
    "celeba": ((3, 84, 84), 2),
    "uci_har": ((561,), 6),
    "pamap2": ((4000,), 12),
    "ppg_dalia": ((192,), 7),
}


def make_synthetic_tensors(name: str, num_samples: int, seed: int = 0, noise: float = 1.0,
                           latent_dim: int = 32, separation: float = 0.6) -> Tuple[torch.Tensor, torch.Tensor]:
    """Deterministic ``(X float32, y int64)`` for workload ``name``."""
    if name not in SHAPES:
        raise ValueError(f"unknown synthetic workload '{name}' (have {sorted(SHAPES)})")
    shape, classes = SHAPES[name]
    dim = int(np.prod(shape))
    g = torch.Generator().manual_seed(1_000_003 * seed + 17)
    protos = torch.randn(classes, latent_dim, generator=g) * separation
    lift = torch.randn(latent_dim, dim, generator=g) / latent_dim ** 0.5
    y = torch.randint(0, classes, (num_samples,), generator=g)
    z = protos[y] + noise * torch.randn(num_samples, latent_dim, generator=g)
    x = torch.tanh(z @ lift) + 0.1 * torch.randn(num_samples, dim, generator=g)
    return x.reshape(num_samples, *shape).contiguous(), y


class SyntheticAdapter(DatasetAdapter):
    """``SyntheticAdapter(name=..., num_nodes=..., samples_per_node=...)``.

    ``partition_method`` ∈ {"dirichlet", "iid"}; ``alpha`` is the Dirichlet concentration.
    ``min_samples_per_client`` defaults to 2 so BatchNorm never sees a single-sample shard.
    """

    def __init__(self, name: str = "mnist", num_nodes: int = 8, samples_per_node: int = 512,
                 partition_method: str = "dirichlet", alpha: float = 0.5, seed: int = 42,
                 noise: float = 1.0, separation: float = 0.6, min_samples_per_client: int = 2,
                 max_samples: Optional[int] = None, **_unused):
        x, y = make_synthetic_tensors(name, num_nodes * samples_per_node, seed=seed, noise=noise,
                                      separation=separation)
        if partition_method == "dirichlet":
            parts = dirichlet_partition(y.numpy(), num_nodes, alpha=alpha,
                                        min_samples_per_client=min_samples_per_client, seed=seed)
        elif partition_method == "iid":
            parts = iid_partition(len(y), num_nodes, seed=seed)
        else:
            raise ValueError(f"Unknown partition method: {partition_method}")
        if max_samples is not None:
            parts = [p[:max_samples] for p in parts]
        super().__init__(TensorDataset(x, y), parts)
        self.name = name
        self.input_shape, self.num_classes = SHAPES[name]


def load_synthetic_adapter(name: str, num_nodes: int, seed: int = 42, **params) -> SyntheticAdapter:
    return SyntheticAdapter(name=name, num_nodes=num_nodes, seed=seed, **params)
