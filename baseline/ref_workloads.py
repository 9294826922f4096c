"""Neutral workload definitions for the reference arm (torch / torchvision / reference package only).

Nothing from ``murmura_b200`` is imported here: the synthetic CIFAR-10-shaped generator below is a
stand-alone copy of the formula in ``murmura_b200/data/synthetic.py`` so both arms train on bit-identical
tensors, the partitioner is the *reference's own* ``murmura.data.partitioners.dirichlet_partition`` and
the model is torchvision's stock ``resnet18(num_classes=10)``.
"""
from __future__ import annotations

import numpy as np
import torch
from torch.utils.data import TensorDataset

SHAPES = {"cifar10": ((3, 32, 32), 10), "mnist": ((784,), 10), "femnist": ((1, 28, 28), 62), "uci_har": ((561,), 6),
          "pamap2": ((4000,), 12), "ppg_dalia": ((192,), 7)}


def synthetic_tensors(name: str, num_samples: int, seed: int = 0, noise: float = 1.0, latent_dim: int = 32,
                      separation: float = 0.6):
    shape, classes = SHAPES[name]
    dim = int(np.prod(shape))
    g = torch.Generator().manual_seed(1_000_003 * seed + 17)
    protos = torch.randn(classes, latent_dim, generator=g) * separation
    lift = torch.randn(latent_dim, dim, generator=g) / latent_dim ** 0.5
    y = torch.randint(0, classes, (num_samples,), generator=g)
    z = protos[y] + noise * torch.randn(num_samples, latent_dim, generator=g)
    x = torch.tanh(z @ lift) + 0.1 * torch.randn(num_samples, dim, generator=g)
    return x.reshape(num_samples, *shape).contiguous(), y


class SyntheticRefAdapter:
    """``data.adapter: baseline.ref_workloads.SyntheticRefAdapter`` (the reference's dotted-path extension point)."""

    def __new__(cls, name="cifar10", num_nodes=8, samples_per_node=512, alpha=0.5, seed=42, partition_method="dirichlet"):
        from murmura.data.adapters import DatasetAdapter
        from murmura.data.partitioners import dirichlet_partition, iid_partition
        x, y = synthetic_tensors(name, num_nodes * samples_per_node, seed=seed)
        if partition_method == "dirichlet":
            parts = dirichlet_partition(y.numpy(), num_nodes, alpha=alpha, min_samples_per_client=2, seed=seed)
        else:
            parts = iid_partition(len(y), num_nodes, seed=seed)
        return DatasetAdapter(TensorDataset(x, y), parts)


def resnet18(num_classes: int = 10):
    import torchvision
    return torchvision.models.resnet18(num_classes=num_classes)


def mlp(input_dim: int = 784, hidden: int = 200, num_classes: int = 10):
    """Stock 2-layer MLP for BASELINE config 1 (plain torch.nn, nothing from this repo)."""
    import torch.nn as nn
    return nn.Sequential(nn.Flatten(), nn.Linear(input_dim, hidden), nn.ReLU(), nn.Linear(hidden, num_classes))


def cifar_cnn(num_classes: int = 10):
    """Stock small CIFAR-10 CNN of BASELINE configs 3 / 4: conv3(3→32)-conv3(32→64)-pool-conv3(64→128)-pool-fc256-fc (plain torch.nn)."""
    import torch.nn as nn
    return nn.Sequential(nn.Conv2d(3, 32, 3, padding=1), nn.ReLU(), nn.Conv2d(32, 64, 3, padding=1), nn.ReLU(), nn.MaxPool2d(2),
                         nn.Conv2d(64, 128, 3, padding=1), nn.ReLU(), nn.MaxPool2d(2), nn.Flatten(), nn.Linear(128 * 8 * 8, 256), nn.ReLU(),
                         nn.Linear(256, num_classes))
