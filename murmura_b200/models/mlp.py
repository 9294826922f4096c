"""MLP families: plain 2-layer MLP (BASELINE config 1) and the evidential wearables MLPs.

Architectures follow reference ``murmura/examples/wearables/models.py:187-347``:
``[Linear → BatchNorm1d → ReLU → Dropout(0.3)] × k → EvidentialHead`` with
``alpha = softplus(Wx+b) + 1`` (``:18-46``).  Module/parameter names are chosen so the
state-dict key order (``feature_extractor.{4i}.weight …``, ``evidential_head.fc.*``) and the
float-state sizes of SURVEY §9 (HAR 179,078; PAMAP2 2,217,868; PPG 92,807) match.
"""
from __future__ import annotations

from typing import Dict, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from murmura_b200.ops import bn_act


class MLP(nn.Module):
    """Plain ``in → hidden… → classes`` ReLU MLP (default 784-200-10, P = 159,010)."""

    def __init__(self, input_dim: int = 784, hidden_dims: Sequence[int] = (200,), num_classes: int = 10):
        super().__init__()
        dims = [input_dim, *hidden_dims]
        layers = []
        for a, b in zip(dims[:-1], dims[1:]):
            layers += [nn.Linear(a, b), nn.ReLU()]
        layers.append(nn.Linear(dims[-1], num_classes))
        self.net = nn.Sequential(*layers)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.net(x.flatten(1))


class EvidentialHead(nn.Module):
    """Linear layer emitting Dirichlet concentrations ``alpha = softplus(z) + 1``."""

    def __init__(self, in_features: int, num_classes: int):
        super().__init__()
        self.fc = nn.Linear(in_features, num_classes)
        self.num_classes = num_classes

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return F.softplus(self.fc(x)) + 1


def compute_uncertainty(alpha: torch.Tensor) -> Dict[str, torch.Tensor]:
    """probs / vacuity (K/S) / entropy / dissonance(=entropy) / strength of a Dirichlet."""
    S = alpha.sum(dim=-1, keepdim=True)
    probs = alpha / S
    entropy = -(probs * torch.log(probs + 1e-10)).sum(dim=-1)
    return {"probs": probs, "vacuity": alpha.shape[-1] / S.squeeze(-1), "dissonance": entropy,
            "entropy": entropy, "strength": S.squeeze(-1)}


class EvidentialMLP(nn.Module):
    """Shared body of the three wearable classifiers."""

    def __init__(self, input_dim: int, hidden_dims: Sequence[int], num_classes: int, dropout: float = 0.3):
        super().__init__()
        blocks = []
        prev = input_dim
        for h in hidden_dims:
            blocks += [nn.Linear(prev, h), nn.BatchNorm1d(h), nn.ReLU(), nn.Dropout(dropout)]
            prev = h
        self.feature_extractor = nn.Sequential(*blocks)
        self.evidential_head = EvidentialHead(prev, num_classes)
        self.num_classes = num_classes
        self.input_dim = input_dim
        self.hidden_dims = tuple(hidden_dims)
        self.dropout = dropout

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        # Same module sequence as ``self.feature_extractor(x)``; BatchNorm1d → ReLU pairs go through ``ops.bn_act``
        # (one fused launch per direction on sm_100a, the stock ops anywhere else).
        mods = list(self.feature_extractor)
        i = 0
        while i < len(mods):
            m = mods[i]
            if isinstance(m, nn.BatchNorm1d) and i + 1 < len(mods) and isinstance(mods[i + 1], nn.ReLU):
                x = bn_act(x, m, relu=True)
                i += 2
            else:
                x = m(x)
                i += 1
        return self.evidential_head(x)

    def predict(self, x: torch.Tensor) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
        alpha = self.forward(x)
        return alpha.argmax(dim=-1), compute_uncertainty(alpha)


class EvidentialHARClassifier(EvidentialMLP):
    def __init__(self, input_dim: int = 561, hidden_dims: Sequence[int] = (256, 128), num_classes: int = 6,
                 dropout: float = 0.3):
        super().__init__(input_dim, hidden_dims, num_classes, dropout)


class EvidentialPAMAP2Classifier(EvidentialMLP):
    def __init__(self, input_dim: int = 4000, hidden_dims: Sequence[int] = (512, 256, 128),
                 num_classes: int = 12, dropout: float = 0.3):
        super().__init__(input_dim, hidden_dims, num_classes, dropout)


class EvidentialPPGDaLiAClassifier(EvidentialMLP):
    def __init__(self, input_dim: int = 192, hidden_dims: Sequence[int] = (256, 128, 64),
                 num_classes: int = 7, dropout: float = 0.3):
        super().__init__(input_dim, hidden_dims, num_classes, dropout)


class EvidentialLoss(nn.Module):
    """``MSE(y, alpha/S) + lambda_t · KL(Dir(alpha~) ‖ Dir(1))`` (Sensoy et al. 2018).

    ``alpha~ = y + (1-y)·alpha`` and ``lambda_t = min(1, epoch/annealing_epochs)·lambda_weight``
    (reference ``examples/wearables/models.py:89-179``).  On CUDA with the extension built the
    forward+backward run as one fused kernel (``ops.evidential_loss``).
    """

    def __init__(self, num_classes: int, annealing_epochs: int = 10, lambda_weight: float = 1.0):
        super().__init__()
        self.num_classes = num_classes
        self.annealing_epochs = annealing_epochs
        self.lambda_weight = lambda_weight

    def anneal(self, epoch: int) -> float:
        return min(1.0, epoch / max(1, self.annealing_epochs)) * self.lambda_weight

    def forward(self, alpha: torch.Tensor, targets: torch.Tensor, epoch: int = 0) -> torch.Tensor:
        lam = self.anneal(epoch)
        if alpha.is_cuda:
            from murmura_b200 import ops
            if ops.available():
                return ops.evidential_loss(alpha, targets, lam)
        return evidential_loss_reference(alpha, targets, lam)

    def _kl_divergence(self, alpha: torch.Tensor) -> torch.Tensor:
        return dirichlet_kl_to_uniform(alpha).mean()


def dirichlet_kl_to_uniform(alpha: torch.Tensor) -> torch.Tensor:
    K = alpha.shape[-1]
    S = alpha.sum(dim=-1)
    return (torch.lgamma(S) - torch.lgamma(torch.tensor(float(K), device=alpha.device))
            - torch.lgamma(alpha).sum(dim=-1)
            + ((alpha - 1) * (torch.digamma(alpha) - torch.digamma(S).unsqueeze(-1))).sum(dim=-1))


def evidential_loss_reference(alpha: torch.Tensor, targets: torch.Tensor, lam: float) -> torch.Tensor:
    """Plain-PyTorch oracle of the evidential loss (also the CPU path)."""
    y = F.one_hot(targets, alpha.shape[-1]).to(alpha.dtype)
    p = alpha / alpha.sum(dim=-1, keepdim=True)
    mse = ((y - p) ** 2).sum(dim=-1).mean()
    kl = dirichlet_kl_to_uniform(y + (1 - y) * alpha).mean()
    return mse + lam * kl
