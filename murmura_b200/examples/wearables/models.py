"""Evidential wearable models + factories (parity: reference ``examples/wearables/models.py:355-481``).

The module classes live in :mod:`murmura_b200.models.mlp`; this file provides the
reference's factory names.
"""
from __future__ import annotations

from typing import Callable, Sequence

import torch.nn as nn

from murmura_b200.models.mlp import (EvidentialHARClassifier, EvidentialHead, EvidentialLoss,
                                     EvidentialPAMAP2Classifier, EvidentialPPGDaLiAClassifier,
                                     compute_uncertainty)


def create_har_model(input_dim: int = 561, hidden_dims: Sequence[int] = (256, 128), num_classes: int = 6,
                     dropout: float = 0.3) -> nn.Module:
    return EvidentialHARClassifier(input_dim, tuple(hidden_dims), num_classes, dropout)


def create_pamap2_model(input_dim: int = 4000, hidden_dims: Sequence[int] = (512, 256, 128),
                        num_classes: int = 12, dropout: float = 0.3) -> nn.Module:
    return EvidentialPAMAP2Classifier(input_dim, tuple(hidden_dims), num_classes, dropout)


def create_ppg_dalia_model(input_dim: int = 192, hidden_dims: Sequence[int] = (256, 128, 64),
                           num_classes: int = 7, dropout: float = 0.3) -> nn.Module:
    return EvidentialPPGDaLiAClassifier(input_dim, tuple(hidden_dims), num_classes, dropout)


_CREATORS = {"uci_har": create_har_model, "pamap2": create_pamap2_model, "ppg_dalia": create_ppg_dalia_model}


def get_wearable_model_factory(dataset_type: str, **kwargs) -> Callable[[], nn.Module]:
    kind = dataset_type.lower().replace("-", "_")
    if kind not in _CREATORS:
        raise ValueError(f"Unknown dataset type: {dataset_type}. Available: {sorted(_CREATORS)}")
    make = _CREATORS[kind]
    return lambda: make(**kwargs)


def get_evidential_loss(num_classes: int, annealing_epochs: int = 10, lambda_weight: float = 1.0) -> EvidentialLoss:
    return EvidentialLoss(num_classes=num_classes, annealing_epochs=annealing_epochs, lambda_weight=lambda_weight)


__all__ = ["create_har_model", "create_pamap2_model", "create_ppg_dalia_model", "get_wearable_model_factory",
           "get_evidential_loss", "EvidentialLoss", "EvidentialHead", "compute_uncertainty",
           "EvidentialHARClassifier", "EvidentialPAMAP2Classifier", "EvidentialPPGDaLiAClassifier"]
