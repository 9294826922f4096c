"""``examples.leaf.<Model>`` factories (parity: reference ``examples/leaf/model_factories.py:9-31``)."""
from __future__ import annotations

from typing import Callable

import torch.nn as nn

from murmura_b200.models.cnn import LEAFCelebAModel, LEAFFEMNISTModel, get_model_variant


def get_leaf_model_factory(model_type: str, **kwargs) -> Callable[[], nn.Module]:
    if model_type == "LEAFFEMNISTModel":
        classes = kwargs.get("num_classes", 62)
        return lambda: LEAFFEMNISTModel(num_classes=classes)
    if model_type == "LEAFCelebAModel":
        classes = kwargs.get("num_classes", 2)
        return lambda: LEAFCelebAModel(num_classes=classes)
    variant = kwargs.get("variant", "baseline")
    classes = kwargs.get("num_classes", 62)
    return lambda: get_model_variant(variant, num_classes=classes)
