"""``examples.leaf.<Model>`` factories.

Behaviour of reference ``examples/leaf/model_factories.py:9-31``: the two canonical LEAF class names map to their model with the
dataset's class count as default; any other name selects a FEMNIST capacity variant (``variant=tiny|small|baseline|large|xlarge``).
"""
from __future__ import annotations

from functools import partial
from typing import Callable

import torch.nn as nn

from murmura_b200.models.cnn import LEAFCelebAModel, LEAFFEMNISTModel, get_model_variant

# canonical name → (class, default number of classes)
_CANONICAL = {"LEAFFEMNISTModel": (LEAFFEMNISTModel, 62), "LEAFCelebAModel": (LEAFCelebAModel, 2)}


def get_leaf_model_factory(model_type: str, **kwargs) -> Callable[[], nn.Module]:
    if model_type in _CANONICAL:
        cls, default_classes = _CANONICAL[model_type]
        return partial(cls, num_classes=kwargs.get("num_classes", default_classes))
    return partial(get_model_variant, kwargs.get("variant", "baseline"), num_classes=kwargs.get("num_classes", 62))
