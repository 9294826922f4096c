"""LEAF benchmark integration (FEMNIST, CelebA)."""
from murmura_b200.examples.leaf.adapter import load_leaf_adapter
from murmura_b200.examples.leaf.model_factories import get_leaf_model_factory

__all__ = ["load_leaf_adapter", "get_leaf_model_factory"]
