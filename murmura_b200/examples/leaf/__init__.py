"""LEAF benchmark integration (FEMNIST, CelebA)."""
from murmura_b200._lazy import lazy_exports

__getattr__, __dir__, __all__ = lazy_exports(__name__, {
    "adapter": ["load_leaf_adapter"],
    "model_factories": ["get_leaf_model_factory"],
})
