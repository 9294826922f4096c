"""Utilities (device, seeding, metrics, factories)."""
from murmura_b200._lazy import lazy_exports

__getattr__, __dir__, __all__ = lazy_exports(__name__, {
    "device": ["get_device"],
    "seed": ["set_seed"],
    "metrics": ["evaluate_model", "compute_accuracy"],
})
