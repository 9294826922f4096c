"""Utilities (device, seeding, metrics).  ``factories`` is imported lazily by callers."""
from murmura_b200.utils.device import get_device
from murmura_b200.utils.seed import set_seed
from murmura_b200.utils.metrics import evaluate_model, compute_accuracy

__all__ = ["get_device", "set_seed", "evaluate_model", "compute_accuracy"]
