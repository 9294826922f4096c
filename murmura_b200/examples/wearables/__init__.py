"""Wearable workloads: UCI-HAR, PAMAP2, PPG-DaLiA + evidential MLPs."""
from murmura_b200._lazy import lazy_exports

__getattr__, __dir__, __all__ = lazy_exports(__name__, {
    "adapter": ["load_wearable_adapter", "get_wearable_dataset_info"],
    "models": ["create_har_model", "create_pamap2_model", "create_ppg_dalia_model", "get_wearable_model_factory", "get_evidential_loss", "EvidentialLoss", "EvidentialHead", "compute_uncertainty"],
})
