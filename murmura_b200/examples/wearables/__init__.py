"""Wearable workloads: UCI-HAR, PAMAP2, PPG-DaLiA + evidential MLPs."""
from murmura_b200.examples.wearables.adapter import load_wearable_adapter, get_wearable_dataset_info
from murmura_b200.examples.wearables.models import (create_har_model, create_pamap2_model, create_ppg_dalia_model,
                                                    get_wearable_model_factory, get_evidential_loss,
                                                    EvidentialLoss, EvidentialHead, compute_uncertainty)

__all__ = ["load_wearable_adapter", "get_wearable_dataset_info", "create_har_model", "create_pamap2_model",
           "create_ppg_dalia_model", "get_wearable_model_factory", "get_evidential_loss", "EvidentialLoss",
           "EvidentialHead", "compute_uncertainty"]
