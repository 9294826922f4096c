"""``wearables.<dataset>`` adapters (parity: reference ``examples/wearables/adapter.py:18-211``)."""
from __future__ import annotations

from typing import Optional

from murmura_b200.data.adapters import DatasetAdapter
from murmura_b200.data.partitioners import dirichlet_partition, iid_partition, natural_partition
from murmura_b200.examples.wearables.datasets import PAMAP2Dataset, PPGDaLiADataset, UCIHARDataset

_INFO = {
    "uci_har": {"name": "UCI Human Activity Recognition", "num_features": 561, "num_classes": 6,
                "natural_clients": 30, "activities": list(UCIHARDataset.ACTIVITIES.values()),
                "description": "Smartphone-based activity recognition from 30 subjects"},
    "pamap2": {"name": "PAMAP2 Physical Activity Monitoring", "num_features": "variable (window_size * 40)",
               "num_classes": 12, "natural_clients": 9, "activities": list(PAMAP2Dataset.ACTIVITY_NAMES.values()),
               "description": "IMU-based activity recognition from 9 subjects"},
    "ppg_dalia": {"name": "PPG-DaLiA Activity Recognition", "num_features": "variable (window_size * 6)",
                  "num_classes": 7, "natural_clients": 15,
                  "activities": list(PPGDaLiADataset.ACTIVITY_NAMES.values()),
                  "description": "PPG/wearable-based activity recognition from 15 subjects"},
}


def _canon(name: str) -> str:
    return name.lower().replace("-", "_")


def _load_dataset(dataset_type: str, data_path: str, split: str, **kw):
    kind = _canon(dataset_type)
    if kind == "uci_har":
        ds = UCIHARDataset(root=data_path, split=split, normalize=kw.get("normalize", True))
    elif kind == "pamap2":
        ds = PAMAP2Dataset(root=data_path, subjects=kw.get("subjects"), activities=kw.get("activities"),
                           window_size=kw.get("window_size", 100), window_stride=kw.get("window_stride", 50),
                           normalize=kw.get("normalize", True),
                           include_heart_rate=kw.get("include_heart_rate", True))
    elif kind == "ppg_dalia":
        ds = PPGDaLiADataset(root=data_path, subjects=kw.get("subjects"), activities=kw.get("activities"),
                             window_size=kw.get("window_size", 32), window_stride=kw.get("window_stride", 16),
                             normalize=kw.get("normalize", True), use_wrist_only=kw.get("use_wrist_only", True))
    else:
        raise ValueError(f"Unknown dataset type: {dataset_type}. Available: 'uci_har', 'pamap2', 'ppg_dalia'")
    return ds, ds.get_labels(), ds.get_subjects()


def load_wearable_adapter(dataset_type: str, data_path: str, num_nodes: int, partition_method: str = "dirichlet",
                          alpha: float = 0.5, seed: int = 42, split: str = "train",
                          max_samples: Optional[int] = None, **kwargs) -> DatasetAdapter:
    """Load a wearable dataset and split it across ``num_nodes`` clients.

    ``data_path: synthetic`` (or a path that does not exist together with
    ``allow_synthetic: true``) swaps in the synthetic generator of the same shape — the GPU
    boxes carry no datasets.
    """
    if data_path == "synthetic" or kwargs.pop("allow_synthetic", False) and not __import__("os").path.exists(data_path):
        from murmura_b200.data.synthetic import SyntheticAdapter
        return SyntheticAdapter(name=_canon(dataset_type), num_nodes=num_nodes, partition_method=partition_method
                                if partition_method != "natural" else "iid", alpha=alpha, seed=seed,
                                max_samples=max_samples, samples_per_node=kwargs.pop("samples_per_node", 512))
    dataset, labels, natural_ids = _load_dataset(dataset_type, data_path, split, **kwargs)
    if partition_method == "dirichlet":
        parts = dirichlet_partition(labels=labels, num_clients=num_nodes, alpha=alpha, seed=seed)
    elif partition_method == "iid":
        parts = iid_partition(num_samples=len(dataset), num_clients=num_nodes, seed=seed)
    elif partition_method == "natural":
        if natural_ids is None:
            raise ValueError(f"Dataset '{dataset_type}' does not support natural partitioning")
        parts, found = natural_partition(client_ids=natural_ids, num_clients=num_nodes)
        if found < num_nodes:
            print(f"Warning: Only {found} natural clients available, requested {num_nodes}")
    else:
        raise ValueError(f"Unknown partition method: {partition_method}")
    if max_samples is not None:
        parts = [p[:max_samples] for p in parts]
    return DatasetAdapter(dataset=dataset, client_partitions=parts)


def get_wearable_dataset_info(dataset_type: str) -> dict:
    kind = _canon(dataset_type)
    if kind not in _INFO:
        raise ValueError(f"Unknown dataset type: {dataset_type}")
    return _INFO[kind]
