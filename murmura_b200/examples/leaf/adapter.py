"""``leaf.<dataset>`` adapter (parity: reference ``examples/leaf/adapter.py:19-61``)."""
from __future__ import annotations

import os

from murmura_b200.data.adapters import DatasetAdapter
from murmura_b200.examples.leaf.datasets import create_leaf_client_partitions, load_leaf_dataset


def load_leaf_adapter(dataset_type: str, **kwargs) -> DatasetAdapter:
    data_path = kwargs.get("data_path", f"leaf/data/{dataset_type}/data")
    split = kwargs.get("split", "train")
    max_samples = kwargs.get("max_samples")
    num_nodes = kwargs.get("num_nodes") or kwargs.get("num_clients")
    seed = kwargs.get("seed", 42)
    if num_nodes is None:
        raise ValueError("num_nodes is required for LEAF adapter (pass topology.num_nodes).")
    if data_path == "synthetic" or (kwargs.get("allow_synthetic") and not os.path.exists(data_path)):
        from murmura_b200.data.synthetic import SyntheticAdapter
        return SyntheticAdapter(name=dataset_type.lower(), num_nodes=num_nodes, seed=seed, max_samples=max_samples,
                                samples_per_node=kwargs.get("samples_per_node", 256),
                                partition_method=kwargs.get("partition_method", "dirichlet"),
                                alpha=kwargs.get("alpha", 0.5))
    train_ds, test_ds, _, _, _ = load_leaf_dataset(dataset_name=dataset_type, data_path=data_path)
    train_parts, test_parts = create_leaf_client_partitions(train_ds, test_ds, num_nodes=num_nodes, seed=seed)
    dataset, parts = (train_ds, train_parts) if split == "train" else (test_ds, test_parts)
    if max_samples is not None:
        parts = [p[:max_samples] for p in parts]
    return DatasetAdapter(dataset=dataset, client_partitions=parts)
