"""Federated data: adapters, partitioners, synthetic workloads."""
from murmura_b200.data.base import DatasetProtocol
from murmura_b200.data.adapters import DatasetAdapter, TorchDatasetAdapter
from murmura_b200.data.partitioners import (dirichlet_partition, iid_partition, natural_partition,
                                            combine_partitions_with_dirichlet)
from murmura_b200.data.synthetic import SyntheticAdapter, load_synthetic_adapter, make_synthetic_tensors

__all__ = ["DatasetProtocol", "DatasetAdapter", "TorchDatasetAdapter", "dirichlet_partition",
           "iid_partition", "natural_partition", "combine_partitions_with_dirichlet",
           "SyntheticAdapter", "load_synthetic_adapter", "make_synthetic_tensors"]
