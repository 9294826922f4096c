"""Federated data: adapters, partitioners, synthetic workloads, dense loaders."""
from murmura_b200._lazy import lazy_exports

__getattr__, __dir__, __all__ = lazy_exports(__name__, {
    "base": ["DatasetProtocol"],
    "adapters": ["DatasetAdapter", "TorchDatasetAdapter"],
    "partitioners": ["dirichlet_partition", "iid_partition", "natural_partition", "combine_partitions_with_dirichlet"],
    "synthetic": ["SyntheticAdapter", "load_synthetic_adapter", "make_synthetic_tensors"],
    "fast_loader": ["FastTensorLoader", "make_loaders"],
})
