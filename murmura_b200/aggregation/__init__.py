"""Byzantine-robust aggregation rules (CPU oracles + device kernel plans)."""
from murmura_b200._lazy import lazy_exports

__getattr__, __dir__, __all__ = lazy_exports(__name__, {
    "base": ["Aggregator"],
    "fedavg": ["FedAvgAggregator"],
    "krum": ["KrumAggregator"],
    "balance": ["BALANCEAggregator"],
    "sketchguard": ["SketchguardAggregator"],
    "ubar": ["UBARAggregator"],
    "evidential_trust": ["EvidentialTrustAggregator"],
})
