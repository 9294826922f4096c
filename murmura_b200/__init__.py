"""murmura_b200 — Blackwell-native decentralized federated learning.

Same public surface as the reference package (``murmura/__init__.py:10-33``) plus the B200 engine (``backend: b200``); names are
resolved lazily through the export table below.
"""
from murmura_b200._lazy import lazy_exports

__version__ = "0.1.0"

__getattr__, __dir__, __all__ = lazy_exports(__name__, {
    "config": ["Config"],
    "core.network": ["Network"],
    "core.node": ["Node"],
    "topology": ["create_topology", "Topology", "MobilityModel"],
    "aggregation": ["FedAvgAggregator", "KrumAggregator", "BALANCEAggregator", "SketchguardAggregator", "UBARAggregator", "EvidentialTrustAggregator"],
})
