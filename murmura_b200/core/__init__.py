"""Core runtime: Node, Network and type aliases."""
from murmura_b200._lazy import lazy_exports

__getattr__, __dir__, __all__ = lazy_exports(__name__, {
    "types": ["ModelState", "DataPartition", "ModelProtocol"],
    "node": ["Node"],
    "network": ["Network"],
})
