"""DMTT — dynamic topology with trusted collaborator selection."""
from murmura_b200._lazy import lazy_exports

__getattr__, __dir__, __all__ = lazy_exports(__name__, {
    "state": ["DMTTNodeState"],
    "node_process": ["DMTTNodeProcess"],
})
