"""Network topologies (static generators + mobility-driven G^t)."""
from murmura_b200._lazy import lazy_exports

__getattr__, __dir__, __all__ = lazy_exports(__name__, {
    "base": ["Topology"],
    "generators": ["create_topology"],
    "dynamic": ["MobilityModel"],
})
