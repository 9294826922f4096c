"""ZeroMQ wall-clock backend (API-parity path; ``backend: b200`` is the performance path)."""
from murmura_b200._lazy import lazy_exports

__getattr__, __dir__, __all__ = lazy_exports(__name__, {
    "runner": ["DistributedRunner"],
})
