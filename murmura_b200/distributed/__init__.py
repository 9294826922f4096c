"""ZeroMQ wall-clock backend (API-parity path; ``backend: b200`` is the performance path)."""
from murmura_b200.distributed.runner import DistributedRunner

__all__ = ["DistributedRunner"]
