"""Network topologies (static generators + mobility-driven G^t)."""
from murmura_b200.topology.base import Topology
from murmura_b200.topology.generators import create_topology
from murmura_b200.topology.dynamic import MobilityModel

__all__ = ["Topology", "create_topology", "MobilityModel"]
