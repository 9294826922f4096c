"""murmura_b200 — Blackwell-native decentralized federated learning.

Same public surface as the reference package (``murmura/__init__.py:10-33``) plus the B200
engine (``backend: b200``).
"""
__version__ = "0.1.0"

from murmura_b200.config import Config
from murmura_b200.core.network import Network
from murmura_b200.core.node import Node
from murmura_b200.topology import create_topology, Topology, MobilityModel
from murmura_b200.aggregation import (FedAvgAggregator, KrumAggregator, BALANCEAggregator,
                                      SketchguardAggregator, UBARAggregator, EvidentialTrustAggregator)

__all__ = ["Config", "Network", "Node", "create_topology", "Topology", "MobilityModel", "FedAvgAggregator",
           "KrumAggregator", "BALANCEAggregator", "SketchguardAggregator", "UBARAggregator",
           "EvidentialTrustAggregator"]
