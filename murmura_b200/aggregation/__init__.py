"""Byzantine-robust aggregation rules (CPU oracles + device kernel plans)."""
from murmura_b200.aggregation.base import Aggregator
from murmura_b200.aggregation.fedavg import FedAvgAggregator
from murmura_b200.aggregation.krum import KrumAggregator
from murmura_b200.aggregation.balance import BALANCEAggregator
from murmura_b200.aggregation.sketchguard import SketchguardAggregator
from murmura_b200.aggregation.ubar import UBARAggregator
from murmura_b200.aggregation.evidential_trust import EvidentialTrustAggregator

__all__ = ["Aggregator", "FedAvgAggregator", "KrumAggregator", "BALANCEAggregator",
           "SketchguardAggregator", "UBARAggregator", "EvidentialTrustAggregator"]
