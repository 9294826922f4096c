"""Core runtime: Node, Network and type aliases."""
from murmura_b200.core.types import ModelState, DataPartition, ModelProtocol
from murmura_b200.core.node import Node
from murmura_b200.core.network import Network

__all__ = ["Network", "Node", "ModelState", "DataPartition", "ModelProtocol"]
