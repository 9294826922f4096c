"""DMTT — dynamic topology with trusted collaborator selection."""
from murmura_b200.dmtt.state import DMTTNodeState
from murmura_b200.dmtt.node_process import DMTTNodeProcess

__all__ = ["DMTTNodeState", "DMTTNodeProcess"]
