"""Byzantine attack injectors."""
from murmura_b200.attacks.base import Attack
from murmura_b200.attacks.gaussian import GaussianAttack
from murmura_b200.attacks.directed import DirectedDeviationAttack
from murmura_b200.attacks.topology_liar import TopologyLiarAttack

__all__ = ["Attack", "GaussianAttack", "DirectedDeviationAttack", "TopologyLiarAttack"]
