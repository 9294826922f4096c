"""Byzantine attack injectors."""
from murmura_b200._lazy import lazy_exports

__getattr__, __dir__, __all__ = lazy_exports(__name__, {
    "base": ["Attack"],
    "gaussian": ["GaussianAttack"],
    "directed": ["DirectedDeviationAttack"],
    "topology_liar": ["TopologyLiarAttack"],
})
