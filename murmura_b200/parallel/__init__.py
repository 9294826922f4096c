"""B200 execution layer: flat peer-mapped arena, node placement, the fused round loop."""
from murmura_b200.parallel.arena import Placement, StateLayout, SymmetricArena

__all__ = ["Placement", "StateLayout", "SymmetricArena", "B200Network"]


def __getattr__(name):          # engine imports torch.distributed lazily; keep `import murmura_b200.parallel` light
    if name == "B200Network":
        from murmura_b200.parallel.engine import B200Network
        return B200Network
    raise AttributeError(name)
