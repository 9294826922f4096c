"""Attack protocol (parity: reference ``murmura/attacks/base.py:8-52``)."""
from __future__ import annotations

import random
from typing import List, Protocol, Set, runtime_checkable

from murmura_b200.core.types import ModelState


@runtime_checkable
class Attack(Protocol):
    """Anything with these three methods can be handed to ``Network(attack=...)``."""

    def is_compromised(self, node_id: int) -> bool: ...

    def get_compromised_nodes(self) -> Set[int]: ...

    def apply_attack(self, node_id: int, model_state: ModelState, round_num: int, **kwargs) -> ModelState: ...


def select_compromised(num_nodes: int, fraction: float, seed: int, *, at_least_one: bool,
                       reseed_global: bool) -> Set[int]:
    """Pick the Byzantine node ids.

    ``reseed_global=True`` reproduces the Gaussian/Directed behaviour of reseeding the
    *global* ``random`` module (reference ``attacks/gaussian.py:37-44``); ``False`` uses a
    private generator like the topology liar (``attacks/topology_liar.py:41-45``).  Both
    yield the same ids for the same seed (SURVEY §9 golden values).
    """
    count = int(num_nodes * fraction)
    if at_least_one:
        count = max(1, count)
    elif count == 0 and fraction > 0:
        count = 1
    count = min(count, num_nodes)
    if reseed_global:
        random.seed(seed)
        picked: List[int] = random.sample(range(num_nodes), count)
    else:
        picked = random.Random(seed).sample(range(num_nodes), count)
    return set(picked)
