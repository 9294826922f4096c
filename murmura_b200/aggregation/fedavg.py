"""FedAvg: uniform mean over {own} ∪ neighbours (reference ``aggregation/fedavg.py:8-42``).

B200 path: one ``ops.weighted_gather`` launch with weights ``1/m``; for a full
mesh the NVLS ``multimem.ld_reduce`` specialisation applies.
"""
from __future__ import annotations

from typing import Dict

from murmura_b200.aggregation.base import Aggregator, average_states
from murmura_b200.core.types import ModelState


class FedAvgAggregator(Aggregator):
    kernel_family = "fedavg"

    def __init__(self, **kwargs):
        super().__init__(**kwargs)

    def aggregate(self, node_id: int, own_state: ModelState, neighbor_states: Dict[int, ModelState],
                  round_num: int, **kwargs) -> ModelState:
        return average_states([own_state, *neighbor_states.values()])
