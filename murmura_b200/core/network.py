"""In-process orchestrator: the ``simulation`` backend and the public ``Network`` API.

Parity: reference ``murmura/core/network.py:16-312`` — train → Jacobi aggregate (all nodes
read the pre-round snapshot, write afterwards) → evaluate; Byzantine nodes skip training
and broadcast the attacked state; history keys and the scraped stdout lines are identical.
Deviation (documented): ``from_config`` builds the attack through
``build_attack`` so ``topology_liar`` with a wrapped model attack is honoured here too.

``Network.from_config`` with ``config.backend == "b200"`` returns the Blackwell engine
(:class:`murmura_b200.parallel.engine.B200Network`), which has the same ``train`` /
``history`` / ``get_node_statistics`` surface.
"""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional

import numpy as np
import torch
import torch.nn as nn
from torch.utils.data import DataLoader  # noqa: F401  (re-exported for user code that builds Nodes by hand)

from murmura_b200.data.fast_loader import make_loaders

from murmura_b200.attacks.base import Attack
from murmura_b200.core.node import Node
from murmura_b200.topology.base import Topology

HISTORY_KEYS = ("round", "mean_accuracy", "std_accuracy", "mean_loss", "honest_accuracy",
                "compromised_accuracy", "mean_vacuity", "mean_entropy", "mean_strength")


def new_history() -> Dict[str, List[Any]]:
    return {k: [] for k in HISTORY_KEYS}


def record_round(history: Dict[str, List[Any]], round_num: int, per_node: List[Dict[str, Any]],
                 compromised: Optional[set], verbose: bool) -> None:
    """Fold per-node eval dicts into ``history`` and print the stdout contract.

    Shared by the simulation backend, the ZMQ monitor and the B200 engine so all three
    produce byte-identical lines.
    """
    acc = [m.get("accuracy", 0.0) for m in per_node]
    loss = [m.get("loss", 0.0) for m in per_node]
    evid = [m for m in per_node if "vacuity" in m]
    honest, comp = [], []
    if compromised is not None:
        for nid, m in enumerate(per_node):
            (comp if m.get("node_id", nid) in compromised else honest).append(m.get("accuracy", 0.0))
    history["round"].append(round_num)
    history["mean_accuracy"].append(np.mean(acc))
    history["std_accuracy"].append(np.std(acc))
    history["mean_loss"].append(np.mean(loss))
    if honest:
        history["honest_accuracy"].append(np.mean(honest))
    if comp:
        history["compromised_accuracy"].append(np.mean(comp))
    if evid:
        history["mean_vacuity"].append(np.mean([m["vacuity"] for m in evid]))
        history["mean_entropy"].append(np.mean([m["entropy"] for m in evid]))
        history["mean_strength"].append(np.mean([m["strength"] for m in evid]))
    if verbose:
        print(f"Round {round_num}: Mean Accuracy = {np.mean(acc):.4f} ± {np.std(acc):.4f}")
        if honest and comp:
            print(f"  Honest: {np.mean(honest):.4f}, Compromised: {np.mean(comp):.4f}")
        if evid:
            print(f"  Uncertainty: Vacuity={history['mean_vacuity'][-1]:.4f}, "
                  f"Entropy={history['mean_entropy'][-1]:.4f}, Strength={history['mean_strength'][-1]:.2f}")


class Network:
    def __init__(self, nodes: List[Node], topology: Topology, attack: Optional[Attack] = None):
        if len(nodes) != topology.num_nodes:
            raise ValueError(f"Number of nodes ({len(nodes)}) must match topology ({topology.num_nodes})")
        self.nodes = nodes
        self.topology = topology
        self.attack = attack
        self.history = new_history()

    # ---- round loop -----------------------------------------------------------------
    def train(self, rounds: int, local_epochs: int = 1, lr: float = 0.01, verbose: bool = False,
              eval_every: int = 1) -> Dict[str, List[Any]]:
        for r in range(rounds):
            if verbose:
                print(f"\n=== Round {r + 1}/{rounds} ===")
            self._local_training_step(local_epochs, lr, r, verbose)
            self._aggregation_step(r, verbose)
            if (r + 1) % eval_every == 0:
                self._evaluation_step(r + 1, verbose)
        return self.history

    def _is_byzantine(self, node_id: int) -> bool:
        return bool(self.attack) and self.attack.is_compromised(node_id)

    def _local_training_step(self, epochs: int, lr: float, round_num: int, verbose: bool) -> None:
        for node in self.nodes:
            if not self._is_byzantine(node.node_id):
                node.local_train(epochs=epochs, lr=lr, round_num=round_num)

    def _aggregation_step(self, round_num: int, verbose: bool) -> None:
        snapshot = [node.get_state() for node in self.nodes]
        for nid in range(len(self.nodes)):
            if self._is_byzantine(nid):
                snapshot[nid] = self.attack.apply_attack(node_id=nid, model_state=snapshot[nid],
                                                         round_num=round_num)
        merged = [
            node.aggregate_with_neighbors(
                neighbor_states={j: snapshot[j] for j in self.topology.neighbors[nid]},
                round_num=round_num)
            for nid, node in enumerate(self.nodes)
        ]
        for node, state in zip(self.nodes, merged):
            node.apply_aggregated_state(state)

    def _evaluation_step(self, round_num: int, verbose: bool) -> None:
        per_node = []
        for node in self.nodes:
            m = dict(node.evaluate())
            m["node_id"] = node.node_id
            per_node.append(m)
        compromised = set(self.attack.get_compromised_nodes()) if self.attack else None
        record_round(self.history, round_num, per_node, compromised, verbose)

    def get_node_statistics(self) -> Dict[int, Dict[str, Any]]:
        return {node.node_id: node.get_aggregator_statistics() for node in self.nodes}

    # ---- construction ---------------------------------------------------------------
    @classmethod
    def from_config(cls, config: Any, model_factory: Callable[[], nn.Module], dataset_adapter: Any,
                    aggregator_factory: Callable[[int], Any], device: Optional[torch.device] = None,
                    criterion: Optional[nn.Module] = None, evidential: bool = False):
        if getattr(config, "backend", "simulation") == "b200":
            from murmura_b200.parallel.engine import B200Network
            return B200Network.from_config(config, model_factory, dataset_adapter, aggregator_factory,
                                           device=device, criterion=criterion, evidential=evidential)
        from murmura_b200.topology import create_topology
        from murmura_b200.utils.factories import build_attack

        topology = create_topology(config.topology.type, config.topology.num_nodes,
                                   p=config.topology.p, k=config.topology.k, seed=config.topology.seed)
        attack = build_attack(config)
        nodes: List[Node] = []
        for nid in range(config.topology.num_nodes):
            train_loader, test_loader, _ = make_loaders(dataset_adapter, nid, config.training.batch_size)
            nodes.append(Node(
                node_id=nid, model=model_factory(), train_loader=train_loader,
                test_loader=test_loader,                                       # evaluates on the training shard
                aggregator=aggregator_factory(nid), device=device, criterion=criterion, evidential=evidential))
        return cls(nodes=nodes, topology=topology, attack=attack)
