"""One federated node as an OS process over ZeroMQ (CPU/any-device compatibility backend).

Parity: reference ``murmura/distributed/node_process.py:60-364`` — wall-clock rounds
(``t_start + k·round_duration_s``), per-node seed ``seed+node_id``, PULL bind + lazy PUSH per
neighbour, training overrun ⇒ skip exchange and report ``skipped``, the same serialized bytes
pushed to every neighbour, partial aggregation with whatever arrived by the deadline, static or
mobility-driven neighbour sets.  Fix: messages are round-tagged and stale ones dropped.
This backend exists for API parity (``murmura run`` with ``backend: distributed``, ``run-node``);
the performance path is ``backend: b200``.
"""
from __future__ import annotations

import time
from pathlib import Path
from typing import Dict, List, Optional, Set

import torch
import zmq
from murmura_b200.data.fast_loader import make_loaders

from murmura_b200.config.loader import load_config
from murmura_b200.config.schema import Config
from murmura_b200.core.node import Node
from murmura_b200.core.types import ModelState
from murmura_b200.distributed.endpoints import Endpoints
from murmura_b200.distributed.messaging import MsgType, decode_full, encode, pack_obj, pack_state, unpack_state
from murmura_b200.utils.factories import (build_aggregator_factory, build_attack, build_criterion,
                                          build_dataset_adapter, build_mobility_model, build_model_factory)
from murmura_b200.utils.seed import set_seed


class NodeProcess:
    log_tag = "Node"

    def __init__(self, node_id: int, config: Config, endpoints: Endpoints, t_start: float, mobility=None):
        self.node_id, self.config, self.endpoints, self.t_start = node_id, config, endpoints, t_start
        self.mobility = mobility
        self._ctx: Optional[zmq.Context] = None
        self._pull = None
        self._push_socks: Dict[int, zmq.Socket] = {}
        self._monitor_push = None
        self._static_neighbors: Optional[List[int]] = None

    @classmethod
    def from_config_path(cls, node_id: int, config_path: str, endpoints: Endpoints, t_start: float):
        config = load_config(Path(config_path))
        return cls(node_id=node_id, config=config, endpoints=endpoints, t_start=t_start,
                   mobility=build_mobility_model(config))

    # ---- lifecycle ------------------------------------------------------------------
    def run(self) -> None:
        set_seed(self.config.experiment.seed + self.node_id)
        device = self._resolve_device()
        node = self._build_node(device)
        attack = build_attack(self.config)
        self._prepare(node, device)
        self._ctx = zmq.Context()
        try:
            self._setup_sockets()
            self._run_all_rounds(node, attack)
        finally:
            self._teardown_sockets()

    def _prepare(self, node: Node, device: torch.device) -> None:
        """Hook for subclasses (DMTT) that need extra state before round 0."""

    def _setup_sockets(self) -> None:
        self._pull = self._ctx.socket(zmq.PULL)
        self._pull.bind(self.endpoints.node_pull_bind(self.node_id))
        for nid in self._get_static_neighbors():
            self._ensure_push_sock(nid)
        self._monitor_push = self._ctx.socket(zmq.PUSH)
        self._monitor_push.connect(self.endpoints.monitor_pull_connect())
        time.sleep(0.1)

    def _ensure_push_sock(self, neighbor_id: int):
        sock = self._push_socks.get(neighbor_id)
        if sock is None:
            sock = self._ctx.socket(zmq.PUSH)
            sock.connect(self.endpoints.node_pull_connect(neighbor_id))
            self._push_socks[neighbor_id] = sock
        return sock

    def _teardown_sockets(self) -> None:
        for sock in (self._pull, self._monitor_push, *self._push_socks.values()):
            if sock is not None:
                sock.close(linger=200)
        self._push_socks.clear()
        if self._ctx is not None:
            self._ctx.term()

    # ---- rounds ---------------------------------------------------------------------
    def _run_all_rounds(self, node: Node, attack) -> None:
        dur = self.config.distributed.round_duration_s
        for r in range(self.config.experiment.rounds):
            opens = self.t_start + r * dur
            wait = opens - time.monotonic()
            if wait > 0:
                time.sleep(wait)
            self._execute_round(node=node, attack=attack, round_idx=r, round_wall_end=opens + dur,
                                current_neighbors=self._get_current_neighbors(r))

    def _outgoing_state(self, node: Node, attack, round_idx: int) -> ModelState:
        state = node.get_state()
        if attack is not None and attack.is_compromised(self.node_id):
            state = attack.apply_attack(node_id=self.node_id, model_state=state, round_num=round_idx)
        return state

    def _train_or_skip(self, node: Node, attack, round_idx: int, round_wall_end: float) -> bool:
        """Local training for honest nodes; returns False when the round budget is blown."""
        cfg = self.config
        if not (attack is not None and attack.is_compromised(self.node_id)):
            node.local_train(epochs=cfg.training.local_epochs, lr=cfg.training.lr, round_num=round_idx)
        if time.monotonic() >= round_wall_end:
            print(f"[{self.log_tag} {self.node_id}] WARNING: training for round {round_idx + 1} exceeded "
                  f"round_duration_s={cfg.distributed.round_duration_s}s. Model exchange will be skipped.",
                  flush=True)
            self._push_metrics(node, round_idx, skipped=True)
            return False
        return True

    def _execute_round(self, node: Node, attack, round_idx: int, round_wall_end: float,
                       current_neighbors: List[int]) -> None:
        if not self._train_or_skip(node, attack, round_idx, round_wall_end):
            return
        blob = pack_state(self._outgoing_state(node, attack, round_idx))
        for nid in current_neighbors:
            self._ensure_push_sock(nid).send_multipart(encode(MsgType.MODEL_STATE, self.node_id, blob, round_idx))
        received = self._collect_neighbor_states(current_neighbors, round_idx, round_wall_end)
        if received:
            node.apply_aggregated_state(node.aggregate_with_neighbors(received, round_idx))
        self._push_metrics(node, round_idx)

    def _collect_neighbor_states(self, expected: List[int], round_idx: int, deadline: float) -> Dict[int, ModelState]:
        got: Dict[int, ModelState] = {}
        want: Set[int] = set(expected)
        while len(got) < len(want):
            left_ms = int((deadline - time.monotonic()) * 1000)
            if left_ms <= 0:
                print(f"[{self.log_tag} {self.node_id}] Round {round_idx + 1}: deadline reached, missing states "
                      f"from {sorted(want - set(got))}. Aggregating with {len(got)}/{len(want)} neighbours.", flush=True)
                break
            if not self._pull.poll(timeout=max(50, left_ms)):
                continue
            kind, sender, rnd, payload = decode_full(self._pull.recv_multipart())
            if kind == MsgType.MODEL_STATE and sender in want and rnd in (-1, round_idx):
                got[sender] = unpack_state(payload)
        return got

    def _push_metrics(self, node: Node, round_idx: int, skipped: bool = False) -> None:
        metrics = {"accuracy": 0.0, "loss": 0.0, "skipped": True} if skipped else dict(node.evaluate())
        metrics["round_idx"] = round_idx
        self._monitor_push.send_multipart(encode(MsgType.METRICS, self.node_id, pack_obj(metrics), round_idx))

    # ---- neighbours -----------------------------------------------------------------
    def _get_static_neighbors(self) -> List[int]:
        if self._static_neighbors is None:
            n = self.config.topology.num_nodes
            if self.mobility is not None:
                self._static_neighbors = [i for i in range(n) if i != self.node_id]
            else:
                from murmura_b200.topology import create_topology
                t = self.config.topology
                self._static_neighbors = create_topology(t.type, n, p=t.p, k=t.k, seed=t.seed).neighbors[self.node_id]
        return self._static_neighbors

    def _get_current_neighbors(self, round_idx: int) -> List[int]:
        if self.mobility is not None:
            return self.mobility.neighbors_at(round_idx).get(self.node_id, [])
        return self._get_static_neighbors()

    # ---- construction ---------------------------------------------------------------
    def _resolve_device(self) -> torch.device:
        from murmura_b200.utils.device import get_device
        dev = get_device()
        if dev.type == "cuda":      # pin node i → GPU i mod G (the reference piles everything on cuda:0)
            return torch.device("cuda", self.node_id % torch.cuda.device_count())
        return dev

    def _build_node(self, device: torch.device) -> Node:
        cfg = self.config
        adapter = build_dataset_adapter(cfg)
        model_factory = build_model_factory(cfg)
        aggregator_factory = build_aggregator_factory(cfg, model_factory, device)
        criterion, evidential = build_criterion(cfg)
        train_loader, test_loader, _ = make_loaders(adapter, self.node_id, cfg.training.batch_size)
        return Node(node_id=self.node_id, model=model_factory().to(device), train_loader=train_loader, test_loader=test_loader,
                    aggregator=aggregator_factory(self.node_id), device=device, criterion=criterion,
                    evidential=evidential)
