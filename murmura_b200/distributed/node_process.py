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

from murmura_b200.config.loader import load_config
from murmura_b200.config.schema import Config
from murmura_b200.core.node import Node
from murmura_b200.core.types import ModelState
from murmura_b200.data.fast_loader import make_loaders
from murmura_b200.distributed.endpoints import Endpoints
from murmura_b200.distributed.messaging import MsgType, decode_full, encode, pack_obj, pack_state, unpack_state
from murmura_b200.utils.factories import (build_aggregator_factory, build_attack, build_criterion,
                                          build_dataset_adapter, build_mobility_model, build_model_factory)
from murmura_b200.utils.seed import set_seed


class Mailbox:
    """The ZeroMQ side of one node: a bound PULL inbox, PUSH outboxes connected on first use, and the uplink to the monitor.
    Frames are the multipart messages of :mod:`murmura_b200.distributed.messaging`."""

    def __init__(self, node_id: int, endpoints: Endpoints):
        self.node_id, self.endpoints = node_id, endpoints
        self.ctx: Optional[zmq.Context] = None
        self.inbox = None
        self.uplink = None
        self.outboxes: Dict[int, zmq.Socket] = {}

    def open(self, peers) -> "Mailbox":
        self.ctx = zmq.Context()
        self.inbox = self.ctx.socket(zmq.PULL)
        self.inbox.bind(self.endpoints.node_pull_bind(self.node_id))
        for peer in peers:
            self._outbox(peer)
        self.uplink = self.ctx.socket(zmq.PUSH)
        self.uplink.connect(self.endpoints.monitor_pull_connect())
        time.sleep(0.1)                                   # let the connects settle before the first round opens
        return self

    def _outbox(self, peer: int):
        if peer not in self.outboxes:
            out = self.ctx.socket(zmq.PUSH)
            out.connect(self.endpoints.node_pull_connect(peer))
            self.outboxes[peer] = out
        return self.outboxes[peer]

    def post(self, peer: int, frames) -> None:
        self._outbox(peer).send_multipart(frames)

    def report(self, frames) -> None:
        self.uplink.send_multipart(frames)

    def receive(self, timeout_ms: int):
        """One decoded message ``(kind, sender, round, payload)`` or ``None`` after ``timeout_ms`` without traffic."""
        if not self.inbox.poll(timeout=timeout_ms):
            return None
        return decode_full(self.inbox.recv_multipart())

    def close(self) -> None:
        for sock in (self.inbox, self.uplink, *self.outboxes.values()):
            if sock is not None:
                sock.close(linger=200)
        self.outboxes.clear()
        if self.ctx is not None:
            self.ctx.term()
            self.ctx = None


class NodeProcess:
    log_tag = "Node"

    def __init__(self, node_id: int, config: Config, endpoints: Endpoints, t_start: float, mobility=None):
        self.node_id, self.config, self.endpoints, self.t_start = node_id, config, endpoints, t_start
        self.mobility = mobility
        self.mail = Mailbox(node_id, endpoints)
        self._fixed_peers: Optional[List[int]] = None

    @classmethod
    def from_config_path(cls, node_id: int, config_path: str, endpoints: Endpoints, t_start: float):
        config = load_config(Path(config_path))
        return cls(node_id=node_id, config=config, endpoints=endpoints, t_start=t_start,
                   mobility=build_mobility_model(config))

    # ---- lifecycle ------------------------------------------------------------------
    def run(self) -> None:
        set_seed(self.config.experiment.seed + self.node_id)
        device = self._pick_device()
        node = self._make_node(device)
        attack = build_attack(self.config)
        self._prepare(node, device)
        try:
            self.mail.open(self._peer_universe())
            self._round_loop(node, attack)
        finally:
            self.mail.close()

    def _prepare(self, node: Node, device: torch.device) -> None:
        """Hook for subclasses (DMTT) that need extra state before round 0."""

    # ---- rounds ---------------------------------------------------------------------
    def _round_loop(self, node: Node, attack) -> None:
        dur = self.config.distributed.round_duration_s
        for r in range(self.config.experiment.rounds):
            opens = self.t_start + r * dur
            wait = opens - time.monotonic()
            if wait > 0:
                time.sleep(wait)
            self._execute_round(node=node, attack=attack, round_idx=r, round_wall_end=opens + dur,
                                current_neighbors=self._neighbors_in_round(r))

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
            self._report(node, round_idx, skipped=True)
            return False
        return True

    def _execute_round(self, node: Node, attack, round_idx: int, round_wall_end: float,
                       current_neighbors: List[int]) -> None:
        if not self._train_or_skip(node, attack, round_idx, round_wall_end):
            return
        frames = encode(MsgType.MODEL_STATE, self.node_id, pack_state(self._outgoing_state(node, attack, round_idx)), round_idx)
        for peer in current_neighbors:                      # one serialisation, the same bytes to every neighbour
            self.mail.post(peer, frames)
        received = self._await_states(current_neighbors, round_idx, round_wall_end)
        if received:
            node.apply_aggregated_state(node.aggregate_with_neighbors(received, round_idx))
        self._report(node, round_idx)

    def _await_states(self, expected: List[int], round_idx: int, deadline: float) -> Dict[int, ModelState]:
        got: Dict[int, ModelState] = {}
        want: Set[int] = set(expected)
        while len(got) < len(want):
            left_ms = int((deadline - time.monotonic()) * 1000)
            if left_ms <= 0:
                print(f"[{self.log_tag} {self.node_id}] Round {round_idx + 1}: deadline reached, missing states "
                      f"from {sorted(want - set(got))}. Aggregating with {len(got)}/{len(want)} neighbours.", flush=True)
                break
            msg = self.mail.receive(max(50, left_ms))
            if msg is None:
                continue
            kind, sender, rnd, payload = msg
            if kind == MsgType.MODEL_STATE and sender in want and rnd in (-1, round_idx):
                got[sender] = unpack_state(payload)
        return got

    def _report(self, node: Node, round_idx: int, skipped: bool = False) -> None:
        metrics = {"accuracy": 0.0, "loss": 0.0, "skipped": True} if skipped else dict(node.evaluate())
        metrics["round_idx"] = round_idx
        self.mail.report(encode(MsgType.METRICS, self.node_id, pack_obj(metrics), round_idx))

    # ---- neighbours -----------------------------------------------------------------
    def _peer_universe(self) -> List[int]:
        """Every node this one may ever talk to (outboxes are connected up front): the static neighbours, or everybody
        under mobility."""
        if self._fixed_peers is None:
            n = self.config.topology.num_nodes
            if self.mobility is not None:
                self._fixed_peers = [i for i in range(n) if i != self.node_id]
            else:
                from murmura_b200.topology import create_topology
                t = self.config.topology
                self._fixed_peers = create_topology(t.type, n, p=t.p, k=t.k, seed=t.seed).neighbors[self.node_id]
        return self._fixed_peers

    def _neighbors_in_round(self, round_idx: int) -> List[int]:
        if self.mobility is not None:
            return self.mobility.neighbors_at(round_idx).get(self.node_id, [])
        return self._peer_universe()

    # ---- construction ---------------------------------------------------------------
    def _pick_device(self) -> torch.device:
        from murmura_b200.utils.device import get_device
        dev = get_device()
        if dev.type == "cuda":      # pin node i → GPU i mod G (the reference piles everything on cuda:0)
            return torch.device("cuda", self.node_id % torch.cuda.device_count())
        return dev

    def _make_node(self, device: torch.device) -> Node:
        cfg = self.config
        adapter = build_dataset_adapter(cfg)
        model_factory = build_model_factory(cfg)
        criterion, evidential = build_criterion(cfg)
        train_loader, test_loader, _ = make_loaders(adapter, self.node_id, cfg.training.batch_size)
        return Node(node_id=self.node_id, model=model_factory().to(device), train_loader=train_loader, test_loader=test_loader,
                    aggregator=build_aggregator_factory(cfg, model_factory, device)(self.node_id), device=device,
                    criterion=criterion, evidential=evidential)
