"""Passive metrics monitor of the ZeroMQ backend.

Parity: reference ``murmura/distributed/monitor.py:28-175`` — one PULL socket, rounds flushed in
order once all N nodes reported, partial rounds flushed at the hard deadline
``t_start + (rounds+2)·round_duration``, same history schema and ``[Monitor] …`` lines.
"""
from __future__ import annotations

import time
from typing import Any, Dict, List, Set

import numpy as np
import zmq

from murmura_b200.core.network import new_history
from murmura_b200.distributed.endpoints import Endpoints
from murmura_b200.distributed.messaging import MsgType, decode, unpack_obj


class Monitor:
    def __init__(self, num_nodes: int, endpoints: Endpoints, rounds: int, t_start: float,
                 round_duration_s: float, compromised_nodes: Set[int], verbose: bool = False):
        self.num_nodes, self.endpoints, self.rounds = num_nodes, endpoints, rounds
        self.t_start, self.round_duration_s = t_start, round_duration_s
        self.compromised_nodes, self.verbose = compromised_nodes, verbose
        self.history: Dict[str, List[Any]] = new_history()

    def run(self) -> Dict[str, List[Any]]:
        ctx = zmq.Context()
        pull = ctx.socket(zmq.PULL)
        try:
            pull.bind(self.endpoints.monitor_pull_bind())
            self._collect(pull)
        finally:
            pull.close(linger=0)
            ctx.term()
        return self.history

    def _collect(self, pull) -> None:
        pending: Dict[int, Dict[int, Dict[str, Any]]] = {}
        expected, received = self.num_nodes * self.rounds, 0
        deadline = self.t_start + (self.rounds + 2) * self.round_duration_s
        while received < expected:
            left_ms = int((deadline - time.monotonic()) * 1000)
            if left_ms <= 0:
                self._log("Deadline reached; some metrics may be missing.")
                break
            if not pull.poll(timeout=max(200, min(left_ms, 2000))):
                continue
            kind, node_id, payload = decode(pull.recv_multipart())
            if kind != MsgType.METRICS:
                continue
            m = unpack_obj(payload)
            pending.setdefault(m.pop("round_idx"), {})[node_id] = m
            received += 1
            while True:
                nxt = len(self.history["round"])
                if len(pending.get(nxt, ())) < self.num_nodes:
                    break
                self._record(nxt + 1, pending.pop(nxt))
        for r in sorted(pending):
            if pending[r]:
                self._record(r + 1, pending[r])

    def _record(self, round_num: int, metrics: Dict[int, Dict[str, Any]]) -> None:
        acc = [m["accuracy"] for m in metrics.values()]
        honest = [m["accuracy"] for n, m in metrics.items() if n not in self.compromised_nodes]
        comp = [m["accuracy"] for n, m in metrics.items() if n in self.compromised_nodes]
        h = self.history
        h["round"].append(round_num)
        h["mean_accuracy"].append(float(np.mean(acc)))
        h["std_accuracy"].append(float(np.std(acc)))
        h["mean_loss"].append(float(np.mean([m["loss"] for m in metrics.values()])))
        if honest:
            h["honest_accuracy"].append(float(np.mean(honest)))
        if comp:
            h["compromised_accuracy"].append(float(np.mean(comp)))
        vac = [m["vacuity"] for m in metrics.values() if "vacuity" in m]
        if vac:
            h["mean_vacuity"].append(float(np.mean(vac)))
            h["mean_entropy"].append(float(np.mean([m["entropy"] for m in metrics.values() if "entropy" in m])))
            h["mean_strength"].append(float(np.mean([m["strength"] for m in metrics.values() if "strength" in m])))
        self._log(f"Round {round_num} ({len(metrics)}/{self.num_nodes} nodes): acc={np.mean(acc):.4f} ± {np.std(acc):.4f}")
        if honest and comp:
            self._log(f"  Honest: {np.mean(honest):.4f}  Compromised: {np.mean(comp):.4f}")

    def _log(self, msg: str) -> None:
        if self.verbose:
            print(f"[Monitor] {msg}", flush=True)
