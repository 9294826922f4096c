"""DMTT node process: dynamic topology + trusted collaborator selection over ZeroMQ.

Parity: reference ``murmura/dmtt/node_process.py:53-406`` (round order: train → state (+model
attack) → truthful or falsified TOPO_CLAIM → send both to ``C_i^{t-1}`` → collect → link EMA →
foreign-model scoring on the local test loader → claim verification against the deterministic
mobility model → aggregate → Top-B over ``G^t`` neighbours → metrics).
"""
from __future__ import annotations

import time
from pathlib import Path
from typing import Any, Dict, List, Optional, Set, Tuple

import torch
import torch.nn as nn

from murmura_b200.config.loader import load_config
from murmura_b200.config.schema import Config, DMTTConfig
from murmura_b200.distributed.endpoints import Endpoints
from murmura_b200.distributed.messaging import (MsgType, encode, pack_obj, pack_state, unpack_obj,
                                                unpack_state)
from murmura_b200.distributed.node_process import NodeProcess
from murmura_b200.dmtt.state import DMTTNodeState
from murmura_b200.topology.dynamic import MobilityModel


def verify_claim(claimed: List[int], true_neighbors: Set[int]) -> Tuple[float, float]:
    """(confirmations d, contradictions x) of one neighbourhood claim."""
    d = float(sum(1 for u in claimed if u in true_neighbors))
    return d, float(len(claimed)) - d


class DMTTNodeProcess(NodeProcess):
    log_tag = "DMTT Node"

    def __init__(self, node_id: int, config: Config, endpoints: Endpoints, t_start: float, mobility: MobilityModel):
        super().__init__(node_id, config, endpoints, t_start, mobility=mobility)
        self.dmtt_cfg: DMTTConfig = config.dmtt  # type: ignore[assignment]
        self._dmtt: Optional[DMTTNodeState] = None
        self._collaborators: Optional[List[int]] = None

    @classmethod
    def from_config_path(cls, node_id: int, config_path: str, endpoints: Endpoints, t_start: float):
        config = load_config(Path(config_path))
        if config.mobility is None:
            raise ValueError("DMTTNodeProcess requires config.mobility to be set.")
        if config.dmtt is None:
            raise ValueError("DMTTNodeProcess requires config.dmtt to be set.")
        from murmura_b200.utils.factories import build_mobility_model
        return cls(node_id=node_id, config=config, endpoints=endpoints, t_start=t_start,
                   mobility=build_mobility_model(config))

    def _prepare(self, node, device) -> None:
        from murmura_b200.utils.factories import build_model_factory
        self._model_factory = build_model_factory(self.config)
        self._eval_device = device
        self._dmtt = DMTTNodeState(self.node_id, self.dmtt_cfg, self.config.topology.num_nodes)

    def _peer_universe(self) -> List[int]:
        return [i for i in range(self.config.topology.num_nodes) if i != self.node_id]

    def _neighbors_in_round(self, round_idx: int) -> List[int]:
        if round_idx == 0 or self._collaborators is None:
            return self.mobility.neighbors_at(0).get(self.node_id, [])
        return self._collaborators

    # ---- one round --------------------------------------------------------------------
    def _execute_round(self, node, attack, round_idx: int, round_wall_end: float, current_neighbors: List[int]) -> None:
        byz = attack is not None and attack.is_compromised(self.node_id)
        if not self._train_or_skip(node, attack, round_idx, round_wall_end):
            if self._collaborators is None:
                self._collaborators = self.mobility.neighbors_at(round_idx).get(self.node_id, [])
            return
        state_blob = pack_state(self._outgoing_state(node, attack, round_idx))
        truth = self.mobility.neighbors_at(round_idx).get(self.node_id, [])
        claimed = (attack.get_false_claims(node_id=self.node_id, true_neighbors=truth, round_num=round_idx)
                   if byz and hasattr(attack, "get_false_claims") else list(truth))
        claim_blob = pack_obj({"round_idx": round_idx, "neighbors": claimed})
        state_frames = encode(MsgType.MODEL_STATE, self.node_id, state_blob, round_idx)
        claim_frames = encode(MsgType.TOPO_CLAIM, self.node_id, claim_blob, round_idx)
        for nid in current_neighbors:
            self.mail.post(nid, state_frames)
            self.mail.post(nid, claim_frames)

        states, claims = self._collect_dmtt_messages(current_neighbors, round_idx, round_wall_end)
        assert self._dmtt is not None
        for nid in current_neighbors:
            self._dmtt.update_link_reliability(nid, nid in states)
        scores = self._score_neighbor_models(states, node, round_idx) if states else {}
        self._process_topo_claims(claims, round_idx)
        if states:
            node.apply_aggregated_state(node.aggregate_with_neighbors(states, round_idx))
        self._collaborators = self._dmtt.top_b(candidates=self.mobility.neighbors_at(round_idx).get(self.node_id, []),
                                               model_scores=scores, B=self.dmtt_cfg.budget_B)
        if self.config.experiment.verbose:
            print(f"[DMTT Node {self.node_id}] Round {round_idx + 1}: collaborators → {self._collaborators}", flush=True)
        self._report(node, round_idx)

    def _collect_dmtt_messages(self, expected: List[int], round_idx: int,
                               deadline: float) -> Tuple[Dict[int, Any], Dict[int, dict]]:
        states: Dict[int, Any] = {}
        claims: Dict[int, dict] = {}
        want = set(expected)
        while not (want <= set(states) and want <= set(claims)):
            left_ms = int((deadline - time.monotonic()) * 1000)
            if left_ms <= 0:
                print(f"[DMTT Node {self.node_id}] Round {round_idx + 1}: deadline — missing states from "
                      f"{sorted(want - set(states))}, claims from {sorted(want - set(claims))}.", flush=True)
                break
            msg = self.mail.receive(max(50, left_ms))
            if msg is None:
                continue
            kind, sender, rnd, payload = msg
            if sender not in want or rnd not in (-1, round_idx):
                continue
            if kind == MsgType.MODEL_STATE and sender not in states:
                states[sender] = unpack_state(payload)
            elif kind == MsgType.TOPO_CLAIM and sender not in claims:
                claims[sender] = unpack_obj(payload)
        return states, claims

    # ---- scoring / trust ----------------------------------------------------------------
    def _score_neighbor_models(self, neighbor_states: Dict[int, Any], node, round_idx: int) -> Dict[int, float]:
        if node.test_loader is None:
            return {j: 0.5 for j in neighbor_states}
        out: Dict[int, float] = {}
        for j, state in neighbor_states.items():
            try:
                probe = self._model_factory().to(self._eval_device)
                probe.load_state_dict({k: v.to(self._eval_device) for k, v in state.items()})
                acc, u_bar = self._evaluate_foreign(probe, node)
                out[j] = self._dmtt.model_score(acc, u_bar)  # type: ignore[union-attr]
            except Exception:
                out[j] = 0.0
        return out

    def _evaluate_foreign(self, model: nn.Module, node) -> Tuple[float, float]:
        model.eval()
        correct = torch.zeros((), device=self._eval_device)
        vac = torch.zeros((), device=self._eval_device)
        total = 0
        with torch.no_grad():
            for xb, yb in node.test_loader:
                xb, yb = xb.to(self._eval_device), yb.to(self._eval_device)
                out = model(xb)
                if node.evidential:
                    vac += (out.shape[-1] / out.sum(dim=-1)).sum()
                correct += (out.argmax(dim=-1) == yb).sum()
                total += yb.size(0)
        if total == 0:
            return 0.0, 0.0
        return float(correct) / total, float(vac) / total

    def _process_topo_claims(self, topo_claims: Dict[int, dict], round_idx: int) -> None:
        assert self._dmtt is not None
        truth = self.mobility.neighbors_at(round_idx)
        for j, claim in topo_claims.items():
            d, x = verify_claim(claim.get("neighbors", []), set(truth.get(j, [])))
            self._dmtt.update_trust(j, d=d, x=x)
