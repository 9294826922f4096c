"""B200 engine: one SPMD process per GPU, virtual nodes in a flat peer-mapped arena, fused kernels.

Public surface mirrors the reference orchestrator (``murmura/core/network.py:16-312``):
``B200Network.from_config(...)``, ``train(rounds, local_epochs, lr, verbose, eval_every) -> history``,
``get_node_statistics()``; the stdout contract and history schema are shared with the simulation
backend through :func:`murmura_b200.core.network.record_round`.

One round (per rank; all device work is enqueued without host synchronisation):

1. **local training** of every honest virtual node hosted here — parameters are views into the arena's
   ``live`` plane, each node's step (gather batch → forward → loss → backward → fused flat SGD kernel)
   is a CUDA graph replayed ``epochs × batches`` times on one of ``b200.streams`` streams;
2. **publish** — ``live → published[parity]`` with the Byzantine attack fused in (Philox Gaussian noise /
   directed-deviation scale), then a release-store of the round epoch into every peer's control page;
3. **filter** — the aggregator's decision kernels (distances, Gram on tcgen05, Count-Sketch, trust …)
   wait on the epoch flags *inside the kernel* and read neighbour tiles straight from peer memory;
4. **weighted_gather** — streams the accepted neighbours' tiles over NVLink and writes ``live`` in place;
5. **evaluation** (every ``eval_every`` rounds) with device-side accumulators, one small D2H per eval.

NCCL is only used for bootstrap (IPC-handle exchange) and for tiny control tensors (metrics, the
128×128 Gram partials, DMTT collaborator rows).
"""
from __future__ import annotations

import contextlib
import io
import math
import os
from contextlib import nullcontext
from dataclasses import dataclass, field
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from murmura_b200 import ops as _ops
from murmura_b200.aggregation.balance import decayed_factor
from murmura_b200.core.network import new_history, record_round
from murmura_b200.parallel.arena import Placement, StateLayout, SymmetricArena
from murmura_b200.topology.base import Topology

_STAT_COLS = 8


def _dist():
    import torch.distributed as dist
    return dist


def init_distributed() -> Tuple[int, int, int]:
    """(rank, world, local_rank) — initialises the NCCL group when launched under torchrun."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", str(rank)))
    if world > 1:
        dist = _dist()
        if torch.cuda.is_available():
            torch.cuda.set_device(local_rank)
        if not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29511")
            dist.init_process_group(backend="nccl" if torch.cuda.is_available() else "gloo",
                                    device_id=torch.device("cuda", local_rank) if torch.cuda.is_available() else None)
    return rank, world, local_rank


@dataclass
class VirtualNode:
    gid: int
    slot: int
    model: nn.Module
    X: torch.Tensor
    y: torch.Tensor
    n: int
    eb: int                 # effective batch size  min(bs, max(2, n))
    nb: int                 # batches per epoch (drop_last when n > eb)
    byzantine: bool = False
    nhwc: bool = False
    params: List[nn.Parameter] = field(default_factory=list)
    perm_buf: Optional[torch.Tensor] = None
    step: Optional[torch.Tensor] = None
    ticket: Optional[torch.Tensor] = None    # scratch word of gather_batch's last-CTA step increment
    launches_per_step: int = 3               # our kernels per SGD step (measured when the step is traced / run eagerly)
    autotuned: bool = False                  # first (serial, un-split) step done → cuDNN algorithm cache is warm
    arange: Optional[torch.Tensor] = None
    loss_sum: Optional[torch.Tensor] = None
    train_graph: Any = None
    steps_per_replay: int = 1
    eval_graph: Any = None
    graph_key: Any = None
    split_bwd: Any = None   # SplitBackward of this node (side stream for weight gradients)


def mlp_plan(model: nn.Module, layout: StateLayout) -> Optional[List[Dict[str, Any]]]:
    """Layer list of the MLP families for the grouped tcgen05 forward; ``None`` for anything else (conv nets …)."""
    from murmura_b200.models.mlp import MLP, EvidentialMLP
    ent = {e.name: e for e in layout.entries}
    plan: List[Dict[str, Any]] = []

    def linear(prefix: str) -> Dict[str, Any]:
        w = ent[prefix + ".weight"]
        return {"w": w.offset, "b": ent[prefix + ".bias"].offset if prefix + ".bias" in ent else None, "bn": None, "act": 0,
                "N": w.shape[0], "K": w.shape[1]}

    if isinstance(model, EvidentialMLP):
        mods = list(model.feature_extractor.named_children())
        i = 0
        while i < len(mods):
            name, m = mods[i]
            if not isinstance(m, nn.Linear):
                return None
            layer = linear(f"feature_extractor.{name}")
            j = i + 1
            while j < len(mods) and not isinstance(mods[j][1], nn.Linear):
                n2, m2 = mods[j]
                if isinstance(m2, nn.BatchNorm1d):
                    pre = f"feature_extractor.{n2}"
                    layer["bn"] = (ent[pre + ".running_mean"].offset, ent[pre + ".running_var"].offset,
                                   ent[pre + ".weight"].offset if m2.affine else None, ent[pre + ".bias"].offset if m2.affine else None)
                    layer["eps"] = m2.eps
                elif isinstance(m2, nn.ReLU):
                    layer["act"] = 1
                elif not isinstance(m2, (nn.Dropout, nn.Identity)):
                    return None
                j += 1
            plan.append(layer)
            i = j
        head = linear("evidential_head.fc")
        head["act"] = 2
        plan.append(head)
        return plan
    if isinstance(model, MLP):
        mods = list(model.net.named_children())
        for idx, (name, m) in enumerate(mods):
            if isinstance(m, nn.Linear):
                layer = linear(f"net.{name}")
                if idx + 1 < len(mods) and isinstance(mods[idx + 1][1], nn.ReLU):
                    layer["act"] = 1
                plan.append(layer)
            elif not isinstance(m, nn.ReLU):
                return None
        return plan
    return None


class ForeignEval:
    """Score a *foreign* weight row on one node's data (UBAR stage 2, EvidentialTrust, DMTT model scores).

    The reference deep-copies / ``load_state_dict``s a model per neighbour (``aggregation/ubar.py:170-190``,
    ``aggregation/evidential_trust.py:236-281``, ``dmtt/node_process.py:309-363``).  Here the candidate row is pulled with ONE
    device-to-device copy (over NVLink when it lives on a peer) into a per-node staging row whose views are the functional
    state of the node's module, and the whole forward + metric kernel is ONE CUDA-graph replay: 3 launches per candidate.
    """

    def __init__(self, eng: "B200Network", vn: "VirtualNode", rows: int, kind: str):
        L = eng.layout
        self.eng, self.vn, self.kind, self.rows = eng, vn, kind, rows
        self.stage = torch.zeros(L.stride, device=eng.device)
        self.state = dict(L.row_views(self.stage, None))
        for e in L.int_entries():
            self.state[e.name] = eng.ints[vn.slot][e.offset:e.offset + e.numel].view(e.shape)
        self.x = torch.zeros(rows, *vn.X.shape[1:], device=eng.device)
        self.y = torch.zeros(rows, dtype=torch.long, device=eng.device)
        self.stats = torch.zeros(_STAT_COLS, device=eng.device)
        self.graph = None

    def load_inputs(self, x: torch.Tensor, y: torch.Tensor) -> None:
        self.x.copy_(x); self.y.copy_(y)

    def _body(self) -> None:
        eng, vn = self.eng, self.vn
        self.stats.zero_()
        out = eng._forward_with(vn, self.state, eng._inputs(vn, self.x)).contiguous()
        if self.kind == "dirichlet":
            eng.ext.dirichlet_eval(out, self.y, None, self.stats)
        else:
            eng.ext.ce_eval(out, self.y, None, self.stats)

    def run(self, src_row: torch.Tensor, out_stats: torch.Tensor) -> None:
        n = self.eng.layout.Pf_pad
        self.stage[:n].copy_(src_row[:n], non_blocking=True)
        if self.eng.opt.cuda_graphs:
            if self.graph is None:
                side = self.eng.capture_streams[self.eng.stream_of[self.vn.slot]]
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    self._body()
                torch.cuda.current_stream().wait_stream(side)
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph, stream=side):
                    self._body()
            self.graph.replay()
        else:
            self._body()
        out_stats.copy_(self.stats, non_blocking=True)
        self.eng.kernel_launches += 1


class B200Network:
    """Blackwell-native counterpart of ``Network`` (same ``train``/``history`` contract)."""

    def __init__(self, config, model_factory: Callable[[], nn.Module], dataset_adapter, aggregator_factory,
                 device: Optional[torch.device] = None, criterion: Optional[nn.Module] = None, evidential: bool = False):
        from murmura_b200 import ops
        from murmura_b200.topology import create_topology
        from murmura_b200.utils.factories import build_attack, build_mobility_model

        if not torch.cuda.is_available():
            raise RuntimeError("backend 'b200' needs a CUDA device (use backend: simulation on CPU)")
        self.ext = ops.ext()                       # fail loudly if the sm_100a extension is missing
        self.cfg = config
        self.opt = config.b200
        self.rank, self.world, local_rank = init_distributed()
        if self.world > 1:
            self.device = torch.device("cuda", local_rank)
        elif device is not None and device.type == "cuda" and device.index is not None:
            self.device = device
        else:
            self.device = torch.device("cuda", torch.cuda.current_device())
        torch.cuda.set_device(self.device)
        self.is_primary = self.rank == 0
        self.evidential = evidential
        self.criterion = criterion
        self.N = config.topology.num_nodes
        self.history = new_history()
        self._apply_math_mode()

        # ---- topology / attack / mobility ----------------------------------------------------
        t = config.topology
        self.mobility = build_mobility_model(config)
        self.topology: Topology = create_topology(t.type, t.num_nodes, p=t.p, k=t.k, seed=t.seed)
        with (contextlib.redirect_stdout(io.StringIO()) if not self.is_primary else nullcontext()):
            self.attack = build_attack(config)
        self.compromised = set(self.attack.get_compromised_nodes()) if self.attack else set()
        self.aggregator = aggregator_factory(0)     # hyper-parameters + host-side statistics container
        self.family = getattr(self.aggregator, "kernel_family", "generic")

        # ---- arena ---------------------------------------------------------------------------
        weights = None
        if self.opt.placement == "balanced" and self.world > 1:
            bs_ = config.training.batch_size
            sizes = [len(p) for p in dataset_adapter.get_client_partitions()]
            weights = [0.0 if i in self.compromised else float(max(1, n // max(1, min(bs_, max(2, n))))) for i, n in enumerate(sizes[: self.N])]
        self.placement = Placement(self.N, self.world, weights)
        B200Network._NETS_BUILT += 1
        self._net_id = B200Network._NETS_BUILT
        self._barrier_until = int(self.opt.host_barrier_rounds)
        rng_state = torch.get_rng_state()
        probe = model_factory()
        if self.opt.seed_parity:
            torch.set_rng_state(rng_state)              # the layout probe must not advance the stream the node models are drawn from
        self.layout = StateLayout.from_model(probe, channels_last=bool(self.opt.channels_last))
        sketch_k = int(getattr(self.aggregator, "sketch_size", 0)) if self.family == "sketchguard" else 0
        auto = self.opt.transport == "auto"
        if auto:
            # full-mesh FedAvg over several GPUs: every node computes the same sum → let the NVSwitch add the ranks' copies
            # (each GPU ingests S·P instead of (N − V)·P bytes); everything else reads neighbour rows in-kernel over P2P
            full = all(len(set(self.topology.neighbors[i]) - {i}) == self.N - 1 for i in range(self.N))
            # (measured at 2 GPUs: one-shot peer loads of the rank-sum rows beat the multicast path, 0.09 vs 0.15 ms; the switch pays off
            # once G − 1 peer rows would have to cross the links)
            nvls = self.world >= 4 and self.family == "fedavg" and full and self.mobility is None and not self.opt.fault_drop_edges
            self.opt.transport = "nvls" if nvls else "p2p"
        try:
            self.arena = SymmetricArena(self.layout, self.placement, self.rank, self.device, sketch_size=sketch_k,
                                        backend="symm" if self.opt.transport == "nvls" else "ipc")
            ok = 1 if (self.opt.transport != "nvls" or self.arena.mc_base) else 0
        except Exception:  # noqa: BLE001 - symmetric memory / multicast unavailable on this box
            if not auto:
                raise
            ok = 0
        if auto and self.opt.transport == "nvls":
            if self.world > 1:                          # every rank must take the same path
                t = torch.tensor([ok], device=self.device)
                _dist().all_reduce(t, op=_dist().ReduceOp.MIN)
                ok = int(t.item())
            if not ok:
                self.opt.transport = "p2p"
                self.arena = SymmetricArena(self.layout, self.placement, self.rank, self.device, sketch_size=sketch_k, backend="ipc")
        S, L = self.placement.slots_per_rank, self.layout
        self.S = S
        self.live = self.arena.live
        self.ints = torch.zeros(S, max(L.Pi, 1), dtype=torch.int64, device=self.device)
        self.local_gids = self.placement.local_nodes(self.rank)
        self.V = len(self.local_gids)

        # ---- virtual nodes -------------------------------------------------------------------
        self.nodes: List[VirtualNode] = []
        self._host_shards: List[Tuple[torch.Tensor, torch.Tensor]] = []
        self.h2d_bytes_per_round = 0
        bs = config.training.batch_size
        parity_models = {}
        if self.opt.seed_parity:
            # the simulation backend builds node 0 … N−1 in order from the caller's global torch stream (core/network.py from_config,
            # reference murmura/core/network.py:262-300): replay exactly that on every rank, keep the local nodes' models
            mine = set(self.local_gids)
            for gid in range(self.N):
                m = model_factory()
                if gid in mine:
                    parity_models[gid] = m
        for slot, gid in enumerate(self.local_gids):
            if self.opt.seed_parity:
                model = parity_models.pop(gid).to(self.device)
            else:
                torch.manual_seed(config.experiment.seed * 1000003 + gid)   # per-node init stream, rank-layout independent
                model = model_factory().to(self.device)
            L.bind(model, self.live[slot], None, self.ints[slot] if L.Pi else None)
            X, y = dataset_adapter.client_tensors(gid)
            nhwc = bool(self.opt.channels_last) and X.dim() == 4
            if nhwc:
                X = X.permute(0, 2, 3, 1)                   # keep image shards physically NHWC
            if self.opt.stream_inputs:                      # end-to-end mode: shards live in pinned host memory
                xh, yh = X.float().contiguous().pin_memory(), y.long().contiguous().pin_memory()
                self._host_shards.append((xh, yh))
                self.h2d_bytes_per_round += xh.numel() * 4 + yh.numel() * 8
            X = X.to(self.device, non_blocking=True).float().contiguous()
            y = y.to(self.device, non_blocking=True).long().contiguous()
            n = int(y.shape[0])
            eb = min(bs, max(2, n))
            nb = (n // eb) if n > eb else (1 if n >= 2 else 0)
            eb = min(eb, n) if n >= 2 else eb
            self.nodes.append(VirtualNode(gid=gid, slot=slot, model=model, X=X, y=y, n=n, eb=eb, nb=nb,
                                          byzantine=gid in self.compromised, nhwc=nhwc,
                                          params=[p for p in model.parameters() if p.requires_grad]))
        del probe
        if not self.opt.seed_parity:                    # parity mode keeps consuming the caller's stream (shuffles below)
            torch.manual_seed(config.experiment.seed + 7919 * self.rank)
        self._shard_sizes = [len(p) for p in dataset_adapter.get_client_partitions()][: self.N] if self.opt.seed_parity else None

        # ---- attack parameters on the device (fused into publish) ------------------------------
        spec = self.attack.device_spec() if self.attack is not None and hasattr(self.attack, "device_spec") else None
        self.attack_spec = spec
        scale = torch.ones(max(self.V, 1)); noise = torch.zeros(max(self.V, 1))
        if spec is not None:
            for vn in self.nodes:
                if vn.byzantine:
                    scale[vn.slot] = spec["scale"]; noise[vn.slot] = spec["noise_std"]
        self.atk_scale = scale.to(self.device); self.atk_noise = noise.to(self.device)
        self.node_gid = torch.tensor(self.local_gids or [0], dtype=torch.int32, device=self.device)
        self.ticket = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.custom_attack = self.attack is not None and spec is None and any(vn.byzantine for vn in self.nodes) \
            and not (hasattr(self.attack, "device_spec"))

        # ---- streams / buffers ---------------------------------------------------------------
        self.main = torch.cuda.current_stream(self.device)
        # Worker streams.  A round's training phase is bounded by its critical path (the node with the most batches: steps of
        # one node are inherently sequential), so nodes are mapped to streams longest-first and the streams carrying the
        # longest chains get the highest CUDA priority — their kernels are scheduled ahead of the short chains' kernels.
        K = max(1, self.opt.streams) if self.opt.streams > 0 else max(1, min(16, len(self.nodes)))
        lo, hi = -5, 0
        try:
            import ctypes
            _lo, _hi = ctypes.c_int(), ctypes.c_int()
            if ctypes.CDLL("libcudart.so.12").cudaDeviceGetStreamPriorityRange(ctypes.byref(_lo), ctypes.byref(_hi)) == 0:
                hi, lo = _lo.value, _hi.value          # API returns (least, greatest); greatest priority is the smallest number
        except Exception:
            pass
        self.streams = [torch.cuda.Stream(self.device, priority=max(lo, min(hi, lo + k))) for k in range(K)]
        order = sorted(range(len(self.nodes)), key=lambda i: -(self.nodes[i].nb * max(self.nodes[i].eb, 1)))
        load = [0] * K
        self.stream_of = [0] * len(self.nodes)
        for i in order:                                   # LPT: next-longest node → least-loaded stream (ties → higher priority)
            k = min(range(K), key=lambda j: (load[j], j))
            self.stream_of[i] = k
            load[k] += self.nodes[i].nb * max(self.nodes[i].eb, 1) + 1
        self.launch_order = order
        # One capture stream per worker stream.  cuBLAS keeps one workspace per (handle, stream); graphs captured on
        # the same stream bake in the same workspace pointer, so graphs that may REPLAY concurrently must have been
        # captured on different streams (graphs of one worker stream replay back-to-back and may share).  The
        # workspaces are created here, outside any capture, so they are not owned by a graph's private pool.
        self.capture_streams = [torch.cuda.Stream(self.device) for _ in self.streams]
        for cs in self.capture_streams:
            with torch.cuda.stream(cs):
                torch.mm(torch.ones(64, 64, device=self.device), torch.ones(64, 64, device=self.device))
        torch.cuda.synchronize(self.device)
        self.eval_stats = torch.zeros(max(self.V, 1), _STAT_COLS, device=self.device)
        self.lam_t = torch.zeros((), device=self.device)
        self.round_idx = 0
        self.epoch = 0
        self._lr = None
        self._edge_cache: Dict[Any, Dict[str, torch.Tensor]] = {}
        self._evaluators: Dict[Any, ForeignEval] = {}
        self._mlp_plan = mlp_plan(self.nodes[0].model, self.layout) if (self.nodes and self.opt.grouped_mlp) else None
        self._mlp_bufs: Dict[Any, torch.Tensor] = {}
        self._row_cache: Dict[Any, torch.Tensor] = {}
        self._stat_log: List[torch.Tensor] = []
        self._agg_state: Dict[str, torch.Tensor] = {}
        self.timers: Dict[str, float] = {"train_ms": 0.0, "aggregate_ms": 0.0, "eval_ms": 0.0, "rounds": 0}
        self.kernel_launches = 0
        self._dmtt_init()
        self._sketch_init()

    # =========================================================================================
    # construction helpers
    # =========================================================================================
    @classmethod
    def from_config(cls, config, model_factory, dataset_adapter, aggregator_factory, device=None, criterion=None,
                    evidential=False) -> "B200Network":
        return cls(config, model_factory, dataset_adapter, aggregator_factory, device=device, criterion=criterion,
                   evidential=evidential)

    def _apply_math_mode(self) -> None:
        mode = self.opt.compute_dtype
        torch.backends.cudnn.benchmark = True
        _ops.set_fused_bn(self.opt.fused_bn and mode != "bf16")
        if mode in ("tf32", "bf16"):
            torch.backends.cuda.matmul.allow_tf32 = True
            torch.backends.cudnn.allow_tf32 = True
        self._autocast = (lambda: torch.autocast("cuda", dtype=torch.bfloat16)) if mode == "bf16" else nullcontext

    # =========================================================================================
    # edge tables
    # =========================================================================================
    def _edge_table(self, neighbors: List[List[int]], key: Any = None) -> Dict[str, torch.Tensor]:
        """Device CSR over this rank's destination slots; row = [self, *neighbours]."""
        if key is not None and key in self._edge_cache:
            return self._edge_cache[key]
        pl = self.placement
        row_ptr = [0]; rk: List[int] = []; sl: List[int] = []; gid: List[int] = []
        for vn in self.nodes:
            srcs = [vn.gid] + [j for j in neighbors[vn.gid] if j != vn.gid]
            if len(srcs) > 128:
                raise ValueError("degree + 1 > 128 is not supported by the fused kernels")
            rk += [int(pl.rank_of[j]) for j in srcs]; sl += [int(pl.slot_of[j]) for j in srcs]; gid += srcs
            row_ptr.append(len(rk))
        dev = self.device
        E = max(len(rk), 1)
        et = {
            "row_ptr": torch.tensor(row_ptr if self.V else [0, 0], dtype=torch.int32, device=dev),
            "src_rank": torch.tensor(rk or [0], dtype=torch.int32, device=dev),
            "src_slot": torch.tensor(sl or [0], dtype=torch.int32, device=dev),
            "src_gid": torch.tensor(gid or [0], dtype=torch.int32, device=dev),
            "mask": torch.ones(E, device=dev),
            "w": torch.zeros(E, device=dev), "w_tail": torch.zeros(E, device=dev),
            "d2": torch.zeros(E, device=dev), "dist": torch.zeros(E, device=dev),
            "n2": torch.zeros(max(self.V, 1), device=dev),
            "stats": torch.zeros(max(self.V, 1), 4, device=dev),
            "aux": torch.zeros(E, device=dev), "aux2": torch.zeros(E, device=dev), "aux3": torch.zeros(E, device=dev),
            "host_rows": row_ptr, "host_gid": gid, "host_rank": rk, "host_slot": sl,
            "max_m": max([row_ptr[i + 1] - row_ptr[i] for i in range(len(row_ptr) - 1)] or [1]),
        }
        if key is not None:
            self._edge_cache[key] = et
        return et

    def _apply_fault_mask(self, et: Dict[str, torch.Tensor], round_idx: int) -> None:
        drops = self.opt.fault_drop_edges.get(round_idx) or self.opt.fault_drop_edges.get(str(round_idx))
        et["mask"].fill_(1.0)
        if not drops:
            return
        dropped = {(int(a), int(b)) for a, b in drops}
        rows, gids = et["host_rows"], et["host_gid"]
        mask = torch.ones(len(gids))
        for vi, vn in enumerate(self.nodes):
            for e in range(rows[vi] + 1, rows[vi + 1]):
                if (gids[e], vn.gid) in dropped:
                    mask[e] = 0.0
        et["mask"].copy_(mask.to(self.device))

    def _neighbors_for_round(self, r: int) -> Tuple[List[List[int]], Any]:
        if self.dmtt_on:
            return self._dmtt_neighbors(r), None
        if self.mobility is not None:
            adj = self.mobility.neighbors_at(r)
            return [adj[i] for i in range(self.N)], None
        return self.topology.neighbors, "static"

    # =========================================================================================
    # training
    # =========================================================================================
    def _loss(self, out: torch.Tensor, yb: torch.Tensor) -> torch.Tensor:
        from murmura_b200 import ops
        from murmura_b200.models.mlp import EvidentialLoss
        if self.evidential and isinstance(self.criterion, EvidentialLoss):
            return ops.evidential_loss(out.float(), yb, self.lam_t)
        if self.criterion is not None and not self.evidential:
            return self.criterion(out.float(), yb)
        if self.criterion is not None:
            return self.criterion(out.float(), yb, epoch=self.round_idx)
        return F.cross_entropy(out.float(), yb)

    @staticmethod
    def _inputs(vn: VirtualNode, x: torch.Tensor) -> torch.Tensor:
        return x.permute(0, 3, 1, 2) if vn.nhwc else x        # zero-copy logical NCHW view of the NHWC shard

    def _split_backward(self, vn: VirtualNode):
        mode = self.opt.split_backward
        if mode == "auto":
            # Measured on B200: with ≥ 4 nodes training concurrently on a GPU the extra side streams alias onto the same
            # hardware queues (CUDA_DEVICE_MAX_CONNECTIONS) and serialise (flagship 36 → 21 rounds/s); with 1–2 nodes per GPU the
            # GPU is idle enough for the parallel branch to pay (ResNet-18 step 940 → 856 µs).
            mode = self.V <= 2
        if not mode or self.opt.compute_dtype == "bf16":
            return None
        if not vn.autotuned:
            # The first step of every node runs stock single-stream autograd on an idle GPU: cudnn.benchmark times its candidate
            # algorithms during that step, and timing them while a sibling kernel runs on the side stream picks bad ones.
            return None
        sb = getattr(vn, "split_bwd", None)
        if sb is None:
            from murmura_b200.parallel.split_backward import SplitBackward
            sb = vn.split_bwd = SplitBackward(self.device)
        return sb

    def _plain_ce(self) -> bool:
        c = self.criterion
        return (not self.evidential) and (c is None or (type(c) is nn.CrossEntropyLoss and c.weight is None and c.reduction == "mean"
                                                         and c.label_smoothing == 0.0 and c.ignore_index < 0))

    def _backward(self, vn: VirtualNode, out: torch.Tensor, yb: torch.Tensor) -> bool:
        """Loss + backward.  The two bundled losses are fused forward+backward kernels whose gradient seeds ``out.backward``
        directly (no autograd nodes for the loss, no ones-like seed, no ``loss_sum`` add); anything else is stock autograd."""
        from murmura_b200.models.mlp import EvidentialLoss
        fp32 = out.dtype == torch.float32 and out.dim() == 2
        if fp32 and self.evidential and isinstance(self.criterion, EvidentialLoss):
            loss, grad = self.ext.evidential_loss_fwd_bwd(out.detach().contiguous(), yb, 0.0, self.lam_t)
            out.backward(grad)
            vn.loss_sum += loss
            return True
        if fp32 and self._plain_ce():
            _, grad = self.ext.ce_loss_fwd_bwd(out.detach().contiguous(), yb, vn.loss_sum)
            out.backward(grad)
            return True
        loss = self._loss(out, yb)
        loss.backward()
        vn.loss_sum += loss.detach()
        return False

    def _train_step(self, vn: VirtualNode, lr: float) -> None:
        c0 = _ops.counters["launches"]
        # one launch: rows perm[step·eb …] of the shard → (xb, yb); the device-side step counter advances inside the kernel
        xb, yb = self.ext.gather_batch(vn.X, vn.y, vn.perm_buf, vn.step, vn.ticket, vn.eb)
        xb = self._inputs(vn, xb)
        for p in vn.params:
            p.grad = None                                      # autograd hands us its own grad buffers: no accumulate pass
        sb = self._split_backward(vn)
        with self._autocast(), (sb if sb is not None else nullcontext()):
            out = vn.model(xb)
        fused_loss = self._backward(vn, out, yb)
        if sb is not None:
            grads = sb.join(vn.params)                         # weight gradients were computed on the side stream
        else:
            grads = []
            for p in vn.params:
                g = p.grad
                if g.stride() != p.stride() or g.dtype != p.dtype:
                    g = torch.empty_like(p).copy_(g)
                grads.append(g)
        self.ext.sgd_multi(vn.params, grads, lr)              # one launch: θ -= lr·g for every tensor of the node
        vn.launches_per_step = (_ops.counters["launches"] - c0) + 2 + (1 if fused_loss else 0)   # + gather_batch, sgd_multi, loss
        if not vn.autotuned:
            torch.cuda.synchronize(self.device)
            vn.autotuned = True

    def _graphs_ok(self) -> bool:
        from murmura_b200.models.mlp import EvidentialLoss
        if not self.opt.cuda_graphs:
            return False
        return not (self.evidential and self.criterion is not None and not isinstance(self.criterion, EvidentialLoss))

    def _prepare_training(self, epochs: int, lr: float) -> None:
        key = (epochs, lr)
        cache = self.__dict__.setdefault("_fused_cache", {})
        if epochs not in cache:
            cache[epochs] = self._fused_setup(epochs)
        if cache[epochs] is not None:
            return                                              # the fused program replaces the per-node autograd graphs
        for vn in self.nodes:
            if vn.byzantine or vn.nb == 0:
                continue
            need = epochs * vn.nb * vn.eb
            if vn.perm_buf is None or vn.perm_buf.numel() != need:
                vn.perm_buf = torch.zeros(need, dtype=torch.int64, device=self.device)
                vn.step = torch.zeros((), dtype=torch.int64, device=self.device)
                vn.ticket = torch.zeros((), dtype=torch.int32, device=self.device)
                vn.arange = torch.arange(vn.eb, device=self.device)
                vn.loss_sum = torch.zeros((), device=self.device)
                vn.train_graph = None
            if self._graphs_ok() and (vn.train_graph is None or vn.graph_key != key):
                self._capture_train(vn, lr)
                vn.graph_key = key

    def _capture_train(self, vn: VirtualNode, lr: float) -> None:
        snap, snap_i = self.live[vn.slot].clone(), self.ints[vn.slot].clone()
        rng = torch.cuda.get_rng_state(self.device)
        vn.model.train()
        vn.perm_buf.copy_(torch.arange(vn.perm_buf.numel(), device=self.device) % vn.n)
        side = self.capture_streams[self.stream_of[vn.slot]]
        torch.cuda.synchronize(self.device)                # idle GPU while cuDNN autotunes in the first warm-up step
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                vn.step.zero_()
                self._train_step(vn, lr)
        torch.cuda.current_stream().wait_stream(side)
        vn.step.zero_()
        graph = torch.cuda.CUDAGraph()
        vn.steps_per_replay = (vn.perm_buf.numel() // vn.eb) if self.opt.unroll_round else 1
        with torch.cuda.graph(graph, stream=side):
            for _ in range(vn.steps_per_replay):          # the device-side step counter walks the permutation buffer
                self._train_step(vn, lr)
        vn.train_graph = graph
        self.live[vn.slot].copy_(snap); self.ints[vn.slot].copy_(snap_i)
        vn.step.zero_(); vn.loss_sum.zero_()
        torch.cuda.set_rng_state(rng, self.device)

    def _fork(self) -> None:
        ev = torch.cuda.Event(); ev.record(self.main)
        for s in self.streams:
            s.wait_event(ev)

    def _join(self) -> None:
        for s in self.streams:
            ev = torch.cuda.Event(); ev.record(s); self.main.wait_event(ev)

    def _parity_orders(self, epochs: int) -> Dict[int, torch.Tensor]:
        """Seed-parity mode: this round's sample orders of the LOCAL nodes, drawn from the global host stream exactly as the
        simulation backend's loaders draw them — nodes in id order, attackers skipped (they do not train), per epoch one int64
        for the iterator's base seed, one int64 seeding a private generator, ``randperm(n)`` from it (``data/fast_loader.py``,
        i.e. ``DataLoader(shuffle=True)``); every rank replays the whole sequence and keeps its own nodes."""
        from murmura_b200.data.fast_loader import replay_round_orders
        mine = {vn.gid for vn in self.nodes}
        orders = replay_round_orders(self._shard_sizes, int(self.cfg.training.batch_size), epochs, skip=self.compromised)
        return {gid: o for gid, o in orders.items() if gid in mine}

    def _local_training(self, epochs: int, lr: float) -> None:
        from murmura_b200.models.mlp import EvidentialLoss
        self._round_orders = self._parity_orders(epochs) if self.opt.seed_parity else None
        self._fused_evidential = self.evidential and isinstance(self.criterion, EvidentialLoss)
        if isinstance(self.criterion, EvidentialLoss):
            self.lam_t.fill_(self.criterion.anneal(self.round_idx))
        if self._fused_training(epochs, lr):
            return
        if self.opt.batched_mlp_train and self._batched_training(epochs, lr):
            return
        self._fork()
        for i in self.launch_order:
            vn = self.nodes[i]
            stream = self.streams[self.stream_of[i]]
            if self._host_shards:                             # per-round H2D of this node's inputs (pinned → HBM)
                with torch.cuda.stream(stream):
                    vn.X.copy_(self._host_shards[i][0], non_blocking=True)
                    vn.y.copy_(self._host_shards[i][1], non_blocking=True)
            if vn.byzantine or vn.nb == 0:
                continue
            with torch.cuda.stream(stream):
                take = vn.nb * vn.eb
                if self._round_orders is not None:
                    vn.perm_buf.copy_(self._round_orders[vn.gid].reshape(-1).to(self.device, non_blocking=True))
                else:
                    keys = torch.rand(epochs, vn.n, device=self.device)
                    vn.perm_buf.copy_(keys.argsort(dim=1)[:, :take].reshape(-1))
                vn.step.zero_(); vn.loss_sum.zero_()
                vn.model.train()
                if vn.train_graph is not None:
                    for _ in range(epochs * vn.nb // vn.steps_per_replay):
                        vn.train_graph.replay()
                else:
                    for _ in range(epochs * vn.nb):
                        self._train_step(vn, lr)
                self.kernel_launches += epochs * vn.nb * vn.launches_per_step
        self._join()

    # ---- fused program: every layer of every node of this GPU in ONE launch (parallel/fused_trainer.py) -------------------------
    def _fused_setup(self, epochs: int):
        """Build the fused trainer for this GPU's nodes; ``None`` when the model family / loss / layout is not supported."""
        from murmura_b200.models.mlp import EvidentialLoss
        from murmura_b200.parallel.fused_trainer import FusedTrainer
        mode = self.opt.fused_train
        if not mode or not self.nodes or self.opt.compute_dtype == "bf16" or not self.opt.cuda_graphs:
            return None
        if self.evidential and not isinstance(self.criterion, EvidentialLoss):
            return None
        if not self.evidential and not self._plain_ce():
            return None
        live = [vn for vn in self.nodes if not vn.byzantine and vn.nb > 0]
        if not live:
            return None
        first = self.nodes[0]
        if first.X.dim() == 4 and not first.nhwc:
            return None
        shape = (first.X.shape[3], first.X.shape[1], first.X.shape[2]) if first.X.dim() == 4 else tuple(first.X.shape[1:])
        trainers = []
        # one program per effective batch size (a shard smaller than the batch trains with batch = shard size, reference
        # core/network.py:280-287): the nodes of the other sizes simply have zero steps in it
        for eb in sorted({vn.eb for vn in live}, reverse=True):
            steps = [epochs * vn.nb if (not vn.byzantine and vn.nb > 0 and vn.eb == eb) else 0 for vn in self.nodes]
            tr = FusedTrainer(first.model, self.layout, self.live, self.ints if self.layout.Pi else None, [(vn.X, vn.y) for vn in self.nodes],
                              steps, eb, shape, evidential=self.evidential, seed=int(self.cfg.experiment.seed),
                              side_stream=bool(self.opt.fused_side_stream))
            if not tr.supported:
                if mode is True and self.is_primary:
                    print(f"[b200] fused_train unavailable for this model ({getattr(tr, 'unsupported_reason', 'unsupported family')}); "
                          "using per-node autograd graphs")
                return None
            tr.lam_t = self.lam_t                                # annealing coefficient is read on the device
            trainers.append(tr)
        return trainers

    def _fused_training(self, epochs: int, lr: float) -> bool:
        cache = self.__dict__.setdefault("_fused_cache", {})
        if epochs not in cache:
            cache[epochs] = self._fused_setup(epochs)
        trainers = cache[epochs]
        if trainers is None:
            return False
        if self._host_shards:                                    # end-to-end mode: this round's inputs come from pinned host memory
            for i, vn in enumerate(self.nodes):
                vn.X.copy_(self._host_shards[i][0], non_blocking=True)
                vn.y.copy_(self._host_shards[i][1], non_blocking=True)
        for tr in trainers:
            tr.loss_acc.zero_()
            before = tr.be.launches
            perms = None
            orders = getattr(self, "_round_orders", None)
            if orders is not None:
                perms = {vn.slot: orders[vn.gid] for vn in self.nodes if vn.gid in orders}
            tr.run_round(epochs, lr, perms=perms)
            if tr.be.launches > before:                          # the round was (re)captured: remember its launch count
                tr.launches_in_graph = (tr.be.launches - before) * tr.max_steps // (tr.max_steps + 1) if tr.max_steps else 0
            self.kernel_launches += getattr(tr, "launches_in_graph", 0)
            for vn in self.nodes:
                if tr.steps[vn.slot] > 0:
                    vn.loss_sum = tr.loss_acc[vn.slot]
        self.fused = trainers[0]
        return True

    # ---- K8b: all MLP nodes of this GPU in one batched step (opt-in: b200.batched_mlp_train) ------------------------------
    def _batched_setup(self, epochs: int, lr: float):
        """Build the batched trainer + the CUDA graph of a whole round; returns None when this GPU's nodes do not qualify."""
        from murmura_b200.models.mlp import MLP, EvidentialLoss, EvidentialMLP
        from murmura_b200.parallel.batched_mlp import BatchedMLPTrainer
        live = [vn for vn in self.nodes if not vn.byzantine and vn.nb > 0]
        if (self.V < 2 or not live or not isinstance(self.nodes[0].model, (MLP, EvidentialMLP)) or self._host_shards
                or len({vn.eb for vn in live}) != 1 or self.opt.compute_dtype == "bf16"):
            return None
        if self.evidential != isinstance(self.nodes[0].model, EvidentialMLP):
            return None
        if self.evidential and not isinstance(self.criterion, EvidentialLoss):
            return None
        if not self.evidential and not self._plain_ce():
            return None
        eb = live[0].eb
        steps = max(vn.nb for vn in live) * epochs
        n_max = max(max(vn.n for vn in self.nodes), 1)
        st: Dict[str, Any] = {"eb": eb, "steps": steps}
        st["xpad"] = torch.zeros(self.V, n_max, *self.nodes[0].X.shape[1:], device=self.device)
        st["ypad"] = torch.zeros(self.V, n_max, dtype=torch.long, device=self.device)
        for vi, vn in enumerate(self.nodes):
            st["xpad"][vi, :vn.n].copy_(vn.X); st["ypad"][vi, :vn.n].copy_(vn.y)
        act = torch.zeros(steps, self.V, device=self.device)
        for vi, vn in enumerate(self.nodes):
            if not vn.byzantine and vn.nb > 0:
                act[: vn.nb * epochs, vi] = 1.0
        st["act"] = act
        st["perm"] = torch.zeros(self.V, steps * eb, dtype=torch.long, device=self.device)
        st["trainer"] = trainer = BatchedMLPTrainer(self.nodes[0].model, self.layout, self.live[: self.V], self.ints[: self.V] if self.layout.Pi else None)
        lam = self.lam_t if self.evidential else 0.0

        def body():
            for t in range(steps):
                xb, yb = trainer.gather(st["xpad"], st["ypad"], st["perm"][:, t * eb:(t + 1) * eb])
                trainer.step(xb, yb, act[t], lr, lam)

        st["body"] = body
        st["graph"] = None
        if self._graphs_ok():
            snap, snap_i = self.live.clone(), self.ints.clone()
            rng = torch.cuda.get_rng_state(self.device)
            side = self.capture_streams[0]
            torch.cuda.synchronize(self.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                body()
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                body()
            st["graph"] = graph
            self.live.copy_(snap); self.ints.copy_(snap_i)
            torch.cuda.set_rng_state(rng, self.device)
        return st

    def _batched_training(self, epochs: int, lr: float) -> bool:
        key = (epochs, lr)
        cache = self.__dict__.setdefault("_batched_cache", {})
        if key not in cache:
            try:
                cache[key] = self._batched_setup(epochs, lr)
            except Exception as exc:  # noqa: BLE001 - opt-in fast path: report and use the per-node graphs
                if self.is_primary:
                    print(f"[b200] batched MLP training unavailable ({type(exc).__name__}: {exc}); using per-node graphs")
                cache[key] = None
        st = cache[key]
        if st is None:
            return False
        eb = st["eb"]
        for vi, vn in enumerate(self.nodes):                       # fresh shuffles: epochs × nb batches of every node
            if vn.byzantine or vn.nb == 0:
                continue
            take = vn.nb * eb
            keys = torch.rand(epochs, vn.n, device=self.device)
            st["perm"][vi, : epochs * take].copy_(keys.argsort(dim=1)[:, :take].reshape(-1))
        for vn in self.nodes:
            vn.model.train()
        if st["graph"] is not None:
            st["graph"].replay()
        else:
            st["body"]()
        self.kernel_launches += st["steps"]
        return True

    # =========================================================================================
    # publish + aggregation plans
    # =========================================================================================
    def _sync_args(self) -> Tuple[int, int, int, float, int]:
        if self.world == 1:
            return 0, 1, 0, 0.0, 0
        if getattr(self, "_liveness_frozen", False):
            # liveness was decided ONCE for this round by the wait_epoch launch: later kernels do not spin, they only apply the
            # frozen timed-out mask (no torn rows when a flag lands near the deadline)
            return 0, self.world, 0, 0.0, self.arena.timed_out_ptr()
        return self.arena.flags_ptr(), self.world, self.epoch, float(self.opt.flag_timeout_ms), self.arena.timed_out_ptr()

    def _publish(self, parity: int, with_sum: bool = False) -> None:
        L = self.layout
        self.epoch += 1
        if self.custom_attack:
            self._publish_custom(parity)
        if with_sum and self.V:
            # full-mesh FedAvg: the publish pass also leaves Σ_v published_v in this rank's rsum row
            self.ext.publish_sum(self.live, self.arena.pub_plane_ptr(parity), self.arena.rsum_ptr(parity), L.stride, self.V, L.Pf, L.Pf_pad,
                                 self.ints if L.Pi else None, self.atk_scale, self.atk_noise, self.node_gid,
                                 int(self.cfg.experiment.seed), int(self.round_idx),
                                 self.arena.tbl_flags.data_ptr() if self.world > 1 else 0, self.world, self.rank, self.epoch, self.ticket)
            self.kernel_launches += 1
            return
        if self.V == 0:                         # a rank hosting no node still has to raise its epoch flag
            self.ext.publish(self.live, self.arena.pub_plane_ptr(parity), L.stride, 1, 0, 0, None, self.atk_scale,
                             self.atk_noise, self.node_gid, 0, 0, self.arena.tbl_flags.data_ptr() if self.world > 1 else 0,
                             self.world, self.rank, self.epoch, self.ticket)
            return
        self.ext.publish(self.live, self.arena.pub_plane_ptr(parity), L.stride, self.V, L.Pf, L.Pf_pad,
                         self.ints if L.Pi else None, self.atk_scale, self.atk_noise, self.node_gid,
                         int(self.cfg.experiment.seed), int(self.round_idx),
                         self.arena.tbl_flags.data_ptr() if self.world > 1 else 0, self.world, self.rank, self.epoch, self.ticket)
        self.kernel_launches += 1
        if self.custom_attack:
            self._publish_custom_fixup(parity)

    def _publish_custom(self, parity: int) -> None:
        """Arbitrary user ``Attack`` objects (no ``device_spec``): run them on state-dict views."""
        self._custom_rows = {}
        for vn in self.nodes:
            if vn.byzantine:
                state = {k: v.clone() for k, v in self.layout.row_views(self.live[vn.slot], self.ints[vn.slot]).items()}
                self._custom_rows[vn.slot] = self.attack.apply_attack(node_id=vn.gid, model_state=state, round_num=self.round_idx)

    def _publish_custom_fixup(self, parity: int) -> None:
        for slot, state in self._custom_rows.items():
            row = self.arena.pub[parity, slot]
            for e in self.layout.float_entries():
                row[e.offset:e.offset + e.numel].copy_(state[e.name].reshape(-1).to(row.dtype))

    def _gather(self, et, parity: int, renorm: bool, sync: bool = True) -> None:
        L = self.layout
        fp, G, ep, to, tp = self._sync_args() if sync else (0, 1, 0, 0.0, 0)
        self.ext.weighted_gather(self.live, self.arena.tbl_pub.data_ptr(), self.arena.parity_off(parity), L.stride, self.V,
                                 et["row_ptr"], et["src_rank"], et["src_slot"], et["mask"], et["w"], L.Pf_pad, renorm,
                                 fp, G, ep, to, tp, self.opt.gather_impl == "tma" or (self.opt.gather_impl == "auto" and self.world > 1))
        self.kernel_launches += 1
        if L.Pi:
            self.ext.tail_blend(self.live, self.arena.tbl_pub.data_ptr(), self.arena.parity_off(parity), L.stride, self.V,
                                et["row_ptr"], et["src_rank"], et["src_slot"], et["mask"], et["w_tail"], L.Pf_pad, self.ints,
                                tp)
            self.kernel_launches += 1

    def _edge_dist(self, et, parity: int, length: int) -> None:
        L = self.layout
        fp, G, ep, to, tp = self._sync_args()
        self.ext.edge_distances(self.live, self.arena.tbl_pub.data_ptr(), self.arena.parity_off(parity), L.stride, self.V,
                                et["row_ptr"], et["src_rank"], et["src_slot"], et["mask"], length, et["d2"], et["n2"],
                                fp, G, ep, to, tp)
        self.kernel_launches += 3

    def _et_args(self, et):
        return (self.V, et["row_ptr"], et["src_rank"], et["src_slot"], et["mask"], et["w"], et["w_tail"], et["stats"])

    def _log_stats(self, et) -> None:
        self._stat_log.append(et["stats"][: self.V].clone())

    # ---- FedAvg ------------------------------------------------------------------------------
    def _two_shot(self) -> bool:
        v = self.opt.fullmesh_two_shot
        return self.world >= 4 if v == "auto" else (bool(v) and self.world > 1)

    def _fullmesh_fedavg(self, et) -> bool:
        """True when every node averages the SAME set (fully connected FedAvg, no dropped edges, device-side attack): the
        round then moves one per-rank sum row instead of every node's row (``publish_sum`` → ``fedavg_fullmesh``)."""
        if self.family != "fedavg" or self.custom_attack or self.opt.fault_drop_edges or self.opt.transport == "nccl":
            return False
        if "full_mesh" not in et:
            rows = et["host_rows"]
            et["full_mesh"] = all(rows[i + 1] - rows[i] == self.N for i in range(len(rows) - 1))
        return bool(et["full_mesh"]) and self.opt.fullmesh_rank_sum

    def _agg_fedavg(self, et, parity: int) -> None:
        if "fedavg_ready" not in et:
            self.ext.fedavg_weights(*self._et_args(et)); et["fedavg_ready"] = True
            rows = et["host_rows"]
            et["full_mesh"] = all(rows[i + 1] - rows[i] == self.N for i in range(len(rows) - 1))
            et["byz"] = torch.tensor([1 if vn.byzantine else 0 for vn in self.nodes] or [0], dtype=torch.uint8, device=self.device)
        if et.get("rank_sum"):
            if not hasattr(self, "_rank_nodes"):
                self._rank_nodes = torch.tensor([int(c) for c in self.placement.counts] + [0] * 16, dtype=torch.int32, device=self.device)[: max(self.world, 1)].contiguous()
            L = self.layout
            use_mc = self.opt.transport == "nvls" and self.world > 1 and self.arena.mc_base
            tot = 0
            if self._two_shot():
                # reduce-scatter + all-gather of the rank sums in one kernel over peer / multicast memory, second epoch flag, then the
                # apply kernel reads the reduced row from LOCAL memory
                self.epoch += 1
                self.ext.fullmesh_reduce_scatter(self.live, self.arena.tbl_rsum[parity].data_ptr(), self.arena.mc_rsum_ptr(parity) if use_mc else 0,
                                                 self.arena.tbl_tot[parity].data_ptr(), self.arena.mc_tot_ptr(parity) if use_mc else 0,
                                                 L.Pf_pad, self.world, self.rank, self._sync_args()[4], self.arena.tbl_flags.data_ptr(),
                                                 self.epoch, self.ticket)
                self.kernel_launches += 1
                self._freeze_liveness()
                tot = self.arena.tot_ptr(parity)
            self.ext.fedavg_fullmesh(self.live, self.arena.pub_plane_ptr(parity), self.arena.tbl_rsum[parity].data_ptr(),
                                     self.arena.mc_rsum_ptr(parity) if use_mc else 0, L.stride, self.V, L.Pf_pad, self.N, self.world,
                                     et["byz"], self._rank_nodes, self._sync_args()[4], tot)
            self.kernel_launches += 1
            return                                            # int buffers keep own under FedAvg: nothing to blend
        if (self.opt.transport == "nvls" and self.world > 1 and et["full_mesh"] and self.arena.mc_base
                and not self.opt.fault_drop_edges):
            # full mesh ⇒ every node computes the same mean: let the NVSwitch do the cross-GPU sum (multimem.ld_reduce)
            L = self.layout
            fp, G, ep, to, tp = self._sync_args()
            self.ext.nvls_fedavg(self.live, self.arena.pub_plane_ptr(parity), self.arena.mc_pub_plane_ptr(parity), L.stride, self.V,
                                 self.S, L.Pf_pad, self.N, et["byz"], fp, G, ep, to, tp)
            self.kernel_launches += 1
            return                                            # int buffers keep own under FedAvg: nothing to blend
        self._gather(et, parity, renorm=True)

    # ---- BALANCE -----------------------------------------------------------------------------
    def _agg_balance(self, et, parity: int) -> None:
        a = self.aggregator
        self._edge_dist(et, parity, self.layout.stride)          # all keys: float region + int tail
        factor = decayed_factor(a.gamma, a.kappa, self.round_idx, a.total_rounds)
        self.ext.balance_filter(*self._et_args(et), et["d2"], et["n2"], et["dist"], factor, a.alpha, a.min_neighbors,
                                self._sync_args()[4])
        self.kernel_launches += 1
        self._log_stats(et)
        self._gather(et, parity, renorm=False, sync=False)

    # ---- Sketchguard -------------------------------------------------------------------------
    def _sketch_init(self) -> None:
        if self.family != "sketchguard":
            return
        from murmura_b200.aggregation.sketchguard import pack_sketch_tables
        a, L = self.aggregator, self.layout
        perm = L.ref_permutation()
        packed_ref = pack_sketch_tables(a.hash_table, a.sign_table)
        table = np.zeros(_ceil4(L.Pf), dtype=np.uint16)
        ok = perm >= 0
        table[: L.Pf][ok] = packed_ref[perm[ok]]
        self.sk_table = torch.from_numpy(table.view(np.int16)).to(self.device)
        self.sk_own = torch.zeros(max(self.V, 1), a.sketch_size, device=self.device)
        self.sk_hist = torch.zeros(max(self.V, 1), 4, device=self.device)
        self.sk_slots = torch.arange(max(self.V, 1), dtype=torch.int32, device=self.device)
        self.sk_fp8 = self.opt.sketch_dtype == "fp8"

    def _agg_sketchguard(self, et, parity: int) -> None:
        a, L, ar = self.aggregator, self.layout, self.arena
        K = a.sketch_size
        # sketches of the published rows go into the symmetric region (read by neighbours), sketches of
        # the live rows stay local; both in one pass each over the data already resident in L2/HBM
        self.ext.count_sketch(ar.pub_plane_ptr(parity), L.stride, self.sk_slots[: self.V], self.sk_table, L.Pf, K,
                              ar.sketch[parity, : self.V])
        self.ext.count_sketch(self.live.data_ptr(), L.stride, self.sk_slots[: self.V], self.sk_table, L.Pf, K,
                              self.sk_own[: self.V])
        if self.sk_fp8:
            self.ext.sketch_quant_mxfp8(ar.sketch[parity, : self.V], ar.sketch_q_ptr() + parity * self.S * ar.Kpad,
                                        ar.sketch_sc_ptr() + parity * self.S * (ar.Kpad // 32), ar.Kpad)
        self.kernel_launches += 5
        if self.world > 1:                              # sketches are written after the publish flag → second epoch
            self._signal_epoch()
        factor = decayed_factor(a.gamma, a.kappa, self.round_idx, a.total_rounds)
        fp, G, ep, to, tp = self._sync_args()
        self.ext.sketchguard_filter(*self._et_args(et), self.sk_own, ar.tbl_sketch.data_ptr(), ar.tbl_sketch_q.data_ptr(),
                                    ar.tbl_sketch_sc.data_ptr(), parity * self.S, K, max(ar.Kpad, 32), self.sk_fp8, factor,
                                    a.alpha, a.min_neighbors, self.sk_hist, et["dist"], fp, G, ep, to, tp)
        self.kernel_launches += 1
        self._log_stats(et)
        self._gather(et, parity, renorm=False, sync=False)

    def _signal_epoch(self) -> None:
        """Extra release of the epoch counter (zero-length publish) after auxiliary symmetric writes."""
        L = self.layout
        self.epoch += 1
        self.ext.publish(self.live, self.arena.pub_plane_ptr(0), L.stride, 1, 0, 0, None,
                         self.atk_scale, self.atk_noise, self.node_gid, 0, 0, self.arena.tbl_flags.data_ptr(), self.world,
                         self.rank, self.epoch, self.ticket)
        self.kernel_launches += 1
        self._freeze_liveness()                           # the consumers of the auxiliary data wait for THIS epoch

    # ---- Krum --------------------------------------------------------------------------------
    def _krum_tables(self, et) -> None:
        if "krum_D" not in et:
            et["krum_D"] = torch.zeros(max(self.V, 1), 32, 32, device=self.device)
            et["krum_win"] = torch.zeros(max(self.V, 1), dtype=torch.int32, device=self.device)

    def _gram_plan(self, et) -> Optional[Dict[str, Any]]:
        """Tile plan for the tcgen05 Gram: rows = live+published planes of every rank."""
        if self.opt.krum_gram != "tcgen05":
            return None                                  # auto = exact fp32 differences (the reference's distances); the Gram is opt-in
        if "gram_plan" in et:
            return et["gram_plan"]
        S, G, L = self.S, self.world, self.layout
        gpr = (S + 7) // 8                              # 8-row groups per plane per rank
        ngroups = 2 * G * gpr
        plan = None
        if ngroups <= 16 and et["max_m"] <= 32:
            row_of_live = lambda r, s: (2 * r * gpr + s // 8) * 8 + s % 8
            row_of_pub = lambda r, s: ((2 * r + 1) * gpr + s // 8) * 8 + s % 8
            idx = torch.zeros(max(self.V, 1), 32, dtype=torch.int64)
            rows, rk, sl = et["host_rows"], et["host_rank"], et["host_slot"]
            for vi, vn in enumerate(self.nodes):
                for c, e in enumerate(range(rows[vi], rows[vi + 1])):
                    idx[vi, c] = row_of_live(self.rank, vn.slot) if c == 0 else row_of_pub(rk[e], sl[e])
            kbs = int(self.ext.gram_kb_per_stage(ngroups))
            maps = self.ext.gram_make_maps([self.arena.base_ptr(r) for r in range(G)], 3 * S, L.stride, L.Pf_pad, kbs)
            nkb = L.Pf_pad // 32
            lo, hi = nkb * self.rank // G, nkb * (self.rank + 1) // G
            lo, hi = lo // kbs * kbs, (hi // kbs * kbs if self.rank + 1 < G else nkb)      # stage-aligned K shards
            plan = {"maps": maps, "gpr": gpr, "idx": idx.to(self.device), "kb": (lo, hi), "R": ngroups * 8,
                    "out": torch.zeros(128 * 128, device=self.device)}
        et["gram_plan"] = plan
        return plan

    def _agg_krum(self, et, parity: int) -> None:
        L = self.layout
        self._krum_tables(et)
        plan = self._gram_plan(et)
        fp, G, ep, to, tp = self._sync_args()
        if plan is not None:
            if self.world > 1:
                self._host_wait_epoch()
            S, gpr = self.S, plan["gpr"]
            group_map, group_y = [], []
            for r in range(self.world):                   # per rank: live-plane groups, then published[parity] groups
                for j in range(gpr):
                    group_map.append(r); group_y.append(8 * j)
                for j in range(gpr):
                    group_map.append(r); group_y.append((1 + parity) * S + 8 * j)
            self.ext.gram_tf32(plan["maps"], group_map, group_y, plan["kb"][0], plan["kb"][1], plan["R"], plan["out"], True, 0)
            self.kernel_launches += 2
            Gm = plan["out"].view(128, 128)
            if self.world > 1:
                _dist().all_reduce(Gm)
            diag = Gm.diagonal()
            idx = plan["idx"]
            norms = diag[idx].contiguous()                # ‖θ‖² of every candidate of every local node, [V, 32]
            D = norms.unsqueeze(2) + norms.unsqueeze(1) - 2.0 * Gm[idx.unsqueeze(2), idx.unsqueeze(1)]
            et["krum_D"].copy_(D.clamp_min_(0.0))
            # TF32 Gram error ∝ ‖θ‖²: pairs whose distance drowns in it (converged honest models) are recomputed exactly
            if "krum_scratch" not in et:
                et["krum_scratch"] = torch.zeros_like(et["krum_D"]); et["krum_refined"] = torch.zeros(max(self.V, 1), device=self.device)
            self.ext.krum_refine(self.live, self.arena.tbl_pub.data_ptr(), self.arena.parity_off(parity), L.stride, self.V,
                                 et["row_ptr"], et["src_rank"], et["src_slot"], et["mask"], L.Pf_pad, et["krum_D"], norms,
                                 float(self.opt.krum_refine_tau), et["krum_scratch"], et["krum_refined"])
            self.kernel_launches += 2
        else:
            self.ext.pairwise_distances(self.live, self.arena.tbl_pub.data_ptr(), self.arena.parity_off(parity), L.stride, self.V,
                                        et["row_ptr"], et["src_rank"], et["src_slot"], et["mask"], L.Pf_pad, et["krum_D"],
                                        et["max_m"], fp, G, ep, to, tp)
            self.kernel_launches += 2
        self.ext.krum_select(*self._et_args(et), et["krum_D"], int(self.aggregator.num_compromised), et["krum_win"], tp)
        self.kernel_launches += 1
        self._log_stats(et)
        self._gather(et, parity, renorm=False, sync=plan is not None)

    def _host_wait_epoch(self) -> None:
        """Block the *stream* (not the host) until all ranks published (1-warp spin kernel)."""
        fp, G, ep, to, tp = self._sync_args()
        if fp:
            self.ext.wait_epoch(self.live, fp, G, ep, to, tp)
            self.kernel_launches += 1

    _NETS_BUILT = 0                                     # construction order is identical on every rank: names the host barriers

    def _host_barrier(self) -> None:
        """CPU-only rendezvous between publish and the flag wait during warm-up rounds (``parallel/hostsync.py`` explains why)."""
        from torch.distributed import distributed_c10d as c10d
        from murmura_b200.parallel.hostsync import store_barrier
        try:
            store = c10d._get_default_store()
        except Exception:  # noqa: BLE001 - no store (single process): nothing to do
            return
        self._hb_count = getattr(self, "_hb_count", 0) + 1
        store_barrier(store, f"murmura_b200/hb/{self._net_id}/{self._hb_count}", self.world, max(1.0, self.opt.flag_timeout_ms / 1000.0))

    def _freeze_liveness(self) -> None:
        """One wait on the epoch flags decides which peers arrived; every later kernel of the round consumes that mask."""
        if self.world == 1:
            return
        if self.round_idx < self._barrier_until or getattr(self, "_warm_rounds", 0) > 0:
            self._host_barrier()
        self._liveness_frozen = False
        self._host_wait_epoch()
        self._liveness_frozen = True
        if not hasattr(self, "_timeout_acc"):
            self._timeout_acc = torch.zeros(1, dtype=torch.int32, device=self.device)
            self._timeout_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        self._timeout_acc |= self.arena.timed_out

    # ---- forward evaluation of foreign weights (UBAR stage 2, EvidentialTrust, DMTT scoring) ----------------
    def _forward_with(self, vn: VirtualNode, state: Optional[Dict[str, torch.Tensor]], xb: torch.Tensor) -> torch.Tensor:
        from murmura_b200.ops import fast_eval_batchnorm
        vn.model.eval()
        with torch.no_grad(), self._autocast(), fast_eval_batchnorm():
            if state is None:
                return vn.model(xb).float()
            return torch.func.functional_call(vn.model, state, (xb,)).float()

    # ---- grouped tcgen05 forward of foreign MLP weights ---------------------------------------------------------
    def _row_ptr(self, rank: int, parity: Optional[int], slot: int) -> int:
        """Device address of a published row (``parity`` 0/1) or of a local live row (``parity`` None)."""
        ar, L = self.arena, self.layout
        if parity is None:
            return ar.base_ptr(self.rank) + ar.off_live + slot * L.stride * 4
        return ar.base_ptr(rank) + ar.off_pub + (parity * self.S + slot) * L.stride * 4

    def _buf(self, key: Any, shape: Tuple[int, ...], dtype=torch.float32) -> torch.Tensor:
        t = self._mlp_bufs.get(key)
        if t is None or tuple(t.shape) != tuple(shape):
            t = self._mlp_bufs[key] = torch.zeros(*shape, dtype=dtype, device=self.device)
        return t

    def _ints(self, values: Sequence[int], dtype=torch.int64) -> torch.Tensor:
        """Device copy of a small host integer list WITHOUT a host synchronisation.  ``torch.tensor(list, device=cuda)`` is a
        pageable H2D copy that blocks the host until the stream has drained (≈ 0.5 ms each behind a training phase — most of a
        round for the small models).  Lists repeat round after round on static topologies, so the copies are cached by content;
        misses go through a fresh pinned buffer with a non-blocking copy."""
        key = (dtype, tuple(values))
        cache = self.__dict__.setdefault("_ints_cache", {})
        t = cache.get(key)
        if t is None:
            if len(cache) >= 256:                            # dynamic topologies: bounded, oldest entries go first
                for k in list(cache)[:64]:
                    del cache[k]
            host = torch.tensor(list(values) or [0], dtype=dtype).pin_memory()
            t = cache[key] = host.to(self.device, non_blocking=True)
            self.__dict__.setdefault("_ints_pinned", []).append(host)     # keep the staging buffers alive until their copies ran
            if len(self._ints_pinned) > 512:
                del self._ints_pinned[:256]
        return t

    def _grouped_mlp_scores(self, jobs: List[Tuple[int, int, int]], inputs: Dict[int, Tuple[torch.Tensor, torch.Tensor]],
                            kind: str, stats: torch.Tensor) -> None:
        """``jobs`` = (stats_row, destination vi, weight-row address).  One launch per layer for ALL jobs."""
        if not jobs:
            return
        plan = self._mlp_plan
        G = len(jobs)
        max_m = max(inputs[vi][0].shape[0] for _, vi, _ in jobs)
        host = np.zeros((G, 9), dtype=np.int64)
        prev = None
        for li, layer in enumerate(plan):
            out = self._buf(("act", li, G, max_m), (G, max_m, layer["N"]))
            for gi, (_, vi, base) in enumerate(jobs):
                x, _ = inputs[vi]
                host[gi, 0] = x.data_ptr() if li == 0 else prev[gi].data_ptr()
                host[gi, 1] = base + layer["w"] * 4
                host[gi, 2] = base + layer["b"] * 4 if layer["b"] is not None else 0
                if layer["bn"] is not None:
                    mean, var, gam, bet = layer["bn"]
                    host[gi, 3], host[gi, 4] = base + mean * 4, base + var * 4
                    host[gi, 5] = base + gam * 4 if gam is not None else 0
                    host[gi, 6] = base + bet * 4 if bet is not None else 0
                else:
                    host[gi, 3:7] = 0
                host[gi, 7] = out[gi].data_ptr()
                host[gi, 8] = x.shape[0]
            desc = torch.from_numpy(host.copy()).to(self.device, non_blocking=True)
            self.ext.grouped_linear_tf32(desc, G, max_m, layer["K"], layer["N"], layer["K"], layer["N"], layer["act"],
                                         float(layer.get("eps", 1e-5)))
            prev = out
        ev = np.zeros((G, 3), dtype=np.int64)
        for gi, (_, vi, _) in enumerate(jobs):
            ev[gi] = (prev[gi].data_ptr(), inputs[vi][1].data_ptr(), inputs[vi][0].shape[0])
        tmp = self._buf(("stats", G), (G, _STAT_COLS))
        tmp.zero_()
        self.ext.grouped_eval(torch.from_numpy(ev).to(self.device, non_blocking=True), G, max_m, plan[-1]["N"], plan[-1]["N"],
                              kind == "dirichlet", tmp)
        stats.index_copy_(0, self._ints([j[0] for j in jobs]), tmp)
        self.kernel_launches += len(plan) + 1

    def _flat_inputs(self, vn: VirtualNode, idx: Optional[torch.Tensor], tag: str) -> Tuple[torch.Tensor, torch.Tensor]:
        x = vn.X if idx is None else vn.X.index_select(0, idx)
        y = vn.y if idx is None else vn.y.index_select(0, idx)
        xb = self._buf((tag, vn.slot, "x"), (x.shape[0], int(np.prod(x.shape[1:]))))
        yb = self._buf((tag, vn.slot, "y"), (x.shape[0],), dtype=torch.long)
        xb.copy_(x.reshape(x.shape[0], -1)); yb.copy_(y)
        return xb, yb

    def _arange(self, n: int) -> torch.Tensor:
        t = self._row_cache.get(("arange", n))
        if t is None:
            t = self._row_cache[("arange", n)] = torch.arange(n, device=self.device)
        return t

    def _evaluator(self, vn: VirtualNode, rows: int, kind: str) -> ForeignEval:
        key = (vn.slot, rows, kind)
        ev = self._evaluators.get(key)
        if ev is None:
            ev = self._evaluators[key] = ForeignEval(self, vn, rows, kind)
        return ev

    def _src_row(self, rank: int, parity: int, slot: int) -> torch.Tensor:
        key = (rank, parity, slot)
        row = self._row_cache.get(key)
        if row is None:
            row = self._row_cache[key] = self.arena.peer_row(rank, parity, slot)
        return row

    # ---- fused scoring of foreign weights: ONE grouped forward program for every (destination, candidate) pair -------------------
    def _score_program(self, rows: int, capacity: int):
        """``FusedForward`` over ``capacity`` (destination, candidate) groups of ``rows`` samples each, or ``None`` when the model
        family is not covered by the fused tape (then the per-candidate graph replays of :class:`ForeignEval` are used)."""
        from murmura_b200.parallel.fused_trainer import FusedForward
        if not self.opt.fused_train or not self.nodes or self.opt.compute_dtype == "bf16":
            return None
        key = ("score", rows)
        st = self._evaluators.get(key)
        if st is not None and st["cap"] >= capacity:
            return st
        first = self.nodes[0]
        if first.X.dim() == 4 and not first.nhwc:
            return None
        shape = (first.X.shape[3], first.X.shape[1], first.X.shape[2]) if first.X.dim() == 4 else tuple(first.X.shape[1:])
        cap = max(capacity, 16)
        fe = FusedForward(first.model, self.layout, self.live, rows, shape, cap, evidential=self.evidential)
        if not fe.supported:
            self._evaluators[key] = {"cap": 1 << 30, "fe": None}
            return self._evaluators[key]
        dev = self.device
        st = {"cap": cap, "fe": fe, "row_tab": torch.zeros(cap, dtype=torch.int64, device=dev),
              "perm": torch.zeros(self.V, rows, dtype=torch.int64, device=dev),
              "x_tab": torch.tensor([vn.X.data_ptr() for vn in self.nodes], dtype=torch.int64, device=dev),
              "y_tab": torch.tensor([vn.y.data_ptr() for vn in self.nodes], dtype=torch.int64, device=dev),
              "stats": torch.zeros(cap, _STAT_COLS, device=dev)}
        fe.row_tab = st["row_tab"]
        self._evaluators[key] = st
        return st

    def _fused_scores(self, jobs: List[Tuple[int, int, int]], samples: Dict[int, torch.Tensor], rows: int, kind: str, stats: torch.Tensor) -> bool:
        """``jobs`` = (stats row, destination vi, candidate row address); ``samples[vi]`` = sample indices of the destination
        (≤ ``rows``).  Everything is enqueued on the current stream; no host synchronisation."""
        if not jobs:
            return True
        st = self._score_program(rows, len(jobs))
        if st is None or st["fe"] is None:
            return False
        fe, G = st["fe"], len(jobs)
        # group the jobs by the GPU that holds the candidate row: the TMA weight maps are per source arena (local or peer-mapped)
        ar, plane_bytes = self.arena, 3 * self.S * self.layout.stride * 4
        bases = [ar.base_ptr(r) for r in range(self.world)]

        def locate(addr: int) -> Tuple[int, int]:
            for r, b in enumerate(bases):
                if 0 <= addr - b < plane_bytes:
                    return r, (addr - b) // (self.layout.stride * 4)
            raise RuntimeError("candidate row outside every arena")

        located = [locate(int(j[2])) for j in jobs]
        remote = sorted({loc for loc in located if loc[0] != self.rank})
        if remote and self.opt.score_tma:
            # candidates that live on peer GPUs are pulled over NVLink ONCE per round into a local mirror (every destination of
            # this GPU that scores the same candidate, and every M tile of its GEMMs, then reads local HBM / L2)
            self._mirror_cap = max(self.N - self.V, len(remote), 1)
            mirror = self._buf(("score_mirror",), (self._mirror_cap, self.layout.stride))
            if getattr(self, "_mirror_epoch", None) != self.epoch:       # UBAR and DMTT score the same published rows: copy once
                self._mirror_epoch, self._mirror_where = self.epoch, {}
            where = self._mirror_where
            for loc in remote:
                if loc not in where:
                    r, ps = loc
                    where[loc] = len(where)
                    mirror[where[loc]].copy_(ar._view(r, ps * self.layout.stride * 4, [self.layout.stride], torch.float32), non_blocking=True)
            mbase = mirror.data_ptr()
            bases = bases + [mbase]
            jobs = [(j[0], j[1], mbase + where[loc] * self.layout.stride * 4) if loc in where else j for j, loc in zip(jobs, located)]
            located = [(self.world, where[loc]) if loc in where else loc for loc in located]
        order = sorted(range(G), key=lambda i: (located[i][0], i))
        jobs = [jobs[i] for i in order]
        located = [located[i] for i in order]
        parts, g0 = [], 0
        for g in range(1, G + 1):
            if g == G or located[g][0] != located[g0][0]:
                src = located[g0][0]
                parts.append((bases[src], 3 * self.S if src < self.world else self._mirror_cap, g0, g)); g0 = g
        fe.score_parts = parts if (ar.off_live == 0 and self.opt.score_tma) else None
        fe.wslot[:G] = self._ints([ps for _, ps in located], torch.int32)
        valid = []
        for vi, idx in samples.items():
            k = int(idx.numel())
            st["perm"][vi, :k] = idx
            if k < rows:
                st["perm"][vi, k:] = idx[-1]
        for _, vi, _ in jobs:
            valid.append(int(samples[vi].numel()))
        fe.gmap[:G] = self._ints([vi for _, vi, _ in jobs], torch.int32)
        st["row_tab"][:G] = self._ints([a for _, _, a in jobs])
        fe.load(G, st["x_tab"], st["y_tab"], st["perm"], 0)
        fe.forward(G)
        tmp = st["stats"][:G]
        tmp.zero_()
        fe.metrics(G, fe.eval_descriptors(valid), tmp, dirichlet=(kind == "dirichlet"))
        stats.index_copy_(0, self._ints([j[0] for j in jobs]), tmp)
        self.kernel_launches += len(fe.ops) + 2
        return True

    # ---- UBAR --------------------------------------------------------------------------------
    def _ubar_prepare(self, et, parity: int) -> None:
        """Stage 1 (distances → shortlist) + an ASYNC copy of the shortlist to pinned memory; the host only waits for it
        in ``_agg_ubar``, after it has enqueued whatever else the round needs (e.g. the DMTT scoring)."""
        a = self.aggregator
        self._edge_dist(et, parity, self.layout.stride)
        self.ext.ubar_stage1(*self._et_args(et), et["d2"], a.rho, a.min_neighbors, et["aux"], et["aux2"], self._sync_args()[4])
        self.kernel_launches += 1
        E = et["aux"].numel()
        if getattr(self, "_cand_pinned", None) is None or self._cand_pinned.numel() < E:
            self._cand_pinned = torch.zeros(max(E, 1024)).pin_memory()
            self._cand_event = torch.cuda.Event()
        self._cand_pinned[:E].copy_(et["aux"], non_blocking=True)
        self._cand_event.record()
        et["ubar_prepared"] = self.epoch

    def _agg_ubar(self, et, parity: int) -> None:
        a = self.aggregator
        if et.get("ubar_prepared") != self.epoch:
            self._ubar_prepare(et, parity)
        tp = self._sync_args()[4]
        cand, rank_t, loss, own_loss = et["aux"], et["aux2"], et["aux3"], et["n2"]
        rows, rk, sl = et["host_rows"], et["host_rank"], et["host_slot"]
        stats = torch.zeros(max(len(rk), 1), _STAT_COLS, device=self.device)
        # fused path: score EVERY neighbour in one grouped forward program and let ubar_stage2 mask with the device-side
        # shortlist — no host round-trip between stage 1 and stage 2 (reference aggregation/ubar.py:152-202)
        jobs, samples = [], {}
        for vi, vn in enumerate(self.nodes):
            if rows[vi + 1] - rows[vi] <= 1 or vn.n == 0:
                continue
            samples[vi] = torch.randperm(vn.n, device=self.device)[: vn.eb]
            jobs.append((rows[vi], vi, self._row_ptr(self.rank, None, vn.slot)))
            jobs += [(e, vi, self._row_ptr(rk[e], parity, sl[e])) for e in range(rows[vi] + 1, rows[vi + 1])]
        eb_max = max([vn.eb for vn in self.nodes] or [1])
        if self._fused_scores(jobs, samples, eb_max, "ce", stats):
            mean_loss = stats[:, 0] / stats[:, 2].clamp_min(1.0)
            loss.copy_(mean_loss[: loss.numel()])
            if self.V:
                own_loss[: self.V] = mean_loss[self._ints(rows[: self.V])]
            self.ext.ubar_stage2(*self._et_args(et), cand, rank_t, loss, own_loss, a.alpha, True)
            self.kernel_launches += 1
            self._log_stats(et)
            return self._gather(et, parity, renorm=False, sync=False)
        self._cand_event.synchronize()                               # tiny D2H issued earlier: which candidates to evaluate
        cand_host = self._cand_pinned[: cand.numel()]
        if self._mlp_plan is not None:
            jobs, inputs = [], {}
            for vi, vn in enumerate(self.nodes):
                if rows[vi + 1] - rows[vi] <= 1 or vn.n == 0:
                    continue
                inputs[vi] = self._flat_inputs(vn, torch.randperm(vn.n, device=self.device)[: vn.eb], "ubar")
                jobs.append((rows[vi], vi, self._row_ptr(self.rank, None, vn.slot)))           # own loss from the live row
                jobs += [(e, vi, self._row_ptr(rk[e], parity, sl[e])) for e in range(rows[vi] + 1, rows[vi + 1]) if cand_host[e] != 0]
            self._grouped_mlp_scores(jobs, inputs, "ce", stats)
        self._fork()
        for vi, vn in enumerate(self.nodes):
            if rows[vi + 1] - rows[vi] <= 1 or vn.n == 0 or self._mlp_plan is not None:
                continue
            with torch.cuda.stream(self.streams[self.stream_of[vi]]):
                ev = self._evaluator(vn, vn.eb, "ce")            # CrossEntropyLoss on raw outputs even for evidential models
                pick = torch.randperm(vn.n, device=self.device)[: vn.eb]
                ev.load_inputs(vn.X.index_select(0, pick), vn.y.index_select(0, pick))
                ev.run(self.live[vn.slot], stats[rows[vi]])       # own loss (self edge slot)
                for e in range(rows[vi] + 1, rows[vi + 1]):
                    if cand_host[e] != 0:
                        ev.run(self._src_row(rk[e], parity, sl[e]), stats[e])
        self._join()
        mean_loss = stats[:, 0] / stats[:, 2].clamp_min(1.0)
        loss.copy_(mean_loss[: loss.numel()])
        own_loss[: self.V] = mean_loss[self._ints(rows[: self.V])] if self.V else own_loss[: self.V]
        self.ext.ubar_stage2(*self._et_args(et), cand, rank_t, loss, own_loss, a.alpha, True)
        self.kernel_launches += 1
        self._log_stats(et)
        self._gather(et, parity, renorm=False, sync=False)

    # ---- EvidentialTrust ----------------------------------------------------------------------
    def _agg_evidential_trust(self, et, parity: int) -> None:
        a = self.aggregator
        if self.world > 1:
            self._host_wait_epoch()
        if "ema" not in self._agg_state:
            self._agg_state["ema"] = torch.zeros(max(self.V, 1), self.N, device=self.device)
            self._agg_state["ema_valid"] = torch.zeros(max(self.V, 1), self.N, device=self.device)
        vac, acc, trust = et["aux"], et["aux2"], et["aux3"]
        rows, rk, sl = et["host_rows"], et["host_rank"], et["host_slot"]
        stats = torch.zeros(max(len(rk), 1), _STAT_COLS, device=self.device)
        jobs, samples, take_max = [], {}, 1
        for vi, vn in enumerate(self.nodes):
            if vn.n == 0:
                continue
            nbatch = max(1, math.ceil(a.max_eval_samples / max(vn.eb, 1)))
            take_n = min(vn.n, nbatch * vn.eb)
            if vn.n > vn.eb:
                take_n = (take_n // vn.eb) * vn.eb               # the reference iterates a drop_last loader: whole batches only
            samples[vi] = torch.randperm(vn.n, device=self.device)[:take_n]
            take_max = max(take_max, take_n)
            jobs += [(e, vi, self._row_ptr(rk[e], parity, sl[e])) for e in range(rows[vi] + 1, rows[vi + 1])]
        fused_done = self._fused_scores(jobs, samples, take_max, "dirichlet", stats)
        if fused_done:
            pass
        elif self._mlp_plan is not None:
            jobs, inputs = [], {}
            for vi, vn in enumerate(self.nodes):
                if vn.n == 0:
                    continue
                nbatch = max(1, math.ceil(a.max_eval_samples / max(vn.eb, 1)))
                take_n = min(vn.n, nbatch * vn.eb)
                inputs[vi] = self._flat_inputs(vn, torch.randperm(vn.n, device=self.device)[:take_n], "et")
                jobs += [(e, vi, self._row_ptr(rk[e], parity, sl[e])) for e in range(rows[vi] + 1, rows[vi + 1])]
            self._grouped_mlp_scores(jobs, inputs, "dirichlet", stats)
        self._fork()
        for vi, vn in enumerate(self.nodes):
            if vn.n == 0 or self._mlp_plan is not None or fused_done:
                continue
            with torch.cuda.stream(self.streams[self.stream_of[vi]]):
                nbatch = max(1, math.ceil(a.max_eval_samples / max(vn.eb, 1)))
                take_n = min(vn.n, nbatch * vn.eb)
                ev = self._evaluator(vn, take_n, "dirichlet")
                take = torch.randperm(vn.n, device=self.device)[:take_n]
                ev.load_inputs(vn.X.index_select(0, take), vn.y.index_select(0, take))
                for e in range(rows[vi] + 1, rows[vi + 1]):
                    ev.run(self._src_row(rk[e], parity, sl[e]), stats[e])
        self._join()
        cnt = stats[:, 2].clamp_min(1.0)
        vac.copy_((stats[:, 3] / cnt)[: vac.numel()]); acc.copy_((stats[:, 1] / cnt)[: acc.numel()])
        self.ext.trust_filter(*self._et_args(et), vac, acc, et["src_gid"], self.N, self._agg_state["ema"],
                              self._agg_state["ema_valid"], a.accuracy_weight, a.vacuity_threshold, a.trust_momentum,
                              bool(a.use_adaptive_trust), a.current_threshold(self.round_idx), a.self_weight, trust,
                              self._sync_args()[4])
        self.kernel_launches += 2
        self._log_stats(et)
        self._gather(et, parity, renorm=False, sync=False)

    # ---- generic (user-defined Aggregator subclasses) -------------------------------------------
    def _agg_generic(self, et, parity: int) -> None:
        if self.world > 1:
            self._host_wait_epoch()
            torch.cuda.synchronize()
        rows, rk, sl, gids = et["host_rows"], et["host_rank"], et["host_slot"], et["host_gid"]
        results = []
        for vi, vn in enumerate(self.nodes):
            own = {k: v.clone() for k, v in self.layout.row_views(self.live[vn.slot], self.ints[vn.slot]).items()}
            nbrs = {}
            for e in range(rows[vi] + 1, rows[vi + 1]):
                row = self.arena.peer_row(rk[e], parity, sl[e])
                st = {k: v.clone() for k, v in self.layout.row_views(row, None).items()}
                for en in self.layout.int_entries():
                    st[en.name] = row[self.layout.Pf_pad + en.offset: self.layout.Pf_pad + en.offset + en.numel].round().long().view(en.shape)
                nbrs[gids[e]] = st
            results.append(self.aggregator.aggregate(node_id=vn.gid, own_state=own, neighbor_states=nbrs,
                                                     round_num=self.round_idx, model_template=vn.model, device=self.device,
                                                     train_loader=None))
        for vn, st in zip(self.nodes, results):
            views = self.layout.row_views(self.live[vn.slot], self.ints[vn.slot])
            for k, v in st.items():
                views[k].copy_(v.to(views[k].dtype))

    # ---- NCCL + stock-PyTorch baseline (b200.transport: nccl) -----------------------------------------
    def _aggregate_nccl(self, neighbors: List[List[int]]) -> None:
        """The comparison baseline named in BASELINE.json (``parallel/nccl_baseline.py``): same placement and training, flat-row
        exchange over NCCL (one ``all_reduce`` on a full mesh, ``batch_isend_irecv`` along the edge list otherwise) and
        vectorised stock-PyTorch aggregation.  None of the fused exchange / aggregation kernels run on this path."""
        if not hasattr(self, "_nccl"):
            from murmura_b200.parallel.nccl_baseline import NcclBaseline
            self._nccl = NcclBaseline(self)
        self._nccl.aggregate(neighbors)

    def _aggregate(self, parity: int) -> None:
        neighbors, key = self._neighbors_for_round(self.round_idx)
        if self.opt.transport == "nccl":
            if self.dmtt_on:                       # trust bookkeeping is shared; only exchange+aggregation differ
                et = self._edge_table(neighbors, key)
                if self.world > 1:
                    self.arena.timed_out.zero_()
                    self._liveness_frozen = False
                self._publish(parity)
                self._freeze_liveness()
                self._dmtt_score_and_update(et, parity)
            return self._aggregate_nccl(neighbors)
        et = self._edge_table(neighbors, key)
        if self.opt.fault_drop_edges:
            self._apply_fault_mask(et, self.round_idx)
        if self.world > 1:
            self.arena.timed_out.zero_()
            self._liveness_frozen = False
        et["rank_sum"] = self._fullmesh_fedavg(et)
        self._publish(parity, with_sum=et["rank_sum"])
        self._freeze_liveness()
        if self.family == "ubar":
            self._ubar_prepare(et, parity)
        if self.dmtt_on:
            self._dmtt_score_and_update(et, parity)
        plan = getattr(self, f"_agg_{self.family}", self._agg_generic)
        self._last_et = et
        self._account_traffic(et)
        plan(et, parity)

    _FILTER_ROW_PASSES = {"balance": 1.0, "ubar": 1.0}        # families whose filter streams every neighbour row once more

    def _account_traffic(self, et) -> None:
        """Algorithmic bytes of this round's exchange+aggregate on this GPU, assuming every edge is accepted
        (publish: read+write V rows; filter pass; gather: one read per edge + one write per node). Feeds ``perf_summary``."""
        rk = et["host_rank"]
        edges = len(rk)
        remote = sum(1 for r in rk if r != self.rank)
        row = self.layout.Pf_pad * 4
        if et.get("rank_sum"):
            # publish_sum: V rows read, V + 1 written; fedavg_fullmesh: own sum row read, V rows written; fabric: one row per peer
            # (or ONE multicast row when the NVSwitch reduces)
            nvls = self.opt.transport == "nvls" and self.arena.mc_base
            G = self.world
            if self._two_shot():
                # inbound per GPU: the (G−1)/G remote parts of its own slice (or the switch-reduced slice) + the G−1 slices the peers scatter
                link = row * ((1.0 / G + (G - 1.0) / G) if nvls else 2.0 * (G - 1.0) / G)
                hbm = row * (3 * self.V + 4)
            else:
                link = row * (0 if G == 1 else (1 if nvls else G - 1))
                hbm = row * (3 * self.V + 2)
            self.timers["hbm_bytes"] = self.timers.get("hbm_bytes", 0.0) + hbm
            self.timers["nvlink_bytes"] = self.timers.get("nvlink_bytes", 0.0) + link
            return
        passes = 1.0 + self._FILTER_ROW_PASSES.get(self.family, 0.0)
        extra = self.V if self.family == "sketchguard" else (self.N / max(self.world, 1) if self.family == "krum" else 0.0)
        self.timers["hbm_bytes"] = self.timers.get("hbm_bytes", 0.0) + row * (3 * self.V + passes * (edges - remote) + extra)
        self.timers["nvlink_bytes"] = self.timers.get("nvlink_bytes", 0.0) + row * passes * remote

    # =========================================================================================
    # evaluation
    # =========================================================================================
    def _eval_node(self, vn: VirtualNode) -> None:
        stats = self.eval_stats[vn.slot]
        stats.zero_()
        vn.model.eval()
        EB = max(1, self.opt.eval_batch)
        from murmura_b200.ops import fast_eval_batchnorm
        with torch.no_grad(), fast_eval_batchnorm():
            for a in range(0, vn.n, EB):
                xb, yb = self._inputs(vn, vn.X[a:a + EB]), vn.y[a:a + EB]
                with self._autocast():
                    out = vn.model(xb).float().contiguous()
                if self.evidential:
                    self.ext.dirichlet_eval(out, yb, None, stats)
                else:
                    self.ext.ce_eval(out, yb, None, stats)

    # ---- fused evaluation: every node's forward in ONE program (conv epilogues carry eval-BN / residual / ReLU) -----------------
    def _fused_eval_setup(self):
        from murmura_b200.parallel.fused_trainer import FusedForward
        if not self.opt.fused_train or not self.nodes or self.opt.compute_dtype == "bf16" or not self.opt.cuda_graphs:
            return None
        first = self.nodes[0]
        if (first.X.dim() == 4 and not first.nhwc) or any(vn.n == 0 for vn in self.nodes):
            return None
        shape = (first.X.shape[3], first.X.shape[1], first.X.shape[2]) if first.X.dim() == 4 else tuple(first.X.shape[1:])
        n_max = max(vn.n for vn in self.nodes)
        # rows per group and launch: ≥ fused_eval_rows, and ≈ 2048 rows per launch over all groups so that a GPU hosting one or
        # two nodes (launch-latency bound) evaluates its shard in one or two passes instead of many small ones
        EB = max(32, int(self.opt.fused_eval_rows), (2048 // max(self.V, 1) + 31) // 32 * 32)
        EB = min(EB, (n_max + 31) // 32 * 32)
        fe = FusedForward(first.model, self.layout, self.live, EB, shape, self.V, evidential=self.evidential)
        if not fe.supported or (self.evidential and not fe.evidential_head):
            return None
        order = sorted(range(self.V), key=lambda i: (-self.nodes[i].n, i))
        chunks = (n_max + EB - 1) // EB
        dev = self.device
        fe.gmap = torch.tensor(order, dtype=torch.int32, device=dev)
        perm = torch.zeros(self.V, chunks * EB, dtype=torch.int64)
        for vn in self.nodes:
            perm[vn.slot] = torch.arange(chunks * EB).clamp_max(vn.n - 1)
        st = {"fe": fe, "order": torch.tensor(order, dtype=torch.int64, device=dev), "perm": perm.to(dev), "chunks": [],
              "x_tab": torch.tensor([vn.X.data_ptr() for vn in self.nodes], dtype=torch.int64, device=dev),
              "y_tab": torch.tensor([vn.y.data_ptr() for vn in self.nodes], dtype=torch.int64, device=dev), "graph": None}
        for c in range(chunks):
            valid = [min(EB, max(0, self.nodes[i].n - c * EB)) for i in order]
            G = sum(1 for v in valid if v > 0)
            st["chunks"].append((G, fe.eval_descriptors(valid[:G])))
        return st

    def _fused_eval_body(self, st) -> None:
        fe = st["fe"]
        fe.stats.zero_()
        for c, (G, desc) in enumerate(st["chunks"]):
            fe.load(G, st["x_tab"], st["y_tab"], st["perm"], c)
            fe.forward(G)
            fe.metrics(G, desc)
        self.eval_stats.index_copy_(0, st["order"], fe.stats[: self.V])

    def _evaluate_fused(self) -> bool:
        if "_fused_eval" not in self.__dict__:
            self._fused_eval = self._fused_eval_setup()
        st = self._fused_eval
        if st is None:
            return False
        fe = st["fe"]
        if st["graph"] is None:
            side = self.capture_streams[0]
            side.wait_stream(torch.cuda.current_stream())
            before = fe.be.launches
            with torch.cuda.stream(side):
                self._fused_eval_body(st)
            st["launches"] = fe.be.launches - before
            torch.cuda.current_stream().wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                self._fused_eval_body(st)
            st["graph"] = g
        st["graph"].replay()
        self.kernel_launches += st["launches"]
        return True

    def _evaluate(self, enqueue_only: bool = False) -> Optional[List[Dict[str, Any]]]:
        if self.opt.seed_parity:
            for _ in range(self.N):                     # the simulation's N evaluation loaders each draw their iterator's base seed
                torch.empty((), dtype=torch.int64).random_()
        if self._evaluate_fused():
            return None if enqueue_only else self._collect_metrics()
        self._fork()
        for i in self.launch_order:
            vn = self.nodes[i]
            stream = self.streams[self.stream_of[i]]
            with torch.cuda.stream(stream):
                if self.opt.cuda_graphs:
                    if vn.eval_graph is None:
                        self._capture_eval(vn)
                    vn.eval_graph.replay()
                else:
                    self._eval_node(vn)
                self.kernel_launches += (vn.n + max(1, self.opt.eval_batch) - 1) // max(1, self.opt.eval_batch)
        self._join()
        return None if enqueue_only else self._collect_metrics()

    @property
    def d2h_bytes_per_round(self) -> int:
        """Bytes of the metric table an evaluated round copies to the host (pinned ring slot + the timed-out mask)."""
        return int(self.placement.slots_per_rank * self.world * _STAT_COLS * 4 + 4)

    # ---- metrics ring: evaluated rounds leave their [N, 8] table in a pinned host ring; the host reads it LATER ----
    _RING = 64

    def _metrics_enqueue(self) -> Tuple[int, "torch.cuda.Event"]:
        """Stream-ordered: gather every rank's metric rows and copy them (non-blocking) into the next ring slot.  No host sync —
        the nodes' critical path never waits for the monitor (reference ``distributed/monitor.py:81-128`` is equally passive)."""
        S = self.placement.slots_per_rank
        if not hasattr(self, "_ring"):
            self._ring = torch.zeros(self._RING, S * self.world, _STAT_COLS).pin_memory()
            self._ring_to = torch.zeros(self._RING, dtype=torch.int32).pin_memory()
            self._ring_next = 0
        local = torch.zeros(S, _STAT_COLS, device=self.device)
        local[: self.V] = self.eval_stats[: self.V]
        if self.world > 1:
            full = torch.zeros(S * self.world, _STAT_COLS, device=self.device)
            _dist().all_gather_into_tensor(full, local)
        else:
            full = local
        slot = self._ring_next % self._RING
        self._ring_next += 1
        self._ring[slot].copy_(full, non_blocking=True)
        if self.world > 1 and hasattr(self, "_timeout_acc"):
            self._ring_to[slot: slot + 1].copy_(self._timeout_acc, non_blocking=True)
            self._timeout_acc.zero_()
        ev = torch.cuda.Event()
        ev.record()
        return slot, ev

    def _metrics_from_slot(self, slot: int, round_no: int) -> List[Dict[str, Any]]:
        S = self.placement.slots_per_rank
        if self.world > 1 and int(self._ring_to[slot]):
            mask = int(self._ring_to[slot]); self._ring_to[slot] = 0
            late = [r for r in range(self.world) if mask >> r & 1]
            self.timeout_events = getattr(self, "timeout_events", []) + [(round_no, late)]
            print(f"[b200 rank {self.rank}] round {round_no}: rank(s) {late} did not publish within "
                  f"{self.opt.flag_timeout_ms:.0f} ms; their nodes were aggregated as missing neighbours")
        host = self._ring[slot].numpy()
        per_node = []
        for gid in range(self.N):
            r, s = int(self.placement.rank_of[gid]), int(self.placement.slot_of[gid])
            row = host[r * S + s]
            total = max(float(row[2]), 1.0)
            m = {"node_id": gid, "accuracy": float(row[1]) / total, "loss": float(row[0]) / total,
                 "correct": int(row[1]), "total": int(row[2])}
            if self.evidential:
                m.update(vacuity=float(row[3]) / total, entropy=float(row[4]) / total, strength=float(row[5]) / total)
            per_node.append(m)
        return per_node

    def _collect_metrics(self) -> List[Dict[str, Any]]:
        """Synchronous form (direct ``_evaluate()`` callers): enqueue, wait for that slot, decode."""
        slot, ev = self._metrics_enqueue()
        ev.synchronize()
        return self._metrics_from_slot(slot, self.round_idx + 1)

    def _drain_metrics(self, block: bool, verbose: bool = False) -> None:
        """Move finished ring slots into ``history`` in round order; ``block`` waits for all of them (end of ``train``, a full
        ring, verbose printing, checkpoints)."""
        pend = self.__dict__.setdefault("_pending_metrics", [])
        while pend:
            round_no, slot, ev = pend[0]
            if not block and not ev.query():
                break
            ev.synchronize()
            record_round(self.history, round_no, self._metrics_from_slot(slot, round_no), self.compromised if self.attack else None, verbose)
            pend.pop(0)

    def _capture_eval(self, vn: VirtualNode) -> None:
        side = self.capture_streams[self.stream_of[vn.slot]]
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            self._eval_node(vn)
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            self._eval_node(vn)
        vn.eval_graph = g

    # =========================================================================================
    # DMTT (dynamic topology + trust) on the device
    # =========================================================================================
    def _dmtt_init(self) -> None:
        self.dmtt_on = self.cfg.dmtt is not None and self.mobility is not None
        if self.mobility is None:
            return
        R = max(int(self.cfg.experiment.rounds), 1)
        self.pos_dev = torch.from_numpy(self.mobility.positions_tensor(R + 1)).to(self.device)
        self.adj_dev = torch.zeros(self.N, self.N, dtype=torch.uint8, device=self.device)
        if not self.dmtt_on:
            return
        N, V = self.N, max(self.V, 1)
        liar = torch.zeros(N, dtype=torch.uint8)
        if self.attack is not None and hasattr(self.attack, "get_false_claims"):
            for i in self.compromised:
                liar[i] = 1
        self.is_liar = liar.to(self.device)
        self.claims_dev = torch.zeros(N, N, dtype=torch.uint8, device=self.device)
        self.collab = None                                    # [N,N] uint8, C^{t-1}
        self.c_hat = torch.full((V, N), 0.5, device=self.device)
        self.t_alpha = torch.ones(V, N, device=self.device); self.t_beta = torch.ones(V, N, device=self.device)
        self.next_collab = torch.zeros(V, N, dtype=torch.uint8, device=self.device)
        self.q_out = torch.zeros(V, N, device=self.device)

    def _adjacency(self, r: int) -> torch.Tensor:
        m = self.cfg.mobility
        self.ext.mobility_adjacency(self.pos_dev, min(r, self.pos_dev.shape[0] - 1), m.area_size, m.comm_range,
                                    bool(m.ensure_connected), self.adj_dev)
        self.kernel_launches += 1
        return self.adj_dev

    def _dmtt_neighbors(self, r: int) -> List[List[int]]:
        adj = self._adjacency(r)
        self.ext.liar_claims(adj, self.is_liar, self.claims_dev)
        if self.collab is None:                              # round 0: C_i = G^0 neighbours
            self.collab = adj.clone()
            self._stage_collab()
        self._collab_event.synchronize()                     # N×N bytes, copied asynchronously when C^t was produced
        c = self._collab_pinned.numpy().astype(bool).copy()  # next round's directed edge list
        self._received = c & c.T                             # j's state reaches i iff i∈C_j and j∈C_i (i only accepts expected senders)
        return [np.flatnonzero(self._received[i]).tolist() for i in range(self.N)]

    def _dmtt_score_and_update(self, et, parity: int) -> None:
        d, N = self.cfg.dmtt, self.N
        if self.world > 1:
            self._host_wait_epoch()
        V = max(self.V, 1)
        score = torch.zeros(V, N, device=self.device); valid = torch.zeros(V, N, dtype=torch.uint8, device=self.device)
        received = torch.zeros(V, N, dtype=torch.uint8, device=self.device)
        rows, rk, sl, gids = et["host_rows"], et["host_rank"], et["host_slot"], et["host_gid"]
        stats = torch.zeros(max(len(rk), 1), _STAT_COLS, device=self.device)
        jobs, samples = [], {}
        for vi, vn in enumerate(self.nodes):
            if vn.n == 0 or rows[vi + 1] - rows[vi] <= 1:
                continue
            samples[vi] = self._arange(vn.n)                         # the whole local test set (= training shard), reference dmtt/node_process.py:309-363
            jobs += [(e, vi, self._row_ptr(rk[e], parity, sl[e])) for e in range(rows[vi] + 1, rows[vi + 1])]
        n_max = max([vn.n for vn in self.nodes] or [1])
        fused_done = self._fused_scores(jobs, samples, n_max, "dirichlet" if self.evidential else "ce", stats)
        if fused_done:
            pass
        elif self._mlp_plan is not None:
            jobs, inputs = [], {}
            for vi, vn in enumerate(self.nodes):
                if vn.n == 0 or rows[vi + 1] - rows[vi] <= 1:
                    continue
                inputs[vi] = self._flat_inputs(vn, None, "dmtt")
                jobs += [(e, vi, self._row_ptr(rk[e], parity, sl[e])) for e in range(rows[vi] + 1, rows[vi + 1])]
            self._grouped_mlp_scores(jobs, inputs, "dirichlet" if self.evidential else "ce", stats)
        self._fork()
        for vi, vn in enumerate(self.nodes):
            if vn.n == 0 or rows[vi + 1] - rows[vi] <= 1 or self._mlp_plan is not None or fused_done:
                continue
            with torch.cuda.stream(self.streams[self.stream_of[vi]]):
                ev = self._evaluator(vn, vn.n, "dirichlet" if self.evidential else "ce")
                if not getattr(ev, "_full_loaded", False):
                    ev.load_inputs(vn.X, vn.y); ev._full_loaded = not bool(self._host_shards)
                for e in range(rows[vi] + 1, rows[vi + 1]):
                    ev.run(self._src_row(rk[e], parity, sl[e]), stats[e])
        self._join()
        cnt = stats[:, 2].clamp_min(1.0)
        acc_e = stats[:, 1] / cnt
        u_e = stats[:, 3] / cnt if self.evidential else torch.zeros_like(acc_e)
        s_e = (1.0 - u_e) * (d.w_a * acc_e + (1.0 - d.w_a))
        s_e = torch.where(u_e > d.tau_u, s_e * torch.exp(-(u_e - d.tau_u)), s_e).clamp_min(0.0)
        ev_i = [vi for vi in range(self.V) for _ in range(rows[vi] + 1, rows[vi + 1])]
        ee_i = [e for vi in range(self.V) for e in range(rows[vi] + 1, rows[vi + 1])]
        if ee_i:
            vi_t = self._ints(ev_i); e_t = self._ints(ee_i)
            g_t = et["src_gid"].long().index_select(0, e_t)
            score[vi_t, g_t] = s_e.index_select(0, e_t); valid[vi_t, g_t] = 1; received[vi_t, g_t] = 1
        self.ext.dmtt_update(self.adj_dev, self.claims_dev, self.collab, received, score, valid, self.c_hat, self.t_alpha,
                             self.t_beta, self.next_collab, self.q_out, d.rho, d.lambda_forget, d.w_d, d.w_x, d.tau_U, d.eta,
                             d.lambda1, d.lambda2, d.lambda3, int(d.budget_B), self.node_gid)
        self.kernel_launches += 1
        S = self.placement.slots_per_rank
        local = torch.zeros(S, N, dtype=torch.uint8, device=self.device)
        local[: self.V] = self.next_collab[: self.V]
        if self.world > 1:
            full = torch.zeros(S * self.world, N, dtype=torch.uint8, device=self.device)
            _dist().all_gather_into_tensor(full, local)
        else:
            full = local
        rows_idx = torch.tensor([int(self.placement.rank_of[g]) * S + int(self.placement.slot_of[g]) for g in range(N)],
                                device=self.device)
        self.collab = full.index_select(0, rows_idx).contiguous()
        self._stage_collab()

    def _stage_collab(self) -> None:
        if getattr(self, "_collab_pinned", None) is None:
            self._collab_pinned = torch.zeros(self.N, self.N, dtype=torch.uint8).pin_memory()
            self._collab_event = torch.cuda.Event()
        self._collab_pinned.copy_(self.collab, non_blocking=True)
        self._collab_event.record()

    # =========================================================================================
    # round loop
    # =========================================================================================
    def train(self, rounds: int, local_epochs: int = 1, lr: float = 0.01, verbose: bool = False,
              eval_every: int = 1) -> Dict[str, List[Any]]:
        verbose = verbose and self.is_primary
        seen = self.__dict__.setdefault("_train_keys", set())
        if (local_epochs, lr, eval_every) not in seen:            # new graphs / workspaces get allocated in the next rounds
            seen.add((local_epochs, lr, eval_every))
            self._warm_rounds = max(getattr(self, "_warm_rounds", 0), int(self.opt.host_barrier_rounds))
        self._prepare_training(local_epochs, lr)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        total_rounds = self.round_idx + rounds
        for _ in range(rounds):
            r = self.round_idx
            if verbose:
                print(f"\n=== Round {r + 1}/{total_rounds} ===")
            prof = self.opt.profile
            nvtx = torch.cuda.nvtx
            if prof:
                ev[0].record(); nvtx.range_push(f"round {r + 1}"); nvtx.range_push("local_training")
            self._local_training(local_epochs, lr)
            if prof:
                ev[1].record(); nvtx.range_pop(); nvtx.range_push("exchange+aggregate")
            self._aggregate(parity=r & 1)
            if prof:
                ev[2].record(); nvtx.range_pop(); nvtx.range_push("evaluate")
            if (r + 1) % eval_every == 0:
                # evaluation is enqueued; its metric table lands in the pinned ring and is folded into ``history`` lazily, so the
                # host keeps enqueueing the next round while the GPU is still busy (verbose runs print round by round: blocking)
                self._evaluate(enqueue_only=True)
                pend = self.__dict__.setdefault("_pending_metrics", [])
                pend.append((r + 1, *self._metrics_enqueue()))
                lazy = self.opt.lazy_metrics and not verbose and not prof
                self._drain_metrics(block=not lazy or len(pend) >= self._RING - 2, verbose=verbose)
            if prof:
                nvtx.range_pop(); nvtx.range_pop()
                ev[3].record(); torch.cuda.synchronize()
                self.timers["train_ms"] += ev[0].elapsed_time(ev[1]); self.timers["aggregate_ms"] += ev[1].elapsed_time(ev[2])
                self.timers["eval_ms"] += ev[2].elapsed_time(ev[3])
            self.timers["rounds"] += 1
            self._warm_rounds = max(0, getattr(self, "_warm_rounds", 0) - 1)
            self.round_idx += 1
            if self.opt.checkpoint_every and self.round_idx % self.opt.checkpoint_every == 0:
                self._drain_metrics(block=True)
                self.save_checkpoint(os.path.join(self.opt.checkpoint_dir, f"round_{self.round_idx:05d}"))
        self._drain_metrics(block=True)
        return self.history

    # =========================================================================================
    # statistics / checkpoint / teardown
    # =========================================================================================
    def reset_timers(self) -> None:
        """Forget the per-phase timings collected so far (benchmarks call this after their warm-up rounds)."""
        for k in ("train_ms", "aggregate_ms", "eval_ms", "rounds", "hbm_bytes", "nvlink_bytes"):
            if k in self.timers:
                self.timers[k] = 0.0

    def get_node_statistics(self) -> Dict[int, Dict[str, Any]]:
        out: Dict[int, Dict[str, Any]] = {}
        log = torch.stack(self._stat_log).cpu().numpy() if self._stat_log else np.zeros((0, max(self.V, 1), 4))
        for vi, vn in enumerate(self.nodes):
            st: Dict[str, Any] = {"total_rounds_processed": int(log.shape[0])}
            if log.shape[0]:
                acc, ev_ = log[:, vi, 0], np.maximum(log[:, vi, 1], 1.0)
                if self.family in ("balance", "sketchguard", "evidential_trust"):
                    st["mean_acceptance_rate"] = float(np.mean(acc / ev_)); st["current_threshold"] = float(log[-1, vi, 2])
                if self.family == "ubar":
                    st["stage1_mean_acceptance_rate"] = float(np.mean(acc / ev_))
                    st["stage2_mean_acceptance_rate"] = float(np.mean(log[:, vi, 2] / np.maximum(log[:, vi, 3], 1.0)))
                if self.family == "krum":
                    st["winner_history"] = log[:, vi, 0].astype(int).tolist()
            for k in ("train_ms", "aggregate_ms", "eval_ms"):
                st[k] = self.timers[k]
            out[vn.gid] = st
        return out

    def perf_summary(self) -> str:
        t = self.timers
        if not t["rounds"] or not self.opt.profile:
            return f"{int(t['rounds'])} rounds on {self.world} GPU(s); set b200.profile: true for the per-phase split"
        r = t["rounds"]
        line = (f"train {t['train_ms'] / r:.2f} ms  aggregate {t['aggregate_ms'] / r:.3f} ms  eval {t['eval_ms'] / r:.2f} ms "
                f"per round ({self.world} GPU(s))")
        if t["aggregate_ms"] > 0 and t.get("hbm_bytes"):
            peak = _measured_hbm_gbs()
            hbm = t["hbm_bytes"] / t["aggregate_ms"] / 1e6
            line += f"; exchange+aggregate ≈ {hbm:.0f} GB/s HBM ({hbm / peak:.2f} of {peak:.0f} GB/s measured)"
            if t.get("nvlink_bytes"):
                line += f" + {t['nvlink_bytes'] / t['aggregate_ms'] / 1e6:.0f} GB/s over NVLink"
            line += " [algorithmic bytes, all edges accepted; includes flag waits and filter kernels]"
        return line

    def state_dict_of(self, gid: int) -> Dict[str, torch.Tensor]:
        vn = next(v for v in self.nodes if v.gid == gid)
        return {k: v.detach().clone() for k, v in self.layout.row_views(self.live[vn.slot], self.ints[vn.slot]).items()}

    def save_checkpoint(self, path: str) -> None:
        """Flat-arena checkpoint: live plane, int table, RNG, trust state, round index."""
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
        blob = {"round_idx": self.round_idx, "epoch": self.epoch, "live": self.live[: self.V].cpu(), "ints": self.ints.cpu(),
                "rng": torch.cuda.get_rng_state(self.device), "history": self.history, "local_gids": self.local_gids,
                "agg_state": {k: v.cpu() for k, v in self._agg_state.items()}}
        if self.dmtt_on:
            blob["dmtt"] = {"c_hat": self.c_hat.cpu(), "alpha": self.t_alpha.cpu(), "beta": self.t_beta.cpu(),
                            "collab": None if self.collab is None else self.collab.cpu()}
        if self.family == "sketchguard":
            blob["sk_hist"] = self.sk_hist.cpu()
        torch.save(blob, f"{path}.rank{self.rank}.pt")

    def load_checkpoint(self, path: str) -> None:
        blob = torch.load(f"{path}.rank{self.rank}.pt", map_location="cpu", weights_only=False)
        if blob["local_gids"] != self.local_gids:
            raise ValueError("checkpoint was written with a different node placement")
        self.round_idx, self.history = blob["round_idx"], blob["history"]
        self.live[: self.V].copy_(blob["live"]); self.ints.copy_(blob["ints"])
        torch.cuda.set_rng_state(blob["rng"], self.device)
        for k, v in blob.get("agg_state", {}).items():
            self._agg_state[k] = v.to(self.device)
        if self.dmtt_on and "dmtt" in blob:
            d = blob["dmtt"]
            self.c_hat.copy_(d["c_hat"]); self.t_alpha.copy_(d["alpha"]); self.t_beta.copy_(d["beta"])
            self.collab = None if d["collab"] is None else d["collab"].to(self.device)
            if self.collab is not None:
                self._stage_collab()
        if self.family == "sketchguard" and "sk_hist" in blob:
            self.sk_hist.copy_(blob["sk_hist"])
        if self.world > 1:      # epochs must stay monotone across the job; re-align on the max
            t = torch.tensor([max(self.epoch, blob["epoch"])], device=self.device)
            _dist().all_reduce(t, op=_dist().ReduceOp.MAX)
            self.epoch = int(t.item())

    def close(self) -> None:
        torch.cuda.synchronize()
        for vn in self.nodes:
            vn.train_graph = None; vn.eval_graph = None; vn.split_bwd = None
        self._evaluators.clear()                      # CUDA graphs / views that point into the arena must go before it does
        self.__dict__.pop("_batched_cache", None)
        self.__dict__.pop("_fused_cache", None); self.__dict__.pop("fused", None); self.__dict__.pop("_fused_eval", None)
        if self.world > 1:
            _dist().barrier()
        self.arena.close()


def _measured_hbm_gbs() -> float:
    """Copy bandwidth from the driver-written ``MEASURED_PEAKS.json`` (fallback: the profiling recipe's 6567 GB/s)."""
    try:
        import json
        root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        return float(json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:  # noqa: BLE001
        return 6567.0


def copy_aggregator(agg):
    import copy
    return copy.deepcopy(agg)


def _ceil4(x: int) -> int:
    return (x + 3) // 4 * 4
