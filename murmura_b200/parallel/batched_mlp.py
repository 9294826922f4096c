"""Batched-over-virtual-nodes training of the MLP families.

The per-node CUDA graphs of the engine replay ≈ 45 launches per node and step; a GPU hosting V small MLP nodes therefore
issues V× that many latency-bound launches.  :class:`BatchedMLPTrainer` advances ALL nodes of a GPU by one SGD step with one
batched forward/backward: every layer is a strided-batched GEMM whose V weight matrices are *views into the arena rows*
(``rows[:, off:off+N·K].view(V, N, K)`` — batch stride = row stride, so no staging copy and SGD writes the arena in place),
BatchNorm1d uses per-node batch statistics and updates the per-node running statistics / ``num_batches_tracked`` under an
``active`` mask (nodes whose shard has fewer batches simply sit a step out), and the loss is the sum of the per-node means, so
each node receives exactly the gradient it would get from its own ``loss.backward()``.

Semantics are those of the reference's ``Node.local_train`` (``murmura/core/node.py:59-109``: fresh plain SGD, no momentum /
weight decay) applied node by node; ``tests/test_data_config.py`` checks the equality on CPU.  The GEMMs are ``torch.baddbmm``
(cuBLAS strided-batched on the GPU) — the grouped tcgen05 kernel of ``ops/csrc/mlp_tcgen05.cu`` covers the forward/scoring
direction only, its dX/dW variants are future work — hence the engine keeps this path opt-in
(``b200.batched_mlp_train``).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from murmura_b200.parallel.arena import StateLayout


@dataclass
class _Stage:
    kind: str                      # "linear" | "bn" | "relu" | "dropout" | "softplus1"
    tensors: Dict[str, torch.Tensor]
    eps: float = 1e-5
    momentum: float = 0.1
    p: float = 0.0


class BatchedMLPTrainer:
    """One instance per GPU; ``rows`` = ``live[:V]`` (float plane), ``ints`` = the int64 side table ``[V, Pi]``."""

    def __init__(self, model: nn.Module, layout: StateLayout, rows: torch.Tensor, ints: Optional[torch.Tensor]):
        from murmura_b200.models.mlp import MLP, EvidentialMLP
        self.V = int(rows.shape[0])
        self.rows, self.ints, self.layout = rows, ints, layout
        self.ent = {e.name: e for e in layout.entries}
        self.params: List[torch.Tensor] = []
        self.stages: List[_Stage] = []
        self.flatten = False
        self._pending: List[Tuple[_Stage, torch.Tensor, torch.Tensor]] = []
        if isinstance(model, EvidentialMLP):
            self.evidential = True
            self._walk(model.feature_extractor, "feature_extractor")
            self._linear("evidential_head.fc", model.evidential_head.fc)
            self.stages.append(_Stage("softplus1", {}))
        elif isinstance(model, MLP):
            self.evidential = False
            self.flatten = True
            self._walk(model.net, "net")
        else:
            raise TypeError(f"BatchedMLPTrainer supports the MLP / EvidentialMLP families, not {type(model).__name__}")

    # ---- arena views --------------------------------------------------------------------------------------------------------
    def _stack(self, name: str, leaf: bool) -> torch.Tensor:
        e = self.ent[name]
        t = self.rows[:, e.offset:e.offset + e.numel].view(self.V, *e.shape).detach()
        if leaf:
            t.requires_grad_(True)
            self.params.append(t)
        return t

    def _linear(self, prefix: str, m: nn.Linear) -> None:
        t = {"w": self._stack(prefix + ".weight", True)}
        if m.bias is not None:
            t["b"] = self._stack(prefix + ".bias", True)
        self.stages.append(_Stage("linear", t))

    def _walk(self, seq: nn.Sequential, prefix: str) -> None:
        for name, m in seq.named_children():
            key = f"{prefix}.{name}"
            if isinstance(m, nn.Linear):
                self._linear(key, m)
            elif isinstance(m, nn.BatchNorm1d):
                if not (m.affine and m.track_running_stats and m.momentum is not None):
                    raise TypeError("batched training needs affine BatchNorm1d with running statistics and a fixed momentum")
                t = {"gamma": self._stack(key + ".weight", True), "beta": self._stack(key + ".bias", True),
                     "rm": self._stack(key + ".running_mean", False), "rv": self._stack(key + ".running_var", False)}
                e = self.ent.get(key + ".num_batches_tracked")
                if e is not None and self.ints is not None:
                    t["nbt"] = self.ints[:, e.offset]
                self.stages.append(_Stage("bn", t, eps=float(m.eps), momentum=float(m.momentum)))
            elif isinstance(m, nn.ReLU):
                self.stages.append(_Stage("relu", {}))
            elif isinstance(m, nn.Dropout):
                self.stages.append(_Stage("dropout", {}, p=float(m.p)))
            elif not isinstance(m, nn.Identity):
                raise TypeError(f"unsupported layer {type(m).__name__} in batched MLP training")

    # ---- one SGD step of every node ---------------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor, active: torch.Tensor) -> torch.Tensor:
        """x: [V, B, …] → outputs [V, B, C] (training mode: batch statistics, dropout, running-stat updates of active nodes)."""
        h = x.flatten(2) if self.flatten or x.dim() > 3 else x
        self._pending = []
        for st in self.stages:
            t = st.tensors
            if st.kind == "linear":
                wt = t["w"].transpose(1, 2)
                h = torch.baddbmm(t["b"].unsqueeze(1), h, wt) if "b" in t else torch.bmm(h, wt)
            elif st.kind == "bn":
                B = h.shape[1]
                mean = h.mean(dim=1)
                var = h.var(dim=1, unbiased=False)
                h = (h - mean.unsqueeze(1)) * torch.rsqrt(var + st.eps).unsqueeze(1) * t["gamma"].unsqueeze(1) + t["beta"].unsqueeze(1)
                # every stacked tensor is a view of ONE storage (the arena plane) and shares its autograd version counter, so
                # the in-place running-stat update must wait until backward has consumed the saved views (see step()).
                self._pending.append((st, mean.detach(), var.detach() * (B / max(B - 1, 1))))
            elif st.kind == "relu":
                h = F.relu(h)
            elif st.kind == "dropout":
                h = F.dropout(h, st.p, training=True)
            elif st.kind == "softplus1":
                h = F.softplus(h) + 1
        return h

    def loss(self, out: torch.Tensor, y: torch.Tensor, lam: Any = 0.0) -> torch.Tensor:
        """Σ_nodes mean_batch loss — every node's gradient equals the one of its own mean loss."""
        V, B, C = out.shape
        flat, yf = out.reshape(V * B, C), y.reshape(V * B)
        if self.evidential:
            if flat.is_cuda:
                from murmura_b200 import ops
                if ops.available():
                    return ops.evidential_loss(flat, yf, lam) * V
            from murmura_b200.models.mlp import evidential_loss_reference
            return evidential_loss_reference(flat, yf, float(lam)) * V
        return F.cross_entropy(flat, yf) * V

    def step(self, x: torch.Tensor, y: torch.Tensor, active: torch.Tensor, lr: float, lam: Any = 0.0) -> torch.Tensor:
        """One plain-SGD step of every *active* node (``active``: float [V] of 0/1).  Returns the summed loss (detached)."""
        for p in self.params:
            p.grad = None
        out = self.forward(x, active)
        loss = self.loss(out, y, lam)
        loss.backward()
        with torch.no_grad():
            for st, mean, unbiased in self._pending:                     # BatchNorm bookkeeping of the nodes that took the step
                t = st.tensors
                a = (active * st.momentum).unsqueeze(1)
                t["rm"].add_(a * (mean - t["rm"]))
                t["rv"].add_(a * (unbiased - t["rv"]))
                if "nbt" in t:
                    t["nbt"].add_(active.to(t["nbt"].dtype))
            self._pending = []
            for p in self.params:
                coef = (-lr * active).view(self.V, *([1] * (p.dim() - 1)))
                p.add_(p.grad * coef)
        return loss.detach()

    @staticmethod
    def gather(xpad: torch.Tensor, ypad: torch.Tensor, idx: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """xpad [V, n_max, …], ypad [V, n_max], idx [V, B] → ([V, B, …], [V, B])."""
        ar = torch.arange(xpad.shape[0], device=xpad.device).unsqueeze(1)
        return xpad[ar, idx], ypad[ar, idx]
