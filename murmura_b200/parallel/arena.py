"""Flat parameter arena: layout manifest, node→GPU placement, peer-mapped symmetric region.

Design: every virtual node's whole ``state_dict`` lives in ONE fp32 row

    [ parameters (Pp, padded to 4) | float buffers (Pb) | pad to 256 | int buffers as float (Pi) | pad to 256 ]

so local SGD writes parameters *in place* in the arena and the fused exchange+aggregate kernels see a
node as ``base + slot*stride``.  Each rank owns one symmetric region (cudaMalloc + cudaIpc, mapped
into every peer) holding three planes of ``S`` rows — ``live``, ``published[0]``, ``published[1]``
(double-buffered by round parity ⇒ the reference's Jacobi semantics, ``core/network.py:108,138-139``) —
followed by the published Count-Sketches and a control page (per-rank epoch flags, timeout mask).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn


def _ceil(x: int, m: int) -> int:
    return (x + m - 1) // m * m


@dataclass
class Entry:
    name: str
    shape: Tuple[int, ...]
    numel: int
    offset: int          # element offset inside the row (float region) or inside the int table
    kind: str            # "param" | "fbuf" | "int"
    ref_offset: int = -1  # offset in the reference's flatten order (float tensors, state_dict order)
    channels_last: bool = False   # 4-D weights stored physically as (O, H, W, I) for the NHWC tensor-core conv path

    def view(self, row: torch.Tensor) -> torch.Tensor:
        flat = row[self.offset:self.offset + self.numel]
        if self.channels_last:
            o, i, h, w = self.shape
            return flat.view(o, h, w, i).permute(0, 3, 1, 2)      # logical (O,I,H,W) with channels_last strides
        return flat.view(self.shape)


@dataclass
class StateLayout:
    """Maps ``state_dict`` keys ↔ offsets of one arena row."""
    entries: List[Entry] = field(default_factory=list)
    Pp: int = 0           # real parameter elements
    Pp4: int = 0          # parameter region padded to a multiple of 4 (SGD kernel extent)
    Pf: int = 0           # end of the float region (params + float buffers), real elements
    Pf_pad: int = 0       # float region padded to 256 → int tail starts here
    Pi: int = 0           # int-buffer elements
    stride: int = 0       # elements per row
    P_float_real: int = 0  # number of float state elements (reference's "model_dim")
    key_order: List[str] = field(default_factory=list)   # state_dict key order (what the reference flattens in)

    @classmethod
    def from_model(cls, model: nn.Module, channels_last: bool = False) -> "StateLayout":
        lay = cls()
        param_names = {n for n, _ in model.named_parameters()}
        state = model.state_dict()
        ref_off = 0
        ref_offsets: Dict[str, int] = {}
        for name, t in state.items():
            if t.is_floating_point():
                ref_offsets[name] = ref_off
                ref_off += t.numel()
        lay.P_float_real = ref_off
        lay.key_order = list(state.keys())
        off = 0
        for name, t in state.items():
            if name in param_names:
                lay.entries.append(Entry(name, tuple(t.shape), t.numel(), off, "param", ref_offsets[name],
                                         channels_last=channels_last and t.dim() == 4))
                off += t.numel()
        lay.Pp = off
        lay.Pp4 = _ceil(off, 4)
        off = lay.Pp4
        for name, t in state.items():
            if name not in param_names and t.is_floating_point():
                lay.entries.append(Entry(name, tuple(t.shape), t.numel(), off, "fbuf", ref_offsets[name]))
                off += t.numel()
        lay.Pf = off
        lay.Pf_pad = _ceil(max(off, 4), 256)
        ioff = 0
        for name, t in state.items():
            if not t.is_floating_point():
                lay.entries.append(Entry(name, tuple(t.shape), t.numel(), ioff, "int"))
                ioff += t.numel()
        lay.Pi = ioff
        lay.stride = lay.Pf_pad + (_ceil(ioff, 256) if ioff else 0)
        return lay

    # ---- views ------------------------------------------------------------------------------
    def float_entries(self) -> List[Entry]:
        return [e for e in self.entries if e.kind != "int"]

    def int_entries(self) -> List[Entry]:
        return [e for e in self.entries if e.kind == "int"]

    def row_views(self, row: torch.Tensor, ints: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """state-dict-shaped views of one arena row (+ optional int64 table row), in ``state_dict`` key order
        (aggregators that flatten a state — Sketchguard — depend on that order)."""
        found: Dict[str, torch.Tensor] = {}
        for e in self.entries:
            if e.kind == "int":
                if ints is not None:
                    found[e.name] = ints[e.offset:e.offset + e.numel].view(e.shape)
            else:
                found[e.name] = e.view(row)
        return {k: found[k] for k in self.key_order if k in found}

    def ref_permutation(self) -> np.ndarray:
        """``perm[arena_pos] = reference flatten index`` for real float elements, -1 for padding."""
        perm = np.full(self.Pf, -1, dtype=np.int64)
        for e in self.float_entries():
            logical = np.arange(e.ref_offset, e.ref_offset + e.numel)
            if e.channels_last:                       # physical order is (O,H,W,I); the reference flattens (O,I,H,W)
                logical = logical.reshape(e.shape).transpose(0, 2, 3, 1).ravel()
            perm[e.offset:e.offset + e.numel] = logical
        return perm

    def bind(self, model: nn.Module, row: torch.Tensor, grad_row: Optional[torch.Tensor], ints: Optional[torch.Tensor]) -> None:
        """Re-point ``model``'s parameters/buffers at the arena (values are copied in first)."""
        params = dict(model.named_parameters())
        modules = dict(model.named_modules())
        with torch.no_grad():
            for e in self.entries:
                if e.kind == "param":
                    p = params[e.name]
                    view = e.view(row)
                    view.copy_(p.detach())
                    p.data = view
                    if grad_row is not None:
                        p.grad = e.view(grad_row)
                else:
                    mod_name, _, leaf = e.name.rpartition(".")
                    mod = modules[mod_name]
                    old = mod._buffers[leaf]
                    if e.kind == "fbuf":
                        view = row[e.offset:e.offset + e.numel].view(e.shape)
                    else:
                        view = ints[e.offset:e.offset + e.numel].view(e.shape)
                    view.copy_(old)
                    mod._buffers[leaf] = view


class Placement:
    """Node → (rank, slot) map.

    ``weights=None``: contiguous packing (first ``N % G`` ranks host one extra node).  With per-node work weights
    (local SGD steps per round) nodes are assigned longest-first to the least-loaded rank (LPT) under the same
    per-rank capacity, so the per-round critical path ``max_rank Σ work`` is balanced — a round ends when the slowest GPU
    has finished training its virtual nodes.
    """

    def __init__(self, num_nodes: int, world: int, weights: Optional[List[float]] = None):
        self.num_nodes, self.world = num_nodes, world
        base, rem = divmod(num_nodes, world)
        self.counts = [base + (1 if r < rem else 0) for r in range(world)]
        self.slots_per_rank = max(self.counts) if self.counts else 0
        self.rank_of = np.zeros(num_nodes, dtype=np.int32)
        self.slot_of = np.zeros(num_nodes, dtype=np.int32)
        self._members: List[List[int]] = [[] for _ in range(world)]
        if weights is None:
            g = 0
            for r in range(world):
                for _ in range(self.counts[r]):
                    self._members[r].append(g); g += 1
        else:
            cap = list(self.counts)
            load = [0.0] * world
            for g in sorted(range(num_nodes), key=lambda i: (-float(weights[i]), i)):
                r = min((r for r in range(world) if len(self._members[r]) < cap[r]), key=lambda r: (load[r], r))
                self._members[r].append(g); load[r] += float(weights[g])
            for r in range(world):
                self._members[r].sort()
        for r in range(world):
            for s, g in enumerate(self._members[r]):
                self.rank_of[g] = r; self.slot_of[g] = s
        self.starts = np.concatenate([[0], np.cumsum(self.counts)]).tolist()

    def local_nodes(self, rank: int) -> List[int]:
        return list(self._members[rank])


class SymmetricArena:
    """One rank's peer-mapped region + typed views + device pointer tables for the kernels."""

    CTRL_BYTES = 4096

    def __init__(self, layout: StateLayout, placement: Placement, rank: int, device: torch.device,
                 sketch_size: int = 0, group=None, backend: str = "ipc"):
        from murmura_b200 import ops
        self.layout, self.placement, self.rank, self.device = layout, placement, rank, device
        self.world = placement.world
        S, stride = placement.slots_per_rank, layout.stride
        self.S = S
        self.K = sketch_size
        self.Kpad = _ceil(sketch_size, 32) if sketch_size else 0
        plane = S * stride * 4
        self.off_live = 0
        self.off_pub = plane                               # two planes follow
        self.off_sketch = self.off_pub + 2 * plane
        sk_bytes = _ceil(2 * S * self.K * 4, 256)
        self.off_sketch_q = self.off_sketch + sk_bytes
        q_bytes = _ceil(2 * S * self.Kpad, 256)
        self.off_sketch_sc = self.off_sketch_q + q_bytes
        sc_bytes = _ceil(2 * S * (self.Kpad // 32 if self.Kpad else 0), 256)
        self.off_rsum = _ceil(self.off_sketch_sc + sc_bytes, 256)     # per-rank column sums of the published rows, one per parity
        self.off_tot = _ceil(self.off_rsum + 2 * stride * 4, 256)      # all-reduced sum row (two-shot full-mesh FedAvg), one per parity
        self.off_ctrl = _ceil(self.off_tot + 2 * stride * 4, 4096)
        total = self.off_ctrl + self.CTRL_BYTES
        index = device.index if device.index is not None else torch.cuda.current_device()
        self.backend = backend if self.world > 1 else "ipc"
        self.mc_base = 0
        if self.backend == "symm":
            # torch symmetric memory: CUDA VMM allocation bound to an NVLS multicast object (multimem.* addresses);
            # peers are mapped like with cudaIpc, so every kernel works unchanged on this backend.
            import torch.distributed as dist
            import torch.distributed._symmetric_memory as symm_mem
            self._symm_t = symm_mem.empty(total, dtype=torch.uint8, device=device)
            self._symm_t.zero_()
            self._hdl = symm_mem.rendezvous(self._symm_t, (group or dist.group.WORLD))
            self._bases = [int(p) for p in self._hdl.buffer_ptrs]
            self.mc_base = int(self._hdl.multicast_ptr or 0)
            torch.cuda.synchronize(device)
            dist.barrier(group=group)
            self._arena = None
        else:
            self._arena = ops.ext().PeerArena(index, total, rank, self.world)
            if self.world > 1:
                import torch.distributed as dist
                handles: List[Optional[bytes]] = [None] * self.world
                dist.all_gather_object(handles, self._arena.ipc_handle(), group=group)
                self._arena.open_peers([bytes(h) for h in handles])
                dist.barrier(group=group)
            self._bases = [self._arena.base_ptr(r) for r in range(self.world)]
        self.live = self._view(rank, self.off_live, [S, stride], torch.float32)
        self.pub = self._view(rank, self.off_pub, [2, S, stride], torch.float32)
        self.sketch = self._view(rank, self.off_sketch, [2, S, self.K], torch.float32) if self.K else None
        self.flags = self._view(rank, self.off_ctrl, [16], torch.int32)
        self.timed_out = self._view(rank, self.off_ctrl + 64, [1], torch.int32)
        # device pointer tables (int64[G]) handed to kernels as `const T* const*`
        self.tbl_pub = self._ptr_table(self.off_pub)
        self.tbl_live = self._ptr_table(self.off_live)
        self.tbl_sketch = self._ptr_table(self.off_sketch)
        self.tbl_sketch_q = self._ptr_table(self.off_sketch_q)
        self.tbl_sketch_sc = self._ptr_table(self.off_sketch_sc)
        self.tbl_flags = self._ptr_table(self.off_ctrl)
        self.rsum = self._view(rank, self.off_rsum, [2, stride], torch.float32)
        self.rsum.zero_()
        self.tbl_rsum = [self._ptr_table(self.off_rsum + q * stride * 4) for q in (0, 1)]
        self.tbl_tot = [self._ptr_table(self.off_tot + q * stride * 4) for q in (0, 1)]

    def _view(self, rank: int, byte_offset: int, sizes: List[int], dtype: torch.dtype) -> torch.Tensor:
        if self._arena is not None:
            return self._arena.view(rank, byte_offset, sizes, dtype)
        esize = torch.empty((), dtype=dtype).element_size()
        if rank == self.rank:
            n = int(np.prod(sizes)) * esize
            return self._symm_t[byte_offset:byte_offset + n].view(dtype).view(sizes)
        return self._hdl.get_buffer(rank, sizes, dtype, byte_offset // esize)

    def _ptr_table(self, byte_offset: int) -> torch.Tensor:
        return torch.tensor([b + byte_offset for b in self._bases], dtype=torch.int64, device=self.device)

    # pointers -----------------------------------------------------------------------------------
    def pub_plane_ptr(self, parity: int) -> int:
        return self._bases[self.rank] + self.off_pub + parity * self.S * self.layout.stride * 4

    def mc_pub_plane_ptr(self, parity: int) -> int:
        """Multicast (NVLS) address of the published plane — 0 when the arena has no multicast mapping."""
        return self.mc_base + self.off_pub + parity * self.S * self.layout.stride * 4 if self.mc_base else 0

    def rsum_ptr(self, parity: int) -> int:
        return self._bases[self.rank] + self.off_rsum + parity * self.layout.stride * 4

    def tot_ptr(self, parity: int) -> int:
        return self._bases[self.rank] + self.off_tot + parity * self.layout.stride * 4

    def mc_tot_ptr(self, parity: int) -> int:
        return self.mc_base + self.off_tot + parity * self.layout.stride * 4 if self.mc_base else 0

    def mc_rsum_ptr(self, parity: int) -> int:
        return self.mc_base + self.off_rsum + parity * self.layout.stride * 4 if self.mc_base else 0

    def parity_off(self, parity: int) -> int:
        return parity * self.S * self.layout.stride

    def base_ptr(self, rank: int) -> int:
        return self._bases[rank]

    def peer_row(self, rank: int, parity: int, slot: int) -> torch.Tensor:
        """Published row of (rank, slot) as a tensor — local or peer-mapped (read in place over NVLink)."""
        off = self.off_pub + ((parity * self.S + slot) * self.layout.stride) * 4
        return self._view(rank, off, [self.layout.stride], torch.float32)

    def sketch_q_ptr(self) -> int:
        return self._bases[self.rank] + self.off_sketch_q

    def sketch_sc_ptr(self) -> int:
        return self._bases[self.rank] + self.off_sketch_sc

    def flags_ptr(self) -> int:
        return self._bases[self.rank] + self.off_ctrl

    def timed_out_ptr(self) -> int:
        return self._bases[self.rank] + self.off_ctrl + 64

    def close(self) -> None:
        if self._arena is not None:
            self._arena.close()
