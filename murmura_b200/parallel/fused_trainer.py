"""Fused local-SGD engine: one explicit forward/backward/SGD program for ALL virtual nodes of a GPU, no autograd.

Replaces the reference's hot loop #1 (``murmura/core/node.py:59-109``: fresh ``SGD`` → epochs × batches of
forward / loss / backward / step through stock autograd, one node after the other) for the bundled model families
(ResNet-18, the LEAF / CIFAR CNNs, the plain and evidential MLPs):

* every layer is ONE launch for all nodes hosted on the GPU (``blockIdx.z`` = node): convolutions and linear layers run the
  grouped implicit-GEMM kernels of ``ops/csrc/conv_tcgen05.cu`` (tcgen05 + TMEM; fprop / dgrad / wgrad), everything else the
  grouped kernels of ``ops/csrc/layers.cu``;
* weights are read and updated IN PLACE in the arena rows — the wgrad epilogue *is* the SGD step (``W += −lr·dW`` with
  ``red.global.add``; γ/β are stepped inside the BatchNorm backward), so there is no gradient buffer and no optimizer pass;
* nodes are ordered by their number of local steps, so the nodes still training at step *t* are a prefix of the group list
  and an unrolled round (all steps of all nodes) is captured once as a single CUDA graph.

The program is a static tape of ops built from the module tree (``build_program``); the same tape runs on a CPU through
:class:`EmuBackend` (NumPy emulation of the kernels' index formulas), which is how ``tests/test_fused_trainer.py`` checks it
against autograd without a GPU.
"""
from __future__ import annotations

import math
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn

from murmura_b200.ops import conv_plan as cp


# =====================================================================================================================
# buffers
# =====================================================================================================================
class Buf:
    """Activation (or gradient) of every group: ``t[g]`` is the flat [rows × ld] fp32 tensor of group ``g``."""

    def __init__(self, name: str, rows: int, C: int, ld: Optional[int] = None, dtype=torch.float32):
        self.name, self.rows, self.C, self.ld, self.dtype = name, rows, C, ld or cp.ceil4(C), dtype
        self.t: Optional[torch.Tensor] = None       # [Gmax, rows·ld] (a column slice of the zero pool for pooled buffers)
        self.gs = 0                                  # elements between groups
        self.pooled = False                          # member of the per-step zero pool (split-K / multi-writer targets)
        self.relu_fused = False                      # values are relu(·) of a conv epilogue: gradients w.r.t. it must be masked
        self.grad: Optional["Buf"] = None
        self.base: Optional["Buf"] = None            # reinterpretation of another buffer's storage

    @property
    def size(self) -> int:
        return self.rows * self.ld

    def ptr(self) -> int:
        return self.t.data_ptr()


# =====================================================================================================================
# backends
# =====================================================================================================================
class CudaBackend:
    """Launches the sm_100a kernels; every call is graph-capturable (no allocation, no synchronisation)."""

    name = "cuda"

    def __init__(self, device: torch.device):
        from murmura_b200 import ops
        self.ext = ops.ext()
        self.device = device
        self.ones = torch.ones(16, device=device)
        self._ptabs: Dict[Tuple[int, int, int], torch.Tensor] = {}
        self._tma: Dict[Any, Optional[Dict]] = {}
        self.launches = 0
        self.tma_launches = 0

    def ptab(self, shape: Tuple[int, int, int]) -> torch.Tensor:
        t = self._ptabs.get(shape)
        if t is None:
            t = self._ptabs[shape] = torch.from_numpy(cp.pixel_table(*shape)).to(self.device)
        return t

    def conv(self, tr: "FusedTrainer", plan: Dict, G: int, X: Buf, Y: Buf, R: Optional[Buf] = None, row_tab: Optional[torch.Tensor] = None,
             geom: Optional[cp.ConvGeom] = None, warena: Optional[torch.Tensor] = None) -> None:
        """``warena`` [groups, row] replaces the arena as the weight source (packed first-layer weights, indexed by group)."""
        d = {k: v for k, v in plan.items() if k != "ptab_shape"}
        arena = tr.live if warena is None else warena
        d.update(G=G, X=X.ptr(), x_gs=X.gs, Y=Y.ptr(), y_gs=Y.gs, arena=arena.data_ptr(), arena_gs=int(arena.shape[1]),
                 gmap=tr.gmap.data_ptr() if warena is None else 0, ones=self.ones.data_ptr())
        if R is not None:
            d.update(R=R.ptr(), r_gs=R.gs)
        self.launches += 1
        if warena is not None:
            row_tab = None                               # the packing kernel already resolved foreign rows
        parts = getattr(tr, "score_parts", None) if row_tab is not None else None
        if parts and geom is not None and getattr(tr, "use_tma", True):
            # foreign-weight scoring on the TMA path: one launch per source GPU, weight maps over that GPU's (peer-mapped) arena
            # [live | pub 0 | pub 1] planes, ``wslot[g]`` = plane·S + slot of group g's candidate row
            maps = []
            for base, slots, g0, g1 in parts:
                key = (geom, plan["mode"], X.ptr(), Y.ptr(), plan["w_off"], int(base), int(slots))
                if key not in self._tma:
                    from murmura_b200.ops.conv_launch import encode_tma
                    self._tma[key] = encode_tma(self.ext, plan["mode"], geom, x_ptr=X.ptr(), x_gs=X.gs, y_ptr=Y.ptr(), y_gs=Y.gs,
                                                w_ptr=int(base) + plan["w_off"] * 4, arena_stride=int(tr.stride), slots=int(slots),
                                                groups=int(X.t.shape[0]))
                maps.append(self._tma[key])
            if all(m is not None for m in maps):
                for (base, slots, g0, g1), extra in zip(parts, maps):
                    dd = dict(d); dd.update(extra)
                    dd.update(G=int(g1 - g0), g_off=int(g0), arena=int(base), arena_gs=int(tr.stride), gmap=tr.wslot.data_ptr())
                    self.ext.conv_tma(dd)
                    self.tma_launches += 1
                self.launches += len(parts) - 1
                return
        if geom is not None and row_tab is None and getattr(tr, "use_tma", True):
            key = (geom, plan["mode"], X.ptr(), Y.ptr(), plan["w_off"], arena.data_ptr())
            if key not in self._tma:
                from murmura_b200.ops.conv_launch import encode_tma
                self._tma[key] = encode_tma(self.ext, plan["mode"], geom, x_ptr=X.ptr(), x_gs=X.gs, y_ptr=Y.ptr(), y_gs=Y.gs,
                                            w_ptr=arena.data_ptr() + plan["w_off"] * 4, arena_stride=int(arena.shape[1]),
                                            slots=int(arena.shape[0]), groups=int(X.t.shape[0]))
            extra = self._tma[key]
            if extra is not None:
                d.update(extra)
                self.ext.conv_tma(d)
                self.tma_launches += 1
                return
        d.update(ptab=self.ptab(plan["ptab_shape"]).data_ptr())
        if row_tab is not None:
            d.update(row_tab=row_tab.data_ptr(), gmap=0)
        self.ext.conv_gemm(d)

    def call(self, fn: str, d: Dict) -> None:
        getattr(self.ext, fn)(d)
        self.launches += 1


class EmuBackend:
    """CPU stand-in with the kernels' semantics (NumPy / PyTorch), used by the CPU tests of the program builder."""

    name = "emu"

    def __init__(self, device: torch.device = torch.device("cpu")):
        self.device = device
        self.launches = 0

    @staticmethod
    def _np(buf: Buf, g: int) -> np.ndarray:
        return buf.t[g].numpy()

    def conv(self, tr: "FusedTrainer", plan: Dict, G: int, X: Buf, Y: Buf, R: Optional[Buf] = None, row_tab=None, geom=None, warena=None) -> None:
        for g in range(G):
            row = tr.live[int(tr.gmap[g])].numpy() if warena is None else warena[g].numpy()
            cp.emulate(plan, self._np(X, g), self._np(Y, g), row, None if R is None else self._np(R, g))
        self.launches += 1

    # ---- layers.cu equivalents (tensor arguments instead of addresses) ----
    def gather(self, tr: "FusedTrainer", G: int, t: int) -> None:
        """Mini-batch gather + im2col of the first layer + packing of its weights (``im2col_pack_kernel``)."""
        import torch.nn.functional as F
        f = tr.first
        for g in range(G):
            slot = int(tr.gmap[g])
            X, y = tr.shards[slot]
            idx = tr.perm[slot, t * tr.eb:(t + 1) * tr.eb]
            xb = X.index_select(0, idx).reshape(tr.eb, f["IH"], f["IW"], f["Cin"]).permute(0, 3, 1, 2)
            cols = F.unfold(xb, (f["KH"], f["KW"]), padding=f["pad"], stride=f["stride"])          # [eb, Cin·KH·KW, L], (c, kh, kw) order
            L = cols.shape[-1]
            cols = cols.view(tr.eb, f["Cin"], f["KH"] * f["KW"], L).permute(0, 3, 2, 1).reshape(tr.eb * L, f["Kreal"])
            out = tr.xb.t[g].view(tr.eb * L, f["Kpad"])
            out.zero_(); out[:, :f["Kreal"]] = cols
            tr.yb[g].copy_(y.index_select(0, idx))
            row = tr.live[slot]
            wp = tr.wpack[g]
            wp.zero_()
            wp[: f["Cout"] * f["Kpad"]].view(f["Cout"], f["Kpad"])[:, :f["Kreal"]] = row[f["w_off"]:f["w_off"] + f["Cout"] * f["Kreal"]].view(f["Cout"], f["Kreal"])
            base, Cp = f["Cout"] * f["Kpad"], cp.ceil4(f["Cout"])
            if f["bias_off"] >= 0:
                wp[base: base + f["Cout"]] = row[f["bias_off"]:f["bias_off"] + f["Cout"]]
            for j, key in enumerate(("bn_mean_off", "bn_var_off", "bn_gamma_off", "bn_beta_off")):
                if f.get(key, -1) >= 0:
                    wp[base + (j + 1) * Cp: base + (j + 1) * Cp + f["Cout"]] = row[f[key]:f[key] + f["Cout"]]
        tr.rng_step += 1
        self.launches += 1

    def bn_fwd(self, tr, G, op) -> None:
        M, C = op.x.rows, op.x.C
        for g in range(G):
            slot = int(tr.gmap[g]); row = tr.live[slot]
            x = op.x.t[g].view(M, C)
            mean = x.mean(0); var = x.var(0, unbiased=False)
            invstd = torch.rsqrt(var + op.eps)
            op.save_mean[g].copy_(mean); op.save_invstd[g].copy_(invstd)
            gam, bet = row[op.gamma_off:op.gamma_off + C], row[op.beta_off:op.beta_off + C]
            y = (x - mean) * invstd * gam + bet
            if op.res is not None:
                y = y + op.res.t[g].view(M, C)
            if op.relu:
                y = torch.relu(y)
            op.y.t[g].view(M, C).copy_(y)
            if op.rmean_off >= 0:
                rm, rv = row[op.rmean_off:op.rmean_off + C], row[op.rvar_off:op.rvar_off + C]
                rm.mul_(1 - op.momentum).add_(op.momentum * mean)
                rv.mul_(1 - op.momentum).add_(op.momentum * var * (M / max(M - 1, 1)))
            if op.nbt_off >= 0 and tr.ints is not None:
                tr.ints[slot, op.nbt_off] += 1
        self.launches += 1

    def bn_bwd(self, tr, G, op, lr: float) -> None:
        M, C = op.x.rows, op.x.C
        for g in range(G):
            slot = int(tr.gmap[g]); row = tr.live[slot]
            x = op.x.t[g].view(M, C); y = op.y.t[g].view(M, C); dy = op.y.grad.t[g].view(M, C).clone()
            mean, istd = op.save_mean[g], op.save_invstd[g]
            gam = row[op.gamma_off:op.gamma_off + C].clone()
            if op.relu:
                dy = dy * (y > 0)
            if op.res is not None:
                op.res.grad.t[g].view(M, C).copy_(dy)
            xh = (x - mean) * istd
            s1, s2 = dy.sum(0), (dy * xh).sum(0)
            op.x.grad.t[g].view(M, C).copy_(gam * istd * (dy - s1 / M - xh * s2 / M))
            row[op.beta_off:op.beta_off + C] -= lr * s1
            row[op.gamma_off:op.gamma_off + C] -= lr * s2
        self.launches += 1

    def maxpool_fwd(self, tr, G, op) -> None:
        import torch.nn.functional as F
        for g in range(G):
            x = op.x.t[g].view(op.B, op.H, op.W, op.C).permute(0, 3, 1, 2)
            y, idx = F.max_pool2d(x, op.k, op.stride, op.pad, return_indices=True)
            op.idx_t[g] = idx
            out = y if op.nchw_out else y.permute(0, 2, 3, 1)
            op.y.t[g][: out.numel()].copy_(out.reshape(-1))
        self.launches += 1

    def maxpool_bwd(self, tr, G, op) -> None:
        import torch.nn.functional as F
        for g in range(G):
            dy = op.y.grad.t[g][: op.B * op.C * op.OH * op.OW]
            dy = dy.view(op.B, op.C, op.OH, op.OW) if op.nchw_out else dy.view(op.B, op.OH, op.OW, op.C).permute(0, 3, 1, 2)
            dx = F.max_unpool2d(dy.contiguous(), op.idx_t[g], op.k, op.stride, op.pad, output_size=(op.H, op.W)) if op.k == op.stride and op.pad == 0 else \
                _unpool_overlapping(dy.contiguous(), op.idx_t[g], op.H, op.W)
            dx = dx.permute(0, 2, 3, 1).reshape(-1)
            if op.x.relu_fused:
                dx = dx * (op.x.t[g][: dx.numel()] > 0)
            op.x.grad.t[g][: dx.numel()].copy_(dx)
        self.launches += 1

    def avgpool(self, tr, G, op, backward: bool) -> None:
        for g in range(G):
            if not backward:
                op.y.t[g].view(op.B, op.C).copy_(op.x.t[g].view(op.B, op.HW, op.C).mean(1))
            else:
                op.x.grad.t[g].view(op.B, op.HW, op.C).copy_(op.y.grad.t[g].view(op.B, 1, op.C).expand(op.B, op.HW, op.C) / op.HW)
        self.launches += 1

    def loss(self, tr, G, op) -> None:
        import torch.nn.functional as F
        B, C, ld = op.x.rows, op.x.C, op.x.ld
        for g in range(G):
            slot = int(tr.gmap[g])
            out = op.x.t[g].view(B, ld)[:, :C].clone().requires_grad_(True)
            y = tr.yb[g]
            if op.evidential:
                from murmura_b200.models.mlp import evidential_loss_reference
                loss = evidential_loss_reference(out, y, float(tr.lam_t))
                (ga,) = torch.autograd.grad(loss, out)
                ga = ga * (1 - torch.exp(-(out.detach() - 1)))          # chain through alpha = softplus(z) + 1
            else:
                loss = F.cross_entropy(out, y)
                (ga,) = torch.autograd.grad(loss, out)
            gbuf = op.x.grad.t[g].view(B, ld); gbuf.zero_(); gbuf[:, :C] = ga
            tr.loss_acc[slot] += loss.detach()
        self.launches += 1


def _unpool_overlapping(dy: torch.Tensor, idx: torch.Tensor, H: int, W: int) -> torch.Tensor:
    B, C = dy.shape[:2]
    dx = torch.zeros(B, C, H * W, dtype=dy.dtype)
    dx.scatter_add_(2, idx.reshape(B, C, -1), dy.reshape(B, C, -1))
    return dx.view(B, C, H, W)


# =====================================================================================================================
# ops of the tape
# =====================================================================================================================
class ConvOp:
    """Conv2d / Linear (+bias) (+ReLU | softplus+1) — fprop, dgrad, wgrad+SGD launches of one layer."""

    def __init__(self, name: str, x: Buf, y: Buf, geom: cp.ConvGeom, w_off: int, bias_off: int, relu: bool = False, act: int = 0,
                 first: bool = False, bn: Optional[Dict[str, Any]] = None, res: Optional[Buf] = None):
        self.name, self.x, self.y, self.geom, self.w_off, self.bias_off = name, x, y, geom, w_off, bias_off
        self.relu, self.act, self.first = relu, act, first
        self.bn, self.res = bn, res                    # inference only: eval-mode BatchNorm (+residual) folded into the epilogue
        if first:
            # ``x`` is the im2col buffer [rows, Kpad]; fprop reads the per-step packed weights (rows of Kpad floats + bias),
            # wgrad steps the real weights in the arena (rows of Kreal floats)
            self.geom_f = cp.ConvGeom(B=geom.B, IH=1, IW=1, Cin=geom.Cin_pad, Cout=geom.Cout)
            self.pf = cp.plan_fprop(self.geom_f)
            self.f_bias_off = geom.Cout * geom.Cin_pad if bias_off >= 0 else -1
        else:
            self.geom_f = geom
            self.pf = cp.plan_fprop(geom, aligned_weights=(w_off % 4 == 0))
            self.f_bias_off = bias_off
        self.pd = None if first else cp.plan_dgrad(geom)
        self.pw = cp.plan_wgrad(geom, bias=bias_off >= 0)
        self.fused_out = relu or act != 0 or bn is not None or res is not None
        self.dgrad_accumulate = False                 # decided by the backward planner
        y.relu_fused = relu

    @staticmethod
    def _split(plan: Dict, G: int, target: int) -> int:
        gx, gy, _ = cp.grid_of(plan, 1)
        return cp.choose_splitk(gx * gy * G, cp.kb_total(plan), target)

    def may_split_fwd(self, target: int) -> bool:
        return not self.fused_out and self._split(self.pf, 1, target) > 1

    def may_split_bwd(self, target: int) -> bool:
        return self.pd is not None and not self.x.relu_fused and self._split(self.pd, 1, target) > 1

    def fwd(self, tr: "FusedTrainer", G: int) -> None:
        p = dict(self.pf)
        split = 1 if self.fused_out else self._split(p, G, tr.target_ctas)
        p.update(w_off=0 if self.first else self.w_off, bias_off=self.f_bias_off, relu=int(self.relu), act=self.act, splitk=split)
        if self.bn is not None:
            if self.first:                               # the packing kernel copied the statistics behind the packed weights
                base, Cp = self.geom.Cout * self.geom.Cin_pad, cp.ceil4(self.geom.Cout)
                p.update(bn_mean_off=base + Cp, bn_var_off=base + 2 * Cp, bn_gamma_off=base + 3 * Cp, bn_beta_off=base + 4 * Cp, eps=self.bn["eps"])
            else:
                p.update(bn_mean_off=self.bn["running_mean"], bn_var_off=self.bn["running_var"], bn_gamma_off=self.bn["weight"],
                         bn_beta_off=self.bn["bias"], eps=self.bn["eps"])
        assert split == 1 or self.y.pooled
        tr.be.conv(tr, p, G, self.x, self.y, R=self.res, row_tab=getattr(tr, "row_tab", None), geom=self.geom_f,
                   warena=tr.wpack if self.first else None)

    def bwd(self, tr: "FusedTrainer", G: int, lr: float) -> None:
        dy = self.y.grad
        if self.pd is not None:
            p = dict(self.pd)
            mask = self.x.relu_fused
            strided = self.geom.stride > 1                  # parity-class launch: no split-K, classes may not cover every pixel
            split = 1 if (mask or strided) else self._split(p, G, tr.target_ctas)
            p.update(w_off=self.w_off, splitk=split, accumulate=int(self.dgrad_accumulate or split > 1 or strided))
            assert not ((split > 1 or strided) and not self.dgrad_accumulate) or self.x.grad.pooled
            if mask:
                p.update(rmode=2)
            tr.be.conv(tr, p, G, dy, self.x.grad, R=self.x if mask else None, geom=self.geom)
        p = dict(self.pw)
        gx, gy, _ = cp.grid_of(p, 1)
        p.update(w_off=self.w_off, bias_off=self.bias_off, alpha=-lr, splitk=cp.choose_splitk(gx * gy * G, cp.kb_total(p), tr.target_ctas, min_kb=2))
        tr.wgrad_launch(lambda: tr.be.conv(tr, p, G, self.x, dy, geom=self.geom))


class BNOp:
    """Training BatchNorm (+residual) (+ReLU) (+dropout) forward / backward (γ, β stepped in the backward)."""

    def __init__(self, name: str, x: Buf, y: Buf, res: Optional[Buf], offs: Dict[str, int], relu: bool, eps: float, momentum: float,
                 p_drop: float = 0.0, layer_id: int = 0):
        self.name, self.x, self.y, self.res, self.relu, self.eps, self.momentum = name, x, y, res, relu, eps, momentum
        self.gamma_off, self.beta_off = offs["weight"], offs["bias"]
        self.rmean_off, self.rvar_off, self.nbt_off = offs.get("running_mean", -1), offs.get("running_var", -1), offs.get("num_batches_tracked", -1)
        self.p_drop, self.layer_id = p_drop, layer_id
        self.save_mean: Optional[torch.Tensor] = None
        self.save_invstd: Optional[torch.Tensor] = None
        assert res is None or p_drop == 0.0

    def _common(self, tr: "FusedTrainer", G: int) -> Dict[str, Any]:
        return dict(G=G, x_gs=self.x.gs, y_gs=self.y.gs, arena=tr.live.data_ptr(), arena_gs=tr.stride, gmap=tr.gmap.data_ptr(), gamma_off=self.gamma_off,
                    beta_off=self.beta_off, M=self.x.rows, C=self.x.C, relu=int(self.relu), p_drop=self.p_drop, layer_id=self.layer_id,
                    seed=tr.seed, rng_step=tr.rng_step.data_ptr(), save_mean=self.save_mean.data_ptr(), save_invstd=self.save_invstd.data_ptr())

    def fwd(self, tr: "FusedTrainer", G: int) -> None:
        if tr.be.name == "emu":
            return tr.be.bn_fwd(tr, G, self)
        d = self._common(tr, G)
        d.update(x=self.x.ptr(), y=self.y.ptr(), res=self.res.ptr() if self.res is not None else 0, res_gs=self.res.gs if self.res is not None else 0,
                 rmean_off=self.rmean_off, rvar_off=self.rvar_off, eps=self.eps, momentum=self.momentum)
        if self.nbt_off >= 0 and tr.ints is not None:
            d.update(nbt=tr.ints.data_ptr(), nbt_gs=tr.ints.shape[1], nbt_off=self.nbt_off)
        tr.be.call("bn_fwd_grouped", d)

    def bwd(self, tr: "FusedTrainer", G: int, lr: float) -> None:
        if tr.be.name == "emu":
            return tr.be.bn_bwd(tr, G, self, lr)
        d = self._common(tr, G)
        d.update(dy=self.y.grad.ptr(), dy_gs=self.y.grad.gs, x=self.x.ptr(), y=self.y.ptr(), dx=self.x.grad.ptr(), dx_gs=self.x.grad.gs,
                 dres=self.res.grad.ptr() if self.res is not None else 0, dres_gs=self.res.grad.gs if self.res is not None else 0, lr=lr)
        tr.be.call("bn_bwd_grouped", d)


class MaxPoolOp:
    def __init__(self, name: str, x: Buf, y: Buf, B: int, H: int, W: int, C: int, k: int, stride: int, pad: int, nchw_out: bool):
        self.name, self.x, self.y, self.B, self.H, self.W, self.C, self.k, self.stride, self.pad = name, x, y, B, H, W, C, k, stride, pad
        self.OH, self.OW = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
        self.nchw_out = nchw_out
        self.idx: Optional[torch.Tensor] = None
        self.idx_t: Dict[int, torch.Tensor] = {}

    def _geo(self, G: int) -> Dict[str, Any]:
        return dict(G=G, B=self.B, H=self.H, W=self.W, C=self.C, OH=self.OH, OW=self.OW, k=self.k, stride=self.stride, pad=self.pad,
                    nchw_out=int(self.nchw_out), idx=self.idx.data_ptr() if self.idx is not None else 0,
                    idx_gs=self.idx.shape[1] if self.idx is not None else 0)

    def fwd(self, tr: "FusedTrainer", G: int) -> None:
        if tr.be.name == "emu":
            return tr.be.maxpool_fwd(tr, G, self)
        d = self._geo(G); d.update(x=self.x.ptr(), x_gs=self.x.gs, y=self.y.ptr(), y_gs=self.y.gs)
        tr.be.call("maxpool_fwd_grouped", d)

    def bwd(self, tr: "FusedTrainer", G: int, lr: float) -> None:
        if tr.be.name == "emu":
            return tr.be.maxpool_bwd(tr, G, self)
        d = self._geo(G)
        d.update(dy=self.y.grad.ptr(), dy_gs=self.y.grad.gs, x=self.x.ptr(), x_gs=self.x.gs, dx=self.x.grad.ptr(), dx_gs=self.x.grad.gs,
                 relu_mask=int(self.x.relu_fused))
        tr.be.call("maxpool_bwd_grouped", d)


class AvgPoolOp:
    def __init__(self, name: str, x: Buf, y: Buf, B: int, HW: int, C: int):
        self.name, self.x, self.y, self.B, self.HW, self.C = name, x, y, B, HW, C

    def fwd(self, tr: "FusedTrainer", G: int) -> None:
        if tr.be.name == "emu":
            return tr.be.avgpool(tr, G, self, False)
        tr.be.call("avgpool_grouped", dict(G=G, B=self.B, HW=self.HW, C=self.C, x=self.x.ptr(), x_gs=self.x.gs, y=self.y.ptr(), y_gs=self.y.gs))

    def bwd(self, tr: "FusedTrainer", G: int, lr: float) -> None:
        if tr.be.name == "emu":
            return tr.be.avgpool(tr, G, self, True)
        tr.be.call("avgpool_grouped", dict(G=G, B=self.B, HW=self.HW, C=self.C, backward=1, dy=self.y.grad.ptr(), dy_gs=self.y.grad.gs,
                                           dx=self.x.grad.ptr(), dx_gs=self.x.grad.gs))


class DropoutOp:
    """Stand-alone dropout after a fused ReLU (FEMNIST-xlarge); CUDA only (the mask is a device-side Philox stream)."""

    def __init__(self, name: str, x: Buf, y: Buf, p: float, layer_id: int):
        self.name, self.x, self.y, self.p, self.layer_id = name, x, y, p, layer_id

    def _d(self, tr: "FusedTrainer", G: int) -> Dict[str, Any]:
        return dict(G=G, n=self.x.size, gmap=tr.gmap.data_ptr(), rng_step=tr.rng_step.data_ptr(), seed=tr.seed,
                    layer_id=self.layer_id, p_drop=self.p)

    def fwd(self, tr: "FusedTrainer", G: int) -> None:
        assert tr.be.name == "cuda"
        d = self._d(tr, G); d.update(x=self.x.ptr(), x_gs=self.x.gs, y=self.y.ptr(), y_gs=self.y.gs)
        tr.be.call("dropout_grouped", d)

    def bwd(self, tr: "FusedTrainer", G: int, lr: float) -> None:
        d = self._d(tr, G); d.update(x=self.y.grad.ptr(), x_gs=self.y.grad.gs, y=self.x.grad.ptr(), y_gs=self.x.grad.gs)
        if self.x.relu_fused:
            d.update(mask=self.x.ptr(), m_gs=self.x.gs)
        tr.be.call("dropout_grouped", d)


class LossOp:
    def __init__(self, x: Buf, evidential: bool):
        self.name, self.x, self.evidential = "loss", x, evidential

    def fwd(self, tr: "FusedTrainer", G: int) -> None:        # forward value + gradient w.r.t. the pre-activation in one launch
        if tr.be.name == "emu":
            return tr.be.loss(tr, G, self)
        d = dict(G=G, B=self.x.rows, C=self.x.C, ld=self.x.ld, out=self.x.ptr(), gs=self.x.gs, targets=tr.yb.data_ptr(), t_gs=tr.yb.shape[1],
                 grad=self.x.grad.ptr(), grad_gs=self.x.grad.gs, loss_acc=tr.loss_acc.data_ptr(), gmap=tr.gmap.data_ptr(),
                 evidential=int(self.evidential), lam=tr.lam_t.data_ptr())
        tr.be.call("loss_grouped", d)

    def bwd(self, tr: "FusedTrainer", G: int, lr: float) -> None:
        pass


# =====================================================================================================================
# program builder
# =====================================================================================================================
class _Unsupported(Exception):
    pass


class _Builder:
    def __init__(self, tr: "FusedTrainer", B: int, training: bool = True):
        self.tr, self.B, self.ops, self.training = tr, B, [], training
        self.layer_id = 0

    def off(self, name: str) -> int:
        return self.tr.offsets[name]

    def buf(self, name: str, rows: int, C: int, ld: Optional[int] = None) -> Buf:
        b = Buf(name, rows, C, ld)
        self.tr.bufs.append(b)
        return b

    def conv_bn(self, prefix: str, m: nn.Module, bn_prefix: str, bn: nn.Module, x: Buf, H: int, W: int, res: Optional[Buf], relu: bool,
                first: bool = False, p_drop: float = 0.0) -> Tuple[Buf, int, int]:
        """conv → BatchNorm (+residual) (+ReLU) (+dropout): two launches when training (batch statistics), ONE for inference
        (running statistics, residual and ReLU folded into the conv epilogue; dropout is the identity)."""
        if self.training:
            r, OH, OW = self.conv(prefix, m, x, H, W, first=first)
            return self.bn(bn_prefix, bn, r, res, relu=relu, p_drop=p_drop), OH, OW
        offs = {k: self.off(f"{bn_prefix}.{k}") for k in ("weight", "bias", "running_mean", "running_var")}
        offs["eps"] = bn.eps
        return self.conv(prefix, m, x, H, W, relu=relu, first=first, bn=offs, res=res)

    def conv(self, prefix: str, m: nn.Module, x: Buf, H: int, W: int, relu: bool = False, act: int = 0, first: bool = False,
             bn: Optional[Dict[str, Any]] = None, res: Optional[Buf] = None) -> Tuple[Buf, int, int]:
        bias = self.off(prefix + ".bias") if m.bias is not None else -1
        if isinstance(m, nn.Conv2d):
            assert m.groups == 1 and m.dilation == (1, 1) and m.kernel_size[0] == m.kernel_size[1] and m.stride[0] == m.stride[1] \
                and m.padding[0] == m.padding[1] and m.padding_mode == "zeros"
            g = cp.ConvGeom(B=self.B, IH=H, IW=W, Cin=m.in_channels, Cout=m.out_channels, KH=m.kernel_size[0], KW=m.kernel_size[1],
                            stride=m.stride[0], pad=m.padding[0], Cin_pad=cp.ceil4(m.in_channels) if first else x.ld)
        else:
            g = cp.ConvGeom(B=self.B, IH=1, IW=1, Cin=m.in_features, Cout=m.out_features, Cin_pad=cp.ceil4(m.in_features) if first else x.ld)
        if first:
            # the gather kernel writes the im2col matrix of the batch: the layer becomes a GEMM over rows of Kpad floats
            Kreal = g.KH * g.KW * g.Cin
            Kpad = cp.ceil32(Kreal)
            rows = self.B * g.OH * g.OW
            xcol = self.buf("input.col", rows, Kreal, Kpad)
            self.tr.xb = xcol
            self.tr.first = dict(IH=g.IH, IW=g.IW, Cin=g.Cin, KH=g.KH, KW=g.KW, stride=g.stride, pad=g.pad, OH=g.OH, OW=g.OW, Kreal=Kreal,
                                 Kpad=Kpad, Cout=g.Cout, w_off=self.off(prefix + ".weight"), bias_off=bias)
            if bn is not None:
                self.tr.first.update(bn_mean_off=bn["running_mean"], bn_var_off=bn["running_var"], bn_gamma_off=bn["weight"], bn_beta_off=bn["bias"])
            gl = cp.ConvGeom(B=rows, IH=1, IW=1, Cin=Kreal, Cout=g.Cout, Cin_pad=Kpad)
            y = self.buf(prefix + ".out", rows, g.Cout)
            self.ops.append(ConvOp(prefix, xcol, y, gl, self.off(prefix + ".weight"), bias, relu=relu, act=act, first=True, bn=bn, res=res))
            return y, g.OH, g.OW
        assert x.ld == g.Cin, f"{prefix}: input buffer has {x.ld} floats per pixel for {g.Cin} channels"
        if self.off(prefix + ".weight") % 4 or g.wrow % 4:
            raise _Unsupported(f"{prefix}: weights are not 16-byte aligned in the arena row")
        y = self.buf(prefix + ".out", self.B * g.OH * g.OW, g.Cout)
        self.ops.append(ConvOp(prefix, x, y, g, self.off(prefix + ".weight"), bias, relu=relu, act=act, first=False, bn=bn, res=res))
        return y, g.OH, g.OW

    def bn(self, prefix: str, m: nn.Module, x: Buf, res: Optional[Buf], relu: bool, p_drop: float = 0.0) -> Buf:
        assert m.affine and m.track_running_stats and m.momentum is not None
        y = self.buf(prefix + ".out", x.rows, x.C)
        offs = {k: self.off(f"{prefix}.{k}") for k in ("weight", "bias", "running_mean", "running_var")}
        offs["num_batches_tracked"] = self.tr.int_offsets.get(f"{prefix}.num_batches_tracked", -1)
        self.layer_id += 1
        op = BNOp(prefix, x, y, res, offs, relu, m.eps, m.momentum, p_drop, self.layer_id)
        self.tr.bn_ops.append(op)
        self.ops.append(op)
        return y

    def maxpool(self, name: str, x: Buf, H: int, W: int, C: int, k: int, s: int, p: int, nchw_out: bool = False) -> Tuple[Buf, int, int]:
        OH, OW = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
        y = self.buf(name + ".out", self.B, C * OH * OW) if nchw_out else self.buf(name + ".out", self.B * OH * OW, C)
        op = MaxPoolOp(name, x, y, self.B, H, W, C, k, s, p, nchw_out)
        self.tr.pool_ops.append(op)
        self.ops.append(op)
        return y, OH, OW

    def flat_view(self, x: Buf, rows: int, C: int) -> Buf:
        assert rows * C == x.rows * x.ld and x.ld == x.C
        v = self.buf(x.name + ".flat", rows, C)
        v.base = x
        return v


def _pool_params(m) -> Tuple[int, int, int]:
    k = m.kernel_size if isinstance(m.kernel_size, int) else m.kernel_size[0]
    s = m.stride if isinstance(m.stride, int) else m.stride[0]
    p = m.padding if isinstance(m.padding, int) else m.padding[0]
    return k, s or k, p


def build_program(tr: "FusedTrainer", model: nn.Module, B: int, sample_shape: Sequence[int], training: bool = True) -> bool:
    """Fill ``tr.ops`` / ``tr.bufs`` for ``model``; returns False when the model is not one of the supported families.
    ``training=False`` builds the inference tape (no loss, BatchNorm folded into the conv epilogues, dropout skipped)."""
    from murmura_b200.models.cnn import CIFARCNN, FEMNISTXLarge, _TwoConvNet
    from murmura_b200.models.mlp import MLP, EvidentialMLP
    from murmura_b200.models.resnet import BasicBlock, ResNet18
    b = _Builder(tr, B, training)
    image = len(sample_shape) == 3
    if image:
        Cin, H, W = sample_shape                      # logical (C, H, W); shards are stored NHWC
    else:
        H = W = 1
    x0 = None                                         # the first conv / linear layer creates the im2col input buffer (tr.xb)
    tr.first = None

    def check_channels(*convs) -> bool:
        return all(c.in_channels % 4 == 0 for c in convs)

    if isinstance(model, ResNet18) and image:
        x, H, W = b.conv_bn("conv1", model.conv1, "bn1", model.bn1, x0, H, W, None, True, first=True)
        k, s, p = _pool_params(model.maxpool)
        x, H, W = b.maxpool("maxpool", x, H, W, 64, k, s, p)
        for li in range(1, 5):
            layer = getattr(model, f"layer{li}")
            for bi, blk in enumerate(layer):
                assert isinstance(blk, BasicBlock)
                pre = f"layer{li}.{bi}"
                identity = x
                if blk.downsample is not None:
                    identity, _, _ = b.conv_bn(pre + ".downsample.0", blk.downsample[0], pre + ".downsample.1", blk.downsample[1], x, H, W, None, False)
                a1, H1, W1 = b.conv_bn(pre + ".conv1", blk.conv1, pre + ".bn1", blk.bn1, x, H, W, None, True)
                x, H, W = b.conv_bn(pre + ".conv2", blk.conv2, pre + ".bn2", blk.bn2, a1, H1, W1, identity, True)
        C = model.fc.in_features
        if H * W > 1:
            f = b.buf("avgpool.out", B, C)
            b.ops.append(AvgPoolOp("avgpool", x, f, B, H * W, C))
            x = f
        logits, _, _ = b.conv("fc", model.fc, x, 1, 1)
    elif isinstance(model, _TwoConvNet) and image:
        if not check_channels(model.conv2):
            return False
        x, H, W = b.conv("conv1", model.conv1, x0, H, W, relu=True, first=True)
        k, s, p = _pool_params(model.pool1)
        x, H, W = b.maxpool("pool1", x, H, W, model.conv1.out_channels, k, s, p)
        x, H, W = b.conv("conv2", model.conv2, x, H, W, relu=True)
        k, s, p = _pool_params(model.pool2)
        x, H, W = b.maxpool("pool2", x, H, W, model.conv2.out_channels, k, s, p, nchw_out=True)
        if x.C != model.fc1.in_features or x.C % 4:
            return False
        x, _, _ = b.conv("fc1", model.fc1, x, 1, 1, relu=True)
        logits, _, _ = b.conv("fc2", model.fc2, x, 1, 1)
    elif isinstance(model, CIFARCNN) and image:
        x, H, W = b.conv("conv1", model.conv1, x0, H, W, relu=True, first=True)
        x, H, W = b.conv("conv2", model.conv2, x, H, W, relu=True)
        x, H, W = b.maxpool("pool1", x, H, W, 64, 2, 2, 0)
        x, H, W = b.conv("conv3", model.conv3, x, H, W, relu=True)
        x, H, W = b.maxpool("pool2", x, H, W, 128, 2, 2, 0, nchw_out=True)
        if x.C != model.fc1.in_features:
            return False
        x, _, _ = b.conv("fc1", model.fc1, x, 1, 1, relu=True)
        logits, _, _ = b.conv("fc2", model.fc2, x, 1, 1)
    elif isinstance(model, FEMNISTXLarge) and image and (tr.be.name == "cuda" or not training):
        x, H, W = b.conv("conv1", model.conv1, x0, H, W, relu=True, first=True)
        x, H, W = b.conv("conv2", model.conv2, x, H, W, relu=True)
        x, H, W = b.maxpool("pool1", x, H, W, 128, 2, 2, 0)
        x, H, W = b.conv("conv3", model.conv3, x, H, W, relu=True)
        x, H, W = b.maxpool("pool2", x, H, W, 256, 2, 2, 0, nchw_out=True)
        if x.C != model.fc1.in_features:
            return False
        for i, fc in enumerate((model.fc1, model.fc2)):
            x, _, _ = b.conv(f"fc{i + 1}", fc, x, 1, 1, relu=True)
            if training:
                y = b.buf(f"drop{i + 1}.out", x.rows, x.C)
                b.layer_id += 1
                b.ops.append(DropoutOp(f"drop{i + 1}", x, y, model.dropout.p, b.layer_id))
                x = y
        logits, _, _ = b.conv("fc3", model.fc3, x, 1, 1)
    elif isinstance(model, MLP) and not image:
        mods = list(model.net.named_children())
        x, first = x0, True
        for idx, (name, m) in enumerate(mods):
            if isinstance(m, nn.Linear):
                relu = idx + 1 < len(mods) and isinstance(mods[idx + 1][1], nn.ReLU)
                if not first and m.in_features % 4:
                    return False
                x, _, _ = b.conv(f"net.{name}", m, x, 1, 1, relu=relu, first=first)
                first = False
            elif not isinstance(m, nn.ReLU):
                return False
        logits = x
    elif isinstance(model, EvidentialMLP) and not image:
        mods = list(model.feature_extractor.named_children())
        x, first, i = x0, True, 0
        while i < len(mods):
            name, m = mods[i]
            if not isinstance(m, nn.Linear) or (not first and m.in_features % 4) or m.out_features % 4:
                return False
            lin_name, lin, lin_first = name, m, first
            first = False
            j = i + 1
            bn_mod, bn_name, relu, p_drop = None, None, False, 0.0
            while j < len(mods) and not isinstance(mods[j][1], nn.Linear):
                n2, m2 = mods[j]
                if isinstance(m2, nn.BatchNorm1d):
                    bn_mod, bn_name = m2, n2
                elif isinstance(m2, nn.ReLU):
                    relu = True
                elif isinstance(m2, nn.Dropout):
                    p_drop = float(m2.p)
                elif not isinstance(m2, nn.Identity):
                    return False
                j += 1
            if bn_mod is None:
                return False
            if p_drop > 0 and tr.be.name != "cuda":
                p_drop = 0.0                           # the CPU emulation has no device Philox stream: tests run with p = 0
            x, _, _ = b.conv_bn(f"feature_extractor.{lin_name}", lin, f"feature_extractor.{bn_name}", bn_mod, x, 1, 1, None, relu,
                                first=lin_first, p_drop=p_drop)
            i = j
        if model.evidential_head.fc.in_features % 4:
            return False
        logits, _, _ = b.conv("evidential_head.fc", model.evidential_head.fc, x, 1, 1, act=2)
        tr.evidential_head = True
    else:
        return False
    if training:
        b.ops.append(LossOp(logits, tr.evidential_head))
    tr.ops = b.ops
    tr.logits = logits
    return True



def first_layer_args(tr, G: int, t: int, perm: torch.Tensor, x_tab: torch.Tensor, y_tab: torch.Tensor, row_tab: Optional[torch.Tensor] = None,
                     rng: bool = False) -> Dict[str, Any]:
    """Arguments of ``im2col_pack`` for a trainer / forward program ``tr`` (gather + im2col + first-layer weight packing)."""
    f = tr.first
    d = dict(G=G, x_tab=x_tab.data_ptr(), y_tab=y_tab.data_ptr(), perm=perm.data_ptr(), perm_ld=perm.shape[1], gmap=tr.gmap.data_ptr(),
             xcol=tr.xb.ptr(), xcol_gs=tr.xb.gs, yb=tr.yb.data_ptr(), yb_gs=tr.yb.shape[1], t=t, eb=tr.eb,
             arena=tr.live.data_ptr(), arena_gs=tr.stride, wmap=tr.gmap.data_ptr(), wpack=tr.wpack.data_ptr(), wpack_gs=tr.wpack.shape[1])
    d.update({k: v for k, v in f.items()})
    if row_tab is not None:
        d.update(row_tab=row_tab.data_ptr())
    if rng:
        d.update(rng_step=tr.rng_step.data_ptr(), ticket=tr.ticket.data_ptr())
    return d


def wpack_row_floats(first: Dict[str, Any]) -> int:
    return first["Cout"] * first["Kpad"] + 5 * cp.ceil4(first["Cout"])

# =====================================================================================================================
# the trainer
# =====================================================================================================================
class FusedTrainer:
    """Explicit fused training program over the nodes of one GPU.

    ``live`` [S, stride] is the arena's live plane (weights are read / stepped in place), ``ints`` the int64 buffer table,
    ``shards[slot] = (X [n, …] fp32 NHWC / flat, y [n] int64)`` the resident local datasets, ``steps[slot]`` the number of SGD
    steps of the node per round (0 = does not train: Byzantine or empty shard).
    """

    def __init__(self, model: nn.Module, layout, live: torch.Tensor, ints: Optional[torch.Tensor], shards: List[Tuple[torch.Tensor, torch.Tensor]],
                 steps: List[int], batch: int, sample_shape: Sequence[int], *, evidential: bool = False, seed: int = 0,
                 backend: Optional[Any] = None, target_ctas: int = 148, side_stream: bool = False):
        self.device = live.device
        self.be = backend or (CudaBackend(self.device) if live.is_cuda else EmuBackend())
        self.live, self.ints, self.stride = live, ints, int(live.shape[1])
        self.shards, self.eb, self.seed = shards, int(batch), int(seed)
        self.target_ctas = target_ctas
        self.offsets = {e.name: e.offset for e in layout.entries if e.kind != "int"}
        self.int_offsets = {e.name: e.offset for e in layout.entries if e.kind == "int"}
        self.bufs: List[Buf] = []
        self.ops: List[Any] = []
        self.bn_ops: List[BNOp] = []
        self.pool_ops: List[MaxPoolOp] = []
        self.evidential_head = False
        self.logits: Optional[Buf] = None
        self.xb: Optional[Buf] = None
        self.npix = self.Csrc = self.Cdst = 0
        try:
            self.supported = build_program(self, model, self.eb, sample_shape)
        except (_Unsupported, AssertionError) as exc:
            self.supported, self.unsupported_reason = False, str(exc)
        if self.supported and evidential != self.evidential_head:
            self.supported = False                      # criterion / head mismatch: leave it to the autograd path
        if not self.supported:
            return
        # ---- node order: longest first, so the nodes active at step t are the prefix [0, A_t) ------------------------
        S = len(steps)
        self.steps = list(steps)
        order = sorted([s for s in range(S) if steps[s] > 0], key=lambda s: (-steps[s], s))
        self.order = order
        self.Gmax = max(len(order), 1)
        self.max_steps = max([steps[s] for s in order], default=0)
        self.active = [sum(1 for s in order if steps[s] > t) for t in range(self.max_steps)]
        self.gmap = torch.tensor(order or [0], dtype=torch.int32, device=self.device)
        dev = self.device
        self.yb = torch.zeros(self.Gmax, self.eb, dtype=torch.int64, device=dev)
        self.loss_acc = torch.zeros(S, device=dev)
        self.lam_t = torch.zeros((), device=dev)
        self.rng_step = torch.zeros((), dtype=torch.int64, device=dev)
        self.ticket = torch.zeros((), dtype=torch.int32, device=dev)
        self.perm = torch.zeros(S, max(self.max_steps * self.eb, 1), dtype=torch.int64, device=dev)
        self.x_tab = torch.tensor([x.data_ptr() for x, _ in shards] or [0], dtype=torch.int64, device=dev)
        self.y_tab = torch.tensor([y.data_ptr() for _, y in shards] or [0], dtype=torch.int64, device=dev)
        self._plan_backward()
        self._allocate()
        self.side = torch.cuda.Stream(dev) if (side_stream and self.be.name == "cuda") else None
        self._side_pending = False
        self.graph = None
        self.graph_key = None

    # ---- static backward plan: who writes which gradient first ------------------------------------------------------
    def _grad(self, b: Buf) -> Buf:
        if b.base is not None:
            g = self._grad(b.base)
            if b.grad is None:
                b.grad = Buf(b.name + ".grad", b.rows, b.C, b.ld)
                b.grad.base = g
                self.bufs.append(b.grad)
            return b.grad
        if b.grad is None:
            b.grad = Buf(b.name + ".grad", b.rows, b.C, b.ld)
            self.bufs.append(b.grad)
        return b.grad

    def _plan_backward(self) -> None:
        written: set = set()

        def root(b: Buf) -> Buf:
            return root(b.base) if b.base is not None else b

        def first_write(b: Buf) -> bool:
            r = root(b)
            new = r not in written
            written.add(r)
            return new

        for op in self.ops:                              # forward outputs that may be produced by split-K reductions
            if isinstance(op, ConvOp) and op.may_split_fwd(self.target_ctas):
                op.y.pooled = True
        self._grad(self.logits)
        first_write(self.logits)
        for op in reversed(self.ops):
            if isinstance(op, ConvOp):
                self._grad(op.y)
                if op.pd is not None:
                    g = self._grad(op.x)
                    op.dgrad_accumulate = not first_write(op.x)
                    if (op.may_split_bwd(self.target_ctas) or op.geom.stride > 1) and not op.dgrad_accumulate:
                        root(g).pooled = True             # split-K slices / the parity classes of a strided dgrad add into zeros
            elif isinstance(op, BNOp):
                self._grad(op.y); self._grad(op.x)
                assert first_write(op.x), f"{op.name}: BatchNorm input has another gradient producer"
                if op.res is not None:
                    self._grad(op.res)
                    assert first_write(op.res), f"{op.name}: residual gradient must be the first contribution"
            elif isinstance(op, (MaxPoolOp, AvgPoolOp, DropoutOp)):
                self._grad(op.y); self._grad(op.x)
                assert first_write(op.x), f"{op.name}: input has another gradient producer"

    def _allocate(self) -> None:
        dev, G = self.device, self.Gmax
        roots = [b for b in self.bufs if b.base is None]
        pooled = [b for b in roots if b.pooled]
        off = 0
        for b in pooled:
            b._pool_off = off
            off += (b.size + 63) // 64 * 64
        self.pool_row = max(off, 64)
        self.zero_pool = torch.zeros(G, self.pool_row, device=dev)
        for b in pooled:
            b.t = self.zero_pool[:, b._pool_off:b._pool_off + b.size]
            b.gs = self.pool_row
        for b in roots:
            if not b.pooled:
                b.t = torch.zeros(G, b.size, dtype=b.dtype, device=dev)
                b.gs = b.size
        for b in self.bufs:
            if b.base is not None:
                r = b.base
                while r.base is not None:
                    r = r.base
                b.t, b.gs, b.pooled = r.t, r.gs, r.pooled
        self.wpack = torch.zeros(G, wpack_row_floats(self.first), device=dev)
        for op in self.bn_ops:
            op.save_mean = torch.zeros(G, op.x.C, device=dev)
            op.save_invstd = torch.zeros(G, op.x.C, device=dev)
        if self.be.name == "cuda":
            for op in self.pool_ops:
                op.idx = torch.zeros(G, op.B * op.OH * op.OW * op.C, dtype=torch.uint8, device=dev)
        self.workspace_bytes = sum(b.t.numel() * 4 for b in roots if not b.pooled) + self.zero_pool.numel() * 4

    # ---- one SGD step of the first G groups -------------------------------------------------------------------------
    def wgrad_launch(self, fn) -> None:
        """Weight-gradient (+SGD) launches go to the side stream when enabled: only the data-gradient chain is on the critical path."""
        if self.side is None:
            return fn()
        main = torch.cuda.current_stream(self.device)
        self.side.wait_stream(main)                      # the layer's dgrad (reads the same weights) was enqueued before us
        with torch.cuda.stream(self.side):
            fn()
        self._side_pending = True

    def _gather(self, G: int, t: int) -> None:
        if self.be.name == "emu":
            return self.be.gather(self, G, t)
        self.be.call("im2col_pack", first_layer_args(self, G, t, self.perm, self.x_tab, self.y_tab, rng=True))

    def step(self, G: int, t: int, lr: float) -> None:
        if G <= 0:
            return
        self.zero_pool[:G].zero_()
        self._gather(G, t)
        for op in self.ops:
            op.fwd(self, G)
        for op in reversed(self.ops):
            op.bwd(self, G, lr)
        if self.side is not None and self._side_pending:  # next step's forward reads the updated weights
            torch.cuda.current_stream(self.device).wait_stream(self.side)
            self._side_pending = False

    def run_steps(self, lr: float) -> None:
        for t in range(self.max_steps):
            self.step(self.active[t], t, lr)

    # ---- per-round driver ----------------------------------------------------------------------------------------------
    def refresh_permutations(self, epochs: int, generator: Optional[torch.Generator] = None) -> None:
        """Fresh shuffles: for every node ``epochs`` permutations of its shard, keeping ``nb·eb`` samples per epoch (drop_last)."""
        if not self.order:
            return
        ns = [int(self.shards[s][1].shape[0]) for s in range(len(self.shards))]
        n_max = max(ns)
        keys = torch.rand(len(ns), epochs, n_max, device=self.device, generator=generator)
        if getattr(self, "_valid_mask", None) is None:      # shard sizes are fixed: build the mask once (a per-round host list → device
            self._valid_mask = torch.arange(n_max, device=self.device)[None, None, :] < torch.tensor(ns, device=self.device)[:, None, None]
        valid = self._valid_mask                            # copy would block the host until the stream drains)
        keys = torch.where(valid, keys, torch.full_like(keys, 2.0))
        order = keys.argsort(dim=2)                                   # [S, epochs, n_max], invalid indices last
        for s in self.order:
            per_epoch = self.steps[s] // max(epochs, 1) * self.eb
            self.perm[s, : epochs * per_epoch] = order[s, :, :per_epoch].reshape(-1)

    def set_permutations(self, perms: Dict[int, torch.Tensor]) -> None:
        """Host-provided sample orders (seed-parity mode): ``perms[slot]`` = [epochs, nb·eb] indices of that node's shard."""
        for s in self.order:
            if s in perms:
                flat = perms[s].reshape(-1).to(self.device, non_blocking=True)
                self.perm[s, : flat.numel()] = flat

    def run_round(self, epochs: int, lr: float, use_graph: bool = True, perms: Optional[Dict[int, torch.Tensor]] = None) -> None:
        if not self.supported or self.max_steps == 0:
            return
        if perms is not None:
            self.set_permutations(perms)
        else:
            self.refresh_permutations(epochs)
        if self.be.name != "cuda" or not use_graph:
            return self.run_steps(lr)
        key = (epochs, lr)
        if self.graph is None or self.graph_key != key:
            self._capture(lr)
            self.graph_key = key
        self.graph.replay()

    def _capture(self, lr: float) -> None:
        snap = self.live.clone()
        snap_i = self.ints.clone() if self.ints is not None else None
        loss, rng = self.loss_acc.clone(), self.rng_step.clone()
        cap = torch.cuda.Stream(self.device)
        cap.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(cap):
            self.step(self.active[0], 0, lr)                          # warm-up: loads the kernels, sets the smem attributes
        torch.cuda.current_stream(self.device).wait_stream(cap)
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=cap):
            self.run_steps(lr)
        self.graph = graph
        self.live.copy_(snap)
        if snap_i is not None:
            self.ints.copy_(snap_i)
        self.loss_acc.copy_(loss); self.rng_step.copy_(rng)

    @property
    def launches_per_round(self) -> int:
        """Kernel launches of one round (all steps), counted while the program was captured / run."""
        per_step = len(self.ops) * 2 + 1
        return per_step * self.max_steps


# =====================================================================================================================
# inference tape: evaluation of the nodes' own models, scoring of foreign weight rows (UBAR stage 2 / EvidentialTrust / DMTT)
# =====================================================================================================================
class FusedForward:
    """Forward-only program over groups.

    A group is (weight row, data source): ``gmap[g]`` is the arena slot whose shard feeds the group (and whose weights are used
    unless ``row_tab`` holds per-group row addresses — foreign candidates read in place from the published planes, local HBM or a
    peer GPU over NVLink).  Eval-mode BatchNorm, residual adds, ReLU and the Dirichlet head are conv epilogues, so a ResNet-18
    forward of every group is 21 conv launches + 1 pooling launch; ``metrics`` reduces the outputs to the engine's
    ``[loss / sq-err, correct, count, vacuity, entropy, strength]`` rows on the device (``grouped_eval_kernel``).

    Reference sites replaced: ``murmura/core/node.py:111-196`` (evaluation), ``murmura/aggregation/ubar.py:152-202``,
    ``murmura/aggregation/evidential_trust.py:236-281``, ``murmura/dmtt/node_process.py:309-363`` (foreign-model scoring).
    """

    def __init__(self, model: nn.Module, layout, live: torch.Tensor, rows: int, sample_shape: Sequence[int], Gmax: int, *,
                 evidential: bool = False, backend: Optional[Any] = None):
        self.device = live.device
        self.be = backend or (CudaBackend(self.device) if live.is_cuda else EmuBackend())
        self.live, self.ints, self.stride = live, None, int(live.shape[1])
        self.eb, self.seed = int(rows), 0
        self.target_ctas = 0                            # every conv has its complete sum in one CTA (fused epilogues): no split-K
        self.offsets = {e.name: e.offset for e in layout.entries if e.kind != "int"}
        self.int_offsets = {e.name: e.offset for e in layout.entries if e.kind == "int"}
        self.bufs, self.ops, self.bn_ops, self.pool_ops = [], [], [], []
        self.evidential_head = False
        self.logits = self.xb = None
        self.npix = self.Csrc = self.Cdst = 0
        self.row_tab: Optional[torch.Tensor] = None
        self.score_parts = None                         # [(arena base of a source GPU, slots, g0, g1)]: scoring on the TMA path
        try:
            self.supported = build_program(self, model, self.eb, sample_shape, training=False)
        except (_Unsupported, AssertionError) as exc:
            self.supported, self.unsupported_reason = False, str(exc)
        if not self.supported:
            return
        self.dirichlet = bool(evidential and self.evidential_head)
        self.Gmax = max(int(Gmax), 1)
        dev = self.device
        for b in self.bufs:
            b.t = torch.zeros(self.Gmax, b.size, device=dev)
            b.gs = b.size
        if self.be.name == "cuda":
            for op in self.pool_ops:
                op.idx = torch.zeros(self.Gmax, op.B * op.OH * op.OW * op.C, dtype=torch.uint8, device=dev)
        self.gmap = torch.arange(self.Gmax, dtype=torch.int32, device=dev)
        self.wslot = torch.zeros(self.Gmax, dtype=torch.int32, device=dev)
        self.yb = torch.zeros(self.Gmax, self.eb, dtype=torch.int64, device=dev)
        self.stats = torch.zeros(self.Gmax, 8, device=dev)
        self.wpack = torch.zeros(self.Gmax, wpack_row_floats(self.first), device=dev)
        self.workspace_bytes = sum(b.t.numel() * 4 for b in self.bufs)

    def load(self, G: int, x_tab: torch.Tensor, y_tab: torch.Tensor, perm: torch.Tensor, t: int) -> None:
        """Inputs of the first ``G`` groups: rows ``perm[gmap[g], t·rows : (t+1)·rows]`` of shard ``gmap[g]`` (im2col'ed for the first
        layer) + labels, and the packed first-layer weights of the groups' weight rows (``row_tab`` or the arena slot)."""
        self.be.call("im2col_pack", first_layer_args(self, G, t, perm, x_tab, y_tab, row_tab=self.row_tab))

    def forward(self, G: int) -> None:
        for op in self.ops:
            op.fwd(self, G)

    def eval_descriptors(self, valid_rows: Sequence[int]) -> torch.Tensor:
        """Device table for :meth:`metrics`: (outputs, labels, #valid rows) of every group."""
        key = tuple(int(v) for v in valid_rows)
        cache = self.__dict__.setdefault("_desc_cache", {})
        if key not in cache:                            # content-keyed: no per-round (host-synchronising) pageable upload
            host = [[self.logits.ptr() + g * self.logits.gs * 4, self.yb.data_ptr() + g * self.yb.shape[1] * 8, v] for g, v in enumerate(key)]
            t = torch.tensor(host or [[0, 0, 0]], dtype=torch.int64)
            if self.device.type == "cuda":
                t = t.pin_memory()
                self.__dict__.setdefault("_desc_pinned", []).append(t)
            if len(cache) > 128:
                cache.clear()
            cache[key] = t.to(self.device, non_blocking=True)
        return cache[key]

    def metrics(self, G: int, desc: torch.Tensor, stats: Optional[torch.Tensor] = None, dirichlet: Optional[bool] = None) -> None:
        """Accumulate the per-group metric rows (zero ``stats`` first).  ``dirichlet=False`` forces softmax-CE on the raw outputs
        (UBAR scores even evidential models that way, reference ``aggregation/ubar.py:54,219``)."""
        st = self.stats if stats is None else stats
        self.be.ext.grouped_eval(desc, G, self.eb, self.logits.C, self.logits.ld, self.dirichlet if dirichlet is None else bool(dirichlet), st)
        self.be.launches += 1
