"""Launch plans for ``conv_tcgen05.cu`` (grouped implicit-GEMM conv / linear layers) + a NumPy emulator of the kernel's addressing.

A plan is a plain ``dict`` whose keys are the fields of ``ConvGemmParams``; the geometry part is built here (pure
integer arithmetic, testable on a CPU), the pointer part (``X``, ``Y``, ``arena`` …) is filled in by the caller right
before the launch.  :func:`emulate` executes a plan with the *same* index formulas as the CUDA kernel (k → (tap, channel)
decode, live-tap list, pixel table, padded-channel remap, output mapping) on NumPy arrays, so the formulas are checked
against ``torch.nn.functional.conv2d`` and its autograd gradients without a GPU (``tests/test_conv_plan.py``); on the GPU
the kernel is then compared with this emulator's oracle (``tests/test_conv_gpu.py``).

Reference hot loop replaced: ``murmura/core/node.py:59-109`` (forward / backward / SGD step through autograd).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np

MODE_F, MODE_D, MODE_W = 0, 1, 2
BM, BK = 128, 32


def ceil4(x: int) -> int:
    return (x + 3) // 4 * 4


def ceil32(x: int) -> int:
    return (x + 31) // 32 * 32


@dataclass(frozen=True)
class ConvGeom:
    """One conv (or linear: 1×1 map, 1×1 kernel) layer at a fixed batch size.  Activations are NHWC with ``Cin_pad`` /
    ``Cout_pad`` floats per pixel (multiples of 4, zero padded); weights are (Cout, KH, KW, Cin) as stored in the arena."""
    B: int
    IH: int
    IW: int
    Cin: int
    Cout: int
    KH: int = 1
    KW: int = 1
    stride: int = 1
    pad: int = 0
    Cin_pad: int = 0
    Cout_pad: int = 0

    def __post_init__(self):
        object.__setattr__(self, "Cin_pad", self.Cin_pad or ceil4(self.Cin))
        object.__setattr__(self, "Cout_pad", self.Cout_pad or ceil4(self.Cout))
        assert self.Cin_pad % 4 == 0 and self.Cout_pad % 4 == 0 and self.Cin_pad >= self.Cin and self.Cout_pad >= self.Cout
        assert self.IH <= 256 and self.IW <= 256, "feature maps are at most 256 x 256"

    @property
    def OH(self) -> int:
        return (self.IH + 2 * self.pad - self.KH) // self.stride + 1

    @property
    def OW(self) -> int:
        return (self.IW + 2 * self.pad - self.KW) // self.stride + 1

    @property
    def wrow(self) -> int:
        return self.KH * self.KW * self.Cin

    def live_taps(self) -> List[int]:
        """Taps (kh·KW + kw) that land on at least one real input pixel for some output pixel."""
        rows = [kh for kh in range(self.KH) if any(0 <= oh * self.stride + kh - self.pad < self.IH for oh in range(self.OH))]
        cols = [kw for kw in range(self.KW) if any(0 <= ow * self.stride + kw - self.pad < self.IW for ow in range(self.OW))]
        return [kh * self.KW + kw for kh in rows for kw in cols]


def linear_geom(B: int, K: int, N: int) -> ConvGeom:
    return ConvGeom(B=B, IH=1, IW=1, Cin=K, Cout=N)


def pixel_table(B: int, H: int, W: int) -> np.ndarray:
    """``b << 16 | y << 8 | x`` for every pixel of a [B, H, W] plane, row-major."""
    assert B < 32768 and H <= 256 and W <= 256, "pixel table packs (b, y, x) as 15/8/8 bits"
    b, y, x = np.meshgrid(np.arange(B), np.arange(H), np.arange(W), indexing="ij")
    return ((b << 16) | (y << 8) | x).astype(np.int32).ravel()


def choose_splitk(ctas: int, kb_total: int, target_ctas: int = 148, min_kb: int = 4) -> int:
    """Split-K so that a launch has ≈ ``target_ctas`` CTAs while every slice keeps ≥ ``min_kb`` k-blocks."""
    if ctas >= target_ctas or kb_total <= min_kb:
        return 1
    return max(1, min((target_ctas + ctas - 1) // ctas, kb_total // min_kb))


def _bn_tile(n: int) -> int:
    return 128 if n > 64 else 64


def plan_fprop(g: ConvGeom, *, aligned_weights: bool = True) -> Dict:
    taps = g.live_taps()
    K = len(taps) * g.Cin_pad
    vec = 4 if (aligned_weights and g.Cin == g.Cin_pad and g.wrow % 4 == 0) else 1
    return {"mode": MODE_F, "M": g.B * g.OH * g.OW, "N": g.Cout, "K": K, "SH": g.IH, "SW": g.IW, "C": g.Cin_pad,
            "lds": g.Cin_pad, "KW": g.KW, "stride": g.stride, "pad": g.pad, "taps": taps, "Cw_real": g.Cin, "Ck_real": g.Cin_pad,
            "wrow": g.wrow, "ldy": g.Cout_pad, "vecB": vec, "BN": _bn_tile(g.Cout), "ptab_shape": (g.B, g.OH, g.OW)}


def plan_dgrad(g: ConvGeom) -> Dict:
    assert g.Cin == g.Cin_pad, "dgrad is never needed for (channel-padded) first layers"
    taps = g.live_taps()
    return {"mode": MODE_D, "M": g.B * g.IH * g.IW, "N": g.Cin, "K": len(taps) * g.Cout_pad, "SH": g.OH, "SW": g.OW,
            "C": g.Cout_pad, "lds": g.Cout_pad, "KW": g.KW, "stride": g.stride, "pad": g.pad, "taps": taps, "Cw_real": g.Cin,
            "Ck_real": g.Cout, "wrow": g.wrow, "ldy": g.Cin_pad, "vecB": 4, "BN": _bn_tile(g.Cin), "ptab_shape": (g.B, g.IH, g.IW)}


def plan_wgrad(g: ConvGeom, *, bias: bool) -> Dict:
    taps = g.live_taps()
    return {"mode": MODE_W, "M": len(taps) * g.Cin_pad, "N": g.Cout, "K": g.B * g.OH * g.OW, "SH": g.IH, "SW": g.IW,
            "C": g.Cin_pad, "lds": g.Cin_pad, "KW": g.KW, "stride": g.stride, "pad": g.pad, "taps": taps, "Cw_real": g.Cin,
            "Ck_real": g.Cin_pad, "wrow": g.wrow, "ldy": g.Cout_pad, "vecB": 4, "ones_row": 1 if bias else 0, "BN": _bn_tile(g.Cout),
            "ptab_shape": (g.B, g.OH, g.OW)}


# =====================================================================================================================
# TMA geometry (conv_tma.cu): M tiles = whole images or strips of image rows, so one im2col tile of one tap is one TMA box
# =====================================================================================================================
SWIZZLE_128B, SWIZZLE_128B_ATOM_32B = 3, 4


def tma_plan(mode: int, g: ConvGeom, arena_stride: int, S: int) -> Optional[Dict]:
    """Extra launch fields + tensor-map specs of the TMA-fed kernel, or ``None`` when this layer needs the cp.async gather
    (first layers with padded channels, channel counts that are not multiples of 32, dgrad of strided convolutions, wide maps).

    ``maps`` entries are ``(operand, dims, strides_bytes, box, elem_strides, swizzle)`` with the activation group stride left
    symbolic (``None``): the caller substitutes its buffers' group strides and base addresses."""
    s = g.stride
    if g.Cin_pad % 32 or (mode != MODE_W and (g.Cin != g.Cin_pad or g.wrow % 4)):
        return None

    def tile(RH: int, RW: int):
        hw = RH * RW
        if hw <= 128:
            Bt = 128 // hw
            return dict(Bt=Bt, TH=RH, tpi=1, RT=Bt * hw, mtiles=(g.B + Bt - 1) // Bt)
        if RW > 128:
            return None
        TH = 128 // RW
        tpi = (RH + TH - 1) // TH
        return dict(Bt=1, TH=TH, tpi=tpi, RT=TH * RW, mtiles=g.B * tpi)

    act = lambda C, W, H: ([C, W, H, g.B, None], [C * 4, W * C * 4, H * W * C * 4, None])
    live = g.live_taps()
    if mode == MODE_F:
        t = tile(g.OH, g.OW)
        if t is None or g.OW * s > 256 or t["TH"] * s > 256 or len(live) > 32:
            return None
        dims, strides = act(g.Cin_pad, g.IW, g.IH)
        cls = dict(py=0, px=0, taps=live, dx=[tp % g.KW - g.pad for tp in live], dy=[tp // g.KW - g.pad for tp in live])
        t.update(RH=g.OH, RW=g.OW, OPH=g.OH, OPW=g.OW, osc=1, ystep=s, classes=[cls], kb_total=len(live) * (g.Cin // 32),
                 maps=[("X", dims, strides, [32, g.OW * s, t["TH"] * s, t["Bt"], 1], [1, s, s, 1, 1], SWIZZLE_128B),
                       ("W", [g.wrow, g.Cout, S], [g.wrow * 4, arena_stride * 4], [32, _bn_tile(g.Cout), 1], [1, 1, 1], SWIZZLE_128B)])
        return t
    if mode == MODE_D:
        if g.Cout != g.Cout_pad or g.Cout % 32 or g.IH % s or g.IW % s or s > 2:
            return None
        # a stride-s dgrad = s² stride-1 problems over the parity classes of the input pixels
        classes = []
        for py in range(s):
            for px in range(s):
                taps = [tp for tp in live if (py + g.pad - tp // g.KW) % s == 0 and (px + g.pad - tp % g.KW) % s == 0]
                if taps:
                    classes.append(dict(py=py, px=px, taps=taps, dx=[(px + g.pad - tp % g.KW) // s for tp in taps],
                                        dy=[(py + g.pad - tp // g.KW) // s for tp in taps]))
        if not classes or max(len(c["taps"]) for c in classes) > 32:
            return None
        t = tile(g.IH // s, g.IW // s)
        if t is None:
            return None
        dims, strides = act(g.Cout_pad, g.OW, g.OH)
        T = g.KH * g.KW
        t.update(RH=g.IH // s, RW=g.IW // s, OPH=g.IH, OPW=g.IW, osc=s, ystep=1, classes=classes,
                 kb_total=max(len(c["taps"]) for c in classes) * (g.Cout // 32), partial=(len(classes) < s * s or s > 1),
                 maps=[("X", dims, strides, [32, g.IW // s, t["TH"], t["Bt"], 1], [1, 1, 1, 1, 1], SWIZZLE_128B),
                       ("W", [g.Cin, T, g.Cout, S], [g.Cin * 4, g.wrow * 4, arena_stride * 4], [32, 1, 32, 1], [1, 1, 1, 1], SWIZZLE_128B_ATOM_32B)])
        return t
    # ---- W: reduction over boxes of <= 32 output pixels ----
    if g.OW > 32 or g.OW * s > 256:
        return None
    hw = g.OH * g.OW
    if hw <= 32:
        bb, bh, bpi = 32 // hw, g.OH, 1
        kb = (g.B + bb - 1) // bb
    else:
        bb, bh = 1, max(1, 32 // g.OW)
        bpi = (g.OH + bh - 1) // bh
        kb = g.B * bpi
    if bh * s > 256:
        return None
    dx, sx = act(g.Cin_pad, g.IW, g.IH)
    dy, sy = act(g.Cout_pad, g.OW, g.OH)
    return dict(PK=bb * bh * g.OW, bh=bh, bb=bb, bpi=bpi, kb_total=kb,
                maps=[("X", dx, sx, [32, g.OW * s, bh * s, bb, 1], [1, s, s, 1, 1], SWIZZLE_128B_ATOM_32B),
                      ("Y", dy, sy, [32, g.OW, bh, bb, 1], [1, 1, 1, 1, 1], SWIZZLE_128B_ATOM_32B)])


def grid_of(plan: Dict, G: int = 1) -> tuple:
    m_ext = plan["M"] + (1 if plan["mode"] == MODE_W and plan.get("ones_row") else 0)
    return ((m_ext + BM - 1) // BM, (plan["N"] + plan["BN"] - 1) // plan["BN"], G)


def kb_total(plan: Dict) -> int:
    return (plan["K"] + BK - 1) // BK


# =====================================================================================================================
# NumPy emulator (same formulas as conv_tcgen05.cu; no tiling, no swizzle)
# =====================================================================================================================
def _gather_pixel(plan: Dict, pk: np.ndarray, kh: int, kw: int, dgrad: bool):
    b, y, x = pk >> 16, (pk >> 8) & 255, pk & 255
    s, pad = plan["stride"], plan["pad"]
    if dgrad:
        ty, tx = y + pad - kh, x + pad - kw
        ok = (ty >= 0) & (tx >= 0) & (ty % s == 0) & (tx % s == 0)
        sy, sx = ty // s, tx // s
    else:
        sy, sx = y * s + kh - pad, x * s + kw - pad
        ok = (sy >= 0) & (sx >= 0)
    ok &= (sy < plan["SH"]) & (sx < plan["SW"])
    pix = (b * plan["SH"] + sy) * plan["SW"] + sx
    return ok, np.where(ok, pix, 0)


def _im2col(plan: Dict, X: np.ndarray, ptab: np.ndarray, nrows: int, dgrad: bool) -> np.ndarray:
    """A[r, k] of modes F / D (rows = ptab entries, k = (live tap, channel))."""
    C, K = plan["C"], plan["K"]
    A = np.zeros((nrows, K), dtype=np.float64)
    for lt, tap in enumerate(plan["taps"]):
        kh, kw = divmod(tap, plan["KW"])
        ok, pix = _gather_pixel(plan, ptab[:nrows], kh, kw, dgrad)
        cols = X[(pix[:, None] * plan["lds"] + np.arange(C)[None, :])]
        A[:, lt * C:(lt + 1) * C] = np.where(ok[:, None], cols, 0.0)
    return A


def emulate(plan: Dict, X: np.ndarray, Y: np.ndarray, row: np.ndarray, R: Optional[np.ndarray] = None) -> None:
    """Run one group of ``plan``: ``X`` / ``Y`` / ``R`` are flat activation buffers, ``row`` the node's flat arena row.
    F, D write (or accumulate into) ``Y``; W updates ``row`` in place."""
    mode, M, N, K, C = plan["mode"], plan["M"], plan["N"], plan["K"], plan["C"]
    taps, Cw, wrow, ldy, w_off = plan["taps"], plan["Cw_real"], plan["wrow"], plan["ldy"], plan["w_off"]
    alpha = float(plan.get("alpha", 1.0))
    ptab = pixel_table(*plan["ptab_shape"])
    if mode in (MODE_F, MODE_D):
        A = _im2col(plan, X, ptab, M, dgrad=(mode == MODE_D))
        Bm = np.zeros((N, K), dtype=np.float64)
        for lt, tap in enumerate(taps):
            if mode == MODE_F:
                for cc in range(min(C, Cw)):
                    Bm[:, lt * C + cc] = row[w_off + np.arange(N) * wrow + tap * Cw + cc]
            else:
                for co in range(min(C, plan["Ck_real"])):
                    Bm[:, lt * C + co] = row[w_off + co * wrow + tap * Cw + np.arange(N)]
        out = alpha * (A @ Bm.T)
        if plan.get("bias_off", -1) >= 0:
            out = out + row[plan["bias_off"] + np.arange(N)][None, :]
        if plan.get("bn_mean_off", -1) >= 0:
            mean, var = row[plan["bn_mean_off"] + np.arange(N)], row[plan["bn_var_off"] + np.arange(N)]
            gam = row[plan["bn_gamma_off"] + np.arange(N)] if plan.get("bn_gamma_off", -1) >= 0 else 1.0
            bet = row[plan["bn_beta_off"] + np.arange(N)] if plan.get("bn_beta_off", -1) >= 0 else 0.0
            out = (out - mean) / np.sqrt(var + plan.get("eps", 1e-5)) * gam + bet
        idx = (np.arange(M)[:, None] * ldy + np.arange(N)[None, :])
        if R is not None:
            out = np.where(R[idx] > 0, out, 0.0) if plan.get("rmode", 1) == 2 else out + R[idx]
        if plan.get("relu"):
            out = np.maximum(out, 0.0)
        elif plan.get("act") == 2:
            out = np.where(out > 20, out, np.log1p(np.exp(np.minimum(out, 20)))) + 1.0
        if plan.get("accumulate"):
            Y[idx] += out.astype(Y.dtype)
        else:
            Y[idx] = out.astype(Y.dtype)
        return
    # ---- W: rows m = (live tap, channel), reduction over output pixels ----
    P = K
    Mreal = len(taps) * C
    A = np.zeros((Mreal, P), dtype=np.float64)
    for lt, tap in enumerate(taps):
        kh, kw = divmod(tap, plan["KW"])
        ok, pix = _gather_pixel(plan, ptab[:P], kh, kw, dgrad=False)
        cols = X[(pix[:, None] * plan["lds"] + np.arange(C)[None, :])]
        A[lt * C:(lt + 1) * C, :] = np.where(ok[:, None], cols, 0.0).T
    dY = Y[(np.arange(P)[:, None] * ldy + np.arange(N)[None, :])].astype(np.float64)
    D = alpha * (A @ dY)                                        # [Mreal, N]
    for lt, tap in enumerate(taps):
        for cc in range(min(C, Cw)):
            row[w_off + np.arange(N) * wrow + tap * Cw + cc] += D[lt * C + cc].astype(row.dtype)
    if plan.get("ones_row") and plan.get("bias_off", -1) >= 0:
        row[plan["bias_off"] + np.arange(N)] += (alpha * dY.sum(axis=0)).astype(row.dtype)


# =====================================================================================================================
# NumPy emulator of the TMA-fed kernel's GEOMETRY (conv_tma.cu): tiles of whole images / row strips, one tensor-map box per tap and
# k-block (out-of-bounds elements read as zero, ``elementStrides`` = conv stride), parity classes of strided dgrads, pixel-block
# reductions of wgrad, and the epilogue's tile-row → output-pixel map.  No swizzle / descriptors (shape independent, GPU-tested).
# =====================================================================================================================
def _box(T: np.ndarray, start: Sequence[int], box: Sequence[int], es: Sequence[int]) -> np.ndarray:
    """``cp.async.bulk.tensor`` tile mode on an N-d array given slowest-dim-first (NumPy order); ``start`` / ``box`` / ``es`` are
    innermost-first like the tensor map.  box[i] is the bounding box in tensor elements, es[i] the traversal stride; elements
    outside the tensor read as zero.  Returns the box slowest-dim-first with ceil(box / es) entries per dim."""
    nd = T.ndim
    out_shape = [-(-int(box[i]) // int(es[i])) for i in range(nd)][::-1]
    out = np.zeros(out_shape, dtype=np.float64)
    idx = [int(start[i]) + np.arange(out_shape[nd - 1 - i]) * int(es[i]) for i in range(nd)]       # innermost-first coordinates
    ok = [(ix >= 0) & (ix < T.shape[nd - 1 - i]) for i, ix in enumerate(idx)]
    sel = np.ix_(*[np.where(ok[i], idx[i], 0) for i in range(nd - 1, -1, -1)])
    vals = T[sel]
    mask = np.ones(out_shape, dtype=bool)
    for axis, i in enumerate(range(nd - 1, -1, -1)):                # axis 0 = slowest dim = innermost-first index nd-1
        shape = [1] * nd
        shape[axis] = out_shape[axis]
        mask = mask & ok[i].reshape(shape)
    out[...] = np.where(mask, vals, 0.0)
    return out


def emulate_tma(mode: int, g: ConvGeom, tp: Dict, X: np.ndarray, Yb: np.ndarray, W: np.ndarray, bn: Optional[int] = None) -> np.ndarray:
    """One group through the TMA geometry ``tp = tma_plan(mode, g, …)``.

    F: ``X`` [B, IH, IW, Cin_pad] activations, ``W`` [Cout, KH·KW·Cin] → returns Y [B, OH, OW, Cout_pad].
    D: ``X`` = dY [B, OH, OW, Cout_pad], ``W`` as above → returns dX [B, IH, IW, Cin_pad].
    W: ``X`` activations, ``Yb`` = dY [B, OH, OW, Cout_pad] → returns dW [Cout, KH·KW·Cin] (without the bias row)."""
    BNt = bn or _bn_tile(g.Cout if mode != MODE_D else g.Cin)
    s = g.stride
    if mode in (MODE_F, MODE_D):
        (_, _, _, boxA, esA, _), (_, _, _, boxB, _, _) = tp["maps"]
        C = g.Cin_pad if mode == MODE_F else g.Cout_pad
        N = g.Cout if mode == MODE_F else g.Cin
        OP = np.zeros((g.B, tp["OPH"], tp["OPW"], g.Cout_pad if mode == MODE_F else g.Cin_pad), dtype=np.float64)
        written = np.zeros((g.B, tp["OPH"], tp["OPW"]), dtype=np.int32)
        W3 = W.reshape(g.Cout, g.KH * g.KW, g.Cin)
        for cls in tp["classes"]:
            for mt in range(tp["mtiles"]):
                if tp["tpi"] == 1:
                    b0, y0 = mt * tp["Bt"], 0
                else:
                    b0, y0 = divmod(mt, tp["tpi"]); y0 *= tp["TH"]
                for n0 in range(0, N, BNt):
                    acc = np.zeros((tp["RT"], BNt), dtype=np.float64)
                    for lt, tap in enumerate(cls["taps"]):
                        for c0 in range(0, C, 32):
                            xs, ys = cls["dx"][lt], y0 * tp["ystep"] + cls["dy"][lt]
                            A = _box(X, [c0, xs, ys, b0], boxA[:4], esA[:4]).reshape(-1, 32)[: tp["RT"]]       # rows = (image, y, x)
                            if mode == MODE_F:          # box {32, BN, 1} over {wrow, Cout, S}: B[n][c] = W[n0+n][tap·Cin + c0 + c]
                                Wm = W.reshape(g.Cout, g.wrow)
                                Bt_ = _box(Wm, [tap * g.Cin + c0, n0], [32, BNt], [1, 1])                       # [BN, 32]
                            else:                       # boxes {32 cin, 1 tap, 32 cout} over {Cin, T, Cout, S}: B[n][c] = W[c0+c][tap][n0+n]
                                Bt_ = np.zeros((BNt, 32))
                                for a in range(BNt // 32):
                                    blk = _box(W3, [n0 + 32 * a, tap, c0], [32, 1, 32], [1, 1, 1])               # [cout 32, 1, cin 32]
                                    Bt_[32 * a: 32 * a + 32] = blk[:, 0, :].T
                            acc += A @ Bt_.T
                    hw = tp["RH"] * tp["RW"]
                    for r in range(tp["RT"]):
                        bi, rem = divmod(r, hw)
                        y, x = y0 + rem // tp["RW"], rem % tp["RW"]
                        b = b0 + bi
                        if y < tp["RH"] and b < g.B:
                            oy, ox = y * tp["osc"] + cls["py"], x * tp["osc"] + cls["px"]
                            ncol = min(BNt, N - n0)
                            OP[b, oy, ox, n0: n0 + ncol] += acc[r, :ncol]
                            if n0 == 0:
                                written[b, oy, ox] += 1
        expect = 1 if mode == MODE_F or not tp.get("partial") else None
        if expect is not None:
            assert (written == 1).all(), "every output pixel is produced by exactly one tile row"
        else:
            assert (written <= 1).all()              # parity classes without taps leave their pixels to the zero fill
        return OP
    # ---- W ----
    (_, _, _, boxA, esA, _), (_, _, _, boxB, esB, _) = tp["maps"]
    taps = g.live_taps()
    dW = np.zeros((g.Cout, g.KH * g.KW, g.Cin), dtype=np.float64)
    for kb in range(tp["kb_total"]):
        if tp["bpi"] == 1:
            pb0, py0 = kb * tp["bb"], 0
        else:
            pb0, py0 = divmod(kb, tp["bpi"]); py0 *= tp["bh"]
        dy_blk = _box(Yb, [0, 0, py0, pb0], [g.Cout_pad, boxB[1], boxB[2], boxB[3]], [1, 1, 1, 1]).reshape(-1, g.Cout_pad)[: tp["PK"]]
        for tap in taps:
            kh, kw = divmod(tap, g.KW)
            xw = _box(X, [0, kw - g.pad, py0 * s + kh - g.pad, pb0], [g.Cin_pad, boxA[1], boxA[2], boxA[3]], [1, esA[1], esA[2], 1])
            xw = xw.reshape(-1, g.Cin_pad)[: tp["PK"]]
            dW[:, tap, :] += (dy_blk[:, : g.Cout].T @ xw[:, : g.Cin])
    return dW.reshape(g.Cout, g.wrow)
