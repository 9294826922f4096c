"""Launch glue for the TMA-fed conv kernels: tensor-map encoding from the geometry specs of ``conv_plan.tma_plan``."""
from __future__ import annotations

from typing import Dict, Optional, Tuple

from murmura_b200.ops import conv_plan as cp


def encode_tma(ext, mode: int, geom: cp.ConvGeom, *, x_ptr: int, x_gs: int, y_ptr: int, y_gs: int, w_ptr: int, arena_stride: int,
               slots: int, groups: int) -> Optional[Dict]:
    """Launch fields (``mapA``, ``mapB`` + tile geometry) of one layer/mode over concrete buffers, or ``None`` when the layer
    has to take the cp.async path.  ``x`` is the gather source of the mode (X for F / W, dY for D), ``y`` the dY operand of W."""
    tp = cp.tma_plan(mode, geom, arena_stride, slots)
    if tp is None:
        return None
    out = {k: v for k, v in tp.items() if k != "maps"}
    for which, (tag, dims, strides, box, es, swz) in zip(("mapA", "mapB"), tp["maps"]):
        if tag == "W":
            ptr = w_ptr
        else:
            ptr, gs = (x_ptr, x_gs) if tag == "X" else (y_ptr, y_gs)
            dims = dims[:4] + [groups]
            strides = strides[:3] + [gs * 4]
        if ptr % 16 or any(s % 16 for s in strides):
            return None
        out[which] = ext.tma_encode(int(ptr), [int(v) for v in dims], [int(v) for v in strides], [int(v) for v in box], [int(v) for v in es], int(swz))
    return out
