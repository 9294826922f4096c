"""On-device self-check + micro-benchmark of the grouped tcgen05 conv / linear kernels (``conv_tcgen05.cu``).

    python -m murmura_b200.ops.selfcheck [--modes F,D,W] [--bench] [--json out.json]

Every case runs the kernel on random data and compares it with exact-fp32 PyTorch (``conv2d`` and its autograd
gradients, TF32 disabled); the tolerance is the TF32 operand rounding (2⁻¹¹ per product).  ``--bench`` times the ResNet-18 /
LEAF / MLP layer shapes with CUDA events for 1 and 8 grouped nodes.  Used by ``tests/test_conv_gpu.py`` and to produce
``profiles/conv_selfcheck.json``.
"""
from __future__ import annotations

import argparse
import json
import sys
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from murmura_b200.ops import conv_plan as cp

# name, B, H, W, Cin, Cout, k, stride, pad
CASES: List[Tuple] = [
    ("tiny3x3", 2, 8, 8, 8, 16, 3, 1, 1),
    ("stride2", 3, 4, 4, 16, 8, 3, 2, 1),
    ("map1x1", 4, 1, 1, 16, 8, 3, 1, 1),
    ("first7x7", 2, 9, 9, 3, 8, 7, 2, 3),
    ("first5x5", 2, 6, 6, 1, 8, 5, 1, 2),
    ("ds1x1s2", 2, 5, 5, 4, 6, 1, 2, 0),
    ("linear", 5, 1, 1, 20, 10, 1, 1, 0),
    ("rn.conv1", 64, 32, 32, 3, 64, 7, 2, 3),
    ("rn.layer1", 64, 8, 8, 64, 64, 3, 1, 1),
    ("rn.layer2.0", 64, 8, 8, 64, 128, 3, 2, 1),
    ("rn.layer2.ds", 64, 8, 8, 64, 128, 1, 2, 0),
    ("rn.layer2", 64, 4, 4, 128, 128, 3, 1, 1),
    ("rn.layer3", 64, 2, 2, 256, 256, 3, 1, 1),
    ("rn.layer4.0", 64, 2, 2, 256, 512, 3, 2, 1),
    ("rn.layer4", 64, 1, 1, 512, 512, 3, 1, 1),
    ("rn.fc", 64, 1, 1, 512, 10, 1, 1, 0),
    ("leaf.conv2", 32, 14, 14, 32, 64, 5, 1, 2),
    ("leaf.fc1", 32, 1, 1, 3136, 2048, 1, 1, 0),
    ("har.fc0", 32, 1, 1, 561, 256, 1, 1, 0),
    ("har.fc1", 32, 1, 1, 256, 128, 1, 1, 0),
]
BENCH = [c for c in CASES if c[0].startswith(("rn.", "leaf.", "har."))]


def _nhwc(x: torch.Tensor, cpad: int) -> torch.Tensor:
    b, c, h, w = x.shape
    out = torch.zeros(b, h, w, cpad, device=x.device)
    out[..., :c] = x.permute(0, 2, 3, 1)
    return out.contiguous()


class Harness:
    def __init__(self, device: torch.device, tma: bool = False):
        from murmura_b200 import ops
        self.ext = ops.ext()
        self.dev = device
        self.ones = torch.ones(16, device=device)
        self.ptabs: Dict = {}
        self.tma = tma
        self._tma: Dict = {}
        self.tma_used = 0

    def ptab(self, shape):
        if shape not in self.ptabs:
            self.ptabs[shape] = torch.from_numpy(cp.pixel_table(*shape)).to(self.dev)
        return self.ptabs[shape]

    def launch(self, plan: Dict, G: int, X: torch.Tensor, Y: torch.Tensor, arena: torch.Tensor, gmap: Optional[torch.Tensor] = None,
               R: Optional[torch.Tensor] = None, geom: Optional[cp.ConvGeom] = None, **kw) -> int:
        if self.tma and geom is not None:
            from murmura_b200.ops.conv_launch import encode_tma
            key = (geom, plan["mode"], X.data_ptr(), Y.data_ptr(), arena.data_ptr())
            if key not in self._tma:
                self._tma[key] = encode_tma(self.ext, plan["mode"], geom, x_ptr=X.data_ptr(), x_gs=X.stride(0), y_ptr=Y.data_ptr(), y_gs=Y.stride(0),
                                            w_ptr=arena.data_ptr() + plan["w_off"] * 4, arena_stride=arena.stride(0), slots=arena.shape[0],
                                            groups=X.shape[0])
            extra = self._tma[key]
            if extra is not None:
                d = {k: v for k, v in plan.items() if k != "ptab_shape"}
                d.update(G=G, X=X.data_ptr(), x_gs=X.stride(0), Y=Y.data_ptr(), y_gs=Y.stride(0), arena=arena.data_ptr(), arena_gs=arena.stride(0),
                         ones=self.ones.data_ptr())
                if gmap is not None:
                    d["gmap"] = gmap.data_ptr()
                if R is not None:
                    d.update(R=R.data_ptr(), r_gs=R.stride(0))
                d.update(kw); d.update(extra)
                self.tma_used += 1
                return int(self.ext.conv_tma(d))
        d = {k: v for k, v in plan.items() if k != "ptab_shape"}
        d.update(G=G, X=X.data_ptr(), x_gs=X.stride(0) if X.dim() > 1 else 0, Y=Y.data_ptr(), y_gs=Y.stride(0) if Y.dim() > 1 else 0,
                 arena=arena.data_ptr(), arena_gs=arena.stride(0), ptab=self.ptab(plan["ptab_shape"]).data_ptr(), ones=self.ones.data_ptr())
        if gmap is not None:
            d["gmap"] = gmap.data_ptr()
        if R is not None:
            d.update(R=R.data_ptr(), r_gs=R.stride(0))
        d.update(kw)
        return int(self.ext.conv_gemm(d))


def make_case(case: Tuple, G: int, dev: torch.device, seed: int = 0):
    name, B, H, W, Cin, Cout, k, s, p = case
    g = cp.ConvGeom(B=B, IH=H, IW=W, Cin=Cin, Cout=Cout, KH=k, KW=k, stride=s, pad=p)
    gen = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(G, B, Cin, H, W, generator=gen).to(dev)
    w = (torch.randn(G, Cout, Cin, k, k, generator=gen) / (Cin * k * k) ** 0.5).to(dev)
    b = torch.randn(G, Cout, generator=gen).to(dev)
    dy = torch.randn(G, B, Cout, g.OH, g.OW, generator=gen).to(dev)
    wsz = Cout * k * k * Cin
    w_off, b_off = 8, 8 + (wsz + 3) // 4 * 4
    stride = (b_off + Cout + 255) // 256 * 256
    arena = torch.zeros(G, stride, device=dev)
    arena[:, w_off:w_off + wsz] = w.permute(0, 1, 3, 4, 2).reshape(G, -1)
    arena[:, b_off:b_off + Cout] = b
    return g, x, w, b, dy, arena, w_off, b_off


def rel_err(got: torch.Tensor, ref: torch.Tensor) -> float:
    return float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-6))


def check_case(h: Harness, case: Tuple, mode: str, G: int = 1, splitk: int = 1, mn_swap: int = 0, perm_groups: bool = False) -> Dict:
    dev = h.dev
    g, x, w, b, dy, arena, w_off, b_off = make_case(case, G, dev)
    name, B, H, W, Cin, Cout, k, s, p = case
    gmap = None
    order = list(range(G))
    if perm_groups and G > 1:
        order = order[::-1]
        gmap = torch.tensor(order, dtype=torch.int32, device=dev)
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
    try:
        X = torch.stack([_nhwc(x[i], g.Cin_pad) for i in range(G)]).reshape(G, -1)
        DY = torch.stack([_nhwc(dy[i], g.Cout_pad) for i in range(G)]).reshape(G, -1)
        if mode == "F":
            plan = cp.plan_fprop(g); plan.update(w_off=w_off, bias_off=b_off, relu=1 if splitk == 1 else 0, splitk=splitk)
            Y = torch.zeros(G, plan["M"] * plan["ldy"], device=dev)
            ctas = h.launch(plan, G, X, Y, arena, gmap, geom=g)
            ref = torch.stack([F.conv2d(x[i], w[order[i]], b[order[i]], stride=s, padding=p) for i in range(G)])
            if splitk == 1:
                ref = F.relu(ref)
            got = Y.view(G, B, g.OH, g.OW, g.Cout_pad)[..., :Cout].permute(0, 1, 4, 2, 3)
            pad_ok = bool((Y.view(G, -1, g.Cout_pad)[..., Cout:] == 0).all())
        elif mode == "D":
            if Cin != g.Cin_pad:
                return {"case": name, "mode": mode, "skipped": "first layer"}
            plan = cp.plan_dgrad(g); plan.update(w_off=w_off, splitk=splitk, accumulate=1, mn_swap=mn_swap)
            Y = torch.full((G, plan["M"] * plan["ldy"]), 0.25, device=dev)
            ctas = h.launch(plan, G, DY, Y, arena, gmap, geom=g)
            xs = x.clone().requires_grad_(True)
            ref = torch.stack([torch.autograd.grad(F.conv2d(xs[i], w[order[i]], None, stride=s, padding=p), xs, dy[i])[0][i] for i in range(G)]) + 0.25
            got = Y.view(G, B, H, W, g.Cin_pad)[..., :Cin].permute(0, 1, 4, 2, 3)
            pad_ok = True
        else:
            plan = cp.plan_wgrad(g, bias=True); plan.update(w_off=w_off, bias_off=b_off, alpha=-1.0, splitk=splitk, mn_swap=mn_swap)
            before = arena.clone()
            ctas = h.launch(plan, G, X, DY, arena, gmap, geom=g)
            ws = w.clone().requires_grad_(True); bs = b.clone().requires_grad_(True)
            dws, dbs = [], []
            for i in range(G):
                dw, db = torch.autograd.grad(F.conv2d(x[i], ws[order[i]], bs[order[i]], stride=s, padding=p), (ws, bs), dy[i])
                dws.append(dw[order[i]]); dbs.append(db[order[i]])
            wsz = Cout * k * k * Cin
            delta = arena - before
            got_w = torch.stack([delta[order[i], w_off:w_off + wsz].view(Cout, k, k, Cin).permute(0, 3, 1, 2) for i in range(G)])
            got_b = torch.stack([delta[order[i], b_off:b_off + Cout] for i in range(G)])
            ref, got = -torch.stack(dws), got_w
            eb = rel_err(got_b, -torch.stack(dbs))
            touched = torch.zeros_like(arena, dtype=torch.bool); touched[:, w_off:w_off + wsz] = True; touched[:, b_off:b_off + Cout] = True
            pad_ok = bool((delta[~touched] == 0).all()) and eb < 5e-3
        torch.cuda.synchronize()
        err = rel_err(got, ref)
        out = {"case": name, "mode": mode, "G": G, "splitk": splitk, "ctas": ctas, "tma": h.tma_used, "rel_err": err, "ok": bool(err < 5e-3 and pad_ok), "pad_ok": pad_ok,
               "taps": len(plan["taps"]), "K": plan["K"], "vecB": plan["vecB"]}
        if not out["ok"]:
            diff = (got - ref).abs()
            bad = (diff > 5e-3 * ref.abs().max()).float()
            out["bad_frac"] = float(bad.mean())
            out["got_absmax"] = float(got.abs().max()); out["ref_absmax"] = float(ref.abs().max())
            nz = bad.nonzero()
            out["first_bad"] = nz[:4].tolist()
        return out
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def prepare_bench(h: Harness, case: Tuple, mode: str, G: int, splitk: Optional[int] = None):
    """(launch closure, plan, geometry) of a benchmark launch of ``case`` in ``mode`` for ``G`` grouped nodes."""
    dev = h.dev
    g, x, w, b, dy, arena, w_off, b_off = make_case(case, G, dev)
    name, B, H, W, Cin, Cout, k, s, p = case
    X = torch.stack([_nhwc(x[i], g.Cin_pad) for i in range(G)]).reshape(G, -1)
    DY = torch.stack([_nhwc(dy[i], g.Cout_pad) for i in range(G)]).reshape(G, -1)
    if mode == "F":
        plan = cp.plan_fprop(g); plan.update(w_off=w_off); src, dst = X, torch.zeros(G, plan["M"] * plan["ldy"], device=dev)
    elif mode == "D":
        if Cin != g.Cin_pad:
            return None, None, g
        plan = cp.plan_dgrad(g); plan.update(w_off=w_off); src, dst = DY, torch.zeros(G, plan["M"] * plan["ldy"], device=dev)
    else:
        plan = cp.plan_wgrad(g, bias=True); plan.update(w_off=w_off, bias_off=b_off, alpha=-1e-6); src, dst = X, DY
    gx, gy, _ = cp.grid_of(plan, 1)
    split = cp.choose_splitk(gx * gy * G, cp.kb_total(plan), 148, 2 if mode == "W" else 4) if splitk is None else splitk
    plan["splitk"] = split
    if split > 1 and mode != "W":
        plan["accumulate"] = 1
    return (lambda: h.launch(plan, G, src, dst, arena, geom=g)), plan, g


def bench_case(h: Harness, case: Tuple, mode: str, G: int, iters: int = 30, splitk: Optional[int] = None) -> Dict:
    dev = h.dev
    name, B, H, W, Cin, Cout, k, s, p = case
    launch, plan, g = prepare_bench(h, case, mode, G, splitk)
    if launch is None:
        return {}
    split = plan["splitk"]
    for _ in range(3):
        ctas = launch()
    torch.cuda.synchronize()
    # the Python launch path costs more than these kernels run: time a CUDA graph of `reps` back-to-back launches
    reps = 20
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        for _ in range(reps):
            launch()
    graph.replay(); torch.cuda.synchronize()
    ts = []
    for _ in range(iters // 3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); graph.replay(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3 / reps)
    ts.sort()
    flops = 2.0 * G * B * g.OH * g.OW * Cout * len(plan["taps"]) * Cin
    us = ts[len(ts) // 2]
    return {"case": name, "mode": mode, "G": G, "tma": h.tma_used > 0, "us": round(us, 2), "us_min": round(ts[0], 2), "ctas": ctas, "splitk": split,
            "tflops": round(flops / us / 1e6, 1), "kb": cp.kb_total(plan)}


def main(argv=None) -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--modes", default="F,D,W")
    ap.add_argument("--bench", action="store_true")
    ap.add_argument("--json", default=None)
    ap.add_argument("--mn-swap", type=int, default=0)
    ap.add_argument("--tma", action="store_true", help="use the TMA-fed kernel wherever the geometry allows it")
    ap.add_argument("--cases", default=None, help="comma separated case-name prefixes")
    args = ap.parse_args(argv)
    dev = torch.device("cuda", 0)
    h = Harness(dev, tma=args.tma)
    results, bench = [], []
    cases = [c for c in CASES if args.cases is None or c[0].startswith(tuple(args.cases.split(",")))]
    for mode in args.modes.split(","):
        for case in cases:
            variants = [(1, 1, False)]
            if case[0] in ("tiny3x3", "rn.layer3", "rn.layer2", "har.fc1"):
                variants += [(3, 1, True), (1, 3, False), (2, 2, True)]
            for G, split, permg in variants:
                try:
                    r = check_case(h, case, mode, G=G, splitk=split, mn_swap=args.mn_swap, perm_groups=permg)
                except Exception as exc:  # noqa: BLE001 - report and stop: a trapped kernel poisons the context
                    r = {"case": case[0], "mode": mode, "G": G, "splitk": split, "ok": False, "exception": f"{type(exc).__name__}: {exc}"[:300]}
                    results.append(r); print(json.dumps(r), flush=True)
                    if args.json:
                        json.dump({"results": results, "bench": bench}, open(args.json, "w"), indent=1)
                    return 2
                results.append(r)
                print(json.dumps(r), flush=True)
    if args.bench:
        for mode in args.modes.split(","):
            for case in BENCH:
                for G in (1, 8):
                    r = bench_case(h, case, mode, G)
                    if r:
                        bench.append(r); print(json.dumps(r), flush=True)
    if args.json:
        json.dump({"results": results, "bench": bench}, open(args.json, "w"), indent=1)
    bad = [r for r in results if not r.get("ok", True) and "skipped" not in r]
    print(f"[selfcheck] {len(results) - len(bad)}/{len(results)} ok" + (f"; FAILED: {[(r['case'], r['mode'], r.get('G'), r.get('splitk')) for r in bad]}" if bad else ""))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
