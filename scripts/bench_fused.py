#!/usr/bin/env python
"""Time the fused ResNet-18 SGD step (all nodes of a GPU in one program) — per launch (eager, CUDA events) and per step
(CUDA-graph replay), for 1 and 8 grouped nodes, with and without the side stream for weight gradients.

    python scripts/bench_fused.py [--out gpurun_out/bench_fused.json] [--model resnet18|femnist|har]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from murmura_b200.models import EvidentialHARClassifier, LEAFFEMNISTModel, ResNet18  # noqa: E402
from murmura_b200.parallel.arena import StateLayout  # noqa: E402
from murmura_b200.parallel.fused_trainer import FusedTrainer  # noqa: E402

MODELS = {"resnet18": (lambda: ResNet18(10), (3, 32, 32), 64, False), "femnist": (lambda: LEAFFEMNISTModel(62), (1, 28, 28), 64, False),
          "har": (lambda: EvidentialHARClassifier(), (561,), 32, True)}


def build(name, G, steps, side, target_ctas=148):
    factory, shape, batch, evid = MODELS[name]
    dev = torch.device("cuda", 0)
    probe = factory()
    layout = StateLayout.from_model(probe, channels_last=True)
    live = torch.zeros(G, layout.stride, device=dev)
    ints = torch.zeros(G, max(layout.Pi, 1), dtype=torch.int64, device=dev)
    shards, models = [], []
    for s in range(G):
        m = factory().to(dev)
        layout.bind(m, live[s], None, ints[s] if layout.Pi else None)
        models.append(m)
        n = batch * steps
        x = torch.randn(n, *shape, device=dev)
        shards.append((x.permute(0, 2, 3, 1).contiguous() if len(shape) == 3 else x, torch.randint(0, 6, (n,), device=dev)))
    tr = FusedTrainer(models[0], layout, live, ints if layout.Pi else None, shards, [steps] * G, batch, shape, evidential=evid, side_stream=side,
                      target_ctas=target_ctas)
    assert tr.supported
    tr._models = models
    return tr


def per_op(tr, G, lr=0.01):
    """Eager, one CUDA-event pair per launch (serialised: shares, not absolutes)."""
    rows = []

    def timed(label, fn):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        rows.append((label, a.elapsed_time(b) * 1e3))

    for _ in range(2):
        rows.clear()
        timed("zero_pool", lambda: tr.zero_pool[:G].zero_())
        timed("gather", lambda: tr._gather(G, 0))
        for op in tr.ops:
            timed(f"fwd {type(op).__name__} {op.name}", lambda op=op: op.fwd(tr, G))
        for op in reversed(tr.ops):
            timed(f"bwd {type(op).__name__} {op.name}", lambda op=op: op.bwd(tr, G, lr))
    return rows


def timeline(name, G, side, out_path):
    """CUPTI records of one graph replay: per-kernel start/duration, busy time vs wall time (ranking only — profiler overhead)."""
    from torch.profiler import ProfilerActivity, profile
    steps = 4
    tr = build(name, G, steps, side)
    tr.refresh_permutations(1); tr._capture(0.01)
    for _ in range(3):
        tr.graph.replay()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        tr.graph.replay(); torch.cuda.synchronize()
    evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
    t0 = evs[0].time_range.start
    wall = evs[-1].time_range.end - t0
    busy, cur_end, rows = 0.0, t0, []
    for e in evs:
        s, t = e.time_range.start, e.time_range.end
        busy += max(0.0, t - max(s, cur_end)); cur_end = max(cur_end, t)
        rows.append((round(s - t0, 2), round(t - s, 2), e.name[:70]))
    agg = {}
    for _, d, n in rows:
        a = agg.setdefault(n, [0.0, 0]); a[0] += d; a[1] += 1
    res = {"model": name, "G": G, "side": side, "steps": steps, "wall_us_per_step": wall / steps, "busy_us_per_step": busy / steps,
           "kernels_per_step": len(rows) / steps,
           "by_kernel": {n: {"us_per_step": round(v[0] / steps, 1), "n_per_step": v[1] / steps, "avg_us": round(v[0] / v[1], 2)} for n, v in sorted(agg.items(), key=lambda kv: -kv[1][0])},
           "first_step": rows[: len(rows) // steps]}
    os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
    json.dump(res, open(out_path, "w"), indent=1)
    print(json.dumps({k: v for k, v in res.items() if k not in ("first_step",)}, indent=1))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--timeline", type=int, default=0, help="G: dump the CUPTI kernel timeline of one graph replay")
    ap.add_argument("--out", default="gpurun_out/bench_fused.json")
    ap.add_argument("--model", default="resnet18")
    ap.add_argument("--target-ctas", type=int, default=148)
    ap.add_argument("--quick", action="store_true", help="graph timings only (no eager per-launch pass)")
    ap.add_argument("--pdl", type=int, default=1, help="programmatic dependent launch on / off (A/B)")
    args = ap.parse_args()
    from murmura_b200 import ops
    ops.load(required=True).set_pdl(bool(args.pdl))
    if args.timeline:
        return timeline(args.model, args.timeline, True, args.out)
    out = {"model": args.model, "runs": []}
    for G in (1, 8):
        for side in (False, True):
            steps = 8
            tr = build(args.model, G, steps, side, args.target_ctas)
            tr.refresh_permutations(1)
            tr._capture(0.01)
            for _ in range(3):
                tr.graph.replay()
            torch.cuda.synchronize()
            reps = 10
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                tr.graph.replay()
            b.record(); torch.cuda.synchronize()
            us_step = a.elapsed_time(b) * 1e3 / (reps * steps)
            run = {"G": G, "side_stream": side, "us_per_step_graph": round(us_step, 1), "launches_per_step": len(tr.ops) * 2 + 1,
                   "workspace_mb": round(tr.workspace_bytes / 2 ** 20, 1), "finite": bool(torch.isfinite(tr.live).all())}
            if not side and not args.quick:
                rows = per_op(tr, G)
                run["eager_sum_us"] = round(sum(t for _, t in rows), 1)
                agg = {}
                for label, t in rows:
                    k = " ".join(label.split()[:2])
                    agg[k] = agg.get(k, 0.0) + t
                run["by_kind_us"] = {k: round(v, 1) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])}
                run["top_launches"] = [(l, round(t, 1)) for l, t in sorted(rows, key=lambda r: -r[1])[:12]]
            out["runs"].append(run)
            print(json.dumps(run), flush=True)
            del tr
            torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    json.dump(out, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
