#!/usr/bin/env python
"""Two eager fused SGD steps (ResNet-18 by default) — meant to be run under
``ncu --metrics gpu__time_duration.sum --clock-control none --csv`` to list every launch of a step with its device time."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from bench_fused import build  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--G", type=int, default=1)
ap.add_argument("--model", default="resnet18")
ap.add_argument("--steps", type=int, default=2)
args = ap.parse_args()
tr = build(args.model, args.G, args.steps, side=False)
tr.refresh_permutations(1)
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("fused_steps")
tr.run_steps(0.01)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print("done", bool(torch.isfinite(tr.live).all()))
