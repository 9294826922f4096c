#!/usr/bin/env python
"""Per-CTA phase timeline of conv_gemm launches (debug stamps of %globaltimer): where does a small launch spend its time?"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from murmura_b200.ops import selfcheck as sc  # noqa: E402

h = sc.Harness(torch.device("cuda", 0))
names = ["entry", "setup done", "1st loads issued", "loads done (t0)", "1st stage full (mma)", "accum ready", "epilogue done", "exit"]
for name, mode, G in [("rn.layer2.ds", "F", 1), ("rn.fc", "W", 1), ("rn.fc", "F", 1), ("har.fc1", "F", 1), ("rn.layer2", "D", 1), ("rn.layer1", "F", 8)]:
    case = next(c for c in sc.CASES if c[0] == name)
    launch, plan, g = sc.prepare_bench(h, case, mode, G)
    dbg = torch.zeros(4096 * 8, dtype=torch.int64, device="cuda")
    plan["dbg"] = dbg.data_ptr()
    for rep in range(3):
        dbg.zero_()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ctas = launch(); b.record()
        torch.cuda.synchronize()
    t = dbg[: ctas * 8].view(ctas, 8).cpu()
    t0 = int(t[:, 0].min())
    rel = (t - t0).float() / 1e3
    print(f"== {name} {mode} G={G} ctas={ctas} splitk={plan['splitk']}  event-time {a.elapsed_time(b) * 1e3:.1f} us (includes host launch)")
    for i, nme in enumerate(names):
        col = rel[:, i]
        print(f"   {nme:22s} min {col.min():7.2f}  median {col.median():7.2f}  max {col.max():7.2f} us")
