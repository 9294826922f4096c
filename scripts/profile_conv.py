#!/usr/bin/env python
"""Launch a few representative conv_gemm cases (2 launches each) — the target of
``ncu --set full --import-source on -k regex:conv_gemm`` captures (see profiles/ncu/)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from murmura_b200.ops import conv_plan as cp  # noqa: E402
from murmura_b200.ops import selfcheck as sc  # noqa: E402

WHAT = [("rn.layer2.ds", "F", 1), ("rn.layer1", "F", 8), ("rn.layer2", "D", 8), ("rn.layer2", "W", 8), ("rn.fc", "W", 1)]
if len(sys.argv) > 1:
    WHAT = [tuple(w.split(":")) for w in sys.argv[1:]]
    WHAT = [(a, b, int(c)) for a, b, c in WHAT]
h = sc.Harness(torch.device("cuda", 0))
for name, mode, G in WHAT:
    case = next(c for c in sc.CASES if c[0] == name)
    launch, plan, g = sc.prepare_bench(h, case, mode, G)
    for _ in range(2):
        ctas = launch()
    torch.cuda.synchronize()
    print(name, mode, G, "ctas", ctas, "splitk", plan["splitk"])
