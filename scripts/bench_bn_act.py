"""Fused BatchNorm(+residual)(+ReLU) kernels vs the stock cuDNN + element-wise composition.

    python scripts/bench_bn_act.py            # per-layer fwd+bwd and a whole ResNet-18 SGD step, both inside CUDA graphs
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn, torch.nn.functional as F
from murmura_b200 import ops
from murmura_b200.models.resnet import ResNet18

dev = torch.device("cuda")
torch.backends.cudnn.benchmark = True
ext = ops.ext()


def graph_time(fn, iters=50, inner=10):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(inner):
            fn()
    for _ in range(3):
        g.replay()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(iters):
        g.replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters / inner * 1e3          # µs per call


out = []
for shape in [(64, 64, 16, 16), (64, 64, 8, 8), (64, 128, 4, 4), (64, 256, 2, 2), (64, 512, 1, 1)]:
    bn = nn.BatchNorm2d(shape[1]).to(dev)
    x = torch.randn(*shape, device=dev).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    r = torch.randn(*shape, device=dev).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    g = torch.randn(*shape, device=dev).contiguous(memory_format=torch.channels_last)

    def stock():
        y = F.relu(bn(x) + r); y.backward(g); x.grad = None; r.grad = None; bn.weight.grad = None; bn.bias.grad = None

    def fused():
        y = ops.bn_act(x, bn, residual=r, relu=True); y.backward(g); x.grad = None; r.grad = None; bn.weight.grad = None; bn.bias.grad = None

    t0, t1 = graph_time(stock), graph_time(fused)
    rec = {"layer": f"bn+add+relu fwd+bwd {shape}", "stock_us": round(t0, 2), "fused_us": round(t1, 2), "speedup": round(t0 / t1, 2)}
    out.append(rec); print(json.dumps(rec), flush=True)

# whole ResNet-18 SGD step (batch 64, 32x32), one stream, CUDA graph of 4 steps
from murmura_b200.parallel.split_backward import SplitBackward
from contextlib import nullcontext
for fused_on, split_on in ((False, False), (True, False), (True, True)):
    ops.set_fused_bn(fused_on)
    sb = SplitBackward(dev) if split_on else None
    torch.manual_seed(0)
    m = ResNet18().to(dev).to(memory_format=torch.channels_last).train()
    params = [p for p in m.parameters()]
    X = torch.randn(64, 3, 32, 32, device=dev).contiguous(memory_format=torch.channels_last); Y = torch.randint(0, 10, (64,), device=dev)

    def step():
        for p in params:
            p.grad = None
        with (sb if sb is not None else nullcontext()):
            out = m(X)
        F.cross_entropy(out, Y).backward()
        grads = sb.join(params) if sb is not None else [p.grad for p in params]
        ext.sgd_multi(params, grads, 0.01)

    us = graph_time(step, iters=20, inner=4)
    rec = {"resnet18_sgd_step_batch64": ("fused_bn" if fused_on else "stock") + ("+split_backward" if split_on else ""), "us_per_step": round(us, 1)}
    out.append(rec); print(json.dumps(rec), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/bn_act_bench.json", "w"), indent=1)
