"""Tiny driver for `ncu -k regex:bn_act`: one fused BatchNorm+residual+ReLU forward and backward on a layer1-sized tensor."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn
from murmura_b200 import ops
bn = nn.BatchNorm2d(64).cuda()
x = torch.randn(64, 64, 8, 8, device="cuda").contiguous(memory_format=torch.channels_last).requires_grad_(True)
r = torch.randn_like(x).requires_grad_(True)
for _ in range(3):
    y = ops.bn_act(x, bn, residual=r, relu=True); y.backward(torch.ones_like(y)); x.grad = None; r.grad = None
torch.cuda.synchronize()
