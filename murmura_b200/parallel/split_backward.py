"""Weight gradients off the critical path.

In back-propagation the *data* gradient of a layer feeds the next layer's backward, while its *weight* gradient is
consumed only by the optimizer at the end of the step.  Autograd nevertheless issues both on one stream, so in a
launch-latency-bound step (ResNet-18 on 32×32 inputs: ≈ 10 µs kernels on an otherwise idle B200) the ≈ 20 ``wgrad``
kernels sit on the critical path for nothing.  :class:`SplitBackward` reroutes them:

* while it is active, ``F.conv2d`` / ``F.linear`` are autograd Functions whose backward computes the input gradient on the
  current stream and forks the weight/bias gradient onto a side stream (``aten.convolution_backward`` with an output mask);
* the weight gradients are *not* returned to autograd (no ``AccumulateGrad`` clone on the wrong stream) but stashed per
  parameter; :meth:`join` makes the current stream wait for the side stream and hands them to the multi-tensor SGD kernel.

Under CUDA-graph capture the fork/join become graph edges, so every replay runs the two chains concurrently.
The reference trains each node with stock autograd on one stream (``murmura/core/node.py:59-109``).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

_aten = torch.ops.aten


def _pair(v) -> List[int]:
    return [int(v), int(v)] if isinstance(v, int) else [int(a) for a in v]


class SplitBackward:
    """Context manager; one instance per training stream (it owns the side stream and the per-step stash)."""

    def __init__(self, device: torch.device, side: Optional[torch.cuda.Stream] = None):
        self.device = device
        self.side = side or torch.cuda.Stream(device=device)
        self.stash: Dict[int, torch.Tensor] = {}
        self._keep: List[Tuple[torch.Tensor, ...]] = []          # operands read by the side stream: alive until join()
        self._orig_conv2d = None
        self._orig_linear = None
        owner = self

        class _Conv2d(torch.autograd.Function):
            @staticmethod
            def forward(ctx, x, w, b, stride, padding, dilation, groups):
                ctx.save_for_backward(x, w)
                ctx.conf = (stride, padding, dilation, groups)
                ctx.bias_sizes = [int(b.shape[0])] if b is not None else None
                ctx.keys = (w.data_ptr(), b.data_ptr() if b is not None else None)
                return owner._orig_conv2d(x, w, b, stride, padding, dilation, groups)

            @staticmethod
            def backward(ctx, gy):
                x, w = ctx.saved_tensors
                stride, padding, dilation, groups = ctx.conf
                cur = torch.cuda.current_stream(owner.device)
                if ctx.needs_input_grad[1] or (ctx.bias_sizes is not None and ctx.needs_input_grad[2]):
                    owner.side.wait_stream(cur)
                    with torch.cuda.stream(owner.side):
                        _, gw, gb = _aten.convolution_backward(gy, x, w, ctx.bias_sizes, stride, padding, dilation, False, [0, 0],
                                                               groups, [False, True, ctx.bias_sizes is not None])
                        owner._put(ctx.keys[0], gw)
                        if ctx.bias_sizes is not None:
                            owner._put(ctx.keys[1], gb)
                    owner._keep.append((gy, x, w))
                gx = None
                if ctx.needs_input_grad[0]:
                    gx = _aten.convolution_backward(gy, x, w, None, stride, padding, dilation, False, [0, 0], groups,
                                                    [True, False, False])[0]
                return gx, None, None, None, None, None, None

        class _Linear(torch.autograd.Function):
            @staticmethod
            def forward(ctx, x, w, b):
                ctx.save_for_backward(x, w)
                ctx.has_bias = b is not None
                ctx.keys = (w.data_ptr(), b.data_ptr() if b is not None else None)
                return owner._orig_linear(x, w, b)

            @staticmethod
            def backward(ctx, gy):
                x, w = ctx.saved_tensors
                cur = torch.cuda.current_stream(owner.device)
                gy2 = gy.reshape(-1, gy.shape[-1])
                owner.side.wait_stream(cur)
                with torch.cuda.stream(owner.side):
                    owner._put(ctx.keys[0], gy2.t().mm(x.reshape(-1, x.shape[-1])))
                    if ctx.has_bias:
                        owner._put(ctx.keys[1], gy2.sum(0))
                owner._keep.append((gy, gy2, x))
                gx = gy.matmul(w) if ctx.needs_input_grad[0] else None
                return gx, None, None

        self._Conv2d, self._Linear = _Conv2d, _Linear

    def _put(self, key: int, grad: torch.Tensor) -> None:
        """Stash a weight gradient (called on the side stream); a parameter used by several layers accumulates."""
        prev = self.stash.get(key)
        self.stash[key] = grad if prev is None else prev + grad

    # ---- patched entry points -------------------------------------------------------------------------------------------
    def _conv2d(self, input, weight, bias=None, stride=1, padding=0, dilation=1, groups=1):
        if (isinstance(padding, str) or not input.is_cuda or not torch.is_grad_enabled() or not weight.requires_grad
                or input.dim() != 4 or torch.is_autocast_enabled()):
            return self._orig_conv2d(input, weight, bias, stride, padding, dilation, groups)
        return self._Conv2d.apply(input, weight, bias, _pair(stride), _pair(padding), _pair(dilation), int(groups))

    def _linear(self, input, weight, bias=None):
        if not input.is_cuda or not torch.is_grad_enabled() or not weight.requires_grad or torch.is_autocast_enabled():
            return self._orig_linear(input, weight, bias)
        return self._Linear.apply(input, weight, bias)

    def __enter__(self):
        self._orig_conv2d, self._orig_linear = F.conv2d, F.linear
        F.conv2d, F.linear = self._conv2d, self._linear
        return self

    def __exit__(self, *exc):
        F.conv2d, F.linear = self._orig_conv2d, self._orig_linear
        return False

    # ---- end of backward -------------------------------------------------------------------------------------------------
    def join(self, params: Sequence[torch.Tensor]) -> List[torch.Tensor]:
        """Wait for the side stream; return one gradient per parameter (stashed weight gradient or ``p.grad``), laid out
        like the parameter."""
        if self.stash:
            torch.cuda.current_stream(self.device).wait_stream(self.side)
        grads = []
        for p in params:
            g = self.stash.get(p.data_ptr())
            if g is None:
                g = p.grad
            if g is None:
                raise RuntimeError("SplitBackward.join: a parameter received no gradient")
            if g.shape != p.shape:
                g = g.reshape(p.shape)
            if g.stride() != p.stride() or g.dtype != p.dtype:
                g = torch.empty_like(p).copy_(g)
            grads.append(g)
        self.stash.clear()
        self._keep.clear()
        return grads
