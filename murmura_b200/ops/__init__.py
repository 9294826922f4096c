"""sm_100a CUDA extension: build, load and thin Python wrappers.

The extension is compiled **in-tree** (``murmura_b200/ops/_build/murmura_b200_ext.so``) with
``nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo`` so the binary travels with the repo to
GPU boxes.  ``available()`` is True only when the ``.so`` is loaded *and* a CUDA device exists; on a
GPU machine a missing extension is a hard error for the B200 engine (no silent eager fallback).

Every wrapper has a plain-PyTorch oracle in :mod:`murmura_b200.ops.reference` used by the tests.
"""
from __future__ import annotations

import hashlib
import importlib.machinery
import importlib.util
import os
import sys
from pathlib import Path
from typing import Optional

import torch

_HERE = Path(__file__).resolve().parent
_SRC = _HERE / "csrc"
_BUILD = _HERE / "_build"
_NAME = "murmura_b200_ext"
_SOURCES = ["bindings.cpp", "arena.cu", "aggregate.cu", "train.cu", "gram_tcgen05.cu", "mlp_tcgen05.cu", "dmtt.cu", "bn_train.cu", "conv_tcgen05.cu", "conv_tma.cu", "layers.cu"]
_CUDA_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "--use_fast_math",
               "-std=c++17", "--expt-relaxed-constexpr", "-Xptxas", "-v"]

_ext = None
_load_error: Optional[str] = None
counters = {"launches": 0}          # launches of our kernels issued through the Python wrappers below (engine's launch accounting)


def _source_digest() -> str:
    h = hashlib.sha256()
    for name in sorted(os.listdir(_SRC)):
        h.update(name.encode())
        h.update((_SRC / name).read_bytes())
    h.update(" ".join(_CUDA_FLAGS).encode())
    h.update(torch.__version__.encode())
    return h.hexdigest()[:16]


def _so_path() -> Path:
    return _BUILD / f"{_NAME}.so"


def _import_so(path: Path):
    loader = importlib.machinery.ExtensionFileLoader(_NAME, str(path))
    spec = importlib.util.spec_from_loader(_NAME, loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    sys.modules[_NAME] = mod
    return mod


def build_extension(verbose: bool = False, force: bool = False) -> str:
    """Compile the extension for sm_100a (cross-compiles without a GPU). Returns the .so path."""
    from torch.utils import cpp_extension
    _BUILD.mkdir(exist_ok=True)
    stamp = _BUILD / "stamp.txt"
    digest = _source_digest()
    if not force and _so_path().exists() and stamp.exists() and stamp.read_text().strip() == digest:
        return str(_so_path())
    os.environ.setdefault("MAX_JOBS", str(min(8, os.cpu_count() or 4)))
    cpp_extension.load(
        name=_NAME, sources=[str(_SRC / s) for s in _SOURCES], build_directory=str(_BUILD),
        extra_cuda_cflags=_CUDA_FLAGS, extra_cflags=["-O3", "-std=c++17"], verbose=verbose, is_python_module=False)
    stamp.write_text(digest)
    return str(_so_path())


def load(required: bool = False):
    """Load the prebuilt ``.so`` (building it first if sources changed and nvcc is present)."""
    global _ext, _load_error
    if _ext is not None:
        return _ext
    try:
        stamp = _BUILD / "stamp.txt"
        fresh = _so_path().exists() and stamp.exists() and stamp.read_text().strip() == _source_digest()
        if not fresh:
            build_extension()
        _ext = _import_so(_so_path())
    except Exception as exc:  # noqa: BLE001 - surfaced through required / load_error()
        _load_error = f"{type(exc).__name__}: {exc}"
        if required:
            raise RuntimeError(f"murmura_b200 CUDA extension unavailable: {_load_error}") from exc
    return _ext


def load_error() -> Optional[str]:
    return _load_error


def available() -> bool:
    """True when kernels can actually run (CUDA device present and extension loaded)."""
    return torch.cuda.is_available() and load() is not None


def ext():
    """The loaded extension module; raises loudly when it is missing."""
    mod = load(required=True)
    return mod


# ---- autograd wrapper: fused evidential loss -------------------------------------------------------

class _EvidentialLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, alpha: torch.Tensor, targets: torch.Tensor, lam):
        lam_t = lam if isinstance(lam, torch.Tensor) else None
        loss, grad = ext().evidential_loss_fwd_bwd(alpha.contiguous().float(), targets.contiguous(),
                                                   0.0 if lam_t is not None else float(lam), lam_t)
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None


def evidential_loss(alpha: torch.Tensor, targets: torch.Tensor, lam) -> torch.Tensor:
    """Fused forward+backward evidential loss (``train.cu::evidential_loss_kernel``).

    ``lam`` is a Python float or a 0-dim CUDA tensor (read on the device, so a CUDA graph that
    captured the loss follows the annealing schedule without re-capture)."""
    return _EvidentialLossFn.apply(alpha, targets, lam)


class _CELossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits: torch.Tensor, targets: torch.Tensor, loss_acc):
        loss, grad = ext().ce_loss_fwd_bwd(logits.contiguous().float(), targets.contiguous(), loss_acc)
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None


def ce_loss(logits: torch.Tensor, targets: torch.Tensor, loss_acc: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Mean softmax cross-entropy, forward+backward in one launch (``train.cu::ce_loss_kernel``); ``loss_acc`` (0-dim fp32
    CUDA tensor) is incremented by the loss inside the kernel."""
    return _CELossFn.apply(logits, targets, loss_acc)


# ---- fused BatchNorm (+residual) (+ReLU): training fwd/bwd in one launch each ------------------------

class _BNActFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, nbt, residual, momentum, eps, relu):
        y, mean, invstd = ext().bn_act_fwd(x, residual, weight, bias, running_mean, running_var, nbt, float(momentum), float(eps), bool(relu))
        counters["launches"] += 1
        ctx.save_for_backward(x, y, weight, mean, invstd)
        ctx.relu, ctx.has_res = bool(relu), residual is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, weight, mean, invstd = ctx.saved_tensors
        fmt = torch.channels_last if x.dim() == 4 else torch.contiguous_format
        counters["launches"] += 1
        dx, dw, db, dres = ext().bn_act_bwd(dy.contiguous(memory_format=fmt), x, y, weight, mean, invstd, ctx.relu,
                                            ctx.has_res and ctx.needs_input_grad[6])
        return dx, dw, db, None, None, None, (dres if ctx.has_res else None), None, None, None


_fused_bn = os.environ.get("MURMURA_B200_FUSED_BN", "1") != "0"


def set_fused_bn(enabled: bool) -> None:
    """Globally enable/disable the fused BatchNorm path of :func:`bn_act` (``b200.fused_bn``)."""
    global _fused_bn
    _fused_bn = bool(enabled)


def _bn_layout_ok(x: torch.Tensor) -> bool:
    if x.dim() == 2:
        return x.is_contiguous()
    return x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last)


def bn_act_fusable(x: torch.Tensor, bn) -> bool:
    """True when ``bn_act`` will take the fused sm_100a path for this input/module."""
    return (_fused_bn and x.is_cuda and x.dtype == torch.float32 and bn.affine and bn.track_running_stats and bn.momentum is not None
            and x.dim() in (2, 4) and x.shape[1] % 4 == 0 and _bn_layout_ok(x) and available()
            and (not bn.training or x.numel() // x.shape[1] >= 2))


def bn_act(x: torch.Tensor, bn, residual: Optional[torch.Tensor] = None, relu: bool = True) -> torch.Tensor:
    """``relu(bn(x) + residual)`` of an ``nn.BatchNorm{1,2}d`` module ``bn``.

    On CUDA (fp32, [B,C] or channels_last activations) the training forward and backward are ONE launch each
    (``bn_train.cu``: thread-block-cluster reduction through distributed shared memory; running statistics and
    ``num_batches_tracked`` updated in the same kernel) and evaluation is one ``bn_eval`` launch; anything else falls back
    to the stock PyTorch ops with identical semantics."""
    if bn_act_fusable(x, bn):
        if residual is not None and (residual.dtype != x.dtype or residual.stride() != x.stride()):
            residual = torch.empty_like(x).copy_(residual)
        if bn.training:
            return _BNActFn.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.num_batches_tracked, residual,
                                  bn.momentum, bn.eps, relu)
        if not torch.is_grad_enabled():
            counters["launches"] += 1
            return ext().bn_eval(x, bn.running_mean, bn.running_var, bn.weight, bn.bias, float(bn.eps), bool(relu), residual)
    import torch.nn.functional as F
    out = torch.nn.modules.batchnorm._BatchNorm.forward(bn, x)
    if residual is not None:
        out = out + residual
    return F.relu(out) if relu else out


class fast_eval_batchnorm:
    """Context manager: route inference-mode ``F.batch_norm`` on fp32 CUDA tensors to ``train.cu::bn_eval_kernel``.

    Used while the engine runs / captures evaluation forwards (CUDA graphs bake the replacement in).  Training-mode
    batch norm (batch statistics + running-stat updates) is untouched and stays on cuDNN.
    """

    def __enter__(self):
        import torch.nn.functional as F
        self._F, self._orig = F, F.batch_norm
        orig, e = self._orig, ext()

        def batch_norm(input, running_mean, running_var, weight=None, bias=None, training=False, momentum=0.1, eps=1e-5):
            if (not training and running_mean is not None and input.is_cuda and input.dtype == torch.float32
                    and not torch.is_grad_enabled()):
                return e.bn_eval(input, running_mean, running_var, weight, bias, float(eps), False)
            return orig(input, running_mean, running_var, weight, bias, training, momentum, eps)

        F.batch_norm = batch_norm
        return self

    def __exit__(self, *exc):
        self._F.batch_norm = self._orig
        return False


__all__ = ["available", "build_extension", "load", "load_error", "ext", "evidential_loss", "ce_loss", "fast_eval_batchnorm", "bn_act", "bn_act_fusable", "set_fused_bn"]
