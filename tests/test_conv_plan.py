"""CPU checks of the implicit-GEMM plans: the NumPy emulator (same index formulas as ``conv_tcgen05.cu``) against
``torch.nn.functional.conv2d`` and its autograd gradients — fprop, dgrad, wgrad(+SGD), live taps, padded first layers."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from murmura_b200.ops import conv_plan as cp

CASES = [
    # B, H, W, Cin, Cout, k, stride, pad
    (2, 8, 8, 8, 16, 3, 1, 1),
    (3, 4, 4, 16, 8, 3, 2, 1),
    (2, 2, 2, 8, 12, 3, 1, 1),
    (4, 1, 1, 16, 8, 3, 1, 1),       # 3×3 on a 1×1 map: only the centre tap is live
    (2, 2, 2, 8, 8, 3, 2, 1),        # stride 2, 2×2 → 1×1: four live taps
    (2, 9, 9, 3, 8, 7, 2, 3),        # first layer, channels padded 3 → 4
    (2, 6, 6, 1, 8, 5, 1, 2),        # FEMNIST-style first layer, 1 → 4
    (2, 5, 5, 4, 6, 1, 2, 0),        # 1×1 stride-2 (downsample), Cout not a multiple of 4
    (5, 1, 1, 20, 10, 1, 1, 0),      # linear layer
]


def _nhwc(x, cpad):
    b, c, h, w = x.shape
    out = np.zeros((b, h, w, cpad), dtype=np.float32)
    out[..., :c] = x.permute(0, 2, 3, 1).numpy()
    return out.ravel()


def _row_with_weights(w, bias):
    """Arena row: weights physically (Cout, KH, KW, Cin), then the bias."""
    wp = w.permute(0, 2, 3, 1).contiguous().numpy().ravel()
    row = np.concatenate([np.zeros(8, np.float32), wp, bias.numpy() if bias is not None else np.zeros(0, np.float32)])
    return row.astype(np.float32), 8, 8 + wp.size


@pytest.mark.parametrize("case", CASES)
def test_live_taps_are_exactly_the_taps_that_touch_pixels(case):
    B, H, W, Cin, Cout, k, s, p = case
    g = cp.ConvGeom(B=B, IH=H, IW=W, Cin=Cin, Cout=Cout, KH=k, KW=k, stride=s, pad=p)
    live = set(g.live_taps())
    for kh in range(k):
        for kw in range(k):
            touches = any(0 <= oh * s + kh - p < H and 0 <= ow * s + kw - p < W for oh in range(g.OH) for ow in range(g.OW))
            assert ((kh * k + kw) in live) == touches
    if (H, W, k, p) == (1, 1, 3, 1):
        assert g.live_taps() == [4]


@pytest.mark.parametrize("case", CASES)
def test_fprop_matches_conv2d(case):
    B, H, W, Cin, Cout, k, s, p = case
    torch.manual_seed(0)
    x = torch.randn(B, Cin, H, W); w = torch.randn(Cout, Cin, k, k); b = torch.randn(Cout)
    g = cp.ConvGeom(B=B, IH=H, IW=W, Cin=Cin, Cout=Cout, KH=k, KW=k, stride=s, pad=p)
    row, w_off, b_off = _row_with_weights(w, b)
    plan = cp.plan_fprop(g)
    plan.update(w_off=w_off, bias_off=b_off, relu=1)
    Y = np.zeros(plan["M"] * plan["ldy"], dtype=np.float32)
    cp.emulate(plan, _nhwc(x, g.Cin_pad), Y, row)
    ref = F.relu(F.conv2d(x, w, b, stride=s, padding=p)).permute(0, 2, 3, 1).reshape(-1, Cout).numpy()
    got = Y.reshape(plan["M"], plan["ldy"])
    np.testing.assert_allclose(got[:, :Cout], ref, rtol=1e-4, atol=1e-4)
    assert np.all(got[:, Cout:] == 0)
    assert (plan["vecB"] == 4) == (Cin % 4 == 0)


@pytest.mark.parametrize("case", [c for c in CASES if c[3] % 4 == 0])
def test_dgrad_matches_autograd(case):
    B, H, W, Cin, Cout, k, s, p = case
    torch.manual_seed(1)
    x = torch.randn(B, Cin, H, W, requires_grad=True); w = torch.randn(Cout, Cin, k, k)
    y = F.conv2d(x, w, None, stride=s, padding=p)
    dy = torch.randn_like(y)
    (dx,) = torch.autograd.grad(y, x, dy)
    g = cp.ConvGeom(B=B, IH=H, IW=W, Cin=Cin, Cout=Cout, KH=k, KW=k, stride=s, pad=p)
    row, w_off, _ = _row_with_weights(w, None)
    plan = cp.plan_dgrad(g)
    plan.update(w_off=w_off)
    dX = np.full(plan["M"] * plan["ldy"], 0.5, dtype=np.float32)
    plan["accumulate"] = 1
    cp.emulate(plan, _nhwc(dy, g.Cout_pad), dX, row)
    ref = dx.permute(0, 2, 3, 1).reshape(-1, Cin).numpy() + 0.5
    np.testing.assert_allclose(dX.reshape(plan["M"], plan["ldy"])[:, :Cin], ref, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("case", CASES)
def test_wgrad_is_the_sgd_step(case):
    B, H, W, Cin, Cout, k, s, p = case
    torch.manual_seed(2)
    x = torch.randn(B, Cin, H, W)
    w = torch.randn(Cout, Cin, k, k, requires_grad=True); b = torch.randn(Cout, requires_grad=True)
    y = F.conv2d(x, w, b, stride=s, padding=p)
    dy = torch.randn_like(y)
    dw, db = torch.autograd.grad(y, (w, b), dy)
    lr = 0.05
    g = cp.ConvGeom(B=B, IH=H, IW=W, Cin=Cin, Cout=Cout, KH=k, KW=k, stride=s, pad=p)
    row, w_off, b_off = _row_with_weights(w.detach(), b.detach())
    plan = cp.plan_wgrad(g, bias=True)
    plan.update(w_off=w_off, bias_off=b_off, alpha=-lr)
    cp.emulate(plan, _nhwc(x, g.Cin_pad), _nhwc(dy, g.Cout_pad), row)
    w_new = (w - lr * dw).detach().permute(0, 2, 3, 1).numpy().ravel()
    np.testing.assert_allclose(row[w_off:w_off + w_new.size], w_new, rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(row[b_off:b_off + Cout], (b - lr * db).detach().numpy(), rtol=1e-4, atol=1e-4)


def test_eval_epilogue_batchnorm_residual_relu():
    torch.manual_seed(3)
    B, H, W, Cin, Cout = 2, 4, 4, 8, 8
    x = torch.randn(B, Cin, H, W); w = torch.randn(Cout, Cin, 3, 3); res = torch.randn(B, Cout, H, W)
    mean, var, gam, bet = torch.randn(Cout), torch.rand(Cout) + 0.5, torch.randn(Cout), torch.randn(Cout)
    g = cp.ConvGeom(B=B, IH=H, IW=W, Cin=Cin, Cout=Cout, KH=3, KW=3, stride=1, pad=1)
    row, w_off, _ = _row_with_weights(w, None)
    base = row.size
    row = np.concatenate([row, mean.numpy(), var.numpy(), gam.numpy(), bet.numpy()]).astype(np.float32)
    plan = cp.plan_fprop(g)
    plan.update(w_off=w_off, bn_mean_off=base, bn_var_off=base + Cout, bn_gamma_off=base + 2 * Cout, bn_beta_off=base + 3 * Cout, relu=1)
    Y = np.zeros(plan["M"] * plan["ldy"], dtype=np.float32)
    cp.emulate(plan, _nhwc(x, Cin), Y, row, R=_nhwc(res, Cout))
    ref = F.relu(F.batch_norm(F.conv2d(x, w, padding=1), mean, var, gam, bet, False, 0.0, 1e-5) + res)
    np.testing.assert_allclose(Y.reshape(-1, Cout), ref.permute(0, 2, 3, 1).reshape(-1, Cout).numpy(), rtol=1e-4, atol=1e-4)


def test_split_k_choice():
    assert cp.choose_splitk(ctas=8, kb_total=72) == 18
    assert cp.choose_splitk(ctas=256, kb_total=18) == 1
    assert cp.choose_splitk(ctas=8, kb_total=4) == 1


# ---- geometry of the TMA-fed kernel (conv_tma.cu): tiles, tensor-map boxes, parity classes, pixel blocks ------------------------
def _tma_case(B, H, W, Cin, Cout, k, s, p):
    g = cp.ConvGeom(B=B, IH=H, IW=W, Cin=Cin, Cout=Cout, KH=k, KW=k, stride=s, pad=p)
    torch.manual_seed(B * 131 + H * 17 + Cin + Cout + k + s)
    x = torch.randn(B, Cin, H, W, dtype=torch.float64, requires_grad=True)
    w = torch.randn(Cout, Cin, k, k, dtype=torch.float64, requires_grad=True)
    y = F.conv2d(x, w, stride=s, padding=p)
    gy = torch.randn_like(y)
    y.backward(gy)
    Xn = x.detach().permute(0, 2, 3, 1).numpy()
    Wn = w.detach().permute(0, 2, 3, 1).reshape(Cout, -1).numpy()
    dYn = gy.permute(0, 2, 3, 1).numpy()
    refs = {cp.MODE_F: y.detach().permute(0, 2, 3, 1).numpy(), cp.MODE_D: x.grad.permute(0, 2, 3, 1).numpy(),
            cp.MODE_W: w.grad.permute(0, 2, 3, 1).reshape(Cout, -1).numpy()}
    ran = []
    for mode in (cp.MODE_F, cp.MODE_D, cp.MODE_W):
        tp = cp.tma_plan(mode, g, 1 << 20, 1)
        if tp is None:
            continue
        out = cp.emulate_tma(mode, g, tp, dYn if mode == cp.MODE_D else Xn, dYn if mode == cp.MODE_W else None, Wn)
        ref = refs[mode]
        np.testing.assert_allclose(out[..., : ref.shape[-1]], ref, rtol=1e-9, atol=1e-9)
        ran.append(mode)
    return ran


TMA_CASES = [
    (2, 8, 8, 32, 64, 3, 1, 1),        # ResNet layer1-style: one image = 64 rows, two images per tile
    (3, 4, 4, 64, 32, 3, 2, 1),        # stride 2: dgrad as four parity classes
    (4, 1, 1, 64, 32, 3, 1, 1),        # 3×3 on a 1×1 map (centre tap only)
    (2, 2, 2, 32, 32, 3, 2, 1),        # 2×2 → 1×1
    (2, 16, 16, 32, 64, 3, 1, 1),      # 256 pixels per image: strips of 8 rows
    (3, 8, 8, 32, 64, 1, 2, 0),        # 1×1 stride-2 downsample: a single parity class has taps
    (5, 1, 1, 96, 40, 1, 1, 0),        # linear layer, Cout not a multiple of 32 (dgrad falls back to the gather kernel)
    (2, 14, 14, 32, 64, 5, 1, 2),      # LEAF conv2: 5×5, 196 pixels per image (strips of 9 rows, the last one partial)
    (130, 2, 2, 32, 32, 3, 1, 1),      # more images than one tile holds, partial last tile
    (2, 16, 16, 64, 128, 3, 2, 1),     # stride-2 with strips, BN = 128
]


@pytest.mark.parametrize("case", TMA_CASES)
def test_tma_geometry_matches_conv2d_and_autograd(case):
    ran = _tma_case(*case)
    assert cp.MODE_F in ran and cp.MODE_W in ran
    if case[4] % 32 == 0:
        assert cp.MODE_D in ran


def test_tma_geometry_random_shapes():
    """Property test over random layer geometries: whenever the planner accepts a layer for the TMA kernels, its tiles / boxes /
    classes reproduce conv2d and its gradients exactly (fp64)."""
    hyp = pytest.importorskip("hypothesis")
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=40, deadline=None)
    @given(B=st.integers(1, 5), H=st.integers(1, 12), W=st.integers(1, 12), cin=st.sampled_from([32, 64]), cout=st.sampled_from([32, 64, 96]),
           k=st.sampled_from([1, 3, 5]), s=st.sampled_from([1, 2]), same=st.booleans())
    def prop(B, H, W, cin, cout, k, s, same):
        p = k // 2 if same else 0
        if (H + 2 * p - k) < 0 or (W + 2 * p - k) < 0:
            return
        _tma_case(B, H, W, cin, cout, k, s, p)

    prop()
