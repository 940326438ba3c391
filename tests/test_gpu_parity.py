"""Parity tests proper: the sm_100a kernels, called through the C-ABI, against
  (1) the per-kernel CPU contracts (oracle/kernel_contracts.py) on seeded inputs,
  (2) the golden fixtures produced by the UNMODIFIED reference (tests/golden),
  (3) size-independent properties at the BASELINE.json full shapes (linearity / adjointness / closure).
Tolerance for the fp32 path: <= 1e-4 relative to the tensor's max-norm (north star), stated per assert.
Run on the B200 box:  python -m pytest tests -m gpu
"""
import math

import os

import pytest
import torch

import gan_lab_b200._growth as growth
import gan_lab_b200._kernels as K
from gan_lab_b200.utils.latent_utils import set_random_source
from oracle import kernel_contracts as KC

import parity_cases as PC

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 1e-4


@pytest.fixture(autouse=True)
def small_fmaps(monkeypatch):
    monkeypatch.setattr(growth, "FMAP_MAX", 32)
    K.set_conv_impl("fp32")
    yield
    set_random_source(None)


def rel(a, b):
    a = a.detach().float().cpu()
    b = b.detach().float().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-20))


def rn(*shape, seed=0):
    g = torch.Generator().manual_seed(seed + sum(shape))
    return torch.randn(*shape, generator=g)


def cl(t):
    return t.contiguous(memory_format=torch.channels_last)


def both(name, *args, tol=TOL, **kw):
    """Run launcher `name` on the GPU and its CPU contract on the same inputs; compare every output."""
    gargs = [a.to(DEV) if torch.is_tensor(a) else a for a in args]
    out_g = getattr(K, name)(*gargs, **kw)
    out_c = getattr(KC, name)(*[a.double() if torch.is_tensor(a) else a for a in args], **kw)
    if torch.is_tensor(out_g):
        out_g, out_c = (out_g,), (out_c,)
    for i, (g, c) in enumerate(zip(out_g, out_c)):
        if c is None or g is None:
            assert c is None and g is None
            continue
        assert tuple(g.shape) == tuple(c.shape), (name, i, g.shape, c.shape)
        assert rel(g, c) < tol, (name, i, rel(g, c))
    return out_g


# ------------------------------------------------------------------------------------------ kernels vs contracts
CONV_SHAPES = [  # N, H, W, Ci, Co, R, pad
    (8, 4, 4, 32, 32, 3, 1), (2, 16, 16, 64, 32, 3, 1), (4, 8, 8, 33, 32, 3, 1), (8, 4, 4, 32, 48, 4, 0),
    (3, 5, 7, 20, 24, 3, 1), (2, 8, 8, 16, 16, 1, 0), (1, 32, 32, 128, 128, 3, 1), (8, 4, 4, 513, 64, 3, 1),
]


@pytest.mark.parametrize("shape", CONV_SHAPES)
def test_conv_family_vs_contract(shape):
    N, H, W, Ci, Co, R, pad = shape
    x, w, b = cl(rn(N, Ci, H, W)), cl(rn(Co, Ci, R, R, seed=1)), rn(Co, seed=2)
    Ho = H + 2 * pad - R + 1
    gy = cl(rn(N, Co, Ho, W + 2 * pad - R + 1, seed=3))
    both("conv_fprop", x, w, b, pad, 0.37, 0.5, K.ACT_LRELU, 0.2)
    both("conv_fprop", x, w, None, pad, 1.0, 1.0, K.ACT_NONE, 0.2)
    both("conv_dgrad", gy, w, (H, W), pad, 0.37)
    both("conv_wgrad", x, gy, (R, R), pad, 0.37)


# TF32 tensor-core path (tcgen05): operands are truncated to 10 mantissa bits, accumulation is fp32.  Stated bound: 3e-3 of
# the tensor's max-norm (measured ~8e-4 on B200 over K = 288..4608); the fp32 FFMA path above keeps the 1e-4 bar.
TOL_TF32 = 3e-3
RESNET_GRAD_TOL = 5e-2      # whole-net ResNet gradients on the fp32 path (see parity_cases.case_resnet_nets_modules)
TC_SHAPES = [  # N, H, W, Ci, Co, R, pad
    (8, 4, 4, 512, 512, 3, 1), (8, 16, 16, 512, 512, 3, 1), (2, 32, 32, 256, 128, 3, 1), (2, 64, 64, 128, 256, 3, 1),
    (4, 8, 8, 96, 128, 3, 1), (3, 5, 7, 32, 64, 3, 1), (2, 16, 16, 64, 32, 3, 1), (2, 32, 32, 128, 128, 1, 0),
    (1, 128, 128, 128, 128, 3, 1),
    (3, 128, 128, 128, 128, 3, 1),      # >= one wave of M-tile pairs: the two-accumulator (BN = 128) kernel
]


@pytest.mark.parametrize("shape", TC_SHAPES)
def test_conv_family_tf32_vs_contract(shape):
    N, H, W, Ci, Co, R, pad = shape
    x, w, b = cl(rn(N, Ci, H, W)), cl(rn(Co, Ci, R, R, seed=1)), rn(Co, seed=2)
    gy = cl(rn(N, Co, H + 2 * pad - R + 1, W + 2 * pad - R + 1, seed=3))
    K.set_conv_impl("tf32")
    try:
        assert K.tc_covers("fprop", N, H, W, Ci, Co, R, R, pad) and K.tc_covers("wgrad", N, H, W, Ci, Co, R, R, pad)
        both("conv_fprop", x, w, b, pad, 0.37, 0.5, K.ACT_LRELU, 0.2, tol=TOL_TF32)
        both("conv_fprop", x, w, None, pad, 1.0, 1.0, K.ACT_NONE, 0.2, tol=TOL_TF32)
        both("conv_dgrad", gy, w, (H, W), pad, 0.37, tol=TOL_TF32)
        both("conv_wgrad", x, gy, (R, R), pad, 0.37, tol=TOL_TF32)
    finally:
        K.set_conv_impl("fp32")


# Upsample folded into the convolution (glb_upconv_*): N, H, W (low resolution), Ci, Co.  Covers the split-K path (4^2, 8^2), the
# single-CTA kernels at BN 256 / 128 / 64 / 32, the CTA-pair kernel (>= 37 pair items at BN 256), ragged maps (5 x 7) and every
# cfg2 layer shape at a reduced batch.
UPCONV_SHAPES = [
    (8, 4, 4, 512, 512), (8, 8, 8, 512, 512), (2, 16, 16, 512, 512), (8, 16, 16, 512, 512), (2, 32, 32, 512, 256),
    (2, 64, 64, 256, 128), (3, 5, 7, 32, 64), (2, 16, 16, 64, 32), (4, 8, 8, 96, 128), (1, 32, 32, 128, 64),
]


@pytest.mark.parametrize("shape", UPCONV_SHAPES)
def test_upconv_family_vs_contract(shape):
    """conv3x3(upsample2x(x)) as four 2x2 phase convolutions on the low-resolution map vs the literal composition in fp64."""
    N, H, W, Ci, Co = shape
    x, w, b = cl(rn(N, Ci, H, W)), cl(rn(Co, Ci, 3, 3, seed=1)), rn(Co, seed=2)
    gy = cl(rn(N, Co, 2 * H, 2 * W, seed=3))
    K.set_conv_impl("tf32")
    try:
        assert K.upconv_covers("fprop", N, H, W, Ci, Co) and K.upconv_covers("wgrad", N, H, W, Ci, Co)
        n0 = K.launch_count()
        both("upconv_fprop", x, w, b, 0.37, 0.5, K.ACT_LRELU, 0.2, tol=TOL_TF32)
        both("upconv_fprop", x, w, None, 1.0, 1.0, K.ACT_NONE, 0.2, tol=TOL_TF32)
        if K.upconv_covers("dgrad", N, H, W, Ci, Co):         # GEMM N = Ci: 32, 64 or a multiple of 128 (ops.upconv2d falls back otherwise)
            both("upconv_dgrad", gy, w, 0.37, tol=TOL_TF32)
        else:
            assert Ci == 96
        both("upconv_wgrad", x, gy, 0.37, tol=TOL_TF32)
        assert K.launch_count() > n0
    finally:
        K.set_conv_impl("fp32")


# Average pool folded into the convolution (glb_downconv_*): N, H, W (LOW-resolution output), Ci, Co
DOWNCONV_SHAPES = [
    (8, 4, 4, 512, 512), (8, 8, 8, 512, 512), (2, 16, 16, 512, 512), (2, 32, 32, 256, 512), (2, 64, 64, 128, 256),
    (3, 5, 7, 64, 32), (2, 16, 16, 32, 64), (1, 32, 32, 128, 128),
]


@pytest.mark.parametrize("shape", DOWNCONV_SHAPES)
def test_downconv_family_vs_contract(shape):
    """avgpool2x2(conv3x3(x)) (+ bias + lrelu) as one stride-2 4x4 convolution vs the literal composition in fp64."""
    N, H, W, Ci, Co = shape
    x, w, b = cl(rn(N, Ci, 2 * H, 2 * W)), cl(rn(Co, Ci, 3, 3, seed=1)), rn(Co, seed=2)
    gy = cl(rn(N, Co, H, W, seed=3))
    K.set_conv_impl("tf32")
    try:
        assert K.downconv_covers(N, H, W, Ci, Co)
        n0 = K.launch_count()
        both("downconv_fprop", x, w, b, 0.37, 0.5, K.ACT_LRELU, 0.2, tol=TOL_TF32)
        both("downconv_fprop", x, w, None, 1.0, 1.0, K.ACT_NONE, 0.2, tol=TOL_TF32)
        both("downconv_dgrad", gy, w, 0.37, tol=TOL_TF32)
        both("downconv_wgrad", x, gy, 0.37, tol=TOL_TF32)
        assert K.launch_count() > n0
    finally:
        K.set_conv_impl("fp32")


FOLD_BF16_SHAPES = [  # N, H, W (low resolution), Ci, Co -- channel counts multiples of 64
    (8, 4, 4, 512, 512), (2, 16, 16, 512, 512), (8, 16, 16, 512, 512), (2, 32, 32, 512, 256), (2, 64, 64, 256, 128), (2, 16, 16, 64, 64),
]


@pytest.mark.parametrize("shape", FOLD_BF16_SHAPES)
def test_folded_conv_families_bf16_vs_contract(shape):
    """bf16-operand variants of glb_upconv_* and glb_downconv_* (the layer's channel roles swapped for the latter)."""
    N, H, W, Ci, Co = shape
    x, w, b = cl(rn(N, Ci, H, W)), cl(rn(Co, Ci, 3, 3, seed=1)), rn(Co, seed=2)
    gy = cl(rn(N, Co, 2 * H, 2 * W, seed=3))
    xd, wd, bd = cl(rn(N, Co, 2 * H, 2 * W, seed=4)), cl(rn(Ci, Co, 3, 3, seed=5)), rn(Ci, seed=6)
    gyd = cl(rn(N, Ci, H, W, seed=7))
    K.set_conv_impl("bf16")
    try:
        for kind in ("fprop", "dgrad", "wgrad"):
            assert K.upconv_covers(kind, N, H, W, Ci, Co), kind
        assert K.downconv_covers(N, H, W, Co, Ci)
        n0 = K.launch_count()
        both("upconv_fprop", x, w, b, 0.37, 0.5, K.ACT_LRELU, 0.2, tol=TOL_BF16)
        both("upconv_dgrad", gy, w, 0.37, tol=TOL_BF16)
        both("upconv_wgrad", x, gy, 0.37, tol=TOL_BF16)
        both("downconv_fprop", xd, wd, bd, 0.37, 0.5, K.ACT_LRELU, 0.2, tol=TOL_BF16)
        both("downconv_dgrad", gyd, wd, 0.37, tol=TOL_BF16)
        both("downconv_wgrad", xd, gyd, 0.37, tol=TOL_BF16)
        assert K.launch_count() > n0
    finally:
        K.set_conv_impl("fp32")


def test_downconv_op_matches_two_kernel_sequence_incl_double_backward():
    """ops.downconv2d (fused) against conv2d -> pool_bias_act on the same TF32 path: output with the fused bias + LeakyReLU, and
    -- without the activation, whose mask would flip on outputs near zero between two differently rounded TF32 paths --
    first-order gradients and the gradients of an R1-style penalty (double backward through the fused family)."""
    from gan_lab_b200 import ops
    N, H, W, Ci, Co = 2, 16, 16, 128, 128
    x0, w0 = cl(rn(N, Ci, 2 * H, 2 * W)).to(DEV), cl(rn(Co, Ci, 3, 3, seed=1)).to(DEV)
    b0 = rn(1, Co, 1, 1, seed=2).to(DEV)
    K.set_conv_impl("tf32")
    try:
        ya = ops.downconv2d(x0, w0, b0, 0.05, 1.0, ops.ACT_LRELU, 0.2)
        yb = ops.pool_bias_act(ops.conv2d(x0, w0, None, 1, 0.05), b0, 1.0, ops.ACT_LRELU, 0.2)
        assert rel(ya, yb) < TOL_TF32
        outs = []
        for fused in (True, False):
            x, w, b = x0.clone().requires_grad_(True), w0.clone().requires_grad_(True), b0.clone().requires_grad_(True)
            n0 = K.launch_count()
            if fused:
                y = ops.downconv2d(x, w, b, 0.05, 1.0, ops.ACT_NONE, 0.2)
            else:
                y = ops.pool_bias_act(ops.conv2d(x, w, None, 1, 0.05), b, 1.0, ops.ACT_NONE, 0.2)
            (gx,) = torch.autograd.grad(y.square().sum(), x, create_graph=True)
            pen = (gx * gx).sum()
            (y.square().sum() + 10.0 * pen).backward()
            outs.append((y.detach(), gx.detach(), x.grad, w.grad, b.grad))
        for i, (a, c) in enumerate(zip(*outs)):
            assert rel(a, c) < TOL_TF32, (i, rel(a, c))
    finally:
        K.set_conv_impl("fp32")


def test_upconv_op_matches_two_kernel_sequence():
    """ops.upconv2d (fused) against ops.conv2d(ops.upsample2x(x)) on the same TF32 path: outputs and all three gradients;
    with create_graph the backward falls back to the differentiable composition (second-order derivative exists)."""
    from gan_lab_b200 import ops
    N, H, W, Ci, Co = 2, 16, 16, 128, 128
    x0, w0, b0 = cl(rn(N, Ci, H, W)).to(DEV), cl(rn(Co, Ci, 3, 3, seed=1)).to(DEV), rn(Co, seed=2).to(DEV)
    K.set_conv_impl("tf32")
    try:
        outs = []
        for fused in (True, False):
            x, w, b = x0.clone().requires_grad_(True), w0.clone().requires_grad_(True), b0.clone().requires_grad_(True)
            if fused:
                y = ops.upconv2d(x, w, b, 0.5, 1.0, ops.ACT_LRELU, 0.2)
            else:
                y = ops.conv2d(ops.upsample2x(x), w, b, 1, 0.5, 1.0, ops.ACT_LRELU, 0.2)
            (y * y).sum().backward()
            outs.append((y.detach(), x.grad, w.grad, b.grad))
        for a, c in zip(*outs):
            assert rel(a, c) < TOL_TF32, rel(a, c)
        x = x0.clone().requires_grad_(True)
        w = w0.clone().requires_grad_(True)
        y = ops.upconv2d(x, w, None, 0.5)
        (gx,) = torch.autograd.grad(y.sum(), x, create_graph=True)
        assert gx.requires_grad
        (gx * gx).sum().backward()
        assert w.grad is not None and torch.isfinite(w.grad).all()
    finally:
        K.set_conv_impl("fp32")


# bf16 operand path (tcgen05 kind::f16 on round-to-nearest bf16 copies of x / gy / w, fp32 accumulation): 8 mantissa bits per
# operand -> stated bound 1e-2 of the tensor's max-norm (expected ~3e-3 over K = 576..4608); channel counts multiples of 64.
TOL_BF16 = 1e-2
BF16_SHAPES = [  # N, H, W, Ci, Co, R, pad
    (8, 4, 4, 512, 512, 3, 1), (8, 16, 16, 512, 512, 3, 1), (2, 32, 32, 256, 128, 3, 1), (2, 64, 64, 128, 256, 3, 1),
    (2, 16, 16, 64, 64, 3, 1), (2, 32, 32, 128, 128, 1, 0), (8, 32, 32, 512, 512, 3, 1), (8, 64, 64, 256, 256, 3, 1),
    (3, 128, 128, 128, 128, 3, 1), (8, 4, 4, 512, 512, 4, 0),
]


@pytest.mark.parametrize("shape", BF16_SHAPES)
def test_conv_family_bf16_vs_contract(shape):
    N, H, W, Ci, Co, R, pad = shape
    x, w, b = cl(rn(N, Ci, H, W)), cl(rn(Co, Ci, R, R, seed=1)), rn(Co, seed=2)
    gy = cl(rn(N, Co, H + 2 * pad - R + 1, W + 2 * pad - R + 1, seed=3))
    K.set_conv_impl("bf16")
    try:
        for kind in ("fprop", "dgrad", "wgrad"):
            assert K.bf16_covers(kind, N, H, W, Ci, Co, R, R, pad), kind
        n0 = K.launch_count()
        both("conv_fprop", x, w, b, pad, 0.37, 0.5, K.ACT_LRELU, 0.2, tol=TOL_BF16)
        both("conv_fprop", x, w, None, pad, 1.0, 1.0, K.ACT_NONE, 0.2, tol=TOL_BF16)
        both("conv_dgrad", gy, w, (H, W), pad, 0.37, tol=TOL_BF16)
        both("conv_wgrad", x, gy, (R, R), pad, 0.37, tol=TOL_BF16)
        assert K.launch_count() > n0
    finally:
        K.set_conv_impl("fp32")


def test_bf16_conversion_is_round_to_nearest_even():
    x = torch.randn(3, 7, 5, 11, device=DEV) * 3
    x.view(-1)[:4] = torch.tensor([1.0, 1.00390625, 1.01171875, -65504.0], device=DEV)     # exact, tie -> even, tie -> even
    got = K.cvt_bf16(x.contiguous())
    assert got.dtype == torch.bfloat16 and torch.equal(got, x.to(torch.bfloat16))
    xc = cl(torch.randn(2, 64, 9, 9, device=DEV))
    assert torch.equal(K.cvt_bf16(xc), xc.to(torch.bfloat16)) and K.cvt_bf16(xc).is_contiguous(memory_format=torch.channels_last)


@pytest.mark.parametrize("shape", [(4, 32, 32, 3, 64, 3, 1), (4, 32, 32, 64, 3, 3, 1), (2, 16, 16, 3, 128, 3, 1)])
def test_conv_three_channel_sides_take_the_tensor_core_path(shape):
    """The 3 -> C and C -> 3 3x3 convolutions of the ResNet nets: the narrow side is zero-padded to 32 channels by the
    launchers; fprop / dgrad / wgrad must equal the contract of the unpadded convolution."""
    N, H, W, Ci, Co, R, pad = shape
    x, w, b = cl(rn(N, Ci, H, W)), cl(rn(Co, Ci, R, R, seed=1)), rn(Co, seed=2)
    gy = cl(rn(N, Co, H, W, seed=3))
    K.set_conv_impl("tf32")
    try:
        assert not K.tc_covers("fprop", N, H, W, Ci, Co, R, R, pad)
        assert (K._pad_ci("fprop", N, H, W, Ci, Co, R, R, pad) or K._pad_co("fprop", N, H, W, Ci, Co, R, R, pad)) == 32
        y, = both("conv_fprop", x, w, b, pad, 0.37, 0.5, K.ACT_LRELU, 0.2, tol=TOL_TF32)
        gx, = both("conv_dgrad", gy, w, (H, W), pad, 0.37, tol=TOL_TF32)
        gw, = both("conv_wgrad", x, gy, (R, R), pad, 0.37, tol=TOL_TF32)
        assert y.shape[1] == Co and gx.shape[1] == Ci and tuple(gw.shape) == (Co, Ci, R, R)
    finally:
        K.set_conv_impl("fp32")


@pytest.mark.parametrize("shape", [(8, 4, 4, 513, 512, 3, 1), (4, 4, 4, 257, 128, 3, 1)])
def test_conv_ragged_channels_take_the_tensor_core_path(shape):
    """Ci = 513 (behind the minibatch-stddev concat) is zero-padded to a multiple of 128 by the launchers and runs on the
    tcgen05 kernels; results (incl. the sliced dgrad / wgrad) must equal the contract of the UNPADDED convolution."""
    N, H, W, Ci, Co, R, pad = shape
    x, w, b = cl(rn(N, Ci, H, W)), cl(rn(Co, Ci, R, R, seed=1)), rn(Co, seed=2)
    gy = cl(rn(N, Co, H, W, seed=3))
    K.set_conv_impl("tf32")
    try:
        assert not K.tc_covers("fprop", N, H, W, Ci, Co, R, R, pad)
        assert K._pad_ci("fprop", N, H, W, Ci, Co, R, R, pad) == (Ci + 127) // 128 * 128
        y, = both("conv_fprop", x, w, b, pad, 0.37, 0.5, K.ACT_LRELU, 0.2, tol=TOL_TF32)
        gx, = both("conv_dgrad", gy, w, (H, W), pad, 0.37, tol=TOL_TF32)
        gw, = both("conv_wgrad", x, gy, (R, R), pad, 0.37, tol=TOL_TF32)
        assert gx.shape[1] == Ci and gw.shape[1] == Ci and gw.is_contiguous(memory_format=torch.channels_last)
        wd = w.to(DEV)
        a = K.conv_fprop(x.to(DEV), wd, None, pad, 1.0, 1.0, K.ACT_NONE, 0.0)
        wd.mul_(2.0)                                   # the cached padded copy must follow in-place weight updates
        assert rel(K.conv_fprop(x.to(DEV), wd, None, pad, 1.0, 1.0, K.ACT_NONE, 0.0), a * 2) < 1e-6
    finally:
        K.set_conv_impl("fp32")


@pytest.mark.parametrize("shape", [(2, 32, 32, 16, 16, 3), (1, 64, 64, 32, 16, 3), (2, 16, 32, 16, 32, 3), (2, 32, 32, 16, 16, 1),
                                   (1, 128, 128, 16, 16, 3)])
def test_conv_narrow_layers_pixel_pair_packing(shape):
    """16-channel layers (512^2 / 1024^2 blocks of the FFHQ-shape nets) run on the tcgen05 kernels through the pixel-pair
    packed view (two pixels = one 32-channel row, expanded weight); all three of fprop / dgrad / wgrad must equal the
    contract of the ORIGINAL convolution."""
    N, H, W, Ci, Co, R = shape
    pad = (R - 1) // 2
    x, w, b = cl(rn(N, Ci, H, W)), cl(rn(Co, Ci, R, R, seed=1)), rn(Co, seed=2)
    gy = cl(rn(N, Co, H, W, seed=3))
    K.set_conv_impl("tf32")
    try:
        kinds = [k for k in ("fprop", "dgrad", "wgrad") if K._pack_ok(k, N, H, W, Ci, Co, R, R, pad)]
        assert kinds, "no kind of this shape takes the packed path"
        both("conv_fprop", x, w, b, pad, 0.37, 0.5, K.ACT_LRELU, 0.2, tol=TOL_TF32)
        both("conv_dgrad", gy, w, (H, W), pad, 0.37, tol=TOL_TF32)
        both("conv_wgrad", x, gy, (R, R), pad, 0.37, tol=TOL_TF32)
    finally:
        K.set_conv_impl("fp32")


@pytest.mark.parametrize("shape", [(8, 64, 64, 128, 128), (8, 64, 64, 256, 256), (8, 128, 128, 128, 128), (8, 128, 128, 128, 256),
                                   (16, 256, 256, 64, 64)])
def test_tf32_highres_kernels_vs_exact_fp32_kernels(shape):
    """The kernels only full-size layers select -- CTA pair + tap reuse fprop/dgrad (>= 74 pair items), three-tap wgrad
    (>= 2048 K blocks) -- against the exact fp32 FFMA kernels on the same device tensors (the fp32 kernels themselves are
    pinned to the fp64 contracts at small shapes above).  TF32 bound as for every tensor-core conv."""
    N, H, W, Ci, Co = shape
    g = torch.Generator().manual_seed(H + Ci + Co)
    x = cl(torch.randn(N, Ci, H, W, generator=g)).to(DEV)
    w = cl(torch.randn(Co, Ci, 3, 3, generator=g)).to(DEV)
    b = torch.randn(Co, generator=g).to(DEV)
    gy = cl(torch.randn(N, Co, H, W, generator=g)).to(DEV)
    outs = {}
    for impl in ("fp32", "tf32"):
        K.set_conv_impl(impl)
        try:
            outs[impl] = (K.conv_fprop(x, w, b, 1, 0.37, 0.5, K.ACT_LRELU, 0.2), K.conv_dgrad(gy, w, (H, W), 1, 0.37),
                          K.conv_wgrad(x, gy, (3, 3), 1, 0.37))
        finally:
            K.set_conv_impl("fp32")
    for name, a, r in zip(("fprop", "dgrad", "wgrad"), outs["tf32"], outs["fp32"]):
        assert rel(a, r) < TOL_TF32, (name, rel(a, r))


def test_tf32_dgrad_sees_weight_updates():
    """The tensor-core dgrad multiplies by a cached re-layout of the weight: an in-place update must invalidate it."""
    K.set_conv_impl("tf32")
    try:
        w, gy = cl(rn(64, 64, 3, 3)).to(DEV), cl(rn(2, 64, 8, 8, seed=1)).to(DEV)
        a = K.conv_dgrad(gy, w, (8, 8), 1, 1.0)
        w.mul_(2.0)
        b = K.conv_dgrad(gy, w, (8, 8), 1, 1.0)
        assert rel(b, a * 2) < 1e-6
    finally:
        K.set_conv_impl("fp32")


@pytest.mark.parametrize("M,Kf,Nout", [(8, 32, 32), (8, 512, 512), (4, 512, 8192), (8, 512, 1), (64, 8192, 1), (5, 37, 11)])
def test_linear_family_vs_contract(M, Kf, Nout):
    x, w, b, gy = rn(M, Kf), rn(Nout, Kf, seed=1), rn(Nout, seed=2), rn(M, Nout, seed=3)
    both("linear_fwd", x, w, b, 0.01, 0.01, K.ACT_LRELU, 0.2)
    both("linear_dgrad", gy, w, 0.3)
    both("linear_wgrad", x, gy, 0.3)


@pytest.mark.parametrize("N,C,H,W", [(2, 64, 40, 52), (2, 32, 33, 47), (8, 128, 128, 128), (3, 256, 64, 64), (1, 512, 32, 32)])
def test_blur_tile_kernels_vs_contract(N, C, H, W):
    """The TMA-staged shared-memory blur kernels (maps >= 32 x 32, channel counts multiples of 32), ragged tile edges included:
    plain blur, and blur + activation mask + bias gradient."""
    x, y, gy = cl(rn(N, C, H, W)), cl(rn(N, C, H, W, seed=1)), cl(rn(N, C, H, W, seed=2))
    both("blur3x3", x)
    both("blur_act_bwd", gy, y, True, 0.7, K.ACT_LRELU, 0.2, tol=3e-4)       # (bias gradient: fp32 atomics over N*H*W terms)
    both("blur_act_bwd", gy, y, False, 1.0, K.ACT_LRELU, 0.2)


@pytest.mark.parametrize("N,C,H,W", [(2, 16, 8, 8), (4, 512, 4, 4), (1, 128, 32, 32), (3, 2048, 2, 2), (2, 32, 6, 10)])
def test_glue_kernels_vs_contract(N, C, H, W):
    x, y, gy, b = cl(rn(N, C, H, W)), cl(rn(N, C, H, W, seed=1)), cl(rn(N, C, H, W, seed=2)), rn(C, seed=3)
    both("bias_act_fwd", x, b, 0.7, K.ACT_LRELU, 0.2)
    both("act_bwd", gy, y, True, 0.7, K.ACT_LRELU, 0.2)
    both("act_bwd", gy, y, False, 1.0, K.ACT_LRELU, 0.2)
    both("colsum", x, 0.5)
    both("axpby", x, y, 0.3, 0.7)
    both("sumsq", x, 0.01)
    both("blur3x3", x)
    both("blur_act_bwd", gy, y, True, 0.7, K.ACT_LRELU, 0.2)
    both("blur_act_bwd", gy, y, False, 1.0, K.ACT_LRELU, 0.2)
    both("upsample2x_fwd", x)
    both("pool_bias_act_fwd", x, b, 0.7, K.ACT_LRELU, 0.2)
    both("pool_bias_act_fwd", x, None, 1.0, K.ACT_NONE, 0.2)
    gyp, yp = cl(rn(N, C, H // 2, W // 2, seed=4)), cl(rn(N, C, H // 2, W // 2, seed=5))
    both("upsample2x_bwd", gy)
    both("pool_bias_act_bwd", gyp, yp, True, 0.7, K.ACT_LRELU, 0.2)
    both("pool_bias_act_bwd", gyp, None, False, 1.0, K.ACT_NONE, 0.2)
    if C <= 1024:
        both("pixelnorm_fwd", x, 1e-8)
        both("pixelnorm_bwd", gy, x, 1e-8)


@pytest.mark.parametrize("N,C,H,W", [(4, 8, 64, 64), (3, 64, 8, 8), (64, 64, 16, 16), (2, 512, 4, 4), (5, 16, 6, 10)])
def test_norm_kernels_vs_contract(N, C, H, W):
    """LayerNorm([C,H,W]) + ReLU with its first- and second-order backward, BatchNorm2d (batch statistics, running-buffer
    update) + ReLU with its backward, Tanh -- each launcher against its fp64 contract on the same inputs.  The mask of the
    backward kernels is the saved forward OUTPUT, passed in identically on both sides, so no bit can flip here."""
    x, gy, u = cl(rn(N, C, H, W) * 1.5 + 0.7), cl(rn(N, C, H, W, seed=1)), cl(rn(N, C, H, W, seed=2))
    gam, bet = 1.0 + 0.3 * rn(C, H, W, seed=3), 0.3 * rn(C, H, W, seed=4)
    for act, slope in ((K.ACT_LRELU, 0.0), (K.ACT_NONE, 0.0), (K.ACT_LRELU, 0.2)):
        y, stats = both("layernorm_fwd", x, gam, bet, 1e-5, act, slope)
        yc = y.cpu()
        both("layernorm_bwd", gy, yc, x, gam, stats.cpu(), act, slope, tol=2e-4)
        both("layernorm_bwd", gy, yc, x, gam, stats.cpu(), act, slope, want_gx=False, tol=2e-4)
        both("layernorm_bwdbwd", u, gy, yc, x, gam, stats.cpu(), act, slope, tol=5e-4)
    # parameters stored the way the LayerNorm module stores them: (H,W,C) memory behind the (C,H,W) shape
    gam_p = gam.permute(1, 2, 0).contiguous().permute(2, 0, 1)
    y2, _ = K.layernorm_fwd(x.to(DEV), gam_p.to(DEV), bet.to(DEV), 1e-5, K.ACT_LRELU, 0.0)
    yc, _ = KC.layernorm_fwd(x.double(), gam.double(), bet.double(), 1e-5, K.ACT_LRELU, 0.0)
    assert rel(y2, yc) < TOL
    g1, b1 = 1.0 + 0.3 * rn(C, seed=5), 0.3 * rn(C, seed=6)
    for act, slope in ((K.ACT_LRELU, 0.0), (K.ACT_NONE, 0.0)):
        rm, rv, nbt = rn(C, seed=7) * 0.1, rn(C, seed=8).abs() + 0.5, torch.tensor(3, dtype=torch.int64)
        rm_g, rv_g, nbt_g = rm.to(DEV), rv.to(DEV), nbt.to(DEV)
        y, stats = K.batchnorm_fwd(x.to(DEV), g1.to(DEV), b1.to(DEV), rm_g, rv_g, nbt_g, 1e-5, 0.1, act, slope)
        rm_c, rv_c, nbt_c = rm.double(), rv.double(), nbt.clone()
        y_c, stats_c = KC.batchnorm_fwd(x.double(), g1.double(), b1.double(), rm_c, rv_c, nbt_c, 1e-5, 0.1, act, slope)
        assert rel(y, y_c) < TOL and rel(stats, stats_c) < TOL
        assert rel(rm_g, rm_c) < TOL and rel(rv_g, rv_c) < TOL and int(nbt_g) == int(nbt_c) == 4
        both("batchnorm_bwd", gy, y.cpu(), x, g1, stats.cpu(), act, slope, tol=5e-4)
    both("tanh_fwd", x)
    both("tanh_bwd", gy, torch.tanh(x))


@pytest.mark.parametrize("M,Kf,nouts", [(8, 512, (1024, 1024, 512, 256)), (4, 32, (64, 64, 32, 16, 8)), (16, 512, (1024, 256))])
def test_grouped_linear_matches_per_layer_linears(M, Kf, nouts):
    """All style affines of a generator pass as ONE grouped launch per direction == the per-layer LinearEx kernels:
    outputs, the gradient of the whole dlatent stack, every weight and bias gradient."""
    from gan_lab_b200 import ops
    L = len(nouts)
    ws = rn(L, M, Kf).to(DEV).requires_grad_(True)
    layers = [(rn(n, Kf, seed=10 + i).to(DEV).requires_grad_(True), rn(n, seed=20 + i).to(DEV).requires_grad_(True),
               0.05 + 0.01 * i, 1.0 if i % 2 else 0.5) for i, n in enumerate(nouts)]
    gs = [rn(M, n, seed=30 + i).to(DEV) for i, n in enumerate(nouts)]
    outs = ops.grouped_linear(ws, K.GroupedLinearTable(), layers)
    torch.autograd.backward(list(outs), gs)
    got = (ws.grad.clone(), [w.grad.clone() for w, _b, _a, _s in layers], [b.grad.clone() for _w, b, _a, _s in layers])
    ws.grad = None
    for w, b, _a, _s in layers:
        w.grad = None; b.grad = None
    refs = [ops.linear(ws[l], w, b, a, bs) for l, (w, b, a, bs) in enumerate(layers)]
    torch.autograd.backward(refs, gs)
    for o, r in zip(outs, refs):
        assert rel(o, r) < TOL
    assert rel(got[0], ws.grad) < TOL
    for (w, b, _a, _s), gw, gb in zip(layers, got[1], got[2]):
        assert rel(gw, w.grad) < TOL and rel(gb, b.grad) < TOL


def test_pixelnorm_latents():
    x, gy = rn(8, 512), rn(8, 512, seed=1)
    both("pixelnorm_fwd", x, 1e-8)
    both("pixelnorm_bwd", gy, x, 1e-8)


@pytest.mark.parametrize("mode", ["0", "1", "2"])       # three kernels / one sample-ordered kernel (default) / cluster kernels
@pytest.mark.parametrize("N,C,H,W", [(4, 32, 8, 8), (8, 512, 4, 4), (2, 128, 64, 64), (2, 16, 128, 128), (3, 64, 16, 16),
                                     (8, 128, 128, 128), (5, 8, 32, 32)])
def test_style_epilogue_vs_contract(N, C, H, W, mode, monkeypatch):
    monkeypatch.setenv("GLB_SE_MODE", mode)
    x, noise, nw, b = cl(rn(N, C, H, W)), rn(N, 1, H, W, seed=1), rn(C, seed=2) * .3, rn(C, seed=3) * .3
    style, gout = rn(N, 2 * C, seed=4), cl(rn(N, C, H, W, seed=5))
    out_g, stats = K.style_epilogue_fwd(x.to(DEV), noise.to(DEV), nw.to(DEV), b.to(DEV), style.to(DEV), 0.2, 1e-8)
    out_c, st_c = KC.style_epilogue_fwd(x.double(), noise.double(), nw.double(), b.double(), style.double(), 0.2, 1e-8)
    assert rel(out_g, out_c) < TOL
    g = K.style_epilogue_bwd(gout.to(DEV), x.to(DEV), noise.to(DEV), nw.to(DEV), b.to(DEV), style.to(DEV), stats, 0.2)
    c = KC.style_epilogue_bwd(gout.double(), x.double(), noise.double(), nw.double(), b.double(), style.double(), st_c, 0.2)
    for name, a, r in zip(("gx", "gstyle", "g_nw", "g_b"), g, c):
        # g_b / g_nw pass through the mean-removing InstanceNorm: cancellation-dominated -> 5e-4
        assert rel(a, r) < (5e-4 if name in ("g_nw", "g_b") else TOL), (name, rel(a, r))
    # InstanceNorm alone (no noise / bias, identity activation, zero style)
    o2, _ = K.style_epilogue_fwd(x.to(DEV), None, None, None, torch.zeros(N, 2 * C, device=DEV), 1.0, 1e-8)
    from oracle import gan_oracle as O
    assert rel(o2, O.instance_norm(x.double(), 1e-8)) < TOL


@pytest.mark.parametrize("N,C,group", [(8, 16, 4), (4, 512, 4), (6, 8, 6), (1, 8, 1), (16, 512, 4)])
def test_mbstd_vs_contract(N, C, group):
    x, gy, v = cl(rn(N, C, 4, 4)), cl(rn(N, C + 1, 4, 4, seed=1)), cl(rn(N, C, 4, 4, seed=2))
    both("mbstd_fwd", x, group)
    both("mbstd_bwd", gy, x, group)
    if N > 1:
        both("mbstd_bwdbwd", v, gy, x, group, tol=5e-4)


@pytest.mark.parametrize("N,C,H,W,pool", [(2, 32, 16, 16, False), (2, 128, 8, 8, True), (4, 16, 32, 32, False),
                                           (1, 512, 4, 4, False), (2, 2048, 2, 2, False)])
def test_rgb_kernels_vs_contract(N, C, H, W, pool):
    s = 2 if pool else 1
    img, feat = rn(N, 3, H * s, W * s), cl(rn(N, C, H, W, seed=1))
    w_from, w_to, b = rn(C, 3, 1, 1, seed=2), rn(3, C, 1, 1, seed=3), rn(C, seed=4)
    both("rgb_expand", img, w_from, 1, 3, C, b, pool, 0.4, 0.9, K.ACT_LRELU, 0.2)        # fromRGB
    both("rgb_expand", img, w_to, C, 1, C, None, pool, 0.4, 1.0, K.ACT_NONE, 0.2)       # toRGB dgrad
    both("rgb_contract", feat, w_to, C, 1, rn(3, seed=5), pool, 0.4, 0.9)              # toRGB
    both("rgb_contract", feat, w_from, 1, 3, None, pool, 0.4, 1.0)                     # fromRGB dgrad
    both("rgb_wgrad", img, feat, (C, 3, 1, 1), 1, 3, pool, 0.4)
    both("rgb_wgrad", img, feat, (3, C, 1, 1), C, 1, pool, 0.4)
    both("plane_sum", img, 0.5)


def test_misc_kernels_vs_contract():
    lo, hi, gout = rn(2, 3, 8, 8), rn(2, 3, 16, 16, seed=1), rn(2, 3, 16, 16, seed=2)
    both("fade_up_blend", lo, hi, 0.3)
    both("fade_up_blend_bwd", gout, 0.3)
    both("fade_real", hi, 0.3)
    dg, dr = rn(8), rn(8, seed=1)
    for kind in ("wgan", "nonsaturating", "minimax"):
        if kind != "minimax":
            both("d_logit_loss", dg, dr, kind, 0.001)
        both("g_logit_loss", dg, kind)
    g = rn(4, 3, 16, 16)
    s = torch.tensor(0.7)
    both("scale_by", g, s, 1.3)
    both("gp_norm_fwd", g, 1.0, 0.01)
    both("gp_norm_bwd", g, s, 1.0, 0.01)
    both("interp_rows", rn(4, 3, 8, 8), rn(4, 3, 8, 8, seed=1), torch.rand(4, 1, 1, 1))
    w = rn(8, 512)
    e1 = torch.zeros(512, device=DEV)
    K.w_ewma_update(w.to(DEV), e1, 0.0)
    assert rel(e1, w.mean(0)) < TOL
    K.w_ewma_update(w.to(DEV) * 2, e1, 0.995)
    assert rel(e1, w.mean(0) * 2 * .005 + w.mean(0) * .995) < TOL


def test_fused_adam_ewma_matches_torch_adam():
    from gan_lab_b200.optim import FusedAdam
    torch.manual_seed(0)
    shapes = [(64, 32, 3, 3), (32,), (1, 32, 1, 1), (7, 5)]
    ps = [torch.randn(s, device=DEV).requires_grad_(True) for s in shapes]
    ps[0].data = ps[0].data.contiguous(memory_format=torch.channels_last)
    qs = [p.detach().clone().requires_grad_(True) for p in ps]
    lag = {str(i): p.detach().clone() for i, p in enumerate(ps)}
    ref_lag = None
    mine = FusedAdam(ps, lr=1e-3, betas=(0.0, 0.99), eps=1e-8)
    mine.attach_ewma([(str(i), p) for i, p in enumerate(ps)], lag, 0.999)
    ref = torch.optim.Adam(qs, lr=1e-3, betas=(0.0, 0.99), eps=1e-8)
    for step in range(3):
        for p, q in zip(ps, qs):
            g = torch.randn_like(p)
            p.grad = g.clone(); q.grad = g.clone()
        mine.step(); ref.step()
        ref_lag = [q.detach() * (1 - 0.999) + (q.detach() if ref_lag is None else ref_lag[i]) * 0.999 for i, q in enumerate(qs)]
        for i, (p, q) in enumerate(zip(ps, qs)):
            assert rel(p, q) < 1e-6, (step, i)
            assert rel(lag[str(i)], ref_lag[i]) < 1e-6, (step, i)


# ------------------------------------------------------------------------------------------ modules / nets / train vs golden
@pytest.mark.parametrize("name", PC.CONV_CASES)
def test_conv2d_ex_module(golden, name):
    PC.case_conv2d_ex_module(golden, DEV, name)


@pytest.mark.parametrize("name", PC.LINEAR_CASES)
def test_linear_ex_module(golden, name):
    PC.case_linear_ex_module(golden, DEV, name)


def test_small_modules(golden):
    PC.case_small_modules(golden, DEV)


@pytest.mark.parametrize("name", PC.MBSTD_CASES)
def test_mbstd_module(golden, name):
    PC.case_mbstd_module(golden, DEV, name)


def test_style_epilogue_op(golden):
    PC.case_style_epilogue_op(golden, DEV)


@pytest.mark.parametrize("fname", PC.STYLE_NETS)
def test_style_nets_modules(golden, fname):
    PC.case_style_nets_modules(golden, DEV, fname)


@pytest.mark.parametrize("fname", PC.PRO_NETS)
def test_pro_nets_modules(golden, fname):
    PC.case_pro_nets_modules(golden, DEV, fname)


@pytest.mark.parametrize("fname,model", PC.TRAIN_CASES)
def test_learner_train(golden, fname, model):
    PC.case_learner_train(golden, DEV, fname, model)


def test_cuda_graph_replay_of_train_steps():
    """D step + G step captured into CUDA graphs (device-side mixing decision, device-side Adam step count) and replayed:
    the optimiser state advances per replay, fresh latents/noise are drawn per replay, and everything stays finite."""
    from gan_lab_b200.config import default_config
    from gan_lab_b200.stylegan.learner import StyleGANLearner
    torch.manual_seed(0)
    cfg = default_config("StyleGAN", res=16, batch_size=4, dev=DEV, len_latent=32, len_dlatent=32, cutoff_trunc_trick=2)
    L = StyleGANLearner(cfg)
    L.gen_model.train(); L.disc_model.train()
    L.beta = L.get_smoothing_ewma_beta(10.)
    L._init_lagged(); L._attach_ewma()
    L.enable_cuda_graphs(True, warmup_iters=2)
    x = torch.rand(4, 3, 16, 16, device=DEV) * 2 - 1
    for _ in range(2):
        L.main_iteration(x)
    assert L._graph is None
    ld, lg = L.main_iteration(x)                      # capture + first replay
    assert L._graph is not None
    torch.cuda.synchronize()
    step0 = float(next(iter(L.opt_disc._hyper.values()))['t'][3])
    w0 = L.disc_model.state_dict()["fromrgb.0.conv2d.weight"].clone()
    l1 = (float(ld), float(lg))
    ld, lg = L.main_iteration(x)
    torch.cuda.synchronize()
    l2 = (float(ld), float(lg))
    assert float(next(iter(L.opt_disc._hyper.values()))['t'][3]) == step0 + 1
    assert not torch.equal(w0, L.disc_model.state_dict()["fromrgb.0.conv2d.weight"])
    assert all(math.isfinite(v) for v in l1 + l2) and l1 != l2
    for p in list(L.gen_model.parameters()) + list(L.disc_model.parameters()) + list(L.gen_model_lagged.parameters()):
        assert torch.isfinite(p).all()


@pytest.mark.gpu
@pytest.mark.parametrize("gp", ["r1", "r2"])
def test_shared_penalty_forward(gp):
    PC.case_shared_penalty_forward(DEV, gp)


@pytest.mark.parametrize("fname", PC.RESNET_NETS)
def test_resnet_nets_modules(golden, fname):
    report = {}
    try:
        PC.case_resnet_nets_modules(golden, DEV, fname, grad_tol=RESNET_GRAD_TOL, report=report, d_grad_tol=2e-4)
    finally:
        d = os.environ.get("GLB_DUMP_PARITY")
        if d:
            import json
            os.makedirs(d, exist_ok=True)
            with open(os.path.join(d, "resnet_nets_" + fname.replace(".pt", ".json")), "w") as f:
                json.dump(report, f)


def test_resnet_train(golden):
    PC.case_resnet_train(golden, DEV)


def test_style_generator_eval_mode(golden):
    PC.case_style_eval(golden, DEV)


def test_library_was_loaded():
    from gan_lab_b200._lib import LIB
    assert LIB._dll is not None and K.launch_count() > 0
