"""Differentiable operators over the sm_100a kernels.

Every op is a `torch.autograd.Function` whose backward is written in terms of *other* ops of this
module, so the set is closed under differentiation: the R1 / WGAN-GP penalty
(`autograd.grad(..., create_graph=True)` followed by `loss.backward()`, reference
resnetgan/learner.py:811-825) differentiates straight through the discriminator's custom kernels.
Generator-only ops (StyleGAN epilogue, PixelNorm, image fade blend) are first-order.

`input_grads_only()` is a hint context for gradient-penalty style calls (`autograd.grad` w.r.t. the
input image only): inside it, weight/bias gradients that autograd would compute and immediately
discard are skipped -- what ATen's `convolution_backward` output_mask does for the reference.
"""
from __future__ import annotations

import contextlib

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _kernels as K

ACT_NONE, ACT_LRELU = K.ACT_NONE, K.ACT_LRELU

_hints = {"input_only": False}


@contextlib.contextmanager
def input_grads_only():
    old = _hints["input_only"]
    _hints["input_only"] = True
    try:
        yield
    finally:
        _hints["input_only"] = old


def _param_grads_wanted() -> bool:
    return not _hints["input_only"]


# ----------------------------------------------------------------------------------------- elementwise
class _Scale(Function):
    @staticmethod
    def forward(ctx, x, s):
        ctx.s = s
        return K.axpby(x, None, s, 0.0)

    @staticmethod
    def backward(ctx, g):
        return _Scale.apply(g, ctx.s), None


class _Axpby(Function):
    """a*alpha + b*beta (fade-in blend of feature maps, reference progan/architectures.py:312-313)."""

    @staticmethod
    def forward(ctx, a, b, alpha, beta):
        ctx.ab = (alpha, beta)
        return K.axpby(a, b, alpha, beta)

    @staticmethod
    def backward(ctx, g):
        alpha, beta = ctx.ab
        ga = _Scale.apply(g, alpha) if ctx.needs_input_grad[0] else None
        gb = _Scale.apply(g, beta) if ctx.needs_input_grad[1] else None
        return ga, gb, None, None


def axpby(a, b, alpha, beta):
    return _Axpby.apply(a, b, K.coef_or_float(alpha), K.coef_or_float(beta))


class _ActBwd(Function):
    """(gy, y) -> gx = gy * act'(y)  [+ per-channel sum = bias gradient].  Linear in gy; y is a mask."""

    @staticmethod
    def forward(ctx, gy, y, want_bias, bias_scale, act, slope):
        ctx.save_for_backward(y)
        ctx.cfg = (want_bias, bias_scale, act, slope)
        ctx.set_materialize_grads(False)
        gx, gb = K.act_bwd(gy, y, want_bias, bias_scale, act, slope)
        return gx, gb

    @staticmethod
    def backward(ctx, ggx, ggb):
        (y,) = ctx.saved_tensors
        want_bias, bias_scale, act, slope = ctx.cfg
        g = ggx
        if ggb is not None:
            # someone differentiates the bias gradient itself (never on the training path): d(gb)/d(gy)
            shape = (1, -1, 1, 1) if y.dim() == 4 else (1, -1)
            extra = (ggb.view(shape) * bias_scale).expand(y.shape)
            g = extra if g is None else g + extra
        if g is None:
            return None, None, None, None, None, None
        out, _ = _ActBwd.apply(g.contiguous() if g.dim() != 4 else g, y, False, 1.0, act, slope)
        return out, None, None, None, None, None


def act_bwd(gy, y, want_bias, bias_scale, act, slope):
    gx, gb = _ActBwd.apply(gy, y, want_bias, bias_scale, act, slope)
    return gx, (gb if want_bias else None)


class _ColSum(Function):
    @staticmethod
    def forward(ctx, x, scale):
        ctx.shape, ctx.scale, ctx.nd = x.shape, scale, x.dim()
        return K.colsum(x, scale)

    @staticmethod
    def backward(ctx, g):
        view = (1, -1, 1, 1) if ctx.nd == 4 else (1, -1)
        return (g.view(view) * ctx.scale).expand(ctx.shape), None


class _BiasAct(Function):
    """y = act(x + bias_scale*bias): reference Conv2dBias + LeakyReLU (utils/custom_layers.py:213-226)."""

    @staticmethod
    def forward(ctx, x, bias, bias_scale, act, slope):
        y = K.bias_act_fwd(x, bias, bias_scale, act, slope)
        ctx.save_for_backward(y)
        ctx.cfg = (bias_scale, act, slope, bias is not None, None if bias is None else bias.shape)
        return y

    @staticmethod
    def backward(ctx, gy):
        (y,) = ctx.saved_tensors
        bias_scale, act, slope, has_bias, bshape = ctx.cfg
        want_b = has_bias and ctx.needs_input_grad[1] and _param_grads_wanted()
        gx, gb = act_bwd(gy, y, want_b, bias_scale, act, slope)
        return gx, (gb.view(bshape) if gb is not None else None), None, None, None


def bias_act(x, bias, bias_scale=1.0, act=ACT_NONE, slope=0.2):
    return _BiasAct.apply(x, bias, float(bias_scale), int(act), float(slope))


class _BlurActBwd(Function):
    """(gz, y) -> g = blur(gz) * act'(y)  [+ bias gradient]: backward of `blur(act(.))` fused.  Linear in gz; its own
    backward (only reached by the R1 / WGAN-GP double backward) is the unfused adjoint blur(gg * act'(y))."""

    @staticmethod
    def forward(ctx, gz, y, want_bias, bias_scale, act, slope):
        ctx.save_for_backward(y)
        ctx.cfg = (bias_scale, act, slope)
        ctx.set_materialize_grads(False)
        return K.blur_act_bwd(gz, y, want_bias, bias_scale, act, slope)

    @staticmethod
    def backward(ctx, gg, ggb):
        (y,) = ctx.saved_tensors
        bias_scale, act, slope = ctx.cfg
        if ggb is not None:
            raise NotImplementedError("differentiating the bias gradient of the fused blur/activation backward")
        if gg is None:
            return None, None, None, None, None, None
        masked, _ = _ActBwd.apply(gg, y, False, 1.0, act, slope)
        return _Blur.apply(masked), None, None, None, None, None


# ----------------------------------------------------------------------------------------- convolution
class _ConvFprop(Function):
    """y = act(alpha * conv(x, w) + bias_scale * bias)   (Conv2dEx.forward, utils/custom_layers.py:202-211)."""

    @staticmethod
    def forward(ctx, x, w, bias, pad, alpha, bias_scale, act, slope, blur=False):
        y = K.conv_fprop(x, w, bias, pad, alpha, bias_scale, act, slope)
        blur = bool(blur) and act != ACT_NONE
        ctx.cfg = (pad, alpha, bias_scale, act, slope, bias is not None, x.shape, w.shape, blur)
        ctx.save_for_backward(x, w, y if act != ACT_NONE else None)
        # blur=True: the binomial FIR that follows the activation in the D blocks rides in this op, so that the backward can
        # run blur + activation mask + bias gradient as ONE pass over the gradient (glb_blur_act_bwd)
        return K.blur3x3(y) if blur else y

    @staticmethod
    def backward(ctx, gy):
        x, w, y = ctx.saved_tensors
        pad, alpha, bias_scale, act, slope, has_bias, xshape, wshape, blur = ctx.cfg
        pg = _param_grads_wanted()
        want_b = has_bias and ctx.needs_input_grad[2] and pg
        if blur:
            g, gb = _BlurActBwd.apply(gy, y, want_b, bias_scale, act, slope)
        elif act != ACT_NONE:
            g, gb = act_bwd(gy, y, want_b, bias_scale, act, slope)
        else:
            g = gy
            gb = _ColSum.apply(g, bias_scale) if want_b else None
        gx = _ConvDgrad.apply(g, w, xshape[2], xshape[3], pad, alpha) if ctx.needs_input_grad[0] else None
        gw = _ConvWgrad.apply(x, g, wshape[2], wshape[3], pad, alpha) if (ctx.needs_input_grad[1] and pg) else None
        return gx, gw, gb, None, None, None, None, None, None


class _ConvDgrad(Function):
    """gx = alpha * conv_transpose(gy, w).  Bilinear in (gy, w)."""

    @staticmethod
    def forward(ctx, gy, w, H, W, pad, alpha):
        ctx.cfg = (H, W, pad, alpha, w.shape)
        ctx.save_for_backward(gy, w)
        return K.conv_dgrad(gy, w, (H, W), pad, alpha)

    @staticmethod
    def backward(ctx, ggx):
        gy, w = ctx.saved_tensors
        H, W, pad, alpha, wshape = ctx.cfg
        g_gy = _ConvFprop.apply(ggx, w, None, pad, alpha, 1.0, ACT_NONE, 0.0) if ctx.needs_input_grad[0] else None
        g_w = _ConvWgrad.apply(ggx, gy, wshape[2], wshape[3], pad, alpha) if ctx.needs_input_grad[1] else None
        return g_gy, g_w, None, None, None, None


class _ConvWgrad(Function):
    """gw = alpha * sum_pixels gy (x) x.  Bilinear in (x, gy)."""

    @staticmethod
    def forward(ctx, x, gy, R, S, pad, alpha):
        ctx.cfg = (R, S, pad, alpha, x.shape)
        ctx.save_for_backward(x, gy)
        return K.conv_wgrad(x, gy, (R, S), pad, alpha)

    @staticmethod
    def backward(ctx, ggw):
        x, gy = ctx.saved_tensors
        R, S, pad, alpha, xshape = ctx.cfg
        g_x = _ConvDgrad.apply(gy, ggw, xshape[2], xshape[3], pad, alpha) if ctx.needs_input_grad[0] else None
        g_gy = _ConvFprop.apply(x, ggw, None, pad, alpha, 1.0, ACT_NONE, 0.0) if ctx.needs_input_grad[1] else None
        return g_x, g_gy, None, None, None, None


def conv2d(x, w, bias=None, pad=0, alpha=1.0, bias_scale=1.0, act=ACT_NONE, slope=0.2, blur=False):
    return _ConvFprop.apply(x, w, bias, int(pad), float(alpha), float(bias_scale), int(act), float(slope), bool(blur))


# ----------------------------------------------------------------------------------------- linear
class _LinearFwd(Function):
    """y = act(alpha * x.w^T + bias_scale * bias)   (LinearEx.forward, utils/custom_layers.py:282-291)."""

    @staticmethod
    def forward(ctx, x, w, bias, alpha, bias_scale, act, slope):
        y = K.linear_fwd(x, w, bias, alpha, bias_scale, act, slope)
        ctx.cfg = (alpha, bias_scale, act, slope, bias is not None)
        ctx.save_for_backward(x, w, y if act != ACT_NONE else None)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w, y = ctx.saved_tensors
        alpha, bias_scale, act, slope, has_bias = ctx.cfg
        pg = _param_grads_wanted()
        want_b = has_bias and ctx.needs_input_grad[2] and pg
        if act != ACT_NONE:
            g, gb = act_bwd(gy, y, want_b, bias_scale, act, slope)
        else:
            g = gy
            gb = _ColSum.apply(g, bias_scale) if want_b else None
        gx = _LinearDgrad.apply(g, w, alpha) if ctx.needs_input_grad[0] else None
        gw = _LinearWgrad.apply(x, g, alpha) if (ctx.needs_input_grad[1] and pg) else None
        return gx, gw, gb, None, None, None, None


class _LinearDgrad(Function):
    @staticmethod
    def forward(ctx, gy, w, alpha):
        ctx.alpha = alpha
        ctx.save_for_backward(gy, w)
        return K.linear_dgrad(gy, w, alpha)

    @staticmethod
    def backward(ctx, ggx):
        gy, w = ctx.saved_tensors
        g_gy = _LinearFwd.apply(ggx, w, None, ctx.alpha, 1.0, ACT_NONE, 0.0) if ctx.needs_input_grad[0] else None
        g_w = _LinearWgrad.apply(ggx, gy, ctx.alpha) if ctx.needs_input_grad[1] else None
        return g_gy, g_w, None


class _LinearWgrad(Function):
    @staticmethod
    def forward(ctx, x, gy, alpha):
        ctx.alpha = alpha
        ctx.save_for_backward(x, gy)
        return K.linear_wgrad(x, gy, alpha)

    @staticmethod
    def backward(ctx, ggw):
        x, gy = ctx.saved_tensors
        g_x = _LinearDgrad.apply(gy, ggw, ctx.alpha) if ctx.needs_input_grad[0] else None
        g_gy = _LinearFwd.apply(x, ggw, None, ctx.alpha, 1.0, ACT_NONE, 0.0) if ctx.needs_input_grad[1] else None
        return g_x, g_gy, None


def linear(x, w, bias=None, alpha=1.0, bias_scale=1.0, act=ACT_NONE, slope=0.2):
    return _LinearFwd.apply(x, w, bias, float(alpha), float(bias_scale), int(act), float(slope))


class _GroupedLinear(Function):
    """styles_l = alpha_l * ws[l] . W_l^T + bias_scale_l * b_l for all layers l in one launch per direction (the style affines of
    a StyleGenerator pass, stylegan/architectures.py:460 / 524).  Generator side only -> first order."""

    @staticmethod
    def forward(ctx, ws, tab, *params):
        y = K.glinear_fwd(ws, tab)
        M = ws.shape[1]
        ctx.tab = tab
        ctx.has_bias = [params[2 * i + 1] is not None for i in range(len(tab.nouts))]
        ctx.save_for_backward(ws)
        return tuple(y[M * o:M * (o + n)].view(M, n) for o, n in zip(tab.offs, tab.nouts))

    @staticmethod
    @once_differentiable
    def backward(ctx, *gs):
        (ws,) = ctx.saved_tensors
        tab = ctx.tab
        L, M, Kf = ws.shape
        g_all = torch.cat([g if g is not None else ws.new_zeros((M, n)) for g, n in zip(gs, tab.nouts)], dim=1)
        g_ws = K.glinear_dgrad(g_all, tab, L, M, Kf) if ctx.needs_input_grad[0] else None
        gw, gb = K.glinear_wgrad(ws, g_all, tab)
        out = [g_ws, None]
        for i, (o, n) in enumerate(zip(tab.offs, tab.nouts)):
            out.append(gw[o:o + n] if ctx.needs_input_grad[2 + 2 * i] else None)
            out.append(gb[o:o + n] if (ctx.has_bias[i] and ctx.needs_input_grad[3 + 2 * i]) else None)
        return tuple(out)


def grouped_linear(ws, tab, layers):
    """layers: list of (weight, bias or None, alpha, bias_scale); ws [L,M,K].  -> tuple of [M, nout_l] tensors."""
    tab.update(layers, ws.device)
    flat = []
    for w, b, _a, _bs in layers:
        flat += [w, b]
    return _GroupedLinear.apply(ws, tab, *flat)


# ----------------------------------------------------------------------------------------- stencil / resampling
class _Blur(Function):
    """Self-adjoint 3x3 binomial FIR (get_blur_op, utils/custom_layers.py:36-53): fwd = bwd = double-bwd."""

    @staticmethod
    def forward(ctx, x):
        return K.blur3x3(x)

    @staticmethod
    def backward(ctx, g):
        return _Blur.apply(g)


def blur3x3(x):
    return _Blur.apply(x)


class _Up2(Function):
    @staticmethod
    def forward(ctx, x):
        return K.upsample2x_fwd(x)

    @staticmethod
    def backward(ctx, g):
        return _Up2Bwd.apply(g)


class _Up2Bwd(Function):
    @staticmethod
    def forward(ctx, g):
        return K.upsample2x_bwd(g)

    @staticmethod
    def backward(ctx, gg):
        return _Up2.apply(gg)


def upsample2x(x):
    return _Up2.apply(x)


class _UpConvFprop(Function):
    """y = act(alpha * conv3x3_same(upsample2x(x), w) + bias_scale * bias) with the upsample folded into the convolution
    (glb_upconv_*: four 2x2 phase convolutions on the low-resolution map; reference `nn.Upsample` + `Conv2dEx`,
    stylegan/architectures.py:155-156).  Backward: fused data / weight gradients on the plain backward pass; when a graph of
    the backward is asked for (double backward) the differentiable two-kernel composition is used instead."""

    @staticmethod
    def forward(ctx, x, w, bias, alpha, bias_scale, act, slope):
        y = K.upconv_fprop(x, w, bias, alpha, bias_scale, act, slope, need_dgrad=ctx.needs_input_grad[0])
        ctx.cfg = (alpha, bias_scale, act, slope, bias is not None, x.shape, w.shape)
        ctx.bshape = None if bias is None else bias.shape
        ctx.save_for_backward(x, w, y if act != ACT_NONE else None)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w, y = ctx.saved_tensors
        alpha, bias_scale, act, slope, has_bias, xshape, wshape = ctx.cfg
        pg = _param_grads_wanted()
        want_b = has_bias and ctx.needs_input_grad[2] and pg
        if act != ACT_NONE:
            g, gb = act_bwd(gy, y, want_b, bias_scale, act, slope)
        else:
            g = gy
            gb = _ColSum.apply(g, bias_scale) if want_b else None
        N, Ci, H, W = xshape
        Co = wshape[0]
        fused = not torch.is_grad_enabled()
        gx = gw = None
        if ctx.needs_input_grad[0]:
            if fused and K.upconv_covers("dgrad", N, H, W, Ci, Co):
                gx = K.upconv_dgrad(g, w, alpha)
            else:
                gx = _Up2Bwd.apply(_ConvDgrad.apply(g, w, 2 * H, 2 * W, 1, alpha))
        if ctx.needs_input_grad[1] and pg:
            if fused and K.upconv_covers("wgrad", N, H, W, Ci, Co):
                gw = K.upconv_wgrad(x, g, alpha)
            else:
                gw = _ConvWgrad.apply(_Up2.apply(x), g, 3, 3, 1, alpha)
        return gx, gw, (gb.view(ctx.bshape) if gb is not None else None), None, None, None, None


def upconv2d(x, w, bias=None, alpha=1.0, bias_scale=1.0, act=ACT_NONE, slope=0.2):
    """conv2d(upsample2x(x), w, pad=1) for a 3x3 weight; one fused launch where the tensor-core path covers the shape."""
    N, Ci, H, W = x.shape
    if w.shape[2] == 3 and w.shape[3] == 3 and x.is_cuda and K.upconv_covers("fprop", N, H, W, Ci, w.shape[0]):
        return _UpConvFprop.apply(x, w, bias, float(alpha), float(bias_scale), int(act), float(slope))
    return conv2d(upsample2x(x), w, bias, 1, alpha, bias_scale, act, slope)


class _DownConvFprop(Function):
    """y = act(alpha * avgpool2x2(conv3x3_same(x, w)) + bias_scale * bias): the discriminator's conv -> AvgPool2d -> bias ->
    LeakyReLU (reference progan/architectures.py:267-284) as one stride-2 4x4 convolution (glb_downconv_*).  {fprop, dgrad,
    wgrad} of it is closed under differentiation like the plain convolution family, so the R1 / WGAN-GP double backward
    stays on the fused kernels."""

    @staticmethod
    def forward(ctx, x, w, bias, alpha, bias_scale, act, slope):
        y = K.downconv_fprop(x, w, bias, alpha, bias_scale, act, slope, need_dgrad=ctx.needs_input_grad[0])
        ctx.cfg = (alpha, bias_scale, act, slope, bias is not None)
        ctx.bshape = None if bias is None else bias.shape
        ctx.save_for_backward(x, w, y if act != ACT_NONE else None)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w, y = ctx.saved_tensors
        alpha, bias_scale, act, slope, has_bias = ctx.cfg
        pg = _param_grads_wanted()
        want_b = has_bias and ctx.needs_input_grad[2] and pg
        if act != ACT_NONE:
            g, gb = act_bwd(gy, y, want_b, bias_scale, act, slope)
        else:
            g = gy
            gb = _ColSum.apply(g, bias_scale) if want_b else None
        gx = _DownConvDgrad.apply(g, w, alpha) if ctx.needs_input_grad[0] else None
        gw = _DownConvWgrad.apply(x, g, alpha) if (ctx.needs_input_grad[1] and pg) else None
        return gx, gw, (gb.view(ctx.bshape) if gb is not None else None), None, None, None, None


class _DownConvDgrad(Function):
    """gx = alpha * conv_transpose(0.25 * upsample2x(gy), w).  Bilinear in (gy, w)."""

    @staticmethod
    def forward(ctx, gy, w, alpha):
        ctx.alpha = alpha
        ctx.save_for_backward(gy, w)
        return K.downconv_dgrad(gy, w, alpha)

    @staticmethod
    def backward(ctx, ggx):
        gy, w = ctx.saved_tensors
        g_gy = _DownConvFprop.apply(ggx, w, None, ctx.alpha, 1.0, ACT_NONE, 0.0) if ctx.needs_input_grad[0] else None
        g_w = _DownConvWgrad.apply(ggx, gy, ctx.alpha) if ctx.needs_input_grad[1] else None
        return g_gy, g_w, None


class _DownConvWgrad(Function):
    """gw = alpha * sum_pixels (0.25 * upsample2x(gy)) (x) x.  Bilinear in (x, gy)."""

    @staticmethod
    def forward(ctx, x, gy, alpha):
        ctx.alpha = alpha
        ctx.save_for_backward(x, gy)
        return K.downconv_wgrad(x, gy, alpha)

    @staticmethod
    def backward(ctx, ggw):
        x, gy = ctx.saved_tensors
        g_x = _DownConvDgrad.apply(gy, ggw, ctx.alpha) if ctx.needs_input_grad[0] else None
        g_gy = _DownConvFprop.apply(x, ggw, None, ctx.alpha, 1.0, ACT_NONE, 0.0) if ctx.needs_input_grad[1] else None
        return g_x, g_gy, None


def downconv2d(x, w, bias=None, alpha=1.0, bias_scale=1.0, act=ACT_NONE, slope=0.2):
    """act(alpha * avgpool2x2(conv2d(x, w, pad=1)) + bias_scale * bias) for a 3x3 weight: one fused launch where the
    tensor-core path covers the layer, conv2d -> pool_bias_act otherwise."""
    N, Ci, H, W = x.shape
    if (w.shape[2] == 3 and w.shape[3] == 3 and x.is_cuda and H % 2 == 0 and W % 2 == 0
            and K.downconv_covers(N, H // 2, W // 2, Ci, w.shape[0])):
        return _DownConvFprop.apply(x, w, bias, float(alpha), float(bias_scale), int(act), float(slope))
    return pool_bias_act(conv2d(x, w, None, 1, alpha), bias, bias_scale, act, slope)


class _AvgPoolBwd(Function):
    """g [N,C,H/2,W/2] -> 0.25 * nearest-upsample(g): adjoint of the 2x2 average pool."""

    @staticmethod
    def forward(ctx, g):
        gx, _ = K.pool_bias_act_bwd(g, None, False, 1.0, ACT_NONE, 0.0)
        return gx

    @staticmethod
    def backward(ctx, gg):
        return _PoolBiasAct.apply(gg, None, 1.0, ACT_NONE, 0.0)


class _PoolBiasAct(Function):
    """y = act(avgpool2(x) + bias_scale*bias): D downsampling tail, reference progan/architectures.py:267-284."""

    @staticmethod
    def forward(ctx, x, bias, bias_scale, act, slope):
        y = K.pool_bias_act_fwd(x, bias, bias_scale, act, slope)
        ctx.cfg = (bias_scale, act, slope, bias is not None, None if bias is None else bias.shape)
        ctx.save_for_backward(y if act != ACT_NONE else None)
        return y

    @staticmethod
    def backward(ctx, gy):
        (y,) = ctx.saved_tensors
        bias_scale, act, slope, has_bias, bshape = ctx.cfg
        want_b = has_bias and ctx.needs_input_grad[1] and _param_grads_wanted()
        if act != ACT_NONE:
            g, gb = act_bwd(gy, y, want_b, bias_scale, act, slope)
        else:
            g = gy
            gb = _ColSum.apply(g, bias_scale) if want_b else None
        gx = _AvgPoolBwd.apply(g) if ctx.needs_input_grad[0] else None
        return gx, (gb.view(bshape) if gb is not None else None), None, None, None


def pool_bias_act(x, bias=None, bias_scale=1.0, act=ACT_NONE, slope=0.2):
    return _PoolBiasAct.apply(x, bias, float(bias_scale), int(act), float(slope))


def avgpool2(x):
    return _PoolBiasAct.apply(x, None, 1.0, ACT_NONE, 0.0)


# ----------------------------------------------------------------------------------------- norms
class _PixelNorm(Function):
    """PixelNorm2d (utils/custom_layers.py:81-86); generator-side only -> first order."""

    @staticmethod
    def forward(ctx, x, eps):
        ctx.eps = eps
        ctx.save_for_backward(x)
        return K.pixelnorm_fwd(x, eps)

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        (x,) = ctx.saved_tensors
        return K.pixelnorm_bwd(gy, x, ctx.eps), None


def pixelnorm(x, eps=1e-8):
    return _PixelNorm.apply(x, float(eps))


class _StyleEpilogue(Function):
    """noise + bias + lrelu + InstanceNorm + AdaIN in one op (stylegan/architectures.py:112-119, 255/263/333,
    460-462/524-526).  Generator-side only -> first order."""

    @staticmethod
    def forward(ctx, x, noise, noise_weight, bias, style, slope, eps):
        out, stats = K.style_epilogue_fwd(x, noise, noise_weight, bias, style, slope, eps)
        ctx.slope = slope
        ctx.shapes = (None if noise_weight is None else noise_weight.shape, None if bias is None else bias.shape)
        ctx.save_for_backward(x, noise, noise_weight, bias, style, stats)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, gout):
        x, noise, nw, bias, style, stats = ctx.saved_tensors
        gx, gstyle, g_nw, g_b = K.style_epilogue_bwd(gout, x, noise, nw, bias, style, stats, ctx.slope)
        nws, bs = ctx.shapes
        return (gx, None, g_nw.view(nws) if g_nw is not None else None, g_b.view(bs) if g_b is not None else None,
                gstyle, None, None)


def style_epilogue(x, noise, noise_weight, bias, style, slope=0.2, eps=1e-8):
    return _StyleEpilogue.apply(x, noise, noise_weight, bias, style, float(slope), float(eps))


def instance_norm(x, eps=1e-8):
    """nn.InstanceNorm2d(eps=1e-8, affine=False) through the same fused kernel (identity act, zero style)."""
    style = x.new_zeros((x.shape[0], 2 * x.shape[1]))
    return _StyleEpilogue.apply(x, None, None, None, style, 1.0, float(eps))


# ----------------------------------------------------------------------------------------- ResNet-GAN norms
class _LayerNormAct(Function):
    """y = act(LayerNorm([C,H,W], elementwise affine)(x)): discriminator ResBlocks, resnetgan/resblocks.py:43-46 through
    NormalizeLayer (utils/custom_layers.py:103-106) + the shared ReLU.  Twice differentiable (WGAN-GP)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps, act, slope):
        y, stats = K.layernorm_fwd(x, gamma, beta, eps, act, slope)
        ctx.cfg = (act, slope)
        ctx.save_for_backward(x, gamma, y, stats)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, gamma, y, stats = ctx.saved_tensors
        act, slope = ctx.cfg
        want_p = (ctx.needs_input_grad[1] or ctx.needs_input_grad[2]) and _param_grads_wanted()
        gx, gg, gb = _LayerNormActBwd.apply(gy, y, x, gamma, stats, act, slope, ctx.needs_input_grad[0], want_p)
        return gx, gg, gb, None, None, None


class _LayerNormActBwd(Function):
    """(gy, x, gamma) -> (gx, ggamma, gbeta); y only provides the activation mask, stats are constants of x's forward
    that the second-order kernel differentiates through analytically (csrc/norm.cu)."""

    @staticmethod
    def forward(ctx, gy, y, x, gamma, stats, act, slope, want_gx, want_params):
        ctx.cfg = (act, slope)
        ctx.save_for_backward(gy, y, x, gamma, stats)
        ctx.set_materialize_grads(False)
        gx, gg, gb = K.layernorm_bwd(gy, y, x, gamma, stats, act, slope, want_gx, want_params)
        return gx, gg, gb

    @staticmethod
    @once_differentiable
    def backward(ctx, u, u_gamma, u_beta):
        if u_gamma is not None or u_beta is not None:
            raise NotImplementedError("differentiating LayerNorm's parameter gradients is not on any gan-lab path")
        if u is None:
            return (None,) * 9
        gy, y, x, gamma, stats = ctx.saved_tensors
        act, slope = ctx.cfg
        g_gy, g_x, g_gamma = K.layernorm_bwdbwd(u, gy, y, x, gamma, stats, act, slope, ctx.needs_input_grad[0],
                                                ctx.needs_input_grad[2], ctx.needs_input_grad[3] and _param_grads_wanted())
        return g_gy, None, g_x, g_gamma, None, None, None, None, None


def layernorm_act(x, gamma, beta, eps=1e-5, act=ACT_NONE, slope=0.0):
    return _LayerNormAct.apply(x, gamma, beta, float(eps), int(act), float(slope))


class _BatchNormAct(Function):
    """y = act(BatchNorm2d(x)) with batch statistics + running-buffer update: generator ResBlocks (resblocks.py:43-46 through
    utils/custom_layers.py:100-102).  Generator side only -> first order."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, nbt, eps, momentum, act, slope):
        y, stats = K.batchnorm_fwd(x, gamma, beta, running_mean, running_var, nbt, eps, momentum, act, slope)
        ctx.cfg = (act, slope)
        ctx.save_for_backward(x, gamma, y, stats)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        x, gamma, y, stats = ctx.saved_tensors
        act, slope = ctx.cfg
        gx, gg, gb = K.batchnorm_bwd(gy, y, x, gamma, stats, act, slope)
        return gx, gg, gb, None, None, None, None, None, None, None


def batchnorm_act(x, gamma, beta, running_mean, running_var, num_batches_tracked, eps=1e-5, momentum=0.1, act=ACT_NONE,
                  slope=0.0):
    return _BatchNormAct.apply(x, gamma, beta, running_mean, running_var, num_batches_tracked, float(eps), float(momentum),
                               int(act), float(slope))


class _Tanh(Function):
    """nn.Tanh on the ResNet generator's output (resnetgan/architectures.py:58, 96); first order."""

    @staticmethod
    def forward(ctx, x):
        y = K.tanh_fwd(x)
        ctx.save_for_backward(y)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        (y,) = ctx.saved_tensors
        return K.tanh_bwd(gy, y)


def tanh(x):
    return _Tanh.apply(x)


# ----------------------------------------------------------------------------------------- minibatch stddev
def mbstd_group(n: int, group_size: int) -> int:
    """Group-size rule of concat_mbstd_layer (utils/custom_layers.py:121-126)."""
    g = min(n, group_size)
    if n % g != 0:
        g = n
    return g


class _Mbstd(Function):
    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        ctx.save_for_backward(x)
        return K.mbstd_fwd(x, group)

    @staticmethod
    def backward(ctx, gy):
        (x,) = ctx.saved_tensors
        return _MbstdBwd.apply(gy, x, ctx.group), None


class _MbstdBwd(Function):
    @staticmethod
    def forward(ctx, gy, x, group):
        ctx.group = group
        ctx.save_for_backward(gy, x)
        return K.mbstd_bwd(gy, x, group)

    @staticmethod
    @once_differentiable
    def backward(ctx, v):
        gy, x = ctx.saved_tensors
        ggx, ggy = K.mbstd_bwdbwd(v, gy, x, ctx.group)
        return ggy, ggx, None


def mbstd_concat(x, group_size=4):
    return _Mbstd.apply(x, mbstd_group(x.shape[0], group_size))


# ----------------------------------------------------------------------------------------- RGB 1x1 convs
class _RgbExpand(Function):
    """img NCHW (3 ch) -> NHWC features: y = act(alpha * img.w + bias_scale*bias); w element (j,c) at w[j*ws_j+c*ws_c]."""

    @staticmethod
    def forward(ctx, img, w, bias, ws_j, ws_c, C, pool, alpha, bias_scale, act, slope):
        y = K.rgb_expand(img, w, ws_j, ws_c, C, bias, pool, alpha, bias_scale, act, slope)
        ctx.cfg = (ws_j, ws_c, C, pool, alpha, bias_scale, act, slope, bias is not None, w.shape,
                   None if bias is None else bias.shape)
        ctx.save_for_backward(img, w, y if act != ACT_NONE else None)
        return y

    @staticmethod
    def backward(ctx, gy):
        img, w, y = ctx.saved_tensors
        ws_j, ws_c, C, pool, alpha, bias_scale, act, slope, has_bias, wshape, bshape = ctx.cfg
        pg = _param_grads_wanted()
        want_b = has_bias and ctx.needs_input_grad[2] and pg
        if act != ACT_NONE:
            g, gb = act_bwd(gy, y, want_b, bias_scale, act, slope)
        else:
            g = gy
            gb = _ColSum.apply(g, bias_scale) if want_b else None
        gimg = _RgbContract.apply(g, w, None, ws_j, ws_c, pool, alpha, 1.0) if ctx.needs_input_grad[0] else None
        gw = _RgbWgrad.apply(img, g, wshape, ws_j, ws_c, pool, alpha) if (ctx.needs_input_grad[1] and pg) else None
        return gimg, gw, (gb.view(bshape) if gb is not None else None), None, None, None, None, None, None, None, None


class _RgbContract(Function):
    """NHWC features -> img NCHW (3 ch): img = alpha * sum_c x*w + bias_scale*bias (pool: avg-pool adjoint scatter)."""

    @staticmethod
    def forward(ctx, x, w, bias, ws_j, ws_c, pool, alpha, bias_scale):
        ctx.cfg = (ws_j, ws_c, pool, alpha, bias_scale, bias is not None, w.shape, x.shape[1],
                   None if bias is None else bias.shape)
        ctx.save_for_backward(x, w)
        return K.rgb_contract(x, w, ws_j, ws_c, bias, pool, alpha, bias_scale)

    @staticmethod
    def backward(ctx, gimg):
        x, w = ctx.saved_tensors
        ws_j, ws_c, pool, alpha, bias_scale, has_bias, wshape, C, bshape = ctx.cfg
        pg = _param_grads_wanted()
        gx = _RgbExpand.apply(gimg, w, None, ws_j, ws_c, C, pool, alpha, 1.0, ACT_NONE, 0.0) if ctx.needs_input_grad[0] else None
        gw = _RgbWgrad.apply(gimg, x, wshape, ws_j, ws_c, pool, alpha) if (ctx.needs_input_grad[1] and pg) else None
        gb = None
        if has_bias and ctx.needs_input_grad[2] and pg:
            gb = K.plane_sum(gimg, bias_scale).view(bshape)
        return gx, gw, gb, None, None, None, None, None


class _RgbWgrad(Function):
    @staticmethod
    def forward(ctx, img, g, wshape, ws_j, ws_c, pool, alpha):
        ctx.cfg = (ws_j, ws_c, pool, alpha, g.shape[1])
        ctx.save_for_backward(img, g)
        return K.rgb_wgrad(img, g, wshape, ws_j, ws_c, pool, alpha)

    @staticmethod
    def backward(ctx, ggw):
        img, g = ctx.saved_tensors
        ws_j, ws_c, pool, alpha, C = ctx.cfg
        g_img = _RgbContract.apply(g, ggw, None, ws_j, ws_c, pool, alpha, 1.0) if ctx.needs_input_grad[0] else None
        g_g = _RgbExpand.apply(img, ggw, None, ws_j, ws_c, C, pool, alpha, 1.0, ACT_NONE, 0.0) if ctx.needs_input_grad[1] else None
        return g_img, g_g, None, None, None, None, None


def fromrgb(img, w, bias, alpha, bias_scale=1.0, act=ACT_LRELU, slope=0.2, pool=False):
    """Conv2dEx 1x1 (3 -> C) [+ lrelu]; weight [C,3,1,1].  pool=True averages 2x2 first (fade-in skip branch)."""
    C = w.shape[0]
    return _RgbExpand.apply(img, w, bias, 1, 3, C, bool(pool), float(alpha), float(bias_scale), int(act), float(slope))


def torgb(x, w, bias, alpha, bias_scale=1.0):
    """Conv2dEx 1x1 (C -> 3); weight [3,C,1,1]; output NCHW image."""
    C = w.shape[1]
    return _RgbContract.apply(x, w, bias, C, 1, False, float(alpha), float(bias_scale))


# ----------------------------------------------------------------------------------------- image-space fade-in
class _FadeUpBlend(Function):
    """(1-alpha) * up2x(lo) + alpha * hi on NCHW images (stylegan/architectures.py:482-487). Generator-side."""

    @staticmethod
    def forward(ctx, lo, hi, alpha):
        ctx.alpha = alpha
        return K.fade_up_blend(lo, hi, alpha)

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        glo, ghi = K.fade_up_blend_bwd(g, ctx.alpha)
        return glo, ghi, None


def fade_up_blend(lo, hi, alpha):
    return _FadeUpBlend.apply(lo, hi, K.coef_or_float(alpha))


def fade_real(x, alpha):
    """Real-image fade-in (progan/learner.py:770-779) on the device; no gradient."""
    return K.fade_real(x.detach(), K.coef_or_float(alpha))


# ----------------------------------------------------------------------------------------- losses
class _DLogitLoss(Function):
    @staticmethod
    def forward(ctx, d_gen, d_real, kind, eps_drift):
        loss, gg, gr = K.d_logit_loss(d_gen, d_real, kind, eps_drift)
        ctx.save_for_backward(gg, gr)
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, gl):
        gg, gr = ctx.saved_tensors
        return K.scale_by(gg, gl, 1.0), K.scale_by(gr, gl, 1.0), None, None


class _GLogitLoss(Function):
    @staticmethod
    def forward(ctx, d_out, kind):
        loss, g = K.g_logit_loss(d_out, kind)
        ctx.save_for_backward(g)
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, gl):
        (g,) = ctx.saved_tensors
        return K.scale_by(g, gl, 1.0), None


def d_logit_loss(d_gen, d_real, kind="nonsaturating", eps_drift=0.0):
    """progan/learner.py:791-800 (+ drift term :811-812)."""
    return _DLogitLoss.apply(d_gen, d_real, kind, float(eps_drift))


def g_logit_loss(d_out, kind="nonsaturating"):
    """progan/learner.py:883-896."""
    return _GLogitLoss.apply(d_out, kind)


class _SumSq(Function):
    """scale * sum(x^2): the R1/R2 penalty reduction (resnetgan/learner.py:825; norm over channels, mean over N,H,W)."""

    @staticmethod
    def forward(ctx, x, scale):
        ctx.scale = scale
        ctx.save_for_backward(x)
        return K.sumsq(x, scale)

    @staticmethod
    @once_differentiable
    def backward(ctx, gl):
        (x,) = ctx.saved_tensors
        return K.scale_by(x, gl, 2.0 * ctx.scale), None


def sumsq(x, scale):
    return _SumSq.apply(x, float(scale))


class _GpNorm(Function):
    """scale * sum_{n,h,w} (||g||_2 over channels - gamma)^2   (WGAN-GP, resnetgan/learner.py:817-823)."""

    @staticmethod
    def forward(ctx, g, gamma, scale):
        ctx.cfg = (gamma, scale)
        ctx.save_for_backward(g)
        return K.gp_norm_fwd(g, gamma, scale)

    @staticmethod
    @once_differentiable
    def backward(ctx, gl):
        (g,) = ctx.saved_tensors
        gamma, scale = ctx.cfg
        return K.gp_norm_bwd(g, gl, gamma, scale), None, None


def gp_norm(g, gamma, scale):
    return _GpNorm.apply(g, float(gamma), float(scale))
