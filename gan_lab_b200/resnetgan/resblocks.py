"""ResBlocks of the ResNet GANs (reference gan_lab/resnetgan/resblocks.py:15-120) on the sm_100a kernels.

Same constructor signatures, the same `nn.Sequential` containers in the same order (so `state_dict` keys such as
`conv_layer_1.0.norm.weight` / `conv_layer_1.3.conv2d.weight` / `skip_connection.1.conv2d.bias` match the reference's) and
the same arithmetic.  What differs is execution: `run_fused()` walks a container and folds the nonlinearity that follows
a Batch/LayerNorm (or a convolution) into the producing kernel; the residual sum is one axpby kernel.
"""
from torch import nn

from .. import ops
from ..utils.custom_layers import (Lambda, get_blur_op, NormalizeLayer, Conv2dEx, LeakyReLU, AvgPool2x, Upsample2x, as_native_nl,
                                   as_native_upsampler, as_native_pooler)


def run_fused(seq, x):
    """Apply the modules of an nn.Sequential in order, fusing `[NormalizeLayer | Conv2dEx] -> nonlinearity` pairs."""
    mods = list(seq)
    i = 0
    while i < len(mods):
        m = mods[i]
        nxt = mods[i + 1] if i + 1 < len(mods) else None
        fuse = isinstance(nxt, LeakyReLU) and ((isinstance(m, NormalizeLayer) and m.fuses_act) or isinstance(m, Conv2dEx))
        if (isinstance(m, AvgPool2x) and isinstance(nxt, Conv2dEx) and nxt.ks == 1 and nxt.ni == 3 and nxt.nf % 4 == 0
                and x.shape[1] == 3):
            # avg-pool of the 3-channel image + 1x1 conv (FastResBlock2dDownsample's skip branch): one fromRGB kernel
            x = ops.fromrgb(x, nxt.conv2d.weight, nxt.conv2d.bias, nxt.alpha, nxt.lrmul if nxt.use_lrmul else 1.,
                            ops.ACT_NONE, 0.0, pool=True)
            i += 2
        elif isinstance(m, Upsample2x) and isinstance(nxt, Conv2dEx) and nxt.ks == 3 and nxt.padding == 1:
            x = nxt(x, up=True)                         # upsample folded into the 3x3 convolution (ops.upconv2d)
            i += 2
        elif isinstance(m, Upsample2x) and isinstance(nxt, Conv2dEx) and nxt.ks == 1 and nxt.padding == 0:
            x = m(nxt(x))                               # a 1x1 convolution commutes with nearest upsampling: same values, 1/4 of the work
            i += 2
        elif isinstance(m, Conv2dEx) and m.ks == 3 and m.padding == 1 and isinstance(nxt, AvgPool2x) and x.dim() == 4:
            # conv -> 2x2 average pool as one stride-2 4x4 convolution (ops.downconv2d falls back to the two kernels by itself);
            # the bias commutes with the average
            x = ops.downconv2d(x, m.conv2d.weight, m.conv2d.bias, m.alpha, m.lrmul if m.use_lrmul else 1., ops.ACT_NONE, 0.2)
            i += 2
        elif fuse:
            x = m(x, act=ops.ACT_LRELU, slope=nxt.negative_slope)
            i += 2
        else:
            x = m(x)
            i += 1
    return x


class ResBlock2d(nn.Module):
    """Pre-activation residual block (reference resblocks.py:15-65): [norm, nl, (up,) conv0, (blur)] -> [norm, nl, (blur,)
    conv1, (pool)] plus a 1x1-conv skip branch resampled the same way."""

    def __init__(self, ni, nf, ks, norm_type, upsampler=None, pooler=None, init='He', nl=None, res=None,
                 flip_sampling=False, equalized_lr=False, blur_type=None):
        super(ResBlock2d, self).__init__()
        assert not (upsampler is not None and pooler is not None)
        nl = as_native_nl(nl if nl is not None else nn.ReLU())
        upsampler = as_native_upsampler(upsampler) if upsampler is not None else None
        pooler = as_native_pooler(pooler) if pooler is not None else None
        padding = (ks - 1) // 2  # 'SAME' padding for stride 1 conv

        if not flip_sampling:
            self.nif = nf if (upsampler is not None and pooler is None) else ni
        else:
            self.nif = ni if (upsampler is None and pooler is not None) else nf
        self.convs = (
            Conv2dEx(ni, self.nif, ks=ks, stride=1, padding=padding, init=init, equalized_lr=equalized_lr),
            Conv2dEx(self.nif, nf, ks=ks, stride=1, padding=padding, init=init, equalized_lr=equalized_lr),
            Conv2dEx(ni, nf, ks=1, stride=1, padding=0, init='Xavier', equalized_lr=equalized_lr),  # same as a FC layer
        )
        blur_op = get_blur_op(blur_type=blur_type, num_channels=self.convs[0].nf) if blur_type is not None else None
        _norm_nls = (
            [NormalizeLayer(norm_type, ni=ni, res=res), nl],
            [NormalizeLayer(norm_type, ni=self.convs[0].nf, res=res), nl],
        )
        if upsampler is not None:
            _op1 = [upsampler, self.convs[0], blur_op] if blur_type is not None else [upsampler, self.convs[0]]
            _op2 = [upsampler, self.convs[2], blur_op] if blur_type is not None else [upsampler, self.convs[2]]
            _ops = (_op1, [self.convs[1]], _op2,)
        elif pooler is not None:
            _op1 = [blur_op, self.convs[1], pooler] if blur_type is not None else [self.convs[1], pooler]
            _op2 = [blur_op, pooler, self.convs[2]] if blur_type is not None else [pooler, self.convs[2]]
            _ops = ([self.convs[0]], _op1, _op2,)
        else:
            _ops = ([self.convs[0]], [self.convs[1]], [self.convs[2]],)

        self.conv_layer_1 = nn.Sequential(*(_norm_nls[0] + _ops[0]))
        self.conv_layer_2 = nn.Sequential(*(_norm_nls[1] + _ops[1]))
        if (upsampler is not None or pooler is not None) or ni != nf:
            self.skip_connection = nn.Sequential(*(_ops[2]))
        else:
            self.skip_connection = Lambda(lambda x: x)

    def forward(self, x):
        skip = run_fused(self.skip_connection, x) if isinstance(self.skip_connection, nn.Sequential) else x
        return ops.axpby(skip, run_fused(self.conv_layer_2, run_fused(self.conv_layer_1, x)), 1.0, 1.0)


class ResBlock2d32Pix(ResBlock2d):
    """reference resblocks.py:68-79: flipped sampling order; a downsampling block pools AFTER the skip conv."""

    def __init__(self, ni, nf, ks, norm_type, upsampler=None, pooler=None, init='He', nl=None, res=None,
                 flip_sampling=True, equalized_lr=False, blur_type=None):
        super(ResBlock2d32Pix, self).__init__(ni, nf, ks, norm_type, upsampler, pooler, init, nl, res, flip_sampling,
                                              equalized_lr, blur_type)
        if upsampler is None and pooler is not None:
            self.skip_connection = nn.Sequential(self.convs[2], as_native_pooler(pooler))


class FastResBlock2dDownsample(nn.Module):
    """Downsampling ResBlock without normalization and activation before the first conv (reference resblocks.py:82-120)."""

    def __init__(self, ni, nf, ks, pooler=None, init='He', nl=None, equalized_lr=False, blur_type=None):
        super(FastResBlock2dDownsample, self).__init__()
        nl = as_native_nl(nl if nl is not None else nn.ReLU())
        pooler = as_native_pooler(pooler if pooler is not None else nn.AvgPool2d(kernel_size=2, stride=2))
        padding = (ks - 1) // 2
        self.conv_layer_1 = nn.Sequential(
            Conv2dEx(ni, nf, ks=ks, stride=1, padding=padding, init='he', equalized_lr=equalized_lr),
            nl
        )
        self.conv_layer_2 = nn.Sequential()
        self.skip_connection = nn.Sequential()
        _seq_n = 0
        if blur_type is not None:
            blur_op = get_blur_op(blur_type=blur_type, num_channels=nf)
            self.conv_layer_2.add_module(str(_seq_n), blur_op)
            self.skip_connection.add_module(str(_seq_n), blur_op)
            _seq_n += 1
        self.conv_layer_2.add_module(
            str(_seq_n), Conv2dEx(nf, nf, ks=ks, stride=1, padding=padding, init='he', equalized_lr=equalized_lr))
        self.skip_connection.add_module(str(_seq_n), pooler)
        _seq_n += 1
        self.conv_layer_2.add_module(str(_seq_n), pooler)
        self.skip_connection.add_module(
            str(_seq_n), Conv2dEx(ni, nf, ks=1, stride=1, padding=0, init='xavier', equalized_lr=equalized_lr))

    def forward(self, x):
        return ops.axpby(run_fused(self.skip_connection, x), run_fused(self.conv_layer_2, run_fused(self.conv_layer_1, x)),
                         1.0, 1.0)
