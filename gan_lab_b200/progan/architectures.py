"""ProGAN generator and the ProGAN/StyleGAN discriminator (reference gan_lab/progan/architectures.py).

Same module tree and state_dict keys as the reference (`gen_blocks.{b}...`, `disc_blocks.{b}.0.0.conv2d.*`,
`disc_blocks.{b}.1.1.conv2d.weight`, `disc_blocks.{b}.1.3.bias`, `fromrgb.0.conv2d.*`, `prev_fromrgb...`,
`torgb.conv2d.*`, `prev_torgb...`); the `nn.Sequential` containers are subclassed so that their forward runs
the fused kernels (conv + bias + lrelu in one epilogue, avg-pool + bias + lrelu in one pass, fromRGB as a
bandwidth kernel) instead of one ATen call per child.
"""
import copy

import functools

import torch
from torch import nn

from .. import ops
from .._growth import GrowthState
from ..utils.custom_layers import (Lambda, get_blur_op, NormalizeLayer, concat_mbstd_layer, Conv2dEx, LinearEx,
                                   Conv2dBias, Blur3x3, Upsample2x, AvgPool2x, LeakyReLU, as_native_nl,
                                   as_native_upsampler, as_native_pooler)
from .base import ProGAN

FMAP_SAMPLES = 3   # reference _int.py:46
RES_INIT = 4       # reference _int.py:47
FMAP_G_INIT_FCTR = 1
FMAP_D_END_FCTR = 1


def _slope(nl):
    return nl.negative_slope


class ConvLayer(nn.Sequential):
    """One `get_conv_layer` of the reference (progan/architectures.py:109-148 G side, :261-284 D side):
    children in the reference's order  [upsampler?] [blur?] conv [blur?] [pooler?] [bias?] [nl?] [norm?];
    forward fuses them:
      G:  up -> conv(no bias) -> blur -> bias+lrelu -> pixelnorm      |  conv+bias+lrelu -> pixelnorm
      D:  blur -> conv(no bias) -> avgpool+bias+lrelu                 |  conv+bias+lrelu
    """

    def forward(self, x, blur_after=False, skip_pre_blur=False):
        """blur_after: the binomial FIR of the NEXT layer rides in this layer's conv op (DiscBlock);
        skip_pre_blur: ... and that next layer then skips its own pre-blur."""
        mods = list(self)
        i = 0
        n = len(mods)
        up = False
        while i < n and not isinstance(mods[i], Conv2dEx):     # upsampler / pre-blur
            if isinstance(mods[i], Upsample2x) and i + 1 < n and isinstance(mods[i + 1], Conv2dEx):
                up = True                                      # rides in the convolution (ops.upconv2d)
            elif not (skip_pre_blur and isinstance(mods[i], Blur3x3)):
                x = mods[i](x)
            i += 1
        conv = mods[i]
        if up:
            conv = functools.partial(conv, up=True)
        i += 1
        rest = mods[i:]
        post_blur = rest[0] if rest and isinstance(rest[0], Blur3x3) else None
        if post_blur is not None:
            rest = rest[1:]
        pool = rest[0] if rest and isinstance(rest[0], AvgPool2x) else None
        if pool is not None:
            rest = rest[1:]
        bias = rest[0] if rest and isinstance(rest[0], Conv2dBias) else None
        if bias is not None:
            rest = rest[1:]
        nl = rest[0] if rest and isinstance(rest[0], LeakyReLU) else None
        if nl is not None:
            rest = rest[1:]
        act = ops.ACT_LRELU if nl is not None else ops.ACT_NONE
        slope = _slope(nl) if nl is not None else 0.2
        if post_blur is None and pool is None and bias is None:
            x = conv(x, act=act, slope=slope, blur=blur_after)                    # conv + bias + lrelu (+ blur), one op
        elif (pool is not None and post_blur is None and not up and conv.ks == 3 and conv.padding == 1 and conv.conv2d.bias is None
              and (bias is None or not (bias.bias is None))):
            # conv -> avgpool -> bias -> lrelu as one stride-2 4x4 convolution (ops.downconv2d; falls back by itself)
            x = ops.downconv2d(x, conv.conv2d.weight, bias.bias if bias is not None else None, conv.alpha,
                               bias.bias_scale if bias is not None else 1., act, slope)
        else:
            x = conv(x)
            if post_blur is not None:
                x = post_blur(x)
            if pool is not None:
                x = ops.pool_bias_act(x, bias.bias if bias is not None else None,
                                      bias.bias_scale if bias is not None else 1., act, slope)
            elif bias is not None or nl is not None:
                x = ops.bias_act(x, bias.bias if bias is not None else None,
                                 bias.bias_scale if bias is not None else 1., act, slope)
        for m in rest:                                                            # pixelnorm / non-fusable tail
            x = m(x)
        return x


class DiscBlock(nn.Sequential):
    """One discriminator block (reference progan/architectures.py:236-259): [conv + bias + lrelu] -> [blur -> conv ->
    avgpool + bias + lrelu].  Same children / state_dict keys as the reference's nn.Sequential of two conv layers; the
    forward moves the second layer's pre-blur into the first layer's op, whose backward then runs blur + activation mask +
    bias gradient as one pass (glb_blur_act_bwd) instead of blur3x3 followed by act_bwd."""

    fuse_blur = True      # class-wide switch (A/B runs)

    def forward(self, x):
        first, second = self[0], self[1]
        fuse = (self.fuse_blur and len(self) == 2 and isinstance(first, ConvLayer) and isinstance(second, ConvLayer) and len(second) > 0 and
                isinstance(second[0], Blur3x3) and len(first) == 2 and isinstance(first[0], Conv2dEx) and
                isinstance(first[1], LeakyReLU))
        if not fuse:
            for m in self:
                x = m(x)
            return x
        return second(first(x, blur_after=True), skip_pre_blur=True)


class FromRGB(nn.Sequential):
    """Sequential(Conv2dEx 1x1 3->C, nl) (reference progan/architectures.py:286-292) as one bandwidth kernel;
    `pool=True` folds the fade-in skip branch's avg_pool2d of the image (:312) into the same kernel."""

    def forward(self, x, pool=False):
        conv, nl = self[0], self[1]
        bscale = conv.lrmul if conv.use_lrmul else 1.
        return ops.fromrgb(x, conv.conv2d.weight, conv.conv2d.bias, conv.alpha, bscale, ops.ACT_LRELU, _slope(nl), pool=pool)


class GenFirstBlock(nn.Sequential):
    """ProGAN generator block 0 (reference progan/architectures.py:78-93):
    LinearEx -> view(512,4,4) -> nl -> [PN] -> Conv2dEx -> nl -> [PN]."""

    def forward(self, x):
        mods = list(self)
        fc, view, nl = mods[0], mods[1], mods[2]
        x = view(fc(x))
        # the lrelu sits after the view (channel-major), apply it on the NHWC tensor
        x = ops.bias_act(x, None, 1.0, ops.ACT_LRELU, _slope(nl))
        i = 3
        if isinstance(mods[i], NormalizeLayer):
            x = mods[i](x)
            i += 1
        conv = mods[i]
        x = conv(x, act=ops.ACT_LRELU, slope=_slope(mods[i + 1]))
        for m in mods[i + 2:]:
            x = m(x)
        return x


class DiscLastBlock(nn.Sequential):
    """Last discriminator block (reference progan/architectures.py:219-233):
    [mbstd] -> Conv2dEx 3x3 (+1 ch) -> nl -> Conv2dEx 4x4 valid -> nl -> view -> LinearEx(->1)."""

    def forward(self, x):
        mods = list(self)
        i = 0
        if isinstance(mods[0], Lambda):
            x = mods[0](x)
            i = 1
        x = mods[i](x, act=ops.ACT_LRELU, slope=_slope(mods[i + 1]))
        x = mods[i + 2](x, act=ops.ACT_LRELU, slope=_slope(mods[i + 3]))
        x = mods[i + 4](x)
        return mods[i + 5](x)


# ++++++++++++++++++++++++++++++++++++++++++++++++++++++++++++++++++++++++++++ #
# Generator
# ++++++++++++++++++++++++++++++++++++++++++++++++++++++++++++++++++++++++++++ #
class ProGenerator(ProGAN):
    """Progressively Growing GAN generator (reference progan/architectures.py:31-167)."""

    def __init__(self, final_res, len_latent=512, upsampler=None, blur_type=None, nl=None, num_classes=0,
                 equalized_lr=True, normalize_z=True, use_pixelnorm=True, state: GrowthState = None):
        super(ProGenerator, self).__init__(final_res, state=state)
        self.gen_blocks = nn.ModuleList()
        self.upsampler = as_native_upsampler(upsampler) if upsampler is not None else Upsample2x()
        self.gen_blur_type = blur_type
        self.nl = as_native_nl(nl) if nl is not None else LeakyReLU(.2)
        self.len_latent = len_latent
        self.num_classes = num_classes
        self.equalized_lr = equalized_lr

        norms = []
        self.use_pixelnorm = use_pixelnorm
        if use_pixelnorm:
            norms.append(NormalizeLayer('PixelNorm'))
        if normalize_z:
            self.preprocess_z = nn.Sequential(Lambda(lambda x: x.view(-1, len_latent + num_classes)),
                                              NormalizeLayer('PixelNorm'))
        else:
            self.preprocess_z = Lambda(lambda x: x.view(-1, len_latent + num_classes))

        _fmap_init = len_latent * FMAP_G_INIT_FCTR
        self.gen_blocks.append(
            GenFirstBlock(
                LinearEx(nin_feat=len_latent + num_classes, nout_feat=_fmap_init * RES_INIT ** 2, init='He',
                         init_type='ProGAN', gain_sq_base=2. / 16, equalized_lr=equalized_lr),
                Lambda(lambda x: x.view(-1, _fmap_init, RES_INIT, RES_INIT)),
                self.nl,
                *norms,
                Conv2dEx(ni=_fmap_init, nf=self.fmap, ks=3, stride=1, padding=1, init='He', init_type='ProGAN',
                         gain_sq_base=2., equalized_lr=equalized_lr),
                self.nl,
                *norms))
        self.prev_torgb = None
        self._update_torgb(ni=self.fmap)

    def increase_scale(self):
        """reference progan/architectures.py:95-107."""
        if not self.scale_inc_metadata_updated:
            super(ProGenerator, self).increase_scale()
        else:
            self.scale_inc_metadata_updated = False
        blur_op = get_blur_op(blur_type=self.gen_blur_type, num_channels=self.fmap) if self.gen_blur_type is not None else None
        self.gen_blocks.append(nn.Sequential(
            self.get_conv_layer(ni=self.fmap_prev, upsample=True, blur_op=blur_op),
            self.get_conv_layer(ni=self.fmap)))
        self.prev_torgb = copy.deepcopy(self.torgb)
        self._update_torgb(ni=self.fmap)
        self.to(self.torgb_device())

    def torgb_device(self):
        return next(self.gen_blocks[0].parameters()).device

    def get_conv_layer(self, ni, upsample=False, blur_op=None, append_nl=True):
        """reference progan/architectures.py:109-148."""
        upsampler = [self.upsampler] if upsample else []
        if blur_op is not None:
            conv = Conv2dEx(ni=ni, nf=self.fmap, ks=3, stride=1, padding=1, init='He', init_type='ProGAN',
                            gain_sq_base=2., equalized_lr=self.equalized_lr, include_bias=False)
            blur = [blur_op]
            bias = [Conv2dBias(nf=self.fmap, device='cpu')]
        else:
            conv = Conv2dEx(ni=ni, nf=self.fmap, ks=3, stride=1, padding=1, init='He', init_type='ProGAN',
                            gain_sq_base=2., equalized_lr=self.equalized_lr, include_bias=True)
            blur = []
            bias = []
        nl = [self.nl] if append_nl else []
        norm = [NormalizeLayer('PixelNorm')] if self.use_pixelnorm else []
        return ConvLayer(*upsampler, conv, *(blur + bias + nl + norm))

    def _update_torgb(self, ni):
        self.torgb = Conv2dEx(ni=ni, nf=FMAP_SAMPLES, ks=1, stride=1, padding=0, init='He', init_type='ProGAN',
                              gain_sq_base=1., equalized_lr=self.equalized_lr)

    def forward(self, x):
        """reference progan/architectures.py:159-167."""
        x = self.preprocess_z(x)
        for gen_block in self.gen_blocks[:-1]:
            x = gen_block(x)
        if self.fade_in_phase:
            return ops.fade_up_blend(self.prev_torgb(x), self.torgb(self.gen_blocks[-1](x)), self._state.blend_coefs()[0])
        return self.torgb(self.gen_blocks[-1](x))


# ++++++++++++++++++++++++++++++++++++++++++++++++++++++++++++++++++++++++++++ #
# Discriminator
# ++++++++++++++++++++++++++++++++++++++++++++++++++++++++++++++++++++++++++++ #
class _DiscriminatorImpl(object):
    """ProDiscriminator (reference progan/architectures.py:175-318); StyleGAN's discriminator is the same class body
    re-parented onto the StyleGAN state base (reference stylegan/learner.py:142)."""

    def _init_disc(self, pooler, blur_type, nl, num_classes, equalized_lr, mbstd_group_size, init_type):
        self.init_type = init_type
        self.disc_blocks = nn.ModuleList()
        self.num_classes = num_classes
        self.preprocess_x = Lambda(lambda x: x.view(-1, FMAP_SAMPLES + num_classes, self.curr_res, self.curr_res))
        self.pooler = as_native_pooler(pooler) if pooler is not None else AvgPool2x()
        self.disc_blur_type = blur_type
        self.nl = as_native_nl(nl) if nl is not None else LeakyReLU(.2)
        self.equalized_lr = equalized_lr
        self.mbstd_group_size = mbstd_group_size
        mbstd_layer = self.get_mbstd_layer()
        self.prev_fromrgb = None
        self._update_fromrgb(nf=self.fmap)
        _fmap_end = self.fmap * FMAP_D_END_FCTR
        self.disc_blocks.insert(0, DiscLastBlock(
            *mbstd_layer,
            Conv2dEx(ni=self.fmap + (1 if mbstd_layer else 0), nf=self.fmap, ks=3, stride=1, padding=1, init='He',
                     init_type=self.init_type, gain_sq_base=2., equalized_lr=equalized_lr),
            self.nl,
            Conv2dEx(ni=self.fmap, nf=_fmap_end, ks=4, stride=1, padding=0, init='He', init_type=self.init_type,
                     gain_sq_base=2., equalized_lr=equalized_lr),
            self.nl,
            Lambda(lambda x: x.view(-1, _fmap_end)),
            LinearEx(nin_feat=_fmap_end, nout_feat=1, init='He', init_type=self.init_type, gain_sq_base=1.,
                     equalized_lr=equalized_lr)))

    def increase_scale(self):
        """reference progan/architectures.py:236-259."""
        if not self.scale_inc_metadata_updated:
            super(_DiscriminatorImpl, self).increase_scale()
        else:
            self.scale_inc_metadata_updated = False
        self.preprocess_x = Lambda(lambda x: x.view(-1, FMAP_SAMPLES + self.num_classes, self.curr_res, self.curr_res))
        self.prev_fromrgb = copy.deepcopy(self.fromrgb)
        self._update_fromrgb(nf=self.fmap)
        blur_op = get_blur_op(blur_type=self.disc_blur_type, num_channels=self.fmap) if self.disc_blur_type is not None else None
        self.disc_blocks.insert(0, DiscBlock(
            self.get_conv_layer(nf=self.fmap),
            self.get_conv_layer(nf=self.fmap_prev, downsample=True, blur_op=blur_op)))
        self.to(next(self.disc_blocks[-1].parameters()).device)

    def get_conv_layer(self, nf, downsample=False, blur_op=None, append_nl=True):
        """reference progan/architectures.py:261-284."""
        blur = [blur_op] if blur_op is not None else []
        if downsample:
            conv = Conv2dEx(ni=self.fmap, nf=nf, ks=3, stride=1, padding=1, init='He', init_type=self.init_type,
                            gain_sq_base=2., equalized_lr=self.equalized_lr, include_bias=False)
            pooler = [self.pooler]
            bias = [Conv2dBias(nf=nf, device='cpu')]
        else:
            conv = Conv2dEx(ni=self.fmap, nf=nf, ks=3, stride=1, padding=1, init='He', init_type=self.init_type,
                            gain_sq_base=2., equalized_lr=self.equalized_lr, include_bias=True)
            pooler = []
            bias = []
        nl = [self.nl] if append_nl else []
        return ConvLayer(*blur, conv, *(pooler + bias + nl))

    def _update_fromrgb(self, nf):
        self.fromrgb = FromRGB(
            Conv2dEx(ni=FMAP_SAMPLES + self.num_classes, nf=nf, ks=1, stride=1, padding=0, init='He',
                     init_type=self.init_type, gain_sq_base=2., equalized_lr=self.equalized_lr),
            self.nl)

    def get_mbstd_layer(self):
        if self.mbstd_group_size == -1:
            return []
        return [Lambda(lambda x, group_size: concat_mbstd_layer(x, group_size), group_size=self.mbstd_group_size)]

    def forward(self, x):
        """reference progan/architectures.py:309-318."""
        x = self.preprocess_x(x)
        if self.fade_in_phase:
            # prev_fromrgb(avg_pool2d(x))*(1-alpha) + disc_blocks[0](fromrgb(x))*alpha
            alpha, one_minus_alpha = self._state.blend_coefs()
            x = ops.axpby(self.prev_fromrgb(x, pool=True), self.disc_blocks[0](self.fromrgb(x)), one_minus_alpha, alpha)
        else:
            x = self.disc_blocks[0](self.fromrgb(x))
        for disc_block in self.disc_blocks[1:]:
            x = disc_block(x)
        return x.view(-1)


class ProDiscriminator(_DiscriminatorImpl, ProGAN):
    def __init__(self, final_res, pooler=None, blur_type=None, nl=None, num_classes=0, equalized_lr=True,
                 mbstd_group_size=4, state: GrowthState = None):
        ProGAN.__init__(self, final_res, state=state)
        self._init_disc(pooler, blur_type, nl, num_classes, equalized_lr, mbstd_group_size, 'ProGAN')
