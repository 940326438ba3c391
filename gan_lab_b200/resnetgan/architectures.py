"""ResNet GAN generators and discriminators (reference gan_lab/resnetgan/architectures.py:29-213).

Same class names, constructor signatures, module tree and `state_dict` keys as the reference; forward passes run on the
sm_100a kernels (feature maps NHWC in memory).  The (N, C*H*W) <-> (N,C,H,W) reshapes at the FC boundaries keep the
reference's C-major flattening order, so FC weights mean the same thing.
"""
import torch
from torch import nn

from .. import ops
from .base import GAN
from .resblocks import ResBlock2d, ResBlock2d32Pix, FastResBlock2dDownsample, run_fused
from ..utils.custom_layers import (Lambda, NormalizeLayer, Conv2dEx, LinearEx, Tanh, Upsample2x, AvgPool2x, ReLU,
                                   as_native_nl, as_native_upsampler, as_native_pooler)

FMAP_SAMPLES = 3
RES_INIT = 4

FMAP_G = 64
FMAP_D = 64
FMAP_G_INIT_32_FCTR = 1
FMAP_G_INIT_64_FCTR = 4

RES_FEATURE_SPACE = 4


def _run(seq, x):
    """run_fused + residual blocks / Lambdas called as they are."""
    return run_fused(seq, x)


def _global_avgpool(x, k):
    """nn.AvgPool2d(k, k) for k a power of two = log2(k) exact 2x2 average pools (equal-size groups)."""
    while k > 1:
        x = ops.avgpool2(x)
        k //= 2
    return x


class _GlobalAvgPool(nn.Module):
    def __init__(self, k):
        super(_GlobalAvgPool, self).__init__()
        assert k & (k - 1) == 0
        self.k = k

    def forward(self, x):
        return _global_avgpool(x, self.k)


# ---------------------------------------------------------------------------------------------- generators
class Generator32PixResnet(GAN):
    """ResNet GAN Generator for 32-pixel samples (reference architectures.py:29-61)."""

    def __init__(self, len_latent=128, fmap=FMAP_G * 2, upsampler=None, blur_type=None, nl=None, num_classes=0,
                 equalized_lr=False):
        super(Generator32PixResnet, self).__init__(32)
        upsampler = as_native_upsampler(upsampler) if upsampler is not None else Upsample2x()
        nl = as_native_nl(nl) if nl is not None else ReLU()
        self.len_latent = len_latent
        self.num_classes = num_classes
        self.equalized_lr = equalized_lr
        _fmap_init_32 = len_latent * FMAP_G_INIT_32_FCTR
        self.generator_model = nn.Sequential(
            Lambda(lambda x: x.view(-1, len_latent + num_classes)),
            LinearEx(nin_feat=len_latent + num_classes, nout_feat=_fmap_init_32 * RES_INIT ** 2, init='Xavier',
                     equalized_lr=equalized_lr),
            Lambda(lambda x: x.view(-1, _fmap_init_32, RES_INIT, RES_INIT)),
            ResBlock2d32Pix(ni=_fmap_init_32, nf=fmap, ks=3, norm_type='BatchNorm', upsampler=upsampler, init='He', nl=nl,
                            equalized_lr=equalized_lr, blur_type=blur_type),
            ResBlock2d32Pix(ni=fmap, nf=fmap, ks=3, norm_type='BatchNorm', upsampler=upsampler, init='He', nl=nl,
                            equalized_lr=equalized_lr, blur_type=blur_type),
            ResBlock2d32Pix(ni=fmap, nf=fmap, ks=3, norm_type='BatchNorm', upsampler=upsampler, init='He', nl=nl,
                            equalized_lr=equalized_lr, blur_type=blur_type),
            NormalizeLayer('BatchNorm', ni=fmap),
            nl,
            Conv2dEx(ni=fmap, nf=FMAP_SAMPLES, ks=3, stride=1, padding=1, init='Xavier', equalized_lr=equalized_lr),
            Tanh()
        )

    def forward(self, x):
        return _run(self.generator_model, x)


class Generator64PixResnet(GAN):
    """ResNet GAN Generator for 64-pixel samples (reference architectures.py:63-99)."""

    def __init__(self, len_latent=128, fmap=FMAP_G, upsampler=None, blur_type=None, nl=None, num_classes=0,
                 equalized_lr=False):
        super(Generator64PixResnet, self).__init__(64)
        upsampler = as_native_upsampler(upsampler) if upsampler is not None else Upsample2x()
        nl = as_native_nl(nl) if nl is not None else ReLU()
        self.len_latent = len_latent
        self.num_classes = num_classes
        self.equalized_lr = equalized_lr
        _fmap_init_64 = len_latent * FMAP_G_INIT_64_FCTR
        self.generator_model = nn.Sequential(
            Lambda(lambda x: x.view(-1, len_latent + num_classes)),
            LinearEx(nin_feat=len_latent + num_classes, nout_feat=_fmap_init_64 * RES_INIT ** 2, init='Xavier',
                     equalized_lr=equalized_lr),
            Lambda(lambda x: x.view(-1, _fmap_init_64, RES_INIT, RES_INIT)),
            ResBlock2d(ni=_fmap_init_64, nf=8 * fmap, ks=3, norm_type='BatchNorm', upsampler=upsampler, init='He', nl=nl,
                       equalized_lr=equalized_lr, blur_type=blur_type),
            ResBlock2d(ni=8 * fmap, nf=4 * fmap, ks=3, norm_type='BatchNorm', upsampler=upsampler, init='He', nl=nl,
                       equalized_lr=equalized_lr, blur_type=blur_type),
            ResBlock2d(ni=4 * fmap, nf=2 * fmap, ks=3, norm_type='BatchNorm', upsampler=upsampler, init='He', nl=nl,
                       equalized_lr=equalized_lr, blur_type=blur_type),
            ResBlock2d(ni=2 * fmap, nf=1 * fmap, ks=3, norm_type='BatchNorm', upsampler=upsampler, init='He', nl=nl,
                       equalized_lr=equalized_lr, blur_type=blur_type),
            NormalizeLayer('BatchNorm', ni=1 * fmap),
            nl,
            Conv2dEx(ni=1 * fmap, nf=FMAP_SAMPLES, ks=3, stride=1, padding=1, init='He', equalized_lr=equalized_lr),
            Tanh()
        )

    def forward(self, x):
        return _run(self.generator_model, x)


# ---------------------------------------------------------------------------------------------- discriminators
class Discriminator32PixResnet(GAN):
    """ResNet GAN Discriminator for 32-pixel samples (reference architectures.py:105-133)."""

    def __init__(self, fmap=FMAP_D * 2, pooler=None, blur_type=None, nl=None, num_classes=0, equalized_lr=False):
        super(Discriminator32PixResnet, self).__init__(32)
        pooler = as_native_pooler(pooler) if pooler is not None else AvgPool2x()
        nl = as_native_nl(nl) if nl is not None else ReLU()
        self.num_classes = num_classes
        self.equalized_lr = equalized_lr
        self.view1 = Lambda(lambda x: x.view(-1, FMAP_SAMPLES + num_classes, self.res, self.res))
        self.conv1 = FastResBlock2dDownsample(ni=FMAP_SAMPLES + num_classes, nf=fmap, ks=3, pooler=pooler, init='Xavier',
                                              nl=nl, equalized_lr=equalized_lr, blur_type=blur_type)
        self.resblocks = nn.Sequential(
            ResBlock2d32Pix(ni=fmap, nf=fmap, ks=3, norm_type='LayerNorm', pooler=pooler, init='He', nl=nl,
                            res=self.res // 2, equalized_lr=equalized_lr, blur_type=blur_type),
            ResBlock2d32Pix(ni=fmap, nf=fmap, ks=3, norm_type='LayerNorm', init='He', nl=nl, res=self.res // 4,
                            equalized_lr=equalized_lr, blur_type=blur_type),
            ResBlock2d32Pix(ni=fmap, nf=fmap, ks=3, norm_type='LayerNorm', init='He', nl=nl, res=self.res // 4,
                            equalized_lr=equalized_lr, blur_type=blur_type),
            nl,
            _GlobalAvgPool(self.res // 4),                      # nn.AvgPool2d(res//4, res//4): no parameters
            Lambda(lambda x: x.reshape(-1, fmap))               # final feature space
        )
        self.linear1 = LinearEx(nin_feat=fmap, nout_feat=1, init='Xavier', equalized_lr=equalized_lr)

    def forward(self, x):
        return (self.linear1(_run(self.resblocks, self.conv1(self.view1(x))))).view(-1)


class Discriminator64PixResnet(GAN):
    """ResNet GAN Discriminator for 64-pixel samples (reference architectures.py:150-182)."""

    def __init__(self, fmap=FMAP_D, pooler=None, blur_type=None, nl=None, num_classes=0, equalized_lr=False):
        super(Discriminator64PixResnet, self).__init__(64)
        pooler = as_native_pooler(pooler) if pooler is not None else AvgPool2x()
        nl = as_native_nl(nl) if nl is not None else ReLU()
        self.num_classes = num_classes
        self.equalized_lr = equalized_lr
        self.view1 = Lambda(lambda x: x.view(-1, FMAP_SAMPLES + num_classes, self.res, self.res))
        self.conv1 = Conv2dEx(ni=FMAP_SAMPLES + num_classes, nf=1 * fmap, ks=3, stride=1, padding=1, init='Xavier',
                              equalized_lr=equalized_lr)
        self.resblocks = nn.Sequential(
            ResBlock2d(ni=1 * fmap, nf=2 * fmap, ks=3, norm_type='LayerNorm', pooler=pooler, init='He', nl=nl,
                       res=self.res // 1, equalized_lr=equalized_lr, blur_type=blur_type),
            ResBlock2d(ni=2 * fmap, nf=4 * fmap, ks=3, norm_type='LayerNorm', pooler=pooler, init='He', nl=nl,
                       res=self.res // 2, equalized_lr=equalized_lr, blur_type=blur_type),
            ResBlock2d(ni=4 * fmap, nf=8 * fmap, ks=3, norm_type='LayerNorm', pooler=pooler, init='He', nl=nl,
                       res=self.res // 4, equalized_lr=equalized_lr, blur_type=blur_type),
            ResBlock2d(ni=8 * fmap, nf=8 * fmap, ks=3, norm_type='LayerNorm', pooler=pooler, init='He', nl=nl,
                       res=self.res // 8, equalized_lr=equalized_lr, blur_type=blur_type),
            # final feature space, flattened C-major as the reference's .view on an NCHW tensor does
            Lambda(lambda x: x.reshape(-1, RES_FEATURE_SPACE ** 2 * 8 * fmap))
        )
        self.linear1 = LinearEx(nin_feat=RES_FEATURE_SPACE ** 2 * 8 * fmap, nout_feat=1, init='Xavier',
                                equalized_lr=equalized_lr)

    def forward(self, x):
        return (self.linear1(_run(self.resblocks, self.conv1(self.view1(x))))).view(-1)
