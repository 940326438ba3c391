"""Drop-in layer API of gan-lab (reference gan_lab/utils/custom_layers.py) on sm_100a kernels.

Same class names, constructor signatures, attributes (`ni`, `nf`, `wscale`, `conv2d`, `linear`, `nout_feat`)
and state_dict keys (`conv2d.weight (Cout,Cin,kh,kw)`, `conv2d.bias`, `linear.weight`, `linear.bias`,
`bias (1,C,1,1)`) as the reference, so its checkpoints load.  The `nn.Conv2d` / `nn.Linear` children only
hold the parameters: forward() goes to the kernels in `gan_lab_b200.ops` with the equalized-LR scale folded
into the GEMM epilogue (reference multiplies the *input* tensor, custom_layers.py:204, 284) and, through
the extra `act=` argument used by the fused architectures, bias + leaky-ReLU in the same epilogue.

Feature maps are logical NCHW tensors kept in `torch.channels_last` memory (NHWC); any other layout is
converted on entry.  Conv weights are stored channels-last too (KRSC) so kernels read them in place.
"""
import torch
from torch import nn

from .. import ops
from .initializer import Initializer


class Lambda(nn.Module):
    """Converts any function into a Module (reference custom_layers.py:18-29)."""

    def __init__(self, func, **kwargs):
        super(Lambda, self).__init__()
        self.func = func
        self.kwargs = kwargs if kwargs else {}

    def forward(self, x):
        return self.func(x, **self.kwargs)


class Blur3x3(nn.Module):
    """Module form of the depthwise binomial FIR returned by `get_blur_op('binomial', C)`."""

    def __init__(self, num_channels=None):
        super(Blur3x3, self).__init__()
        self.num_channels = num_channels

    def forward(self, x):
        return ops.blur3x3(x)


def get_blur_op(blur_type, num_channels):
    """reference custom_layers.py:36-53.  Only the stride-1 3x3 binomial filter is on any training path."""
    if blur_type.casefold() == 'binomial':
        return Blur3x3(num_channels)
    elif blur_type.casefold() == 'box':
        raise NotImplementedError('box blur (stride 3) is not exercised by any gan-lab configuration; not built')
    elif blur_type.casefold() == 'gaussian':
        raise NotImplementedError('Gaussian blur not yet implemented.')
    raise ValueError(blur_type)


class Upsample2x(nn.Module):
    """nn.Upsample(scale_factor=2, mode='nearest') (reference resnetgan/learner.py:154-158)."""

    def forward(self, x):
        return ops.upsample2x(x)


class AvgPool2x(nn.Module):
    """nn.AvgPool2d(kernel_size=2, stride=2) (reference resnetgan/learner.py:160-164)."""

    def forward(self, x):
        return ops.avgpool2(x)


class LeakyReLU(nn.LeakyReLU):
    """The shared nonlinearity instance (reference resnetgan/learner.py:176-177) as a standalone kernel call;
    inside the fused layers it is folded into the producing kernel's epilogue instead."""

    def forward(self, x):
        return ops.bias_act(x, None, 1.0, ops.ACT_LRELU, self.negative_slope)


class ReLU(LeakyReLU):
    def __init__(self):
        super(ReLU, self).__init__(negative_slope=0.0)


def as_native_nl(nl):
    """Map a torch activation module onto the native one (slope carried over)."""
    if isinstance(nl, LeakyReLU):
        return nl
    if isinstance(nl, nn.LeakyReLU):
        return LeakyReLU(nl.negative_slope)
    if isinstance(nl, nn.ReLU):
        return ReLU()
    raise NotImplementedError(f'activation {type(nl).__name__} has no sm_100a kernel here')


def as_native_upsampler(up):
    if isinstance(up, Upsample2x):
        return up
    if isinstance(up, nn.Upsample) and up.mode == 'nearest' and float(up.scale_factor) == 2.0:
        return Upsample2x()
    raise NotImplementedError('only nearest x2 upsampling is built (the default of every gan-lab configuration)')


def as_native_pooler(p):
    if isinstance(p, AvgPool2x):
        return p
    if isinstance(p, nn.AvgPool2d):
        return AvgPool2x()
    if isinstance(p, (NearestPool2d, BilinearPool2d)):
        return p
    raise NotImplementedError(f'pooler {type(p).__name__} not supported')


class NearestPool2d(nn.Module):
    """reference custom_layers.py:59-65 (F.interpolate(scale_factor=.5, mode='nearest')): pure data movement."""

    def forward(self, x):
        if x.dim() == 3:
            x = x.view(-1, *x.shape)
        return x[..., ::2, ::2].contiguous(memory_format=torch.channels_last)


class BilinearPool2d(nn.Module):
    """reference custom_layers.py:67-75; at scale 0.5 without align_corners bilinear == 2x2 average."""

    def __init__(self, align_corners):
        super(BilinearPool2d, self).__init__()
        self.align_corners = align_corners

    def forward(self, x):
        if x.dim() == 3:
            x = x.view(-1, *x.shape)
        if self.align_corners:
            raise NotImplementedError('bilinear pooling with align_corners=True is not built')
        return ops.avgpool2(x)


class PixelNorm2d(nn.Module):
    """reference custom_layers.py:81-86 (5 ATen kernels) -> one warp-per-pixel kernel."""

    def forward(self, x, eps=1.e-8):
        return ops.pixelnorm(x, eps)


class InstanceNorm2d(nn.Module):
    """nn.InstanceNorm2d(ni, eps=1e-8): no affine, no running stats (reference custom_layers.py:98-99)."""

    def __init__(self, ni=None, eps=1.e-8):
        super(InstanceNorm2d, self).__init__()
        self.eps = eps

    def forward(self, x):
        return ops.instance_norm(x, self.eps)


class BatchNorm2d(nn.BatchNorm2d):
    """nn.BatchNorm2d(ni) of the ResNet generator blocks (reference custom_layers.py:100-102): same parameters / buffers /
    state_dict keys (weight, bias, running_mean, running_var, num_batches_tracked); training-mode forward (batch
    statistics, running-buffer momentum update) is one fused kernel family, optionally with the following ReLU."""

    def forward(self, x, act=None, slope=0.0):
        if self.momentum is None or not self.affine or not self.track_running_stats:
            raise NotImplementedError('only the nn.BatchNorm2d defaults used by gan-lab are built')
        if not self.training:
            return self._forward_eval(x, act, slope)
        return ops.batchnorm_act(x, self.weight, self.bias, self.running_mean, self.running_var, self.num_batches_tracked,
                                 self.eps, self.momentum, ops.ACT_NONE if act is None else act, slope)


    def _forward_eval(self, x, act, slope):
        """Evaluation mode (validation metrics, image grids -- off the training path): running statistics, i.e. a per-channel
        affine map y = x*s + t, s = weight / sqrt(running_var + eps), t = bias - running_mean*s.  No kernel of its own: it is
        issued as a 1x1 convolution with the diagonal weight diag(s), bias t and the fused activation on the exact fp32
        convolution kernel (C <= 512 here, so the C/1 excess of multiplies is immaterial for an inference-only call)."""
        from .. import _kernels as K
        with torch.no_grad():
            s = self.weight.detach() / torch.sqrt(self.running_var + self.eps)
            t = self.bias.detach() - self.running_mean * s
            w = torch.diag(s).view(s.numel(), s.numel(), 1, 1).contiguous(memory_format=torch.channels_last)
        impl = K.get_conv_impl()
        K.set_conv_impl('fp32')                     # not the TF32 tensor-core path: its operand truncation would show in eval outputs
        try:
            return ops.conv2d(x, w, t, 0, 1.0, 1.0, ops.ACT_NONE if act is None else act, slope)
        finally:
            K.set_conv_impl(impl)


class LayerNorm(nn.LayerNorm):
    """nn.LayerNorm([ni, res, res]) of the ResNet discriminator blocks (reference custom_layers.py:103-106).  weight / bias
    keep the reference's (C,H,W) shape and keys but are STORED in (H,W,C) order -- the order of the NHWC activations they
    multiply -- so the kernels read them in place (load_state_dict / Adam / deepcopy all preserve that layout)."""

    def __init__(self, normalized_shape, eps=1e-5):
        super(LayerNorm, self).__init__(list(normalized_shape), eps=eps, elementwise_affine=True)
        assert len(self.normalized_shape) == 3
        for p in (self.weight, self.bias):
            p.data = p.data.permute(1, 2, 0).contiguous().permute(2, 0, 1)

    def forward(self, x, act=None, slope=0.0):
        if tuple(x.shape[1:]) != tuple(self.normalized_shape):
            raise ValueError(f'LayerNorm{tuple(self.normalized_shape)} got input of shape {tuple(x.shape)}')
        return ops.layernorm_act(x, self.weight, self.bias, self.eps, ops.ACT_NONE if act is None else act, slope)


class Tanh(nn.Tanh):
    def forward(self, x):
        return ops.tanh(x)


class NormalizeLayer(nn.Module):
    """All normalization methods in one place (reference custom_layers.py:88-111).  `forward(x, act=...)` lets the blocks fold
    the nonlinearity that follows a Batch/LayerNorm into the same kernel."""

    def __init__(self, norm_type, ni=None, res=None):
        super(NormalizeLayer, self).__init__()
        norm_type = norm_type.lower()
        self.norm_type = norm_type
        if norm_type in ('pixelnorm', 'pixel norm'):
            self.norm = PixelNorm2d()
        elif norm_type in ('instancenorm', 'instance norm'):
            self.norm = InstanceNorm2d(ni, eps=1.e-8)
        elif norm_type in ('batchnorm', 'batch norm'):
            assert isinstance(ni, int)
            self.norm = BatchNorm2d(ni)
        elif norm_type in ('layernorm', 'layer norm'):
            assert isinstance(ni, int); assert isinstance(res, int)
            self.norm = LayerNorm([ni, res, res])
        else:
            raise Exception(f'`norm_type` == "{norm_type}" not supported.')
        self.fuses_act = isinstance(self.norm, (BatchNorm2d, LayerNorm))

    def forward(self, x, act=None, slope=0.0):
        if self.fuses_act:
            return self.norm(x, act=act, slope=slope)
        assert act is None
        return self.norm(x)


def concat_mbstd_layer(x, group_size=4):
    """Minibatch standard deviation layer (reference custom_layers.py:117-140)."""
    return ops.mbstd_concat(x, group_size)


class Conv2dEx(nn.Module):
    """Equalized-LR convolution (reference custom_layers.py:147-211)."""

    def __init__(self, ni, nf, ks, stride=1, padding=0, groups=1, init='he', init_type='default',
                 gain_sq_base=2., equalized_lr=False, lrmul=1., include_bias=True):
        super(Conv2dEx, self).__init__()
        if stride != 1 or groups != 1:
            raise NotImplementedError('only stride-1 dense convolutions exist on the gan-lab training path')
        self.ni = ni; self.nf = nf
        self.ks = ks; self.padding = padding
        init_cf = init.casefold() if init is not None else None
        init_type = init_type.casefold()
        self.equalized_lr = equalized_lr
        self.use_lrmul = True if lrmul != 1. else False
        self.lrmul = lrmul

        self.initializer = None
        if init_type != 'standard normal':
            self.initializer = Initializer(init=init_cf, init_type=init_type, gain_sq_base=gain_sq_base,
                                           equalized_lr=equalized_lr)
        self.conv2d = nn.Conv2d(ni, nf, kernel_size=ks, stride=stride, padding=padding, groups=groups, bias=include_bias)

        self.wscale = None
        # NB the reference tests `init_type == ('default', 'resnet')` (custom_layers.py:173), which is never true:
        # 'default'/'resnet' convs keep nn.Conv2d's own initialisation and wscale stays None.
        if init_type in ('progan', 'stylegan') and init_cf is not None:
            bound = self.initializer.get_init_bound_layer(tensor=self.conv2d.weight, distribution_type='Normal', stride=stride)
            if equalized_lr:
                self.wscale = bound
                self.conv2d.weight.data.normal_(0., 1. / lrmul)
            else:
                self.conv2d.weight.data.normal_(0., bound / lrmul)
        elif init_type == 'standard normal' and init is None and not equalized_lr and not self.use_lrmul:
            self.conv2d.weight.data.normal_(0., 1.)
        if equalized_lr and self.wscale is None:
            raise TypeError('equalized_lr needs a ProGAN/StyleGAN init_type (the reference fails the same way: wscale is None)')

        self.bias = None
        if include_bias:
            self.conv2d.bias.data.fill_(0)
        # KRSC in memory, OIHW in shape: kernels read the parameter in place
        self.conv2d.weight.data = self.conv2d.weight.data.contiguous(memory_format=torch.channels_last)

    @property
    def alpha(self):
        a = self.wscale if self.equalized_lr else 1.
        return a * self.lrmul if self.use_lrmul else a

    def forward(self, x, act=None, slope=0.2, blur=False, up=False):
        """`act=None` reproduces the reference module; the fused architectures pass act=ops.ACT_LRELU to fold
        the following LeakyReLU into the conv epilogue, and blur=True to let the binomial FIR that follows ride in the op."""
        act = ops.ACT_NONE if act is None else act
        w, b = self.conv2d.weight, self.conv2d.bias
        bscale = self.lrmul if self.use_lrmul else 1.
        if up:                                         # the nearest-neighbour 2x upsampler in front of this layer rides in the conv
            if self.ks == 3 and self.padding == 1 and not blur:
                return ops.upconv2d(x, w, b, self.alpha, bscale, act, slope)
            x = ops.upsample2x(x)
        if self.ks == 1 and self.padding == 0 and x.dim() == 4:
            if self.ni == 3 and self.nf % 4 == 0:
                return ops.fromrgb(x, w, b, self.alpha, bscale, act, slope)
            if self.nf == 3 and self.ni % 4 == 0 and act == ops.ACT_NONE:
                return ops.torgb(x, w, b, self.alpha, bscale)
        return ops.conv2d(x, w, b, self.padding, self.alpha, bscale, act, slope, blur=blur)


class Conv2dBias(nn.Module):
    """reference custom_layers.py:213-226."""

    def __init__(self, nf, lrmul=1., device='cuda' if torch.cuda.is_available() else 'cpu'):
        super(Conv2dBias, self).__init__()
        self.use_lrmul = True if lrmul != 1. else False
        self.lrmul = lrmul
        self.bias = nn.Parameter(torch.zeros(1, nf, 1, 1, dtype=torch.float32, device=device))

    @property
    def bias_scale(self):
        return self.lrmul if self.use_lrmul else 1.

    def forward(self, x, act=None, slope=0.2):
        return ops.bias_act(x, self.bias, self.bias_scale, ops.ACT_NONE if act is None else act, slope)


class LinearEx(nn.Module):
    """Equalized-LR fully connected layer (reference custom_layers.py:230-291)."""

    def __init__(self, nin_feat, nout_feat, init='xavier', init_type='default', gain_sq_base=2.,
                 equalized_lr=False, lrmul=1., include_bias=True):
        super(LinearEx, self).__init__()
        self.nin_feat = int(nin_feat); self.nout_feat = int(nout_feat)
        init_cf = init.casefold() if init is not None else None
        init_type = init_type.casefold()
        self.equalized_lr = equalized_lr
        self.use_lrmul = True if lrmul != 1. else False
        self.lrmul = lrmul

        self.initializer = None
        if init_type != 'standard normal':
            self.initializer = Initializer(init=init_cf, init_type=init_type, gain_sq_base=gain_sq_base,
                                           equalized_lr=equalized_lr)
        self.linear = nn.Linear(self.nin_feat, self.nout_feat, bias=include_bias)

        self.wscale = None
        if init_type in ('default', 'resnet') and init_cf is not None:
            bound = self.initializer.get_init_bound_layer(tensor=self.linear.weight, distribution_type='Uniform')
            if equalized_lr:
                self.wscale = bound
                self.linear.weight.data.uniform_(-1. / lrmul, 1. / lrmul)
            else:
                self.linear.weight.data.uniform_(-bound / lrmul, bound / lrmul)
        elif init_type in ('progan', 'stylegan') and init_cf is not None:
            bound = self.initializer.get_init_bound_layer(tensor=self.linear.weight, distribution_type='Normal')
            if equalized_lr:
                self.wscale = bound
                self.linear.weight.data.normal_(0., 1. / lrmul)
            else:
                self.linear.weight.data.normal_(0., bound / lrmul)
        elif init_type == 'standard normal' and init is None and not equalized_lr and not self.use_lrmul:
            self.linear.weight.data.normal_(0., 1.)

        self.bias = None
        if include_bias:
            self.linear.bias.data.fill_(0)

    @property
    def alpha(self):
        a = self.wscale if self.equalized_lr else 1.
        return a * self.lrmul if self.use_lrmul else a

    def forward(self, x, act=None, slope=0.2):
        bscale = self.lrmul if self.use_lrmul else 1.
        return ops.linear(x, self.linear.weight, self.linear.bias, self.alpha, bscale,
                          ops.ACT_NONE if act is None else act, slope)


class LinearBias(nn.Module):
    """reference custom_layers.py:293-306."""

    def __init__(self, nout_feat, lrmul=1., device='cuda' if torch.cuda.is_available() else 'cpu'):
        super(LinearBias, self).__init__()
        self.use_lrmul = True if lrmul != 1. else False
        self.lrmul = lrmul
        self.bias = nn.Parameter(torch.zeros(1, nout_feat, dtype=torch.float32, device=device))

    def forward(self, x):
        return ops.bias_act(x, self.bias, self.lrmul if self.use_lrmul else 1., ops.ACT_NONE, 0.2)
