"""StyleGAN generator (reference gan_lab/stylegan/architectures.py) and discriminator on sm_100a kernels.

Module tree and state_dict keys are the reference's (`const_input`, `gen_layers.{n}.{0|0.0|0.1}.conv2d.weight`,
`gen_layers.{n}.1.noise_weight`, `gen_layers.{n}.2.0.bias`, `gen_layers.{n}.3.linear.*`,
`z_to_w.fc_mapping_model.fc_{i}.linear.*`, `torgb.conv2d.*`, `prev_torgb.conv2d.*`).  forward() keeps the
reference's control flow (mixing regularisation, w_ewma side effect, fade-in) but each layer's
noise + bias + leaky-ReLU + InstanceNorm + AdaIN chain (~8 ATen kernels) is one fused op.
"""
import copy

import numpy as np
import torch
from torch import nn

from .. import ops
from .. import _kernels as K
from .._growth import GrowthState
from ..utils.latent_utils import gen_rand_latent_vars, RANDOM, RandomSource
from ..utils.custom_layers import (Lambda, get_blur_op, NormalizeLayer, Conv2dEx, LinearEx, Conv2dBias, Blur3x3,
                                   Upsample2x, LeakyReLU, InstanceNorm2d, PixelNorm2d, as_native_nl,
                                   as_native_upsampler)
from ..progan.architectures import _DiscriminatorImpl, FMAP_SAMPLES, RES_INIT
from .base import StyleGAN

FMAP_G_INIT_FCTR = 1


class StyleMappingNetwork(nn.Module):
    """Mapping network z -> w (reference stylegan/architectures.py:27-59): PixelNorm + 8 x (LinearEx, lrelu),
    each FC a single kernel with lrmul, bias and lrelu in the epilogue."""

    def __init__(self, len_latent=512, len_dlatent=512, num_fcs=8, lrmul=.01, nl=None, equalized_lr=True,
                 normalize_z=True):
        super(StyleMappingNetwork, self).__init__()
        nl = as_native_nl(nl) if nl is not None else LeakyReLU(.2)
        self.len_latent = len_latent
        if normalize_z:
            self.preprocess_z = nn.Sequential(Lambda(lambda x: x.view(-1, len_latent)), NormalizeLayer('PixelNorm'))
        else:
            self.preprocess_z = Lambda(lambda x: x.view(-1, len_latent))
        self.dims = np.linspace(len_latent, len_dlatent, num_fcs + 1).astype(np.int64)
        self.fc_mapping_model = nn.Sequential()
        for seq_n in range(num_fcs):
            self.fc_mapping_model.add_module(
                'fc_' + str(seq_n),
                LinearEx(nin_feat=self.dims[seq_n], nout_feat=self.dims[seq_n + 1], init='He', init_type='StyleGAN',
                         gain_sq_base=2., equalized_lr=equalized_lr, lrmul=lrmul))
            self.fc_mapping_model.add_module('nl_' + str(seq_n), nl)

    def forward(self, x):
        x = self.preprocess_z(x)
        mods = list(self.fc_mapping_model)
        for fc, nl in zip(mods[0::2], mods[1::2]):
            x = fc(x, act=ops.ACT_LRELU, slope=nl.negative_slope)
        return x


class StyleAddNoise(nn.Module):
    """Adds weighted uncorrelated Gaussian noise (reference stylegan/architectures.py:105-119).  Standalone it is a
    bias-free call of the fused epilogue's first stage; inside StyleGenerator it is fused with what follows."""

    def __init__(self, nf):
        super(StyleAddNoise, self).__init__()
        self.noise_weight = nn.Parameter(torch.zeros(1, nf, 1, 1, dtype=torch.float32))

    def draw(self, x):
        return RANDOM.source.randn((x.shape[0], 1, x.shape[2], x.shape[3]), x.device)

    def forward(self, x, noise=None):
        if self.training or noise is None:
            noise = self.draw(x)
        return x + self.noise_weight * noise      # unfused reference semantics (not used on the training path)


class StyleGenerator(StyleGAN):
    """StyleGAN (Karras et al. 2019) generator (reference stylegan/architectures.py:126-528)."""

    def __init__(self, final_res, latent_distribution='normal', len_latent=512, len_dlatent=512, mapping_num_fcs=8,
                 mapping_lrmul=.01, use_instancenorm=True, use_noise=True, upsampler=None, blur_type=None, nl=None,
                 num_classes=0, equalized_lr=True, normalize_z=True, use_pixelnorm=False, pct_mixing_reg=.9,
                 truncation_trick_params={'beta': .995, 'psi': .7, 'cutoff_stage': 4}, state: GrowthState = None):
        super(StyleGenerator, self).__init__(final_res, state=state)
        self.gen_layers = nn.ModuleList()
        self.upsampler = as_native_upsampler(upsampler) if upsampler is not None else Upsample2x()
        self.gen_blur_type = blur_type
        self.nl = as_native_nl(nl) if nl is not None else LeakyReLU(.2)
        self.equalized_lr = equalized_lr
        self.pct_mixing_reg = pct_mixing_reg
        self._use_mixing_reg = True if pct_mixing_reg else False
        self.latent_distribution = latent_distribution
        self.len_latent = len_latent
        self.len_dlatent = len_dlatent
        assert isinstance(num_classes, int)
        self.num_classes = num_classes
        if num_classes:
            raise NotImplementedError('class-conditioned mapping network is off the benchmarked path; not built')
        self.z_to_w = StyleMappingNetwork(len_latent=len_latent, len_dlatent=len_dlatent, num_fcs=mapping_num_fcs,
                                          lrmul=mapping_lrmul, nl=self.nl, equalized_lr=equalized_lr,
                                          normalize_z=normalize_z)
        _fmap_init = len_latent * FMAP_G_INIT_FCTR
        self.const_input = nn.Parameter(torch.ones(1, _fmap_init, RES_INIT, RES_INIT, dtype=torch.float32))

        self._use_noise = use_noise
        self._trained_with_noise = use_noise
        if use_noise:
            conv = Conv2dEx(ni=_fmap_init, nf=self.fmap, ks=3, stride=1, padding=1, init='He', init_type='StyleGAN',
                            gain_sq_base=2., equalized_lr=equalized_lr, include_bias=False)
            noise = [StyleAddNoise(nf=_fmap_init), StyleAddNoise(nf=self.fmap)]
            bias = ([Conv2dBias(nf=_fmap_init, device='cpu')], [Conv2dBias(nf=self.fmap, device='cpu')])
        else:
            conv = Conv2dEx(ni=_fmap_init, nf=self.fmap, ks=3, stride=1, padding=1, init='He', init_type='StyleGAN',
                            gain_sq_base=2., equalized_lr=equalized_lr, include_bias=True)
            noise = [None, None]
            bias = ([], [])

        self.use_pixelnorm = use_pixelnorm
        self.use_instancenorm = use_instancenorm

        def norms():
            n = []
            if use_pixelnorm:
                n.append(NormalizeLayer('PixelNorm'))
            if use_instancenorm:
                n.append(NormalizeLayer('InstanceNorm'))
            return n

        w_to_styles = (
            LinearEx(nin_feat=self.z_to_w.dims[-1], nout_feat=2 * _fmap_init, init='He', init_type='StyleGAN',
                     gain_sq_base=1., equalized_lr=equalized_lr),
            LinearEx(nin_feat=self.z_to_w.dims[-1], nout_feat=2 * self.fmap, init='He', init_type='StyleGAN',
                     gain_sq_base=1., equalized_lr=equalized_lr))
        assert 0. <= truncation_trick_params['beta'] <= 1.
        self.w_ewma_beta = truncation_trick_params['beta']
        self._w_eval_psi = truncation_trick_params['psi']
        assert ((isinstance(truncation_trick_params['cutoff_stage'], int) and
                 0 < truncation_trick_params['cutoff_stage'] <= int(np.log2(self.final_res)) - 2) or
                truncation_trick_params['cutoff_stage'] is None)
        self._trunc_cutoff_stage = truncation_trick_params['cutoff_stage']
        self.use_truncation_trick = True if self._trunc_cutoff_stage else False
        self.w_ewma = None
        self.grouped_styles = True              # device-mixing path: all style affines as one grouped launch
        self._style_table = K.GroupedLinearTable()

        self.gen_layers.append(nn.ModuleList([None, noise[0], nn.Sequential(*bias[0], self.nl, *norms()), w_to_styles[0]]))
        self.gen_layers.append(nn.ModuleList([conv, noise[1], nn.Sequential(*bias[1], self.nl, *norms()), w_to_styles[1]]))
        self.prev_torgb = None
        self._update_torgb(ni=self.fmap)

    # ------------------------------------------------------------------ growth
    def increase_scale(self):
        """reference stylegan/architectures.py:271-290."""
        if not self.scale_inc_metadata_updated:
            super(StyleGenerator, self).increase_scale()
        else:
            self.scale_inc_metadata_updated = False
        blur_op = get_blur_op(blur_type=self.gen_blur_type, num_channels=self.fmap) if self.gen_blur_type is not None else None
        self.gen_layers.append(self.get_conv_layer(ni=self.fmap_prev, upsample=True, blur_op=blur_op))
        self.gen_layers.append(self.get_conv_layer(ni=self.fmap))
        self.prev_torgb = copy.deepcopy(self.torgb)
        self._update_torgb(ni=self.fmap)
        self.to(self.const_input.device)

    def get_conv_layer(self, ni, upsample=False, blur_op=None, append_nl=True):
        """reference stylegan/architectures.py:292-334."""
        upsampler = [self.upsampler] if upsample else []
        if self.use_noise or blur_op is not None:
            conv = Conv2dEx(ni=ni, nf=self.fmap, ks=3, stride=1, padding=1, init='He', init_type='StyleGAN',
                            gain_sq_base=2., equalized_lr=self.equalized_lr, include_bias=False)
            bias = [Conv2dBias(nf=self.fmap, device='cpu')]
        else:
            conv = Conv2dEx(ni=ni, nf=self.fmap, ks=3, stride=1, padding=1, init='He', init_type='StyleGAN',
                            gain_sq_base=2., equalized_lr=self.equalized_lr, include_bias=True)
            bias = []
        blur = [blur_op] if blur_op is not None else []
        noise = StyleAddNoise(nf=self.fmap) if self.use_noise else None
        nl = [self.nl] if append_nl else []
        norms = []
        if self.use_pixelnorm:
            norms.append(NormalizeLayer('PixelNorm'))
        if self.use_instancenorm:
            norms.append(NormalizeLayer('InstanceNorm'))
        w_to_style = LinearEx(nin_feat=self.z_to_w.dims[-1], nout_feat=2 * self.fmap, init='He', init_type='StyleGAN',
                              gain_sq_base=1., equalized_lr=self.equalized_lr)
        return nn.ModuleList([nn.Sequential(*upsampler, conv, *blur), noise, nn.Sequential(*(bias + nl + norms)), w_to_style])

    def _update_torgb(self, ni):
        self.torgb = Conv2dEx(ni=ni, nf=FMAP_SAMPLES, ks=1, stride=1, padding=0, init='He', init_type='StyleGAN',
                              gain_sq_base=1., equalized_lr=self.equalized_lr)

    # ------------------------------------------------------------------ mode switches (reference :343-408)
    def train(self, mode=True):
        super(StyleGenerator, self).train(mode=mode)
        self._use_noise = self._trained_with_noise
        self._use_mixing_reg = True if self.pct_mixing_reg else False
        return self

    def eval(self):
        super(StyleGenerator, self).eval()
        self._use_mixing_reg = False
        return self

    def to(self, *args, **kwargs):
        super(StyleGenerator, self).to(*args, **kwargs)
        for arg in args:
            if arg in ('cpu', 'cuda') or isinstance(arg, torch.device):
                if self.w_ewma is not None:
                    self.w_ewma = self.w_ewma.to(arg)
                    break
        return self

    @property
    def use_noise(self):
        return self._use_noise

    @use_noise.setter
    def use_noise(self, mode):
        if self.training:
            raise Exception('Once use_noise argument is set, it cannot be changed for training purposes. '
                            'It can, however, be changed in eval mode.')
        elif not self._trained_with_noise:
            raise Exception('Model was not trained with noise, so cannot use noise in eval mode.')
        self._use_noise = mode

    @property
    def w_eval_psi(self):
        return self._w_eval_psi

    @w_eval_psi.setter
    def w_eval_psi(self, new_w_eval_psi):
        if not self.training:
            self._w_eval_psi = new_w_eval_psi
        else:
            raise Exception('Can only alter psi value for truncation trick on w during evaluation mode.')

    @property
    def trunc_cutoff_stage(self):
        return self._trunc_cutoff_stage

    @trunc_cutoff_stage.setter
    def trunc_cutoff_stage(self, new_trunc_cutoff_stage):
        if not self.training:
            _final_stage = int(np.log2(self.final_res)) - 1
            if (isinstance(new_trunc_cutoff_stage, int) and 0 < new_trunc_cutoff_stage <= _final_stage) or \
                    new_trunc_cutoff_stage is None:
                self._trunc_cutoff_stage = new_trunc_cutoff_stage
            else:
                raise ValueError(f'Input cutoff stage for truncation trick on w must be of type `int` in range '
                                 f'(0,{_final_stage}] or `None`.')
        else:
            raise Exception('Can only alter cutoff stage for truncation trick on w during evaluation mode.')

    # ------------------------------------------------------------------ fused layer pieces
    def _layer_conv(self, layer, out):
        """layer[0]: Sequential([upsampler], Conv2dEx, [blur]) (reference :331) or the bare Conv2dEx of layer 1."""
        op = layer[0]
        if isinstance(op, Conv2dEx):
            return op(out)
        mods, i = list(op), 0
        while i < len(mods):
            if isinstance(mods[i], Upsample2x) and i + 1 < len(mods) and isinstance(mods[i + 1], Conv2dEx):
                out = mods[i + 1](out, up=True)             # the upsampler rides in the convolution (ops.upconv2d)
                i += 2
            else:
                out = mods[i](out)
                i += 1
        return out

    def _layer_tail(self, layer, out, w, noise, style=None):
        """layer[1] noise -> layer[2] (bias, lrelu, norms) -> layer[3] style affine + AdaIN (reference :500-526):
        one fused kernel pair when the layer has the default composition, the unfused sequence otherwise."""
        tail = layer[2]
        if style is None:
            style = layer[3](w)                                         # [N, 2C] = [ys ; yb]
        mods = list(tail)
        bias = mods[0] if mods and isinstance(mods[0], Conv2dBias) else None
        fusable = (self.use_instancenorm and not self.use_pixelnorm and bias is not None and bias.bias_scale == 1.
                   and (layer[1] is not None) == bool(self.use_noise))
        if fusable:
            nz = nw = None
            if self.use_noise:
                if noise is not None and not self.training:
                    nz = noise
                elif getattr(self, '_noise_pool', None) is not None:
                    nz = self._take_noise(out)
                else:
                    nz = layer[1].draw(out)
                nw = layer[1].noise_weight
            return ops.style_epilogue(out, nz, nw, bias.bias, style, self.nl.negative_slope, 1.e-8)
        if self.use_noise:
            out = layer[1](out, noise=noise)
        out = tail(out)
        y = style.view(-1, 2, layer[3].nout_feat // 2, 1, 1)
        return out * (y[:, 0].contiguous().add(1)) + y[:, 1].contiguous()

    # ------------------------------------------------------------------ pooled noise
    # The reference draws one N(0,1) map per layer (`StyleAddNoise`, :112-119): 14 `normal_` launches per generator pass at 128^2.
    # With the default random source the maps of a training pass are slices of ONE draw (same distribution, one launch); a taped
    # source (parity tests) keeps the reference's per-layer draws.
    def _begin_noise_pool(self, bs, device):
        self._noise_pool = None
        if not (self.training and self.use_noise and type(RANDOM.source) is RandomSource):
            return
        total = sum(bs * (4 << (n // 2)) ** 2 for n in range(len(self.gen_layers)))
        self._noise_pool = RANDOM.source.randn((total,), device)
        self._noise_off = 0

    def _take_noise(self, out):
        n_el = out.shape[0] * out.shape[2] * out.shape[3]
        if self._noise_off + n_el > self._noise_pool.numel():          # (a layer list the pool was not sized for)
            return RANDOM.source.randn((out.shape[0], 1, out.shape[2], out.shape[3]), out.device)
        nz = self._noise_pool[self._noise_off:self._noise_off + n_el].view(out.shape[0], 1, out.shape[2], out.shape[3])
        self._noise_off += n_el
        return nz

    def _second_w(self, bs, device):
        z2 = gen_rand_latent_vars(num_samples=bs, length=self.len_latent, distribution=self.latent_distribution, device=device)
        z2.requires_grad_(True)
        return self.z_to_w(z2)

    # ------------------------------------------------------------------ forward
    def _device_mixing_ws(self, w1, w2):
        """Mixing regularisation decided ON THE DEVICE (CUDA-graph replayable; reference :417-422, :505-511 draw the
        decision and the cut-off with host RNGs, which would bake one outcome into a captured graph): the second
        latent is always mapped, `fire ~ U(0,1) < pct` and `cutoff ~ randint(1, hi)` are device scalars, and layer n
        uses w2 iff fire and n >= cutoff.  Same distribution as the reference, different random stream."""
        hi = 2 * self.scale_stage if self.alpha != 0 else 2 * self.scale_stage - 2
        fire = torch.rand((), device=w1.device) < self.pct_mixing_reg
        cut = torch.randint(1, max(hi, 2), (), device=w1.device)
        layers = torch.arange(len(self.gen_layers), device=w1.device)
        sel = (fire & (layers >= cut)).view(-1, 1, 1)
        return torch.where(sel, w2.unsqueeze(0), w1.unsqueeze(0))           # [L, N, len_dlatent]

    def forward(self, x, x_mixing=None, style_mixing_stage: int = None, noise=None):
        """reference stylegan/architectures.py:411-528."""
        cutoff_idx = None
        device_mix = self._use_mixing_reg and self.training and getattr(self, 'device_mixing', False)
        if self._use_mixing_reg and not device_mix:
            if RANDOM.source.host_uniform() < self.pct_mixing_reg:
                if self.alpha != 0:
                    cutoff_idx = RANDOM.source.host_randint(1, 2 * self.scale_stage)
                else:
                    cutoff_idx = RANDOM.source.host_randint(1, 2 * self.scale_stage - 2)

        bs = x.shape[0]
        w2 = None
        if device_mix and 2 * bs <= 16:
            # both latents through the mapping network as ONE batch of 2N rows: half the (launch-bound) small linear
            # kernels of the mapping network, forward and backward
            z2 = gen_rand_latent_vars(num_samples=bs, length=self.len_latent, distribution=self.latent_distribution,
                                      device=x.device)
            z2.requires_grad_(True)
            both = self.z_to_w(torch.cat([x.view(bs, -1), z2], dim=0))
            x, w2 = both[:bs], both[bs:]
        else:
            x = self.z_to_w(x)
            if device_mix:
                w2 = self._second_w(bs, x.device)

        if self.use_truncation_trick:
            if self.training:
                with torch.no_grad():
                    if self.w_ewma is None:
                        self.w_ewma = torch.empty(x.shape[1], device=x.device, dtype=torch.float32)
                        K.w_ewma_update(x.detach(), self.w_ewma, 0.0)
                    else:
                        K.w_ewma_update(x.detach(), self.w_ewma, self.w_ewma_beta)
            elif self.trunc_cutoff_stage is not None:
                x = self.w_ewma.expand_as(x) + self.w_eval_psi * (x - self.w_ewma.expand_as(x))

        ws = self._device_mixing_ws(x, w2) if device_mix else None
        out = self.const_input.expand(bs, -1, -1, -1)

        # All style affines depend only on the dlatents: with `side_stream_styles` they run on a second stream (a parallel
        # branch of the captured graph, forward and -- because autograd replays each node on its forward's stream --
        # backward), so these launch-bound small kernels overlap the convolution chain instead of sitting in it.
        styles = None
        if ws is not None and self.grouped_styles and bs <= 16 and bs * ws.shape[2] <= 8192 and not self.fade_in_phase:
            # all style affines of the pass in ONE launch (and one dgrad + one wgrad launch in backward, which also hand
            # back the gradient of the whole [L,N,K] dlatent stack: no per-layer select / accumulate kernels)
            lins = [layer[3] for layer in self.gen_layers]
            sts = ops.grouped_linear(ws, self._style_table,
                                     [(l.linear.weight, l.linear.bias, l.alpha, l.lrmul if l.use_lrmul else 1.) for l in lins])
            styles = [(st, None) for st in sts]
        elif ws is not None and getattr(self, 'side_stream_styles', False) and x.is_cuda and not self.fade_in_phase:
            main = torch.cuda.current_stream()
            if getattr(self, '_style_stream', None) is None:
                self._style_stream = torch.cuda.Stream()
            side = self._style_stream
            side.wait_stream(main)
            styles = []
            with torch.cuda.stream(side):
                for n, layer in enumerate(self.gen_layers):
                    st = layer[3](ws[n])
                    ev = torch.cuda.Event()
                    ev.record(side)
                    styles.append((st, ev))

        self._begin_noise_pool(bs, x.device)
        if self.fade_in_phase:
            for n, layer in enumerate(self.gen_layers[:-2]):
                if n:
                    out = self._layer_conv(layer, out)
                if n == cutoff_idx:
                    x = self._second_w(bs, x.device)
                out = self._layer_tail(layer, out, x if ws is None else ws[n], noise[n] if noise is not None else None)
            skip = self.prev_torgb(out)
            n += 1
            if n == cutoff_idx:
                x = self._second_w(bs, x.device)
            out = self._layer_conv(self.gen_layers[-2], out)
            out = self._layer_tail(self.gen_layers[-2], out, x if ws is None else ws[n], noise[-2] if noise is not None else None)
            n += 1
            if n == cutoff_idx:
                x = self._second_w(bs, x.device)
            out = self._layer_conv(self.gen_layers[-1], out)
            out = self._layer_tail(self.gen_layers[-1], out, x if ws is None else ws[n], noise[-1] if noise is not None else None)
            self._noise_pool = None                     # (the slices handed out keep the storage alive for backward)
            return ops.fade_up_blend(skip, self.torgb(out), self._state.blend_coefs()[0])

        for n, layer in enumerate(self.gen_layers):
            if n:
                out = self._layer_conv(layer, out)
            if n == cutoff_idx:
                x = self._second_w(bs, x.device)
            if n == style_mixing_stage:
                assert (style_mixing_stage and not self.training and isinstance(x_mixing, torch.Tensor))
                x = self.z_to_w(x_mixing)
                if self.use_truncation_trick and self.trunc_cutoff_stage is not None and n < 2 * self.trunc_cutoff_stage:
                    x = self.w_ewma.expand_as(x) + self.w_eval_psi * (x - self.w_ewma.expand_as(x))
            elif self.use_truncation_trick and not self.training and self.trunc_cutoff_stage is not None and \
                    n == 2 * self.trunc_cutoff_stage:
                x = (x - self.w_ewma.expand_as(x)).div(self.w_eval_psi) + self.w_ewma.expand_as(x)
            st = None
            if styles is not None:
                st, ev = styles[n]
                if ev is not None:
                    torch.cuda.current_stream().wait_event(ev)
                    st.record_stream(torch.cuda.current_stream())
            out = self._layer_tail(layer, out, x if ws is None else ws[n], noise[n] if noise is not None else None, style=st)
        self._noise_pool = None
        return self.torgb(out)


class StyleDiscriminator(_DiscriminatorImpl, StyleGAN):
    """ProDiscriminator's body on the StyleGAN state base (reference stylegan/learner.py:142)."""

    def __init__(self, final_res, pooler=None, blur_type=None, nl=None, num_classes=0, equalized_lr=True,
                 mbstd_group_size=4, state: GrowthState = None):
        StyleGAN.__init__(self, final_res, state=state)
        self._init_disc(pooler, blur_type, nl, num_classes, equalized_lr, mbstd_group_size, 'StyleGAN')
