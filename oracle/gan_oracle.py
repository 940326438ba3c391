"""CPU oracle for the gan-lab G/D training step.  TEST INFRASTRUCTURE ONLY.

A functional (stateless, dtype-generic) restatement in plain PyTorch of every
function on the hot path of sidward14/gan-lab.  Each function
cites the reference file:line it follows (paths relative to the reference's
`gan_lab/`).  Parameters are passed as a dict keyed by the *reference's own
state_dict names* so the same dict can be loaded into the reference modules, this
oracle and the B200 modules.

Pinning: `oracle/make_golden.py` runs the unmodified reference (imported from
/root/reference in the build container) and writes `tests/golden/*.pt`;
`tests/test_oracle_golden.py` checks this restatement against those fixtures.

All randomness (z, second mixing latent, per-layer noise, mixing cutoff, WGAN-GP
eps) is an explicit argument: the reference draws them inside forward().
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]

FMAP_SAMPLES = 3   # _int.py:46
RES_INIT = 4       # _int.py:47


# --------------------------------------------------------------------------- #
# initializer.py
# --------------------------------------------------------------------------- #
def init_std(fan_in: int, gain_sq_base: float = 2.0, init: str = "he",
             fan_out: Optional[int] = None) -> float:
    """utils/initializer.py:65-80 (`_calculate_init_weight_std`)."""
    gain_sq = gain_sq_base / 2.0
    if fan_out is not None:
        fan = fan_in + fan_out
        gain_sq *= 2
    else:
        fan = fan_in
    if init == "he":
        gain_sq = 2.0 * gain_sq
    return math.sqrt(gain_sq / fan)


def conv_wscale(ni: int, ks: int, gain_sq_base: float = 2.0) -> float:
    """Runtime equalized-LR scale of a ProGAN/StyleGAN-type Conv2dEx.
    utils/custom_layers.py:183-190 + utils/initializer.py:35-37,42-62."""
    return init_std(ni * ks * ks, gain_sq_base)


def linear_wscale(nin: int, gain_sq_base: float = 2.0) -> float:
    """utils/custom_layers.py:263-270 + utils/initializer.py:48-52."""
    return init_std(nin, gain_sq_base)


def fmap_for_res(res: int, fmap_base: int = 8192, fmap_max: int = 512) -> int:
    """stylegan/base.py:16-17,76-77 / progan/base.py (`get_fmap`); scale_stage = log2(res)-1."""
    stage = int(math.log2(res)) - 1
    return min(int(fmap_base / (2 ** stage)), fmap_max)


# --------------------------------------------------------------------------- #
# utils/custom_layers.py
# --------------------------------------------------------------------------- #
def conv2d_ex(x, w, b, wscale: Optional[float], lrmul: float = 1.0, padding: int = 0):
    """Conv2dEx.forward, utils/custom_layers.py:202-211: the scale multiplies the INPUT;
    lrmul post-scales output *and* bias."""
    if wscale is not None:
        x = x * wscale
    y = F.conv2d(x, w, b, stride=1, padding=padding)
    if lrmul != 1.0:
        y = y * lrmul
    return y


def linear_ex(x, w, b, wscale: Optional[float], lrmul: float = 1.0):
    """LinearEx.forward, utils/custom_layers.py:282-291."""
    if wscale is not None:
        x = x * wscale
    y = F.linear(x, w, b)
    if lrmul != 1.0:
        y = y * lrmul
    return y


def conv2d_bias(x, bias, lrmul: float = 1.0):
    """Conv2dBias.forward, utils/custom_layers.py:222-226."""
    return x + (bias * lrmul if lrmul != 1.0 else bias)


def pixelnorm(x, eps: float = 1e-8):
    """PixelNorm2d.forward, utils/custom_layers.py:85-86."""
    return x * ((x ** 2).mean(dim=1, keepdim=True) + eps).rsqrt()


def blur3x3(x):
    """get_blur_op('binomial'), utils/custom_layers.py:41-51: depthwise [1 2 1]x[1 2 1]/16, zero pad 1."""
    c = x.shape[1]
    k = torch.tensor([[1., 2., 1.], [2., 4., 2.], [1., 2., 1.]], dtype=x.dtype, device=x.device) / 16.0
    return F.conv2d(x, k.expand(c, 1, 3, 3), stride=1, padding=1, groups=c)


def instance_norm(x, eps: float = 1e-8):
    """NormalizeLayer('InstanceNorm') = nn.InstanceNorm2d(ni, eps=1e-8), utils/custom_layers.py:98-99:
    per-(n,c) mean / biased variance, no affine, no running stats."""
    mu = x.mean(dim=(2, 3), keepdim=True)
    var = x.var(dim=(2, 3), unbiased=False, keepdim=True)
    return (x - mu) * (var + eps).rsqrt()


def mbstd_concat(x, group_size: int = 4):
    """concat_mbstd_layer, utils/custom_layers.py:117-140."""
    n, c, h, w = x.shape
    group_size = min(n, group_size)
    if n % group_size != 0:
        group_size = n
    g = n // group_size
    if group_size > 1:
        m = x.view(g, group_size, c, h, w)
        m = torch.var(m, dim=1)                      # unbiased (custom_layers.py:130)
        m = torch.sqrt(m + 1e-8)
        m = m.view(g, -1).mean(dim=1).view(g, 1)
        m = m.expand(g, h * w).view(g, 1, 1, h, w).expand(g, group_size, -1, -1, -1)
        m = m.contiguous().view(-1, 1, h, w)
    else:
        m = torch.zeros(n, 1, h, w, dtype=x.dtype, device=x.device)
    return torch.cat((x, m), dim=1)


def upsample2x(x):
    """nn.Upsample(scale_factor=2, mode='nearest'), resnetgan/learner.py:154-158."""
    return F.interpolate(x, scale_factor=2, mode="nearest")


def avgpool2(x):
    """nn.AvgPool2d(2, 2), resnetgan/learner.py:160-164."""
    return F.avg_pool2d(x, kernel_size=2, stride=2)


def lrelu(x, slope: float = 0.2):
    return F.leaky_relu(x, slope)


# --------------------------------------------------------------------------- #
# stylegan/architectures.py
# --------------------------------------------------------------------------- #
def style_add_noise(x, noise_weight, noise):
    """StyleAddNoise.forward, stylegan/architectures.py:112-119 (noise (N,1,H,W) supplied)."""
    return x + noise_weight * noise


def style_mapping(p: Params, z, *, num_fcs: int = 8, lrmul: float = 0.01, normalize_z: bool = True,
                  slope: float = 0.2, prefix: str = "z_to_w.fc_mapping_model."):
    """StyleMappingNetwork.forward, stylegan/architectures.py:29-59."""
    x = z.view(z.shape[0], -1)
    if normalize_z:
        x = pixelnorm(x)
    for i in range(num_fcs):
        w = p[f"{prefix}fc_{i}.linear.weight"]
        b = p[f"{prefix}fc_{i}.linear.bias"]
        x = lrelu(linear_ex(x, w, b, linear_wscale(w.shape[1], 2.0), lrmul), slope)
    return x


def adain(x, style, nf):
    """stylegan/architectures.py:460-462, 524-526: out*(ys+1)+yb, ys = first half."""
    y = style.view(-1, 2, nf, 1, 1)
    return x * (y[:, 0] + 1) + y[:, 1]


def _style_layer_tail(p, n, out, w, noise_n, slope, use_noise, use_instancenorm, use_pixelnorm):
    """layer[1] (noise) -> layer[2] (bias, lrelu, norms) -> layer[3] (style affine) + AdaIN.
    stylegan/architectures.py:500-526."""
    pre = f"gen_layers.{n}."
    if use_noise:
        out = style_add_noise(out, p[pre + "1.noise_weight"], noise_n)
        out = conv2d_bias(out, p[pre + "2.0.bias"])
    out = lrelu(out, slope)
    if use_pixelnorm:
        out = pixelnorm(out)
    if use_instancenorm:
        out = instance_norm(out)
    sw = p[pre + "3.linear.weight"]
    sb = p[pre + "3.linear.bias"]
    y = linear_ex(w, sw, sb, linear_wscale(sw.shape[1], 1.0))
    return adain(out, y, sw.shape[0] // 2)


def _style_layer_conv(p, n, out, blur: bool, use_noise: bool):
    """layer[0]: [upsample] -> Conv2dEx 3x3 (bias only when neither noise nor blur) -> [blur].
    stylegan/architectures.py:292-334."""
    if n == 1:
        key = f"gen_layers.{n}.0.conv2d."
        up = False
    elif n % 2 == 0:
        key = f"gen_layers.{n}.0.1.conv2d."
        up = True
    else:
        key = f"gen_layers.{n}.0.0.conv2d."
        up = False
    if up:
        out = upsample2x(out)
    w = p[key + "weight"]
    b = p.get(key + "bias")
    out = conv2d_ex(out, w, b, conv_wscale(w.shape[1], 3, 2.0), padding=1)
    if up and blur:
        out = blur3x3(out)
    return out


def style_generator_forward(p: Params, z, *, res: int, noise: Sequence[torch.Tensor],
                            z2=None, cutoff_idx: Optional[int] = None,
                            alpha: float = 1.0, fade_in: bool = False,
                            blur: bool = True, use_noise: bool = True, use_instancenorm: bool = True,
                            use_pixelnorm: bool = False, normalize_z: bool = True,
                            num_fcs: int = 8, mapping_lrmul: float = 0.01, slope: float = 0.2,
                            return_w: bool = False):
    """StyleGenerator.forward (training mode), stylegan/architectures.py:411-528.

    `noise[n]` is the (N,1,H,W) tensor layer n's StyleAddNoise would have drawn; `z2` the second
    latent drawn when `n == cutoff_idx` (mixing regularisation, :417-422, :505-511)."""
    num_layers = 2 * (int(math.log2(res)) - 1)
    w = style_mapping(p, z, num_fcs=num_fcs, lrmul=mapping_lrmul, normalize_z=normalize_z, slope=slope)
    w_first = w
    bs = w.shape[0]
    out = p["const_input"].expand(bs, -1, -1, -1)
    n_loop = num_layers - 2 if fade_in else num_layers
    for n in range(n_loop):
        if n:
            out = _style_layer_conv(p, n, out, blur, use_noise)
        # the style of layer n is computed AFTER the w switch at n == cutoff_idx (:505-524)
        if n == cutoff_idx:
            w = style_mapping(p, z2, num_fcs=num_fcs, lrmul=mapping_lrmul, normalize_z=normalize_z, slope=slope)
        out = _style_layer_tail(p, n, out, w, noise[n] if use_noise else None, slope,
                                use_noise, use_instancenorm, use_pixelnorm)
    if not fade_in:
        img = conv2d_ex(out, p["torgb.conv2d.weight"], p["torgb.conv2d.bias"],
                        conv_wscale(p["torgb.conv2d.weight"].shape[1], 1, 1.0))
        return (img, w_first) if return_w else img
    # fade-in (:444-494): both branch outputs blended
    skip = conv2d_ex(out, p["prev_torgb.conv2d.weight"], p["prev_torgb.conv2d.bias"],
                     conv_wscale(p["prev_torgb.conv2d.weight"].shape[1], 1, 1.0))
    skip = upsample2x(skip)
    for n in (num_layers - 2, num_layers - 1):
        # NB (:468-481): the reference evaluates both w switches before running the two convs, which
        # is equivalent to switching at layer n because the mapping only depends on z2.
        if n == cutoff_idx:
            w = style_mapping(p, z2, num_fcs=num_fcs, lrmul=mapping_lrmul, normalize_z=normalize_z, slope=slope)
        out = _style_layer_conv(p, n, out, blur, use_noise)
        out = _style_layer_tail(p, n, out, w, noise[n] if use_noise else None, slope,
                                use_noise, use_instancenorm, use_pixelnorm)
    new = conv2d_ex(out, p["torgb.conv2d.weight"], p["torgb.conv2d.bias"],
                    conv_wscale(p["torgb.conv2d.weight"].shape[1], 1, 1.0))
    img = skip * (1.0 - alpha) + new * alpha
    return (img, w_first) if return_w else img


def w_ewma_update(w_ewma, w, beta: float = 0.995):
    """stylegan/architectures.py:427-437."""
    if w_ewma is None:
        return w.detach().clone().mean(dim=0)
    return w.detach().mean(dim=0) * (1.0 - beta) + w_ewma * beta


# --------------------------------------------------------------------------- #
# progan/architectures.py
# --------------------------------------------------------------------------- #
def pro_generator_forward(p: Params, z, *, res: int, alpha: float = 1.0, fade_in: bool = False,
                          blur: bool = True, use_pixelnorm: bool = True, normalize_z: bool = True,
                          slope: float = 0.2):
    """ProGenerator.forward, progan/architectures.py:159-167 (blocks :78-93, :109-148)."""
    pn = pixelnorm if use_pixelnorm else (lambda t: t)
    x = z.view(z.shape[0], -1)
    if normalize_z:
        x = pixelnorm(x)
    # block 0: FC -> view -> nl -> PN -> conv -> nl -> PN   (module indices shift if PN is absent)
    w0 = p["gen_blocks.0.0.linear.weight"]
    x = linear_ex(x, w0, p["gen_blocks.0.0.linear.bias"], linear_wscale(w0.shape[1], 2.0 / 16))
    x = pn(lrelu(x.view(x.shape[0], -1, RES_INIT, RES_INIT), slope))
    ci = 4 if use_pixelnorm else 3
    wc = p[f"gen_blocks.0.{ci}.conv2d.weight"]
    x = pn(lrelu(conv2d_ex(x, wc, p[f"gen_blocks.0.{ci}.conv2d.bias"], conv_wscale(wc.shape[1], 3), padding=1), slope))
    nblocks = int(math.log2(res)) - 1

    def block(x, b):
        pre = f"gen_blocks.{b}."
        x = upsample2x(x)
        w1 = p[pre + "0.1.conv2d.weight"]
        if blur:
            x = conv2d_ex(x, w1, None, conv_wscale(w1.shape[1], 3), padding=1)
            x = blur3x3(x)
            x = conv2d_bias(x, p[pre + "0.3.bias"])
        else:
            x = conv2d_ex(x, w1, p[pre + "0.1.conv2d.bias"], conv_wscale(w1.shape[1], 3), padding=1)
        x = pn(lrelu(x, slope))
        w2 = p[pre + "1.0.conv2d.weight"]
        x = conv2d_ex(x, w2, p[pre + "1.0.conv2d.bias"], conv_wscale(w2.shape[1], 3), padding=1)
        return pn(lrelu(x, slope))

    for b in range(1, nblocks - 1):
        x = block(x, b)

    def torgb(x, key):
        w = p[key + ".conv2d.weight"]
        return conv2d_ex(x, w, p[key + ".conv2d.bias"], conv_wscale(w.shape[1], 1, 1.0))

    if nblocks == 1:
        return torgb(x, "torgb")
    if fade_in:
        return upsample2x(torgb(x, "prev_torgb")) * (1.0 - alpha) + torgb(block(x, nblocks - 1), "torgb") * alpha
    return torgb(block(x, nblocks - 1), "torgb")


def pro_discriminator_forward(p: Params, x, *, res: int, alpha: float = 1.0, fade_in: bool = False,
                              blur: bool = True, mbstd_group_size: int = 4, slope: float = 0.2):
    """ProDiscriminator.forward (= StyleDiscriminator), progan/architectures.py:309-318;
    blocks :261-284, last block :219-233, fromrgb :286-292."""
    nblocks = int(math.log2(res)) - 1          # len(disc_blocks)
    x = x.view(-1, FMAP_SAMPLES, res, res)

    def fromrgb(x, key):
        w = p[key + ".0.conv2d.weight"]
        return lrelu(conv2d_ex(x, w, p[key + ".0.conv2d.bias"], conv_wscale(w.shape[1], 1)), slope)

    def block(x, b):
        pre = f"disc_blocks.{b}."
        w1 = p[pre + "0.0.conv2d.weight"]
        x = lrelu(conv2d_ex(x, w1, p[pre + "0.0.conv2d.bias"], conv_wscale(w1.shape[1], 3), padding=1), slope)
        if blur:
            x = blur3x3(x)
            w2 = p[pre + "1.1.conv2d.weight"]
            x = conv2d_ex(x, w2, None, conv_wscale(w2.shape[1], 3), padding=1)
            x = avgpool2(x)
            x = conv2d_bias(x, p[pre + "1.3.bias"])
        else:
            w2 = p[pre + "1.0.conv2d.weight"]
            x = conv2d_ex(x, w2, None, conv_wscale(w2.shape[1], 3), padding=1)
            x = avgpool2(x)
            x = conv2d_bias(x, p[pre + "1.2.bias"])
        return lrelu(x, slope)

    def last_block(x):
        L = nblocks - 1
        pre = f"disc_blocks.{L}."
        i = 0
        if mbstd_group_size != -1:
            x = mbstd_concat(x, mbstd_group_size)
            i = 1
        w1 = p[pre + f"{i}.conv2d.weight"]
        x = lrelu(conv2d_ex(x, w1, p[pre + f"{i}.conv2d.bias"], conv_wscale(w1.shape[1], 3), padding=1), slope)
        w2 = p[pre + f"{i + 2}.conv2d.weight"]
        x = lrelu(conv2d_ex(x, w2, p[pre + f"{i + 2}.conv2d.bias"], conv_wscale(w2.shape[1], 4), padding=0), slope)
        x = x.view(x.shape[0], -1)
        wl = p[pre + f"{i + 5}.linear.weight"]
        return linear_ex(x, wl, p[pre + f"{i + 5}.linear.bias"], linear_wscale(wl.shape[1], 1.0))

    if nblocks == 1:
        return last_block(fromrgb(x, "fromrgb")).view(-1)
    if fade_in:
        x = fromrgb(avgpool2(x), "prev_fromrgb") * (1.0 - alpha) + block(fromrgb(x, "fromrgb"), 0) * alpha
    else:
        x = block(fromrgb(x, "fromrgb"), 0)
    for b in range(1, nblocks - 1):
        x = block(x, b)
    return last_block(x).view(-1)


# --------------------------------------------------------------------------- #
# ResNet GAN: resnetgan/resblocks.py:15-120, resnetgan/architectures.py:29-182
# --------------------------------------------------------------------------- #
def _conv(p: Params, key: str, x, padding: int):
    """Conv2dEx without equalized LR (ResNet GAN default, config.py:213): plain nn.Conv2d, custom_layers.py:202-211."""
    return F.conv2d(x, p[key + ".conv2d.weight"], p.get(key + ".conv2d.bias"), 1, padding)


def _batchnorm_train(p: Params, key: str, x, buffers: Optional[Params] = None, momentum: float = 0.1, eps: float = 1e-5):
    """nn.BatchNorm2d in training mode (custom_layers.py:100-102): batch statistics, biased variance for the
    normalisation; running buffers (if given) updated with the unbiased variance and momentum 0.1."""
    mean = x.mean(dim=(0, 2, 3))
    var = x.var(dim=(0, 2, 3), unbiased=False)
    if buffers is not None:
        n = x.numel() // x.shape[1]
        with torch.no_grad():
            buffers[key + ".norm.running_mean"].mul_(1 - momentum).add_(momentum * mean)
            buffers[key + ".norm.running_var"].mul_(1 - momentum).add_(momentum * var * n / (n - 1))
            buffers[key + ".norm.num_batches_tracked"] += 1
    xh = (x - mean.view(1, -1, 1, 1)) * (var.view(1, -1, 1, 1) + eps).rsqrt()
    return xh * p[key + ".norm.weight"].view(1, -1, 1, 1) + p[key + ".norm.bias"].view(1, -1, 1, 1)


def _layernorm(p: Params, key: str, x, eps: float = 1e-5):
    """nn.LayerNorm([C,H,W]) with elementwise affine (custom_layers.py:103-106)."""
    mean = x.mean(dim=(1, 2, 3), keepdim=True)
    var = x.var(dim=(1, 2, 3), unbiased=False, keepdim=True)
    return (x - mean) * (var + eps).rsqrt() * p[key + ".norm.weight"] + p[key + ".norm.bias"]


def resblock2d(p: Params, key: str, x, *, norm: str, up: bool = False, pool: bool = False, ni: int, nf: int,
               pix32: bool = False, buffers: Optional[Params] = None):
    """ResBlock2d / ResBlock2d32Pix.forward (resblocks.py:15-79): skip(x) + conv_layer_2(conv_layer_1(x)) with the module
    indices the reference's Sequentials give the convs (blur_type = None)."""
    nrm = (lambda k, t: _batchnorm_train(p, k, t, buffers)) if norm == "BatchNorm" else (lambda k, t: _layernorm(p, k, t))
    h = torch.relu(nrm(key + ".conv_layer_1.0", x))
    if up:
        h = _conv(p, key + ".conv_layer_1.3", upsample2x(h), 1)
    else:
        h = _conv(p, key + ".conv_layer_1.2", h, 1)
    h = torch.relu(nrm(key + ".conv_layer_2.0", h))
    h = _conv(p, key + ".conv_layer_2.2", h, 1)
    if pool:
        h = avgpool2(h)
    if up:
        s = _conv(p, key + ".skip_connection.1", upsample2x(x), 0)
    elif pool and pix32:
        s = avgpool2(_conv(p, key + ".skip_connection.0", x, 0))                      # resblocks.py:75-76
    elif pool:
        s = _conv(p, key + ".skip_connection.1", avgpool2(x), 0)
    elif ni != nf:
        s = _conv(p, key + ".skip_connection.0", x, 0)
    else:
        s = x
    return s + h


def resnet_generator_forward(p: Params, z, *, res: int, fmap: int = 64, buffers: Optional[Params] = None):
    """Generator64PixResnet / Generator32PixResnet.forward (architectures.py:63-99 / 29-61), training mode."""
    len_latent = z.shape[1]
    if res == 64:
        c0, plan, blocks = len_latent * 4, [8 * fmap, 4 * fmap, 2 * fmap, fmap], (3, 4, 5, 6)
    elif res == 32:
        c0, plan, blocks = len_latent, [fmap, fmap, fmap], (3, 4, 5)
    else:
        raise ValueError(res)
    x = F.linear(z, p["generator_model.1.linear.weight"], p["generator_model.1.linear.bias"]).view(-1, c0, RES_INIT, RES_INIT)
    ni = c0
    for idx, nf in zip(blocks, plan):
        x = resblock2d(p, f"generator_model.{idx}", x, norm="BatchNorm", up=True, ni=ni, nf=nf, buffers=buffers)
        ni = nf
    n = blocks[-1] + 1
    x = torch.relu(_batchnorm_train(p, f"generator_model.{n}", x, buffers))
    return torch.tanh(_conv(p, f"generator_model.{n + 2}", x, 1))


def resnet_discriminator_forward(p: Params, x, *, res: int, fmap: int = 64):
    """Discriminator64PixResnet / Discriminator32PixResnet.forward (architectures.py:150-182 / 105-133)."""
    x = x.view(-1, FMAP_SAMPLES, res, res)
    if res == 64:
        x = _conv(p, "conv1", x, 1)
        ni = fmap
        for b, nf in enumerate([2 * fmap, 4 * fmap, 8 * fmap, 8 * fmap]):
            x = resblock2d(p, f"resblocks.{b}", x, norm="LayerNorm", pool=True, ni=ni, nf=nf)
            ni = nf
        x = x.reshape(-1, 16 * 8 * fmap)
    elif res == 32:
        # FastResBlock2dDownsample (resblocks.py:82-120): conv-relu-conv-pool + pool-conv1x1 skip
        h = torch.relu(_conv(p, "conv1.conv_layer_1.0", x, 1))
        h = avgpool2(_conv(p, "conv1.conv_layer_2.0", h, 1))
        x = _conv(p, "conv1.skip_connection.1", avgpool2(x), 0) + h
        x = resblock2d(p, "resblocks.0", x, norm="LayerNorm", pool=True, ni=fmap, nf=fmap, pix32=True)
        x = resblock2d(p, "resblocks.1", x, norm="LayerNorm", ni=fmap, nf=fmap, pix32=True)
        x = resblock2d(p, "resblocks.2", x, norm="LayerNorm", ni=fmap, nf=fmap, pix32=True)
        x = F.avg_pool2d(torch.relu(x), res // 4).view(-1, fmap)
    else:
        raise ValueError(res)
    return F.linear(x, p["linear1.linear.weight"], p["linear1.linear.bias"]).view(-1)


# --------------------------------------------------------------------------- #
# losses / penalties: progan/learner.py:791-812, 883-896; resnetgan/learner.py:780-827
# --------------------------------------------------------------------------- #
def gradient_penalty(d_fn: Callable, gp_type: str, real, fake, lda: float = 10.0, gamma: float = 1.0,
                     eps=None):
    """GANLearner.calc_gp (the METHOD the train loops call), resnetgan/learner.py:780-827.
    NB the 2-norm is over the channel axis only (:820-825)."""
    if gp_type == "wgan-gp":
        xb = eps * fake.detach() + (1 - eps) * real.detach()
    elif gp_type == "r1":
        xb = real.detach()
    elif gp_type == "r2":
        xb = fake.detach()
    else:
        raise ValueError(gp_type)
    xb = xb.clone().requires_grad_(True)
    outb = d_fn(xb)
    g = torch.autograd.grad(outb, xb, grad_outputs=torch.ones_like(outb), create_graph=True,
                            retain_graph=True, only_inputs=True)[0]
    if gp_type == "wgan-gp":
        if gamma != 1.0:
            return ((g.norm(2, dim=1) - gamma) ** 2 / gamma ** 2).mean() * lda
        return ((g.norm(2, dim=1) - 1.0) ** 2).mean() * lda / 2.0
    return (g.norm(2, dim=1) ** 2).mean() * lda / 2.0


def disc_loss(d_fn: Callable, fake, real, *, loss: str = "nonsaturating", gp_type: Optional[str] = "r1",
              lda: float = 10.0, gamma: float = 1.0, eps_drift: float = 0.001, gp_eps=None):
    """D-step loss, progan/learner.py:788-812."""
    d_gen = d_fn(fake)
    d_real = d_fn(real)
    if loss == "wgan":
        l = (d_gen - d_real).mean()
    else:
        l = F.binary_cross_entropy_with_logits(d_gen, torch.zeros_like(d_gen)) + \
            F.binary_cross_entropy_with_logits(d_real, torch.ones_like(d_real))
    if gp_type is not None:
        l = l + gradient_penalty(d_fn, gp_type, real, fake, lda, gamma, gp_eps)
    if eps_drift > 0:
        l = l + (d_real ** 2).mean() * eps_drift
    return l


def gen_loss(d_out, loss: str = "nonsaturating"):
    """G-step loss, progan/learner.py:883-896."""
    if loss == "wgan":
        return -d_out.mean()
    if loss == "nonsaturating":
        return F.binary_cross_entropy_with_logits(d_out, torch.ones_like(d_out))
    if loss == "minimax":
        return -F.binary_cross_entropy_with_logits(d_out, torch.zeros_like(d_out))
    raise ValueError(loss)


# --------------------------------------------------------------------------- #
# optimiser / EWMA: utils/backprop_utils.py:109-120 (torch.optim.Adam), progan/learner.py:909-916
# --------------------------------------------------------------------------- #
def adam_step(p, g, m, v, step: int, lr: float, beta1: float = 0.0, beta2: float = 0.99,
              eps: float = 1e-8, wd: float = 0.0):
    """One torch.optim.Adam update (non-amsgrad), returning (p, m, v); `step` is 1-based."""
    if wd != 0.0:
        g = g + wd * p
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    p = p - (lr / bc1) * (m / denom)
    return p, m, v


def ewma_step(lagged, p, beta: float):
    """progan/learner.py:909-916."""
    return p * (1.0 - beta) + lagged * beta


def ewma_beta(batch_size: int, gen_bs_mult: int = 1, half_life: float = 10.0) -> float:
    """progan/learner.py:1124-1127."""
    return 0.5 ** ((batch_size * gen_bs_mult) / (half_life * 1000.0)) if half_life > 0 else 0.0


def fade_real_images(x, alpha: float):
    """Host-side real-image fade, progan/learner.py:770-779 (avg-pool down, nearest up)."""
    return upsample2x(avgpool2(x)) * (1.0 - alpha) + x * alpha
