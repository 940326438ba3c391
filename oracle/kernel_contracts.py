"""CPU restatement of every kernel launcher's CONTRACT in gan_lab_b200/_kernels.py.  TEST INFRASTRUCTURE ONLY.

Two uses, both from tests/ only:
  * `-m gpu`: each CUDA launcher is compared with the function of the same name here on the same inputs
    (per-kernel parity, fp64 available);
  * `-m "not gpu"`: `install(monkeypatch)` swaps these in for the launchers so that the host-side logic above
    the C-ABI (autograd wiring incl. the double backward, module/learner control flow, optimiser tables) is
    checked on CPU against the golden fixtures.  The product never imports this module and has no such path.

Everything is written with plain torch ops on the reference's arithmetic (cf. oracle/gan_oracle.py).
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch
import torch.nn.functional as F

from . import gan_oracle as O

ACT_NONE, ACT_LRELU = 0, 1


def _act(v, act, slope):
    return torch.where(v > 0, v, v * slope) if act == ACT_LRELU else v


def _dact(y, act, slope):
    return torch.where(y > 0, torch.ones_like(y), torch.full_like(y, slope)) if act == ACT_LRELU else torch.ones_like(y)


def _cl(t):
    return t.contiguous(memory_format=torch.channels_last) if t.dim() == 4 else t.contiguous()


def _b(bias, nd):
    if bias is None:
        return None
    return bias.reshape(1, -1, 1, 1) if nd == 4 else bias.reshape(1, -1)


# ---- conv / linear -------------------------------------------------------------------------------------
def conv_fprop(x, w, bias, pad, alpha, bias_scale, act, slope):
    y = alpha * F.conv2d(x, w, None, 1, pad)
    if bias is not None:
        y = y + bias_scale * _b(bias, 4)
    return _cl(_act(y, act, slope))


def conv_dgrad(gy, w, x_hw, pad, alpha):
    n = gy.shape[0]
    return _cl(alpha * torch.nn.grad.conv2d_input((n, w.shape[1], x_hw[0], x_hw[1]), w, gy, 1, pad))


def conv_wgrad(x, gy, rs, pad, alpha):
    return _cl(alpha * torch.nn.grad.conv2d_weight(x, (gy.shape[1], x.shape[1], rs[0], rs[1]), gy, 1, pad))


# conv3x3_same(upsample2x_nearest(x)) and its two gradients, stated as the literal composition the reference runs
# (nn.Upsample + Conv2dEx, stylegan/architectures.py:155-156): what glb_upconv_* must reproduce without the upsampled copy
def upconv_fprop(x, w, bias, alpha, bias_scale, act, slope):
    return conv_fprop(O.upsample2x(x), w, bias, 1, alpha, bias_scale, act, slope)


def upconv_dgrad(gy, w, alpha):
    return upsample2x_bwd(conv_dgrad(gy, w, (gy.shape[2], gy.shape[3]), 1, alpha))


def upconv_wgrad(x, gy, alpha):
    return conv_wgrad(O.upsample2x(x), gy, (3, 3), 1, alpha)


# avgpool2x2(conv3x3_same(x)) + bias + activation and its gradients, as the literal composition the reference runs
# (Conv2dEx + nn.AvgPool2d + Conv2dBias + LeakyReLU, progan/architectures.py:267-284): what glb_downconv_* must reproduce
def downconv_fprop(x, w, bias, alpha, bias_scale, act, slope):
    v = F.avg_pool2d(alpha * F.conv2d(x, w, None, 1, 1), 2, 2)
    if bias is not None:
        v = v + bias_scale * _b(bias, 4)
    return _cl(_act(v, act, slope))


def downconv_dgrad(gy, w, alpha):
    g = 0.25 * O.upsample2x(gy)
    return conv_dgrad(g, w, (g.shape[2], g.shape[3]), 1, alpha)


def downconv_wgrad(x, gy, alpha):
    return conv_wgrad(x, 0.25 * O.upsample2x(gy), (3, 3), 1, alpha)


def linear_fwd(x, w, bias, alpha, bias_scale, act, slope):
    y = alpha * (x @ w.t())
    if bias is not None:
        y = y + bias_scale * bias.reshape(1, -1)
    return _act(y, act, slope)


def linear_dgrad(gy, w, alpha):
    return alpha * (gy @ w)


def linear_wgrad(x, gy, alpha):
    return alpha * (gy.t() @ x)


# ---- elementwise / reductions ---------------------------------------------------------------------------
def bias_act_fwd(x, bias, bias_scale, act, slope):
    v = x if bias is None else x + bias_scale * _b(bias, x.dim())
    return _cl(_act(v, act, slope))


def act_bwd(gy, y, want_bias, bias_scale, act, slope):
    gx = gy * _dact(y, act, slope)
    dims = (0, 2, 3) if gx.dim() == 4 else (0,)
    return _cl(gx), (bias_scale * gx.sum(dims) if want_bias else None)


def colsum(x, scale):
    return scale * x.sum((0, 2, 3) if x.dim() == 4 else (0,))


def _coef(v):
    """A coefficient handed over as a device-side handle (gan_lab_b200._kernels.Coef) is read from the tensor the kernel would
    read -- not from the host-side mirror -- so that a missing DeviceAlpha.set() shows up in the host-wiring tests."""
    if hasattr(v, "ref") and hasattr(v, "idx"):
        return float(v.ref.coef[v.idx])
    return v


def axpby(a, b, alpha, beta):
    alpha, beta = _coef(alpha), _coef(beta)
    return _cl(alpha * a + (beta * b if b is not None else 0))


def scale_by(x, s_dev, scale):
    return x * (scale * s_dev.reshape(()))


def sumsq(x, scale):
    return scale * (x * x).sum()


def gp_norm_fwd(g, gamma, scale):
    return scale * ((g.pow(2).sum(1).sqrt() - gamma) ** 2).sum()


def gp_norm_bwd(g, s_dev, gamma, scale):
    nrm = g.pow(2).sum(1, keepdim=True).sqrt()
    k = torch.where(nrm > 0, 2 * scale * (nrm - gamma) / nrm.clamp_min(1e-30), torch.zeros_like(nrm))
    return g * k * s_dev.reshape(())


def interp_rows(a, b, eps):
    e = eps.reshape(-1, *([1] * (a.dim() - 1)))
    return e * a + (1 - e) * b


def pixelnorm_fwd(x, eps):
    return _cl(O.pixelnorm(x, eps))


def pixelnorm_bwd(gy, x, eps):
    xx = x.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        y = O.pixelnorm(xx, eps)
    return _cl(torch.autograd.grad(y, xx, gy)[0])


def blur3x3(x):
    return _cl(O.blur3x3(x))


def blur_act_bwd(gz, y, want_bias, bias_scale, act, slope):
    g = blur3x3(gz) * _dact(y, act, slope)
    return _cl(g), (bias_scale * g.sum(dim=(0, 2, 3)) if want_bias else None)


def upsample2x_fwd(x):
    return _cl(O.upsample2x(x))


def upsample2x_bwd(gy):
    return _cl(F.avg_pool2d(gy, 2, 2) * 4)


def pool_bias_act_fwd(x, bias, bias_scale, act, slope):
    v = F.avg_pool2d(x, 2, 2)
    if bias is not None:
        v = v + bias_scale * _b(bias, 4)
    return _cl(_act(v, act, slope))


def pool_bias_act_bwd(gy, y, want_bias, bias_scale, act, slope):
    g = gy * _dact(y, act, slope) if y is not None else gy
    gx = 0.25 * F.interpolate(g, scale_factor=2, mode="nearest")
    return _cl(gx), (bias_scale * g.sum((0, 2, 3)) if want_bias else None)


# ---- StyleGAN epilogue -----------------------------------------------------------------------------------
def _se(x, noise, nw, bias, style, slope, eps):
    u = x
    if noise is not None:
        u = u + nw.reshape(1, -1, 1, 1) * noise
    if bias is not None:
        u = u + bias.reshape(1, -1, 1, 1)
    t = torch.where(u > 0, u, u * slope)
    return O.adain(O.instance_norm(t, eps), style, x.shape[1])


def style_epilogue_fwd(x, noise, noise_weight, bias, style, slope, eps):
    out = _se(x, noise, noise_weight, bias, style, slope, eps)
    return _cl(out), torch.tensor([float(eps)])     # "stats" is opaque to the caller; the double keeps eps there


def style_epilogue_bwd(gout, x, noise, noise_weight, bias, style, stats, slope):
    eps = float(stats[0])
    leaves = [x.detach().clone().requires_grad_(True), style.detach().clone().requires_grad_(True)]
    nw = noise_weight.detach().clone().requires_grad_(True) if noise_weight is not None else None
    b = bias.detach().clone().requires_grad_(True) if bias is not None else None
    with torch.enable_grad():
        out = _se(leaves[0], noise, nw, b, leaves[1], slope, eps)
    ins = leaves + [t for t in (nw, b) if t is not None]
    gs = list(torch.autograd.grad(out, ins, gout, allow_unused=True))
    gx, gstyle = gs[0], gs[1]
    rest = gs[2:]
    g_nw = rest.pop(0).reshape(-1) if nw is not None else None
    g_b = rest.pop(0).reshape(-1) if b is not None else None
    return _cl(gx), gstyle, g_nw, g_b


# ---- minibatch stddev ---------------------------------------------------------------------------------------
def mbstd_fwd(x, group):
    return _cl(O.mbstd_concat(x, group))


def mbstd_bwd(gy, x, group):
    xx = x.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        y = O.mbstd_concat(xx, group)
    return _cl(torch.autograd.grad(y, xx, gy)[0])


def mbstd_bwdbwd(v, gy, x, group):
    xx = x.detach().clone().requires_grad_(True)
    gg = gy.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        y = O.mbstd_concat(xx, group)
        gx, = torch.autograd.grad(y, xx, gg, create_graph=True)
    ggx, ggy = torch.autograd.grad(gx, (xx, gg), v, allow_unused=True)
    if ggx is None:
        ggx = torch.zeros_like(x)
    return _cl(ggx), _cl(ggy)


# ---- RGB 1x1 convs ---------------------------------------------------------------------------------------------
def _wmat(w, ws_j, ws_c, C):
    """-> [3, C] matrix with element (j, c) = w.flat[j*ws_j + c*ws_c]."""
    flat = w.reshape(-1)
    j = torch.arange(3).view(3, 1)
    c = torch.arange(C).view(1, C)
    return flat[j * ws_j + c * ws_c]


def rgb_expand(img, w, ws_j, ws_c, C, bias, pool, alpha, bias_scale, act, slope):
    m = _wmat(w, ws_j, ws_c, C)                         # [3, C]
    if pool:
        img = F.avg_pool2d(img, 2, 2)
    y = alpha * torch.einsum("njhw,jc->nchw", img, m)
    if bias is not None:
        y = y + bias_scale * bias.reshape(1, -1, 1, 1)
    return _cl(_act(y, act, slope))


def rgb_contract(x, w, ws_j, ws_c, bias, pool, alpha, bias_scale):
    m = _wmat(w, ws_j, ws_c, x.shape[1])
    img = alpha * torch.einsum("nchw,jc->njhw", x, m)
    if bias is not None:
        img = img + bias_scale * bias.reshape(1, -1, 1, 1)
    if pool:
        img = 0.25 * F.interpolate(img, scale_factor=2, mode="nearest")
    return img.contiguous()


def rgb_wgrad(img, g, w_shape, ws_j, ws_c, pool, alpha):
    if pool:
        img = F.avg_pool2d(img, 2, 2)
    C = g.shape[1]
    m = alpha * torch.einsum("njhw,nchw->jc", img, g)    # [3, C]
    out = torch.zeros(int(np.prod(w_shape)), dtype=m.dtype)
    j = torch.arange(3).view(3, 1)
    c = torch.arange(C).view(1, C)
    out[(j * ws_j + c * ws_c).reshape(-1)] = m.reshape(-1)
    return out.reshape(w_shape)


def plane_sum(img, scale):
    return scale * img.sum((0, 2, 3))


# ---- fade-in / losses / misc ---------------------------------------------------------------------------------------
def fade_up_blend(lo, hi, alpha):
    alpha = _coef(alpha)
    return (1 - alpha) * O.upsample2x(lo) + alpha * hi


def fade_up_blend_bwd(gout, alpha):
    alpha = _coef(alpha)
    return (1 - alpha) * F.avg_pool2d(gout, 2, 2) * 4, alpha * gout


def fade_real(x, alpha):
    return O.fade_real_images(x, _coef(alpha))


def d_logit_loss(d_gen, d_real, kind, eps_drift):
    a = d_gen.detach().clone().requires_grad_(True)
    b = d_real.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        if kind == "wgan":
            l = (a - b).mean()
        else:
            l = F.binary_cross_entropy_with_logits(a, torch.zeros_like(a)) + F.binary_cross_entropy_with_logits(b, torch.ones_like(b))
        l = l + eps_drift * (b ** 2).mean()
    ga, gb = torch.autograd.grad(l, (a, b))
    return l.detach(), ga, gb


def g_logit_loss(d_out, kind):
    a = d_out.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        l = O.gen_loss(a, kind)
    ga, = torch.autograd.grad(l, a)
    return l.detach(), ga


def w_ewma_update(w, ewma, beta):
    m = w.mean(0)
    ewma.copy_(m if beta == 0 else m * (1 - beta) + ewma * beta)
    return ewma


def _from_ptr(ptr, n):
    return torch.from_numpy(np.ctypeslib.as_array((ctypes.c_float * n).from_address(ptr)))


def adam_hyper_advance(hyper, beta1, beta2):
    t = float(hyper[3]) + 1.0
    hyper[3] = t
    hyper[1] = 1.0 - beta1 ** t
    hyper[2] = 1.0 - beta2 ** t


def adam_ewma_multi(ptr_table, sizes, T, max_size, hyper, beta1, beta2, eps, wd, ewma_beta, ewma_mode):
    lr, bc1, bc2 = float(hyper[0]), float(hyper[1]), float(hyper[2])
    ptr_table = ptr_table.view(-1, 5)
    for t in range(T):
        pp, gp, mp, vp, lp = [int(v) for v in ptr_table[t]]
        n = int(sizes[t])
        p = _from_ptr(pp, n)
        if gp != 0:
            g, m, v = _from_ptr(gp, n), _from_ptr(mp, n), _from_ptr(vp, n)
            if wd != 0:
                g = g + wd * p
            m.copy_(beta1 * m + (1 - beta1) * g)
            v.copy_(beta2 * v + (1 - beta2) * g * g)
            p.sub_((lr / bc1) * (m / (v.sqrt() / np.sqrt(bc2) + eps)))
        if ewma_mode != 0 and lp != 0:
            lag = _from_ptr(lp, n)
            prev = p if ewma_mode == 2 else lag
            lag.copy_(p * (1 - ewma_beta) + prev * ewma_beta)


# ---- ResNet-GAN norms (csrc/norm.cu); second-order terms come from torch's own autograd on F.layer_norm ------------
def _ln_core(x, gamma, beta, eps, act, slope):
    return _act(F.layer_norm(x, tuple(x.shape[1:]), gamma, beta, eps), act, slope)


def layernorm_fwd(x, gamma, beta, eps, act, slope):
    y = _ln_core(x, gamma, beta, eps, act, slope)
    mean = x.mean(dim=(1, 2, 3))
    var = x.var(dim=(1, 2, 3), unbiased=False)
    return _cl(y), torch.stack([mean, (var + eps).rsqrt()], dim=1)


def _ln_bwd_graph(gy, y, x, gamma, stats, act, slope):
    """gx, ggamma, gbeta as differentiable functions of (gy, x, gamma) -- the mask act'(y) is a constant."""
    n = x.shape[0]
    mean, rstd = stats[:, 0].view(n, 1, 1, 1), stats[:, 1].view(n, 1, 1, 1)
    eps_eff = (1.0 / rstd ** 2 - x.var(dim=(1, 2, 3), unbiased=False, keepdim=True)).detach()   # recover eps per sample
    mu = x.mean(dim=(1, 2, 3), keepdim=True)
    var = x.var(dim=(1, 2, 3), unbiased=False, keepdim=True)
    r = (var + eps_eff).rsqrt()
    xh = (x - mu) * r
    gm = gy * _dact(y, act, slope)
    g = gm * gamma
    gx = r * (g - g.mean(dim=(1, 2, 3), keepdim=True) - xh * (g * xh).mean(dim=(1, 2, 3), keepdim=True))
    return gx, (gm * xh).sum(0), gm.sum(0)


def layernorm_bwd(gy, y, x, gamma, stats, act, slope, want_gx=True, want_params=True):
    gx, gg, gb = _ln_bwd_graph(gy, y, x, gamma, stats, act, slope)
    return (_cl(gx) if want_gx else None), (gg if want_params else None), (gb if want_params else None)


def layernorm_bwdbwd(u, gy, y, x, gamma, stats, act, slope, want_gy=True, want_x=True, want_gamma=True):
    with torch.enable_grad():
        gy_, x_, gamma_ = (t.detach().clone().requires_grad_(True) for t in (gy, x, gamma))
        gx, _, _ = _ln_bwd_graph(gy_, y.detach(), x_, gamma_, stats.detach(), act, slope)
        g_gy, g_x, g_gamma = torch.autograd.grad(gx, (gy_, x_, gamma_), u)
    return (_cl(g_gy) if want_gy else None), (_cl(g_x) if want_x else None), (g_gamma if want_gamma else None)


def batchnorm_fwd(x, gamma, beta, running_mean, running_var, num_batches_tracked, eps, momentum, act, slope):
    mean = x.mean(dim=(0, 2, 3))
    var = x.var(dim=(0, 2, 3), unbiased=False)
    y = F.batch_norm(x, running_mean, running_var, gamma.reshape(-1), beta.reshape(-1), True, momentum, eps)
    if num_batches_tracked is not None:
        num_batches_tracked += 1
    return _cl(_act(y, act, slope)), torch.stack([mean, (var + eps).rsqrt()], dim=0)


def batchnorm_bwd(gy, y, x, gamma, stats, act, slope):
    c = x.shape[1]
    mean, rstd = stats[0].view(1, c, 1, 1), stats[1].view(1, c, 1, 1)
    xh = (x - mean) * rstd
    g = gy * _dact(y, act, slope)
    gx = gamma.view(1, c, 1, 1) * rstd * (g - g.mean(dim=(0, 2, 3), keepdim=True) - xh * (g * xh).mean(dim=(0, 2, 3), keepdim=True))
    return _cl(gx), (g * xh).sum(dim=(0, 2, 3)), g.sum(dim=(0, 2, 3))


def tanh_fwd(x):
    return torch.tanh(x)


def tanh_bwd(gy, y):
    return gy * (1 - y * y)


# ---- grouped small linears (csrc/linear.cu glinear_*): the table object carries python-side copies of what the kernels read
def _gl_layers(tab):
    return tab._layers_for_contract


def glinear_fwd(ws, tab):
    outs = []
    for l, (w, b, a, bs) in enumerate(_gl_layers(tab)):
        y = a * ws[l] @ w.t()
        if b is not None:
            y = y + bs * b
        outs.append(y.reshape(-1))
    return torch.cat(outs)


def glinear_dgrad(g_all, tab, L, M, Kf):
    out = []
    for l, (w, b, a, bs) in enumerate(_gl_layers(tab)):
        g = g_all[:, tab.offs[l]:tab.offs[l] + tab.nouts[l]]
        out.append(a * g @ w)
    return torch.stack(out)


def glinear_wgrad(ws, g_all, tab):
    gw, gb = [], []
    for l, (w, b, a, bs) in enumerate(_gl_layers(tab)):
        g = g_all[:, tab.offs[l]:tab.offs[l] + tab.nouts[l]]
        gw.append(a * g.t() @ ws[l])
        gb.append(bs * g.sum(0))
    return torch.cat(gw), torch.cat(gb)


# ---- real-image input pipeline ------------------------------------------------------------------------------------------
def u8_box_resize_normalize(src, index, out_hw, mean, std, flip=None):
    from oracle import pil_box
    out = pil_box.input_pipeline(src.cpu().numpy(), None if index is None else index.cpu().numpy(), out_hw, mean, std,
                                 None if flip is None else flip.cpu().numpy())
    return out.to(src.device)


ALL = [n for n, f in list(globals().items()) if callable(f) and not n.startswith("_") and f.__module__ == __name__
       and n not in ("install",)]


def install(monkeypatch):
    """Swap the CUDA launchers of gan_lab_b200._kernels for these CPU contracts (tests only)."""
    import gan_lab_b200._kernels as K
    for name in ALL:
        if hasattr(K, name):
            monkeypatch.setattr(K, name, globals()[name])
    # the grouped-linear table normally holds device pointers; the CPU double keeps the layer tensors themselves
    def update(self, layers, device):
        self._layers_for_contract = [(w.detach(), None if b is None else b.detach(), a, bs) for w, b, a, bs in layers]
        self.nouts = [w.shape[0] for w, _b, _a, _bs in layers]
        self.offs = [sum(self.nouts[:i]) for i in range(len(self.nouts))]
        self.G = sum(self.nouts)
        return self
    monkeypatch.setattr(K.GroupedLinearTable, "update", update)
