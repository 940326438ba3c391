"""Generate tests/golden/*.pt by running the UNMODIFIED reference (CPU, fp32).  TEST INFRASTRUCTURE ONLY.

Run in the build container (needs /root/reference):   python -m oracle.make_golden
The fixtures travel to the GPU box; the reference does not.

Every fixture is a dict of tensors / python scalars saved with torch.save.  Random draws that the
reference makes inside forward()/train() are captured on a tape (by wrapping torch.randn / torch.rand /
torch.randint / np.random.rand while the reference runs -- the reference's code itself is untouched)
so that the oracle and the B200 path can replay exactly the same latents, noise and mixing cut-offs.
"""
from __future__ import annotations

import contextlib
import io
import sys
import zlib
from pathlib import Path

import numpy as np
import torch
from torch.utils.data import BatchSampler, DataLoader, SequentialSampler, TensorDataset

from oracle.reference_loader import load_reference, make_config

GOLDEN_DIR = Path(__file__).resolve().parent.parent / "tests" / "golden"
SMALL_FMAP_MAX = 32      # reference constant stylegan/base.py:17 / progan/base.py:17 patched for small fixtures


# --------------------------------------------------------------------------- #
class Tape:
    """Records the reference's random draws in call order."""

    def __init__(self):
        self.events = []
        self._orig = {}

    def __enter__(self):
        self._orig = dict(randn=torch.randn, rand=torch.rand, randint=torch.randint, nprand=np.random.rand)
        tape = self

        def randn(*a, **k):
            t = tape._orig["randn"](*a, **k)
            tape.events.append(("randn", t.detach().clone()))
            return t

        def rand(*a, **k):
            t = tape._orig["rand"](*a, **k)
            tape.events.append(("rand", t.detach().clone()))
            return t

        def randint(*a, **k):
            t = tape._orig["randint"](*a, **k)
            tape.events.append(("randint", t.detach().clone()))
            return t

        def nprand(*a):
            v = tape._orig["nprand"](*a)
            tape.events.append(("np_rand", float(v)))
            return v

        torch.randn, torch.rand, torch.randint, np.random.rand = randn, rand, randint, nprand
        return self

    def __exit__(self, *exc):
        torch.randn, torch.rand, torch.randint = self._orig["randn"], self._orig["rand"], self._orig["randint"]
        np.random.rand = self._orig["nprand"]
        return False


class ReplayTape:
    """The inverse of Tape: while active, torch.randn / rand / randint / np.random.rand hand out recorded draws in order
    (moved to `dev`), so that a second run of the reference -- on another device -- sees exactly the first run's random numbers."""

    def __init__(self, events, dev):
        self.events, self.dev, self._orig = list(events), dev, {}

    def _pop(self, kind):
        k, v = self.events.pop(0)
        assert k == kind, (k, kind)
        return v

    def __enter__(self):
        self._orig = dict(randn=torch.randn, rand=torch.rand, randint=torch.randint, nprand=np.random.rand)
        tape = self

        def tensor_draw(kind):
            def f(*a, **k):
                t = tape._pop(kind)
                want = k.get("device", None)
                return t.to(want) if want is not None else t.clone()
            return f

        torch.randn, torch.rand, torch.randint = tensor_draw("randn"), tensor_draw("rand"), tensor_draw("randint")
        np.random.rand = lambda *a: tape._pop("np_rand")
        return self

    def __exit__(self, *exc):
        torch.randn, torch.rand, torch.randint = self._orig["randn"], self._orig["rand"], self._orig["randint"]
        np.random.rand = self._orig["nprand"]
        return False


def perturb_zero_params(module, gen):
    """Zero-initialised params (biases, noise_weight) hide whole code paths; perturb them
    (SURVEY.md section 8c 'Determinism hooks')."""
    with torch.no_grad():
        for name, p in module.named_parameters():
            if float(p.abs().max()) == 0.0:
                p.copy_(torch.randn(p.shape, generator=gen) * 0.3)


def sd_clone(module):
    return {k: v.detach().clone() for k, v in module.state_dict().items()}


def grads_of(module):
    return {k: (p.grad.detach().clone() if p.grad is not None else None) for k, p in module.named_parameters()}


# --------------------------------------------------------------------------- #
def golden_layers(ref):
    """Per-module forward / backward (/ double-backward for D-side modules)."""
    cl = ref.custom_layers
    g = torch.Generator().manual_seed(1234)
    out = {}

    def rn(*shape):
        return torch.randn(*shape, generator=g)

    # Conv2dEx variants (custom_layers.py:147-211)
    conv_cases = {
        "conv3x3": dict(ni=32, nf=64, ks=3, padding=1, gain_sq_base=2., lrmul=1., include_bias=True, hw=8, n=4),
        "conv3x3_nobias": dict(ni=64, nf=32, ks=3, padding=1, gain_sq_base=2., lrmul=1., include_bias=False, hw=16, n=2),
        "conv1x1_torgb": dict(ni=32, nf=3, ks=1, padding=0, gain_sq_base=1., lrmul=1., include_bias=True, hw=16, n=4),
        "conv1x1_fromrgb": dict(ni=3, nf=32, ks=1, padding=0, gain_sq_base=2., lrmul=1., include_bias=True, hw=16, n=4),
        "conv4x4_valid": dict(ni=32, nf=32, ks=4, padding=0, gain_sq_base=2., lrmul=1., include_bias=True, hw=4, n=8),
        "conv3x3_c33": dict(ni=33, nf=32, ks=3, padding=1, gain_sq_base=2., lrmul=1., include_bias=True, hw=4, n=8),
        "conv3x3_lrmul": dict(ni=16, nf=16, ks=3, padding=1, gain_sq_base=2., lrmul=.5, include_bias=True, hw=8, n=2),
    }
    for name, c in conv_cases.items():
        torch.manual_seed(zlib.crc32(name.encode()) % 1000)      # (not hash(): str hashes are salted per process)
        m = cl.Conv2dEx(ni=c["ni"], nf=c["nf"], ks=c["ks"], stride=1, padding=c["padding"], init="He",
                        init_type="StyleGAN", gain_sq_base=c["gain_sq_base"], equalized_lr=True,
                        lrmul=c["lrmul"], include_bias=c["include_bias"])
        perturb_zero_params(m, g)
        x = rn(c["n"], c["ni"], c["hw"], c["hw"]).requires_grad_(True)
        y = m(x)
        gy = rn(*y.shape)
        gx, = torch.autograd.grad(y, x, gy, create_graph=True)
        # double backward: d/dtheta of sum(gx * v) (the R1 pattern), plus first-order param grads
        v = rn(*gx.shape)
        (gx * v).sum().backward(retain_graph=True)
        gg_w = m.conv2d.weight.grad.clone()
        m.zero_grad()
        y.backward(gy)
        out[name] = dict(cfg=c, sd=sd_clone(m), wscale=float(m.wscale), x=x.detach(), y=y.detach(), gy=gy,
                         gx=gx.detach(), v=v, gg_w=gg_w, grads=grads_of(m))

    # LinearEx (custom_layers.py:230-291)
    lin_cases = {
        "linear_mapping": dict(nin=32, nout=32, gain_sq_base=2., lrmul=.01, n=8),
        "linear_style": dict(nin=32, nout=128, gain_sq_base=1., lrmul=1., n=8),
        "linear_dhead": dict(nin=32, nout=1, gain_sq_base=1., lrmul=1., n=8),
        "linear_progan_fc": dict(nin=32, nout=512, gain_sq_base=2. / 16, lrmul=1., n=4),
    }
    for name, c in lin_cases.items():
        torch.manual_seed(zlib.crc32(name.encode()) % 1000)      # (not hash(): str hashes are salted per process)
        m = cl.LinearEx(nin_feat=c["nin"], nout_feat=c["nout"], init="He", init_type="StyleGAN",
                        gain_sq_base=c["gain_sq_base"], equalized_lr=True, lrmul=c["lrmul"])
        perturb_zero_params(m, g)
        x = rn(c["n"], c["nin"]).requires_grad_(True)
        y = m(x)
        gy = rn(*y.shape)
        gx, = torch.autograd.grad(y, x, gy, create_graph=True)
        v = rn(*gx.shape)
        (gx * v).sum().backward(retain_graph=True)
        gg_w = m.linear.weight.grad.clone()
        m.zero_grad()
        y.backward(gy)
        out[name] = dict(cfg=c, sd=sd_clone(m), wscale=float(m.wscale), x=x.detach(), y=y.detach(), gy=gy,
                         gx=gx.detach(), v=v, gg_w=gg_w, grads=grads_of(m))

    # Conv2dBias (custom_layers.py:213-226)
    m = cl.Conv2dBias(nf=16, lrmul=1., device="cpu")
    perturb_zero_params(m, g)
    x = rn(2, 16, 4, 4).requires_grad_(True)
    y = m(x); gy = rn(*y.shape); y.backward(gy)
    out["conv2dbias"] = dict(sd=sd_clone(m), x=x.detach(), y=y.detach(), gy=gy, gx=x.grad.clone(), grads=grads_of(m))

    # PixelNorm2d (custom_layers.py:81-86), 2-D (latents) and 4-D (ProGAN features)
    pn = cl.PixelNorm2d()
    for name, shape in (("pixelnorm_z", (8, 32)), ("pixelnorm_feat", (4, 32, 8, 8))):
        x = rn(*shape).requires_grad_(True)
        y = pn(x); gy = rn(*y.shape); y.backward(gy)
        out[name] = dict(x=x.detach(), y=y.detach(), gy=gy, gx=x.grad.clone())

    # blur (custom_layers.py:36-53): fwd, bwd, and double-bwd (linear -> gg = blur(v))
    for name, shape in (("blur", (2, 16, 8, 8)), ("blur_odd", (1, 5, 6, 4))):
        op = cl.get_blur_op("binomial", shape[1])
        x = rn(*shape).requires_grad_(True)
        y = op(x); gy = rn(*y.shape).requires_grad_(True)
        gx, = torch.autograd.grad(y, x, gy, create_graph=True)
        v = rn(*gx.shape)
        ggy, = torch.autograd.grad(gx, gy, v)
        out[name] = dict(x=x.detach(), y=y.detach(), gy=gy.detach(), gx=gx.detach(), v=v, ggy=ggy)

    # minibatch stddev (custom_layers.py:117-140): fwd, bwd, double-bwd wrt x and gy
    for name, (shape, gs) in dict(mbstd_n8=((8, 16, 4, 4), 4), mbstd_n4=((4, 32, 4, 4), 4),
                                  mbstd_n6=((6, 8, 4, 4), 4), mbstd_n1=((1, 8, 4, 4), 4),
                                  mbstd_n16=((16, 16, 4, 4), 4)).items():
        x = rn(*shape).requires_grad_(True)
        y = cl.concat_mbstd_layer(x, gs)
        gy = rn(*y.shape).requires_grad_(True)
        if shape[0] > 1:
            gx, = torch.autograd.grad(y, x, gy, create_graph=True)
            v = rn(*gx.shape)
            ggx, ggy = torch.autograd.grad(gx, (x, gy), v, allow_unused=True)
        else:
            gx, = torch.autograd.grad(y, x, gy)
            v = rn(*gx.shape); ggx = None; ggy = None
        out[name] = dict(group_size=gs, x=x.detach(), y=y.detach(), gy=gy.detach(), gx=gx.detach(), v=v,
                         ggx=ggx, ggy=ggy)

    # InstanceNorm (custom_layers.py:98-99) + the G-layer tail Sequential(bias, lrelu, IN) and AdaIN
    sa = ref.stylegan_arch
    nf = 32
    noise_mod = sa.StyleAddNoise(nf)
    bias_mod = cl.Conv2dBias(nf, device="cpu")
    tail = torch.nn.Sequential(bias_mod, torch.nn.LeakyReLU(.2), cl.NormalizeLayer("InstanceNorm", ni=nf))
    perturb_zero_params(noise_mod, g); perturb_zero_params(bias_mod, g)
    noise_mod.eval()                                     # honour the supplied noise (architectures.py:113)
    x = rn(4, nf, 8, 8).requires_grad_(True)
    noise = rn(4, 1, 8, 8)
    style = rn(4, 2 * nf).requires_grad_(True)
    t = tail(noise_mod(x, noise=noise))
    ysyb = style.view(-1, 2, nf, 1, 1)
    y = t * (ysyb[:, 0].contiguous().add(1)) + ysyb[:, 1].contiguous()
    gy = rn(*y.shape)
    y.backward(gy)
    out["style_epilogue"] = dict(x=x.detach(), noise=noise, style=style.detach(), y=y.detach(), gy=gy,
                                 noise_weight=noise_mod.noise_weight.detach().clone(),
                                 bias=bias_mod.bias.detach().clone(), gx=x.grad.clone(),
                                 gstyle=style.grad.clone(), g_noise_weight=noise_mod.noise_weight.grad.clone(),
                                 g_bias=bias_mod.bias.grad.clone())

    # D downsample tail: AvgPool2d(2,2) -> Conv2dBias -> LeakyReLU (progan/architectures.py:267-284) with 2nd order
    bias_mod = cl.Conv2dBias(16, device="cpu"); perturb_zero_params(bias_mod, g)
    x = rn(2, 16, 8, 8).requires_grad_(True)
    y = torch.nn.functional.leaky_relu(bias_mod(torch.nn.functional.avg_pool2d(x, 2, 2)), .2)
    gy = rn(*y.shape).requires_grad_(True)
    gx, = torch.autograd.grad(y, x, gy, create_graph=True)
    v = rn(*gx.shape)
    ggy, = torch.autograd.grad(gx, gy, v)
    out["pool_bias_lrelu"] = dict(x=x.detach(), bias=bias_mod.bias.detach().clone(), y=y.detach(),
                                  gy=gy.detach(), gx=gx.detach(), v=v, ggy=ggy)

    # nearest x2 upsample / avg-pool (resnetgan/learner.py:154-164)
    x = rn(2, 8, 4, 4).requires_grad_(True)
    y = torch.nn.Upsample(scale_factor=2, mode="nearest")(x); gy = rn(*y.shape); y.backward(gy)
    out["upsample2x"] = dict(x=x.detach(), y=y.detach(), gy=gy, gx=x.grad.clone())
    return out


# --------------------------------------------------------------------------- #
def _patch_small(ref, fmap_max=SMALL_FMAP_MAX, fmap_base=8192):
    """fmap(stage) = min(FMAP_BASE / 2^stage, FMAP_MAX) (stylegan/base.py:16-17, 76-77).  With the default base every stage of a
    small fixture sits at FMAP_MAX; TAPER_FMAP_BASE makes the channel count halve from 16x16 on, like 512 -> 256 -> 128 ... in the
    full-size networks (blocks with ni != nf, prev_torgb / torgb of different widths)."""
    for mod in (ref.stylegan_base, ref.progan_base):
        mod.FMAP_MAX = fmap_max
        mod.FMAP_BASE = fmap_base


TAPER_FMAP_BASE = 128      # 32, 32, 16, 8 feature maps at 4, 8, 16, 32 pixels


def _unpatch(ref):
    for mod in (ref.stylegan_base, ref.progan_base):
        mod.FMAP_MAX = 512
        mod.FMAP_BASE = 8192


def _quiet():
    return contextlib.redirect_stdout(io.StringIO())


def _build_style_learner(ref, res, init_res, bs, **over):
    over.setdefault("cutoff_trunc_trick", int(np.log2(res)) - 2)
    cfg = make_config("StyleGAN", res=res, init_res=init_res, batch_size=bs, len_latent=SMALL_FMAP_MAX,
                      len_dlatent=SMALL_FMAP_MAX, **over)
    with _quiet():
        L = ref.stylegan_learner.StyleGANLearner(cfg)
    return L, cfg


def golden_style_nets(ref, res=16, bs=4, fade_in=False, fmap_base=8192):
    """StyleGenerator / StyleDiscriminator forward+backward(+R1) on the reference modules."""
    _patch_small(ref, fmap_base=fmap_base)
    try:
        torch.manual_seed(7 + int(fade_in)); np.random.seed(7)
        L, cfg = _build_style_learner(ref, res, res // 2 if fade_in else res, bs)
        G, D = L.gen_model, L.disc_model
        alpha = 1.0
        if fade_in:
            G.increase_scale(); D.increase_scale()
            G.alpha = 0.3
            alpha = 0.3
        gen = torch.Generator().manual_seed(99)
        perturb_zero_params(G, gen); perturb_zero_params(D, gen)
        G.train(); D.train()
        z = torch.randn(bs, cfg.len_latent, generator=gen)
        with Tape() as tape:
            img = G(z)
        w_ewma = G.w_ewma.detach().clone()
        gimg = torch.randn(img.shape, generator=gen)
        G.zero_grad(); img.backward(gimg)
        g_grads = grads_of(G)

        x = torch.rand(bs, 3, res, res, generator=gen) * 2 - 1
        D.zero_grad()
        logits = D(x)
        glog = torch.randn(logits.shape, generator=gen)
        logits.backward(glog)
        d_grads = grads_of(D)
        # R1 on its own (invisible inside the whole loss at random init)
        D.zero_grad()
        L.batch_size = bs
        gp = L.calc_gp(img.detach(), x)
        gp.backward()
        d_gp_grads = grads_of(D)
        # grad of D output wrt image (the G-step path)
        xi = img.detach().clone().requires_grad_(True)
        gxi, = torch.autograd.grad(D(xi), xi, glog)
        return dict(res=res, bs=bs, fade_in=fade_in, alpha=alpha, fmap_max=SMALL_FMAP_MAX, fmap_base=fmap_base,
                    len_latent=cfg.len_latent, g_sd=sd_clone(G), d_sd=sd_clone(D), z=z, tape=tape.events,
                    img=img.detach(), w_ewma=w_ewma, gimg=gimg, g_grads=g_grads, x=x, logits=logits.detach(),
                    glog=glog, d_grads=d_grads, gp=gp.detach(), d_gp_grads=d_gp_grads, d_gx_img=gxi,
                    lda=cfg.lda)
    finally:
        _unpatch(ref)


def golden_style_eval(ref, res=32, bs=3):
    """Evaluation-mode StyleGenerator: supplied per-layer noise, truncation trick on w
    (psi, cut-off stage, de-truncation at the cut-off), style mixing at a given stage -- reference
    stylegan/architectures.py:438-440, 505-522."""
    _patch_small(ref)
    try:
        torch.manual_seed(61); np.random.seed(61)
        L, cfg = _build_style_learner(ref, res, res, bs, cutoff_trunc_trick=2, psi_trunc_trick=.6)
        G = L.gen_model
        gen = torch.Generator().manual_seed(63)
        perturb_zero_params(G, gen)
        G.train()
        with torch.no_grad():
            G(torch.randn(bs, cfg.len_latent, generator=gen))          # one training forward creates w_ewma
        w_ewma = G.w_ewma.detach().clone()
        G.eval()
        n_layers = len(G.gen_layers)
        hw = [4 * 2 ** (n // 2) for n in range(n_layers)]
        noise = [torch.randn(bs, 1, h, h, generator=gen) for h in hw]
        z = torch.randn(bs, cfg.len_latent, generator=gen)
        z_mix = torch.randn(bs, cfg.len_latent, generator=gen)
        out = {}
        with torch.no_grad():
            out["img_trunc"] = G(z, noise=noise).clone()
            out["img_mix_low"] = G(z, x_mixing=z_mix, style_mixing_stage=2, noise=noise).clone()     # below the cut-off (2*2)
            out["img_mix_high"] = G(z, x_mixing=z_mix, style_mixing_stage=5, noise=noise).clone()    # above it
            G.trunc_cutoff_stage = None
            out["img_notrunc"] = G(z, noise=noise).clone()
        return dict(res=res, bs=bs, fmap_max=SMALL_FMAP_MAX, len_latent=cfg.len_latent, g_sd=sd_clone(G), w_ewma=w_ewma,
                    z=z, z_mix=z_mix, noise=noise, psi=.6, cutoff=2, **out)
    finally:
        _unpatch(ref)


def golden_pro_nets(ref, res=16, bs=4, fade_in=False, fmap_base=8192):
    _patch_small(ref, fmap_base=fmap_base)
    try:
        torch.manual_seed(11 + int(fade_in)); np.random.seed(11)
        cfg = make_config("ProGAN", res=res, init_res=res // 2 if fade_in else res, batch_size=bs,
                          len_latent=SMALL_FMAP_MAX)
        with _quiet():
            L = ref.progan_learner.ProGANLearner(cfg)
        G, D = L.gen_model, L.disc_model
        alpha = 1.0
        if fade_in:
            G.increase_scale(); D.increase_scale(); G.alpha = 0.6; alpha = 0.6
        gen = torch.Generator().manual_seed(5)
        perturb_zero_params(G, gen); perturb_zero_params(D, gen)
        G.train(); D.train()
        z = torch.randn(bs, cfg.len_latent, generator=gen)
        img = G(z)
        gimg = torch.randn(img.shape, generator=gen)
        G.zero_grad(); img.backward(gimg)
        g_grads = grads_of(G)
        x = torch.rand(bs, 3, res, res, generator=gen) * 2 - 1
        D.zero_grad()
        logits = D(x); glog = torch.randn(logits.shape, generator=gen); logits.backward(glog)
        d_grads = grads_of(D)
        D.zero_grad()
        L.batch_size = bs
        with Tape() as tape:
            gp = L.calc_gp(img.detach(), x)          # wgan-gp: draws eps via torch.rand
        gp.backward()
        d_gp_grads = grads_of(D)
        return dict(res=res, bs=bs, fade_in=fade_in, alpha=alpha, fmap_max=SMALL_FMAP_MAX, fmap_base=fmap_base,
                    len_latent=cfg.len_latent, g_sd=sd_clone(G), d_sd=sd_clone(D), z=z, img=img.detach(),
                    gimg=gimg, g_grads=g_grads, x=x, logits=logits.detach(), glog=glog, d_grads=d_grads,
                    gp=gp.detach(), gp_tape=tape.events, d_gp_grads=d_gp_grads, lda=cfg.lda, gamma=cfg.gamma)
    finally:
        _unpatch(ref)


def golden_train_steps(ref, model="StyleGAN", res=16, bs=4, iters=2):
    """Whole Learner.train() for `iters` main iterations (D step + G step + Adam + EWMA), unmodified loop."""
    _patch_small(ref)
    try:
        torch.manual_seed(21); np.random.seed(21)
        if model == "StyleGAN":
            L, cfg = _build_style_learner(ref, res, res, bs)
        else:
            cfg = make_config("ProGAN", res=res, init_res=res, batch_size=bs, len_latent=SMALL_FMAP_MAX)
            with _quiet():
                L = ref.progan_learner.ProGANLearner(cfg)
        gen = torch.Generator().manual_seed(31)
        perturb_zero_params(L.gen_model, gen); perturb_zero_params(L.disc_model, gen)
        g0, d0 = sd_clone(L.gen_model), sd_clone(L.disc_model)
        data = torch.rand(iters * bs, 3, res, res, generator=gen) * 2 - 1
        ds = TensorDataset(data)
        dl = DataLoader(ds, batch_sampler=BatchSampler(SequentialSampler(ds), batch_size=bs, drop_last=True))
        losses = []
        orig_backward = torch.Tensor.backward

        def rec_backward(self, *a, **k):
            losses.append(float(self.detach()))
            return orig_backward(self, *a, **k)

        torch.Tensor.backward = rec_backward
        try:
            with Tape() as tape, _quiet(), contextlib.redirect_stderr(io.StringIO()):
                L.train(dl, num_main_iters=iters)
        finally:
            torch.Tensor.backward = orig_backward
        lagged = {k: v.detach().clone() for k, v in L.lagged_params.items()}
        out = dict(model=model, res=res, bs=bs, iters=iters, fmap_max=SMALL_FMAP_MAX, len_latent=cfg.len_latent,
                   g_sd0=g0, d_sd0=d0, data=data, tape=tape.events, losses=losses,
                   g_sd1=sd_clone(L.gen_model), d_sd1=sd_clone(L.disc_model), lagged=lagged, beta=float(L.beta),
                   lr=cfg.lr_base * cfg.lr_fctr_dict[res])
        if model == "StyleGAN":
            out["w_ewma"] = L.gen_model.w_ewma.detach().clone()
        return out
    finally:
        _unpatch(ref)


class PILBoxDataset(torch.utils.data.Dataset):
    """Synthetic stand-in for the torchvision dataset the reference's train() drives while it grows: uint8 images behind
    the reference's own transform chain Resize(BOX) -> ToTensor -> Normalize(.5, .5) (data_config.py:312-342); train()
    rewrites `dataset.transforms.transform.transforms` at every resolution increase (progan/learner.py:611-612).  Every
    sample served is recorded, so the B200 path can be fed exactly the same real images."""

    def __init__(self, images_u8, res):
        from PIL import Image
        from torchvision import transforms as T
        from torchvision.datasets.vision import StandardTransform
        self.images = [Image.fromarray(a) for a in images_u8]          # (H, W, 3) uint8
        self.transforms = StandardTransform(T.Compose([T.Resize((res, res), interpolation=Image.BOX), T.ToTensor(),
                                                       T.Normalize(mean=[.5] * 3, std=[.5] * 3)]))
        self.served = []

    def __len__(self):
        return len(self.images)

    def __getitem__(self, i):
        x = self.transforms.transform(self.images[i])
        self.served.append((int(i), x.clone()))
        return (x,)


def golden_grow(ref, model="StyleGAN", init_res=4, res=8, bs_dict=None, nimg_transition=16, iters=11, data_res=8,
                snap_iters=tuple(range(1, 11)), fmap_base=8192):
    """Learner.train() THROUGH a resolution increase: stabilise at `init_res`, grow, fade the new
    block in with a moving alpha (incl. the real-image blend, progan/learner.py:770-779), stabilise at `res`, enter the
    final phase.  Recorded besides the usual train fixture: the parameters right after every increase_scale() (the
    freshly initialised block does not come from taped draws), every real sample served, and a per-iteration trace of
    (curr_res, fade_in_phase, alpha, batch size, learning rates, phase number)."""
    _patch_small(ref, fmap_base=fmap_base)
    try:
        torch.manual_seed(77); np.random.seed(77)
        bs_dict = bs_dict or {r: 4 for r in (4, 8, 16, 32)}     # a batch-size change mid-iterator breaks the reference itself under torch 2.11 (BatchSampler caches it)
        over = dict(bs_dict={**{r: 4 for r in (4, 8, 16, 32, 64, 128, 256, 512, 1024)}, **bs_dict},
                    nimg_transition=nimg_transition, res_dataset=data_res,
                    lr_fctr_dict={4: 1., 8: 1.5, 16: 2., 32: 1., 64: 1., 128: 1., 256: 1., 512: 1., 1024: 1.})
        if model == "StyleGAN":
            L, cfg = _build_style_learner(ref, res, init_res, bs_dict[init_res], **over)
        else:
            cfg = make_config("ProGAN", res=res, init_res=init_res, batch_size=bs_dict[init_res], len_latent=SMALL_FMAP_MAX,
                              **over)
            with _quiet():
                L = ref.progan_learner.ProGANLearner(cfg)
        gen = torch.Generator().manual_seed(78)
        perturb_zero_params(L.gen_model, gen); perturb_zero_params(L.disc_model, gen)
        g0, d0 = sd_clone(L.gen_model), sd_clone(L.disc_model)
        images = torch.randint(0, 256, (24, data_res, data_res, 3), generator=gen, dtype=torch.uint8).numpy()
        ds = PILBoxDataset(images, init_res)
        dl = DataLoader(ds, batch_sampler=BatchSampler(SequentialSampler(ds), batch_size=L.batch_size, drop_last=True))

        after_inc = []
        for net, tag in ((L.gen_model, "g"), (L.disc_model, "d")):
            orig = net.increase_scale

            def wrapped(orig=orig, net=net, tag=tag):
                orig()
                perturb_zero_params(net, gen)          # the new block's zero-initialised biases / noise weights
                after_inc.append((tag, sd_clone(net)))
            net.increase_scale = wrapped

        losses, trace, iter_snaps = [], [], {}
        orig_backward = torch.Tensor.backward

        def rec_backward(self, *a, **k):
            losses.append(float(self.detach()))
            if len(losses) % 2 == 1:                     # D-step backward: the iteration's state is settled here
                if len(trace) in snap_iters:
                    iter_snaps[len(trace)] = (sd_clone(L.gen_model), sd_clone(L.disc_model))
                trace.append(dict(res=int(L.gen_model.curr_res), fade=bool(L.gen_model.fade_in_phase),
                                  alpha=float(L.gen_model.alpha), bs=int(L.batch_size), phase=int(L.curr_phase_num),
                                  lr_d=float(L.opt_disc.param_groups[0]["lr"]), lr_g=float(L.opt_gen.param_groups[0]["lr"]),
                                  beta=float(L.beta), img_num=int(L.curr_img_num)))
            return orig_backward(self, *a, **k)

        torch.Tensor.backward = rec_backward
        try:
            with Tape() as tape, _quiet(), contextlib.redirect_stderr(io.StringIO()):
                L.train(dl, num_main_iters=iters)
        finally:
            torch.Tensor.backward = orig_backward
        lagged = {k: v.detach().clone() for k, v in L.lagged_params.items()}
        out = dict(model=model, init_res=init_res, res=res, iters=iters, fmap_max=SMALL_FMAP_MAX, fmap_base=fmap_base,
                   len_latent=cfg.len_latent, bs_dict=dict(cfg.bs_dict), nimg_transition=nimg_transition, lr_fctr_dict=dict(cfg.lr_fctr_dict),
                   lr_base=cfg.lr_base, data_res=data_res, images_u8=torch.from_numpy(images),
                   g_sd0=g0, d_sd0=d0, after_inc=after_inc, served=ds.served, tape=tape.events, losses=losses, trace=trace,
                   iter_snaps=iter_snaps,      # parameters at the START of those iterations (behind that many updates)
                   g_sd1=sd_clone(L.gen_model), d_sd1=sd_clone(L.disc_model), lagged=lagged, beta=float(L.beta),
                   final=dict(res=int(L.gen_model.curr_res), fade=bool(L.gen_model.fade_in_phase),
                              alpha=float(L.gen_model.alpha), phase=int(L.curr_phase_num), img_num=int(L.curr_img_num),
                              nimg_transition_lst=[float(v) for v in L.nimg_transition_lst],
                              progressively_grow=bool(L.progressively_grow), bs=int(L.batch_size)))
        if model == "StyleGAN":
            out["w_ewma"] = L.gen_model.w_ewma.detach().clone()
        return out
    finally:
        _unpatch(ref)


def golden_resume(ref, model="StyleGAN", iters_before=6, iters_after=3, ckpt_name=None):
    """The reference saves a checkpoint in the MIDDLE of a fade-in (save_model), a fresh reference learner loads it
    (load_model) and trains on.  The checkpoint file itself -- a real reference `.tar` -- is committed
    next to the fixture; the fixture holds what the resumed reference run consumed and produced."""
    _patch_small(ref)
    try:
        torch.manual_seed(91); np.random.seed(91)
        over = dict(bs_dict={r: 4 for r in (4, 8, 16, 32, 64, 128, 256, 512, 1024)}, nimg_transition=16, res_dataset=8,
                    lr_fctr_dict={4: 1., 8: 1.5, 16: 2., 32: 1., 64: 1., 128: 1., 256: 1., 512: 1., 1024: 1.})
        learner_mod = ref.stylegan_learner if model == "StyleGAN" else ref.progan_learner
        learner_cls = learner_mod.StyleGANLearner if model == "StyleGAN" else learner_mod.ProGANLearner

        def build():
            if model == "StyleGAN":
                return _build_style_learner(ref, 8, 4, 4, **over)
            cfg = make_config("ProGAN", res=8, init_res=4, batch_size=4, len_latent=SMALL_FMAP_MAX, **over)
            with _quiet():
                return learner_cls(cfg), cfg

        L, cfg = build()
        gen = torch.Generator().manual_seed(92)
        perturb_zero_params(L.gen_model, gen); perturb_zero_params(L.disc_model, gen)
        images = torch.randint(0, 256, (24, 8, 8, 3), generator=gen, dtype=torch.uint8).numpy()

        def loader(learner):
            ds = PILBoxDataset(images, 4)
            return ds, DataLoader(ds, batch_sampler=BatchSampler(SequentialSampler(ds), batch_size=learner.batch_size,
                                                                 drop_last=True))

        for net in (L.gen_model, L.disc_model):
            orig = net.increase_scale
            net.increase_scale = lambda orig=orig, net=net: (orig(), perturb_zero_params(net, gen))
        ds, dl = loader(L)
        with _quiet(), contextlib.redirect_stderr(io.StringIO()):
            L.train(dl, num_main_iters=iters_before)
        # save_model() reads valid_z, which only exists once an image grid has been computed (resnetgan/learner.py:388-389)
        L.valid_z = torch.zeros(cfg.img_grid_sz ** 2, cfg.len_latent)
        ckpt = GOLDEN_DIR / (ckpt_name or f"{model.lower()}_reference_checkpoint.tar")
        L.save_model(ckpt)
        saved = dict(alpha=float(L.gen_model.alpha), res=int(L.gen_model.curr_res), fade=bool(L.gen_model.fade_in_phase),
                     img_num=int(L.curr_img_num), phase=int(L.curr_phase_num), g_sd=sd_clone(L.gen_model),
                     d_sd=sd_clone(L.disc_model), lagged={k: v.detach().clone() for k, v in L.lagged_params.items()})

        # ---- a fresh reference learner resumes from the file
        # torch >= 2.6 defaults torch.load to weights_only=True, under which the unmodified reference cannot read its own
        # checkpoints (they pickle the config object, nn.Modules and an IndexedOrderedDict); the harness restores the old default
        L2, _ = build()
        orig_load = torch.load
        torch.load = lambda *a, **k: orig_load(*a, **{"weights_only": False, **k})
        try:
            with _quiet():
                L2.load_model(ckpt, dev_of_saved_model="cpu")
        finally:
            torch.load = orig_load
        ds2, dl2 = loader(L2)
        losses, trace, iter_snaps = [], [], {}
        orig_backward = torch.Tensor.backward

        def rec_backward(self, *a, **k):
            if len(losses) and len(losses) % 2 == 0:           # first backward of a later iteration: parameters at its start
                iter_snaps[len(losses) // 2] = (sd_clone(L2.gen_model), sd_clone(L2.disc_model))
            losses.append(float(self.detach()))
            if len(losses) % 2 == 1:
                trace.append(dict(res=int(L2.gen_model.curr_res), fade=bool(L2.gen_model.fade_in_phase),
                                  alpha=float(L2.gen_model.alpha), bs=int(L2.batch_size), phase=int(L2.curr_phase_num),
                                  lr_d=float(L2.opt_disc.param_groups[0]["lr"]), lr_g=float(L2.opt_gen.param_groups[0]["lr"]),
                                  beta=float(L2.beta), img_num=int(L2.curr_img_num)))
            return orig_backward(self, *a, **k)

        torch.Tensor.backward = rec_backward
        try:
            with Tape() as tape, _quiet(), contextlib.redirect_stderr(io.StringIO()):
                L2.train(dl2, num_main_iters=iters_after)
        finally:
            torch.Tensor.backward = orig_backward
        out = dict(model=model, checkpoint=ckpt.name, init_res=4, res=8, iters_before=iters_before, iters_after=iters_after,
                   fmap_max=SMALL_FMAP_MAX, len_latent=cfg.len_latent, bs_dict=dict(cfg.bs_dict), nimg_transition=16,
                   lr_fctr_dict=dict(cfg.lr_fctr_dict), lr_base=cfg.lr_base, data_res=8, saved=saved, served=ds2.served,
                   iter_snaps=iter_snaps,
                   tape=tape.events, losses=losses, trace=trace, g_sd1=sd_clone(L2.gen_model), d_sd1=sd_clone(L2.disc_model),
                   lagged={k: v.detach().clone() for k, v in L2.lagged_params.items()}, beta=float(L2.beta),
                   opt_gen_sd=L2.opt_gen.state_dict(), opt_disc_sd=L2.opt_disc.state_dict(),
                   final=dict(res=int(L2.gen_model.curr_res), fade=bool(L2.gen_model.fade_in_phase),
                              alpha=float(L2.gen_model.alpha), phase=int(L2.curr_phase_num), img_num=int(L2.curr_img_num),
                              nimg_transition_lst=[float(v) for v in L2.nimg_transition_lst], bs=int(L2.batch_size)))
        if model == "StyleGAN":
            out["w_ewma"] = L2.gen_model.w_ewma.detach().clone()
        return out
    finally:
        _unpatch(ref)


def golden_metrics(ref, model="StyleGAN", res=16, bs=4, n_valid=10):
    """compute_metrics() of the reference (progan/learner.py:248-416) after one training iteration:
    generator metrics on a latent validation set whose last batch is short, discriminator metrics on latents + real images."""
    _patch_small(ref)
    try:
        torch.manual_seed(41); np.random.seed(41)
        if model == "StyleGAN":
            L, cfg = _build_style_learner(ref, res, res, bs)
        else:
            cfg = make_config("ProGAN", res=res, init_res=res, batch_size=bs, len_latent=SMALL_FMAP_MAX)
            with _quiet():
                L = ref.progan_learner.ProGANLearner(cfg)
        gen = torch.Generator().manual_seed(43)
        perturb_zero_params(L.gen_model, gen); perturb_zero_params(L.disc_model, gen)
        data = torch.rand(bs, 3, res, res, generator=gen) * 2 - 1
        ds = TensorDataset(data)
        dl = DataLoader(ds, batch_sampler=BatchSampler(SequentialSampler(ds), batch_size=bs, drop_last=True))
        with Tape() as train_tape, _quiet(), contextlib.redirect_stderr(io.StringIO()):
            L.train(dl, num_main_iters=1)
        g_sd, d_sd = sd_clone(L.gen_model), sd_clone(L.disc_model)
        z_valid = torch.randn(n_valid, cfg.len_latent, generator=gen)
        x_valid = torch.rand(n_valid, 3, res, res, generator=gen) * 2 - 1
        zds, xds = TensorDataset(z_valid), TensorDataset(x_valid)
        z_dl = DataLoader(zds, batch_sampler=BatchSampler(SequentialSampler(zds), batch_size=bs, drop_last=False))
        x_dl = DataLoader(xds, batch_sampler=BatchSampler(SequentialSampler(xds), batch_size=bs, drop_last=False))
        gen_metrics = ["fake realness", "generator loss"]
        disc_metrics = ["fake realness", "real realness", "discriminator loss"]
        # the reference returns the values formatted with %.4g; the harness also keeps the raw floats its `.item()` calls produce
        items = []
        orig_item = torch.Tensor.item
        torch.Tensor.item = lambda self: (items.append(orig_item(self)), items[-1])[1]
        try:
            with Tape() as tape_g, _quiet(), contextlib.redirect_stderr(io.StringIO()):
                vals_g = L.compute_metrics(metrics=gen_metrics, metrics_type="Generator", z_valid_dl=z_dl, valid_dl=None)
            raw_g = [float(v) for v in items[-len(gen_metrics):]]
            with Tape() as tape_d, _quiet(), contextlib.redirect_stderr(io.StringIO()):
                vals_d = L.compute_metrics(metrics=disc_metrics, metrics_type="Discriminator", z_valid_dl=z_dl, valid_dl=x_dl)
            raw_d = [float(v) for v in items[-len(disc_metrics):]]
        finally:
            torch.Tensor.item = orig_item
        out = dict(model=model, res=res, bs=bs, fmap_max=SMALL_FMAP_MAX, len_latent=cfg.len_latent, g_sd=g_sd, d_sd=d_sd,
                   z_valid=z_valid, x_valid=x_valid, gen_metrics=gen_metrics, disc_metrics=disc_metrics,
                   tape_g=tape_g.events, tape_d=tape_d.events, vals_g=vals_g, vals_d=vals_d, raw_g=raw_g, raw_d=raw_d,
                   gen_metrics_num=int(L.gen_metrics_num), disc_metrics_num=int(L.disc_metrics_num),
                   modes=(bool(L.gen_model.training), bool(L.disc_model.training)))
        if model == "StyleGAN":
            out["w_ewma"] = L.gen_model.w_ewma.detach().clone()
        return out
    finally:
        _unpatch(ref)


# Configuration switches of the train step that the headline fixtures do not exercise (each is an attribute the reference's loop
# or networks read, SURVEY.md section 8b "Learner surface to keep"); two main iterations each at 8x8 with 16 feature maps.
TRAIN_VARIANTS = {
    "style_minimax_r2": ("StyleGAN", dict(loss="minimax", gradient_penalty="r2")),
    "style_wgan_wgangp_gamma": ("StyleGAN", dict(loss="wgan", gradient_penalty="wgan-gp", gamma=750., lda=5.)),
    # (gen_bs_mult > 1 only works with the wgan loss in the reference: its BCE targets keep the discriminator's batch size)
    "style_two_d_iters_gen_bs_mult": ("StyleGAN", dict(num_disc_iters=2, gen_bs_mult=2, loss="wgan")),
    "style_no_noise_no_in_pixelnorm": ("StyleGAN", dict(use_noise=False, use_instancenorm=False, use_pixelnorm=True)),
    "style_no_mixing_no_ewma_uniform": ("StyleGAN", dict(pct_mixing_reg=0., use_ewma_gen=False, latent_distribution="uniform",
                                                          normalize_z=False)),
    # (gradient_penalty=None is a documented choice, config.py:108, but the reference's property crashes on it, :908)
    "style_linear_decay_no_drift": ("StyleGAN", dict(lr_sched="linear decay", eps_drift=0.)),
    "style_nearest_pool_no_blur": ("StyleGAN", dict(model_downsample_type="nearest", blur_type=None, mbstd_group_size=2)),
    "style_not_equalized_relu": ("StyleGAN", dict(use_equalized_lr=False, nonlinearity="relu", mapping_lrmul=1.)),
    "pro_nonsaturating_r1_no_pixelnorm": ("ProGAN", dict(loss="nonsaturating", gradient_penalty="r1", use_pixelnorm=False)),
    "pro_two_gen_iters_no_sched": ("ProGAN", dict(num_gen_iters=2, lr_sched=None, beta1=.5, wd=1e-3)),
}
VARIANT_FMAP = 16


def golden_train_variants(ref, res=8, bs=4, iters=2):
    out = {}
    _patch_small(ref, VARIANT_FMAP)
    try:
        for name, (model, over) in TRAIN_VARIANTS.items():
            torch.manual_seed(300 + len(out)); np.random.seed(300 + len(out))
            if model == "StyleGAN":
                over_ = dict(over)
                over_.setdefault("cutoff_trunc_trick", int(np.log2(res)) - 2)
                cfg = make_config("StyleGAN", res=res, init_res=res, batch_size=bs, len_latent=VARIANT_FMAP,
                                  len_dlatent=VARIANT_FMAP, **over_)
                with _quiet():
                    L = ref.stylegan_learner.StyleGANLearner(cfg)
            else:
                cfg = make_config("ProGAN", res=res, init_res=res, batch_size=bs, len_latent=VARIANT_FMAP, **over)
                with _quiet():
                    L = ref.progan_learner.ProGANLearner(cfg)
            gen = torch.Generator().manual_seed(400 + len(out))
            perturb_zero_params(L.gen_model, gen); perturb_zero_params(L.disc_model, gen)
            g0, d0 = sd_clone(L.gen_model), sd_clone(L.disc_model)
            n_d = cfg.num_disc_iters
            data = torch.rand(iters * n_d * bs, 3, res, res, generator=gen) * 2 - 1
            ds = TensorDataset(data)
            dl = DataLoader(ds, batch_sampler=BatchSampler(SequentialSampler(ds), batch_size=bs, drop_last=True))
            losses, lrs, iter_snaps = [], [], {}
            orig_backward = torch.Tensor.backward
            per_iter = cfg.num_disc_iters + cfg.num_gen_iters

            def rec_backward(self, *a, **k):
                if len(losses) and len(losses) % per_iter == 0:      # first backward of a later main iteration
                    iter_snaps[len(losses) // per_iter] = (sd_clone(L.gen_model), sd_clone(L.disc_model))
                losses.append(float(self.detach()))
                lrs.append((float(L.opt_disc.param_groups[0]["lr"]), float(L.opt_gen.param_groups[0]["lr"])))
                return orig_backward(self, *a, **k)

            torch.Tensor.backward = rec_backward
            try:
                with Tape() as tape, _quiet(), contextlib.redirect_stderr(io.StringIO()):
                    L.train(dl, num_main_iters=iters)
            finally:
                torch.Tensor.backward = orig_backward
            lagged = {k: v.detach().clone() for k, v in L.lagged_params.items()} if cfg.use_ewma_gen else None
            out[name] = dict(model=model, over=over, res=res, bs=bs, iters=iters, fmap_max=VARIANT_FMAP,
                             len_latent=cfg.len_latent, g_sd0=g0, d_sd0=d0, data=data, tape=tape.events, losses=losses, lrs=lrs,
                             iter_snaps=iter_snaps,      # parameters at the start of main iteration 1, 2, ...
                             g_sd1=sd_clone(L.gen_model), d_sd1=sd_clone(L.disc_model), lagged=lagged,
                             lr=cfg.lr_base * (cfg.lr_fctr_dict[res] if cfg.lr_sched == "resolution dependent" else 1.))
        return out
    finally:
        _unpatch(ref)


RESNET_FMAP = 8          # reference constants resnetgan/architectures.py:19-20 (FMAP_G = FMAP_D = 64) patched for small fixtures
RESNET_LATENT = 16


def _resnet_learner(ref, res, bs, **over):
    cfg = make_config("ResNet GAN", res=res, batch_size=bs, len_latent=RESNET_LATENT, **over)
    ref.resnet_learner.FMAP_G = ref.resnet_learner.FMAP_D = RESNET_FMAP
    try:
        with _quiet():
            L = ref.resnet_learner.GANLearner(cfg)
    finally:
        ref.resnet_learner.FMAP_G = ref.resnet_learner.FMAP_D = 64
    return L, cfg


def perturb_norm_params(module, gen):
    """BatchNorm / LayerNorm affine maps start at weight = 1, bias = 0: move them so that they matter."""
    with torch.no_grad():
        for m in module.modules():
            if isinstance(m, (torch.nn.BatchNorm2d, torch.nn.LayerNorm)):
                m.weight.copy_(1.0 + 0.3 * torch.randn(m.weight.shape, generator=gen))
                m.bias.copy_(0.3 * torch.randn(m.bias.shape, generator=gen))


def golden_resnet_nets(ref, res=64, bs=4):
    """ResNet generator / discriminator forward + backward (+ WGAN-GP through the LayerNorm blocks), reference modules."""
    torch.manual_seed(41 + res); np.random.seed(41)
    L, cfg = _resnet_learner(ref, res, bs)
    G, D = L.gen_model, L.disc_model
    gen = torch.Generator().manual_seed(43)
    perturb_zero_params(G, gen); perturb_zero_params(D, gen)
    perturb_norm_params(G, gen); perturb_norm_params(D, gen)
    G.train(); D.train()
    g_sd, d_sd = sd_clone(G), sd_clone(D)
    z = torch.randn(bs, cfg.len_latent, generator=gen)
    img = G(z)
    gimg = torch.randn(img.shape, generator=gen)
    G.zero_grad(); img.backward(gimg)
    g_grads = grads_of(G)
    g_buffers = {k: v.detach().clone() for k, v in G.named_buffers()}       # running statistics after ONE forward
    x = torch.rand(bs, 3, res, res, generator=gen) * 2 - 1
    D.zero_grad()
    logits = D(x); glog = torch.randn(logits.shape, generator=gen); logits.backward(glog)
    d_grads = grads_of(D)
    D.zero_grad()
    with Tape() as tape:
        gp = L.calc_gp(img.detach(), x)          # wgan-gp: draws eps via torch.rand
    gp.backward()
    d_gp_grads = grads_of(D)
    return dict(res=res, bs=bs, fmap=RESNET_FMAP, len_latent=cfg.len_latent, g_sd=g_sd, d_sd=d_sd, z=z, img=img.detach(),
                gimg=gimg, g_grads=g_grads, g_buffers=g_buffers, x=x, logits=logits.detach(), glog=glog, d_grads=d_grads,
                gp=gp.detach(), gp_tape=tape.events, d_gp_grads=d_gp_grads, lda=cfg.lda, gamma=cfg.gamma)


def golden_resnet_train(ref, res=64, bs=4, iters=2, num_disc_iters=2, **over):
    """GANLearner.train() (ResNet GAN: generator step first, then num_disc_iters discriminator steps), unmodified loop."""
    torch.manual_seed(51); np.random.seed(51)
    # lr 1e-5: with Adam(beta1=0) every step moves each parameter by ~+-lr whatever the gradient's size; at larger lr the sign
    # flips of noise-level gradients make the trajectory chaotic (the reference run twice with 1 vs 8 threads disagrees on
    # 37 % of the discriminator's parameters after 2 iterations at lr 1e-4, on 1.3 % at 1e-5)
    L, cfg = _resnet_learner(ref, res, bs, num_disc_iters=num_disc_iters, lr_base=1e-5, **over)
    gen = torch.Generator().manual_seed(53)
    perturb_zero_params(L.gen_model, gen); perturb_zero_params(L.disc_model, gen)
    perturb_norm_params(L.gen_model, gen); perturb_norm_params(L.disc_model, gen)
    g0, d0 = sd_clone(L.gen_model), sd_clone(L.disc_model)
    data = torch.rand(iters * num_disc_iters * bs, 3, res, res, generator=gen) * 2 - 1
    ds = TensorDataset(data)
    dl = DataLoader(ds, batch_sampler=BatchSampler(SequentialSampler(ds), batch_size=bs, drop_last=True))
    losses, lrs = [], []
    orig_backward = torch.Tensor.backward

    def rec_backward(self, *a, **k):
        losses.append(float(self.detach()))
        lrs.append((float(L.opt_disc.param_groups[0]["lr"]), float(L.opt_gen.param_groups[0]["lr"])))
        return orig_backward(self, *a, **k)

    torch.Tensor.backward = rec_backward
    try:
        with Tape() as tape, _quiet(), contextlib.redirect_stderr(io.StringIO()):
            L.train(dl, num_main_iters=iters)
    finally:
        torch.Tensor.backward = orig_backward
    return dict(model="ResNet GAN", res=res, bs=bs, iters=iters, num_disc_iters=num_disc_iters, fmap=RESNET_FMAP, over=over,
                len_latent=cfg.len_latent, g_sd0=g0, d_sd0=d0, data=data, tape=tape.events, losses=losses, lrs=lrs,
                g_sd1=sd_clone(L.gen_model), d_sd1=sd_clone(L.disc_model), lr=cfg.lr_base)


def golden_resnet_resume(ref, res=32, bs=4, num_disc_iters=2):
    """ResNet GAN: the reference trains one main iteration, saves (resnetgan/learner.py:1076-1137 -- with its Adam state: this
    learner does not rebuild its optimisers before a save), a fresh reference learner loads the file and trains one more."""
    torch.manual_seed(71); np.random.seed(71)
    L, cfg = _resnet_learner(ref, res, bs, num_disc_iters=num_disc_iters, lr_base=1e-5)
    gen = torch.Generator().manual_seed(73)
    perturb_zero_params(L.gen_model, gen); perturb_zero_params(L.disc_model, gen)
    perturb_norm_params(L.gen_model, gen); perturb_norm_params(L.disc_model, gen)
    data = torch.rand(2 * num_disc_iters * bs, 3, res, res, generator=gen) * 2 - 1

    def loader():
        ds = TensorDataset(data)
        return DataLoader(ds, batch_sampler=BatchSampler(SequentialSampler(ds), batch_size=bs, drop_last=True))

    with _quiet(), contextlib.redirect_stderr(io.StringIO()):
        L.train(loader(), num_main_iters=1)
    L.valid_z = torch.zeros(cfg.img_grid_sz ** 2, cfg.len_latent)          # see golden_resume
    ckpt = GOLDEN_DIR / "resnetgan_reference_checkpoint.tar"
    # save_model() writes the module constants FMAP_G / FMAP_D (resnetgan/learner.py:1082-1085): keep them at the small fixtures' value
    ref.resnet_learner.FMAP_G = ref.resnet_learner.FMAP_D = RESNET_FMAP
    try:
        L.save_model(ckpt)
    finally:
        ref.resnet_learner.FMAP_G = ref.resnet_learner.FMAP_D = 64
    saved = dict(g_sd=sd_clone(L.gen_model), d_sd=sd_clone(L.disc_model), img_num=int(L.curr_img_num),
                 opt_disc_steps=sorted({float(st["step"]) for st in L.opt_disc.state_dict()["state"].values()}))
    L2, _ = _resnet_learner(ref, res, bs, num_disc_iters=num_disc_iters, lr_base=1e-5)
    orig_load = torch.load
    torch.load = lambda *a, **k: orig_load(*a, **{"weights_only": False, **k})
    try:
        with _quiet():
            L2.load_model(ckpt, dev_of_saved_model="cpu")
    finally:
        torch.load = orig_load
    losses = []
    orig_backward = torch.Tensor.backward

    def rec_backward(self, *a, **k):
        losses.append(float(self.detach()))
        return orig_backward(self, *a, **k)

    torch.Tensor.backward = rec_backward
    try:
        with Tape() as tape, _quiet(), contextlib.redirect_stderr(io.StringIO()):
            L2.train(loader(), num_main_iters=1)
    finally:
        torch.Tensor.backward = orig_backward
    return dict(model="ResNet GAN", checkpoint=ckpt.name, res=res, bs=bs, num_disc_iters=num_disc_iters, fmap=RESNET_FMAP,
                len_latent=cfg.len_latent, data=data, saved=saved, tape=tape.events, losses=losses, lr=cfg.lr_base,
                g_sd1=sd_clone(L2.gen_model), d_sd1=sd_clone(L2.disc_model), img_num=int(L2.curr_img_num),
                opt_disc_steps=sorted({float(st["step"]) for st in L2.opt_disc.state_dict()["state"].values()}))


def golden_state_dict_shapes(ref):
    """Parameter / buffer names and shapes of the reference's FULL-SIZE networks (no patched constants): the contract that lets
    its checkpoints load into the drop-in modules (SURVEY.md 8b 'Parameter names/shapes')."""
    out = {}
    shapes = lambda m: [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
    for model in ("StyleGAN", "ProGAN"):
        for res, init_res in ((4, 4), (128, 128), (256, 128), (1024, 1024)):
            torch.manual_seed(0)
            cfg = make_config(model, res=res, init_res=init_res, batch_size=8,
                              cutoff_trunc_trick=(min(4, int(np.log2(res)) - 2) or None))
            with _quiet():
                L = (ref.stylegan_learner.StyleGANLearner if model == "StyleGAN" else ref.progan_learner.ProGANLearner)(cfg)
            if init_res != res:                     # cfg3: grown once more, mid-fade-in (prev_torgb / prev_fromrgb present)
                L.gen_model.increase_scale(); L.disc_model.increase_scale()
            out[(model, res, init_res)] = dict(g=shapes(L.gen_model), d=shapes(L.disc_model))
            del L
    for res in (32, 64):
        cfg = make_config("ResNet GAN", res=res, batch_size=8)
        with _quiet():
            L = ref.resnet_learner.GANLearner(cfg)
        out[("ResNet GAN", res, res)] = dict(g=shapes(L.gen_model), d=shapes(L.disc_model))
    return out


def _digest(t):
    import hashlib
    return hashlib.sha256(t.detach().contiguous().cpu().numpy().tobytes()).hexdigest()


def golden_init_digests(ref):
    """sha256 of every freshly initialised parameter / buffer of the reference's networks under a fixed seed (tapering small
    networks, grown once after construction; both ResNets at full size): the drop-in modules draw the same numbers in the same
    order, so `Learner(config)` under the same seed starts from bit-identical weights."""
    out = {}
    _patch_small(ref, SMALL_FMAP_MAX, TAPER_FMAP_BASE)
    try:
        for model in ("StyleGAN", "ProGAN"):
            torch.manual_seed(5); np.random.seed(5)
            kw = dict(res=32, init_res=16, batch_size=4, len_latent=SMALL_FMAP_MAX)
            if model == "StyleGAN":
                kw.update(len_dlatent=SMALL_FMAP_MAX, cutoff_trunc_trick=2)
            with _quiet():
                L = (ref.stylegan_learner.StyleGANLearner if model == "StyleGAN" else ref.progan_learner.ProGANLearner)(
                    make_config(model, **kw))
            L.gen_model.increase_scale(); L.disc_model.increase_scale()
            out[model] = dict(seed=5, kw=kw, fmap_max=SMALL_FMAP_MAX, fmap_base=TAPER_FMAP_BASE,
                              g={k: _digest(v) for k, v in L.gen_model.state_dict().items()},
                              d={k: _digest(v) for k, v in L.disc_model.state_dict().items()})
    finally:
        _unpatch(ref)
    for res in (32, 64):
        torch.manual_seed(7)
        with _quiet():
            L = ref.resnet_learner.GANLearner(make_config("ResNet GAN", res=res, batch_size=4))
        out[f"ResNet GAN {res}"] = dict(seed=7, res=res, g={k: _digest(v) for k, v in L.gen_model.state_dict().items()},
                                        d={k: _digest(v) for k, v in L.disc_model.state_dict().items()})
    return out


def _cfg2_fullwidth_run(ref, res, bs, seed, ulp_perturb=False, dev="cpu", replay=None):
    """ONE main iteration (D step + G step, Adam, EWMA) of the UNMODIFIED reference at cfg2's REAL widths (StyleGAN 128x128,
    FMAP_MAX 512: 512 -> 256 -> 128 channels, batch 8, nonsaturating + R1 + drift, noise, mixing .9) -- the configuration
    bench.py times.  49 M parameters and their gradients do not fit a committed fixture, so: the initial weights are NOT
    stored (the drop-in learner built under the same seed starts from bit-identical weights, see init_digests.pt; their
    sha256 digests are stored and asserted), the real batch is regenerated from its generator seed, the reference's random
    draws are taped as usual, and every gradient / post-Adam parameter / EWMA tensor is stored as an oracle.summaries summary
    (norm, max-norm, strided sample, random projections)."""
    from oracle import summaries as S
    torch.manual_seed(seed); np.random.seed(seed)
    # dev / replay: the same run on another device with the CPU run's taped draws replayed (tests/calibrate_tf32_bounds.py:
    # the unmodified reference through stock PyTorch on the B200); the fixture itself is dev = "cpu", replay = None
    cfg = make_config("StyleGAN", res=res, init_res=res, batch_size=bs, dev=dev, metrics_dev=torch.device(dev))
    with _quiet():
        L = ref.stylegan_learner.StyleGANLearner(cfg)
    gen = torch.Generator().manual_seed(seed + 1)
    perturb_zero_params(L.gen_model, gen); perturb_zero_params(L.disc_model, gen)
    taping = (lambda key: Tape()) if replay is None else (lambda key: ReplayTape(replay[key], dev))
    if ulp_perturb:          # every weight moved by about one unit in the last place (see golden_cfg2_fullwidth_step)
        pg = torch.Generator().manual_seed(seed + 2)
        with torch.no_grad():
            for p in list(L.gen_model.parameters()) + list(L.disc_model.parameters()):
                p.mul_(1. + 6e-8 * (torch.empty(p.shape).random_(0, 2, generator=pg) * 2 - 1))
    g_dig = {k: _digest(v) for k, v in L.gen_model.state_dict().items()}
    d_dig = {k: _digest(v) for k, v in L.disc_model.state_dict().items()}
    p0 = {"g." + k: S.summarize("p0.g." + k, v) for k, v in L.gen_model.state_dict().items()}
    p0.update({"d." + k: S.summarize("p0.d." + k, v) for k, v in L.disc_model.state_dict().items()})
    data = torch.rand(bs, 3, res, res, generator=gen) * 2 - 1
    ds = TensorDataset(data)
    dl = DataLoader(ds, batch_sampler=BatchSampler(SequentialSampler(ds), batch_size=bs, drop_last=True))
    names = {id(p): "g." + n for n, p in L.gen_model.named_parameters()}
    names.update({id(p): "d." + n for n, p in L.disc_model.named_parameters()})
    # the R1 penalty on its own at the initial weights (inside the whole loss it is ~1e-8 of it at random init):
    # value and every discriminator gradient of calc_gp's double backward
    L.disc_model.train(); L.disc_model.zero_grad()
    L.disc_model.to(dev); L.gen_model.to(dev)
    gp_alone = L.calc_gp(data.clone().to(dev), data.clone().to(dev))          # R1 only looks at the real batch (resnetgan/learner.py:799-802)
    gp_alone.backward()
    gp_grads = {"d." + n: S.summarize("gpgrad.d." + n, p.grad) for n, p in L.disc_model.named_parameters() if p.grad is not None}
    L.disc_model.zero_grad()
    # the generator step's forward / backward on the INITIAL weights (inside train() it runs behind the discriminator's first
    # Adam step, whose sign-like updates amplify rounding noise): G(z) with taped draws -> D -> non-saturating loss
    # (progan/learner.py:883-896: binary_cross_entropy_with_logits against ones) -> every generator gradient
    L.gen_model.train(); L.gen_model.zero_grad()
    z_alone = torch.randn(bs, cfg.len_latent, generator=gen)
    with taping("g_alone") as tape_alone:
        img_alone = L.gen_model(z_alone.to(dev))
    for p in L.disc_model.parameters():
        p.requires_grad_(False)
    logits_alone = L.disc_model(img_alone).view(-1)
    g_alone_loss = torch.nn.functional.binary_cross_entropy_with_logits(logits_alone, torch.ones_like(logits_alone))
    g_alone_loss.backward()
    for p in L.disc_model.parameters():
        p.requires_grad_(True)
    g_alone = dict(z=z_alone, tape=tape_alone.events, loss=float(g_alone_loss.detach()), img=S.summarize("galone.img", img_alone),
                   logits=logits_alone.detach().cpu().clone(),
                   grads={"g." + n: S.summarize("galone.g." + n, p.grad) for n, p in L.gen_model.named_parameters()
                          if p.grad is not None})
    L.gen_model.zero_grad(); L.disc_model.zero_grad()
    L.gen_model.w_ewma = None            # the stand-alone forward created it; train() must start like a fresh learner
    losses, grads, gp_vals = [], {}, []
    orig_backward, orig_step, orig_gp = torch.Tensor.backward, torch.optim.Adam.step, L.calc_gp

    def rec_backward(self, *a, **k):
        losses.append(float(self.detach()))
        return orig_backward(self, *a, **k)

    def rec_step(self, *a, **k):
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is not None:
                    grads[names[id(p)]] = S.summarize("grad." + names[id(p)], p.grad)
        return orig_step(self, *a, **k)

    def rec_gp(*a, **k):
        v = orig_gp(*a, **k)
        gp_vals.append(float(v.detach()))
        return v

    torch.Tensor.backward, torch.optim.Adam.step, L.calc_gp = rec_backward, rec_step, rec_gp
    try:
        with taping("train") as tape, _quiet(), contextlib.redirect_stderr(io.StringIO()):
            L.train(dl, num_main_iters=1)
    finally:
        torch.Tensor.backward, torch.optim.Adam.step = orig_backward, orig_step
    p1 = {"g." + k: S.summarize("p1.g." + k, v) for k, v in L.gen_model.state_dict().items()}
    p1.update({"d." + k: S.summarize("p1.d." + k, v) for k, v in L.disc_model.state_dict().items()})
    lagged = {k: S.summarize("lag." + k, v) for k, v in L.lagged_params.items()}
    return dict(model="StyleGAN", res=res, bs=bs, seed=seed, lr=cfg.lr_base * cfg.lr_fctr_dict[res], lda=cfg.lda,
                g_digests=g_dig, d_digests=d_dig, data_digest=_digest(data), tape=tape.events, losses=losses, gp=gp_vals,
                p0=p0, grads=grads, p1=p1, lagged=lagged, beta=float(L.beta), gp_alone=float(gp_alone.detach()), gp_grads=gp_grads, g_alone=g_alone,
                w_ewma=L.gen_model.w_ewma.detach().cpu().clone())



def _flip_fraction(p1a, p1b, lr):
    """Fraction of the sampled post-Adam parameter elements that differ by more than 2 % of the learning rate."""
    bad = tot = 0
    for k, sa in p1a.items():
        d = (sa["sample"] - p1b[k]["sample"]).abs()
        bad += int((d > 0.02 * lr).sum()); tot += d.numel()
    return bad / tot


def golden_cfg2_fullwidth_step(ref, res=128, bs=8, seed=1234):
    """The fixture proper + the reference's OWN noise floor at this size: the same iteration re-run with every weight moved by
    about one fp32 unit in the last place (relative 6e-8, random sign); `self_noise` holds, per class, the worst relative L2
    difference (over tensors) between the two reference runs.  Cancellation-dominated gradients (biases / noise weights in
    front of InstanceNorm, everything behind leaky-ReLU masks that sit within rounding of zero at random init) amplify a 1-ulp
    perturbation by 3-4 orders of magnitude; the GPU tests' bounds are stated as multiples of this floor."""
    from oracle import summaries as S
    a = _cfg2_fullwidth_run(ref, res, bs, seed, False)
    b = _cfg2_fullwidth_run(ref, res, bs, seed, True)

    def worst(x, y):
        out = dict(l2=0.0, smax=0.0)
        for k, sx in x.items():
            sy = y[k]
            scale = max(sx["norm"], 1e-30)
            out["l2"] = max(out["l2"], float((sx["proj"] - sy["proj"]).abs().max()) / scale)
            out["smax"] = max(out["smax"], float((sx["sample"] - sy["sample"]).abs().max()) / max(sx["absmax"], 1e-30))
        return out

    gd = lambda d, tag: {k: v for k, v in d.items() if k.startswith(tag)}
    a["self_noise"] = dict(
        perturbation="weights * (1 +- 6e-8)",
        loss_d=abs(a["losses"][0] - b["losses"][0]) / max(1.0, abs(a["losses"][0])),
        loss_g=abs(a["losses"][1] - b["losses"][1]) / max(1.0, abs(a["losses"][1])),
        gp_value=abs(a["gp_alone"] - b["gp_alone"]) / abs(a["gp_alone"]),
        gp_grads=worst(a["gp_grads"], b["gp_grads"]),
        g_alone_loss=abs(a["g_alone"]["loss"] - b["g_alone"]["loss"]) / max(1.0, abs(a["g_alone"]["loss"])),
        g_alone_img=worst({"img": a["g_alone"]["img"]}, {"img": b["g_alone"]["img"]}),
        g_alone_grads=worst(a["g_alone"]["grads"], b["g_alone"]["grads"]),
        d_grads=worst(gd(a["grads"], "d."), gd(b["grads"], "d.")),
        g_grads=worst(gd(a["grads"], "g."), gd(b["grads"], "g.")),
        p1_flip_fraction=_flip_fraction(a["p1"], b["p1"], a["lr"]))
    return a

def golden_resnet_metrics(ref, res=32, bs=4, n_valid=10):
    """compute_metrics() of the reference's ResNet learner (resnetgan/learner.py:318-460) after one training iteration: its
    generator runs in eval mode, i.e. BatchNorm on the running statistics that iteration left behind."""
    torch.manual_seed(81); np.random.seed(81)
    L, cfg = _resnet_learner(ref, res, bs, num_disc_iters=2, lr_base=1e-5)
    gen = torch.Generator().manual_seed(83)
    perturb_zero_params(L.gen_model, gen); perturb_zero_params(L.disc_model, gen)
    perturb_norm_params(L.gen_model, gen); perturb_norm_params(L.disc_model, gen)
    data = torch.rand(2 * bs, 3, res, res, generator=gen) * 2 - 1
    ds = TensorDataset(data)
    dl = DataLoader(ds, batch_sampler=BatchSampler(SequentialSampler(ds), batch_size=bs, drop_last=True))
    with _quiet(), contextlib.redirect_stderr(io.StringIO()):
        L.train(dl, num_main_iters=1)
    g_sd, d_sd = sd_clone(L.gen_model), sd_clone(L.disc_model)
    z_valid = torch.randn(n_valid, cfg.len_latent, generator=gen)
    x_valid = torch.rand(n_valid, 3, res, res, generator=gen) * 2 - 1
    zds, xds = TensorDataset(z_valid), TensorDataset(x_valid)
    z_dl = DataLoader(zds, batch_sampler=BatchSampler(SequentialSampler(zds), batch_size=bs, drop_last=False))
    x_dl = DataLoader(xds, batch_sampler=BatchSampler(SequentialSampler(xds), batch_size=bs, drop_last=False))
    gen_metrics = ["fake realness", "generator loss"]
    disc_metrics = ["fake realness", "real realness", "discriminator loss"]
    items = []
    orig_item = torch.Tensor.item
    torch.Tensor.item = lambda self: (items.append(orig_item(self)), items[-1])[1]
    try:
        with _quiet(), contextlib.redirect_stderr(io.StringIO()):
            vals_g = L.compute_metrics(metrics=gen_metrics, metrics_type="Generator", z_valid_dl=z_dl, valid_dl=None)
        raw_g = [float(v) for v in items[-len(gen_metrics):]]
        with _quiet(), contextlib.redirect_stderr(io.StringIO()):
            vals_d = L.compute_metrics(metrics=disc_metrics, metrics_type="Discriminator", z_valid_dl=z_dl, valid_dl=x_dl)
        raw_d = [float(v) for v in items[-len(disc_metrics):]]
    finally:
        torch.Tensor.item = orig_item
    return dict(model="ResNet GAN", res=res, bs=bs, fmap=RESNET_FMAP, len_latent=cfg.len_latent, g_sd=g_sd, d_sd=d_sd,
                z_valid=z_valid, x_valid=x_valid, gen_metrics=gen_metrics, disc_metrics=disc_metrics, vals_g=vals_g, vals_d=vals_d,
                raw_g=raw_g, raw_d=raw_d)


def main():
    ref = load_reference()
    GOLDEN_DIR.mkdir(parents=True, exist_ok=True)
    jobs = {
        "layers.pt": lambda: golden_layers(ref),
        "style_nets_res16.pt": lambda: golden_style_nets(ref, 16, 4, False),
        "style_nets_res16_fade.pt": lambda: golden_style_nets(ref, 16, 4, True),
        "style_eval_res32.pt": lambda: golden_style_eval(ref, 32, 3),
        "pro_nets_res16.pt": lambda: golden_pro_nets(ref, 16, 4, False),
        "pro_nets_res8_fade.pt": lambda: golden_pro_nets(ref, 8, 4, True),
        "style_train_res16.pt": lambda: golden_train_steps(ref, "StyleGAN", 16, 4, 2),
        "pro_train_res8.pt": lambda: golden_train_steps(ref, "ProGAN", 8, 4, 2),
        "style_nets_res32_taper_fade.pt": lambda: golden_style_nets(ref, 32, 4, True, fmap_base=TAPER_FMAP_BASE),
        "pro_nets_res32_taper_fade.pt": lambda: golden_pro_nets(ref, 32, 4, True, fmap_base=TAPER_FMAP_BASE),
        "style_grow_8to16_taper.pt": lambda: golden_grow(ref, "StyleGAN", init_res=8, res=16, data_res=16, iters=10,
                                                         snap_iters=(2, 4, 5, 6, 8), fmap_base=TAPER_FMAP_BASE),
        "pro_grow_8to16_taper.pt": lambda: golden_grow(ref, "ProGAN", init_res=8, res=16, data_res=16, iters=10,
                                                       snap_iters=(2, 4, 5, 6, 8), fmap_base=TAPER_FMAP_BASE),
        "style_grow_4to8.pt": lambda: golden_grow(ref, "StyleGAN"),
        "pro_grow_4to8.pt": lambda: golden_grow(ref, "ProGAN"),
        "style_resume.pt": lambda: golden_resume(ref, "StyleGAN"),
        "pro_resume.pt": lambda: golden_resume(ref, "ProGAN"),
        "style_metrics.pt": lambda: golden_metrics(ref, "StyleGAN"),
        "pro_metrics.pt": lambda: golden_metrics(ref, "ProGAN"),
        "train_variants.pt": lambda: golden_train_variants(ref),
        "state_dict_shapes.pt": lambda: golden_state_dict_shapes(ref),
        "init_digests.pt": lambda: golden_init_digests(ref),
        "resnet_nets_res64.pt": lambda: golden_resnet_nets(ref, 64, 4),
        "resnet_nets_res32.pt": lambda: golden_resnet_nets(ref, 32, 4),
        "resnet_train_res64.pt": lambda: golden_resnet_train(ref, 64, 4, 2, 2),
        "resnet_resume_res32.pt": lambda: golden_resnet_resume(ref),
        "resnet_metrics_res32.pt": lambda: golden_resnet_metrics(ref),
        "style_cfg2_fullwidth_step.pt": lambda: golden_cfg2_fullwidth_step(ref),
        "resnet_train_res32_variant.pt": lambda: golden_resnet_train(
            ref, 32, 4, 2, 1, loss="nonsaturating", gradient_penalty="r1", num_gen_iters=2,     # (equalized LR crashes in the reference ResNets: wscale None)
            lr_sched="linear decay", nonlinearity="leaky relu"),
    }
    only = sys.argv[1:]
    for name, fn in jobs.items():
        if only and name not in only:
            continue
        obj = fn()
        torch.save(obj, GOLDEN_DIR / name)
        print(f"wrote {name}: {(GOLDEN_DIR / name).stat().st_size / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
