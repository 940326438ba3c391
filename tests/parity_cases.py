"""Parity case bodies shared by the CPU host-wiring tests (kernels replaced by their CPU contracts) and the
GPU parity tests (real sm_100a kernels through the C-ABI).  Expected values: fixtures from the UNMODIFIED reference."""
import contextlib
import math

import torch

import gan_lab_b200._kernels as K
from gan_lab_b200 import ops
from gan_lab_b200.config import default_config
from gan_lab_b200.utils import custom_layers as CL
from gan_lab_b200.utils.latent_utils import TapeSource, set_random_source


CONV_CASES = ["conv3x3", "conv3x3_nobias", "conv1x1_torgb", "conv1x1_fromrgb", "conv4x4_valid", "conv3x3_c33",
              "conv3x3_lrmul"]
LINEAR_CASES = ["linear_mapping", "linear_style", "linear_dhead", "linear_progan_fc"]
MBSTD_CASES = ["mbstd_n8", "mbstd_n4", "mbstd_n6", "mbstd_n1", "mbstd_n16"]
STYLE_NETS = ["style_nets_res16.pt", "style_nets_res16_fade.pt"]
PRO_NETS = ["pro_nets_res16.pt", "pro_nets_res8_fade.pt"]
# channel counts that halve per stage (32, 32, 16, 8), like 512 -> 256 -> 128 in the full-size networks
STYLE_NETS_TAPER = ["style_nets_res32_taper_fade.pt"]
PRO_NETS_TAPER = ["pro_nets_res32_taper_fade.pt"]
TRAIN_CASES = [("style_train_res16.pt", "StyleGAN"), ("pro_train_res8.pt", "ProGAN")]
GROW_CASES = [("style_grow_4to8.pt", "StyleGAN"), ("pro_grow_4to8.pt", "ProGAN")]
GROW_TAPER_CASES = [("style_grow_8to16_taper.pt", "StyleGAN"), ("pro_grow_8to16_taper.pt", "ProGAN")]
RESUME_CASES = [("style_resume.pt", "StyleGAN"), ("pro_resume.pt", "ProGAN")]
METRICS_CASES = [("style_metrics.pt", "StyleGAN"), ("pro_metrics.pt", "ProGAN")]
TRAIN_VARIANTS = ["style_minimax_r2", "style_wgan_wgangp_gamma", "style_two_d_iters_gen_bs_mult", "style_no_noise_no_in_pixelnorm",
                  "style_no_mixing_no_ewma_uniform", "style_linear_decay_no_drift", "style_nearest_pool_no_blur",
                  "style_not_equalized_relu", "pro_nonsaturating_r1_no_pixelnorm", "pro_two_gen_iters_no_sched"]
RESNET_NETS = ["resnet_nets_res64.pt", "resnet_nets_res32.pt"]


@contextlib.contextmanager
def _fmap_base(g):
    """Networks are built under the FMAP_BASE the fixture's reference run used (default 8192; the `taper` fixtures use a small
    one so that the channel count halves per stage like in the full-size networks).  GrowthState reads it at construction."""
    import gan_lab_b200._growth as growth
    old = growth.FMAP_BASE
    growth.FMAP_BASE = g.get("fmap_base", 8192)
    try:
        yield
    finally:
        growth.FMAP_BASE = old


def _to(g, dev):
    """Move every tensor of a (nested) fixture to `dev`."""
    if torch.is_tensor(g):
        return g.to(dev)
    if isinstance(g, dict):
        return {k: _to(v, dev) for k, v in g.items()}
    if isinstance(g, (list, tuple)):
        return type(g)(_to(v, dev) for v in g)
    return g


def relerr(a, b):
    return float((a.detach() - b).abs().max() / b.abs().max().clamp_min(1e-30))


def close(a, b, rtol=1e-4, atol=1e-5):
    torch.testing.assert_close(a.detach().contiguous(), b.contiguous(), rtol=rtol, atol=atol)


def _load(module, sd):
    missing, unexpected = module.load_state_dict(sd, strict=True)
    assert not missing and not unexpected


def case_conv2d_ex_module(golden, dev, name):
    g = _to(golden("layers.pt")[name], dev)
    c = g["cfg"]
    m = CL.Conv2dEx(ni=c["ni"], nf=c["nf"], ks=c["ks"], stride=1, padding=c["padding"], init="He", init_type="StyleGAN",
                    gain_sq_base=c["gain_sq_base"], equalized_lr=True, lrmul=c["lrmul"], include_bias=c["include_bias"])
    assert abs(m.wscale - g["wscale"]) < 1e-12
    _load(m, g["sd"]); m.to(dev)
    assert m.conv2d.weight.is_contiguous(memory_format=torch.channels_last)
    x = g["x"].clone().requires_grad_(True)
    y = m(x)
    close(y, g["y"])
    gx, = torch.autograd.grad(y, x, g["gy"], create_graph=True)
    close(gx, g["gx"])
    m.zero_grad()
    (gx * g["v"]).sum().backward(retain_graph=True)          # double backward -> weight
    close(m.conv2d.weight.grad, g["gg_w"], rtol=1e-3, atol=1e-5)
    m.zero_grad()
    y.backward(g["gy"])
    for k, p in m.named_parameters():
        close(p.grad, g["grads"][k], rtol=1e-3, atol=1e-5)


def case_linear_ex_module(golden, dev, name):
    g = _to(golden("layers.pt")[name], dev)
    c = g["cfg"]
    m = CL.LinearEx(nin_feat=c["nin"], nout_feat=c["nout"], init="He", init_type="StyleGAN",
                    gain_sq_base=c["gain_sq_base"], equalized_lr=True, lrmul=c["lrmul"])
    _load(m, g["sd"]); m.to(dev)
    x = g["x"].clone().requires_grad_(True)
    y = m(x)
    close(y, g["y"])
    gx, = torch.autograd.grad(y, x, g["gy"], create_graph=True)
    close(gx, g["gx"])
    m.zero_grad()
    (gx * g["v"]).sum().backward(retain_graph=True)
    close(m.linear.weight.grad, g["gg_w"], rtol=1e-3, atol=1e-6)
    m.zero_grad()
    y.backward(g["gy"])
    for k, p in m.named_parameters():
        close(p.grad, g["grads"][k], rtol=1e-3, atol=1e-6)


def case_small_modules(golden, dev):
    L = _to(golden("layers.pt"), dev)
    g = L["conv2dbias"]
    m = CL.Conv2dBias(16, device=dev); _load(m, g["sd"]); m.to(dev)
    x = g["x"].clone().requires_grad_(True)
    y = m(x); close(y, g["y"]); y.backward(g["gy"])
    close(x.grad, g["gx"]); close(m.bias.grad, g["grads"]["bias"])
    for k in ("pixelnorm_z", "pixelnorm_feat"):
        g = L[k]
        x = g["x"].clone().requires_grad_(True)
        y = CL.PixelNorm2d()(x); close(y, g["y"]); y.backward(g["gy"]); close(x.grad, g["gx"])
    for k in ("blur", "blur_odd"):
        g = L[k]
        if g["x"].shape[1] % 4:
            continue                    # NHWC glue kernels need C % 4 == 0 (every gan-lab feature map has it)
        x = g["x"].clone().requires_grad_(True)
        gy = g["gy"].clone().requires_grad_(True)
        y = CL.get_blur_op("binomial", x.shape[1])(x); close(y, g["y"])
        gx, = torch.autograd.grad(y, x, gy, create_graph=True); close(gx, g["gx"])
        close(torch.autograd.grad(gx, gy, g["v"])[0], g["ggy"])
    g = L["upsample2x"]
    x = g["x"].clone().requires_grad_(True)
    y = CL.Upsample2x()(x); close(y, g["y"]); y.backward(g["gy"]); close(x.grad, g["gx"])
    g = L["pool_bias_lrelu"]
    from gan_lab_b200 import ops
    x = g["x"].clone().requires_grad_(True); gy = g["gy"].clone().requires_grad_(True)
    y = ops.pool_bias_act(x, g["bias"], 1.0, ops.ACT_LRELU, 0.2); close(y, g["y"])
    gx, = torch.autograd.grad(y, x, gy, create_graph=True); close(gx, g["gx"])
    close(torch.autograd.grad(gx, gy, g["v"])[0], g["ggy"])


def case_mbstd_module(golden, dev, name):
    g = _to(golden("layers.pt")[name], dev)
    x = g["x"].clone().requires_grad_(True)
    gy = g["gy"].clone().requires_grad_(True)
    y = CL.concat_mbstd_layer(x, g["group_size"])
    close(y, g["y"])
    gx, = torch.autograd.grad(y, x, gy, create_graph=True)
    close(gx, g["gx"])
    if x.shape[0] > 1:
        ggx, ggy = torch.autograd.grad(gx, (x, gy), g["v"])
        close(ggx, g["ggx"], rtol=1e-3, atol=1e-6); close(ggy, g["ggy"], rtol=1e-3, atol=1e-6)


def case_style_epilogue_op(golden, dev):
    from gan_lab_b200 import ops
    g = _to(golden("layers.pt")["style_epilogue"], dev)
    x = g["x"].clone().requires_grad_(True); st = g["style"].clone().requires_grad_(True)
    nw = g["noise_weight"].clone().requires_grad_(True); b = g["bias"].clone().requires_grad_(True)
    y = ops.style_epilogue(x, g["noise"], nw, b, st, 0.2, 1e-8)
    close(y, g["y"], atol=1e-5)
    y.backward(g["gy"])
    close(x.grad, g["gx"], rtol=1e-3); close(st.grad, g["gstyle"], rtol=1e-3)
    close(nw.grad, g["g_noise_weight"], rtol=1e-3); close(b.grad, g["g_bias"], rtol=1e-3)


def _grads_ok(module, ref, rtol, skip_cancelled=False, report=None, tag=""):
    """Per-tensor max-norm relative error.  skip_cancelled: gradients that are analytically zero (a conv bias feeding a
    BatchNorm) hold only rounding noise in the fixture; they are compared against 5e-2 of the largest gradient instead."""
    floor = 0.0
    if skip_cancelled:
        floor = 5e-2 * max(float(r.abs().max()) for r in ref.values() if r is not None)
    for k, p in module.named_parameters():
        r = ref[k]
        if r is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
        else:
            if p.grad is None:                      # autograd may leave an identically-zero gradient undefined
                assert float(r.abs().max()) == 0.0, k
                continue
            r = r.to(p.grad.dtype)
            err = float((p.grad.detach() - r).abs().max() / r.abs().max().clamp_min(max(floor, 1e-30)))
            if report is not None and err > report.get(tag, (0.0, None))[0]:
                report[tag] = (err, k)
            assert err < rtol, (k, err)


def _style_learner(g, fade, dev):
    from gan_lab_b200.stylegan.learner import StyleGANLearner
    res = g["res"]
    cfg = default_config("StyleGAN", res=res, init_res=res // 2 if fade else res, batch_size=g["bs"], dev=dev,
                         len_latent=g["len_latent"], len_dlatent=g["len_latent"], cutoff_trunc_trick=int(math.log2(res)) - 2)
    with _fmap_base(g):
        L = StyleGANLearner(cfg)
    if fade:
        L.gen_model.increase_scale(); L.disc_model.increase_scale()
        L.gen_model.alpha = g["alpha"]
    return L


def case_style_nets_modules(golden, dev, fname):
    g = _to(golden(fname), dev)
    L = _style_learner(g, g["fade_in"], dev)
    G, D = L.gen_model, L.disc_model
    assert list(G.state_dict().keys()) == list(g["g_sd"].keys())      # same names AND order as the reference
    assert list(D.state_dict().keys()) == list(g["d_sd"].keys())
    _load(G, g["g_sd"]); _load(D, g["d_sd"])
    G.train(); D.train()
    set_random_source(TapeSource(g["tape"], dev))
    img = G(g["z"])
    close(img, g["img"], rtol=1e-4, atol=1e-5)
    close(G.w_ewma, g["w_ewma"])
    G.zero_grad(); img.backward(g["gimg"])
    _grads_ok(G, g["g_grads"], 2e-3)
    D.zero_grad()
    logits = D(g["x"])
    close(logits, g["logits"], rtol=1e-4, atol=1e-5)
    logits.backward(g["glog"])
    _grads_ok(D, g["d_grads"], 2e-4)
    D.zero_grad()
    L.batch_size = g["bs"]
    pen = L.calc_gp(g["img"], g["x"])
    assert relerr(pen, g["gp"]) < 1e-4
    pen.backward()
    _grads_ok(D, g["d_gp_grads"], 5e-4)
    for p in D.parameters():
        p.requires_grad_(False)
    xi = g["img"].clone().requires_grad_(True)
    gxi, = torch.autograd.grad(D(xi), xi, g["glog"])
    assert relerr(gxi, g["d_gx_img"]) < 2e-4


def case_pro_nets_modules(golden, dev, fname):
    from gan_lab_b200.progan.learner import ProGANLearner
    g = _to(golden(fname), dev)
    res = g["res"]
    cfg = default_config("ProGAN", res=res, init_res=res // 2 if g["fade_in"] else res, batch_size=g["bs"], dev=dev,
                         len_latent=g["len_latent"])
    with _fmap_base(g):
        L = ProGANLearner(cfg)
    G, D = L.gen_model, L.disc_model
    if g["fade_in"]:
        G.increase_scale(); D.increase_scale(); G.alpha = g["alpha"]
    assert list(G.state_dict().keys()) == list(g["g_sd"].keys())
    assert list(D.state_dict().keys()) == list(g["d_sd"].keys())
    _load(G, g["g_sd"]); _load(D, g["d_sd"])
    G.train(); D.train()
    img = G(g["z"])
    close(img, g["img"], rtol=1e-4, atol=1e-5)
    G.zero_grad(); img.backward(g["gimg"]); _grads_ok(G, g["g_grads"], 2e-4)
    D.zero_grad()
    logits = D(g["x"]); close(logits, g["logits"], rtol=1e-4, atol=1e-5)
    logits.backward(g["glog"]); _grads_ok(D, g["d_grads"], 2e-4)
    D.zero_grad()
    L.batch_size = g["bs"]
    set_random_source(TapeSource(g["gp_tape"], dev))
    pen = L.calc_gp(g["img"], g["x"])
    assert relerr(pen, g["gp"]) < 1e-4
    pen.backward(); _grads_ok(D, g["d_gp_grads"], 5e-4)


def case_learner_train(golden, dev, fname, model):
    """Learner.train() for two main iterations (D step, G step, fused Adam, EWMA, w_ewma) vs the reference's."""
    from torch.utils.data import BatchSampler, DataLoader, SequentialSampler, TensorDataset
    from gan_lab_b200.progan.learner import ProGANLearner
    from gan_lab_b200.stylegan.learner import StyleGANLearner
    g = _to(golden(fname), dev)
    res, bs, iters = g["res"], g["bs"], g["iters"]
    if model == "StyleGAN":
        cfg = default_config("StyleGAN", res=res, batch_size=bs, dev=dev, len_latent=g["len_latent"],
                             len_dlatent=g["len_latent"], cutoff_trunc_trick=int(math.log2(res)) - 2)
        L = StyleGANLearner(cfg)
    else:
        cfg = default_config("ProGAN", res=res, batch_size=bs, dev=dev, len_latent=g["len_latent"])
        L = ProGANLearner(cfg)
    _load(L.gen_model, g["g_sd0"]); _load(L.disc_model, g["d_sd0"])
    _load(L.gen_model_lagged, g["g_sd0"])
    ds = TensorDataset(g["data"])
    dl = DataLoader(ds, batch_sampler=BatchSampler(SequentialSampler(ds), batch_size=bs, drop_last=True))
    set_random_source(TapeSource(g["tape"], dev))
    losses = []
    orig_d, orig_g = L.disc_step, L.gen_step
    L.disc_step = lambda xb: losses.append(float(orig_d(xb))) or torch.tensor(losses[-1])
    L.gen_step = lambda: losses.append(float(orig_g())) or torch.tensor(losses[-1])
    L.train(dl, num_main_iters=iters)
    # the first loss is a pure function of the inputs; later ones sit behind Adam(beta1=0) sign-like updates, where
    # a rounding-level gradient difference moves a parameter by 2*lr -> allow 1e-3 there.
    for i, (a, b) in enumerate(zip(losses, g["losses"])):
        assert abs(a - b) < (1e-4 if i == 0 else 1e-3) * max(1.0, abs(b)), (losses, g["losses"])

    def adam_close(mine, ref, what):         # scattered sign flips, never concentrated; exact per-step Adam bound
        _adam_close(mine, ref, g["lr"], iters, what, frac=0.05)     # (measured on the doubles: <= 1.5 %; a wiring error gives > 50 %)

    adam_close(L.gen_model.state_dict(), g["g_sd1"], "G")
    adam_close(L.disc_model.state_dict(), g["d_sd1"], "D")
    adam_close(dict(L.gen_model_lagged.named_parameters()), g["lagged"], "EWMA-G")
    assert abs(L.beta - g["beta"]) < 1e-12
    if model == "StyleGAN":
        close(L.gen_model.w_ewma, g["w_ewma"], rtol=1e-3, atol=1e-5)


class ReplayLoader(object):
    """Feeds a learner the real samples the reference's own loader served (recorded in the fixture), in order, `bs` at a
    time; the learner may change `batch_sampler.batch_size` and call `set_resolution` as it grows."""

    def __init__(self, served, bs, dev):
        self.dataset = [x for _, x in served]
        self.batch_sampler = type("_BS", (), {"batch_size": bs})()
        self.cursor, self.dev, self.resolutions = 0, dev, []

    def set_resolution(self, res):
        self.resolutions.append(res)

    def __iter__(self):
        while self.cursor < len(self.dataset):
            n = self.batch_sampler.batch_size
            xs = self.dataset[self.cursor:self.cursor + n]
            self.cursor += n
            yield (torch.stack(xs).to(self.dev),)


def _grow_learner(g, dev, model):
    from gan_lab_b200.progan.learner import ProGANLearner
    from gan_lab_b200.stylegan.learner import StyleGANLearner
    kw = dict(res=g["res"], init_res=g["init_res"], batch_size=g["bs_dict"][g["init_res"]], dev=dev,
              len_latent=g["len_latent"], bs_dict=dict(g["bs_dict"]), nimg_transition=g["nimg_transition"],
              lr_fctr_dict=dict(g["lr_fctr_dict"]), res_dataset=g["data_res"])
    with _fmap_base(g):
        if model == "StyleGAN":
            return StyleGANLearner(default_config("StyleGAN", len_dlatent=g["len_latent"],
                                                  cutoff_trunc_trick=int(math.log2(g["res"])) - 2, **kw))
        return ProGANLearner(default_config("ProGAN", **kw))


def _adam_close(mine, ref, lr, steps, what, frac=0.03, big=None):
    """Parameters behind `steps` Adam(beta1=0) updates: a rounding-level gradient difference flips a sign-like update, so allow
    scattered flips (never more than `frac` of a network).  Hard bound per element: step t of Adam(beta2=.99) moves a parameter by
    at most lr*sqrt((1-.99^t)/.01) (a gradient much larger than its own history; typical for gradients that are pure rounding
    noise, e.g. a bias in front of an InstanceNorm), and two trajectories can move apart by twice that."""
    bound = 2.001 * lr * sum(math.sqrt((1. - .99 ** t) / .01) for t in range(1, steps + 1)) + 1e-6
    bad = tot = 0
    per_key = {}
    for k, v in ref.items():
        d = (mine[k].detach() - v).abs()
        assert float(d.max()) <= bound, (what, k, float(d.max()), bound)
        # big: count only elements that moved apart by more than `big` (a fraction of a step): where the gradient itself carries
        # per-cent noise (generator gradients behind the discriminator's fresh Adam step) every element differs a little
        nb = int((d > (2e-5 + 1e-4 * v.abs() if big is None else big)).sum())
        if nb:
            per_key[k] = (nb, v.numel(), float(d.max()))
        bad += nb; tot += v.numel()
    worst = sorted(per_key.items(), key=lambda kv: -kv[1][0])[:8]
    assert bad <= frac * tot, (what, bad, tot, worst)


def case_learner_grow(golden, dev, fname, model, device_alpha=False):
    """Learner.train() through a resolution increase vs the reference: phase bookkeeping, the
    optimiser / scheduler / EWMA rebuild at each phase change, the moving alpha incl. the real-image blend, the final
    phase.  The new block's initial values are taken from the reference (they are not taped draws); everything carried
    over (old blocks, torgb -> prev_torgb, lagged generator) is the learner's own."""
    g = _to(golden(fname), dev)
    L = _grow_learner(g, dev, model)
    if device_alpha:         # the blends read alpha from a device vector (what lets fade-in phases replay as CUDA graphs)
        L.enable_cuda_graphs(False, device_alpha=True)
        assert L.state.alpha_dev is not None
    _load(L.gen_model, g["g_sd0"]); _load(L.disc_model, g["d_sd0"]); _load(L.gen_model_lagged, g["g_sd0"])
    snaps = {"g": [sd for t, sd in g["after_inc"] if t == "g"], "d": [sd for t, sd in g["after_inc"] if t == "d"]}
    lr_max = g["lr_base"] * max(g["lr_fctr_dict"][r] for r in (g["init_res"], g["res"]))
    steps_so_far = [0]
    # one Adam(beta2=.99) step moves an element by at most lr*sqrt((1-.99^t)/.01) <= lr*t, t = steps since the optimiser was
    # (re)built at the last phase change; phases are nimg_transition/bs iterations long
    per = g["nimg_transition"] // min(g["bs_dict"][g["init_res"]], g["bs_dict"][g["res"]])
    for net, tag, fresh_prefix in ((L.gen_model, "g", ("torgb.",)), (L.disc_model, "d", ("fromrgb.",))):
        orig = net.increase_scale

        def wrapped(orig=orig, net=net, tag=tag, fresh_prefix=fresh_prefix):
            alive = list(net.parameters())                 # keeps the old objects alive: ids of collected ones get reused
            before = {id(p) for p in alive}
            old_prev = sum(1 for n, _ in net.named_parameters() if n.startswith("prev_"))   # dropped by this increase
            orig()
            snap = snaps[tag].pop(0)
            named = dict(net.named_parameters())
            assert set(named.keys()) == set(snap.keys()), (set(named.keys()) ^ set(snap.keys()))
            # prev_torgb / prev_fromrgb are new Parameter objects holding the old torgb / fromrgb VALUES: carried, not fresh
            fresh = [k for k, p in named.items() if id(p) not in before and not k.startswith("prev_")]
            assert fresh and any(k.startswith(fresh_prefix) for k in fresh), fresh
            with torch.no_grad():
                for k in fresh:                                # freshly initialised by increase_scale()
                    assert named[k].shape == snap[k].shape, k
                    named[k].copy_(snap[k])
            # what the learner carried over (old blocks, re-indexed; torgb -> prev_torgb) must be what the reference carried
            carried = {k: v for k, v in snap.items() if k not in fresh}
            assert len(carried) == len(before) - old_prev, (len(carried), len(before), old_prev)
            _adam_close(named, carried, lr_max, per, "carried-" + tag, frac=0.01)
        net.increase_scale = wrapped

    dl = ReplayLoader(g["served"], L.batch_size, dev)
    set_random_source(TapeSource(g["tape"], dev))
    losses, trace = [], []
    orig_d, orig_g = L.disc_step, L.gen_step

    def disc_step(xb):
        # Adam(beta1=0) makes the trajectory chaotic at fp32 rounding level (the reference re-run from parameters perturbed by
        # 3e-7 drifts from itself faster than this path drifts from it), so every iteration is checked on its own: compare
        # with the reference's parameters at the start of the iteration, then continue FROM the reference's parameters.
        snap = g["iter_snaps"].get(len(trace))
        if snap is not None:
            # The generator step of the FIRST iteration after an optimiser rebuild is a pure sign step (Adam, t = 1) on gradients
            # taken behind the discriminator's own sign step of the same iteration: the unmodified reference re-run with its
            # weights moved by one ulp differs from itself by ~5 % (L2) in exactly these gradients at cfg2's size
            # (tests/golden/style_cfg2_fullwidth_step.pt, `self_noise`), i.e. a few per cent of sign flips.  Elsewhere 1 %.
            # From the second step on the update lr * g2 / sqrt((g1^2 + g2^2) / 2) is smooth in the gradient, so that per-cent
            # gradient noise moves EVERY element by a few per cent of a step: only elements a quarter of a step apart count.
            # Iterations run since the last re-synchronisation with the reference's parameters widen the gap further -- for the
            # discriminator as well, whose step then sits behind a generator that has already drifted.
            prev = max([k for k in g["iter_snaps"] if k < len(trace)], default=0)
            gap = max(1, len(trace) - prev)
            _adam_close(L.gen_model.state_dict(), snap[0], lr_max, per, "G@%d" % len(trace), frac=0.05 * gap, big=0.25 * lr_max)
            if gap == 1:
                _adam_close(L.disc_model.state_dict(), snap[1], lr_max, per, "D@%d" % len(trace), frac=0.01)
            else:
                _adam_close(L.disc_model.state_dict(), snap[1], lr_max, per, "D@%d" % len(trace), frac=0.05 * gap, big=0.25 * lr_max)
            with torch.no_grad():
                for net, sd in ((L.gen_model, snap[0]), (L.disc_model, snap[1])):
                    for k, v in net.state_dict().items():
                        v.copy_(sd[k])
            K.weights_updated()
        trace.append(dict(res=int(L.gen_model.curr_res), fade=bool(L.gen_model.fade_in_phase), alpha=float(L.gen_model.alpha),
                          bs=int(L.batch_size), phase=int(L.curr_phase_num), lr_d=float(L.opt_disc.param_groups[0]["lr"]),
                          lr_g=float(L.opt_gen.param_groups[0]["lr"]), beta=float(L.beta), img_num=int(L.curr_img_num)))
        losses.append(float(orig_d(xb)))
        steps_so_far[0] += 1
        return torch.tensor(losses[-1])

    L.disc_step = disc_step
    L.gen_step = lambda: losses.append(float(orig_g())) or torch.tensor(losses[-1])
    L.train(dl, num_main_iters=g["iters"])

    assert len(trace) == len(g["trace"])
    for mine, ref in zip(trace, g["trace"]):
        for k, v in ref.items():
            assert (abs(mine[k] - v) < 1e-12) if isinstance(v, float) else (mine[k] == v), (k, mine, ref)
    fin = g["final"]
    G = L.gen_model
    assert (int(G.curr_res), bool(G.fade_in_phase), float(G.alpha), L.curr_phase_num, L.curr_img_num, L.batch_size) == \
           (fin["res"], fin["fade"], fin["alpha"], fin["phase"], fin["img_num"], fin["bs"])
    assert [float(v) for v in L.nimg_transition_lst] == fin["nimg_transition_lst"]
    assert bool(L._progressively_grow) == fin["progressively_grow"]
    assert dl.resolutions == [g["res"]]
    # losses: the first is a pure function of the inputs; later ones sit behind Adam's sign-like updates
    assert len(losses) == len(g["losses"])
    for i, (a, b) in enumerate(zip(losses, g["losses"])):
        forced = (i // 2) in g["iter_snaps"] or i < 2        # this iteration started from the reference's own parameters
        assert abs(a - b) < (1e-4 if i == 0 else (5e-4 if forced else 3e-3)) * max(1.0, abs(b)), (i, a, b)
    free = g["iters"] - max(g["iter_snaps"])           # iterations run since the last re-synchronisation with the reference
    if free <= 1:
        _adam_close(L.gen_model.state_dict(), g["g_sd1"], lr_max, per, "G", frac=0.05, big=0.25 * lr_max)
        _adam_close(L.disc_model.state_dict(), g["d_sd1"], lr_max, per, "D", frac=0.01)
    else:       # free-running iterations: see the per-iteration checks above
        _adam_close(L.gen_model.state_dict(), g["g_sd1"], lr_max, per, "G", frac=0.05 * free, big=0.25 * lr_max)
        _adam_close(L.disc_model.state_dict(), g["d_sd1"], lr_max, per, "D", frac=0.05 * free, big=0.25 * lr_max)
    lag = dict(L.gen_model_lagged.named_parameters())
    assert set(lag.keys()) == set(g["lagged"].keys())
    _adam_close(lag, g["lagged"], lr_max, g["iters"], "EWMA-G", frac=0.01)
    if model == "StyleGAN":
        close(L.gen_model.w_ewma, g["w_ewma"], rtol=1e-3, atol=1e-4)


def case_learner_resume(golden, dev, fname, model, golden_dir):
    """load_model() of a checkpoint file written by the UNMODIFIED reference in the middle of a fade-in, then train() on:
    vs the reference resuming from the same file."""
    g = _to(golden(fname), dev)
    L = _grow_learner(g, dev, model)                      # any learner of the family; load_model rebuilds everything
    L.load_model(golden_dir / g["checkpoint"], dev_of_saved_model="cpu", dev=dev)
    sv = g["saved"]
    G = L.gen_model
    assert (int(G.curr_res), bool(G.fade_in_phase), float(G.alpha), L.curr_img_num, L.curr_phase_num) == \
           (sv["res"], sv["fade"], sv["alpha"], sv["img_num"], sv["phase"])
    assert L.pretrained_model and not L.not_trained_yet
    for mine, ref in ((L.gen_model.state_dict(), sv["g_sd"]), (L.disc_model.state_dict(), sv["d_sd"]),
                      (dict(L.gen_model_lagged.named_parameters()), sv["lagged"])):
        assert set(mine.keys()) == set(ref.keys())
        for k, v in ref.items():
            assert torch.equal(mine[k].detach(), v), k
    # the reference rebuilds its optimisers before every save (progan/learner.py:963,1018): the stored Adam state is empty, and
    # the optimisers load_model() builds for the stored phase (mid-fade-in: prev_torgb / prev_fromrgb included) start fresh
    for opt, net in ((L.opt_gen, L.gen_model), (L.opt_disc, L.disc_model)):
        assert not opt.state_dict()["state"]
        assert len(opt.param_groups[0]["params"]) == len(list(net.parameters()))
    lr_max = g["lr_base"] * max(g["lr_fctr_dict"][r] for r in (g["init_res"], g["res"]))
    dl = ReplayLoader(g["served"], 1, dev)                # train() must set the loader's batch size itself on a resume
    set_random_source(TapeSource(g["tape"], dev))
    losses, trace = [], []
    orig_d, orig_g = L.disc_step, L.gen_step

    per = g["nimg_transition"] // g["bs_dict"][g["res"]]

    def disc_step(xb):
        snap = g["iter_snaps"].get(len(trace))        # every iteration on its own, continued from the reference's parameters
        if snap is not None:
            _adam_close(L.gen_model.state_dict(), snap[0], lr_max, per, "G@%d" % len(trace), frac=0.01)
            _adam_close(L.disc_model.state_dict(), snap[1], lr_max, per, "D@%d" % len(trace), frac=0.01)
            with torch.no_grad():
                for net, sd in ((L.gen_model, snap[0]), (L.disc_model, snap[1])):
                    for k, v in net.state_dict().items():
                        v.copy_(sd[k])
            K.weights_updated()
        trace.append(dict(res=int(L.gen_model.curr_res), fade=bool(L.gen_model.fade_in_phase), alpha=float(L.gen_model.alpha),
                          bs=int(L.batch_size), phase=int(L.curr_phase_num), lr_d=float(L.opt_disc.param_groups[0]["lr"]),
                          lr_g=float(L.opt_gen.param_groups[0]["lr"]), beta=float(L.beta), img_num=int(L.curr_img_num)))
        losses.append(float(orig_d(xb)))
        return torch.tensor(losses[-1])

    L.disc_step = disc_step
    L.gen_step = lambda: losses.append(float(orig_g())) or torch.tensor(losses[-1])
    L.train(dl, num_main_iters=g["iters_after"])
    assert dl.resolutions == [g["res"]] and dl.batch_sampler.batch_size == g["bs_dict"][g["res"]]
    assert len(trace) == len(g["trace"])
    for mine, ref in zip(trace, g["trace"]):
        for k, v in ref.items():
            assert (abs(mine[k] - v) < 1e-12) if isinstance(v, float) else (mine[k] == v), (k, mine, ref)
    fin = g["final"]
    assert (int(G.curr_res), bool(G.fade_in_phase), float(G.alpha), L.curr_phase_num, L.curr_img_num, L.batch_size) == \
           (fin["res"], fin["fade"], fin["alpha"], fin["phase"], fin["img_num"], fin["bs"])
    assert [float(v) for v in L.nimg_transition_lst] == fin["nimg_transition_lst"]
    for i, (a, b) in enumerate(zip(losses, g["losses"])):
        assert abs(a - b) < (1e-4 if i == 0 else 5e-4) * max(1.0, abs(b)), (i, a, b)
    _adam_close(L.gen_model.state_dict(), g["g_sd1"], lr_max, per, "G", frac=0.01)
    _adam_close(L.disc_model.state_dict(), g["d_sd1"], lr_max, per, "D", frac=0.01)
    # the EWMA generator restarted from the live one (reference progan/learner.py:462-472)
    _adam_close(dict(L.gen_model_lagged.named_parameters()), g["lagged"], lr_max, per, "EWMA-G", frac=0.02)
    if model == "StyleGAN":
        close(L.gen_model.w_ewma, g["w_ewma"], rtol=1e-3, atol=1e-4)
    return L


def case_checkpoint_roundtrip(golden, dev, fname, model, tmp_path):
    """save_model() -> load_model() of this package's own file: every tensor and every bookkeeping entry survives, and the
    resumed learner steps exactly like the one that kept running would after the same end-of-train() optimiser rebuild."""
    g = _to(golden(fname), dev)
    L = _grow_learner(g, dev, model)
    _load(L.gen_model, g["g_sd0"]); _load(L.disc_model, g["d_sd0"]); _load(L.gen_model_lagged, g["g_sd0"])
    dl = ReplayLoader(g["served"], L.batch_size, dev)
    L.train(dl, num_main_iters=6)                         # 4 stabilising iterations at 4x4, 2 fading 8x8 in
    path = tmp_path / "model.tar"
    L.save_model(path)
    L2 = _grow_learner(g, dev, model)
    with _fmap_base(g):              # (a module constant on both sides, stylegan/base.py:16: checkpoints do not carry it)
        L2.load_model(path, dev_of_saved_model="cpu", dev=dev)
    for a, b in ((L.gen_model, L2.gen_model), (L.disc_model, L2.disc_model), (L.gen_model_lagged, L2.gen_model_lagged)):
        sa, sb = a.state_dict(), b.state_dict()
        assert list(sa.keys()) == list(sb.keys())
        for k in sa:
            assert torch.equal(sa[k], sb[k]) and sa[k].stride() == sb[k].stride(), k
    for attr in ("batch_size", "curr_img_num", "curr_phase_num", "curr_epoch_num", "curr_dataset_batch_num", "loss",
                 "gradient_penalty", "optimizer", "lr_sched", "latent_distribution", "not_trained_yet", "sched_bool"):
        assert getattr(L, attr) == getattr(L2, attr), attr
    assert [float(v) for v in L.nimg_transition_lst] == [float(v) for v in L2.nimg_transition_lst]
    G, G2 = L.gen_model, L2.gen_model
    assert (G.curr_res, G.fade_in_phase, G.alpha) == (G2.curr_res, G2.fade_in_phase, G2.alpha)
    assert vars(L.config).keys() == vars(L2.config).keys()
    for k, v in vars(L.config).items():
        assert vars(L2.config)[k] == v, k
    if model == "StyleGAN":
        assert torch.equal(G.w_ewma, G2.w_ewma) and G.w_ewma_beta == G2.w_ewma_beta


def case_compute_metrics(golden, dev, fname, model, tmp_path):
    """Learner.compute_metrics() vs the reference's (progan/learner.py:248-416): generator metrics on a latent validation set
    with a short last batch, discriminator metrics on latents + reals; eval-mode generator (truncation trick, fresh noise)."""
    from torch.utils.data import BatchSampler, DataLoader, SequentialSampler, TensorDataset
    from gan_lab_b200.progan.learner import ProGANLearner
    from gan_lab_b200.stylegan.learner import StyleGANLearner
    g = _to(golden(fname), dev)
    res, bs = g["res"], g["bs"]
    kw = dict(res=res, batch_size=bs, dev=dev, len_latent=g["len_latent"], save_samples_dir=tmp_path, img_grid_sz=2)
    if model == "StyleGAN":
        L = StyleGANLearner(default_config("StyleGAN", len_dlatent=g["len_latent"], cutoff_trunc_trick=int(math.log2(res)) - 2, **kw))
        L.gen_model.w_ewma = g["w_ewma"].clone()
    else:
        L = ProGANLearner(default_config("ProGAN", **kw))
    _load(L.gen_model, g["g_sd"]); _load(L.disc_model, g["d_sd"]); _load(L.gen_model_lagged, g["g_sd"])
    L.gen_model.train(); L.disc_model.train()
    zds, xds = TensorDataset(g["z_valid"]), TensorDataset(g["x_valid"])
    z_dl = DataLoader(zds, batch_sampler=BatchSampler(SequentialSampler(zds), batch_size=bs, drop_last=False))
    x_dl = DataLoader(xds, batch_sampler=BatchSampler(SequentialSampler(xds), batch_size=bs, drop_last=False))

    def check(lines, ref_lines, raw, names):
        assert len(lines) == len(ref_lines)
        for name, want in zip(names, raw):
            got = L.last_metrics[name]
            assert abs(got - want) < 2e-4 * max(1.0, abs(want)), (name, got, want)
        for a, b in zip(lines, ref_lines):            # same layout; the %.4g figure may differ in its last digit
            assert a.split(":")[0] == b.split(":")[0] and a.endswith("\n")
            assert abs(float(a.split(":")[1]) - float(b.split(":")[1])) <= 2e-3 * max(1.0, abs(float(b.split(":")[1])))

    set_random_source(TapeSource(g["tape_g"], dev))
    lines = L.compute_metrics(metrics=g["gen_metrics"], metrics_type="Generator", z_valid_dl=z_dl, valid_dl=None)
    check(lines, g["vals_g"], g["raw_g"], g["gen_metrics"])
    set_random_source(TapeSource(g["tape_d"], dev))
    lines = L.compute_metrics(metrics=g["disc_metrics"], metrics_type="Discriminator", z_valid_dl=z_dl, valid_dl=x_dl)
    check(lines, g["vals_d"], g["raw_d"], g["disc_metrics"])
    assert (L.gen_metrics_num, L.disc_metrics_num) == (g["gen_metrics_num"], g["disc_metrics_num"])
    assert (L.gen_model.training, L.disc_model.training) == g["modes"]
    # image grid: 2x2 samples from both generators, written as PNGs where the reference writes them
    set_random_source(None)
    L.compute_metrics(metrics=["image grid"], metrics_type="Generator", z_valid_dl=z_dl)
    assert L.grid_inputs_constructed and L.valid_z.shape == (4, g["len_latent"])
    base = tmp_path / model.casefold() / "image_grid"
    from PIL import Image
    for sub in ("original", "time_averaged"):
        im = Image.open(base / sub / (str(g["gen_metrics_num"]) + ".png"))
        assert im.size == (2 * res, 2 * res) and im.mode == "RGB"


def case_train_variant(golden, dev, name, monkeypatch):
    """Learner.train() under configuration switches the headline fixtures leave at their defaults (losses, penalties, gamma,
    several D / G steps per iteration, gen_bs_mult, noise / InstanceNorm / PixelNorm / mixing / EWMA off, uniform latents, LR
    schedules, poolers, blur off, equalized LR off, ReLU, Adam beta1 / weight decay) vs the unmodified reference."""
    from torch.utils.data import BatchSampler, DataLoader, SequentialSampler, TensorDataset
    import gan_lab_b200._growth as growth
    from gan_lab_b200.progan.learner import ProGANLearner
    from gan_lab_b200.stylegan.learner import StyleGANLearner
    g = _to(golden("train_variants.pt")[name], dev)
    monkeypatch.setattr(growth, "FMAP_MAX", g["fmap_max"])
    res, bs, iters, over = g["res"], g["bs"], g["iters"], dict(g["over"])
    if g["model"] == "StyleGAN":
        over.setdefault("cutoff_trunc_trick", int(math.log2(res)) - 2)
        L = StyleGANLearner(default_config("StyleGAN", res=res, batch_size=bs, dev=dev, len_latent=g["len_latent"],
                                           len_dlatent=g["len_latent"], **over))
    else:
        L = ProGANLearner(default_config("ProGAN", res=res, batch_size=bs, dev=dev, len_latent=g["len_latent"], **over))
    _load(L.gen_model, g["g_sd0"]); _load(L.disc_model, g["d_sd0"])
    if L.gen_model_lagged is not None:
        _load(L.gen_model_lagged, g["g_sd0"])
    ds = TensorDataset(g["data"])
    dl = DataLoader(ds, batch_sampler=BatchSampler(SequentialSampler(ds), batch_size=bs, drop_last=True))
    set_random_source(TapeSource(g["tape"], dev))
    losses, lrs = [], []
    orig_d, orig_g = L.disc_step, L.gen_step

    per_iter = L.config.num_disc_iters + L.config.num_gen_iters

    def rec(fn, *a):
        # every main iteration is checked on its own and continued from the reference's parameters (see case_learner_grow)
        snap = g["iter_snaps"].get(len(losses) // per_iter) if len(losses) % per_iter == 0 else None
        if snap is not None:
            _adam_close(L.gen_model.state_dict(), snap[0], g["lr"], per_iter, "G@%d" % len(losses), frac=0.02)
            _adam_close(L.disc_model.state_dict(), snap[1], g["lr"], per_iter, "D@%d" % len(losses), frac=0.02)
            with torch.no_grad():
                for net, sd in ((L.gen_model, snap[0]), (L.disc_model, snap[1])):
                    for k, v in net.state_dict().items():
                        v.copy_(sd[k])
            K.weights_updated()
        lrs.append((float(L.opt_disc.param_groups[0]["lr"]), float(L.opt_gen.param_groups[0]["lr"])))
        losses.append(float(fn(*a)))
        return torch.tensor(losses[-1])

    L.disc_step = lambda xb: rec(orig_d, xb)
    L.gen_step = lambda: rec(orig_g)
    L.train(dl, num_main_iters=iters)
    assert len(losses) == len(g["losses"]), (len(losses), len(g["losses"]))
    for mine, ref in zip(lrs, g["lrs"]):
        assert abs(mine[0] - ref[0]) < 1e-12 and abs(mine[1] - ref[1]) < 1e-12, (lrs, g["lrs"])
    for i, (a, b) in enumerate(zip(losses, g["losses"])):
        assert abs(a - b) < (1e-4 if i == 0 else 5e-4) * max(1.0, abs(b)), (name, i, a, b)
    steps = iters * max(L.config.num_disc_iters, L.config.num_gen_iters)
    _adam_close(L.gen_model.state_dict(), g["g_sd1"], g["lr"], steps, "G", frac=0.02)
    _adam_close(L.disc_model.state_dict(), g["d_sd1"], g["lr"], steps, "D", frac=0.02)
    if g["lagged"] is not None:
        _adam_close(dict(L.gen_model_lagged.named_parameters()), g["lagged"], g["lr"], steps, "EWMA-G", frac=0.02)
    else:
        assert L.gen_model_lagged is None


def case_shared_penalty_forward(dev, gp):
    """disc_step with the R1/R2 penalty riding on the loss's own D forward == the reference's literal order (a separate
    D forward inside calc_gp), on identical draws; lda inflated so the penalty's gradient is visible in the total."""
    from gan_lab_b200.stylegan.learner import StyleGANLearner
    from gan_lab_b200.optim import FusedAdam
    grads = {}
    for share in (False, True):
        torch.manual_seed(5)
        cfg = default_config("StyleGAN", res=16, batch_size=4, dev=dev, len_latent=32, len_dlatent=32,
                             cutoff_trunc_trick=2, lda=1.e4, gradient_penalty=gp, pct_mixing_reg=0., use_ewma_gen=False)
        L = StyleGANLearner(cfg)
        gen = torch.Generator().manual_seed(6)
        with torch.no_grad():
            for m in (L.gen_model, L.disc_model):
                for prm in m.parameters():
                    if float(prm.abs().max()) == 0.0:
                        prm.copy_((torch.randn(prm.shape, generator=gen) * 0.3).to(dev))
        L.gen_model.train(); L.disc_model.train()
        L.share_penalty_forward = share
        L.opt_disc.step = lambda: None                  # keep the gradients: compare them, not the Adam result
        x = (torch.rand(4, 3, 16, 16, generator=gen) * 2 - 1).to(dev)
        torch.manual_seed(7)
        loss = L.disc_step(x)
        grads[share] = (float(loss), {n: prm.grad.detach().clone() for n, prm in L.disc_model.named_parameters()
                                      if prm.grad is not None})
    (l0, g0), (l1, g1) = grads[False], grads[True]
    assert abs(l0 - l1) <= 1e-5 * max(1.0, abs(l0)), (l0, l1)
    assert set(g0) == set(g1) and len(g0) > 10
    for n in g0:
        assert relerr(g1[n], g0[n]) < 2e-4, (n, relerr(g1[n], g0[n]))


def case_batched_d_passes(dev, gp, bs=4):
    """disc_step with D(fake) and D(real) evaluated as one pass over the concatenated batch == two separate passes (loss and every
    discriminator gradient), with the R1 / R2 penalty riding on its half of the batch; lda inflated so the penalty is visible.
    With a batch smaller than the minibatch-stddev group the learner must fall back to separate passes by itself."""
    from gan_lab_b200.stylegan.learner import StyleGANLearner
    grads = {}
    for batched in (False, True):
        torch.manual_seed(5)
        cfg = default_config("StyleGAN", res=16, batch_size=bs, dev=dev, len_latent=32, len_dlatent=32,
                             cutoff_trunc_trick=2, lda=1.e3, gradient_penalty=gp, pct_mixing_reg=0., use_ewma_gen=False)
        L = StyleGANLearner(cfg)
        gen = torch.Generator().manual_seed(6)
        with torch.no_grad():
            for m in (L.gen_model, L.disc_model):
                for prm in m.parameters():
                    if float(prm.abs().max()) == 0.0:
                        prm.copy_((torch.randn(prm.shape, generator=gen) * 0.3).to(dev))
        L.gen_model.train(); L.disc_model.train()
        L.batch_d_passes = batched
        calls = []
        orig_forward = L.disc_model.forward
        L.disc_model.forward = lambda x: calls.append(x.shape[0]) or orig_forward(x)
        L.opt_disc.step = lambda: None
        x = (torch.rand(bs, 3, 16, 16, generator=gen) * 2 - 1).to(dev)
        torch.manual_seed(7)
        loss = L.disc_step(x)
        want_batched = batched and bs % 4 == 0
        assert calls == ([2 * bs] if want_batched else [bs, bs]), calls
        grads[batched] = (float(loss), {n: prm.grad.detach().clone() for n, prm in L.disc_model.named_parameters()
                                        if prm.grad is not None})
    (l0, g0), (l1, g1) = grads[False], grads[True]
    assert abs(l0 - l1) <= 1e-5 * max(1.0, abs(l0)), (l0, l1)
    assert g0.keys() == g1.keys()
    for n in g0:
        assert relerr(g1[n], g0[n]) < 2e-4, (n, relerr(g1[n], g0[n]))


def _resnet_learner(g, dev, **over):
    """GANLearner for the small ResNet fixtures: the reference's FMAP_G / FMAP_D constants (resnetgan/architectures.py:19-20)
    were patched to g['fmap'] when the fixture was made; do the same to our mirror of them."""
    import gan_lab_b200.resnetgan.architectures as RA
    from gan_lab_b200.resnetgan.learner import GANLearner
    old = (RA.FMAP_G, RA.FMAP_D)
    RA.FMAP_G = RA.FMAP_D = g["fmap"]
    try:
        cfg = default_config("ResNet GAN", res=g["res"], batch_size=g["bs"], dev=dev, len_latent=g["len_latent"], **over)
        return GANLearner(cfg), cfg
    finally:
        RA.FMAP_G, RA.FMAP_D = old


def case_resnet_nets_modules(golden, dev, fname, dtype=torch.float32, grad_tol=5e-2, report=None, d_grad_tol=None):
    """ResNet generator (BatchNorm blocks, Tanh) / discriminator (LayerNorm blocks) forward, backward and the WGAN-GP
    double backward vs the reference's modules; also the BatchNorm running buffers after one forward.

    Gradient tolerance: these nets put ~5e5 ReLU inputs behind Batch/LayerNorms, so a fixture always holds pre-activations
    within fp32 rounding of zero; an implementation that rounds differently flips such a mask bit, which moves every
    gradient behind it by ~1e-3..1e-2 of its max-norm (measured: the reference's own fp32 vs fp64).  fp32 runs therefore
    use grad_tol = 5e-2 (forward values stay at 1e-4); the CPU host-wiring test runs the same case in fp64, where no bit
    flips, at 2e-5 -- i.e. down to the fp32 noise of the fixture itself.  Measured on B200 (fp32 kernels): only the GENERATOR
    (BatchNorm + ReLU) shows such flips (3.7e-2 on one skip-connection weight of the 64x64 net, 8e-6 at 32x32); the
    discriminator's gradients, including the WGAN-GP double backward through its LayerNorm blocks, agree to 2e-6 -- the GPU test
    therefore holds them to d_grad_tol = 2e-4."""
    d_grad_tol = grad_tol if d_grad_tol is None else d_grad_tol
    g = _to(golden(fname), dev)
    L, cfg = _resnet_learner(g, dev)
    G, D = L.gen_model, L.disc_model
    assert list(G.state_dict().keys()) == list(g["g_sd"].keys())
    assert list(D.state_dict().keys()) == list(g["d_sd"].keys())
    _load(G, g["g_sd"]); _load(D, g["d_sd"])
    G.to(dtype); D.to(dtype)
    c = lambda t: t.to(dtype)
    ftol = 1e-4 if dtype == torch.float32 else 2e-5
    G.train(); D.train()
    img = G(c(g["z"]))
    # max-norm relative error (Tanh output, |img| <= 1): an element-wise bound near the zero crossings depends on the CPU's
    # summation order, i.e. on the thread count the doubles happen to run with
    assert relerr(img, c(g["img"])) < ftol
    G.zero_grad(); img.backward(c(g["gimg"])); _grads_ok(G, g["g_grads"], grad_tol, skip_cancelled=True, report=report, tag="g_grads")
    for k, v in G.named_buffers():
        if k.endswith("num_batches_tracked"):
            assert int(v) == int(g["g_buffers"][k])
        else:
            close(v, c(g["g_buffers"][k]), rtol=1e-4, atol=1e-6)
    D.zero_grad()
    logits = D(c(g["x"])); close(logits, c(g["logits"]), rtol=ftol, atol=ftol / 10)
    logits.backward(c(g["glog"])); _grads_ok(D, g["d_grads"], d_grad_tol, report=report, tag="d_grads")
    D.zero_grad()
    set_random_source(TapeSource(g["gp_tape"], dev))
    pen = L.calc_gp(c(g["img"]), c(g["x"]))
    assert relerr(pen, c(g["gp"])) < ftol
    pen.backward(); _grads_ok(D, g["d_gp_grads"], max(d_grad_tol, 2e-4), report=report, tag="d_gp_grads")


def case_resnet_train(golden, dev, fname="resnet_train_res64.pt"):
    """GANLearner.train() (ResNet GAN 64x64, WGAN + WGAN-GP, generator step first then 2 discriminator steps; or the 32x32
    variant: non-saturating loss + R1, two generator steps, LeakyReLU, linear-decay LR) for two main iterations vs the
    reference's: every loss, the learning rates, post-Adam parameters and BatchNorm buffers."""
    from torch.utils.data import BatchSampler, DataLoader, SequentialSampler, TensorDataset
    g = _to(golden(fname), dev)
    bs, iters = g["bs"], g["iters"]
    L, cfg = _resnet_learner(g, dev, num_disc_iters=g["num_disc_iters"], lr_base=g["lr"], **g.get("over", {}))
    _load(L.gen_model, g["g_sd0"]); _load(L.disc_model, g["d_sd0"])
    ds = TensorDataset(g["data"])
    dl = DataLoader(ds, batch_sampler=BatchSampler(SequentialSampler(ds), batch_size=bs, drop_last=True))
    set_random_source(TapeSource(g["tape"], dev))
    losses = []
    orig_d, orig_g = L.disc_step, L.gen_step
    lrs = []
    note_lr = lambda: lrs.append((float(L.opt_disc.param_groups[0]["lr"]), float(L.opt_gen.param_groups[0]["lr"])))
    L.disc_step = lambda xb: note_lr() or losses.append(float(orig_d(xb))) or torch.tensor(losses[-1])
    L.gen_step = lambda: note_lr() or losses.append(float(orig_g())) or torch.tensor(losses[-1])
    L.train(dl, num_main_iters=iters)
    assert len(losses) == len(g["losses"])
    for mine, ref in zip(lrs, g.get("lrs", [])):
        assert abs(mine[0] - ref[0]) < 1e-15 and abs(mine[1] - ref[1]) < 1e-15, (lrs, g["lrs"])
    # the first G and D losses are pure functions of the inputs; later ones sit behind Adam(beta1=0) sign-like updates of
    # every parameter (+-lr wherever a gradient is rounding noise) and an unbounded WGAN critic -> 1e-2 there
    for i, (a, b) in enumerate(zip(losses, g["losses"])):
        assert abs(a - b) < (1e-4 if i < 2 else 1e-3) * max(1.0, abs(b)), (losses, g["losses"])

    def adam_close(mine, ref, what, steps):
        bad = tot = 0
        for k, v in ref.items():
            if k.endswith("num_batches_tracked"):
                assert int(mine[k]) == int(v), k
                continue
            d = (mine[k].detach() - v).abs()
            if "running_" in k:
                assert float(d.max()) <= 5e-3 * max(1.0, float(v.abs().max())), (what, k, float(d.max()))
                continue
            bound = 2.001 * g["lr"] * sum(math.sqrt((1. - .9 ** t) / .1) for t in range(1, steps + 1)) + 1e-6   # Adam beta2 = .9
            assert float(d.max()) <= bound, (what, k, float(d.max()), bound)
            bad += int((d > 0.02 * g["lr"] + 2e-7 * v.abs()).sum()); tot += v.numel()
        # scattered +-2*lr sign flips where a gradient is rounding noise (the reference against itself, 1 vs 8 threads: 1.3 % on
        # the 64x64 case; the two-generator-step variant run with another thread count than its fixture: 7 %)
        assert bad <= (0.05 if "over" not in g or not g["over"] else 0.12) * tot, (what, bad, tot)

    adam_close(L.gen_model.state_dict(), g["g_sd1"], "G", iters * cfg.num_gen_iters)
    adam_close(L.disc_model.state_dict(), g["d_sd1"], "D", iters * g["num_disc_iters"])


def case_resnet_resume(golden, dev, golden_dir):
    """GANLearner.load_model() of a checkpoint the UNMODIFIED reference wrote after one ResNet-GAN iteration -- networks, BatchNorm
    buffers and torch.optim.Adam state (step 2, both moments) -- then one more iteration vs the reference resuming from the file."""
    from torch.utils.data import BatchSampler, DataLoader, SequentialSampler, TensorDataset
    g = _to(golden("resnet_resume_res32.pt"), dev)
    L, _ = _resnet_learner(g, dev)
    L.load_model(golden_dir / g["checkpoint"], dev_of_saved_model="cpu", dev=dev)
    assert (L.config.num_disc_iters, L.config.lr_base, L.curr_img_num, L.batch_size) == (g["num_disc_iters"], g["lr"],
                                                                                          g["saved"]["img_num"], g["bs"])
    for mine, ref in ((L.gen_model.state_dict(), g["saved"]["g_sd"]), (L.disc_model.state_dict(), g["saved"]["d_sd"])):
        assert list(mine.keys()) == list(ref.keys())
        for k, v in ref.items():
            assert torch.equal(mine[k], v), k
    assert sorted({float(st["step"]) for st in L.opt_disc.state_dict()["state"].values()}) == g["saved"]["opt_disc_steps"]
    for p in L.opt_disc.param_groups[0]["params"]:
        st = L.opt_disc.state[p]
        assert st["exp_avg_sq"].stride() == p.stride() and st["exp_avg_sq"].device == p.device
    ds = TensorDataset(g["data"])
    dl = DataLoader(ds, batch_sampler=BatchSampler(SequentialSampler(ds), batch_size=g["bs"], drop_last=True))
    set_random_source(TapeSource(g["tape"], dev))
    losses = []
    orig_d, orig_g = L.disc_step, L.gen_step
    L.disc_step = lambda xb: losses.append(float(orig_d(xb))) or torch.tensor(losses[-1])
    L.gen_step = lambda: losses.append(float(orig_g())) or torch.tensor(losses[-1])
    L.train(dl, num_main_iters=1)
    assert L.curr_img_num == g["img_num"]
    assert sorted({float(st["step"]) for st in L.opt_disc.state_dict()["state"].values()}) == g["opt_disc_steps"]
    for i, (a, b) in enumerate(zip(losses, g["losses"])):
        assert abs(a - b) < (1e-4 if i == 0 else 1e-3) * max(1.0, abs(b)), (losses, g["losses"])
    # steps 3 and 4 of the loaded Adam state: a wrong step count or lost moments would move every element by a different amount
    for mine, ref, steps in ((L.gen_model.state_dict(), g["g_sd1"], 1), (L.disc_model.state_dict(), g["d_sd1"], g["num_disc_iters"])):
        bad = tot = 0
        for k, v in ref.items():
            if not v.is_floating_point() or "running_" in k:
                continue
            d = (mine[k].detach() - v).abs()
            assert float(d.max()) <= 4.0 * g["lr"] * steps + 1e-7, (k, float(d.max()))
            bad += int((d > 0.02 * g["lr"] + 2e-7 * v.abs()).sum()); tot += v.numel()
        assert bad <= 0.05 * tot, (bad, tot)


def case_style_eval(golden, dev):
    """Evaluation-mode StyleGenerator vs the reference's: supplied noise, truncation trick (psi, cut-off stage and the
    de-truncation above it), style mixing below and above the cut-off, and truncation switched off."""
    from gan_lab_b200.stylegan.learner import StyleGANLearner
    g = _to(golden("style_eval_res32.pt"), dev)
    cfg = default_config("StyleGAN", res=g["res"], batch_size=g["bs"], dev=dev, len_latent=g["len_latent"],
                         len_dlatent=g["len_latent"], cutoff_trunc_trick=g["cutoff"], psi_trunc_trick=g["psi"])
    L = StyleGANLearner(cfg)
    G = L.gen_model
    _load(G, g["g_sd"])
    G.w_ewma = g["w_ewma"].clone()
    G.eval()
    with torch.no_grad():
        assert relerr(G(g["z"], noise=g["noise"]), g["img_trunc"]) < 1e-4
        assert relerr(G(g["z"], x_mixing=g["z_mix"], style_mixing_stage=2, noise=g["noise"]), g["img_mix_low"]) < 1e-4
        assert relerr(G(g["z"], x_mixing=g["z_mix"], style_mixing_stage=5, noise=g["noise"]), g["img_mix_high"]) < 1e-4
        G.trunc_cutoff_stage = None
        assert relerr(G(g["z"], noise=g["noise"]), g["img_notrunc"]) < 1e-4
    assert float((g["img_trunc"] - g["img_notrunc"]).abs().max()) > 1e-3        # the truncation trick did something


# ------------------------------------------------------------------------------------------------ cfg2 at its real widths
def _perturb_zero_params(module, gen):
    """Same rule as oracle.make_golden.perturb_zero_params (zero-initialised parameters get 0.3 * N(0,1) from `gen`)."""
    with torch.no_grad():
        for _name, p in module.named_parameters():
            if float(p.abs().max()) == 0.0:
                p.copy_(torch.randn(p.shape, generator=gen) * 0.3)


class _CapturableTape(TapeSource):
    """Taped draws handed out as CLONES made on the current (capturing) stream: a leaf tensor that was created on the legacy
    default stream would make autograd synchronise that stream with the capture."""

    def randn(self, shape, device):
        return TapeSource.randn(self, shape, device).clone()

    def rand(self, shape, device):
        return TapeSource.rand(self, shape, device).clone()


def cfg2_fullwidth_learner(g, dev):
    """The drop-in StyleGANLearner at cfg2's real size rebuilt from the fixture's seeds: bit-identical initial weights and real
    batch (sha256 digests asserted), without the fixture having to carry 49 M parameters."""
    import hashlib
    import numpy as np
    from gan_lab_b200.stylegan.learner import StyleGANLearner
    dig = lambda t: hashlib.sha256(t.detach().contiguous().cpu().numpy().tobytes()).hexdigest()
    torch.manual_seed(g["seed"]); np.random.seed(g["seed"])
    L = StyleGANLearner(default_config("StyleGAN", res=g["res"], init_res=g["res"], batch_size=g["bs"], dev=dev))
    gen = torch.Generator().manual_seed(g["seed"] + 1)
    _perturb_zero_params(L.gen_model, gen); _perturb_zero_params(L.disc_model, gen)
    for net, want in ((L.gen_model, g["g_digests"]), (L.disc_model, g["d_digests"])):
        sd = net.state_dict()
        assert list(sd.keys()) == list(want.keys())
        for k, v in sd.items():
            assert dig(v) == want[k], k
    data = torch.rand(g["bs"], 3, g["res"], g["res"], generator=gen) * 2 - 1
    assert dig(data) == g["data_digest"]
    return L, data


def case_cfg2_fullwidth_step(golden, dev, impl, graph):
    """ONE main iteration of cfg2 (StyleGAN 128x128, 512/256/128 channels, batch 8, nonsaturating + R1 + drift, noise, mixing)
    -- the configuration bench.py times -- against the UNMODIFIED reference's (tests/golden/style_cfg2_fullwidth_step.pt):
    D loss, G loss, the R1 penalty on its own (value + every D gradient of its double backward, through the shared-forward
    path the step uses), every parameter gradient of the D step and of the G step, the post-Adam parameters, the EWMA generator
    and w_ewma.  impl: 'fp32' | 'tf32' (the benchmarked conv path); graph: run the step as a captured-and-replayed CUDA graph
    pair, as bench.py does.  Returns the worst error per class (the callers assert their bounds on it)."""
    from oracle import summaries as S
    g = golden("style_cfg2_fullwidth_step.pt")
    L, data = cfg2_fullwidth_learner(g, dev)
    G, D = L.gen_model, L.disc_model
    x = data.to(dev)
    on_dev = lambda events: [(k, v.to(dev) if (torch.is_tensor(v) and k != "randint") else v) for k, v in events]   # the mixing cut-off is a host draw
    tape = on_dev(g["tape"])
    K.set_conv_impl(impl)
    out = {}
    try:
        G.train(); D.train()
        L.beta = L.get_smoothing_ewma_beta(10.)
        assert abs(L.beta - g["beta"]) < 1e-12
        L._init_lagged(); L._attach_ewma()
        if L.sched_bool:                 # train() attaches the resolution-dependent LR schedule (lr = lr_base * 1.5 at 128x128)
            L.sched_stop_step = 0
            L._set_scheduler()
        assert abs(L.opt_disc.param_groups[0]["lr"] - g["lr"]) < 1e-12
        # ---- the R1 penalty on its own, through the path the step uses (penalty rides on the loss's own D(real) forward)
        xr = x.detach().clone().requires_grad_(True)
        pen = L.gp_from_forward(D(xr), xr)
        D.zero_grad()
        pen.backward()
        out["gp_value"] = abs(float(pen.detach()) - g["gp_alone"]) / abs(g["gp_alone"])
        cmp = {k: S.compare("gpgrad." + k, dict(D.named_parameters())[k[2:]].grad, v) for k, v in g["gp_grads"].items()}
        out["gp_grads_l2"] = max(c["l2"] for c in cmp.values())
        out["gp_grads_smax"] = max(c["smax"] for c in cmp.values())
        out["gp_grads_worst"] = max(cmp.items(), key=lambda kv: kv[1]["l2"])[0]
        out["gp_per_tensor"] = cmp
        D.zero_grad(set_to_none=True)
        del pen, xr         # (a live autograd graph would keep the parameters' AccumulateGrad nodes bound to this stream: no capture later)
        # ---- the generator step's forward / backward on the initial weights (no Adam step in front of it)
        ga = g["g_alone"]
        set_random_source(TapeSource(on_dev(ga["tape"]), dev))
        for p in D.parameters():
            p.requires_grad_(False)
        img = G(ga["z"].to(dev))
        logits = D(img)
        loss_alone = ops.g_logit_loss(logits, L.loss)
        G.zero_grad()
        loss_alone.backward()
        set_random_source(None)
        for p in D.parameters():
            p.requires_grad_(True)
        ci = S.compare("galone.img", img, ga["img"])
        out["g_alone_img_l2"], out["g_alone_img_smax"] = ci["l2"], ci["smax"]
        out["g_alone_logits"] = float((logits.detach().view(-1).cpu() - ga["logits"]).abs().max() / ga["logits"].abs().max())
        out["g_alone_loss"] = abs(float(loss_alone.detach()) - ga["loss"]) / max(1.0, abs(ga["loss"]))
        cmp = {k: S.compare("galone." + k, dict(G.named_parameters())[k[2:]].grad, v) for k, v in ga["grads"].items()}
        out["g_alone_grads_l2"] = max(c["l2"] for c in cmp.values())
        out["g_alone_grads_smax"] = max(c["smax"] for c in cmp.values())
        out["g_alone_grads_worst"] = max(cmp.items(), key=lambda kv: kv[1]["l2"])[0]
        out["g_alone_per_tensor"] = cmp
        G.zero_grad(set_to_none=True)
        G.w_ewma = None
        del img, logits, loss_alone
        p0 = {id(p): p.detach().clone() for p in list(G.parameters()) + list(D.parameters())}

        if graph:
            # one eager iteration builds the optimisers' device-side state and the grouped-linear tables (a capture cannot);
            # then every piece of training state is put back to its initial value and the step is captured WITH the taped
            # draws (host-side mixing decision, taped latents / noise as device tensors) and replayed
            L.enable_cuda_graphs(True, warmup_iters=1)
            G.device_mixing = False
            L.main_iteration(torch.rand_like(x) * 2 - 1)
            assert L._graph is None
            with torch.no_grad():
                for p in list(G.parameters()) + list(D.parameters()):
                    p.copy_(p0[id(p)])
                for opt in (L.opt_disc, L.opt_gen):
                    for st in opt.state.values():
                        st["exp_avg"].zero_(); st["exp_avg_sq"].zero_()
                    for h in opt._hyper.values():
                        h["t"][1:].zero_()
                for n, p in L.gen_model_lagged.named_parameters():
                    p.copy_(dict(G.named_parameters())[n])
            G.w_ewma = None
            K.weights_updated()
            set_random_source(_CapturableTape(tape, dev))
            ld, lg = L.main_iteration(x)                   # capture + first replay
            assert L._graph is not None
            torch.cuda.synchronize()
            lag_mode_first = False                         # the captured step's EWMA update is the steady-state one
        else:
            set_random_source(TapeSource(tape, dev))
            for p in D.parameters():
                p.requires_grad_(True)
            ld = L.disc_step(x)
            d_grads = {n: p.grad.detach().clone() for n, p in D.named_parameters() if p.grad is not None}
            for p in D.parameters():
                p.requires_grad_(False)
            lg = L.gen_step()
            lag_mode_first = True
        set_random_source(None)
        if graph:
            d_grads = {n: p.grad.detach() for n, p in D.named_parameters() if p.grad is not None}
        g_grads = {n: p.grad.detach() for n, p in G.named_parameters() if p.grad is not None}
        out["loss_d"] = abs(float(ld) - g["losses"][0]) / max(1.0, abs(g["losses"][0]))
        out["loss_g"] = abs(float(lg) - g["losses"][1]) / max(1.0, abs(g["losses"][1]))

        def worst(cmp):
            k = max(cmp, key=lambda n: cmp[n]["l2"])
            return max(c["l2"] for c in cmp.values()), max(c["smax"] for c in cmp.values()), max(c["norm"] for c in cmp.values()), k

        want_d = {k[2:]: v for k, v in g["grads"].items() if k.startswith("d.")}
        want_g = {k[2:]: v for k, v in g["grads"].items() if k.startswith("g.")}
        assert set(d_grads) == set(want_d) and set(g_grads) == set(want_g), (set(d_grads) ^ set(want_d), set(g_grads) ^ set(want_g))
        per_d = {n: S.compare("grad.d." + n, t, want_d[n]) for n, t in d_grads.items()}
        per_g = {n: S.compare("grad.g." + n, t, want_g[n]) for n, t in g_grads.items()}
        out["d_grads_l2"], out["d_grads_smax"], out["d_grads_norm"], out["d_grads_worst"] = worst(per_d)
        out["g_grads_l2"], out["g_grads_smax"], out["g_grads_norm"], out["g_grads_worst"] = worst(per_g)
        out["per_tensor"] = {"d." + n: c for n, c in per_d.items()}
        out["per_tensor"].update({"g." + n: c for n, c in per_g.items()})

        # ---- post-Adam parameters: step 1 of Adam(beta1 = 0) moves every element by lr * g / (|g| + eps), i.e. by +-lr unless the
        # gradient is tiny: hard bound 2*lr per element, and only a small fraction of the sampled elements may differ at all
        lr = g["lr"]
        bad = tot = 0
        hard = 0.0
        for net, tag in ((G, "g."), (D, "d.")):
            for n, p in net.named_parameters():
                ref = g["p1"][tag + n]
                f = p.detach().reshape(-1)
                mine = f[S.sample_index(f.numel()).to(f.device)].cpu()
                d = (mine - ref["sample"]).abs()
                hard = max(hard, float(d.max()))
                bad += int((d > 0.02 * lr).sum()); tot += d.numel()
        out["p1_max_abs_diff_over_lr"] = hard / lr
        out["p1_flip_fraction"] = bad / tot
        # ---- EWMA generator: first-step semantics in the eager run (lagged = post-Adam parameter, progan/learner.py:472 aliasing);
        # the replayed graph holds the steady-state update lagged = p1*(1-beta) + lagged0*beta with lagged0 = p0
        lag_err = 0.0
        for n, p in G.named_parameters():
            lag = dict(L.gen_model_lagged.named_parameters())[n].detach()
            prev = p.detach() if lag_mode_first else p0[id(p)]
            want = p.detach() * (1. - L.beta) + prev * L.beta
            lag_err = max(lag_err, float(((lag - want).abs() / want.abs().clamp_min(1.0)).max()))
        out["lagged_rel_err"] = lag_err
        out["w_ewma"] = float((G.w_ewma.cpu() - g["w_ewma"]).abs().max() / g["w_ewma"].abs().max())
        return out
    finally:
        K.set_conv_impl("fp32")
        set_random_source(None)
