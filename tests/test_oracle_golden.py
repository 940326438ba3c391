"""Pin the oracle: the CPU restatement in oracle/ must reproduce the fixtures that the UNMODIFIED
reference produced (tests/golden/*.pt, written by oracle/make_golden.py)."""
import math

import pytest
import torch

from oracle import gan_oracle as O
from oracle.train_oracle import OracleTrainer, parse_gen_forward, parse_train_tape

TOL = dict(rtol=1e-5, atol=1e-6)


def close(a, b, rtol=1e-5, atol=1e-6):
    torch.testing.assert_close(a, b, rtol=rtol, atol=atol)


def relerr(a, b):
    return float((a.detach() - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("name", ["conv3x3", "conv3x3_nobias", "conv1x1_torgb", "conv1x1_fromrgb",
                                  "conv4x4_valid", "conv3x3_c33", "conv3x3_lrmul"])
def test_conv2d_ex(golden, name):
    g = golden("layers.pt")[name]
    c = g["cfg"]
    assert abs(O.conv_wscale(c["ni"], c["ks"], c["gain_sq_base"]) - g["wscale"]) < 1e-12
    w = g["sd"]["conv2d.weight"].clone().requires_grad_(True)
    b = g["sd"].get("conv2d.bias")
    b = b.clone().requires_grad_(True) if b is not None else None
    x = g["x"].clone().requires_grad_(True)
    y = O.conv2d_ex(x, w, b, g["wscale"], c["lrmul"], c["padding"])
    close(y, g["y"])
    gx, = torch.autograd.grad(y, x, g["gy"], create_graph=True)
    close(gx, g["gx"])
    ggw, = torch.autograd.grad((gx * g["v"]).sum(), w, retain_graph=True)
    close(ggw, g["gg_w"], rtol=1e-4, atol=1e-5)
    gw, = torch.autograd.grad(y, w, g["gy"], retain_graph=True)
    close(gw, g["grads"]["conv2d.weight"], rtol=1e-4, atol=1e-5)
    if b is not None:
        gb, = torch.autograd.grad(y, b, g["gy"])
        close(gb, g["grads"]["conv2d.bias"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("name", ["linear_mapping", "linear_style", "linear_dhead", "linear_progan_fc"])
def test_linear_ex(golden, name):
    g = golden("layers.pt")[name]
    c = g["cfg"]
    assert abs(O.linear_wscale(c["nin"], c["gain_sq_base"]) - g["wscale"]) < 1e-12
    w = g["sd"]["linear.weight"].clone().requires_grad_(True)
    b = g["sd"]["linear.bias"].clone().requires_grad_(True)
    x = g["x"].clone().requires_grad_(True)
    y = O.linear_ex(x, w, b, g["wscale"], c["lrmul"])
    close(y, g["y"])
    gx, gw, gb = torch.autograd.grad(y, (x, w, b), g["gy"])
    close(gx, g["gx"]); close(gw, g["grads"]["linear.weight"]); close(gb, g["grads"]["linear.bias"])


def test_small_layers(golden):
    L = golden("layers.pt")
    g = L["conv2dbias"]
    close(O.conv2d_bias(g["x"], g["sd"]["bias"]), g["y"])
    for k in ("pixelnorm_z", "pixelnorm_feat"):
        g = L[k]
        x = g["x"].clone().requires_grad_(True)
        y = O.pixelnorm(x)
        close(y, g["y"])
        close(torch.autograd.grad(y, x, g["gy"])[0], g["gx"])
    for k in ("blur", "blur_odd"):
        g = L[k]
        x = g["x"].clone().requires_grad_(True)
        y = O.blur3x3(x)
        close(y, g["y"])
        close(torch.autograd.grad(y, x, g["gy"])[0], g["gx"])
        close(O.blur3x3(g["v"]), g["ggy"])          # blur is linear and self-adjoint
    g = L["upsample2x"]
    close(O.upsample2x(g["x"]), g["y"])


@pytest.mark.parametrize("name", ["mbstd_n8", "mbstd_n4", "mbstd_n6", "mbstd_n1", "mbstd_n16"])
def test_mbstd(golden, name):
    g = golden("layers.pt")[name]
    x = g["x"].clone().requires_grad_(True)
    gy = g["gy"].clone().requires_grad_(True)
    y = O.mbstd_concat(x, g["group_size"])
    close(y, g["y"])
    if x.shape[0] > 1:
        gx, = torch.autograd.grad(y, x, gy, create_graph=True)
        close(gx, g["gx"])
        ggx, ggy = torch.autograd.grad(gx, (x, gy), g["v"])
        close(ggx, g["ggx"], rtol=1e-4, atol=1e-6); close(ggy, g["ggy"], rtol=1e-4, atol=1e-6)


def test_style_epilogue(golden):
    g = golden("layers.pt")["style_epilogue"]
    x = g["x"].clone().requires_grad_(True)
    st = g["style"].clone().requires_grad_(True)
    nw = g["noise_weight"].clone().requires_grad_(True)
    b = g["bias"].clone().requires_grad_(True)
    t = O.instance_norm(O.lrelu(O.conv2d_bias(O.style_add_noise(x, nw, g["noise"]), b)))
    y = O.adain(t, st, x.shape[1])
    close(y, g["y"], rtol=1e-5, atol=1e-5)
    gx, gst, gnw, gb = torch.autograd.grad(y, (x, st, nw, b), g["gy"])
    close(gx, g["gx"], rtol=1e-4, atol=1e-5); close(gst, g["gstyle"], rtol=1e-4, atol=1e-5)
    close(gnw, g["g_noise_weight"], rtol=1e-4, atol=1e-5); close(gb, g["g_bias"], rtol=1e-4, atol=1e-5)


def test_pool_bias_lrelu(golden):
    g = golden("layers.pt")["pool_bias_lrelu"]
    x = g["x"].clone().requires_grad_(True)
    gy = g["gy"].clone().requires_grad_(True)
    y = O.lrelu(O.conv2d_bias(O.avgpool2(x), g["bias"]))
    close(y, g["y"])
    gx, = torch.autograd.grad(y, x, gy, create_graph=True)
    close(gx, g["gx"])
    close(torch.autograd.grad(gx, gy, g["v"])[0], g["ggy"])


def _req(p):
    return {k: v.clone().requires_grad_(True) for k, v in p.items()}


def _check_grads(names, grads, ref, rtol=2e-4):
    for k, gr in zip(names, grads):
        r = ref[k]
        if r is None:
            assert gr is None or float(gr.abs().max()) == 0.0, k
            continue
        assert gr is not None, k
        assert relerr(gr, r) < rtol, (k, relerr(gr, r))


@pytest.mark.parametrize("fname", ["style_nets_res16.pt", "style_nets_res16_fade.pt", "style_nets_res32_taper_fade.pt"])
def test_style_nets(golden, fname):
    g = golden(fname)
    res = g["res"]
    num_layers = 2 * (int(math.log2(res)) - 1)
    dr = parse_gen_forward(iter([("randn", g["z"])] + list(g["tape"])), "StyleGAN", num_layers)
    gp = _req(g["g_sd"])
    img, w = O.style_generator_forward(gp, dr.z, res=res, noise=dr.noise, z2=dr.z2, cutoff_idx=dr.cutoff_idx,
                                       alpha=g["alpha"], fade_in=g["fade_in"], return_w=True)
    close(img, g["img"], rtol=1e-4, atol=1e-5)
    close(O.w_ewma_update(None, w), g["w_ewma"])
    names = list(gp)
    grads = torch.autograd.grad(img, [gp[k] for k in names], g["gimg"], allow_unused=True)
    # layer-0 bias / noise_weight grads pass through InstanceNorm of a constant plane: mathematically ~0,
    # dominated by fp32 cancellation in the reference itself (7e-4 vs an fp64 evaluation) -> 2e-3 here.
    _check_grads(names, grads, g["g_grads"], rtol=2e-3)

    dp = _req(g["d_sd"])
    d_fn = lambda t: O.pro_discriminator_forward(dp, t, res=res, alpha=g["alpha"], fade_in=g["fade_in"])
    logits = d_fn(g["x"])
    close(logits, g["logits"], rtol=1e-4, atol=1e-5)
    names = list(dp)
    grads = torch.autograd.grad(logits, [dp[k] for k in names], g["glog"], allow_unused=True)
    _check_grads(names, grads, g["d_grads"])
    pen = O.gradient_penalty(d_fn, "r1", g["x"], g["img"], g["lda"])
    assert relerr(pen, g["gp"]) < 1e-4
    grads = torch.autograd.grad(pen, [dp[k] for k in names], allow_unused=True)
    _check_grads(names, grads, g["d_gp_grads"], rtol=5e-4)
    xi = g["img"].clone().requires_grad_(True)
    gxi, = torch.autograd.grad(d_fn(xi), xi, g["glog"])
    assert relerr(gxi, g["d_gx_img"]) < 2e-4


@pytest.mark.parametrize("fname", ["pro_nets_res16.pt", "pro_nets_res8_fade.pt", "pro_nets_res32_taper_fade.pt"])
def test_pro_nets(golden, fname):
    g = golden(fname)
    res = g["res"]
    gp = _req(g["g_sd"])
    img = O.pro_generator_forward(gp, g["z"], res=res, alpha=g["alpha"], fade_in=g["fade_in"])
    close(img, g["img"], rtol=1e-4, atol=1e-5)
    names = list(gp)
    grads = torch.autograd.grad(img, [gp[k] for k in names], g["gimg"], allow_unused=True)
    _check_grads(names, grads, g["g_grads"])
    dp = _req(g["d_sd"])
    d_fn = lambda t: O.pro_discriminator_forward(dp, t, res=res, alpha=g["alpha"], fade_in=g["fade_in"])
    logits = d_fn(g["x"])
    close(logits, g["logits"], rtol=1e-4, atol=1e-5)
    names = list(dp)
    grads = torch.autograd.grad(logits, [dp[k] for k in names], g["glog"], allow_unused=True)
    _check_grads(names, grads, g["d_grads"])
    (kind, eps), = g["gp_tape"]
    assert kind == "rand"
    pen = O.gradient_penalty(d_fn, "wgan-gp", g["x"], g["img"], g["lda"], g["gamma"], eps=eps)
    assert relerr(pen, g["gp"]) < 1e-4
    grads = torch.autograd.grad(pen, [dp[k] for k in names], allow_unused=True)
    _check_grads(names, grads, g["d_gp_grads"], rtol=5e-4)


@pytest.mark.parametrize("fname", ["resnet_nets_res64.pt", "resnet_nets_res32.pt"])
def test_resnet_nets(golden, fname):
    """ResNet generator (BatchNorm batch statistics + running buffers, Tanh) and discriminator (LayerNorm) restatements vs
    the reference modules, in fp64 so that no ReLU mask bit flips: forward, parameter gradients, WGAN-GP and its gradients."""
    g = golden(fname)
    res, fmap = g["res"], (g["fmap"] if g["res"] == 64 else 2 * g["fmap"])     # 32-pixel nets use FMAP*2 (learner.py:192)
    dbl = lambda d: {k: (v.double() if v.is_floating_point() else v.clone()) for k, v in d.items()}
    gp = {k: (v.requires_grad_(True) if v.is_floating_point() and "running" not in k else v) for k, v in dbl(g["g_sd"]).items()}
    buffers = {k: v for k, v in gp.items() if "running" in k or "num_batches" in k}
    img = O.resnet_generator_forward(gp, g["z"].double(), res=res, fmap=fmap, buffers=buffers)
    close(img, g["img"].double(), rtol=2e-5, atol=2e-6)
    for k, v in g["g_buffers"].items():
        if k.endswith("num_batches_tracked"):
            assert int(buffers[k]) == int(v)
        else:
            close(buffers[k], v.double(), rtol=1e-5, atol=1e-6)
    names = [k for k in gp if gp[k].requires_grad]
    grads = torch.autograd.grad(img, [gp[k] for k in names], g["gimg"].double(), allow_unused=True)
    big = max(float(v.abs().max()) for v in g["g_grads"].values() if v is not None)
    for k, gr in zip(names, grads):
        ref = g["g_grads"][k].double()
        assert float((gr - ref).abs().max()) <= 2e-5 * max(float(ref.abs().max()), 5e-2 * big), k
    dp = {k: v.requires_grad_(True) for k, v in dbl(g["d_sd"]).items()}
    d_fn = lambda t: O.resnet_discriminator_forward(dp, t, res=res, fmap=fmap)
    logits = d_fn(g["x"].double())
    close(logits, g["logits"].double(), rtol=2e-5, atol=2e-6)
    names = list(dp)
    grads = torch.autograd.grad(logits, [dp[k] for k in names], g["glog"].double(), allow_unused=True)
    for k, gr in zip(names, grads):
        assert relerr(gr, g["d_grads"][k].double()) < 2e-5, k
    (kind, eps), = g["gp_tape"]
    pen = O.gradient_penalty(d_fn, "wgan-gp", g["x"].double(), g["img"].double(), g["lda"], g["gamma"], eps=eps.double())
    assert relerr(pen, g["gp"].double()) < 2e-5
    grads = torch.autograd.grad(pen, [dp[k] for k in names], allow_unused=True)
    for k, gr in zip(names, grads):
        ref = g["d_gp_grads"][k]
        if gr is None or ref is None:         # parameters the penalty does not depend on (last biases)
            assert (gr is None or float(gr.abs().max()) == 0.0) and (ref is None or float(ref.abs().max()) == 0.0), k
        else:
            assert relerr(gr, ref.double()) < 1e-4, (k, relerr(gr, ref.double()))


@pytest.mark.parametrize("fname,model,gp,loss", [("style_train_res16.pt", "StyleGAN", "r1", "nonsaturating"),
                                                ("pro_train_res8.pt", "ProGAN", "wgan-gp", "wgan")])
def test_train_steps(golden, fname, model, gp, loss):
    """Two whole main iterations of the reference's Learner.train(): losses, post-Adam params,
    EWMA generator and w_ewma."""
    g = golden(fname)
    res, bs, iters = g["res"], g["bs"], g["iters"]
    draws = parse_train_tape(g["tape"], model, res, iters, gp)
    kw = dict(use_pixelnorm=True) if model == "ProGAN" else {}
    T = OracleTrainer(g["g_sd0"], g["d_sd0"], model=model, res=res, lr=g["lr"], loss=loss, gp_type=gp,
                      ewma_beta=g["beta"], g_kwargs=kw)
    for i in range(iters):
        T.main_iter(g["data"][i * bs:(i + 1) * bs], draws[i])
    for a, b in zip(T.losses, g["losses"]):
        assert abs(a - b) < 1e-4 * max(1.0, abs(b)), (T.losses, g["losses"])
    # Adam with beta1=0 is a sign-like update (|dp| = lr per step whatever |g| is): an element whose
    # gradient is mathematically ~0 (e.g. layer-0 bias behind InstanceNorm) moves by +-lr per step on
    # rounding noise alone.  So: every element within 2*lr*steps, and >= 99% of all elements tight.
    def adam_close(mine, ref, what):
        bad = tot = 0
        for k, v in ref.items():
            d = (mine[k] - v).abs()
            # step t of Adam(beta2=.99) moves an element by at most lr*sqrt((1-.99^t)/.01); two trajectories by twice that
            bound = 2.001 * g["lr"] * sum(math.sqrt((1. - .99 ** t) / .01) for t in range(1, iters + 1)) + 1e-6
            assert float(d.max()) <= bound, (what, k, float(d.max()), bound)
            bad += int((d > 2e-5 + 1e-4 * v.abs()).sum()); tot += v.numel()
        assert bad <= 0.02 * tot, (what, bad, tot)

    adam_close(T.g, g["g_sd1"], "G")
    adam_close(T.d, g["d_sd1"], "D")
    adam_close(T.lagged, g["lagged"], "EWMA-G")
    if model == "StyleGAN":
        close(T.w_ewma, g["w_ewma"], rtol=1e-3, atol=1e-5)
