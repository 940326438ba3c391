"""CPU oracle of one main training iteration (D step + G step + Adam + EWMA).  TEST INFRASTRUCTURE ONLY.

Restates the loop body of ProGANLearner.train (progan/learner.py:734-816 D step, :854-916 G step and
EWMA, inherited unchanged by StyleGANLearner, stylegan/learner.py:90) on top of the functional
oracle in `gan_oracle.py`.  All random draws come from a `Draws` record (replayed from a reference
tape or freshly sampled), so the same iteration can be replayed on the B200 path.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch

from . import gan_oracle as O


@dataclass
class GenDraws:
    """Random draws of ONE generator forward (stylegan/architectures.py:417-422, 116, 508)."""
    z: torch.Tensor
    noise: List[torch.Tensor] = field(default_factory=list)
    cutoff_idx: Optional[int] = None
    z2: Optional[torch.Tensor] = None


@dataclass
class IterDraws:
    d_gen: GenDraws
    g_gen: GenDraws
    gp_eps: Optional[torch.Tensor] = None      # WGAN-GP interpolation eps (resnetgan/learner.py:794)


def parse_gen_forward(it, model: str, num_layers: int, pct_mixing: float = 0.9) -> GenDraws:
    """Consume one generator forward's worth of tape events (see make_golden.Tape); `it` is an iterator
    positioned just before the latent draw of the learner (progan/learner.py:750, :873)."""
    kind, z = next(it)
    assert kind == "randn" and z.dim() == 2, (kind, getattr(z, "shape", None))
    d = GenDraws(z=z)
    if model != "StyleGAN":
        return d
    kind, v = next(it)
    assert kind == "np_rand"
    if v < pct_mixing:
        kind, c = next(it)
        assert kind == "randint"
        d.cutoff_idx = int(c.item())
    need_z2 = d.cutoff_idx is not None
    while len(d.noise) < num_layers or (need_z2 and d.z2 is None):
        kind, t = next(it)
        assert kind == "randn", kind
        if t.dim() == 4:
            d.noise.append(t)
        else:
            assert need_z2 and d.z2 is None
            d.z2 = t
    return d


def parse_train_tape(events, model: str, res: int, iters: int, gp_type: str) -> List[IterDraws]:
    num_layers = 2 * (int(math.log2(res)) - 1)
    it = iter(events)
    out = []
    for _ in range(iters):
        dg = parse_gen_forward(it, model, num_layers)
        eps = None
        if gp_type == "wgan-gp":
            kind, eps = next(it)
            assert kind == "rand"
        gg = parse_gen_forward(it, model, num_layers)
        out.append(IterDraws(d_gen=dg, g_gen=gg, gp_eps=eps))
    assert next(it, None) is None, "tape not fully consumed"
    return out


def sample_gen_draws(model: str, res: int, bs: int, len_latent: int, gen: torch.Generator,
                     pct_mixing: float = 0.9, alpha: float = 1.0, dtype=torch.float32) -> GenDraws:
    """Fresh draws with the reference's distributions (for bench / smoke, where no tape exists)."""
    num_layers = 2 * (int(math.log2(res)) - 1)
    d = GenDraws(z=torch.randn(bs, len_latent, generator=gen, dtype=dtype))
    if model != "StyleGAN":
        return d
    if float(torch.rand((), generator=gen)) < pct_mixing:
        hi = num_layers if alpha != 0 else num_layers - 2
        d.cutoff_idx = int(torch.randint(1, hi, (1,), generator=gen))
        d.z2 = torch.randn(bs, len_latent, generator=gen, dtype=dtype)
    for n in range(num_layers):
        r = 4 * 2 ** (n // 2)
        d.noise.append(torch.randn(bs, 1, r, r, generator=gen, dtype=dtype))
    return d


class OracleTrainer:
    """State (params, Adam moments, EWMA copy, w_ewma) + one-iteration update, reference semantics."""

    def __init__(self, g_params: Dict[str, torch.Tensor], d_params: Dict[str, torch.Tensor], *, model: str,
                 res: int, lr: float, loss: str = "nonsaturating", gp_type: Optional[str] = "r1",
                 lda: float = 10.0, gamma: float = 1.0, eps_drift: float = 0.001, beta1: float = 0.0,
                 beta2: float = 0.99, adam_eps: float = 1e-8, ewma_beta: Optional[float] = None,
                 w_ewma_beta: float = 0.995, mbstd_group_size: int = 4, blur: bool = True,
                 alpha: float = 1.0, fade_in: bool = False, g_kwargs: Optional[dict] = None):
        self.model, self.res, self.lr = model, res, lr
        self.loss, self.gp_type, self.lda, self.gamma, self.eps_drift = loss, gp_type, lda, gamma, eps_drift
        self.b1, self.b2, self.adam_eps = beta1, beta2, adam_eps
        self.ewma_beta, self.w_ewma_beta = ewma_beta, w_ewma_beta
        self.mbstd, self.blur, self.alpha, self.fade_in = mbstd_group_size, blur, alpha, fade_in
        self.g_kwargs = g_kwargs or {}
        self.g = {k: v.detach().clone() for k, v in g_params.items()}
        self.d = {k: v.detach().clone() for k, v in d_params.items()}
        # prev_torgb / prev_fromrgb are excluded from the optimiser outside fade-in (progan/learner.py:1072-1086)
        self.g_m = {k: torch.zeros_like(v) for k, v in self.g.items()}
        self.g_v = {k: torch.zeros_like(v) for k, v in self.g.items()}
        self.d_m = {k: torch.zeros_like(v) for k, v in self.d.items()}
        self.d_v = {k: torch.zeros_like(v) for k, v in self.d.items()}
        self.g_step_n = 0
        self.d_step_n = 0
        # progan/learner.py:472: `lagged_params` starts as the dict of the LIVE parameter tensors, so at the
        # first G step "lagged" is already the post-Adam value: lagged_1 = p_1*(1-b) + p_1*b.
        self.use_ewma = ewma_beta is not None
        self.lagged = None
        self.w_ewma = None
        self.losses: List[float] = []

    # ---- forwards ------------------------------------------------------- #
    def gen_forward(self, params, dr: GenDraws):
        if self.model == "StyleGAN":
            img, w = O.style_generator_forward(params, dr.z, res=self.res, noise=dr.noise, z2=dr.z2,
                                               cutoff_idx=dr.cutoff_idx, alpha=self.alpha, fade_in=self.fade_in,
                                               blur=self.blur, return_w=True, **self.g_kwargs)
            self.w_ewma = O.w_ewma_update(self.w_ewma, w, self.w_ewma_beta)
            return img
        return O.pro_generator_forward(params, dr.z, res=self.res, alpha=self.alpha, fade_in=self.fade_in,
                                       blur=self.blur, **self.g_kwargs)

    def disc_forward(self, params, x):
        return O.pro_discriminator_forward(params, x, res=self.res, alpha=self.alpha, fade_in=self.fade_in,
                                           blur=self.blur, mbstd_group_size=self.mbstd)

    def _trainable(self, name: str) -> bool:
        return self.fade_in or not (name.startswith("prev_torgb") or name.startswith("prev_fromrgb"))

    # ---- steps ----------------------------------------------------------- #
    def d_step(self, real: torch.Tensor, dr: GenDraws, gp_eps=None):
        """progan/learner.py:734-816."""
        with torch.no_grad():
            fake = self.gen_forward(self.g, dr)
        if self.fade_in:
            real = O.fade_real_images(real, self.alpha)
        dp = {k: v.clone().requires_grad_(self._trainable(k)) for k, v in self.d.items()}
        loss = O.disc_loss(lambda t: self.disc_forward(dp, t), fake, real, loss=self.loss, gp_type=self.gp_type,
                           lda=self.lda, gamma=self.gamma, eps_drift=self.eps_drift, gp_eps=gp_eps)
        names = [k for k in dp if dp[k].requires_grad]
        grads = torch.autograd.grad(loss, [dp[k] for k in names], allow_unused=True)
        self.d_step_n += 1
        self.last_d_grads = {}
        for k, g in zip(names, grads):
            if g is None:
                continue
            self.last_d_grads[k] = g
            self.d[k], self.d_m[k], self.d_v[k] = O.adam_step(self.d[k], g, self.d_m[k], self.d_v[k], self.d_step_n,
                                                              self.lr, self.b1, self.b2, self.adam_eps)
        self.losses.append(float(loss))
        return float(loss)

    def g_step(self, dr: GenDraws):
        """progan/learner.py:854-916."""
        gp = {k: v.clone().requires_grad_(self._trainable(k)) for k, v in self.g.items()}
        out = self.disc_forward(self.d, self.gen_forward(gp, dr))
        loss = O.gen_loss(out, self.loss)
        names = [k for k in gp if gp[k].requires_grad]
        grads = torch.autograd.grad(loss, [gp[k] for k in names], allow_unused=True)
        self.g_step_n += 1
        self.last_g_grads = {}
        for k, g in zip(names, grads):
            if g is None:
                continue
            self.last_g_grads[k] = g
            self.g[k], self.g_m[k], self.g_v[k] = O.adam_step(self.g[k], g, self.g_m[k], self.g_v[k], self.g_step_n,
                                                              self.lr, self.b1, self.b2, self.adam_eps)
        if self.use_ewma:
            prev = self.lagged if self.lagged is not None else self.g
            self.lagged = {k: (O.ewma_step(prev[k], self.g[k], self.ewma_beta) if self.ewma_beta else self.g[k].clone())
                           for k in self.g}
        self.losses.append(float(loss))
        return float(loss)

    def main_iter(self, real: torch.Tensor, draws: IterDraws):
        ld = self.d_step(real, draws.d_gen, draws.gp_eps)
        lg = self.g_step(draws.g_gen)
        return ld, lg
