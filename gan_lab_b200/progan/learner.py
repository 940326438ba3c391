"""ProGANLearner: the progressive-growing train loop (reference gan_lab/progan/learner.py:86-1130).

`train()` keeps the reference's loop semantics -- phase bookkeeping (grow / fade-in / stabilise, :560-729),
D step (:734-816: G forward, fake + real D forwards, logit loss, gradient penalty with double backward, drift
term, Adam), G step (:854-916: D frozen, G+D forward, loss, backward, Adam, EWMA generator), alpha update and
LR schedule (:951-956) -- on the sm_100a kernels.  Deliberate deviations (all host-side, none numerical):
  * the D-step generator forward runs under no_grad (the reference builds and drops the graph, :752-753);
  * the real-image fade-in blend runs on the device (reference: on the host before upload, :770-779);
  * per-step `.item()` syncs / tqdm strings are replaced by an optional `log_every`;
  * no dependency on a pickled data_config (only used for bit-exact PIL resampling and metrics);
  * Adam + EWMA run as one fused multi-tensor kernel (`optim.FusedAdam`);
  * with `dp=...` (a `parallel.DataParallel` object) gradients are all-reduced over NCCL before each optimiser step.
"""
import copy

import numpy as np
import torch

from .. import ops
from .. import _kernels as K
from .._growth import GrowthState
from ..resnetgan.learner import GANLearner, LearnerConfigCopy
from ..utils.latent_utils import gen_rand_latent_vars
from .architectures import ProGenerator, ProDiscriminator

NONREDEFINABLE_ATTRS = ('model', 'init_res', 'res_samples', 'res_dataset', 'len_latent', 'num_classes',
                        'class_condition', 'use_auxiliary_classifier', 'model_upsample_type', 'model_downsample_type',
                        'align_corners', 'blur_type', 'nonlinearity', 'use_equalized_lr', 'normalize_z',
                        'use_pixelnorm', 'mbstd_group_size', 'use_ewma_gen',)
REDEFINABLE_FROM_LEARNER_ATTRS = ('batch_size', 'loss', 'gradient_penalty', 'optimizer', 'lr_sched', 'latent_distribution',)

COMPUTE_EWMA_VIA_HALFLIFE = True
EWMA_SMOOTHING_HALFLIFE = 10.
EWMA_SMOOTHING_BETA = .999


class ProGANLearner(GANLearner):
    """GAN Learner for ProGAN architectures (and, through StyleGANLearner, StyleGAN)."""

    _model_name = 'ProGAN'

    def __init__(self, config):
        super(ProGANLearner, self).__init__(config)
        self.curr_phase_num = 0
        self._graph_on, self._graph, self._graph_eager_iters, self._graph_warmup = False, None, 0, 3
        self.lagged_params = None
        self._progressively_grow = True
        self.dp = None
        self.share_penalty_forward = True
        # D(fake) and D(real) as ONE pass over the concatenated batch: identical values (samples are independent in D except
        # minibatch-stddev, whose groups of `mbstd_group_size` consecutive samples stay inside one half when the batch size is a
        # multiple of it), half the discriminator launches, twice the work per launch in the latency-bound low-resolution layers,
        # no gradient accumulation between the two passes.  Measured on B200 (cfg2): SLOWER, 568 vs 617 img/s (the R1 double
        # backward then runs over the whole 2N-sample graph) -- kept as an option (bench: GLB_BATCH_D=1), off.
        self.batch_d_passes = False
        self.parallel_d_passes = False     # two-stream D passes: measured neutral on B200 (598 vs 596 img/s at cfg2), kept as an option
        self._side_stream = None
        # data parallelism: overlap the discriminator's gradient all-reduce + Adam pass with the generator forward of the G step
        # that follows (main_iteration() only: there exactly one G step follows every D step).  Measured on B200 (cfg2): +1.2 %
        # on 2 GPUs (1188 vs 1173 img/s), no change on 4; on 8 GPUs two of the eight ranks died with SIGSEGV inside the captured
        # side-stream NCCL group (the same run without the overlap is fine) -- OFF by default until that is understood.
        self.overlap_d_update = False
        self._defer_d_update = False
        self._d_update_pending = False
        if self.model == self._model_name:
            self.config = LearnerConfigCopy(config, self.__class__.__name__, self._nonredefinable(),
                                            REDEFINABLE_FROM_LEARNER_ATTRS)
            self.latent_distribution = self.config.latent_distribution
            self.state = GrowthState()
            self._build_models()
            self._finish_init(config)

    def _nonredefinable(self):
        return NONREDEFINABLE_ATTRS

    # ------------------------------------------------------------------ construction (reference :116-215)
    def _build_models(self):
        c = self.config
        self.gen_model = ProGenerator(final_res=c.res_samples, len_latent=c.len_latent, upsampler=self.gen_model_upsampler,
                                      blur_type=c.blur_type, nl=self.nl, num_classes=0, equalized_lr=c.use_equalized_lr,
                                      normalize_z=c.normalize_z, use_pixelnorm=c.use_pixelnorm, state=self.state)
        self.disc_model = ProDiscriminator(final_res=c.res_samples, pooler=self.disc_model_downsampler, blur_type=c.blur_type,
                                           nl=self.nl, num_classes=0, equalized_lr=c.use_equalized_lr,
                                           mbstd_group_size=c.mbstd_group_size, state=self.state)

    def _finish_init(self, config):
        c = self.config
        assert c.init_res <= c.res_samples
        if getattr(c, 'bit_exact_resampling', False):
            # reference :770-776: the real images' fade-in skip connection through PIL (de-normalise -> uint8 -> BOX down ->
            # NEAREST up -> normalise).  Only the default torch-resampling variant (:777-779) has a kernel (`ops.fade_real`).
            raise NotImplementedError("bit_exact_resampling=True (PIL resampling of the real images' skip connection) is not "
                                      "built; the default (False) is")
        if c.init_res > 4:
            _init_res_log2 = int(np.log2(c.init_res))
            if float(c.init_res) != 2 ** _init_res_log2:
                raise ValueError('Only resolutions that are powers of 2 are supported.')
            for _ in range(_init_res_log2 - 2):
                self.gen_model.increase_scale()
                self.disc_model.increase_scale()
            self.gen_model.fade_in_phase = False
        assert self.gen_model.cls_base is self.disc_model.cls_base
        self.gen_model.to(c.dev)
        self.disc_model.to(c.dev)
        # EWMA generator: the reference deep-copies G (reference :165-174) and keeps `lagged_params`
        self.gen_model_lagged = None
        if c.use_ewma_gen:
            with torch.no_grad():
                self.gen_model_lagged = copy.deepcopy(self.gen_model)
        self.batch_size = c.bs_dict[self.gen_model.curr_res]
        self._loss = config.loss.casefold()
        self._set_loss()
        self._set_optimizer()
        self.eps = c.eps_drift > 0

    def _set_optimizer(self):
        """reference progan/learner.py:1064-1095: prev_torgb / prev_fromrgb are only optimised while fading in."""
        adam_gan = self._adam_factory()
        if self.gen_model.fade_in_phase:
            self.opt_gen = adam_gan(params=self.gen_model.parameters())
            self.opt_disc = adam_gan(params=self.disc_model.parameters())
        else:
            self.opt_gen = adam_gan(params=self.gen_model.most_parameters(
                excluded_params=['prev_torgb.conv2d.weight', 'prev_torgb.conv2d.bias']))
            self.opt_disc = adam_gan(params=self.disc_model.most_parameters(
                excluded_params=['prev_fromrgb.0.conv2d.weight', 'prev_fromrgb.0.conv2d.bias']))
        self._attach_ewma()

    def _attach_ewma(self):
        if getattr(self, 'lagged_params', None) is not None and self.beta is not None:
            self.opt_gen.attach_ewma(list(self.gen_model.named_parameters()), self.lagged_params, self.beta)
            self.opt_gen._ewma_started = self._ewma_started

    def _set_scheduler(self):
        """reference progan/learner.py:1034-1062."""
        if self._lr_sched == 'resolution dependent':
            self.scheduler_fn = lambda _: self.config.lr_fctr_dict[self.gen_model.curr_res]
        elif self._lr_sched == 'linear decay':
            self.scheduler_fn = lambda main_iter: 1. - (main_iter + self.sched_stop_step) * (1. / self.num_main_iters)
        elif self._lr_sched == 'custom':
            self.scheduler_fn = eval(self.config.lr_sched_custom)
        else:
            raise ValueError("Currently supported LR Schedulers are: [ 'resolution dependent', 'linear decay', 'custom' ]")
        self.scheduler_gen = torch.optim.lr_scheduler.LambdaLR(self.opt_gen, self.scheduler_fn, last_epoch=-1)
        self.scheduler_disc = torch.optim.lr_scheduler.LambdaLR(self.opt_disc, self.scheduler_fn, last_epoch=-1)

    def get_smoothing_ewma_beta(self, half_life):
        """reference progan/learner.py:1124-1127."""
        assert isinstance(half_life, float)
        return .5 ** ((self.batch_size * self.config.gen_bs_mult) / (half_life * 1000.)) if half_life > 0. else 0.

    def _init_lagged(self):
        """`lagged_params` starts as an alias of the live parameters in the reference (progan/learner.py:472), i.e. the
        first EWMA update sees lagged == post-Adam param; here the lagged tensors are separate buffers and the
        fused kernel's first-step mode reproduces that."""
        self.lagged_params = {n: p for n, p in self.gen_model_lagged.named_parameters()}
        for p in self.lagged_params.values():
            p.requires_grad_(False)
        self._ewma_started = False

    @torch.no_grad()
    def _update_gen_lagged(self):
        """The lagged tensors ARE gen_model_lagged's parameters here, so it is always current (reference :234-242)."""
        return self.gen_model_lagged

    def _sync_lagged_structure(self):
        """After increase_scale(): give the lagged generator the new blocks (initialised from the live ones) and
        carry torgb -> prev_torgb over, as reference progan/learner.py:660-686 does on `lagged_params`."""
        old = {n: p.detach().clone() for n, p in self.gen_model_lagged.named_parameters()}
        fresh = []
        with torch.no_grad():
            self.gen_model_lagged = copy.deepcopy(self.gen_model)
            for n, p in self.gen_model_lagged.named_parameters():
                src = None
                if n.startswith('prev_torgb.'):
                    src = old.get(n.replace('prev_torgb.', 'torgb.', 1))
                elif not n.startswith('torgb.'):
                    src = old.get(n)
                if src is not None and src.shape == p.shape:
                    p.copy_(src)
                else:
                    fresh.append(n)
        # the reference ALIASES lagged_params[name] to the live Parameter for every new name (:664-666): their first update after
        # the growth yields lagged = post-Adam parameter; reproduced after the next generator step (see gen_step)
        self._fresh_lagged = fresh
        self._init_lagged_keep_started()

    def _init_lagged_keep_started(self):
        started = getattr(self, '_ewma_started', False)
        self._init_lagged()
        self._ewma_started = started

    # ------------------------------------------------------------------ one step each (used by train() and by bench/tests)
    def _next_real(self, train_dl):
        batch = next(self.train_dataiter, None)
        if batch is None:
            self.curr_epoch_num += 1
            self.train_dataiter = iter(train_dl)
            batch = next(self.train_dataiter)
        xb = batch[0] if isinstance(batch, (list, tuple)) else batch
        return xb

    def disc_step(self, xb):
        """One discriminator step on real batch `xb` (reference progan/learner.py:734-816).  Returns the loss tensor."""
        c = self.config
        if str(c.dev).startswith('cuda'):
            K.begin_step('disc', c.dev)
        try:
            return self._disc_step(xb)
        finally:
            K.end_step()

    def _disc_step(self, xb):
        c = self.config
        self.disc_model.zero_grad()
        zb = gen_rand_latent_vars(num_samples=self.batch_size, length=c.len_latent, distribution=self.latent_distribution,
                                  device=c.dev)
        with torch.no_grad():
            _xgenb = self.gen_model(zb).detach()
        xb = xb.to(c.dev, non_blocking=True)
        if self.gen_model.fade_in_phase:
            xb = ops.fade_real(xb, self.state.blend_coefs()[0])
        # R1 / R2 differentiate D at exactly the batch one of the two loss forwards already ran on (reference
        # resnetgan/learner.py:799-802 re-runs D on the same tensor): share that forward, so the penalty's create_graph
        # gradient and its double backward ride on the loss's own graph -- same values, one D forward and one
        # dgrad+wgrad sweep fewer per D step (3 of 14 F_D).  `share_penalty_forward = False` restores the literal order.
        gp = self.gradient_penalty
        share = self.share_penalty_forward and gp in ('r1', 'r2')
        if share and gp == 'r1':
            xb = xb.detach().requires_grad_(True)
        elif share:
            _xgenb = _xgenb.detach().requires_grad_(True)
        # The fake and the real pass are independent until the loss: run D(fake) on a second stream.  Autograd runs each
        # backward node on its forward's stream, so the two passes stay on two streams through backward as well, and in the
        # captured CUDA graph they are parallel branches: the HBM-bound glue kernels of one pass fill in under the
        # tensor-bound convolutions of the other, and the low-resolution layers (a few CTAs each) overlap.
        two_streams = self.parallel_d_passes and xb.is_cuda
        if two_streams:
            main = torch.cuda.current_stream()
            if self._side_stream is None:
                self._side_stream = torch.cuda.Stream()
            side = self._side_stream
            side.wait_stream(main)
            _xgenb.record_stream(side)
            with torch.cuda.stream(side):
                discriminative_gen = self.disc_model(_xgenb)
            discriminative_real = self.disc_model(xb)
            if share:     # the penalty's create_graph gradient rides on the real pass (main stream) while D(fake) is still running
                gp_term = self.gp_from_forward(discriminative_real, xb) if gp == 'r1' else None
            main.wait_stream(side)
            discriminative_gen.record_stream(main)
        elif self._can_batch_d_passes(_xgenb, xb):
            both = self.disc_model(torch.cat((_xgenb, xb), dim=0))
            discriminative_gen, discriminative_real = both[:_xgenb.shape[0]], both[_xgenb.shape[0]:]
            gp_term = None
        else:
            discriminative_gen = self.disc_model(_xgenb)
            discriminative_real = self.disc_model(xb)
            gp_term = None
        loss_train_disc = ops.d_logit_loss(discriminative_gen, discriminative_real, self.loss,
                                           c.eps_drift if self.eps else 0.)
        if share and two_streams and gp == 'r1':
            loss_train_disc = loss_train_disc + gp_term
            torch.autograd.backward(loss_train_disc, inputs=[p for p in self.disc_model.parameters() if p.requires_grad])
        elif share:
            loss_train_disc = loss_train_disc + (self.gp_from_forward(discriminative_real, xb) if gp == 'r1' else
                                                 self.gp_from_forward(discriminative_gen, _xgenb))
            torch.autograd.backward(loss_train_disc, inputs=[p for p in self.disc_model.parameters() if p.requires_grad])
        else:
            if gp is not None:
                loss_train_disc = loss_train_disc + self.calc_gp(_xgenb, xb)
            loss_train_disc.backward()
        if self.dp is not None:
            group = c.mbstd_group_size
            if group and group > 1 and self.batch_size % min(group, self.batch_size) != 0:
                raise ValueError("data parallelism needs a per-GPU batch that is a multiple of the minibatch-stddev group "
                                 f"({self.batch_size} vs {group}): a group must not straddle ranks (SURVEY.md 8e)")
        if self._defer_d_update:
            # main_iteration(): the all-reduce of the discriminator's gradients and its Adam pass do not feed the generator's
            # forward pass of the G step that follows -- they run beside it, on a second stream (see _gen_step)
            self._d_update_pending = True
        else:
            self._finish_disc_update()
        return loss_train_disc.detach()

    def _finish_disc_update(self):
        if self.dp is not None:
            self.dp.allreduce_grads(self.disc_model)
        self.opt_disc.step()
        self._d_update_pending = False

    def _can_batch_d_passes(self, fake, real):
        if not self.batch_d_passes or fake.shape != real.shape:
            return False
        group = self.config.mbstd_group_size
        return group is None or group <= 1 or fake.shape[0] % min(group, fake.shape[0]) == 0 and fake.shape[0] >= group

    def gen_step(self):
        """One generator step (reference progan/learner.py:854-916).  Returns the loss tensor."""
        c = self.config
        if str(c.dev).startswith('cuda'):
            K.begin_step('gen', c.dev)
        try:
            return self._gen_step()
        finally:
            K.end_step()

    def _gen_step(self):
        c = self.config
        self.gen_model.zero_grad()
        zb = gen_rand_latent_vars(num_samples=self.batch_size * c.gen_bs_mult, length=c.len_latent,
                                  distribution=self.latent_distribution, device=c.dev)
        zb.requires_grad_(True)
        if self._d_update_pending:
            if zb.is_cuda:
                main = torch.cuda.current_stream()
                if self._side_stream is None:
                    self._side_stream = torch.cuda.Stream()
                side = self._side_stream
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    self._finish_disc_update()
                img = self.gen_model(zb)                 # generator forward: independent of the discriminator's update
                main.wait_stream(side)
            else:
                self._finish_disc_update()
                img = self.gen_model(zb)
            loss_train_gen = ops.g_logit_loss(self.disc_model(img), self.loss)
        else:
            loss_train_gen = ops.g_logit_loss(self.disc_model(self.gen_model(zb)), self.loss)
        loss_train_gen.backward()
        if self.dp is not None:
            self.dp.allreduce_grads(self.gen_model)
        self.opt_gen.step()      # Adam + EWMA generator in one fused pass
        if c.use_ewma_gen:
            self._ewma_started = True
            fresh = getattr(self, '_fresh_lagged', None)
            if fresh:            # first generator step after a growth (always an eager iteration): new names start from the live values
                live = dict(self.gen_model.named_parameters())
                with torch.no_grad():
                    for n in fresh:
                        if n in self.lagged_params and n in live:
                            self.lagged_params[n].copy_(live[n])
                self._fresh_lagged = []
        return loss_train_gen.detach()

    # ------------------------------------------------------------------ CUDA-graph replay of the two steps
    # One main iteration is ~1000 kernel launches behind ~50 us of Python each; once the kernels are fast the loop is
    # launch bound.  With `enable_cuda_graphs()` the D step and the G step (forward, double
    # backward, fused Adam/EWMA, RNG) are captured once per (resolution, phase, batch size) after a few eager
    # iterations and replayed; the real batch is copied into a static input buffer, the losses are static tensors.
    def enable_cuda_graphs(self, enabled=True, warmup_iters=3, device_alpha=None):
        """device_alpha (default: on with graphs): keep the fade-in alpha in a device vector read by the blend kernels
        (`_kernels.DeviceAlpha`), so that fade-in phases -- whose alpha moves every iteration -- are captured once and replayed
        as well; without it they run eagerly (alpha is then a kernel argument)."""
        if device_alpha is None:
            device_alpha = bool(enabled)
        if device_alpha:
            self.state.enable_device_alpha(torch.device(self.config.dev))
        else:
            self.state.alpha_dev = None
        self._graph_on = bool(enabled)
        self._graph_warmup = int(warmup_iters)
        self._graph = None
        self._graph_eager_iters = 0
        if hasattr(self.gen_model, '_use_mixing_reg'):
            self.gen_model.device_mixing = bool(enabled)   # mixing decision on the device (graph replayable)

    def _push_alpha(self):
        """Write the current alpha to its device-side copy (no-op without one); never inside a capture."""
        if self.state.alpha_dev is not None:
            self.state.alpha_dev.set(self.gen_model.alpha)

    def _graph_key(self):
        g = self.gen_model
        if g.fade_in_phase:      # with a device-side alpha only its being zero matters (host-side control flow depends on it)
            alpha_key = ('dev', g.alpha != 0) if self.state.alpha_dev is not None else float(g.alpha)
        else:
            alpha_key = None
        return (g.curr_res, bool(g.fade_in_phase), alpha_key, self.batch_size,
                id(self.opt_disc), id(self.opt_gen), self.loss, self.gradient_penalty)

    def _graphs_allowed(self):
        if not self._graph_on or self.config.num_disc_iters != 1 or self.config.num_gen_iters != 1:
            return False
        if self.gen_model.fade_in_phase and getattr(self, 'delta_alpha', 0.0) != 0.0 and self.state.alpha_dev is None:
            return False                      # alpha is a kernel argument that changes every iteration while fading in
        return str(self.config.dev).startswith('cuda')

    def _capture_graphs(self):
        from .. import _kernels as K
        c = self.config
        res = self.gen_model.curr_res
        st = {'key': self._graph_key(), 'x': torch.empty(self.batch_size, 3, res, res, device=c.dev)}
        self.opt_disc.prepare_capture(); self.opt_gen.prepare_capture()
        self.disc_model.zero_grad(set_to_none=True); self.gen_model.zero_grad(set_to_none=True)
        for p in self.disc_model.parameters():
            p.requires_grad_(True)
        K.weights_updated()
        gd = torch.cuda.CUDAGraph()
        self._defer_d_update = self._overlap_active()
        try:
            with torch.cuda.graph(gd):
                st['ld'] = self.disc_step(st['x'])
        finally:
            self._defer_d_update = False
        for p in self.disc_model.parameters():
            p.requires_grad_(False)
        K.weights_updated()
        gg = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gg, pool=gd.pool()):
            st['lg'] = self.gen_step()
        K.weights_updated()
        self.opt_disc.finish_capture(); self.opt_gen.finish_capture()
        st['gd'], st['gg'] = gd, gg
        return st

    def main_iteration(self, xb):
        """One D step on real batch `xb` + one G step (the num_disc_iters = num_gen_iters = 1 case); eager for the
        first `warmup_iters` iterations of a phase, CUDA-graph replay afterwards.  Returns (loss_d, loss_g) tensors."""
        self._push_alpha()
        if self._graphs_allowed():
            key = self._graph_key()
            if key != getattr(self, '_graph_last_key', None):
                # a new phase: drop the captured graphs AND restart the eager warm-up (the grouped-linear tables and the
                # optimisers' device-side hyper vectors of the new phase are built by its first eager iterations)
                self._graph, self._graph_eager_iters, self._graph_last_key = None, 0, key
            if self._graph is None and self._graph_eager_iters >= self._graph_warmup:
                self._graph = self._capture_graphs()
            if self._graph is not None:
                g = self._graph
                g['x'].copy_(xb, non_blocking=True)
                self.opt_disc.push_lr(); self.opt_gen.push_lr()
                g['gd'].replay()
                g['gg'].replay()
                # the replayed optimiser kernels rewrote the weights through raw pointers: no torch version bump, so the
                # weight re-layout caches (transposed / padded / packed copies) must not survive into an eager pass
                K.weights_updated()
                return g['ld'], g['lg']
            self._graph_eager_iters += 1
        for p in self.disc_model.parameters():
            p.requires_grad_(True)
        self._defer_d_update = self._overlap_active()
        try:
            loss_d = self.disc_step(xb)
        finally:
            self._defer_d_update = False
        for p in self.disc_model.parameters():
            p.requires_grad_(False)
        loss_g = self.gen_step()
        return loss_d, loss_g

    def _overlap_active(self):
        return bool(self.overlap_d_update and self.dp is not None and getattr(self.dp, 'world', 1) > 1)

    # ------------------------------------------------------------------ validation metrics hooks (reference :248-416)
    def _fade_real_for_metrics(self, xb):
        """Real validation images are faded like the generated ones (reference :371-378)."""
        return ops.fade_real(xb, self.gen_model.alpha) if self.gen_model.fade_in_phase else xb      # (host-side alpha: eval path)

    def _lagged_for_metrics(self):
        return self._update_gen_lagged() if self.config.use_ewma_gen else None

    def compute_metrics(self, metrics, metrics_type, z_valid_dl, valid_dl=None):
        self._sync_replicas()
        return super(ProGANLearner, self).compute_metrics(metrics, metrics_type, z_valid_dl, valid_dl)

    # ------------------------------------------------------------------ checkpoints (reference progan/learner.py:1238-1460)
    @property
    def progressively_grow(self):
        return self._progressively_grow

    def _extra_checkpoint_entries(self):
        return {}

    def _broadcast_replicas(self):
        """Data parallelism: rank 0's parameters to every rank (collective).  New blocks / torgb / fromrgb are initialised
        from each rank's own RNG by increase_scale(), and a checkpoint may have been loaded on one rank only; gradients are
        averaged, so replicas that start different stay different.  Called after every growth and after load_model()."""
        if self.dp is None or getattr(self.dp, 'world', 1) == 1:
            return
        self.dp.broadcast_params(self.gen_model)
        self.dp.broadcast_params(self.disc_model)
        if self.gen_model_lagged is not None:
            self.dp.broadcast_params(self.gen_model_lagged)
        from .. import _kernels as K
        K.weights_updated()

    def _sync_replicas(self):
        """Bring rank-local running statistics to their global-batch value (collective: every rank calls it).  Nothing to do
        for ProGAN; StyleGAN averages `w_ewma`."""

    def save_model(self, save_path):
        """reference progan/learner.py:1238-1298 (stylegan/learner.py:433-506 adds `_extra_checkpoint_entries`); the file is
        readable by the reference's own load_model.  Under data parallelism only rank 0 writes (replicas are identical)."""
        from pathlib import Path
        from .. import checkpoint as ckpt
        self._sync_replicas()
        if self.dp is not None and getattr(self.dp, 'rank', 0) != 0:
            return
        common = self._checkpoint_common()
        gmeta, dmeta = self._model_metadata()
        lagged = None
        if self.config.use_ewma_gen and self.lagged_params is not None:
            lagged = ckpt.IndexedOrderedDict((k, v.detach().clone(memory_format=torch.contiguous_format))
                                             for k, v in self.lagged_params.items())
        save_path = Path(save_path)
        save_path.parents[0].mkdir(parents=True, exist_ok=True)
        ckpt.save({
            **common,
            'curr_res': self.gen_model.curr_res,
            'alpha': self.gen_model.alpha,
            **self._extra_checkpoint_entries(),
            'gen_model_metadata': gmeta,
            'gen_model_lagged_state_dict': ckpt.plain_state_dict(self.gen_model_lagged) if self.config.use_ewma_gen else None,
            'disc_model_metadata': dmeta,
            'nimg_transition_lst': self.nimg_transition_lst,
            'latent_distribution': self.latent_distribution,
            'curr_phase_num': self.curr_phase_num,
            'lagged_params': lagged,
            'progressively_grow': self.progressively_grow,
        }, save_path)

    def _restore_extra_checkpoint_entries(self, checkpoint):
        pass

    def load_model(self, load_path, dev_of_saved_model='cpu', dev=None):
        """reference progan/learner.py:1300-1460 / stylegan/learner.py:508-682, in the reference's order: config, networks
        grown to the stored resolution, alpha (which also settles fade_in_phase), EWMA generator, parameters, optimisers
        (rebuilt for the stored phase by the `optimizer` setter, then their Adam state), bookkeeping.  Reads files written by
        the reference itself.  `train()` afterwards takes the reference's `pretrained_model` branches."""
        from ..utils.custom_layers import as_native_upsampler, as_native_pooler
        checkpoint = self._load_checkpoint_file(load_path, dev_of_saved_model)
        self._adopt_checkpoint_config(checkpoint, dev)
        c = self.config
        self._latent_distribution = checkpoint['latent_distribution']
        gmeta, dmeta = checkpoint['gen_model_metadata'], checkpoint['disc_model_metadata']
        self.gen_model_metadata, self.disc_model_metadata = gmeta, dmeta
        self.gen_model_upsampler = as_native_upsampler(gmeta['gen_model_upsampler'])
        self.disc_model_downsampler = as_native_pooler(dmeta['disc_model_downsampler'])
        self.num_classes_gen, self.num_classes_disc = gmeta['num_classes_gen'], dmeta['num_classes_disc']
        had_device_alpha = getattr(self, 'state', None) is not None and self.state.alpha_dev is not None
        self.state = GrowthState()
        if had_device_alpha:
            self.state.enable_device_alpha(torch.device(c.dev))
        self._build_models()
        self._restore_extra_checkpoint_entries(checkpoint)
        assert c.init_res <= c.res_samples
        _curr_res = checkpoint['curr_res']
        if _curr_res > 4:
            _init_res_log2 = int(np.log2(_curr_res))
            if float(_curr_res) != 2 ** _init_res_log2:
                raise ValueError('Only resolutions that are powers of 2 are supported.')
            for _ in range(_init_res_log2 - 2):
                self.gen_model.increase_scale()
                self.disc_model.increase_scale()
        self.gen_model.alpha = checkpoint['alpha']      # applies to both networks and takes care of fade_in_phase
        assert self.gen_model.cls_base is self.disc_model.cls_base
        self.gen_model_lagged = None
        if c.use_ewma_gen:
            with torch.no_grad():
                self.gen_model_lagged = copy.deepcopy(self.gen_model)
        self.gen_model.to(c.dev)
        self.gen_model.load_state_dict(checkpoint['gen_model_state_dict'])
        self.gen_model.zero_grad()
        if c.use_ewma_gen:
            self.gen_model_lagged.to(c.dev)
            self.gen_model_lagged.load_state_dict(checkpoint['gen_model_lagged_state_dict'])
            self._restore_lagged_extras(checkpoint)
            self.gen_model_lagged.zero_grad()
        self.disc_model.to(c.dev)
        self.disc_model.load_state_dict(checkpoint['disc_model_state_dict'])
        self.disc_model.zero_grad()
        self.lagged_params = None          # train() re-aliases it to the live parameters on resume (reference :462-472)
        self._graph, self._graph_eager_iters = None, 0
        self._load_checkpoint_common(checkpoint)
        self.nimg_transition_lst = checkpoint['nimg_transition_lst']
        self.latent_distribution = checkpoint['latent_distribution']
        self.curr_phase_num = checkpoint['curr_phase_num']
        self.eps = c.eps_drift > 0
        self._progressively_grow = checkpoint['progressively_grow']
        if hasattr(self, 'delta_alpha'):
            del self.delta_alpha
        self._broadcast_replicas()

    def _restore_lagged_extras(self, checkpoint):
        pass

    # ------------------------------------------------------------------ train loop
    def train(self, train_dl, valid_dl=None, z_valid_dl=None, num_main_iters=None, num_gen_iters=None,
              num_disc_iters=None, log_every=0, step_callback=None):
        """reference progan/learner.py:418-1030 (signature kept).  With `z_valid_dl` (and `valid_dl` for the discriminator)
        and non-empty `config.gen_metrics` / `config.disc_metrics`, validation metrics are computed where the reference
        computes them."""
        c = self.config
        num_main_iters = c.num_main_iters if num_main_iters is None else num_main_iters
        num_gen_iters = c.num_gen_iters if num_gen_iters is None else num_gen_iters
        num_disc_iters = c.num_disc_iters if num_disc_iters is None else num_disc_iters
        self.num_main_iters = num_main_iters
        self.dataset_sz = len(train_dl.dataset) if hasattr(train_dl, 'dataset') else None

        self.gen_model.to(c.dev); self.gen_model.train()
        self.disc_model.to(c.dev); self.disc_model.train()
        if c.use_ewma_gen:
            self.gen_model_lagged.to(c.dev); self.gen_model_lagged.train()

        def round_transition():
            if (c.nimg_transition % self.batch_size) != 0:
                return self.batch_size * (int(c.nimg_transition / self.batch_size) + 1)
            return c.nimg_transition

        if self.not_trained_yet or self.pretrained_model:
            self.nimg_transition = round_transition()
        if self.not_trained_yet:
            self.nimg_transition_lst = [self.nimg_transition]
        if self.not_trained_yet or self.pretrained_model:
            # on a resume the reference re-aliases `lagged_params` to the live parameters as well (:462-472): the EWMA
            # generator restarts from the live one at the first step after a load_model()
            self.beta = None
            if c.use_ewma_gen:
                self.beta = self.get_smoothing_ewma_beta(half_life=EWMA_SMOOTHING_HALFLIFE) if COMPUTE_EWMA_VIA_HALFLIFE \
                    else EWMA_SMOOTHING_BETA
                self._init_lagged()
                self._attach_ewma()
        if self.sched_bool:
            if not self.pretrained_model:
                self.sched_stop_step = 0
            self._set_scheduler()
        if self.not_trained_yet:
            self.train_dataiter = iter(train_dl)
        elif self.pretrained_model:                        # reference :488-494
            if hasattr(train_dl, 'batch_sampler') and train_dl.batch_sampler is not None:
                train_dl.batch_sampler.batch_size = self.batch_size
            if hasattr(train_dl, 'set_resolution'):
                train_dl.set_resolution(self.gen_model.curr_res)
            self.train_dataiter = iter(train_dl)
        if self.gen_model.fade_in_phase and (self.pretrained_model or not hasattr(self, 'delta_alpha')):
            self.delta_alpha = self.batch_size / ((self.nimg_transition / num_disc_iters) - self.batch_size)

        self.last_losses = (None, None)
        for itr in range(num_main_iters):
            for p in self.disc_model.parameters():
                p.requires_grad_(True)

            # ---- phase bookkeeping (reference :560-729) ----
            if self.gen_model.curr_res < self.gen_model.final_res:
                if self.curr_img_num == sum(self.nimg_transition_lst):
                    self.curr_phase_num += 1
                    if self.curr_phase_num % 2 == 1:
                        self.gen_model.zero_grad(); self.disc_model.zero_grad()
                        self.gen_model.increase_scale(); self.disc_model.increase_scale()
                        assert self.gen_model.cls_base is self.disc_model.cls_base
                        self.gen_model.to(c.dev); self.disc_model.to(c.dev)
                        self.batch_size = c.bs_dict[self.gen_model.curr_res]
                        if c.use_ewma_gen:
                            self.beta = self.get_smoothing_ewma_beta(half_life=EWMA_SMOOTHING_HALFLIFE) if \
                                COMPUTE_EWMA_VIA_HALFLIFE else self.beta
                            self._sync_lagged_structure()
                        self._broadcast_replicas()
                        self._set_optimizer()
                        if self.sched_bool:
                            self.sched_stop_step += self.scheduler_gen._step_count
                            self._set_scheduler()
                        self._set_loss()
                        if hasattr(train_dl, 'batch_sampler') and train_dl.batch_sampler is not None:
                            train_dl.batch_sampler.batch_size = self.batch_size
                        if hasattr(train_dl, 'set_resolution'):
                            train_dl.set_resolution(self.gen_model.curr_res)
                            self.train_dataiter = iter(train_dl)
                        self.nimg_transition = round_transition()
                        self.delta_alpha = self.batch_size / ((self.nimg_transition / num_disc_iters) - self.batch_size)
                        self.gen_model.alpha = 0
                    else:
                        self._set_optimizer()
                        if self.sched_bool:
                            self.sched_stop_step += self.scheduler_gen._step_count
                            self._set_scheduler()
                    self.nimg_transition_lst.append(self.nimg_transition)
            elif self._progressively_grow and self.curr_img_num == sum(self.nimg_transition_lst):
                # final phase (reference :709-727)
                self._set_optimizer()
                if self.sched_bool:
                    self.sched_stop_step += self.scheduler_gen._step_count
                    self._set_scheduler()
                self.curr_phase_num += 1
                self.nimg_transition_lst.append(np.inf)
                self._progressively_grow = False

            # validation metrics are computed on the first and every `num_iters_valid`-th iteration, between the steps
            # (reference :820-832, 918-930); those iterations take the step-by-step path
            validate_now = ((itr + 1) % c.num_iters_valid == 0 or itr == 0)
            validate_d = validate_now and z_valid_dl is not None and valid_dl is not None and bool(c.disc_metrics)
            validate_g = validate_now and z_valid_dl is not None and bool(c.gen_metrics)
            if num_disc_iters == 1 and num_gen_iters == 1 and not (validate_d or validate_g):
                # ---- D step + G step (CUDA-graph replay once warmed up, see main_iteration) ----
                loss_d, loss_g = self.main_iteration(self._next_real(train_dl))
                self.curr_dataset_batch_num += 1
                self.curr_img_num += self.batch_size
            else:
                # ---- train discriminator ----
                self._push_alpha()
                for disc_iter in range(num_disc_iters):
                    loss_d = self.disc_step(self._next_real(train_dl))
                    if validate_d and disc_iter == num_disc_iters - 1:
                        vals = self.compute_metrics(metrics=c.disc_metrics, metrics_type='Discriminator',
                                                    z_valid_dl=z_valid_dl, valid_dl=valid_dl)
                        print('|\n', 'Discriminator Validation Metrics:\n', *vals)
                    self.curr_dataset_batch_num += 1
                    self.curr_img_num += self.batch_size

                # ---- train generator ----
                for p in self.disc_model.parameters():
                    p.requires_grad_(False)
                for gen_iter in range(num_gen_iters):
                    loss_g = self.gen_step()
                    if validate_g and gen_iter == num_gen_iters - 1:
                        vals = self.compute_metrics(metrics=c.gen_metrics, metrics_type='Generator',
                                                    z_valid_dl=z_valid_dl, valid_dl=None)
                        print('|\n', 'Generator Validation Metrics:\n', *vals)

            self.last_losses = (loss_d, loss_g)
            if step_callback is not None:
                step_callback(itr, loss_d, loss_g)
            if log_every and (itr % log_every == 0):
                print(f'itr {itr:7d}  res {self.gen_model.curr_res:4d}  '
                      f'{"fade-in" if self.gen_model.fade_in_phase else "stab."}  D {float(loss_d):9.4g}  G {float(loss_g):9.4g}')

            if self.gen_model.fade_in_phase:
                self.gen_model.alpha += self.delta_alpha
            if self.sched_bool:
                self.scheduler_gen.step()
                self.scheduler_disc.step()
            if self.not_trained_yet:
                self.not_trained_yet = False

            # ---- periodic checkpoint (reference :960-985).  The reference rebuilds the optimisers before every save "for
            # [the] niche case when training ends right when alpha becomes 1" -- Adam's moments restart there and the LR
            # schedulers stay attached to the replaced optimisers until the next phase change; kept as is.
            if (itr + 1) % c.num_iters_save_model == 0:
                self._set_optimizer()
                self.gen_model.eval(); self.disc_model.eval()
                if c.use_ewma_gen:
                    self.gen_model_lagged.eval()
                self.save_model(c.save_model_dir / (self.model.casefold().replace(' ', '') + '_model.tar'))
                self.gen_model.train(); self.disc_model.train()
                if c.use_ewma_gen:
                    self.gen_model_lagged.train()

        for p in self.disc_model.parameters():
            p.requires_grad_(True)
        # end of train() (reference :1016-1030): fresh optimisers (every train() call starts Adam from zero moments), the
        # networks left in eval mode.  The reference also parks them on the CPU; device placement is left alone here.
        self._set_optimizer()
        self._sync_replicas()
        self.gen_model.eval(); self.disc_model.eval()
        if c.use_ewma_gen:
            self.gen_model_lagged.eval()
        return self.last_losses
