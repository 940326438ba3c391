"""GANLearner: base learner (reference gan_lab/resnetgan/learner.py:88-946).

Keeps the reference's constructor `(config: argparse.Namespace)`, its public attributes (`gen_model`,
`disc_model`, `opt_gen`, `opt_disc`, `batch_size`, `config`, `loss`, `gradient_penalty`, `optimizer`,
`lr_sched`), `calc_gp` and `save_model` / `load_model` (files interchangeable with the reference's); metrics and
plotting are outside the hot path and not built.  With `config.model == 'ResNet GAN'` this class builds the 32 / 64-pixel ResNet generator and
discriminator (reference :183-232) and `train()` runs the reference's loop (:463-776: generator step(s) FIRST, then
`num_disc_iters` discriminator steps per main iteration); it is also the shared machinery the ProGAN / StyleGAN
learners build on.
"""
import argparse
import copy

import torch

from .. import ops
from .. import _kernels as K
from ..optim import FusedAdam
from ..utils.custom_layers import Upsample2x, AvgPool2x, NearestPool2d, BilinearPool2d, LeakyReLU, ReLU
from ..utils.latent_utils import RANDOM, gen_rand_latent_vars

FMAP_SAMPLES = 3

NONREDEFINABLE_ATTRS = ('model', 'res_samples', 'res_dataset', 'len_latent', 'num_classes', 'class_condition',
                        'use_auxiliary_classifier', 'model_upsample_type', 'model_downsample_type', 'align_corners',
                        'blur_type', 'nonlinearity', 'use_equalized_lr',)
REDEFINABLE_FROM_LEARNER_ATTRS = ('batch_size', 'loss', 'gradient_penalty', 'optimizer', 'lr_sched',)


class LearnerConfigCopy(object):
    """Private copy of the config whose architecture-defining attributes are frozen (reference _int.py:102-148)."""

    def __init__(self, config, learner_class: str, nonredefinable_attrs: tuple, redefinable_from_learner_attrs: tuple):
        assert isinstance(config, argparse.Namespace) and 'model' in config.__dict__
        object.__setattr__(self, '__dict__', copy.deepcopy(config.__dict__))
        self.__dict__['model_name'] = {'GANLearner': 'resnetgan', 'ProGANLearner': 'progan',
                                       'StyleGANLearner': 'stylegan'}.get(learner_class, learner_class)
        self.__dict__['learner_class'] = learner_class
        self.__dict__['_nonredefinable_attrs'] = nonredefinable_attrs
        self.__dict__['_redefinable_from_learner_attrs'] = redefinable_from_learner_attrs

    def __setattr__(self, name, value):
        if name in self._nonredefinable_attrs:
            raise AttributeError(f'{self.learner_class}().config.{name} attribute cannot be changed once '
                                 f'{self.learner_class} is instantiated.')
        if name in self._redefinable_from_learner_attrs:
            raise AttributeError(f'{self.learner_class}().config.{name} attribute cannot be changed. Instead, please '
                                 f'change {self.learner_class}().{name}.')
        object.__setattr__(self, name, value)

    def __str__(self):
        hidden = ('_nonredefinable_attrs', '_redefinable_from_learner_attrs', 'model_name', 'learner_class')
        return '\n'.join(f'  {k}: {v}' for k, v in vars(self).items() if k not in hidden)


def configure_adam_for_gan(lr_base, betas: tuple, eps=1.e-8, wd=0):
    """reference utils/backprop_utils.py:109-120, returning the fused optimiser's constructor."""
    assert isinstance(betas, tuple)
    from functools import partial
    return partial(FusedAdam, lr=lr_base, betas=betas, eps=eps, weight_decay=wd)


class GANLearner(object):
    def __init__(self, config):
        super(GANLearner, self).__init__()
        self._model = config.model
        self.pretrained_model = False
        self.config = LearnerConfigCopy(config, self.__class__.__name__, NONREDEFINABLE_ATTRS, REDEFINABLE_FROM_LEARNER_ATTRS)
        self.batch_size = config.batch_size
        self.curr_dataset_batch_num = 0
        self.curr_epoch_num = 1
        if config.use_auxiliary_classifier or config.class_condition:
            raise NotImplementedError('class conditioning / auxiliary classifier are off the benchmarked path; not built')
        self.num_classes = 0
        self.cond_gen = self.cond_disc = self.ac = False

        if not (config.res_samples <= config.res_dataset):
            raise ValueError(f'Resolution of generated images (config.res_samples = {config.res_samples}) must be less '
                             f'than or equal to resolution of dataset (config.res_dataset = {config.res_dataset}).')
        # Generator upsampling / discriminator downsampling / nonlinearity selection (reference :143-181)
        if config.model_upsample_type.casefold() != 'nearest':
            raise NotImplementedError("only model_upsample_type='nearest' (the default) has a kernel")
        self.gen_model_upsampler = Upsample2x()
        ds = config.model_downsample_type.casefold()
        if ds in ('average', 'box'):
            self.disc_model_downsampler = AvgPool2x()
        elif ds == 'nearest':
            self.disc_model_downsampler = NearestPool2d()
        elif ds == 'bilinear':
            self.disc_model_downsampler = BilinearPool2d(align_corners=config.align_corners)
        else:
            raise ValueError("Supported Downsampling Types are: [ 'nearest', 'average', 'box', 'bilinear' ]")
        nlin = config.nonlinearity.casefold()
        if nlin == 'leaky relu':
            self.nl = LeakyReLU(negative_slope=config.leakiness)
        elif nlin == 'relu':
            self.nl = ReLU()
        else:
            raise NotImplementedError(f'nonlinearity {config.nonlinearity} has no kernel')

        self.gen_model = None
        self.disc_model = None
        self.dp = None
        self.share_penalty_forward = True
        self._graph_on, self._graph, self._graph_eager_iters, self._graph_warmup = False, None, 0, 3
        if self._model == 'ResNet GAN':
            from .architectures import (Generator32PixResnet, Generator64PixResnet, Discriminator32PixResnet,
                                        Discriminator64PixResnet, FMAP_G, FMAP_D)
            c = self.config
            if c.res_samples == 64:
                _gen_model, _disc_model, _fmap_g, _fmap_d = Generator64PixResnet, Discriminator64PixResnet, FMAP_G, FMAP_D
            elif c.res_samples == 32:
                _gen_model, _disc_model, _fmap_g, _fmap_d = Generator32PixResnet, Discriminator32PixResnet, FMAP_G * 2, FMAP_D * 2
            else:
                raise ValueError('GANLearner currently only supports 32 pixel and 64 pixel GAN architectures.\n'
                                 'If a different generated sample resolution is desired, please use the\n'
                                 'ProGAN or StyleGAN models featured in this package instead.')
            self._fmap_g, self._fmap_d = _fmap_g, _fmap_d
            self.gen_model = _gen_model(len_latent=c.len_latent, fmap=_fmap_g, upsampler=self.gen_model_upsampler,
                                        blur_type=c.blur_type, nl=self.nl, num_classes=0, equalized_lr=c.use_equalized_lr)
            self.disc_model = _disc_model(fmap=_fmap_d, pooler=self.disc_model_downsampler, blur_type=c.blur_type, nl=self.nl,
                                          num_classes=0, equalized_lr=c.use_equalized_lr)
            self.gen_model.to(c.dev)
            self.disc_model.to(c.dev)
            assert self.gen_model.res == self.disc_model.res
        self._loss = config.loss.casefold()
        self._gradient_penalty = config.gradient_penalty
        self._optimizer = config.optimizer.casefold()
        self.opt_gen = None
        self.opt_disc = None
        self._lr_sched = None
        self.sched_bool = False
        self.sched_stop_step = None
        self.scheduler_gen = None
        self.scheduler_disc = None
        if config.lr_sched is not None:
            self._lr_sched = config.lr_sched.casefold()
            self.sched_bool = True
            self.sched_stop_step = 0
        self.curr_img_num = 0
        self.tot_num_epochs = None
        self.dataset_sz = None
        self.not_trained_yet = True
        # bookkeeping of the reference's metrics code (resnetgan/learner.py:274-290); carried through checkpoints unchanged
        self.num_classes_gen = self.num_classes_disc = 0
        self.valid_z = self.valid_label = self.rand_idxs = None
        self.grid_inputs_constructed = False
        self.gen_metrics_num = self.disc_metrics_num = 0
        self.ds_mean = self.ds_std = None
        if self._model == 'ResNet GAN':
            self._set_loss()
            self._set_optimizer()

    # ------------------------------------------------------------------ one step each (reference :544-592, 626-692)
    def gen_step(self):
        """One generator step; the discriminator's parameters must already be frozen.  Returns the loss tensor."""
        c = self.config
        self.gen_model.zero_grad()
        zb = gen_rand_latent_vars(num_samples=self.batch_size * c.gen_bs_mult, length=c.len_latent,
                                  distribution=c.latent_distribution, device=c.dev)
        zb.requires_grad_(True)
        loss_train_gen = ops.g_logit_loss(self.disc_model(self.gen_model(zb)), self.loss)
        loss_train_gen.backward()
        if self.dp is not None:
            self.dp.allreduce_grads(self.gen_model)
        self.opt_gen.step()
        return loss_train_gen.detach()

    def disc_step(self, xb):
        """One discriminator step on the real batch `xb`.  Returns the loss tensor."""
        c = self.config
        self.disc_model.zero_grad()
        zb = gen_rand_latent_vars(num_samples=self.batch_size, length=c.len_latent, distribution=c.latent_distribution,
                                  device=c.dev)
        # the reference builds (and drops) the generator's graph here (:640-641); its BatchNorm running statistics are
        # updated by this forward too, which no_grad keeps
        with torch.no_grad():
            _xgenb = self.gen_model(zb).detach()
        xb = xb.to(c.dev, non_blocking=True)
        gp = self.gradient_penalty
        share = self.share_penalty_forward and gp in ('r1', 'r2')
        if share and gp == 'r1':
            xb = xb.detach().requires_grad_(True)
        elif share:
            _xgenb = _xgenb.detach().requires_grad_(True)
        discriminative_gen = self.disc_model(_xgenb)
        discriminative_real = self.disc_model(xb)
        loss_train_disc = ops.d_logit_loss(discriminative_gen, discriminative_real, self.loss, 0.)
        if share:
            loss_train_disc = loss_train_disc + (self.gp_from_forward(discriminative_real, xb) if gp == 'r1' else
                                                 self.gp_from_forward(discriminative_gen, _xgenb))
            torch.autograd.backward(loss_train_disc, inputs=[p for p in self.disc_model.parameters() if p.requires_grad])
        else:
            if gp is not None:
                loss_train_disc = loss_train_disc + self.calc_gp(_xgenb, xb)
            loss_train_disc.backward()
        if self.dp is not None:
            self.dp.allreduce_grads(self.disc_model)
        self.opt_disc.step()
        return loss_train_disc.detach()

    # ------------------------------------------------------------------ CUDA-graph replay of the steps (opt-in)
    # The ResNet loop launches ~300 kernels per step behind ~50 us of Python each and is launch bound when run eagerly.  With
    # `enable_cuda_graphs()` one generator step and one discriminator step are captured after a few eager iterations and
    # replayed (the generator graph `num_gen_iters` times, the discriminator graph once per real batch, which is copied into a
    # static input buffer): latents and the WGAN-GP interpolation weights are drawn on the device inside the graphs, BatchNorm's
    # running buffers and Adam's step count advance inside them.  Same machinery as ProGANLearner's (which overrides these).
    def enable_cuda_graphs(self, enabled=True, warmup_iters=3):
        self._graph_on = bool(enabled)
        self._graph_warmup = int(warmup_iters)
        self._graph, self._graph_eager_iters = None, 0

    def _graphs_allowed(self):
        return self._graph_on and str(self.config.dev).startswith('cuda')

    def _graph_key(self):
        return (self.batch_size, id(self.opt_disc), id(self.opt_gen), self.loss, self.gradient_penalty)

    def _capture_graphs(self):
        from .. import _kernels as K
        c = self.config
        st = {'key': self._graph_key(), 'x': torch.empty(self.batch_size, FMAP_SAMPLES, c.res_samples, c.res_samples, device=c.dev)}
        self.opt_disc.prepare_capture(); self.opt_gen.prepare_capture()
        self.disc_model.zero_grad(set_to_none=True); self.gen_model.zero_grad(set_to_none=True)
        for p in self.disc_model.parameters():
            p.requires_grad_(False)
        K.weights_updated()
        gg = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gg):
            st['lg'] = self.gen_step()
        for p in self.disc_model.parameters():
            p.requires_grad_(True)
        K.weights_updated()
        gd = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gd, pool=gg.pool()):
            st['ld'] = self.disc_step(st['x'])
        K.weights_updated()
        self.opt_disc.finish_capture(); self.opt_gen.finish_capture()
        st['gg'], st['gd'] = gg, gd
        return st

    def main_iteration(self, real_batches, num_gen_iters=1):
        """`num_gen_iters` generator steps, then one discriminator step per batch of `real_batches` (reference :544-692); eager
        for the first `warmup_iters` iterations, CUDA-graph replay afterwards when enabled.  Returns (loss_d, loss_g) tensors."""
        if self._graphs_allowed():
            if self._graph is not None and self._graph['key'] != self._graph_key():
                self._graph, self._graph_eager_iters = None, 0
            if self._graph is None and self._graph_eager_iters >= self._graph_warmup:
                self._graph = self._capture_graphs()
            if self._graph is not None:
                g = self._graph
                self.opt_disc.push_lr(); self.opt_gen.push_lr()
                for _ in range(num_gen_iters):
                    g['gg'].replay()
                for xb in real_batches:
                    g['x'].copy_(xb, non_blocking=True)
                    g['gd'].replay()
                K.weights_updated()      # replays rewrite the weights through raw pointers (no torch version bump)
                return g['ld'], g['lg']
            self._graph_eager_iters += 1
        loss_d = loss_g = None
        for p in self.disc_model.parameters():
            p.requires_grad_(False)
        for _ in range(num_gen_iters):
            loss_g = self.gen_step()
        for p in self.disc_model.parameters():
            p.requires_grad_(True)
        for xb in real_batches:
            loss_d = self.disc_step(xb)
        return loss_d, loss_g

    def _set_scheduler(self):
        """reference resnetgan/learner.py:849-864."""
        if self._lr_sched == 'linear decay':
            self.scheduler_fn = lambda main_iter: 1. - (main_iter + self.sched_stop_step) * (1. / self.num_main_iters)
        elif self._lr_sched == 'custom':
            self.scheduler_fn = eval(self.config.lr_sched_custom)
        else:
            raise ValueError("Currently supported LR Schedulers are: [ 'linear decay', 'custom' ]")
        self.scheduler_gen = torch.optim.lr_scheduler.LambdaLR(self.opt_gen, self.scheduler_fn, last_epoch=-1)
        self.scheduler_disc = torch.optim.lr_scheduler.LambdaLR(self.opt_disc, self.scheduler_fn, last_epoch=-1)

    def train(self, train_dl, valid_dl=None, z_valid_dl=None, num_main_iters=None, num_gen_iters=None,
              num_disc_iters=None, log_every=0, step_callback=None):
        """reference resnetgan/learner.py:463-776 (signature kept; valid_dl / z_valid_dl accepted and ignored: metrics are
        outside the hot path).  Per main iteration: `num_gen_iters` generator steps, then `num_disc_iters` discriminator
        steps, LR schedulers stepped once."""
        c = self.config
        num_main_iters = c.num_main_iters if num_main_iters is None else num_main_iters
        num_gen_iters = c.num_gen_iters if num_gen_iters is None else num_gen_iters
        num_disc_iters = c.num_disc_iters if num_disc_iters is None else num_disc_iters
        self.num_main_iters = num_main_iters
        self.dataset_sz = len(train_dl.dataset) if hasattr(train_dl, 'dataset') else None
        self.gen_model.to(c.dev); self.gen_model.train()
        self.disc_model.to(c.dev); self.disc_model.train()
        if self.sched_bool:
            if not self.pretrained_model:
                self.sched_stop_step = 0
            self._set_scheduler()
        if self.not_trained_yet or self.pretrained_model:
            self.train_dataiter = iter(train_dl)
        self.last_losses = (None, None)
        loss_d = loss_g = None
        for itr in range(num_main_iters):
            # ---- generator step(s), then `num_disc_iters` discriminator steps on fresh real batches ----
            reals = []
            for disc_iter in range(num_disc_iters):
                batch = next(self.train_dataiter, None)
                if batch is None:
                    self.curr_epoch_num += 1
                    self.train_dataiter = iter(train_dl)
                    batch = next(self.train_dataiter)
                reals.append(batch[0] if isinstance(batch, (list, tuple)) else batch)
            loss_d, loss_g = self.main_iteration(reals, num_gen_iters)
            self.curr_dataset_batch_num += num_disc_iters
            self.curr_img_num += self.batch_size * num_disc_iters
            self.last_losses = (loss_d, loss_g)
            if step_callback is not None:
                step_callback(itr, loss_d, loss_g)
            if log_every and (itr % log_every == 0):
                print(f'itr {itr:7d}  res {c.res_samples:4d}  D {float(loss_d):9.4g}  G {float(loss_g):9.4g}')
            if self.sched_bool:
                self.scheduler_gen.step()
                self.scheduler_disc.step()
            if self.not_trained_yet:
                self.not_trained_yet = False
            if (itr + 1) % c.num_iters_save_model == 0:          # reference :737-749
                self.gen_model.eval(); self.disc_model.eval()
                self.save_model(c.save_model_dir / (self.model.casefold().replace(' ', '') + '_model.tar'))
                self.gen_model.train(); self.disc_model.train()
        self.gen_model.eval(); self.disc_model.eval()             # reference :771-776 (it also parks them on the CPU)
        return self.last_losses

    # ------------------------------------------------------------------ gradient penalties
    def calc_gp(self, gen_data, real_data):
        """All gradient regularizers (reference resnetgan/learner.py:780-827).  NB the 2-norm is over the channel
        axis only and the mean runs over N*H*W (:820-825) -- matched exactly."""
        gp = self.gradient_penalty
        dev = self.config.dev
        if gp in ('wgan-gp', 'r1'):
            real_data = real_data.view(-1, FMAP_SAMPLES, real_data.shape[2], real_data.shape[3])
        if gp in ('wgan-gp', 'r2'):
            gen_data = gen_data.view(-1, FMAP_SAMPLES, gen_data.shape[2], gen_data.shape[3])
        if gp == 'wgan-gp':
            eps = RANDOM.source.rand((self.batch_size, 1, 1, 1), dev)
            from .. import _kernels as K
            xb = K.interp_rows(gen_data.detach(), real_data.detach(), eps)
        elif gp == 'r1':
            xb = real_data.detach()
        elif gp == 'r2':
            xb = gen_data.detach()
        else:
            raise ValueError(gp)
        xb.requires_grad_(True)
        return self.gp_from_forward(self.disc_model(xb), xb)

    def gp_from_forward(self, outb, xb):
        """Penalty from an existing discriminator forward `outb = D(xb)` (xb requires grad): reference :811-825."""
        gp = self.gradient_penalty
        dev = self.config.dev
        with ops.input_grads_only():
            outb_grads = torch.autograd.grad(outb, xb, grad_outputs=torch.ones(outb.shape[0], device=dev),
                                             create_graph=True, retain_graph=True, only_inputs=True)[0]
        n_hw = outb_grads.shape[0] * outb_grads.shape[2] * outb_grads.shape[3]
        if gp == 'wgan-gp':
            gamma = self.config.gamma
            if gamma != 1.:
                return ops.gp_norm(outb_grads, gamma, self.config.lda / (gamma ** 2 * n_hw))
            return ops.gp_norm(outb_grads, 1., self.config.lda / (2. * n_hw))
        return ops.sumsq(outb_grads, self.config.lda / (2. * n_hw))

    # ------------------------------------------------------------------ validation metrics (reference :318-460 / progan :248-416)
    def _fade_real_for_metrics(self, xb):
        return xb

    def _lagged_for_metrics(self):
        return None

    @torch.no_grad()
    def compute_metrics(self, metrics, metrics_type, z_valid_dl, valid_dl=None):
        """Metric evaluation over a latent validation set (and, for the discriminator, a real one), run periodically by
        train() or by the user: 'fake realness', 'real realness', 'generator loss', 'discriminator loss' (batch means of the
        logit losses without penalty or drift term, weighted by batch length) and 'image grid'.  Returns the reference's
        formatted lines; the raw values are kept in `self.last_metrics`."""
        c = self.config
        if type(self) is GANLearner:
            self._sync_bn_buffers()
        metrics_type = metrics_type.casefold()
        if metrics_type not in ('generator', 'critic', 'discriminator'):
            raise Exception('Invalid metrics_type. Only "generator", "critic", or "discriminator" are accepted.')
        metrics = [metric.casefold() for metric in metrics]
        self.disc_model.eval()
        lagged = None
        if metrics_type == 'generator':
            self.gen_model.train()
            lagged = self._lagged_for_metrics()
            if lagged is not None:
                lagged.eval()
        self.gen_model.eval()
        valid_dataiter = iter(valid_dl) if valid_dl is not None else None
        z_valid_dataiter = iter(z_valid_dl) if z_valid_dl is not None else None
        if z_valid_dl is None:
            raise ValueError('compute_metrics needs a latent validation loader (z_valid_dl)')
        n_batches, n_samples = len(z_valid_dl), len(z_valid_dl.dataset)
        if not self.grid_inputs_constructed and 'image grid' in metrics:
            assert c.img_grid_sz ** 2 <= n_samples
            self.rand_idxs = torch.multinomial(torch.ones(n_samples, dtype=torch.float32), num_samples=c.img_grid_sz ** 2,
                                               replacement=False)
        self._img_grid_constructed = False
        sums = {metric: torch.zeros((), device=c.dev, dtype=torch.float32) for metric in metrics}
        _idx = 0
        for n in range(n_batches):
            zbatch = next(z_valid_dataiter)
            zb = zbatch[0].to(c.dev)
            nb = len(zb)
            _xgenb = self.gen_model(zb)
            fake = None
            if 'fake realness' in metrics:
                fake = self.disc_model(_xgenb)
                sums['fake realness'] += fake.sum()
            if metrics_type == 'generator':
                if 'generator loss' in metrics:
                    _ygenb = fake if fake is not None else self.disc_model(_xgenb)
                    sums['generator loss'] += ops.g_logit_loss(_ygenb, self.loss) * nb
                if 'image grid' in metrics:
                    if self.valid_z is None:
                        self.valid_z = torch.zeros(c.img_grid_sz ** 2, zb.shape[1], device=c.dev)
                    if not self.grid_inputs_constructed:
                        picked = set(int(i) for i in self.rand_idxs)
                        for o in range(nb):
                            if (n * self.batch_size + o) in picked:
                                self.valid_z[_idx] = zb[o]
                                _idx += 1
                        if _idx == c.img_grid_sz ** 2:
                            self.grid_inputs_constructed = True
                    if self.grid_inputs_constructed and not self._img_grid_constructed:
                        base = c.save_samples_dir / self.model.casefold().replace(' ', '') / 'image_grid'
                        if lagged is not None:
                            (base / 'time_averaged').mkdir(parents=True, exist_ok=True)
                            self.make_image_grid(self.valid_z, time_average=True,
                                                 save_path=base / 'time_averaged' / (str(self.gen_metrics_num) + '.png'))
                        (base / 'original').mkdir(parents=True, exist_ok=True)
                        self.make_image_grid(self.valid_z, time_average=False,
                                             save_path=base / 'original' / (str(self.gen_metrics_num) + '.png'))
                        self._img_grid_constructed = True
                self.gen_metrics_num += 1 if n == n_batches - 1 else 0
            else:
                if valid_dl is not None:
                    xb = next(valid_dataiter)[0].to(c.dev)
                    xb = self._fade_real_for_metrics(xb)
                    real = None
                    if 'real realness' in metrics:
                        real = self.disc_model(xb)
                        sums['real realness'] += real.sum()
                    if 'discriminator loss' in metrics:
                        _ygenb = fake if fake is not None else self.disc_model(_xgenb)
                        _yb = real if real is not None else self.disc_model(xb)
                        sums['discriminator loss'] += ops.d_logit_loss(_ygenb, _yb, self.loss, 0.) * nb
                self.disc_metrics_num += 1 if n == n_batches - 1 else 0
        self.last_metrics = {metric: float(v) / n_samples for metric, v in sums.items() if metric != 'image grid'}
        width = '%-' + str(max(len(m) for m in metrics) + 3) + 's'
        lines = ['    ' + (width % (metric + ':')) + '%.4g' % self.last_metrics[metric] + '\n'
                 for metric in metrics if metric != 'image grid']
        self.gen_model.train()
        self.disc_model.train()
        return lines

    @torch.no_grad()
    def make_image_grid(self, zs, labels=None, time_average=True, save_path=None):
        """Grid of generated images for the latent codes `zs` (a perfect-square count), de-normalised with the dataset
        statistics as the reference does before plotting (progan/learner.py:1198-1234).  The reference draws the grid with
        matplotlib; here it is assembled as one uint8 [rows*res, cols*res, 3] array (returned) and written as a PNG."""
        import numpy as np
        if zs.dim() != 2:
            raise IndexError('Incorrect dimensions of input latent vector. Must be `dim == 2`.')
        if zs.shape[1] != self.config.len_latent:
            raise IndexError(f'Input latent vector must be of size {self.config.len_latent}.')
        sz = int(round(np.sqrt(len(zs))))
        if sz * sz != len(zs):
            raise ValueError('Argument `zs` must be a perfect square-length in order to make image grid.')
        net = self._lagged_for_metrics() if time_average else self.gen_model
        if net is None:
            raise ValueError('time_average=True needs the EWMA generator (config.use_ewma_gen)')
        mean = self.ds_mean if self.ds_mean is not None else torch.full((FMAP_SAMPLES, 1, 1), .5)
        std = self.ds_std if self.ds_std is not None else torch.full((FMAP_SAMPLES, 1, 1), .5)
        x = net(zs.to(self.config.dev)).float().cpu()
        x = (x * std + mean).clamp_(0., 1.)                       # imshow clips to [0, 1] as well
        n, ch, h, w = x.shape
        grid = x.view(sz, sz, ch, h, w).permute(0, 3, 1, 4, 2).reshape(sz * h, sz * w, ch)
        grid = (grid * 255.).round().to(torch.uint8).numpy()
        if save_path is not None:
            from pathlib import Path
            from PIL import Image
            Path(save_path).parent.mkdir(parents=True, exist_ok=True)
            Image.fromarray(grid).save(str(save_path))
        return grid

    # ------------------------------------------------------------------ checkpoints (reference :1076-1250)
    def _checkpoint_common(self):
        """Entries every learner writes (reference resnetgan/learner.py:1104-1137) with the reference's value types."""
        from .. import checkpoint as ckpt
        if self.not_trained_yet:
            raise Exception('Please train your model for atleast 1 iteration before saving.')
        c = self.config
        valid_z = self.valid_z if self.valid_z is not None else torch.zeros(c.img_grid_sz ** 2, c.len_latent)
        mean = self.ds_mean if self.ds_mean is not None else torch.full((FMAP_SAMPLES, 1, 1), .5)
        std = self.ds_std if self.ds_std is not None else torch.full((FMAP_SAMPLES, 1, 1), .5)
        return {
            'config': c,
            'gen_model_state_dict': ckpt.plain_state_dict(self.gen_model),
            'disc_model_state_dict': ckpt.plain_state_dict(self.disc_model),
            'nl': ckpt.to_reference_module(self.nl),
            'sched_stop_step': self.sched_stop_step,
            'lr_sched': self.lr_sched,
            'scheduler_gen_state_dict': self.scheduler_gen.state_dict() if self.sched_bool else None,
            'scheduler_disc_state_dict': self.scheduler_disc.state_dict() if self.sched_bool else None,
            'optimizer': self.optimizer,
            'opt_gen_state_dict': self.opt_gen.state_dict(),
            'opt_disc_state_dict': self.opt_disc.state_dict(),
            'loss': self.loss,
            'gradient_penalty': self.gradient_penalty,
            'batch_size': self.batch_size,
            'curr_dataset_batch_num': self.curr_dataset_batch_num,
            'curr_epoch_num': self.curr_epoch_num,
            'tot_num_epochs': self.tot_num_epochs,
            'dataset_sz': self.dataset_sz,
            'ac': self.ac, 'cond_gen': self.cond_gen, 'cond_disc': self.cond_disc,
            'valid_z': valid_z.to('cpu'),
            'valid_label': self.valid_label,
            'grid_inputs_constructed': self.grid_inputs_constructed,
            'rand_idxs': self.rand_idxs,
            'gen_metrics_num': self.gen_metrics_num,
            'disc_metrics_num': self.disc_metrics_num,
            'curr_img_num': self.curr_img_num,
            'not_trained_yet': self.not_trained_yet,
            'ds_mean': mean, 'ds_std': std,
        }

    def _model_metadata(self):
        from .. import checkpoint as ckpt
        return ({'gen_model_upsampler': ckpt.to_reference_module(self.gen_model_upsampler),
                 'num_classes_gen': self.num_classes_gen},
                {'disc_model_downsampler': ckpt.to_reference_module(self.disc_model_downsampler),
                 'num_classes_disc': self.num_classes_disc})

    def _sync_bn_buffers(self):
        """Data parallelism: the generator's BatchNorm uses per-rank batch statistics (standard DDP semantics -- NOT the
        single-GPU result at the N-times larger batch), so its running_mean / running_var buffers drift apart per rank.  Where
        they are observed (checkpoints, validation metrics) they are averaged over the ranks first (collective: every rank
        calls it).  With a per-rank batch that is not a multiple of the minibatch-stddev group the discriminator's groups would
        straddle ranks: refused."""
        dp = self.dp
        if dp is None or getattr(dp, 'world', 1) == 1:
            return
        for net in (self.gen_model, self.disc_model):
            for b in net.buffers():
                if b.is_floating_point():
                    dp.allreduce_mean_(b)

    def save_model(self, save_path):
        """reference resnetgan/learner.py:1076-1137; the file is readable by the reference's own load_model.  Under data
        parallelism only rank 0 writes (parameters are identical on every rank; BatchNorm buffers are averaged first)."""
        from pathlib import Path
        from .. import checkpoint as ckpt
        if type(self) is GANLearner:
            self._sync_bn_buffers()
        if self.dp is not None and getattr(self.dp, 'rank', 0) != 0:
            return
        gmeta, dmeta = self._model_metadata()
        gmeta = {'gen_model': self.gen_model.__class__, 'fmap_g': self._fmap_g, **gmeta}
        dmeta = {'disc_model': self.disc_model.__class__, 'fmap_d': self._fmap_d, **dmeta}
        save_path = Path(save_path)
        save_path.parents[0].mkdir(parents=True, exist_ok=True)
        ckpt.save({**self._checkpoint_common(), 'gen_model_metadata': gmeta, 'disc_model_metadata': dmeta}, save_path)

    def _load_checkpoint_file(self, load_path, dev_of_saved_model):
        from .. import checkpoint as ckpt
        dev_of_saved_model = dev_of_saved_model.casefold()
        assert dev_of_saved_model in ('cpu', 'cuda')
        return ckpt.load(load_path, map_location=(lambda storage, loc: storage) if dev_of_saved_model == 'cpu' else None)

    def _adopt_checkpoint_config(self, checkpoint, dev):
        """`self.config = checkpoint['config']` (reference :1145); `dev` optionally retargets the run (a reference file
        written on a CPU box resumed on a GPU)."""
        from ..utils.custom_layers import as_native_nl
        self.config = checkpoint['config']
        if dev is not None:
            self.config.__dict__['dev'] = torch.device(dev)
        self.nl = as_native_nl(checkpoint['nl'])

    def _load_checkpoint_common(self, checkpoint):
        """reference :1196-1239, in its order (the `optimizer` / `loss` setters rebuild the optimisers / loss first)."""
        # NB on a first load `pretrained_model` is still False here, so the `lr_sched` setter resets the step offset the line
        # before restored (reference :836-847, 1196-1197) -- kept as is
        self.sched_stop_step = checkpoint['sched_stop_step']
        self.lr_sched = checkpoint['lr_sched']
        # the reference keeps the stored scheduler state dicts but overwrites them with a fresh scheduler's before use
        # (progan/learner.py:1055-1062): LambdaLR restarts, `sched_stop_step` carries the offset -- same here
        self._scheduler_gen_state_dict = checkpoint['scheduler_gen_state_dict']
        self._scheduler_disc_state_dict = checkpoint['scheduler_disc_state_dict']
        self.optimizer = checkpoint['optimizer']
        self.opt_gen.load_state_dict(checkpoint['opt_gen_state_dict'])
        self.opt_disc.load_state_dict(checkpoint['opt_disc_state_dict'])
        self.batch_size = checkpoint['batch_size']
        self.loss = checkpoint['loss']
        self.gradient_penalty = checkpoint['gradient_penalty']
        for key in ('curr_dataset_batch_num', 'curr_epoch_num', 'tot_num_epochs', 'dataset_sz', 'ac', 'cond_gen', 'cond_disc',
                    'valid_label', 'grid_inputs_constructed', 'rand_idxs', 'gen_metrics_num', 'disc_metrics_num',
                    'curr_img_num', 'not_trained_yet', 'ds_mean', 'ds_std'):
            setattr(self, key, checkpoint[key])
        if self.ac or self.cond_gen or self.cond_disc:
            raise NotImplementedError('class conditioning / auxiliary classifier are off the benchmarked path; not built')
        self.valid_z = checkpoint['valid_z'].to(self.config.dev)
        self.pretrained_model = True

    def load_model(self, load_path, dev_of_saved_model='cpu', dev=None):
        """reference resnetgan/learner.py:1139-1250: rebuild the networks the file describes, load parameters, optimiser
        state and bookkeeping; `train()` then continues from there.  Reads files written by the reference itself."""
        from ..utils.custom_layers import as_native_upsampler, as_native_pooler
        checkpoint = self._load_checkpoint_file(load_path, dev_of_saved_model)
        self._adopt_checkpoint_config(checkpoint, dev)
        c = self.config
        gmeta, dmeta = checkpoint['gen_model_metadata'], checkpoint['disc_model_metadata']
        self.gen_model_metadata, self.disc_model_metadata = gmeta, dmeta
        self.gen_model_upsampler = as_native_upsampler(gmeta['gen_model_upsampler'])
        self.disc_model_downsampler = as_native_pooler(dmeta['disc_model_downsampler'])
        self.num_classes_gen, self.num_classes_disc = gmeta['num_classes_gen'], dmeta['num_classes_disc']
        self.gen_model = gmeta['gen_model'](len_latent=c.len_latent, fmap=gmeta['fmap_g'], upsampler=self.gen_model_upsampler,
                                            blur_type=c.blur_type, nl=self.nl, num_classes=self.num_classes_gen,
                                            equalized_lr=c.use_equalized_lr)
        self.gen_model.to(c.dev)
        self.gen_model.load_state_dict(checkpoint['gen_model_state_dict'])
        self.gen_model.zero_grad()
        self.disc_model = dmeta['disc_model'](fmap=dmeta['fmap_d'], pooler=self.disc_model_downsampler, blur_type=c.blur_type,
                                              nl=self.nl, num_classes=self.num_classes_disc, equalized_lr=c.use_equalized_lr)
        self.disc_model.to(c.dev)
        self.disc_model.load_state_dict(checkpoint['disc_model_state_dict'])
        self.disc_model.zero_grad()
        assert self.gen_model.res == self.disc_model.res
        self._fmap_g, self._fmap_d = gmeta['fmap_g'], dmeta['fmap_d']
        self._load_checkpoint_common(checkpoint)

    # ------------------------------------------------------------------ properties (reference :831-960)
    @property
    def lr_sched(self):
        return self._lr_sched

    @lr_sched.setter
    def lr_sched(self, new_lr_sched):
        self._lr_sched = None
        self.sched_bool = False
        if not self.pretrained_model:
            self.sched_stop_step = None
        self.scheduler_gen = None
        self.scheduler_disc = None
        if new_lr_sched is not None:
            self._lr_sched = new_lr_sched.casefold()
            self.sched_bool = True
            if not self.pretrained_model:
                self.sched_stop_step = 0

    @property
    def optimizer(self):
        return self._optimizer

    @optimizer.setter
    def optimizer(self, new_optimizer):
        self._optimizer = new_optimizer.casefold()
        self._set_optimizer()

    def _adam_factory(self):
        if self._optimizer != 'adam':
            raise NotImplementedError("Supported Optimizers are: [ 'adam' ] (the reference implements no other)")
        return configure_adam_for_gan(lr_base=self.config.lr_base, betas=(self.config.beta1, self.config.beta2),
                                      eps=self.config.eps, wd=self.config.wd)

    def _set_optimizer(self):
        adam_gan = self._adam_factory()
        self.opt_gen = adam_gan(params=self.gen_model.parameters())
        self.opt_disc = adam_gan(params=self.disc_model.parameters())

    @property
    def gradient_penalty(self):
        return self._gradient_penalty.casefold() if self._gradient_penalty is not None else None

    @gradient_penalty.setter
    def gradient_penalty(self, new_gradient_penalty):
        self._gradient_penalty = new_gradient_penalty.casefold() if new_gradient_penalty is not None else None

    @property
    def loss(self):
        return self._loss

    @loss.setter
    def loss(self, new_loss):
        self._loss = new_loss.casefold()
        self._set_loss()

    def _set_loss(self):
        if self._loss not in ('wgan', 'nonsaturating', 'minimax'):
            raise ValueError("Currently supported Loss Functions are: [ 'wgan', 'nonsaturating', 'minimax' ]")

    @property
    def model(self):
        return self._model

    @model.setter
    def model(self, new_model):
        raise AttributeError(f'{self.__class__.__name__}().model attribute cannot be changed once '
                             f'{self.__class__.__name__} is instantiated.')
