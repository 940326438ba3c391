"""GANLearner: base learner (reference gan_lab/resnetgan/learner.py:88-946).

Keeps the reference's constructor `(config: argparse.Namespace)`, its public attributes (`gen_model`,
`disc_model`, `opt_gen`, `opt_disc`, `batch_size`, `config`, `loss`, `gradient_penalty`, `optimizer`,
`lr_sched`) and `calc_gp`; metrics, plotting and checkpoint I/O are outside
the hot path and not built.  The ResNet-GAN architectures themselves (BatchNorm / LayerNorm blocks) are a
later row of the scope table; this class is the shared machinery the ProGAN / StyleGAN learners build on.
"""
import argparse
import copy

import torch

from .. import ops
from ..optim import FusedAdam
from ..utils.custom_layers import Upsample2x, AvgPool2x, NearestPool2d, BilinearPool2d, LeakyReLU, ReLU
from ..utils.latent_utils import RANDOM

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
        hidden = ('_nonredefinable_attrs', '_redefinable_from_learner_attrs', 'learner_class')
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
        self.not_trained_yet = True

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

    # ------------------------------------------------------------------ properties (reference :831-960)
    @property
    def lr_sched(self):
        return self._lr_sched

    @lr_sched.setter
    def lr_sched(self, new_lr_sched):
        self._lr_sched = None
        self.sched_bool = False
        self.scheduler_gen = None
        self.scheduler_disc = None
        if new_lr_sched is not None:
            self._lr_sched = new_lr_sched.casefold()
            self.sched_bool = True
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
