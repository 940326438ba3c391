"""StyleGANLearner (reference gan_lab/stylegan/learner.py:90-217): StyleGenerator + StyleDiscriminator on the
train loop inherited unchanged from ProGANLearner (reference stylegan/learner.py:90)."""
import torch

from ..progan.learner import ProGANLearner, NONREDEFINABLE_ATTRS as _PRO_ATTRS
from .architectures import StyleGenerator, StyleDiscriminator

NONREDEFINABLE_ATTRS = _PRO_ATTRS + ('use_instancenorm', 'use_noise', 'pct_mixing_reg', 'beta_trunc_trick',
                                     'psi_trunc_trick', 'cutoff_trunc_trick', 'len_dlatent', 'mapping_num_fcs',
                                     'mapping_lrmul',)


class StyleGANLearner(ProGANLearner):
    _model_name = 'StyleGAN'

    def _nonredefinable(self):
        return NONREDEFINABLE_ATTRS

    def _build_models(self):
        c = self.config
        self.gen_model = StyleGenerator(
            final_res=c.res_samples, latent_distribution=self.latent_distribution, len_latent=c.len_latent,
            len_dlatent=c.len_dlatent, mapping_num_fcs=c.mapping_num_fcs, mapping_lrmul=c.mapping_lrmul,
            use_instancenorm=c.use_instancenorm, use_noise=c.use_noise, upsampler=self.gen_model_upsampler,
            blur_type=c.blur_type, nl=self.nl, num_classes=0, equalized_lr=c.use_equalized_lr, normalize_z=c.normalize_z,
            use_pixelnorm=c.use_pixelnorm, pct_mixing_reg=c.pct_mixing_reg,
            truncation_trick_params={'beta': c.beta_trunc_trick, 'psi': c.psi_trunc_trick,
                                     'cutoff_stage': c.cutoff_trunc_trick},
            state=self.state)
        self.disc_model = StyleDiscriminator(
            final_res=c.res_samples, pooler=self.disc_model_downsampler, blur_type=c.blur_type, nl=self.nl, num_classes=0,
            equalized_lr=c.use_equalized_lr, mbstd_group_size=c.mbstd_group_size, state=self.state)

    # ---- checkpoint entries StyleGAN adds (reference stylegan/learner.py:456-464, 545-553, 604-607)
    def _extra_checkpoint_entries(self):
        g = self.gen_model
        w = g.w_ewma if g.w_ewma is not None else torch.zeros(self.config.len_dlatent)
        # the reference re-copies the lagged generator from the live one right before it saves (progan/learner.py:968-971), so the
        # `w_ewma_lagged` it writes is the live average
        w_lag = w.detach().to('cpu').clone() if self.config.use_ewma_gen else None
        return {'use_truncation_trick': g.use_truncation_trick, 'trunc_cutoff_stage': g.trunc_cutoff_stage,
                'w_eval_psi': g.w_eval_psi, 'w_ewma_beta': g.w_ewma_beta, 'w_ewma': w.detach().to('cpu').clone(),
                'w_ewma_lagged': w_lag, 'trained_with_noise': g._trained_with_noise, 'pct_mixing_reg': g.pct_mixing_reg}

    def _restore_extra_checkpoint_entries(self, checkpoint):
        g = self.gen_model
        g.w_ewma_beta = checkpoint['w_ewma_beta']
        g._w_eval_psi = checkpoint['w_eval_psi']
        g._trunc_cutoff_stage = checkpoint['trunc_cutoff_stage']
        g.use_truncation_trick = checkpoint['use_truncation_trick']
        g.w_ewma = checkpoint['w_ewma'].to(self.config.dev)
        g._trained_with_noise = checkpoint['trained_with_noise']
        g._use_noise = checkpoint['trained_with_noise']
        g.pct_mixing_reg = checkpoint['pct_mixing_reg']
        g._use_mixing_reg = True if g.pct_mixing_reg else False

    def _update_gen_lagged(self):
        """The lagged generator's parameters are always current here; what the reference's re-copy (progan/learner.py:235-242)
        additionally refreshes is its copy of the live generator's `w_ewma` (used by the truncation trick in eval mode)."""
        lag = super(StyleGANLearner, self)._update_gen_lagged()
        if lag is not None and self.gen_model.w_ewma is not None:
            lag.w_ewma = self.gen_model.w_ewma.detach().clone()
        return lag

    def _sync_replicas(self):
        """Under data parallelism every rank's generator averages the w of ITS latents into `w_ewma`
        (stylegan/architectures.py:427-437).  The update is linear, so the mean over ranks of the rank-local averages IS the
        average the reference would hold at the global batch -- at any time; it is therefore taken lazily, where `w_ewma` is
        observed (end of train(), checkpoints, validation metrics), not in every step."""
        dp, w = self.dp, self.gen_model.w_ewma
        if dp is not None and dp.world > 1 and w is not None:
            dp.allreduce_mean_(w)

    def _restore_lagged_extras(self, checkpoint):
        # literal: the reference assigns the stored LAGGED average to the live generator here (stylegan/learner.py:607)
        self.gen_model.w_ewma = checkpoint['w_ewma_lagged'].to(self.config.dev)

    @property
    def latent_distribution(self):
        return self._latent_distribution

    @latent_distribution.setter
    def latent_distribution(self, new_latent_distribution):
        self._latent_distribution = new_latent_distribution.casefold()
        if getattr(self, 'gen_model', None) is not None:
            self.gen_model.latent_distribution = self._latent_distribution
            if self.config.use_ewma_gen and getattr(self, 'gen_model_lagged', None) is not None:
                self.gen_model_lagged.latent_distribution = self._latent_distribution
