"""StyleGANLearner (reference gan_lab/stylegan/learner.py:90-217): StyleGenerator + StyleDiscriminator on the
train loop inherited unchanged from ProGANLearner (reference stylegan/learner.py:90)."""
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
