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

    @torch.no_grad()
    def make_stylemixing_grid(self, zs_sourceb, zs_coarse=(), zs_middle=(), zs_fine=(), labels=None, time_average=True,
                              save_path=None):
        """Style-mixed grid in the layout of Figure 3 of Karras et al. 2019 (reference stylegan/learner.py:306-431): first row =
        the source-B images, first column = the source-A images, cell (a, b) = source A with source B's styles mixed in from
        stage 1 (coarse rows), 4 (middle rows) or 8 (fine rows) on.  The reference draws it with matplotlib; here the cells are
        assembled into one uint8 [(1+rows)*res, (1+cols)*res, 3] array (white top-left corner), returned and optionally saved."""
        from ..resnetgan.learner import FMAP_SAMPLES
        groups = [(torch.as_tensor(zs), stage) for zs, stage in ((zs_coarse, 1), (zs_middle, 4), (zs_fine, 8)) if len(zs)]
        assert groups, 'give at least one of zs_coarse / zs_middle / zs_fine'
        net = self._lagged_for_metrics() if time_average else self.gen_model
        if net is None:
            raise ValueError('time_average=True needs the EWMA generator (config.use_ewma_gen)')
        was_training = net.training
        net.eval()
        dev = self.config.dev
        zb = torch.as_tensor(zs_sourceb).reshape(-1, self.config.len_latent).to(dev)
        mean = self.ds_mean if self.ds_mean is not None else torch.full((FMAP_SAMPLES, 1, 1), .5)
        std = self.ds_std if self.ds_std is not None else torch.full((FMAP_SAMPLES, 1, 1), .5)

        def cell(x):                    # [n, 3, H, W] -> uint8 [n, H, W, 3]
            x = (x.float().cpu() * std + mean).clamp_(0., 1.)
            return (x * 255.).round().to(torch.uint8).permute(0, 2, 3, 1)

        top = cell(net(zb))
        res = top.shape[1]
        rows = [torch.cat([torch.full((res, res, 3), 255, dtype=torch.uint8)] + list(top), dim=1)]
        for zs, stage in groups:
            za = zs.reshape(-1, self.config.len_latent).to(dev)
            left = cell(net(za))
            for i in range(za.shape[0]):
                mixed = cell(net(za[i:i + 1].expand(zb.shape[0], -1).contiguous(), x_mixing=zb, style_mixing_stage=stage))
                rows.append(torch.cat([left[i]] + list(mixed), dim=1))
        if was_training:            # (not net.train(False): as in the reference, that would switch mixing regularisation on)
            net.train()
        grid = torch.cat(rows, dim=0).numpy()
        if save_path is not None:
            from pathlib import Path
            from PIL import Image
            Path(save_path).parent.mkdir(parents=True, exist_ok=True)
            Image.fromarray(grid).save(str(save_path))
        return grid

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
