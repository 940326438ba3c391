"""Default configuration Namespaces (host-side only).

The reference builds its config with argparse and pickles it (gan_lab/config.py:48-412); the learners only need a
Namespace with the same attribute names.  `default_config()` returns the reference's defaults for each model
(config.py:81-325) so that scripts, tests and the benchmark do not depend on the pickled-file plumbing.
"""
import argparse
import tempfile
from pathlib import Path

import torch


def default_config(model='StyleGAN', res=128, init_res=None, batch_size=8, dev='cuda', **overrides):
    bs = batch_size
    init_res = res if init_res is None else init_res
    scratch = Path(tempfile.gettempdir())
    cfg = dict(
        model=model, dev=torch.device(dev), n_gpu=1, enable_cudnn_autotuner=False, random_seed=0,
        gen_bs_mult=1, num_gen_iters=1, num_disc_iters=1, loss='nonsaturating', gradient_penalty='r1',
        lda=10., gamma=1., lr_sched_custom=None, optimizer='adam', beta1=0., beta2=.99, eps=1e-8, wd=0.,
        lr_base=.001, lr_sched='resolution dependent',
        lr_fctr_dict={4: 1., 8: 1., 16: 1., 32: 1., 64: 1., 128: 1.5, 256: 2., 512: 3., 1024: 3.},
        align_corners=False, model_upsample_type='nearest', model_downsample_type='average',
        latent_distribution='normal', num_classes=0, class_condition=False, use_auxiliary_classifier=False,
        ac_disc_scale=1., ac_gen_scale=.1, num_iters_valid=1000, metrics_dev=torch.device(dev),
        gen_metrics=[], disc_metrics=[], img_grid_sz=4, img_grid_show_labels=True,
        save_samples_dir=scratch / 'samples', save_model_dir=scratch / 'models',
        num_iters_save_model=10 ** 9, num_workers=0, pin_memory=False, batch_size=bs,
        bs_dict={4: bs, 8: bs, 16: bs, 32: bs, 64: bs, 128: bs, 256: bs, 512: bs // 2, 1024: bs // 4},
        nimg_transition=600000, res_samples=res, res_dataset=res, init_res=init_res,
        blur_type='binomial', bit_exact_resampling=False, eps_drift=.001, len_latent=512,
        nonlinearity='leaky relu', leakiness=.2, use_equalized_lr=True, normalize_z=True,
        mbstd_group_size=4, use_ewma_gen=True, num_main_iters=1,
        len_dlatent=512, mapping_num_fcs=8, mapping_lrmul=.01, use_noise=True, use_pixelnorm=False,
        use_instancenorm=True, pct_mixing_reg=.9, beta_trunc_trick=.995, psi_trunc_trick=.7,
        cutoff_trunc_trick=4)
    if model == 'ProGAN':
        cfg.update(loss='wgan', gradient_penalty='wgan-gp', use_pixelnorm=True,
                   lr_fctr_dict={4: 1., 8: 1., 16: 1., 32: 1., 64: 1., 128: 1., 256: 1., 512: 1., 1024: 1.5})
    elif model == 'ResNet GAN':
        cfg.update(loss='wgan', gradient_penalty='wgan-gp', lr_base=1e-4, lr_sched=None, beta2=.9,
                   blur_type=None, eps_drift=0., len_latent=128, nonlinearity='relu', leakiness=.01,
                   use_equalized_lr=False, num_disc_iters=5)
    cfg.update(overrides)
    return argparse.Namespace(**cfg)
