"""Import the UNMODIFIED reference (sidward14/gan-lab) for oracle pinning.  TEST INFRASTRUCTURE ONLY.

Works only where the reference tree exists (the build container: /root/reference).  Nothing in the
`-m gpu` tests, `smoke()` or `bench.py` calls this -- the GPU box has no /root/reference; they use
the committed fixtures under tests/golden/ instead.

Recipe = SURVEY.md Appendix B: stub the three absent off-path deps (matplotlib, indexed, lmdb), put
`<ref>/gan_lab` on sys.path (the reference uses flat imports), give it a private $HOME with a pickled
data_config (train() insists on one, _int.py:55-84).
"""
from __future__ import annotations

import argparse
import os
import pickle
import sys
import tempfile
import types
from collections import OrderedDict
from pathlib import Path

def _find_reference_root() -> Path:
    """$GANLAB_REFERENCE_ROOT, else /root/reference (build container), else <repo>/baseline/_ref (the unmodified reference
    pip-installed there with `pip install --no-index --no-deps --target baseline/_ref /root/reference`; git-ignored, it
    travels to the GPU box with the snapshot and is what `bench.py --impl reference` times)."""
    cands = [os.environ.get("GANLAB_REFERENCE_ROOT"), "/root/reference",
             str(Path(__file__).resolve().parent.parent / "baseline" / "_ref")]
    for c in cands:
        if c and (Path(c) / "gan_lab" / "utils" / "custom_layers.py").exists():
            return Path(c)
    return Path(cands[1])


REFERENCE_ROOT = _find_reference_root()


def reference_available() -> bool:
    return (REFERENCE_ROOT / "gan_lab" / "utils" / "custom_layers.py").exists()


class IndexedOrderedDict(OrderedDict):
    """Stand-in for the absent `indexed` package (progan/learner.py:472; `.values()[i]` is indexed at :228).  It reports
    itself as `indexed.IndexedOrderedDict`, so a checkpoint the reference pickles here names the same class a real
    installation would."""

    def values(self):
        return list(super().values())


IndexedOrderedDict.__module__ = "indexed"
IndexedOrderedDict.__qualname__ = "IndexedOrderedDict"

_loaded = {}


def load_reference():
    """Returns a namespace with the reference's modules; idempotent."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for name in ("matplotlib", "matplotlib.pyplot", "indexed", "lmdb"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib.pyplot"].rcParams = {}

    sys.modules["indexed"].IndexedOrderedDict = IndexedOrderedDict
    ref_pkg = str(REFERENCE_ROOT / "gan_lab")
    if ref_pkg not in sys.path:
        sys.path.insert(0, ref_pkg)

    scratch = Path(tempfile.mkdtemp(prefix="ganlab_ref_home_"))
    (scratch / "cfg").mkdir()
    os.environ["HOME"] = str(scratch)
    (scratch / ".configs_dir.txt").write_bytes(str(scratch / "cfg").encode())
    with open(scratch / "cfg" / ".data_config.p", "wb") as f:
        pickle.dump(argparse.Namespace(ds_mean=[.5] * 3, ds_std=[.5] * 3, dataset="synthetic",
                                       dataset_downsample_type=4), f)   # 4 == PIL.Image.BOX

    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import utils.custom_layers as custom_layers
        import utils.initializer as initializer
        import utils.backprop_utils as backprop_utils
        import utils.latent_utils as latent_utils
        import progan.base as progan_base
        import progan.architectures as progan_arch
        import stylegan.base as stylegan_base
        import stylegan.architectures as stylegan_arch
        import resnetgan.architectures as resnet_arch
        import resnetgan.learner as resnet_learner
        import progan.learner as progan_learner
        import stylegan.learner as stylegan_learner
    _loaded.update(dict(custom_layers=custom_layers, initializer=initializer, backprop_utils=backprop_utils,
                        latent_utils=latent_utils, progan_base=progan_base, progan_arch=progan_arch,
                        stylegan_base=stylegan_base, stylegan_arch=stylegan_arch, resnet_arch=resnet_arch,
                        resnet_learner=resnet_learner, progan_learner=progan_learner,
                        stylegan_learner=stylegan_learner, scratch=scratch))
    return types.SimpleNamespace(**_loaded)


def make_config(model: str = "StyleGAN", res: int = 128, init_res: int = None, batch_size: int = 8,
                dev: str = "cpu", **overrides) -> argparse.Namespace:
    """A config Namespace with every attribute the learners read; values are the defaults of the
    reference's config.py for the given model (config.py:81-325), as listed in SURVEY.md App. B."""
    import torch
    bs = batch_size
    init_res = res if init_res is None else init_res
    scratch = Path(tempfile.gettempdir())
    cfg = dict(
        model=model, dev=torch.device(dev), n_gpu=1, enable_cudnn_autotuner=False, random_seed=0,
        gen_bs_mult=1, num_gen_iters=1, num_disc_iters=1, loss="nonsaturating", gradient_penalty="r1",
        lda=10., gamma=1., lr_sched_custom=None, optimizer="adam", beta1=0., beta2=.99, eps=1e-8, wd=0.,
        lr_base=.001, lr_sched="resolution dependent",
        lr_fctr_dict={4: 1., 8: 1., 16: 1., 32: 1., 64: 1., 128: 1.5, 256: 2., 512: 3., 1024: 3.},
        align_corners=False, model_upsample_type="nearest", model_downsample_type="average",
        latent_distribution="normal", num_classes=0, class_condition=False, use_auxiliary_classifier=False,
        ac_disc_scale=1., ac_gen_scale=.1, num_iters_valid=1000, metrics_dev=torch.device("cpu"),
        gen_metrics=[], disc_metrics=[], img_grid_sz=4, img_grid_show_labels=True,
        save_samples_dir=scratch / "samples", save_model_dir=scratch / "models",
        num_iters_save_model=10 ** 9, num_workers=0, pin_memory=False, batch_size=bs,
        bs_dict={4: bs, 8: bs, 16: bs, 32: bs, 64: bs, 128: bs, 256: bs, 512: bs // 2, 1024: bs // 4},
        nimg_transition=600000, res_samples=res, res_dataset=res, init_res=init_res,
        blur_type="binomial", bit_exact_resampling=False, eps_drift=.001, len_latent=512,
        nonlinearity="leaky relu", leakiness=.2, use_equalized_lr=True, normalize_z=True,
        mbstd_group_size=4, use_ewma_gen=True, num_main_iters=1,
        len_dlatent=512, mapping_num_fcs=8, mapping_lrmul=.01, use_noise=True, use_pixelnorm=False,
        use_instancenorm=True, pct_mixing_reg=.9, beta_trunc_trick=.995, psi_trunc_trick=.7,
        cutoff_trunc_trick=4)
    if model == "ProGAN":
        cfg.update(loss="wgan", gradient_penalty="wgan-gp", use_pixelnorm=True,
                   lr_fctr_dict={4: 1., 8: 1., 16: 1., 32: 1., 64: 1., 128: 1., 256: 1., 512: 1., 1024: 1.5})
    elif model == "ResNet GAN":
        cfg.update(loss="wgan", gradient_penalty="wgan-gp", lr_base=1e-4, lr_sched=None, beta2=.9,
                   blur_type=None, eps_drift=0., len_latent=128, nonlinearity="relu", leakiness=.01,
                   use_equalized_lr=False, num_disc_iters=5)
    cfg.update(overrides)
    return argparse.Namespace(**cfg)
