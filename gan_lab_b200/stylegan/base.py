"""Base class of the StyleGAN family (reference gan_lab/stylegan/base.py:21-173)."""
from .._growth import ProgressiveBase, GrowthState, FMAP_BASE, FMAP_MAX  # noqa: F401


class StyleGAN(ProgressiveBase):
    _family_root = True
    _default_state = None
