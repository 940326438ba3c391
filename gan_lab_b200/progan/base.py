"""Base class of the ProGAN family (reference gan_lab/progan/base.py:21-173)."""
from .._growth import ProgressiveBase, GrowthState, FMAP_BASE, FMAP_MAX  # noqa: F401


class ProGAN(ProgressiveBase):
    _family_root = True
    _default_state = None
