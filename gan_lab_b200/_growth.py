"""Progressive-growing state shared by a generator/discriminator pair.

The reference keeps alpha / curr_res / fmap / fade_in_phase ... as *class-level* attributes of the `ProGAN` /
`StyleGAN` base classes so that G and D see the same values (gan_lab/progan/base.py:21-170,
gan_lab/stylegan/base.py:21-173), and its learners re-create those classes with `type(...)` to get fresh
state (stylegan/learner.py:114-118).  Here the state is an explicit `GrowthState` object: modules built
without one share their family's default state (same observable behaviour as the reference's class
attributes); a Learner passes one fresh object to both networks.
"""
import numpy as np
from torch import nn

FMAP_BASE = 8192
FMAP_MAX = 512


class GrowthState(object):
    def __init__(self, fmap_base=None, fmap_max=None):
        self.fmap_base = FMAP_BASE if fmap_base is None else fmap_base
        self.fmap_max = FMAP_MAX if fmap_max is None else fmap_max
        self.final_res = None
        self.alpha_dev = None          # _kernels.DeviceAlpha once a learner asks for CUDA-graph replay of fade-in phases
        self.reset()

    def get_fmap(self, scale_stage):
        return min(int(self.fmap_base / (2 ** scale_stage)), self.fmap_max)

    def reset(self):
        self.alpha = 1
        self.alpha_tol = 1.e-8
        self.prev_res = None
        self.curr_res = 4
        if self.final_res is not None:
            assert self.curr_res <= self.final_res
        self.scale_stage = int(np.log2(self.curr_res)) - 1
        self.fmap_prev = None
        self.fmap = self.get_fmap(self.scale_stage)
        self.scale_inc_metadata_updated = False
        self.fade_in_phase = False

    def as_dict(self):
        return dict(self.__dict__)

    def enable_device_alpha(self, device):
        from . import _kernels as K
        if self.alpha_dev is None or self.alpha_dev.coef.device != device:
            self.alpha_dev = K.DeviceAlpha(device)
        self.alpha_dev.set(self.alpha)

    def blend_coefs(self):
        """(alpha, 1 - alpha) for the fade-in blends: floats, or handles of the device-side vector when that is enabled."""
        if self.alpha_dev is not None:
            return self.alpha_dev.alpha, self.alpha_dev.one_minus
        return self.alpha, 1. - self.alpha


def _state_property(name):
    def getter(self):
        return getattr(self._state, name)

    def setter(self, value):
        setattr(self._state, name, value)

    return property(getter, setter)


class ProgressiveBase(nn.Module):
    """Metadata that defines the current state of a progressively grown model (see module docstring)."""

    _default_state = None

    @classmethod
    def default_state(cls):
        # one default per family root (ProGAN / StyleGAN), shared by its G and D subclasses
        for klass in cls.__mro__:
            if '_family_root' in klass.__dict__:
                if klass.__dict__.get('_default_state') is None:
                    klass._default_state = GrowthState()
                return klass._default_state
        raise TypeError('not a ProGAN/StyleGAN family class')

    @classmethod
    def reset_state(cls):
        """Call this to reset the shared state in case one wants to start over (reference base.py:40-54)."""
        cls.default_state().reset()

    def __init__(self, final_res, state=None):
        super(ProgressiveBase, self).__init__()
        object.__setattr__(self, '_state', state if state is not None else type(self).default_state())
        self.final_res = final_res
        assert self.curr_res <= self.final_res

    @property
    def cls_base(self):
        """The reference exposes the state holder as `cls_base` and compares `cls_base.__dict__` of G and D."""
        return self._state

    def increase_scale(self):
        """Use this to increase scale during training or for initial resolution (reference base.py:62-74)."""
        s = self._state
        s.prev_res = s.curr_res
        s.curr_res = int(2 ** (int(np.log2(s.curr_res)) + 1))
        s.scale_stage = int(np.log2(s.curr_res)) - 1
        s.fmap_prev = s.fmap
        s.fmap = self.get_fmap(scale_stage=s.scale_stage)
        s.scale_inc_metadata_updated = True
        s.fade_in_phase = True

    def get_fmap(self, scale_stage):
        return self._state.get_fmap(scale_stage)

    def most_parameters(self, recurse=True, excluded_params: list = []):
        """`parameters()` with the option to exclude named parameters (reference base.py:79-83)."""
        for name, params in self.named_parameters(recurse=recurse):
            if name not in excluded_params:
                yield params

    fade_in_phase = _state_property('fade_in_phase')
    scale_inc_metadata_updated = _state_property('scale_inc_metadata_updated')
    fmap = _state_property('fmap')
    fmap_prev = _state_property('fmap_prev')
    scale_stage = _state_property('scale_stage')
    curr_res = _state_property('curr_res')
    final_res = _state_property('final_res')
    prev_res = _state_property('prev_res')
    alpha_tol = _state_property('alpha_tol')

    @property
    def alpha(self):
        return self._state.alpha

    @alpha.setter
    def alpha(self, new_alpha):
        """reference base.py:157-170."""
        if not (0. <= new_alpha < 1. + self.alpha_tol):
            raise ValueError('Input alpha parameter must be in the range [0,1].')
        if 1. - self.alpha_tol < new_alpha < 1. + self.alpha_tol:
            self._state.fade_in_phase = False
            self._state.alpha = 1
        else:
            self._state.alpha = new_alpha

    def forward(self, x):
        raise NotImplementedError('Can only call `forward` on valid subclasses.')
