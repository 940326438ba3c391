"""Checkpoint files interchangeable with the reference's.

The reference's `save_model` (progan/learner.py:1238-1298, stylegan/learner.py:433-506, resnetgan/learner.py:1076-1137)
`torch.save`s one dict whose pickle stream names classes of the reference's own flat modules -- `_int.LearnerConfigCopy`
(the frozen config), `indexed.IndexedOrderedDict` (`lagged_params`), `utils.custom_layers.*` poolers,
`resnetgan.architectures.*` network classes -- next to stock torch modules (`nn.Upsample`, `nn.AvgPool2d`,
`nn.LeakyReLU`) and torch.optim.Adam state dicts.  This module reads and writes that format without the reference
being importable:

  * `load(path)`  unpickles with a class map from the reference's module paths to this package's drop-in classes;
  * `save(obj, path)` pickles this package's classes UNDER the reference's names (a `pickle._Pickler` whose
    `save_global` writes the mapped module/name), so the reference's own `load_model` reads the file.

Tensors travel through torch's zip serialisation unchanged (storages, strides, devices).  `torch.load` needs
`weights_only=False` for these files (they carry objects, not just tensors) -- as the reference's files always did.
"""
import pickle
from collections import OrderedDict

import torch


class IndexedOrderedDict(OrderedDict):
    """What `lagged_params` is pickled as (reference progan/learner.py:472); `values()` is indexable (:228)."""

    def values(self):
        return list(super().values())


def _class_map():
    """(module, name) in a reference pickle stream  ->  class of this package.  Built lazily (import cycles)."""
    from .resnetgan import architectures as ra
    from .resnetgan.learner import LearnerConfigCopy
    from .utils import custom_layers as cl
    m = {('_int', 'LearnerConfigCopy'): LearnerConfigCopy,
         ('indexed', 'IndexedOrderedDict'): IndexedOrderedDict}
    for name in ('NearestPool2d', 'BilinearPool2d', 'Lambda'):
        m[('utils.custom_layers', name)] = getattr(cl, name)
    for name in ('Generator32PixResnet', 'Generator64PixResnet', 'Discriminator32PixResnet', 'Discriminator64PixResnet'):
        m[('resnetgan.architectures', name)] = getattr(ra, name)
    return m


def _strip_package(mod_name):
    """The reference's modules import each other flat (`from _int import ...`); a copy imported as a package would
    pickle `gan_lab._int` instead."""
    return mod_name[len('gan_lab.'):] if mod_name.startswith('gan_lab.') else mod_name


class _Unpickler(pickle.Unpickler):
    def find_class(self, mod_name, name):
        hit = _class_map().get((_strip_package(mod_name), name))
        if hit is not None:
            return hit
        return super().find_class(mod_name, name)


class _Pickler(pickle._Pickler):
    """Pure-Python pickler (the C one cannot be told to write a global under another name).  Only the object skeleton goes
    through it; tensor payloads are written by torch as zip records."""

    def __init__(self, *args, **kwargs):
        super(_Pickler, self).__init__(*args, **kwargs)
        self._renames = {cls: key for key, cls in _class_map().items()}

    def save_global(self, obj, name=None):
        ref_name = self._renames.get(obj)
        if ref_name is None:
            return super().save_global(obj, name)
        module, qualname = ref_name
        if self.proto >= 4:
            self.save(module)
            self.save(qualname)
            self.write(pickle.STACK_GLOBAL)
        else:
            self.write(pickle.GLOBAL + module.encode('utf-8') + b'\n' + qualname.encode('utf-8') + b'\n')
        self.memoize(obj)


class _PickleModule(object):
    """The `pickle_module` handed to torch.save / torch.load."""
    __name__ = 'gan_lab_b200.checkpoint'
    Unpickler = _Unpickler
    Pickler = _Pickler

    @staticmethod
    def load(f, **kw):
        return _Unpickler(f, **kw).load()

    @staticmethod
    def dump(obj, f, protocol=None):
        _Pickler(f, protocol).dump(obj)


def load(path, map_location=None):
    """Read a checkpoint written by the reference or by `save()`."""
    return torch.load(path, map_location=map_location, weights_only=False, pickle_module=_PickleModule)


def save(obj, path):
    """Write `obj` so that both `load()` and the reference's `torch.load` can read it."""
    torch.save(obj, path, pickle_module=_PickleModule)


# ------------------------------------------------------------------------------------------------ module translation
def to_reference_module(m):
    """This package's shared upsampler / pooler / nonlinearity instance -> the stock torch module the reference holds in
    the same slot (resnetgan/learner.py:154-181), so the reference can rebuild its networks from the file."""
    from torch import nn
    from .utils import custom_layers as cl
    if isinstance(m, cl.Upsample2x):
        return nn.Upsample(scale_factor=2, mode='nearest')
    if isinstance(m, cl.AvgPool2x):
        return nn.AvgPool2d(kernel_size=2, stride=2)
    if isinstance(m, cl.ReLU):
        return nn.ReLU()
    if isinstance(m, cl.LeakyReLU):
        return nn.LeakyReLU(negative_slope=m.negative_slope)
    return m      # NearestPool2d / BilinearPool2d pickle under the reference's names


def plain_state_dict(module):
    """state_dict as contiguous CPU-layout-agnostic tensors: parameters live in channels_last memory here; the values and
    logical (OIHW) shapes are the reference's, and `load_state_dict` on either side copies element-wise."""
    return OrderedDict((k, v.detach().clone(memory_format=torch.contiguous_format)) for k, v in module.state_dict().items())
