"""Base class of the non-progressive (ResNet) GAN architectures (reference gan_lab/resnetgan/base.py:15-36)."""
from torch import nn


class GAN(nn.Module):
    """Base GAN for all non-progressive architectures: a fixed resolution and `most_parameters`."""

    def __init__(self, res):
        super(GAN, self).__init__()
        self._res = res

    def most_parameters(self, recurse=True, excluded_params: list = []):
        """nn.Module.parameters() with the option to exclude parameters by name."""
        for name, params in self.named_parameters(recurse=recurse):
            if name not in excluded_params:
                yield params

    @property
    def res(self):
        return self._res

    @res.setter
    def res(self, new_res):
        raise AttributeError(f'GAN().res cannot be changed, as {self.__class__.__name__} only permits one resolution: '
                             f'{self._res}.')

    def forward(self, x):
        raise NotImplementedError('Can only call `forward` on valid subclasses.')
