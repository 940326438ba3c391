"""Runtime He/Xavier scale of the equalized-LR layers (host-side scalar math).

Mirrors the semantics of the reference's `Initializer` (gan_lab/utils/initializer.py:16-80): same class
name, constructor and `get_init_bound_layer` signature, same numbers bit-for-bit (they are Python floats).
"""
import math


class Initializer(object):
    def __init__(self, init, init_type='default', gain_sq_base=2., equalized_lr=False):
        self.init = init.casefold()
        self.init_type = init_type.casefold()
        self.gain_sq_base = gain_sq_base
        self.equalized_lr = equalized_lr

    def get_init_bound_layer(self, tensor, distribution_type, stride=1):
        """Std (normal) or bound (uniform = sqrt(3)*std) for `tensor` (reference initializer.py:27-39)."""
        distribution_type = distribution_type.casefold()
        if distribution_type not in ('uniform', 'normal'):
            raise ValueError('Only uniform and normal distributions are supported.')
        fan_in, fan_out = self._fans(tensor, stride)
        if self.init_type in ('progan', 'stylegan'):
            fan_out = None                       # fan-in only (initializer.py:35-37)
        std = self._std(fan_in, fan_out)
        return math.sqrt(3) * std if distribution_type == 'uniform' else std

    def _fans(self, tensor, stride=1):
        """reference initializer.py:42-62."""
        if tensor.dim() < 2:
            raise ValueError('Fan in and fan out cannot be computed for tensor with fewer than 2 dimensions.')
        if tensor.dim() == 2:
            return tensor.size(1), tensor.size(0)
        rf = 1
        for d in tensor.shape[2:]:
            rf *= d
        return tensor.size(1) * rf, tensor.size(0) * rf / stride ** 2

    def _std(self, fan_in=None, fan_out=None):
        """reference initializer.py:65-80."""
        gain_sq = self.gain_sq_base / 2.
        if fan_out is not None and fan_in is not None:
            fan = fan_in + fan_out
            gain_sq *= 2
        elif fan_in is not None:
            fan = fan_in
        else:
            fan = fan_out
        if self.init == 'he':
            gain_sq = 2. * gain_sq
        elif self.init == 'xavier':
            gain_sq = 1. * gain_sq
        return math.sqrt(gain_sq / fan)
