"""Latent sampling (reference gan_lab/utils/latent_utils.py:12-31) behind a replaceable random source.

The reference draws z, the style-mixing latent, the per-layer noise, the mixing cut-off and the WGAN-GP
eps inside forward()/train() from global RNGs.  Here every draw goes through `RANDOM` so that parity tests
(and DP ranks) can substitute their own stream; the default draws on the device with torch's generators.
"""
import numpy as np
import torch


class RandomSource(object):
    """Default source: device-side torch RNG for tensors, host RNGs for the mixing decision (as the reference)."""

    def randn(self, shape, device):
        return torch.randn(*shape, dtype=torch.float32, device=device)

    def rand(self, shape, device):
        return torch.rand(*shape, dtype=torch.float32, device=device)

    def host_uniform(self):
        return float(np.random.rand())

    def host_randint(self, lo, hi):
        return int(torch.randint(lo, hi, (1,)).item())


class TapeSource(RandomSource):
    """Replays recorded draws by kind and shape (FIFO per (kind, shape)); used by the parity tests."""

    def __init__(self, events, device):
        self.device = device
        self.q = {}
        for kind, val in events:
            key = (kind, tuple(val.shape)) if torch.is_tensor(val) else (kind, None)
            self.q.setdefault(key, []).append(val)

    def _pop(self, kind, shape):
        lst = self.q.get((kind, tuple(shape) if shape is not None else None))
        if not lst:
            raise RuntimeError(f'TapeSource exhausted for {kind} {shape}')
        return lst.pop(0)

    def randn(self, shape, device):
        return self._pop('randn', shape).to(device=device, dtype=torch.float32)

    def rand(self, shape, device):
        return self._pop('rand', shape).to(device=device, dtype=torch.float32)

    def host_uniform(self):
        return float(self._pop('np_rand', None))

    def host_randint(self, lo, hi):
        return int(self._pop('randint', (1,)).item())


class _Holder(object):
    source = RandomSource()


RANDOM = _Holder()


def set_random_source(src):
    old = RANDOM.source
    RANDOM.source = src if src is not None else RandomSource()
    return old


def gen_rand_latent_vars(num_samples, length, distribution='normal', device='cuda'):
    if distribution == 'normal':
        return RANDOM.source.randn((num_samples, length), device)
    elif distribution == 'uniform':
        return RANDOM.source.rand((num_samples, length), device)
    raise ValueError(distribution)


def concat_rand_classes_to_z(z, num_classes, z_labels=None, device='cuda'):
    if z_labels is None:
        z_labels = torch.randint(0, num_classes, (len(z), 1), dtype=torch.int64, device=device)
    assert z_labels.shape == (len(z), 1)
    labels_one_hot = torch.zeros(len(z), num_classes, dtype=torch.float32, device=device)
    labels_one_hot.scatter_(1, z_labels, 1)
    return torch.cat((z, labels_one_hot), dim=1)
