"""Device-side real-image loader.

The reference feeds train() from a torch DataLoader whose dataset resizes every sample on the host with PIL
(`Resize(curr_res, BOX) -> ToTensor -> Normalize`, data_config.py:312-342) and rewrites that Resize whenever the
networks grow (progan/learner.py:611-612, 1099-1112); at 8 B200s that host pipeline cannot keep up with the train step.
`DeviceImageLoader` keeps the decoded uint8 images (HWC RGB, as PIL decodes them) resident in HBM -- or in pinned host
memory, shipping each batch as uint8, 4x less H2D traffic than fp32 -- and produces every batch with one kernel
(`_kernels.u8_box_resize_normalize`, bit-exact with the PIL chain).  It offers what the learners use of a DataLoader:
iteration yielding `(xb,)`, `.dataset` (len), `.batch_sampler.batch_size` (train() rewrites it when the batch size changes
with the resolution) and `set_resolution(res)` (called by train() at every resolution increase and on a resume).
"""
import torch

from . import _kernels as K


class _BatchSampler(object):
    def __init__(self, batch_size, drop_last):
        self.batch_size = int(batch_size)
        self.drop_last = bool(drop_last)


class DeviceImageLoader(object):
    def __init__(self, images_u8, batch_size, res, mean=(.5, .5, .5), std=(.5, .5, .5), shuffle=True, mirror=False,
                 drop_last=True, device='cuda', seed=0, rank=0, world_size=1):
        """images_u8: uint8 [M, H, W, 3].  A CUDA tensor stays where it is (samples are gathered by the kernel); a CPU tensor
        is pinned and each batch travels as uint8.  rank / world_size: this rank's strided shard of every epoch's order."""
        if images_u8.dtype != torch.uint8 or images_u8.dim() != 4 or images_u8.shape[3] != 3:
            raise ValueError('images_u8 must be uint8 [M, H, W, 3]')
        self.device = torch.device(device)
        if images_u8.is_cuda:
            self.images = images_u8.contiguous()
        else:
            images_u8 = images_u8.contiguous()
            self.images = images_u8.pin_memory() if self.device.type == 'cuda' and torch.cuda.is_available() else images_u8
        self.dataset = self.images                      # len(loader.dataset) as the learners read it
        self.batch_sampler = _BatchSampler(batch_size, drop_last)
        self.res = int(res)
        self.mean, self.std = tuple(float(v) for v in mean), tuple(float(v) for v in std)
        self.shuffle, self.mirror = bool(shuffle), bool(mirror)
        self.rank, self.world_size = int(rank), int(world_size)
        self._gen = torch.Generator().manual_seed(int(seed))

    def set_resolution(self, res):
        self.res = int(res)

    def __len__(self):
        n = len(range(self.rank, self.images.shape[0], self.world_size))
        bs = self.batch_sampler.batch_size
        return n // bs if self.batch_sampler.drop_last else (n + bs - 1) // bs

    def __iter__(self):
        M = self.images.shape[0]
        order = torch.randperm(M, generator=self._gen) if self.shuffle else torch.arange(M)
        order = order[self.rank::self.world_size]
        pos = 0
        while pos < order.numel():
            bs = self.batch_sampler.batch_size           # read every batch: train() changes it as the networks grow
            idx = order[pos:pos + bs]
            pos += bs
            if idx.numel() < bs and self.batch_sampler.drop_last:
                return
            flip = (torch.rand(idx.numel(), generator=self._gen) < 0.5) if self.mirror else None
            yield (self.batch(idx, flip),)

    def batch(self, idx, flip=None):
        """Samples `idx` (int64 CPU tensor) at the current resolution -> fp32 [N, 3, res, res] on the device."""
        if self.images.is_cuda:
            return K.u8_box_resize_normalize(self.images, idx, (self.res, self.res), self.mean, self.std, flip)
        staged = self.images[idx].to(self.device, non_blocking=True)          # uint8 H2D
        return K.u8_box_resize_normalize(staged, None, (self.res, self.res), self.mean, self.std, flip)
