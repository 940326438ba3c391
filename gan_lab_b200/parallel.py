"""Data parallelism over the GPUs of one node: one process per GPU, NCCL over NVLink 5 / NVSwitch.

The reference has no multi-GPU path at all (config.py:88-89 "TODO", README.md:136).  The batch shards
naturally: samples are independent in G (InstanceNorm is per-sample) and in D except for minibatch-stddev,
whose groups are 4 CONSECUTIVE samples -- with a per-GPU batch that is a multiple of 4 every group stays on one
rank, so averaging the per-rank gradients reproduces the single-GPU gradient of the N-times larger batch
(all losses are batch means).  The only exchange step is therefore the gradient all-reduce before each
optimiser step; parameters, Adam state and the EWMA generator are replicated.

Gradients are reduced in place, bucketed (default 32 MiB) so that NCCL launches overlap each other and the tail
of the backward pass; buckets are flat views over the gradients' dense storage (channels-last weights included).
"""
import torch
import torch.distributed as dist


def _flat_view(t):
    """1-D alias of a dense tensor's memory, whatever its logical strides (e.g. channels_last)."""
    return t.as_strided((t.numel(),), (1,))


class DataParallel(object):
    def __init__(self, world_size=None, bucket_bytes=32 * 1024 * 1024):
        self.world = world_size if world_size is not None else dist.get_world_size()
        self.bucket_bytes = bucket_bytes

    def broadcast_params(self, module, src=0):
        with torch.no_grad():
            for p in module.parameters():
                dist.broadcast(_flat_view(p.data), src=src)

    def allreduce_grads(self, module):
        """Average gradients across ranks (sum all-reduce of pre-scaled buckets)."""
        if self.world == 1:
            return
        grads = [p.grad for p in module.parameters() if p.grad is not None]
        bucket, size, works = [], 0, []

        def flush():
            nonlocal bucket, size
            if not bucket:
                return
            flat = torch.cat([_flat_view(g) for g in bucket])
            flat.mul_(1.0 / self.world)
            works.append((dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True), flat, bucket))
            bucket, size = [], 0

        for g in reversed(grads):            # reverse-autograd order: last layers' grads are ready first
            bucket.append(g)
            size += g.numel() * 4
            if size >= self.bucket_bytes:
                flush()
        flush()
        for work, flat, bk in works:
            work.wait()
            off = 0
            for g in bk:
                n = g.numel()
                _flat_view(g).copy_(flat[off:off + n])
                off += n

    def allreduce_mean_(self, t):
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            t.mul_(1.0 / self.world)
        return t
