"""Data parallelism over the GPUs of one node: one process per GPU, NCCL over NVLink 5 / NVSwitch.

The reference has no multi-GPU path at all (config.py:88-89 "TODO", README.md:136).  The batch shards
naturally: samples are independent in G (InstanceNorm is per-sample) and in D except for minibatch-stddev,
whose groups are 4 CONSECUTIVE samples -- with a per-GPU batch that is a multiple of 4 every group stays on one
rank, so averaging the per-rank gradients reproduces the single-GPU gradient of the N-times larger batch
(all losses are batch means).  The only exchange step is therefore the gradient all-reduce before each
optimiser step; parameters, Adam state and the EWMA generator are replicated.  (Not so for the ResNet GAN:
its generator's BatchNorm sees per-rank batch statistics -- standard DDP semantics, not the single-GPU result at the larger
batch -- and its running buffers are averaged over the ranks where they are observed, GANLearner._sync_bn_buffers.)

Overlap with backward: every parameter carries a post-accumulate-grad hook; as soon as the gradients that have become
ready fill a bucket (in autograd order: last layers first) the bucket is all-reduced asynchronously -- NCCL runs on the process
group's own stream while the compute stream keeps executing the rest of backward (including the R1 double-backward tail).  With
NCCL the gradients of a bucket are reduced IN PLACE as one grouped call (ncclGroupStart / ncclAllReduce(ncclAvg) per tensor /
ncclGroupEnd): no pack, pre-scale or scatter-back pass for the large tensors; gradients below 1 MiB travel as ONE packed message
inside the same group (per-collective latency, not bytes, is what they cost); other backends (the gloo CPU tests) pack the bucket with `torch.cat`,
pre-scale by 1/world and copy back.  `allreduce_grads()` after backward only flushes the last partial bucket and makes the compute
stream wait for the collectives.  All of this is stream-ordered, so it is captured into the step's CUDA graph as parallel branches.
Default bucket size: 32 MiB on 2 GPUs, 128 MiB (= ONE group per network, issued when its backward has finished) on more: measured
on 8 B200s (cfg2, round 1) 4432 img/s against 4389 with 32 MiB buckets launched during backward -- the persistent tensor-core
kernels occupy every SM, so an NCCL kernel that starts mid-backward displaces convolution CTAs; on 2 GPUs the overlap wins (1151 vs
1123).  Round 2, in place: 1171 / 2302 / 4502 img/s on 2 / 4 / 8 GPUs.  Bucket composition follows the autograd order, which is
identical on every rank (same graph on every rank).
"""
import torch
import torch.distributed as dist


def _flat_view(t):
    """1-D alias of a dense tensor's memory, whatever its logical strides (e.g. channels_last)."""
    return t.as_strided((t.numel(),), (1,))


class DataParallel(object):
    def __init__(self, world_size=None, bucket_bytes=None, overlap=True, inplace=None):
        self.world = world_size if world_size is not None else dist.get_world_size()
        self.rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0   # rank 0 writes checkpoints
        if bucket_bytes is None:
            # measured on B200 (cfg2, img/s): 2 GPUs 1151 with 32 MiB buckets launched during backward vs 1123 with one
            # 128 MiB bucket per network; 4 GPUs 2184 vs 2275; 8 GPUs 4389 vs 4432
            bucket_bytes = (32 if self.world <= 2 else 128) * 1024 * 1024
        self.bucket_bytes = bucket_bytes
        self.overlap = overlap
        # inplace (NCCL only; default on): the gradients of a bucket are all-reduced WHERE THEY ARE, as one grouped NCCL call
        # (ncclGroupStart / one ncclAllReduce per tensor / ncclGroupEnd, averaging in the collective: ncclAvg) -- no pack
        # (`torch.cat`), no pre-scale and no scatter-back pass, i.e. ~6 bytes of HBM traffic per gradient byte less per step.
        if inplace is None:
            inplace = dist.is_available() and dist.is_initialized() and dist.get_backend() == "nccl"
        self.inplace = bool(inplace)
        import os
        self.pack_below = int(os.environ.get("GLB_DP_PACK_BELOW", 262144))     # elements; 0 = every gradient in place
        self.enabled = True         # False: the hooks stay silent (rank-local passes such as bench.py's roofline replay)
        self._hooked = set()        # ids of parameters that carry our hook
        self._pending = []          # parameters whose gradient is ready but not yet in a bucket
        self._pending_bytes = 0
        self._inflight = []         # (work, flat bucket, parameters)

    def broadcast_params(self, module, src=0):
        with torch.no_grad():
            for p in module.parameters():
                dist.broadcast(_flat_view(p.data), src=src)
            for b in module.buffers():
                if b.is_floating_point():
                    dist.broadcast(_flat_view(b.data), src=src)

    # ------------------------------------------------------------------ overlap machinery
    def attach(self, *modules):
        """Hook every parameter of `modules` (idempotent; call again after the networks have grown)."""
        if self.world == 1 or not self.overlap:
            return
        for m in modules:
            for p in m.parameters():
                if id(p) not in self._hooked and p.requires_grad:     # (frozen for the moment: hooked at a later call)
                    p.register_post_accumulate_grad_hook(self._on_grad)
                    self._hooked.add(id(p))

    def _on_grad(self, p):
        if p.grad is None or not self.enabled:
            return
        self._pending.append(p)
        self._pending_bytes += p.numel() * 4
        if self._pending_bytes >= self.bucket_bytes:
            self._launch()

    @torch.no_grad()
    def _launch(self):
        ps, self._pending, self._pending_bytes = self._pending, [], 0
        if not ps:
            return
        if self.inplace:
            # large gradients are reduced where they are; the many SMALL ones (biases, noise weights, narrow layers: ~half of the
            # tensors, ~1 % of the bytes) are packed into one message first -- every collective of a group costs its own
            # latency, which grows with the rank count (measured on 8 B200s: +1.7 ms per cfg2 iteration with ~160 ops per step)
            big = [p for p in ps if p.numel() >= self.pack_below]
            small = [p for p in ps if p.numel() < self.pack_below]
            flat = torch.cat([_flat_view(p.grad) for p in small]) if len(small) > 1 else None
            if flat is None:
                big, small = ps, []
            with dist._coalescing_manager(device=ps[0].grad.device, async_ops=True) as cm:
                if flat is not None:
                    dist.all_reduce(flat, op=dist.ReduceOp.AVG)
                for p in big:
                    dist.all_reduce(_flat_view(p.grad), op=dist.ReduceOp.AVG)
            self._inflight.append((cm, flat, small))
            return
        flat = torch.cat([_flat_view(p.grad) for p in ps])
        flat.mul_(1.0 / self.world)
        self._inflight.append((dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True), flat, ps))

    def allreduce_grads(self, module):
        """Average the gradients of `module` across ranks; returns once the compute stream is ordered after the
        collectives (sum all-reduce of pre-scaled buckets)."""
        if self.world == 1:
            return
        done = {id(p) for _, _, ps in self._inflight for p in ps} | {id(p) for p in self._pending}
        for p in reversed(list(module.parameters())):       # reverse-autograd order: last layers' grads are ready first
            if p.grad is not None and id(p) not in done:
                self._on_grad(p)
        self._launch()
        with torch.no_grad():
            for work, flat, ps in self._inflight:
                work.wait()
                if flat is not None:
                    torch._foreach_copy_([_flat_view(p.grad) for p in ps], list(flat.split([p.numel() for p in ps])))
        self._inflight = []
        self.attach(module)                                  # from the next backward on, buckets launch from the hooks

    def allreduce_mean_(self, t):
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            t.mul_(1.0 / self.world)
        return t
