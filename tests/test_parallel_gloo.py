"""Data-parallel host logic (gan_lab_b200/parallel.py) on CPU: world_size 2, gloo backend.

What is checked is the N>1 plumbing itself -- parameter broadcast, bucketed gradient averaging incl. channels-last
gradients, parameters without a gradient (prev_torgb / prev_fromrgb outside fade-in), the mean all-reduce used for w_ewma --
with plain torch modules standing in for the CUDA layers."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gan_lab_b200.parallel import DataParallel
        torch.manual_seed(100 + rank)                       # different initial weights per rank
        net = torch.nn.Sequential(torch.nn.Conv2d(4, 8, 3, padding=1), torch.nn.Conv2d(8, 8, 3, padding=1),
                                  torch.nn.Conv2d(8, 3, 1))
        net[0].weight.data = net[0].weight.data.contiguous(memory_format=torch.channels_last)
        unused = torch.nn.Parameter(torch.randn(5))        # never receives a gradient
        net.register_parameter("unused", unused)
        dp = DataParallel(world, bucket_bytes=1024)         # tiny buckets -> several all-reduces
        dp.broadcast_params(net)
        w_after_bcast = [p.detach().clone() for p in net.parameters()]
        torch.manual_seed(7 + rank)                         # different data per rank
        x = torch.randn(4, 4, 6, 6)
        net(x).pow(2).mean().backward()
        local = [None if p.grad is None else p.grad.detach().clone() for p in net.parameters()]
        dp.allreduce_grads(net)
        avg = [None if p.grad is None else p.grad.detach().clone() for p in net.parameters()]
        # second backward: the hooks attached by the first allreduce_grads() now launch the buckets DURING backward
        assert len(dp._hooked) == len(list(net.parameters()))
        net.zero_grad(set_to_none=True)
        net(x).pow(2).mean().backward()
        launched_in_backward = len(dp._inflight)
        dp.allreduce_grads(net)
        avg2 = [None if p.grad is None else p.grad.detach().clone() for p in net.parameters()]
        t = torch.full((3,), float(rank + 1))
        dp.allreduce_mean_(t)
        ret[rank] = dict(bcast=w_after_bcast, local=local, avg=avg, avg2=avg2, mean=t, launched=launched_in_backward)
    finally:
        dist.destroy_process_group()


def test_dataparallel_world2_gloo():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    r0, r1 = ret[0], ret[1]
    for a, b in zip(r0["bcast"], r1["bcast"]):             # rank 0's parameters everywhere
        assert torch.equal(a, b)
    for l0, l1, a0, a1 in zip(r0["local"], r1["local"], r0["avg"], r1["avg"]):
        if l0 is None:
            assert a0 is None and a1 is None and l1 is None
            continue
        assert not torch.equal(l0, l1)
        want = (l0 + l1) / 2
        torch.testing.assert_close(a0, want, rtol=1e-6, atol=1e-7)
        assert torch.equal(a0, a1)
        sig = lambda t: [st for st, sz in zip(t.stride(), t.shape) if sz > 1]
        assert sig(a0) == sig(l0)                           # layout of the gradient (channels_last) preserved
    assert r0["launched"] >= 2 and r0["launched"] == r1["launched"]     # buckets went out from the hooks, during backward
    for a, b in zip(r0["avg"], r0["avg2"]):                               # same data, same weights -> same averaged gradient
        assert (a is None and b is None) or torch.equal(a, b)
    assert torch.allclose(r0["mean"], torch.full((3,), 1.5)) and torch.equal(r0["mean"], r1["mean"])


class _Patch(object):
    """monkeypatch stand-in for a spawned worker process (nothing to undo: the process ends)."""

    @staticmethod
    def setattr(obj, name, value):
        setattr(obj, name, value)


def _learner_worker(rank, world, port, ret, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import numpy as np
        from pathlib import Path
        import gan_lab_b200._growth as growth
        from gan_lab_b200.config import default_config
        from gan_lab_b200.data import DeviceImageLoader
        from gan_lab_b200.parallel import DataParallel
        from gan_lab_b200.stylegan.learner import StyleGANLearner
        from oracle import kernel_contracts
        kernel_contracts.install(_Patch)                     # CPU doubles of the kernels: host logic only
        growth.FMAP_MAX = 32
        torch.manual_seed(100 + rank); np.random.seed(5)     # different weights / z / noise per rank, shared mixing stream
        cfg = default_config("StyleGAN", res=8, batch_size=4, dev="cpu", len_latent=32, len_dlatent=32, cutoff_trunc_trick=1)
        L = StyleGANLearner(cfg)
        L.dp = DataParallel(world, bucket_bytes=4096)
        L.dp.broadcast_params(L.gen_model); L.dp.broadcast_params(L.disc_model)
        images = torch.randint(0, 256, (16, 16, 16, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(9))
        dl = DeviceImageLoader(images, batch_size=4, res=8, shuffle=True, device="cpu", seed=3, rank=rank, world_size=world)
        seen = []
        orig = dl.batch
        dl.batch = lambda idx, flip=None: seen.extend(int(i) for i in idx) or orig(idx, flip)
        local_w = []
        orig_sync = L._sync_replicas
        L._sync_replicas = lambda: local_w.append(L.gen_model.w_ewma.clone()) or orig_sync()
        L.train(dl, num_main_iters=2)
        path = Path(tmp) / f"rank{rank}" / "model.tar"
        L.save_model(path)
        ret[rank] = dict(g=[p.detach().clone() for p in L.gen_model.parameters()],
                         d=[p.detach().clone() for p in L.disc_model.parameters()], seen=seen, saved=path.exists(),
                         w_local=local_w[0], w=L.gen_model.w_ewma.clone())
    finally:
        dist.destroy_process_group()


def test_learner_loader_checkpoint_world2_gloo(tmp_path):
    """The N>1 path of the rows around the step: per-rank shards of the real-image loader, gradient averaging inside the
    learner's steps (replicas stay identical although every rank draws its own latents and noise), rank 0 alone writes the
    checkpoint."""
    world, port = 2, _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_learner_worker, args=(world, port, ret, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = ret[0], ret[1]
    assert len(r0["seen"]) == len(r1["seen"]) == 8 and not set(r0["seen"]) & set(r1["seen"])
    for a, b in zip(r0["g"] + r0["d"], r1["g"] + r1["d"]):
        assert torch.equal(a, b)
    assert r0["saved"] and not r1["saved"]
    # w_ewma: rank-local averages differ (different latents), the synchronised one is their mean on both ranks
    assert not torch.equal(r0["w_local"], r1["w_local"])
    torch.testing.assert_close(r0["w"], (r0["w_local"] + r1["w_local"]) / 2, rtol=1e-6, atol=1e-7)
    assert torch.equal(r0["w"], r1["w"])


def _growth_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import numpy as np
        import gan_lab_b200._growth as growth
        from gan_lab_b200.config import default_config
        from gan_lab_b200.data import DeviceImageLoader
        from gan_lab_b200.parallel import DataParallel
        from gan_lab_b200.stylegan.learner import StyleGANLearner
        from oracle import kernel_contracts
        kernel_contracts.install(_Patch)
        growth.FMAP_MAX = 32
        torch.manual_seed(200 + rank); np.random.seed(5)     # every rank's own RNG: increase_scale() draws the new blocks from it
        cfg = default_config("StyleGAN", res=8, init_res=4, batch_size=4, dev="cpu", len_latent=32, len_dlatent=32,
                             cutoff_trunc_trick=1, nimg_transition=8)
        L = StyleGANLearner(cfg)
        L.dp = DataParallel(world, bucket_bytes=4096)
        L.dp.broadcast_params(L.gen_model); L.dp.broadcast_params(L.disc_model); L.dp.broadcast_params(L.gen_model_lagged)
        images = torch.randint(0, 256, (32, 16, 16, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(9))
        dl = DeviceImageLoader(images, batch_size=4, res=4, shuffle=True, device="cpu", seed=3, rank=rank, world_size=world)
        L.train(dl, num_main_iters=3)                       # 2 iterations at 4x4, growth, 1 iteration of the 8x8 fade-in
        ret[rank] = dict(res=int(L.gen_model.curr_res), fade=bool(L.gen_model.fade_in_phase),
                         g=[p.detach().clone() for p in L.gen_model.parameters()],
                         d=[p.detach().clone() for p in L.disc_model.parameters()],
                         lag=[p.detach().clone() for p in L.gen_model_lagged.parameters()])
    finally:
        dist.destroy_process_group()


def test_growth_keeps_replicas_identical_world2_gloo():
    """increase_scale() initialises the new blocks / torgb / fromrgb from each rank's own RNG: train() must re-broadcast rank 0's
    parameters (and the EWMA generator) after a growth, or the replicas differ from then on although gradients are averaged."""
    world, port = 2, _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_growth_worker, args=(world, port, ret), nprocs=world, join=True)
    r0, r1 = ret[0], ret[1]
    assert r0["res"] == r1["res"] == 8
    for key in ("g", "d", "lag"):
        assert len(r0[key]) == len(r1[key])
        for a, b in zip(r0[key], r1[key]):
            assert torch.equal(a, b), key
