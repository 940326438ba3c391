"""Data-parallel self-check under torchrun (NCCL): a few StyleGAN iterations on per-rank data, then every rank must hold the
same parameters (a gradient that missed its all-reduce, or a packed gradient that was not copied back, makes the replicas
diverge), and the averaged gradient of one backward must equal the mean of the per-rank gradients gathered explicitly."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import gan_lab_b200 as glb  # noqa: E402
from gan_lab_b200.config import default_config  # noqa: E402
from gan_lab_b200.parallel import DataParallel  # noqa: E402
from gan_lab_b200.stylegan.learner import StyleGANLearner  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl")
    glb.set_conv_impl("tf32")
    torch.manual_seed(1234 + rank)
    cfg = default_config("StyleGAN", res=32, batch_size=4, dev=f"cuda:{int(os.environ['LOCAL_RANK'])}", cutoff_trunc_trick=3)
    L = StyleGANLearner(cfg)
    L.dp = DataParallel(world)
    L.dp.broadcast_params(L.gen_model); L.dp.broadcast_params(L.disc_model)
    L.gen_model.train(); L.disc_model.train()
    L.beta = L.get_smoothing_ewma_beta(10.)
    L._init_lagged(); L._attach_ewma()
    if L.gen_model_lagged is not None:
        L.dp.broadcast_params(L.gen_model_lagged)
    x = torch.rand(4, 3, 32, 32, device=cfg.dev) * 2 - 1
    # (1) one D backward: hooked all-reduce vs an explicit gather of the local gradients
    L.dp.enabled = False
    for p in L.disc_model.parameters():
        p.requires_grad_(True)
    L.disc_model.zero_grad(set_to_none=True)
    out = L.disc_model(x).square().mean()
    out.backward()
    params = [p for p in L.disc_model.parameters() if p.grad is not None]
    local = [p.grad.clone() for p in params]
    L.dp.enabled = True
    L.dp.allreduce_grads(L.disc_model)
    worst = 0.0
    for p, g in zip(params, local):
        ref = g.clone()
        dist.all_reduce(ref, op=dist.ReduceOp.SUM)
        ref /= world
        worst = max(worst, float((p.grad - ref).abs().max() / ref.abs().max().clamp_min(1e-20)))
    # (2) replicas stay identical through (eager) training iterations; bench.py runs the same collectives inside captured graphs
    L.disc_model.zero_grad(set_to_none=True)
    for _ in range(4):
        L.main_iteration(x)
    torch.cuda.synchronize()
    div = 0.0
    for m in (L.gen_model, L.disc_model):
        for p in m.parameters():
            ref = p.detach().clone()
            dist.broadcast(ref, src=0)
            div = max(div, float((p.detach() - ref).abs().max()))
    t = torch.tensor([worst, div], device=cfg.dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ok = float(t[0]) < 1e-5 and float(t[1]) == 0.0
        print(f"dp_check world {world}: averaged-gradient rel err {float(t[0]):.2e}, replica divergence {float(t[1]):.2e} -> {'OK' if ok else 'FAIL'}")
    dist.destroy_process_group()
    sys.exit(0 if (float(t[0]) < 1e-5 and float(t[1]) == 0.0) else 1)


if __name__ == "__main__":
    main()
