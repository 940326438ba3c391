#!/usr/bin/env python
"""One main iteration (D step + G step) of a bench config between cudaProfilerStart/Stop, for
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file ... python tools/profile_step.py
Usage: profile_step.py [cfg2|cfg3|cfg4|cfg1] [fp32|tf32]"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import gan_lab_b200 as glb  # noqa: E402
from gan_lab_b200.config import default_config  # noqa: E402
from gan_lab_b200.progan.learner import ProGANLearner  # noqa: E402
from gan_lab_b200.stylegan.learner import StyleGANLearner  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    impl = sys.argv[2] if len(sys.argv) > 2 else "tf32"
    warm = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    glb.set_conv_impl(impl)
    model, res, init_res, bs_cfg, alpha = bench.CONFIGS[name]
    torch.manual_seed(0)
    cfg = default_config(model, res=res, init_res=init_res, batch_size=bs_cfg, dev="cuda:0")
    L = (StyleGANLearner if model == "StyleGAN" else ProGANLearner)(cfg)
    if alpha is not None:
        L.gen_model.increase_scale(); L.disc_model.increase_scale()
        L.gen_model.to(cfg.dev); L.disc_model.to(cfg.dev)
        L.gen_model.alpha = alpha
        L.batch_size = cfg.bs_dict[res]
        if cfg.use_ewma_gen:
            L.beta = L.get_smoothing_ewma_beta(10.)
            L._sync_lagged_structure()             # the EWMA generator grows with the live one (progan/learner.py:660-686)
        L._set_optimizer()
    bs = L.batch_size
    L.gen_model.train(); L.disc_model.train()
    L.beta = L.get_smoothing_ewma_beta(10.)
    L._init_lagged(); L._attach_ewma()
    if hasattr(L.gen_model, "_use_mixing_reg"):
        L.gen_model.device_mixing = True           # the path the CUDA-graph replayed bench runs (grouped style affines)
    x = torch.rand(bs, 3, res, res, device="cuda:0") * 2 - 1

    def main_iter():
        for p in L.disc_model.parameters():
            p.requires_grad_(True)
        L.disc_step(x)
        for p in L.disc_model.parameters():
            p.requires_grad_(False)
        L.gen_step()

    for _ in range(warm):
        main_iter()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    main_iter()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
