#!/usr/bin/env python
"""bench.py -- StyleGAN G+D training throughput (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W              # our arm (sm_100a kernels through the C-ABI)
    python bench.py --impl reference --gpus N --steps K ...    # reference arm: the CPU oracle port on host cores

One "step" = one main iteration of the reference's train loop = 1 D step + 1 G step on a synthetic batch
(cfg2: StyleGAN 128x128, nonsaturating + R1 + drift, noise, InstanceNorm/AdaIN, mixing 0.9, batch 8 per GPU).
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for how each field is obtained.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

CONFIGS = {
    # name: (model, res, init_res, batch per GPU, fade alpha or None)
    "cfg2": ("StyleGAN", 128, 128, 8, None),
    "cfg3": ("StyleGAN", 256, 128, 16, 0.5),
    "cfg4": ("StyleGAN", 1024, 1024, 16, None),     # bs_dict[1024] = 16 // 4 = 4 per GPU
    "cfg1": ("ProGAN", 32, 32, 8, None),
    "cfg5": ("ResNet GAN", 64, 64, 64, None),
}
WORKLOAD_NAME = {
    "cfg2": "StyleGAN 128x128 nonsaturating loss + R1, batch 8 per GPU, fp32 storage (BASELINE.json configs[1])",
    "cfg3": "StyleGAN 256x256 mid-fade-in (alpha=0.5), batch 16 per GPU",
    "cfg4": "StyleGAN 1024x1024, batch 4 per GPU",
    "cfg1": "ProGAN 32x32 fixed resolution, batch 8 (BASELINE.json configs[0])",
    "cfg5": "ResNet GAN 64x64 (WGAN + WGAN-GP, 5 D steps + 1 G step per iteration), batch 64 per GPU",
}


# ----------------------------------------------------------------------------------------------- helpers
def workload_config(name, per_gpu_batch, world):
    """The `config` object of BOTH arms (ours and --impl reference): what is computed, nothing about how."""
    return {"workload": WORKLOAD_NAME[name], "global_batch": per_gpu_batch * world, "parallelism": f"dp{world}"}


def conv_flops_per_iter(model, res, bs):
    """Algorithmic conv/linear FLOPs of one main iteration = 4 F_G + 14 F_D."""
    import math
    fm = lambda r: min(int(8192 / (2 ** (int(math.log2(r)) - 1))), 512)
    FG = FD = 0.0
    r = 4
    # generator
    if model == "StyleGAN":
        FG += 2 * bs * 16 * 512 * 512 * 9                       # layer 1 conv at 4x4 (layer 0 has no conv)
        FG += 2 * bs * 512 * 512 * 8 * 2                        # mapping network (x2 when mixing fires; count once)
    else:
        FG += 2 * bs * 512 * 8192 + 2 * bs * 16 * 512 * 512 * 9
    while r < res:
        r *= 2
        ci, co = fm(r // 2), fm(r)
        FG += 2 * bs * r * r * co * ci * 9 + 2 * bs * r * r * co * co * 9
        if model == "StyleGAN":
            FG += 2 * 2 * bs * 512 * 2 * co
    FG += 2 * bs * res * res * 3 * fm(res)
    # discriminator
    FD += 2 * bs * res * res * 3 * fm(res)
    r = res
    while r > 4:
        c, cn = fm(r), fm(r // 2)
        FD += 2 * bs * r * r * c * c * 9 + 2 * bs * r * r * cn * c * 9
        r //= 2
    FD += 2 * bs * 16 * 512 * 513 * 9 + 2 * bs * 512 * 512 * 16 + 2 * bs * 512
    return 4 * FG + 14 * FD


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, threading.Event(), []

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": d["hbm_gbs"], "tf": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "src": "measured"}
    return {"hbm_gbs": 6650.0, "tf": 1400.0, "src": "fallback"}


def measured_mma_peaks():
    """tcgen05 issue peaks measured on this pool's B200 by tools/micro/mma_peak.cu (every SM issuing back-to-back M128 N256 MMAs
    from shared memory; CUDA events): the committed run in profiles/."""
    p = ROOT / "profiles" / "r2_mma_peak_tcgen05.txt"
    out = {"tf32": 1115.0, "bf16": 2234.0}
    try:
        d = json.loads([ln for ln in p.read_text().splitlines() if ln.startswith("{")][-1])
        out = {"tf32": d["tf32_tflops_sustained"], "bf16": d["bf16_mma_tflops_sustained"]}
    except Exception:
        pass
    return out


DOMINANT_KERNEL = "conv_fprop_tc2_kernel<256"     # CTA-pair implicit GEMM: the largest share of the step when the committed ncu capture was taken
                                                  # (since then conv_fprop_tc2_fold_kernel took its folded layers; the wgrad kernel leads the launch list)


def ncu_traffic():
    """`roofline.traffic`: dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant convolution kernel, read from
    the committed `ncu --set full` capture (profiles/r2_ncu_full_conv_family.csv, raw page, one row per launch; the rows of
    DOMINANT_KERNEL)."""
    import csv
    p = ROOT / "profiles" / "r2_ncu_full_conv_family.csv"
    if not p.exists():
        return {"traffic": None, "traffic_note": "no committed ncu capture"}
    try:
        rows = list(csv.reader(p.open()))
        hdr = rows[0]
        rd, wr, nm = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
        allrows = [r for r in rows[1:] if len(r) == len(hdr) and r[hdr.index("ID")].strip().isdigit()]
        kern = DOMINANT_KERNEL
        body = [r for r in allrows if kern in r[nm]]
        for alt in ("conv_fprop_tc2_fold_kernel<256", "conv_fprop_tc2_halo_kernel<256", "conv_fprop_tc"):
            if body:
                break
            kern, body = alt, [r for r in allrows if alt in r[nm]]
        unit = rows[1] if rows[1][hdr.index("ID")].strip() == "" else None          # units row of the raw page
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        def val(r, i):
            u = unit[i] if unit else "byte"
            return float(r[i].replace(",", "")) * scale.get(u, 1.0)
        tr = [val(r, rd) + val(r, wr) for r in body]
        return {"traffic": sum(tr) / len(tr),
                "traffic_note": f"mean over {len(tr)} launches of {kern}...> in profiles/r2_ncu_full_conv_family.csv "
                                "(dram__bytes_read.sum + dram__bytes_write.sum)"}
    except Exception as e:          # noqa: BLE001
        return {"traffic": None, "traffic_note": "ncu csv unreadable: " + repr(e)[:120]}


# ----------------------------------------------------------------------------------------------- reference arm
def run_reference(args):
    """The reference's own CPU implementation of the path on the box's host cores, all host threads.

    kind = "reference": the UNMODIFIED reference (pip-installed into baseline/_ref) -- its own
    StyleGANLearner / ProGANLearner built from its own config Namespace, `Learner.train(dl, num_main_iters=...)` on a
    synthetic TensorDataset, dev = cpu.  Only stubs for the absent, off-path matplotlib / indexed / lmdb modules are
    installed (oracle/reference_loader.py).  If baseline/_ref is absent: kind = "port", the CPU oracle port."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    model, res, init_res, bs, alpha = CONFIGS[args.config]
    if args.config == "cfg4":
        bs = bs // 4
    if getattr(args, "ref_dev", "cpu") == "cuda":
        return run_reference_on_gpu(args, model, res, bs, alpha)
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sample_bs = bs               # the full per-GPU batch of the workload: same configuration as our arm (~5 s per step on 16 cores)
    from oracle.reference_loader import reference_available
    kind = "reference" if (reference_available() and alpha is None) else "port"
    if kind == "reference":
        one, note = _reference_stepper(model, res, sample_bs, args.steps, args.warmup)
    else:
        one, note = _port_stepper(model, res, sample_bs)
    for _ in range(args.warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one()
    dt = time.perf_counter() - t0
    val = sample_bs * args.steps / dt
    sample = (f"{args.steps} main iterations (1 D step + 1 G step) at the workload's full batch {sample_bs}, after {args.warmup} warm-up, "
              f"{threads} torch threads")
    line = {"impl": "reference", "metric": "StyleGAN G+D train img/s", "value": val, "unit": "img/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.config, bs, 1), "details": {"note": note},
            "cpu_baseline": {"value": val, "unit": "img/s", "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_reference_on_gpu(args, model, res, bs, alpha):
    """`--impl reference --ref-dev cuda` (not the contract's reference arm, which is the CPU path): the UNMODIFIED reference
    with dev=cuda, i.e. stock PyTorch eager (cuDNN / cuBLAS / ATen kernels) on the same B200 -- the number SURVEY.md 8d calls
    'the real bar to beat'.  Full per-GPU batch, one train() call for the warm-up and one for the timed steps (the reference
    parks its networks on the CPU at the end of every train() call), wall clock around a device synchronise."""
    import contextlib
    import io
    import torch
    from torch.utils.data import BatchSampler, DataLoader, SequentialSampler, TensorDataset
    from oracle.reference_loader import load_reference, make_config, reference_available, REFERENCE_ROOT
    if not reference_available() or alpha is not None or model == "ResNet GAN" or not torch.cuda.is_available():
        if not getattr(args, "_embed", False):
            print(json.dumps({"impl": "reference", "unavailable": "reference-on-GPU needs baseline/_ref, a CUDA device and a "
                              "fixed-resolution StyleGAN / ProGAN config"}), flush=True)
        return None
    ref = load_reference()
    torch.manual_seed(0)
    cfg = make_config(model, res=res, init_res=res, batch_size=bs, dev="cuda", metrics_dev=torch.device("cuda"))
    quiet = lambda: contextlib.redirect_stdout(io.StringIO())
    with quiet():
        L = (ref.stylegan_learner.StyleGANLearner if model == "StyleGAN" else ref.progan_learner.ProGANLearner)(cfg)
    gen = torch.Generator().manual_seed(0)
    data = torch.rand(8 * bs, 3, res, res, generator=gen) * 2 - 1
    ds = TensorDataset(data)
    dl = DataLoader(ds, batch_sampler=BatchSampler(SequentialSampler(ds), batch_size=bs, drop_last=True), pin_memory=True)
    with quiet(), contextlib.redirect_stderr(io.StringIO()):
        L.train(dl, num_main_iters=max(args.warmup, 3))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        L.train(dl, num_main_iters=args.steps)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    val = bs * args.steps / dt
    line = {"impl": "reference", "metric": "StyleGAN G+D train img/s", "value": val, "unit": "img/s", "n_gpus": 1,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 (stock PyTorch defaults: cuDNN TF32 allowed, matmul fp32)",
            "data": "synthetic",
            "config": workload_config(args.config, bs, 1),
            "details": {"note": f"UNMODIFIED reference (gan_lab {REFERENCE_ROOT}) Learner.train with dev=cuda: stock PyTorch eager on "
                                f"the B200, torch {torch.__version__}"},
            "e2e": {"value": val, "unit": "img/s", "h2d_bytes_per_step": bs * 3 * res * res * 4, "d2h_bytes_per_step": 8}}
    if not getattr(args, "_embed", False):
        print(json.dumps(line), flush=True)
    return line


def _reference_stepper(model, res, bs, steps, warmup):
    """-> (callable running ONE main iteration of the unmodified reference's Learner.train on CPU, note)."""
    import contextlib
    import io
    import torch
    from torch.utils.data import BatchSampler, DataLoader, SequentialSampler, TensorDataset
    from oracle.reference_loader import load_reference, make_config, REFERENCE_ROOT
    ref = load_reference()
    torch.manual_seed(0)
    cfg = make_config(model, res=res, init_res=res, batch_size=bs, dev="cpu")
    quiet = lambda: contextlib.redirect_stdout(io.StringIO())
    with quiet():
        L = (ref.stylegan_learner.StyleGANLearner if model == "StyleGAN" else ref.progan_learner.ProGANLearner)(cfg)
    gen = torch.Generator().manual_seed(0)
    data = torch.rand((steps + warmup + 1) * bs, 3, res, res, generator=gen) * 2 - 1
    ds = TensorDataset(data)
    dl = DataLoader(ds, batch_sampler=BatchSampler(SequentialSampler(ds), batch_size=bs, drop_last=True))

    def one():
        with quiet(), contextlib.redirect_stderr(io.StringIO()):
            L.train(dl, num_main_iters=1)

    return one, f"UNMODIFIED reference (gan_lab {REFERENCE_ROOT}) Learner.train on host cores, dev=cpu"


def _port_stepper(model, res, bs):
    import torch
    from oracle.train_oracle import OracleTrainer, IterDraws, sample_gen_draws
    from oracle import gan_oracle as O
    gen = torch.Generator().manual_seed(0)
    g_sd, d_sd = synth_params(model, res, gen)
    T = OracleTrainer(g_sd, d_sd, model=model, res=res, lr=0.0015, loss="nonsaturating" if model == "StyleGAN" else "wgan",
                      gp_type="r1" if model == "StyleGAN" else "wgan-gp", ewma_beta=O.ewma_beta(bs),
                      g_kwargs={} if model == "StyleGAN" else dict(use_pixelnorm=True))

    def one():
        real = torch.rand(bs, 3, res, res, generator=gen) * 2 - 1
        eps = torch.rand(bs, 1, 1, 1, generator=gen) if model != "StyleGAN" else None
        T.main_iter(real, IterDraws(sample_gen_draws(model, res, bs, 512, gen), sample_gen_draws(model, res, bs, 512, gen), eps))

    return one, "CPU oracle port of the reference algorithm on host cores"


def synth_params(model, res, gen):
    """Random-init parameter dicts with the reference's names/shapes/initialiser (N(0,1) weights under equalized LR,
    zero biases, ones const input), built by instantiating our drop-in modules on CPU (no kernels run)."""
    import torch
    import gan_lab_b200._growth as growth
    from gan_lab_b200.config import default_config
    from gan_lab_b200.progan.learner import ProGANLearner
    from gan_lab_b200.stylegan.learner import StyleGANLearner
    torch.manual_seed(0)
    cfg = default_config(model, res=res, batch_size=8, dev="cpu", use_ewma_gen=False)
    L = (StyleGANLearner if model == "StyleGAN" else ProGANLearner)(cfg)
    return ({k: v.detach().clone().contiguous() for k, v in L.gen_model.state_dict().items()},
            {k: v.detach().clone().contiguous() for k, v in L.disc_model.state_dict().items()})


# ----------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import gan_lab_b200 as glb
    from gan_lab_b200 import _kernels as K
    from gan_lab_b200.config import default_config
    from gan_lab_b200.progan.learner import ProGANLearner
    from gan_lab_b200.stylegan.learner import StyleGANLearner

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (our arm) needs a CUDA device: gan_lab_b200 has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    glb.set_conv_impl(args.conv_impl)

    model, res, init_res, bs_cfg, alpha = CONFIGS[args.config]
    torch.manual_seed(0 + rank)
    import numpy as np
    np.random.seed(0)                                  # mixing cut-off stream shared by all ranks (SURVEY 8e caveat)
    cfg = default_config(model, res=res, init_res=init_res, batch_size=bs_cfg, dev=str(dev))
    if model == "ResNet GAN":
        return run_resnet(args, cfg, rank, world, dev)
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
    if world > 1:
        from gan_lab_b200.parallel import DataParallel
        mb = os.environ.get("GLB_DP_BUCKET_MB")          # tuning experiments only; default = DataParallel's measured choice
        L.dp = DataParallel(world, overlap=os.environ.get("GLB_DP_OVERLAP", "1") != "0",
                            bucket_bytes=int(float(mb) * 1024 * 1024) if mb else None,
                            inplace=(os.environ["GLB_DP_INPLACE"] != "0") if "GLB_DP_INPLACE" in os.environ else None)
        L.dp.broadcast_params(L.gen_model); L.dp.broadcast_params(L.disc_model)
        if os.environ.get("GLB_DP_OVERLAP_D"):           # A/B: D's all-reduce + Adam beside the generator forward of the G step
            L.overlap_d_update = os.environ["GLB_DP_OVERLAP_D"] != "0"
    def log(msg):
        if args.verbose:
            print(f"[bench rank {rank}] {msg}", file=sys.stderr, flush=True)

    log("models built, parameters broadcast")
    L.gen_model.train(); L.disc_model.train()
    L.beta = L.get_smoothing_ewma_beta(10.)
    L._init_lagged(); L._attach_ewma()

    # synthetic data: device-resident pool for `value`, pinned host pool for `e2e`
    n_pool = 8
    pool_dev = [(torch.rand(bs, 3, res, res, device=dev) * 2 - 1) for _ in range(n_pool)]
    pool_host = [(torch.rand(bs, 3, res, res) * 2 - 1).pin_memory() for _ in range(n_pool)]
    flush = torch.empty(160 * 1024 * 1024 // 4, device=dev)   # > 126 MB L2

    use_graphs = not args.no_graphs
    L.enable_cuda_graphs(use_graphs, warmup_iters=3)
    if os.environ.get("GLB_DISC_FUSE"):
        from gan_lab_b200.progan.architectures import DiscBlock
        DiscBlock.fuse_blur = os.environ["GLB_DISC_FUSE"] != "0"
    if os.environ.get("GLB_SIDE_STYLES"):
        L.gen_model.side_stream_styles = os.environ["GLB_SIDE_STYLES"] != "0"
    if os.environ.get("GLB_BATCH_D"):          # A/B switch: D(fake) and D(real) as one pass over the concatenated batch
        L.batch_d_passes = os.environ["GLB_BATCH_D"] != "0"
    if os.environ.get("GLB_PARALLEL_D"):
        L.parallel_d_passes = os.environ["GLB_PARALLEL_D"] != "0"

    def main_iter_eager(x):
        for p in L.disc_model.parameters():
            p.requires_grad_(True)
        ld = L.disc_step(x)
        for p in L.disc_model.parameters():
            p.requires_grad_(False)
        lg = L.gen_step()
        return ld, lg

    def main_iter(x):
        return L.main_iteration(x)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up: >= 3 eager iterations (the last one also counts this repo's kernel launches per iteration), then -- with
    # CUDA graphs -- the capture itself and replays, all before the timed region
    n_warm = max(args.warmup, 3)
    launches_per_iter = 0
    for i in range(n_warm):
        l0 = K.launch_count()
        main_iter(pool_dev[i % n_pool])
        launches_per_iter = max(launches_per_iter, K.launch_count() - l0)
        log(f"eager warm-up iteration {i} done")
    if use_graphs:
        for i in range(3):
            main_iter(pool_dev[i % n_pool])
        assert L._graph is not None, "CUDA graphs were requested but not captured"
        log("graphs captured and replayed")
    barrier()

    # ---- value: inputs resident in HBM -------------------------------------------------------------
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    step_ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    ev[0].record()
    step_ev[0].record()
    for i in range(args.steps):
        main_iter(pool_dev[i % n_pool])
        step_ev[i + 1].record()
    ev[1].record()
    barrier()
    ms = ev[0].elapsed_time(ev[1])
    per_step = sorted(step_ev[i].elapsed_time(step_ev[i + 1]) for i in range(args.steps))
    ms_median = per_step[len(per_step) // 2]
    log(f"timed region done: {ms / args.steps:.2f} ms/step")
    launches = launches_per_iter * args.steps   # kernels of this repo per main iteration (counted on an eager iteration) x steps
    if sampler:
        sampler.stop_flag.set(); sampler.join(timeout=2)
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t)

    # ---- e2e: public API (Learner.train) with HOST batches; H2D of the batch and D2H of the losses per step ----
    class HostLoader:
        dataset = list(range(n_pool * bs))
        batch_sampler = None

        def __iter__(self):
            return iter([(x,) for x in pool_host])

    loader = HostLoader()
    L.not_trained_yet = False                          # keep optimiser/EWMA state from the warm-up
    L.train_dataiter = iter(loader)
    L.nimg_transition = getattr(L, "nimg_transition", cfg.nimg_transition)
    L.nimg_transition_lst = getattr(L, "nimg_transition_lst", [10 ** 12])
    L.curr_img_num = 0
    sched, L.sched_bool = L.sched_bool, False
    barrier()
    ev2 = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev2[0].record()
    if L.gen_model.fade_in_phase:
        L.delta_alpha = 0.0                            # hold alpha fixed (cfg3: mid-fade-in)
    host_losses = []
    # the step's result (both losses) is read back to the host every step (the reference's tqdm line reads them too): an
    # asynchronous 8-byte device-to-host copy into pinned memory per step, stream-ordered before the next replay overwrites the
    # static loss tensors, consumed after the loop -- no host stall per step.  Falls back to a blocking read if that fails.
    readback = {"buf": None, "n": 0}
    try:
        readback["buf"] = torch.empty(args.steps, 2, dtype=torch.float32).pin_memory()
    except Exception:
        readback["buf"] = None

    def on_step(i, ld, lg):
        buf = readback["buf"]
        if buf is not None and i < buf.shape[0]:
            try:
                buf[i, 0].copy_(ld.detach().reshape(()), non_blocking=True)
                buf[i, 1].copy_(lg.detach().reshape(()), non_blocking=True)
                readback["n"] = i + 1
                return
            except Exception:
                readback["buf"] = None
        host_losses.append((ld.item(), lg.item()))

    L.train(loader, num_main_iters=args.steps, step_callback=on_step)
    if readback["n"]:
        torch.cuda.current_stream().synchronize()       # the last read has landed; inside the timed region
        host_losses.extend(tuple(r) for r in readback["buf"][:readback["n"]].tolist())
    ev2[1].record()
    # train() ends like the reference's (progan/learner.py:1016-1030): optimisers rebuilt, networks in eval mode
    L.gen_model.train(); L.disc_model.train()
    barrier()
    ms_e2e = ev2[0].elapsed_time(ev2[1])
    t = torch.tensor([ms_e2e], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_e2e = float(t)
    L.sched_bool = sched

    log("e2e done")
    # ---- roofline of the dominant kernel family (dense convs), timed live with CUDA events on the launch stream ----
    roof, glue = kernel_rooflines(L, pool_dev[0], main_iter_eager, flush) if rank == 0 else (None, None)

    if rank == 0:
        peaks = measured_peaks()
        imgs = bs * cfg.num_disc_iters * args.steps * world
        value = imgs / (ms / 1e3)
        flops = conv_flops_per_iter(model, res, bs)
        line = {
            "metric": "StyleGAN G+D train img/s", "value": value, "unit": "img/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": {"tf32": "tf32", "bf16": "bf16", "fp32": "f32"}[args.conv_impl], "data": "synthetic",
            "config": workload_config(args.config, bs, world),
            "ms_per_step_median": ms_median, "value_from_median_step": bs * cfg.num_disc_iters * world / (ms_median / 1e3),
            "details": {"conv_impl": args.conv_impl, "cuda_graphs": use_graphs,
                       "r1_shares_real_forward": bool(getattr(L, "share_penalty_forward", False)),
                       "d_fake_real_one_pass": bool(getattr(L, "batch_d_passes", False)),
                       "grad_allreduce": (f"NCCL all-reduce ({('in place, grouped, ncclAvg; gradients below ' + str(L.dp.pack_below * 4 >> 10) + ' KiB packed into one message of the group') if L.dp.inplace else 'packed buckets'}), "
                                          f"{L.dp.bucket_bytes >> 20} MiB buckets launched from grad hooks as they fill"
                                          + ("; D's all-reduce + Adam overlapped with the generator forward of the G step" if getattr(L, "overlap_d_update", False) else "")
                                          if world > 1 else None), "l2": "8-batch input pool; activations per step (>1 GB) exceed the 126 MB L2",
                       "algorithmic_conv_gflop_per_step_per_gpu": flops / 1e9,
                       "achieved_conv_tflops_whole_step": flops / (ms / args.steps / 1e3) / 1e12},
            "e2e": {"value": imgs / (ms_e2e / 1e3), "unit": "img/s", "h2d_bytes_per_step": bs * 3 * res * res * 4,
                    "d2h_bytes_per_step": 8},
            "gpu_launches": launches,
            "clocks": sampler.summary() if sampler else None,
        }
        if roof is not None:
            roof["peak"] = peaks["tf"]
            roof["frac"] = roof["achieved"] / peaks["tf"]
            mma = measured_mma_peaks()
            roof["peak_source"] = ("of " + peaks["src"] + " dense bf16 cuBLAS peak (sustained figure: kernels timed inside a long "
                                   "step); operands are TF32 -> frac_of_tf32_mma_peak is against the MEASURED tcgen05 kind::tf32 issue "
                                   "peak of this chip (tools/micro/mma_peak.cu, profiles/r2_mma_peak_tcgen05.txt)")
            roof["tf32_mma_peak"] = mma["tf32"]
            roof["frac_of_tf32_mma_peak"] = roof["achieved"] / mma["tf32"]
            if args.config == "cfg2":
                roof.update(ncu_traffic())
            line["roofline"] = roof
        if glue is not None and glue["achieved"] is not None:
            glue["peak"] = peaks["hbm_gbs"]
            glue["frac"] = glue["achieved"] / peaks["hbm_gbs"]
            glue["peak_source"] = peaks["src"] + " HBM copy bandwidth (read+write)"
            line["roofline_glue"] = glue
        if world == 1:
            line["roofline_input"] = input_pipeline_roofline(peaks)
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args)
            # the opt-in bf16-operand mode on the same workload, as its own process (same harness)
            if args.conv_impl == "tf32" and os.environ.get("GLB_BENCH_NO_BF16") is None:
                try:
                    sub = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--config", args.config, "--conv-impl", "bf16",
                                          "--steps", str(args.steps), "--warmup", str(args.warmup), "--no-cpu-baseline"],
                                         capture_output=True, text=True, timeout=600)
                    b = json.loads([ln for ln in sub.stdout.splitlines() if ln.startswith("{")][-1])
                    line["opt_in_bf16_operands"] = {
                        "value": b["value"], "unit": "img/s", "ms_per_step": b["ms_per_step"], "e2e": b["e2e"]["value"],
                        "conv_tflops": b["roofline"]["achieved"], "conv_frac_of_bf16_peak": b["roofline"]["frac"],
                        "glue_gbs": b["roofline_glue"]["achieved"],
                        "note": "python bench.py --conv-impl bf16: bf16 copies of the GEMM operands (round to nearest), fp32 accumulation / "
                                "storage; NOT the default -- BASELINE.json names fp32/TF32 for this workload and the whole-step deviation "
                                "is 4-10x the TF32 one (tests/test_cfg2_fullwidth.py)"}
                except Exception as e:          # noqa: BLE001
                    line["opt_in_bf16_operands"] = {"error": repr(e)[:300]}
            # SURVEY.md 8d "the real bar to beat": the UNMODIFIED reference through stock PyTorch (cuDNN / cuBLAS / ATen) on this
            # same B200, same workload, 3 iterations after 3 warm-up (~1.5 s each)
            try:
                L._graph = None
                del L
                import gc
                gc.collect(); torch.cuda.empty_cache()
                sub = argparse.Namespace(**vars(args)); sub.steps, sub.warmup, sub._embed, sub.ref_dev = 3, 3, True, "cuda"
                ref_line = run_reference_on_gpu(sub, model, res, bs, alpha)
                if ref_line is not None:
                    line["torch_eager_b200"] = {"value": ref_line["value"], "unit": "img/s", "ms_per_step": ref_line["ms_per_step"],
                                                "steps": 3, "note": ref_line["details"]["note"],
                                                "speedup_of_this_repo": line["e2e"]["value"] / ref_line["value"]}
            except Exception as e:          # noqa: BLE001 -- an auxiliary figure must not take the bench line down
                line["torch_eager_b200"] = {"error": repr(e)[:300]}
        print(json.dumps(line), flush=True)
    if world > 1:
        # Captured NCCL work keeps the communicator busy: drop the graphs first, then tear down; never let a stuck
        # communicator teardown hang the benchmark after the result line is out.
        L._graph = None
        import gc
        gc.collect()
        torch.cuda.synchronize()
        sys.stdout.flush(); sys.stderr.flush()
        t = threading.Thread(target=dist.destroy_process_group, daemon=True)
        t.start()
        t.join(timeout=20)
        os._exit(0)


def run_resnet(args, cfg, rank, world, dev):
    """cfg5: the ResNet GAN loop (reference resnetgan/learner.py:463-776): per main iteration 1 generator step, then 5
    discriminator steps, each step replayed as a CUDA graph.  Same JSON contract; the algorithmic
    conv FLOPs are the ones the recorded conv launches execute (nothing is shared or skipped on this path)."""
    import torch
    import torch.distributed as dist
    from gan_lab_b200 import _kernels as K
    from gan_lab_b200.resnetgan.learner import GANLearner
    L = GANLearner(cfg)
    bs = L.batch_size
    if world > 1:
        from gan_lab_b200.parallel import DataParallel
        L.dp = DataParallel(world)
        L.dp.broadcast_params(L.gen_model); L.dp.broadcast_params(L.disc_model)
    L.gen_model.train(); L.disc_model.train()
    nd = cfg.num_disc_iters
    res = cfg.res_samples
    pool_dev = [(torch.rand(bs, 3, res, res, device=dev) * 2 - 1) for _ in range(nd)]
    pool_host = [(torch.rand(bs, 3, res, res) * 2 - 1).pin_memory() for _ in range(nd)]
    flush = torch.empty(160 * 1024 * 1024 // 4, device=dev)

    resnet_graphs = os.environ.get("GLB_RESNET_GRAPHS", "1") != "0" and not args.no_graphs   # measured on B200: 3081 vs 2603 img/s
    if resnet_graphs:
        L.enable_cuda_graphs(True, warmup_iters=3)

    def main_iter(pool):
        if resnet_graphs:
            return L.main_iteration(pool)
        for p in L.disc_model.parameters():
            p.requires_grad_(False)
        lg = L.gen_step()
        for p in L.disc_model.parameters():
            p.requires_grad_(True)
        for x in pool:
            ld = L.disc_step(x)
        return ld, lg

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    launches_per_iter = 0
    for _ in range(max(args.warmup, 3)):
        l0 = K.launch_count()
        main_iter(pool_dev)
        launches_per_iter = K.launch_count() - l0
    barrier()
    sampler = ClockSampler(dev.index or 0) if rank == 0 else None
    if sampler:
        sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    for _ in range(args.steps):
        main_iter(pool_dev)
    ev[1].record()
    barrier()
    if sampler:
        sampler.stop_flag.set(); sampler.join(timeout=2)
    ev[2].record()
    for _ in range(args.steps):                         # e2e: host batches, H2D per D step, losses read back per iteration
        ld, lg = main_iter(pool_host)
        _ = (ld.item(), lg.item())
    ev[3].record()
    barrier()
    t = torch.tensor([ev[0].elapsed_time(ev[1]), ev[2].elapsed_time(ev[3])], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    def eager_iter(pool):
        for p in L.disc_model.parameters():
            p.requires_grad_(False)
        lg = L.gen_step()
        for p in L.disc_model.parameters():
            p.requires_grad_(True)
        for x in pool:
            ld = L.disc_step(x)
        return ld, lg

    roof, glue = kernel_rooflines(L, pool_dev, eager_iter, flush) if rank == 0 else (None, None)
    if rank == 0:
        peaks = measured_peaks()
        imgs = bs * nd * args.steps * world
        line = {"metric": "ResNet GAN G+D train img/s", "value": imgs / (ms / 1e3), "unit": "img/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "tf32" if args.conv_impl == "tf32" else "f32", "data": "synthetic",
                "config": workload_config(args.config, bs, world),
                "details": {"conv_impl": args.conv_impl, "cuda_graphs": resnet_graphs,
                            "l2": "activations per step (> 1 GB) exceed the 126 MB L2"},
                "e2e": {"value": imgs / (ms_e2e / 1e3), "unit": "img/s", "h2d_bytes_per_step": nd * bs * 3 * res * res * 4,
                        "d2h_bytes_per_step": 8},
                "gpu_launches": launches_per_iter * args.steps, "clocks": sampler.summary() if sampler else None}
        if roof is not None:
            roof["peak"] = peaks["tf"]; roof["frac"] = roof["achieved"] / peaks["tf"]
            roof["frac_of_tf32_peak"] = roof["achieved"] / (peaks["tf"] / 2)
            line["roofline"] = roof
        if glue is not None and glue["achieved"] is not None:
            glue["peak"] = peaks["hbm_gbs"]; glue["frac"] = glue["achieved"] / peaks["hbm_gbs"]
            line["roofline_glue"] = glue
        print(json.dumps(line), flush=True)
    if world > 1:
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(0)


GLUE_LAUNCHERS = ("cvt_bf16", "bias_act_fwd", "act_bwd", "blur_act_bwd", "axpby", "colsum", "scale_by", "pixelnorm_fwd", "pixelnorm_bwd", "blur3x3",
                  "upsample2x_fwd", "upsample2x_bwd", "pool_bias_act_fwd", "pool_bias_act_bwd", "style_epilogue_fwd",
                  "style_epilogue_bwd", "rgb_expand", "rgb_contract", "rgb_wgrad", "fade_up_blend", "fade_up_blend_bwd",
                  "fade_real", "interp_rows", "batchnorm_fwd", "batchnorm_bwd", "layernorm_fwd", "layernorm_bwd",
                  "layernorm_bwdbwd", "tanh_fwd", "tanh_bwd", "adam_ewma_multi")


def input_pipeline_roofline(peaks):
    """HBM roofline of the real-image input kernel (csrc/input.cu), outside the timed step: 96 uint8
    1024x1024 sources (302 MB > L2) -> the 128x128 fp32 batch the cfg2 step consumes; algorithmic bytes = 3 B per source pixel +
    12 B per output pixel; CUDA events on the launch stream.  Never allowed to take the bench line down with it."""
    try:
        import torch
        from gan_lab_b200 import _kernels as K
        n, hs, ho = 96, 1024, 128
        src = torch.randint(0, 256, (n, hs, hs, 3), dtype=torch.uint8, device="cuda")
        idx = torch.randperm(n)
        for _ in range(3):
            out = K.u8_box_resize_normalize(src, idx, (ho, ho), (.5,) * 3, (.5,) * 3)
        torch.cuda.synchronize()
        ts = []
        for _ in range(7):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = K.u8_box_resize_normalize(src, idx, (ho, ho), (.5,) * 3, (.5,) * 3)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        gbs = (n * hs * hs * 3 + out.numel() * 4) / ms / 1e6
        del src, out
        return {"bound": "hbm", "kernel": "box_resize_normalize_grouped_kernel (Pillow BOX resize + ToTensor + Normalize, uint8 1024x1024 -> "
                "fp32 128x128, 96 images per launch)", "achieved": gbs, "unit": "GB/s", "peak": peaks["hbm_gbs"],
                "frac": gbs / peaks["hbm_gbs"], "traffic": None, "images_per_s": n / ms * 1e3, "launch_us": ms * 1e3}
    except Exception as e:       # noqa: BLE001 -- an auxiliary figure
        return {"error": repr(e)[:300]}


def kernel_rooflines(L, x, main_iter, flush):
    """Both rooflines of BASELINE.json's metric from ONE recorded eager main iteration.

    Every launcher call of the iteration is recorded (arguments kept alive); then, per launcher kind, the recorded
    calls are re-issued back to back as one CUDA graph -- GPU bound, in step order, L2 flushed (160 MB write) before each
    replay -- between one pair of CUDA events on the launch stream.
      * dense convs (fprop / dgrad / wgrad implicit GEMMs): achieved = algorithmic FLOPs / device time  -> tensor roofline
      * glue launchers: achieved = compulsory bytes (every tensor argument read once + every output written once,
        4 B/element; Adam: 28 B/param + 8 B/param for the fused EWMA) / device time                   -> HBM roofline
    """
    import torch
    from gan_lab_b200 import _kernels as K
    calls = []
    orig = {}

    def tensors(obj):
        if torch.is_tensor(obj):
            yield obj
        elif isinstance(obj, (list, tuple)):
            for o in obj:
                yield from tensors(o)

    def nbytes(a, k, out):
        return 4.0 * sum(t.numel() for t in tensors(list(a) + list(k.values()) + [out]))

    def wrap(name, work_fn):
        f = getattr(K, name, None)
        if f is None:
            return
        orig[name] = f

        def g(*a, **k):
            out = f(*a, **k)
            calls.append((name, f, a, k, work_fn(a, k, out)))
            return out

        setattr(K, name, g)

    def f_fprop(a, k, out):
        w_ = a[1]
        return 2.0 * out.shape[0] * out.shape[2] * out.shape[3] * w_.shape[0] * w_.shape[1] * w_.shape[2] * w_.shape[3]

    def f_dgrad(a, k, out):
        gy_, w_ = a[0], a[1]
        return 2.0 * gy_.shape[0] * gy_.shape[2] * gy_.shape[3] * w_.shape[0] * w_.shape[1] * w_.shape[2] * w_.shape[3]

    def f_wgrad(a, k, out):
        gy_ = a[1]
        return 2.0 * gy_.shape[0] * gy_.shape[2] * gy_.shape[3] * out.numel()

    def f_adam(a, k, out):
        sizes, ewma_mode = a[1], a[-1]
        return float(sizes.sum().item()) * (28.0 + (8.0 if ewma_mode else 0.0))

    def f_cvt(a, k, out):
        return 6.0 * out.numel()

    # upsample / average pool folded into the convolution (glb_upconv_* / glb_downconv_*): EXECUTED FLOPs = 16 pre-summed taps per
    # low-resolution pixel (4/9 of the literal upsample -> conv / conv -> pool sequence)
    def f_fold(lo):
        def f(a, k, out):
            t = lo(a, out)
            w_ = a[1] if a[1].dim() == 4 and a[1].shape[2] == 3 else None
            ci_co = (w_.shape[0] * w_.shape[1]) if w_ is not None else a[0].shape[1] * a[1].shape[1]
            return 2.0 * t.shape[0] * t.shape[2] * t.shape[3] * 16.0 * ci_co
        return f

    conv_kinds = ("conv_fprop", "conv_dgrad", "conv_wgrad", "upconv_fprop", "upconv_dgrad", "upconv_wgrad",
                  "downconv_fprop", "downconv_dgrad", "downconv_wgrad")
    wrap("conv_fprop", f_fprop); wrap("conv_dgrad", f_dgrad); wrap("conv_wgrad", f_wgrad)
    wrap("upconv_fprop", f_fold(lambda a, out: a[0])); wrap("upconv_dgrad", f_fold(lambda a, out: out))
    wrap("upconv_wgrad", f_fold(lambda a, out: a[0]))
    wrap("downconv_fprop", f_fold(lambda a, out: out)); wrap("downconv_dgrad", f_fold(lambda a, out: a[0]))
    wrap("downconv_wgrad", f_fold(lambda a, out: a[1]))
    for name in GLUE_LAUNCHERS:
        wrap(name, f_adam if name == "adam_ewma_multi" else (f_cvt if name == "cvt_bf16" else nbytes))
    dp, L.dp = L.dp, None          # rank-local pass: no collective may run here (the other ranks are not in it)
    if dp is not None:
        dp.enabled = False         # ... including the ones the gradient hooks would launch
    try:
        main_iter(x)
        torch.cuda.synchronize()
    finally:
        L.dp = dp
        if dp is not None:
            dp.enabled = True
        for n, f in orig.items():
            setattr(K, n, f)

    def timed(sel, reps=5):
        """Median device time of the recorded launches of one kind, re-issued as ONE CUDA graph (as in the real step: no
        Python between launches), L2 flushed before every replay; eager re-issue if the capture is refused."""
        graph = None
        try:
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.no_grad(), torch.cuda.graph(graph):
                for _, f, a, k, _w in sel:
                    f(*a, **k)
        except Exception:
            graph = None
            torch.cuda.synchronize()
        times = []
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if graph is not None:
                graph.replay()
            else:
                with torch.no_grad():
                    for _, f, a, k, _w in sel:
                        f(*a, **k)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        times.sort()
        return times[len(times) // 2]

    if os.environ.get("GLB_BENCH_DUMP"):          # debugging aid: per-call device time of every recorded conv launch
        rows = []
        for name, f, a, k, work in [c for c in calls if c[0] in conv_kinds]:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.no_grad():
                f(*a, **k); torch.cuda.synchronize()
                e0.record(); f(*a, **k); e1.record()
            torch.cuda.synchronize()
            shp = [tuple(t.shape) for t in a if torch.is_tensor(t)]
            rows.append((e0.elapsed_time(e1) * 1e3, name, shp, work))
        agg = {}
        for us, name, shp, work in rows:
            key = (name, str(shp))
            agg.setdefault(key, [0, 0.0, work]); agg[key][0] += 1; agg[key][1] += us
        for key, (n, us, work) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
            print(f"[dump] {us:9.1f} us total  x{n:3d}  {work / (us / n) / 1e6:7.1f} TF/s  {key[0]} {key[1]}", file=sys.stderr)
    by = {}
    for kind in conv_kinds:
        sel = [c for c in calls if c[0] == kind]
        if sel:
            by[kind] = (sum(c[4] for c in sel), timed(sel), len(sel))
    tot_f = sum(v[0] for v in by.values())
    tot_ms = sum(v[1] for v in by.values())
    n_launch = sum(v[2] for v in by.values())
    # the folded layers execute 16 taps per low-resolution pixel where the literal upsample -> conv / conv -> pool sequence of the
    # reference executes 36: `achieved` counts what the kernels EXECUTE; the literal-sequence figure is reported beside it
    lit_f = sum(v[0] * (2.25 if n.startswith(("upconv", "downconv")) else 1.0) for n, v in by.items())
    conv = {"bound": "tensor", "kernel": "conv2d fprop/dgrad/wgrad family incl. the upsample- / pool-folded layers (tcgen05 implicit GEMM, "
                                       "operands " + K.get_conv_impl() + "; FFMA for uncovered shapes; bf16 mode: the operand conversion "
                                       "passes are counted under glue); FLOPs = executed by the kernels",
            "achieved": tot_f / (tot_ms / 1e3) / 1e12, "unit": "TFLOP/s", "traffic": None, "launches": n_launch,
            "avg_launch_us": 1e3 * tot_ms / max(n_launch, 1), "conv_ms_per_step": tot_ms,
            "executed_conv_gflop_per_step": tot_f / 1e9, "literal_sequence_conv_gflop_per_step": lit_f / 1e9,
            "tflops_on_literal_sequence_flops": lit_f / (tot_ms / 1e3) / 1e12,
            "by_kind_tflops": {n: v[0] / (v[1] / 1e3) / 1e12 for n, v in by.items()},
            "by_kind_ms": {n: round(v[1], 4) for n, v in by.items()}}
    # layers whose arithmetic intensity (algorithmic FLOPs / compulsory bytes: operands + output, 4 B each) lies
    # below the ridge of the machine -- 3-channel / 16-32-channel layers at high resolution, weight-bound layers with M <= 512
    # pixels -- are HBM-bound and belong on the HBM roofline; the rest on the tensor roofline.  Both classes timed separately.
    try:
        pk = measured_peaks()
        ridge = pk["tf"] * 1e12 / (pk["hbm_gbs"] * 1e9)
        cls = {"tensor_bound": [], "hbm_bound": []}
        for c in calls:
            if c[0] in conv_kinds:
                a = c[2]
                nb = 4.0 * sum(t.numel() for t in a if torch.is_tensor(t))
                out_elems = {"conv_fprop": lambda: a[0].shape[0] * a[1].shape[0] * (a[0].shape[2] + 2 * a[3] - a[1].shape[2] + 1) * (a[0].shape[3] + 2 * a[3] - a[1].shape[3] + 1),
                             "conv_dgrad": lambda: a[0].shape[0] * a[1].shape[1] * a[2][0] * a[2][1],
                             "conv_wgrad": lambda: a[1].shape[1] * a[0].shape[1] * a[2][0] * a[2][1],
                             "upconv_fprop": lambda: 4 * a[0].shape[0] * a[1].shape[0] * a[0].shape[2] * a[0].shape[3],
                             "upconv_dgrad": lambda: a[0].shape[0] * a[1].shape[1] * a[0].shape[2] * a[0].shape[3] // 4,
                             "upconv_wgrad": lambda: 9 * a[0].shape[1] * a[1].shape[1],
                             "downconv_fprop": lambda: a[0].shape[0] * a[1].shape[0] * a[0].shape[2] * a[0].shape[3] // 4,
                             "downconv_dgrad": lambda: 4 * a[0].shape[0] * a[1].shape[1] * a[0].shape[2] * a[0].shape[3],
                             "downconv_wgrad": lambda: 9 * a[0].shape[1] * a[1].shape[1]}[c[0]]()
                nb += 4.0 * out_elems
                cls["tensor_bound" if c[4] / nb >= ridge else "hbm_bound"].append((c, nb))
        split = {"ridge_flop_per_byte": ridge}
        for name, sel in cls.items():
            if not sel:
                continue
            ms_ = timed([c for c, _ in sel])
            fl, by_ = sum(c[4] for c, _ in sel), sum(nb for _, nb in sel)
            split[name] = {"launches": len(sel), "ms_per_step": ms_, "tflops": fl / (ms_ / 1e3) / 1e12, "gbs": by_ / (ms_ / 1e3) / 1e9,
                           "frac_of_tensor_peak": fl / (ms_ / 1e3) / 1e12 / pk["tf"], "frac_of_hbm_peak": by_ / (ms_ / 1e3) / 1e9 / pk["hbm_gbs"]}
        conv["split_by_roofline"] = split
    except Exception as e:          # noqa: BLE001 -- an auxiliary breakdown
        conv["split_by_roofline"] = {"error": repr(e)[:200]}

    # glue: launches that move >= 4 MB are HBM-bound and make up the roofline figure; the small ones (latents, 4x4 / 8x8
    # maps, scalars) are launch-latency bound and reported separately as time only
    BIG = 4 << 20
    gby, small_ms, small_n = {}, 0.0, 0
    for kind in GLUE_LAUNCHERS:
        sel = [c for c in calls if c[0] == kind]
        if not sel:
            continue
        big = [c for c in sel if c[4] >= BIG]
        sm = [c for c in sel if c[4] < BIG]
        if big:
            gby[kind] = (sum(c[4] for c in big), timed(big), len(big))
        if sm:
            small_ms += timed(sm, reps=3); small_n += len(sm)
    gb = sum(v[0] for v in gby.values())
    gms = sum(v[1] for v in gby.values())
    glue = {"bound": "hbm", "kernel": "fused glue launchers moving >= 4 MB (bias/act, blur, up/downsample, style epilogue, RGB, Adam+EWMA)",
            "achieved": gb / (gms / 1e3) / 1e9 if gms else None, "unit": "GB/s", "traffic": None,
            "launches": sum(v[2] for v in gby.values()), "glue_ms_per_step": gms, "algorithmic_mb_per_step": gb / 1e6,
            "small_launches": small_n, "small_launches_ms_per_step": small_ms,
            "by_kind_gbs": {n: round(v[0] / (v[1] / 1e3) / 1e9, 1) for n, v in gby.items()},
            "by_kind_ms": {n: round(v[1], 4) for n, v in gby.items()}}
    return conv, glue


def cpu_baseline(args):
    """The reference's CPU path timed on this box's host cores on a bounded sample (2 main iterations of the workload at its
    full batch after 1 warm-up, ~15 s): the unmodified reference from baseline/_ref when present (kind "reference"), else the
    oracle port."""
    import torch
    from oracle.reference_loader import reference_available
    model, res, init_res, bs, alpha = CONFIGS[args.config]
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sbs = bs // 4 if args.config == "cfg4" else bs
    kind = "reference" if (reference_available() and alpha is None) else "port"
    one, note = _reference_stepper(model, res, sbs, 2, 1) if kind == "reference" else _port_stepper(model, res, sbs)
    one()
    t0 = time.perf_counter()
    one(); one()
    dt = (time.perf_counter() - t0) / 2
    return {"value": sbs / dt, "unit": "img/s", "cores": threads, "kind": kind,
            "sample": f"2 main iterations (D step + G step) at the workload's full batch {sbs} after 1 warm-up, {threads} torch threads; {note}"}


def main():
    if os.environ.get("GLB_BENCH_FAULT"):          # debugging aid: dump every thread's Python stack if the run hangs
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["GLB_BENCH_FAULT"]), exit=True)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-dev", default="cpu", choices=["cpu", "cuda"],
                    help="with --impl reference: cpu = the contract's reference arm (host cores); cuda = the unmodified "
                         "reference through stock PyTorch on the GPU (an extra comparison, never the default)")
    ap.add_argument("--config", default="cfg2", choices=list(CONFIGS))
    ap.add_argument("--conv-impl", default=os.environ.get("GLB_CONV_IMPL", "tf32"), choices=["fp32", "tf32", "bf16"],
                    help="tf32 (default; the precision BASELINE.json names) | bf16 (opt-in: bf16 GEMM operands, fp32 everything else) | fp32 (exact FFMA)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--verbose", action="store_true", help="progress lines on stderr")
    ap.add_argument("--no-graphs", action="store_true", help="eager launches instead of CUDA-graph replay of the D/G steps")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_ours(args)


if __name__ == "__main__":
    main()
