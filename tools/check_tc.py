#!/usr/bin/env python
"""Tensor-core (tcgen05, TF32) conv kernels vs the exact fp32 FFMA kernels on the same inputs: error + timing sweep.
Run on the B200 box:  python tools/check_tc.py [--quick]"""
import os
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import gan_lab_b200 as glb  # noqa: E402
from gan_lab_b200 import _kernels as K  # noqa: E402

DEV = "cuda"


def cl(t):
    return t.contiguous(memory_format=torch.channels_last)


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-20))


def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def run(kind, N, H, W, Ci, Co, R=3, pad=1, reps=10):
    g = torch.Generator(device="cpu").manual_seed(N + H + Ci + Co)
    x = cl(torch.randn(N, Ci, H, W, generator=g)).to(DEV)
    w = cl(torch.randn(Co, Ci, R, R, generator=g)).to(DEV)
    b = torch.randn(Co, generator=g).to(DEV)
    Ho, Wo = H + 2 * pad - R + 1, W + 2 * pad - R + 1
    gy = cl(torch.randn(N, Co, Ho, Wo, generator=g)).to(DEV)
    flops = 2.0 * N * Ho * Wo * Co * Ci * R * R
    if not K.tc_covers(kind, N, H, W, Ci, Co, R, R, pad):
        print(f"{kind:6s} N{N} {H}x{W} {Ci}->{Co} k{R}: not covered"); return
    if kind == "fprop":
        f = lambda: K.conv_fprop(x, w, b, pad, 0.37, 0.5, K.ACT_LRELU, 0.2)
    elif kind == "dgrad":
        f = lambda: K.conv_dgrad(gy, w, (H, W), pad, 0.37)
    else:
        f = lambda: K.conv_wgrad(x, gy, (R, R), pad, 0.37)
    glb.set_conv_impl("fp32"); ref = f(); t_ref = timeit(f, 3)
    impl = os.environ.get("GLB_CHECK_IMPL", "tf32")          # tf32 | bf16 (bf16: operand conversions are cached, i.e. not in the time)
    if impl == "bf16" and not K.bf16_covers(kind, N, H, W, Ci, Co, R, R, pad):
        print(f"{kind:6s} N{N} {H}x{W} {Ci}->{Co} k{R}: not covered by bf16"); return
    glb.set_conv_impl(impl); out = f(); t_tc = timeit(f, reps)
    err = rel(out, ref)
    flag = "OK " if err < (5e-3 if impl == "tf32" else 1.5e-2) else "BAD"
    print(f"{flag} {kind:6s} N{N} {H}x{W} {Ci}->{Co} k{R}: rel err {err:.2e}  tc {t_tc*1e3:8.1f} us {flops/t_tc/1e9:7.1f} TF | "
          f"fp32 {t_ref*1e3:8.1f} us {flops/t_ref/1e9:6.1f} TF", flush=True)


def main():
    quick = "--quick" in sys.argv
    kinds = [a for a in sys.argv[1:] if a in ("fprop", "dgrad", "wgrad")] or ["fprop", "dgrad", "wgrad"]
    shapes = [(8, 4, 4, 512, 512), (8, 8, 8, 512, 512), (8, 16, 16, 512, 512), (8, 32, 32, 512, 512),
              (8, 64, 64, 512, 256), (8, 64, 64, 256, 256), (8, 128, 128, 256, 128), (8, 128, 128, 128, 128),
              (8, 64, 64, 256, 512), (8, 128, 128, 128, 256), (2, 16, 16, 64, 32), (3, 5, 7, 32, 64), (4, 8, 8, 96, 128)]
    if quick:
        shapes = shapes[:4] + shapes[-3:]
    for kind in kinds:
        for s in shapes:
            try:
                run(kind, *s)
            except Exception as e:  # keep going: one bad shape should not hide the others
                print(f"ERR {kind} {s}: {e}", flush=True)
        if kind != "wgrad":
            try:
                run(kind, 8, 128, 128, 128, 128, R=1, pad=0)
            except Exception as e:
                print(f"ERR {kind} 1x1: {e}", flush=True)


if __name__ == "__main__":
    main()
