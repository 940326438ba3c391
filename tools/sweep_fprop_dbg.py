#!/usr/bin/env python
"""Experiment: fprop duration when the A (activation) or B (weight) TMA loads are mostly skipped (results are wrong on
purpose) -- an upper bound on what operand reuse in shared memory could buy."""
import os, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import gan_lab_b200 as glb
from gan_lab_b200 import _kernels as K

def cl(t): return t.contiguous(memory_format=torch.channels_last)

def timeit(fn, reps=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

glb.set_conv_impl("tf32")
os.environ["GLB_FPROP_PAIR"] = "0"
for (N, H, Ci, Co) in [(8, 128, 128, 128), (8, 128, 256, 128), (8, 64, 256, 256), (8, 32, 512, 512)]:
    x = cl(torch.randn(N, Ci, H, H, device="cuda")); w = cl(torch.randn(Co, Ci, 3, 3, device="cuda"))
    fl = 2.0 * N * H * H * Co * Ci * 9
    out = []
    for dbg in (0, 1, 2):
        os.environ["GLB_FPROP_DBG"] = str(dbg)
        t = timeit(lambda: K.conv_fprop(x, w, None, 1, 1.0, 1.0, K.ACT_NONE, 0.2))
        out.append(f"dbg{dbg}: {t:6.1f}us {fl/t/1e6:5.0f}TF")
    print(f"N{N} {H}x{H} {Ci}->{Co}: " + " | ".join(out), flush=True)
