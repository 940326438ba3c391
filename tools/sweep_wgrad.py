#!/usr/bin/env python
"""wgrad duration vs split-K factor (GLB_WGRAD_SPLITS) for a few cfg2 shapes."""
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
for (N, H, Ci, Co) in [(8, 32, 512, 512), (8, 64, 256, 256), (8, 128, 128, 128), (8, 128, 128, 256), (8, 16, 512, 512)]:
    x = cl(torch.randn(N, Ci, H, H, device="cuda")); gy = cl(torch.randn(N, Co, H, H, device="cuda"))
    fl = 2.0 * N * H * H * Co * Ci * 9
    out = []
    for sp in (1, 2, 3, 4, 6, 8, 16, 32):
        os.environ["GLB_WGRAD_SPLITS"] = str(sp)
        t = timeit(lambda: K.conv_wgrad(x, gy, (3, 3), 1, 1.0))
        out.append(f"s{sp}: {t:6.1f}us {fl/t/1e6:5.0f}TF")
    print(f"N{N} {H}x{H} {Ci}->{Co}: " + " | ".join(out), flush=True)
