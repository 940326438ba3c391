#!/usr/bin/env python
"""Per-CTA clock64 timeline of the single-CTA tcgen05 fprop kernel (GLB_FPROP_TRACE): where do the cycles go?"""
import os, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import gan_lab_b200 as glb
from gan_lab_b200 import _kernels as K

def cl(t): return t.contiguous(memory_format=torch.channels_last)

glb.set_conv_impl("tf32")
os.environ["GLB_FPROP_PAIR"] = "0"
names = ["start", "setup", "1st operands", "tile0 MMAs issued", "all MMAs issued", "tile0 acc ready", "tile0 epilogue done", "last epilogue done"]
for (N, H, Ci, Co) in [(8, 128, 128, 128), (8, 64, 256, 256), (8, 32, 512, 512), (8, 16, 512, 512)]:
    x = cl(torch.randn(N, Ci, H, H, device="cuda")); w = cl(torch.randn(Co, Ci, 3, 3, device="cuda")); b = torch.randn(Co, device="cuda")
    f = lambda: K.conv_fprop(x, w, b, 1, 1.0, 1.0, K.ACT_LRELU, 0.2)
    os.environ.pop("GLB_FPROP_TRACE", None)
    for _ in range(3): f()
    buf = torch.zeros(148 * 8, dtype=torch.int64, device="cuda")
    os.environ["GLB_FPROP_TRACE"] = str(buf.data_ptr())
    f(); torch.cuda.synchronize()
    os.environ.pop("GLB_FPROP_TRACE", None)
    t = buf.view(148, 8).cpu()
    print(f"--- N{N} {H}x{H} {Ci}->{Co}  (cycles since CTA start; CTA 0 / CTA 73)")
    for cta in (0, 73):
        r = t[cta]
        if int(r[0]) == 0: continue
        print(f"  CTA {cta:3d}: " + " | ".join(f"{names[i]} {int(r[i]-r[0])}" for i in range(1, 8)))
