#!/usr/bin/env python
"""A handful of representative tensor-core conv launches (cfg2 shapes) for `ncu --set full` captures."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import gan_lab_b200 as glb  # noqa: E402
from gan_lab_b200 import _kernels as K  # noqa: E402


def cl(t):
    return t.contiguous(memory_format=torch.channels_last)


def main():
    glb.set_conv_impl("tf32")
    dev = "cuda"
    shapes = [(8, 32, 32, 512, 512), (8, 128, 128, 128, 128), (8, 64, 64, 256, 256), (8, 16, 16, 512, 512), (8, 4, 4, 512, 512)]
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    for (N, H, W, Ci, Co) in shapes:
        x = cl(torch.randn(N, Ci, H, W, device=dev)); w = cl(torch.randn(Co, Ci, 3, 3, device=dev))
        gy = cl(torch.randn(N, Co, H, W, device=dev)); b = torch.randn(Co, device=dev)
        for _ in range(2):
            flush.zero_()
            K.conv_fprop(x, w, b, 1, 0.37, 1.0, K.ACT_LRELU, 0.2)
            flush.zero_()
            K.conv_wgrad(x, gy, (3, 3), 1, 0.37)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
