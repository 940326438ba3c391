"""Achieved HBM bandwidth of the input-pipeline kernel (csrc/input.cu) on dataset-shaped sources, CUDA events on the launch
stream, sources larger than the 126 MB L2.  Algorithmic bytes = 3 B per source pixel read once + 12 B per output pixel."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from gan_lab_b200 import _kernels as K  # noqa: E402


def run(n, hs, ho, reps=10):
    src = torch.randint(0, 256, (n, hs, hs, 3), dtype=torch.uint8, device="cuda")
    idx = torch.randperm(n)
    for _ in range(3):
        out = K.u8_box_resize_normalize(src, idx, (ho, ho), (.5,) * 3, (.5,) * 3)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = K.u8_box_resize_normalize(src, idx, (ho, ho), (.5,) * 3, (.5,) * 3)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    ms = ts[len(ts) // 2]
    nbytes = n * hs * hs * 3 + out.numel() * 4
    return {"images": n, "src": hs, "dst": ho, "ms": round(ms, 4), "GB/s": round(nbytes / ms / 1e6, 1),
            "img/s": round(n / ms * 1e3)}


if __name__ == "__main__":
    rows = [run(96, 1024, 128), run(96, 1024, 1024), run(96, 1024, 32), run(1024, 256, 128), run(2048, 128, 128),
            run(96, 1024, 512)]
    for r in rows:
        print(json.dumps(r))
