#!/usr/bin/env python
"""Achieved HBM GB/s of the glue launchers on the cfg2 shapes (B200, CUDA events, L2 flushed before every launch).

    python tools/glue_bw.py [reps]

Per launcher: compulsory bytes (every tensor argument read once + every output written once) / median device time.
`copy` rows (torch's own elementwise copy of the same volume) show what a plain streaming kernel reaches at that size.
Environment switches read by the library (csrc/glue.cu) select kernel variants for A/B runs."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import gan_lab_b200 as glb  # noqa: E402,F401
from gan_lab_b200 import _kernels as K  # noqa: E402

DEV = "cuda"


def cl(t):
    return t.contiguous(memory_format=torch.channels_last)


def nbytes(args, out):
    def ts(o):
        if torch.is_tensor(o):
            yield o
        elif isinstance(o, (list, tuple)):
            for x in o:
                yield from ts(x)
    return 4.0 * sum(t.numel() for t in ts(list(args) + [out]))


def timeit(fn, reps, flush):
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], out


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 15
    flush = torch.empty(192 * 1024 * 1024 // 4, device=DEV)
    rows = []
    shapes = [(8, 128, 128, 128), (8, 256, 64, 64), (8, 512, 32, 32), (8, 512, 16, 16)]
    for (N, C, H, W) in shapes:
        x, y, gy = (cl(torch.randn(N, C, H, W, device=DEV)) for _ in range(3))
        b = torch.randn(C, device=DEV)
        noise = torch.randn(N, 1, H, W, device=DEV)
        style = torch.randn(N, 2 * C, device=DEV)
        img = torch.randn(N, 3, H, W, device=DEV)
        wf, wt = torch.randn(C, 3, 1, 1, device=DEV), torch.randn(3, C, 1, 1, device=DEV)
        tag = f"N{N} C{C} {H}x{W}"
        cases = {
            "copy (torch)": (lambda: x.clone(), (x,)),
            "bias_act_fwd": (lambda: K.bias_act_fwd(x, b, 1.0, K.ACT_LRELU, 0.2), (x, b)),
            "act_bwd +bias": (lambda: K.act_bwd(gy, y, True, 1.0, K.ACT_LRELU, 0.2), (gy, y)),
            "act_bwd": (lambda: K.act_bwd(gy, y, False, 1.0, K.ACT_LRELU, 0.2), (gy, y)),
            "axpby": (lambda: K.axpby(x, y, 0.5, 0.5), (x, y)),
            "blur3x3": (lambda: K.blur3x3(x), (x,)),
            "upsample2x_fwd": (lambda: K.upsample2x_fwd(x), (x,)),
            "upsample2x_bwd": (lambda: K.upsample2x_bwd(gy), (gy,)),
            "pool_bias_act_fwd": (lambda: K.pool_bias_act_fwd(x, b, 1.0, K.ACT_LRELU, 0.2), (x, b)),
            "rgb_expand": (lambda: K.rgb_expand(img, wf, 1, 3, C, b, False, 0.3, 1.0, K.ACT_LRELU, 0.2), (img, wf, b)),
            "rgb_contract": (lambda: K.rgb_contract(x, wt, C, 1, None, False, 0.3, 1.0), (x, wt)),
            "rgb_wgrad": (lambda: K.rgb_wgrad(img, gy, (C, 3, 1, 1), 1, 3, False, 0.3), (img, gy)),
        }
        se_out = {}

        def se_fwd():
            se_out["o"] = K.style_epilogue_fwd(x, noise, b, b, style, 0.2, 1e-8)
            return se_out["o"][0]

        cases["style_epilogue_fwd"] = (se_fwd, (x, noise, style))
        se_fwd()
        cases["style_epilogue_bwd"] = (lambda: K.style_epilogue_bwd(gy, x, noise, b, b, style, se_out["o"][1], 0.2)[:2],
                                       (gy, x, noise, style))
        for name, (fn, args) in cases.items():
            for _ in range(2):
                fn()
            ms, out = timeit(fn, reps, flush)
            by = nbytes(args, out)
            rows.append((tag, name, by / 1e6, ms * 1e3, by / (ms / 1e3) / 1e9))
    # Adam over cfg2-sized parameter sets
    print(f"{'shape':<18} {'launcher':<20} {'MB':>8} {'us':>8} {'GB/s':>8}")
    for r in rows:
        print(f"{r[0]:<18} {r[1]:<20} {r[2]:8.1f} {r[3]:8.1f} {r[4]:8.0f}")


if __name__ == "__main__":
    main()
