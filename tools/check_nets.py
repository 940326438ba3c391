#!/usr/bin/env python
"""Error budget of the StyleGAN generator on the B200 kernels: ours (fp32 kernels) and the reference's fp32 CPU result
(golden fixture) are both compared with the oracle evaluated in fp64 on the same draws.  Run on the GPU box."""
import math
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import gan_lab_b200 as glb  # noqa: E402
import gan_lab_b200._growth as growth  # noqa: E402
from gan_lab_b200.utils.latent_utils import TapeSource, set_random_source  # noqa: E402
from oracle import gan_oracle as O  # noqa: E402
from oracle.train_oracle import parse_gen_forward  # noqa: E402
import parity_cases as PC  # noqa: E402


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def main():
    dev = "cuda"
    growth.FMAP_MAX = 32
    for fname in PC.STYLE_NETS:
        g = torch.load(ROOT / "tests" / "golden" / fname, weights_only=False)
        res, fade = g["res"], g["fade_in"]
        for impl in ("fp32", "tf32"):
            glb.set_conv_impl(impl)
            L = PC._style_learner(PC._to(g, dev), fade, dev)
            G = L.gen_model
            G.load_state_dict(g["g_sd"]); G.train()
            set_random_source(TapeSource([(k, v.to(dev) if torch.is_tensor(v) else v) for k, v in g["tape"]], dev))
            img = G(g["z"].to(dev))
            G.zero_grad(); img.backward(g["gimg"].to(dev))
            set_random_source(None)
            # fp64 oracle on the same draws
            dr = parse_gen_forward(iter([("randn", g["z"])] + list(g["tape"])), "StyleGAN", 2 * (int(math.log2(res)) - 1))
            sd64 = {k: v.double().requires_grad_(True) for k, v in g["g_sd"].items()}
            ref = O.style_generator_forward(sd64, dr.z.double(), res=res, noise=[n.double() for n in dr.noise],
                                            z2=None if dr.z2 is None else dr.z2.double(), cutoff_idx=dr.cutoff_idx,
                                            alpha=g.get("alpha", 1.0) if fade else 1.0, fade_in=fade)
            ref.backward(g["gimg"].double())
            print(f"{fname} [{impl}] img: ours vs fp64 {rel(img, ref):.2e} | reference-fp32 vs fp64 {rel(g['img'], ref):.2e}")
            worst = []
            for k, p in G.named_parameters():
                r64 = sd64[k].grad
                if r64 is None or p.grad is None or float(r64.abs().max()) == 0:
                    continue
                worst.append((rel(p.grad, r64), rel(g["g_grads"][k], r64) if g["g_grads"][k] is not None else -1, k))
            worst.sort(reverse=True)
            for e_ours, e_ref, k in worst[:6]:
                print(f"    grad {k}: ours {e_ours:.2e} | reference-fp32 {e_ref:.2e}")


if __name__ == "__main__":
    main()
