"""Compact, size-independent summaries of large tensors.  TEST INFRASTRUCTURE ONLY.

The full-width cfg2 fixture (tests/golden/style_cfg2_fullwidth_step.pt) cannot hold 49 M parameters and as many gradients;
per tensor it holds instead: the L2 norm, the max-norm, a strided sample, and a few random +-1 projections.  A projection of
an error vector e onto a random sign vector has magnitude ~ ||e||_2, so |proj_mine - proj_ref| / ||ref||_2 estimates the
relative L2 error of the WHOLE tensor, not just of the sampled elements.
"""
import zlib

import torch

N_SAMPLE = 2048
N_PROJ = 4


def _seed(name: str) -> int:
    return zlib.crc32(name.encode()) & 0x7FFFFFFF


def sample_index(n: int) -> torch.Tensor:
    m = min(n, N_SAMPLE)
    return torch.arange(m, dtype=torch.int64) * (n // m)


def sign_vectors(name: str, n: int) -> torch.Tensor:
    gen = torch.Generator().manual_seed(_seed(name))
    # Tensor.random_ rather than torch.randint: fixture generation tapes the module-level torch.rand* functions
    return torch.empty((N_PROJ, n), dtype=torch.int8).random_(0, 2, generator=gen) * 2 - 1


def summarize(name: str, t: torch.Tensor) -> dict:
    """`t` in its LOGICAL (row-major over its shape) order, whatever its memory format."""
    f = t.detach().reshape(-1).to("cpu", torch.float64)
    n = f.numel()
    return dict(shape=tuple(t.shape), norm=float(f.norm()), absmax=float(f.abs().max()),
                sample=f[sample_index(n)].to(torch.float32), proj=sign_vectors(name, n).to(torch.float64) @ f)


def compare(name: str, t: torch.Tensor, ref: dict) -> dict:
    """-> dict(l2 = estimated relative L2 error of the whole tensor, smax = max-norm relative error on the strided sample,
    norm = relative difference of the L2 norms)."""
    mine = summarize(name, t)
    assert mine["shape"] == tuple(ref["shape"]), (name, mine["shape"], ref["shape"])
    scale = max(ref["norm"], 1e-30)
    return dict(l2=float((mine["proj"] - ref["proj"]).abs().max()) / scale,
                smax=float((mine["sample"] - ref["sample"]).abs().max()) / max(ref["absmax"], 1e-30),
                norm=abs(mine["norm"] - ref["norm"]) / scale)


def compare_summaries(mine: dict, ref: dict) -> dict:
    """Both sides already summarised (same tensor names): -> dict(l2, smax, norm) as `compare`."""
    scale = max(ref["norm"], 1e-30)
    return dict(l2=float((mine["proj"] - ref["proj"]).abs().max()) / scale,
                smax=float((mine["sample"] - ref["sample"]).abs().max()) / max(ref["absmax"], 1e-30),
                norm=abs(mine["norm"] - ref["norm"]) / scale)


def class_stats(per_tensor: dict) -> dict:
    """worst / median of the per-tensor errors of one class of tensors."""
    import statistics
    l2 = [c["l2"] for c in per_tensor.values()]
    sm = [c["smax"] for c in per_tensor.values()]
    return dict(l2=max(l2), smax=max(sm), l2_median=statistics.median(l2), smax_median=statistics.median(sm),
                worst=max(per_tensor, key=lambda k: per_tensor[k]["l2"]))
