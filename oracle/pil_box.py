"""CPU restatement of the reference's real-image input chain.  TEST INFRASTRUCTURE ONLY -- never imported by the product.

The reference resizes every real sample on the host with torchvision `Resize(res, interpolation=PIL.Image.BOX)`, then
`ToTensor()` and `Normalize(mean, std)` (gan_lab/data_config.py:312-342; the Resize is rewritten at every resolution increase,
gan_lab/progan/learner.py:1099-1112).  The arithmetic lives in a third-party dependency that is not under /root/reference:
Pillow (requirements.txt pins nothing; installed here: Pillow 12.2.0).  Restated below from Pillow's published algorithm,
src/libImaging/Resample.c:
    precompute_coeffs        -> `coeffs()`  (double arithmetic, BOX filter = 1 on (-0.5, 0.5])
    normalize_coeffs_8bpc    -> 22-bit fixed point, round half away from zero
    ImagingResampleHorizontal_8bpc, then ImagingResampleVertical_8bpc on the uint8 result of the first pass
Pinned: tests/test_input_pipeline.py compares this restatement with Pillow itself (present on the build container and on the
GPU box) on integer, fractional, identity and up-scaling ratios; the CUDA kernel is then compared with both.
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def coeffs(in_size, out_size):
    """-> bounds int32 [out, 2] (first source index, count), kk int32 [out, ksize]   (Resample.c precompute_coeffs)."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 0.5 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = np.zeros(ksize)
        ww = 0.0
        for x in range(xmax):
            a = (x + xmin - center + 0.5) * ss
            w[x] = 1.0 if (-0.5 < a <= 0.5) else 0.0
            ww += w[x]
        if ww != 0.0:
            w[:xmax] /= ww
        for x in range(ksize):
            v = w[x] * (1 << PRECISION_BITS)
            kk[xx, x] = int(-0.5 + v) if w[x] < 0 else int(0.5 + v)
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _resample_axis(img, out_size, axis):
    in_size = img.shape[axis]
    b, kk = coeffs(in_size, out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.zeros((out_size,) + src.shape[1:], np.uint8)
    for xx in range(out_size):
        xmin, xmax = int(b[xx, 0]), int(b[xx, 1])
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), np.int64)
        for x in range(xmax):
            acc += src[xmin + x] * int(kk[xx, x])
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def box_resize_u8(img, out_h, out_w):
    """uint8 [..., H, W, 3] -> uint8 [..., out_h, out_w, 3]: horizontal pass, then vertical pass on its uint8 result."""
    h_axis, w_axis = img.ndim - 3, img.ndim - 2
    return _resample_axis(_resample_axis(img, out_w, w_axis), out_h, h_axis)


def to_tensor_normalize(img_u8, mean, std):
    """torchvision ToTensor + Normalize on uint8 [N, H, W, 3] -> float32 [N, 3, H, W] (fp32 ops in torchvision's order)."""
    import torch
    x = torch.from_numpy(np.ascontiguousarray(img_u8)).permute(0, 3, 1, 2).contiguous().to(torch.float32).div(255)
    m = torch.tensor(mean, dtype=torch.float32).view(1, 3, 1, 1)
    s = torch.tensor(std, dtype=torch.float32).view(1, 3, 1, 1)
    return x.sub_(m).div_(s)


def input_pipeline(images_u8, index, out_hw, mean, std, flip=None):
    """What the reference's DataLoader hands train() for samples `index` at resolution `out_hw`."""
    imgs = np.asarray(images_u8)
    if index is not None:
        imgs = imgs[np.asarray(index)]
    small = box_resize_u8(imgs, int(out_hw[0]), int(out_hw[1]))
    if flip is not None:
        f = np.asarray(flip).astype(bool)
        small = small.copy()
        small[f] = small[f][:, :, ::-1]
    return to_tensor_normalize(small, mean, std)
