"""The exact re-layouts the launchers apply below autograd (gan_lab_b200/_kernels.py), checked on CPU with plain torch
convolutions: pixel-pair packing of narrow layers (expanded weight, packed views, folded weight gradient) and the
zero-padding helpers.  No kernel is launched here; the GPU parity tests cover the same paths through the tcgen05 kernels."""
import pytest
import torch
import torch.nn.functional as F

from gan_lab_b200 import _kernels as K


def cl(t):
    return t.contiguous(memory_format=torch.channels_last)


@pytest.mark.parametrize("Ci,Co,S", [(16, 16, 3), (32, 16, 3), (16, 32, 3), (16, 16, 1), (8, 4, 3)])
def test_pixel_pair_packing_is_exact(Ci, Co, S):
    torch.manual_seed(Ci + Co + S)
    pad, N, H, W = (S - 1) // 2, 2, 6, 8
    x, w = cl(torch.randn(N, Ci, H, W, dtype=torch.float64)), cl(torch.randn(Co, Ci, S, S, dtype=torch.float64))
    gy = cl(torch.randn(N, Co, H, W, dtype=torch.float64))
    wp, xp, gyp = K.packed_weight(w, pad), K._pack_view(x), K._pack_view(gy)
    assert xp.data_ptr() == x.data_ptr() and xp.is_contiguous(memory_format=torch.channels_last)      # a view, no copy
    assert tuple(xp.shape) == (N, 2 * Ci, H, W // 2) and tuple(wp.shape) == (2 * Co, 2 * Ci, S, S)
    y = K._unpack_view(cl(F.conv2d(xp, wp, None, 1, pad)), Co)
    torch.testing.assert_close(y, F.conv2d(x, w, None, 1, pad), rtol=1e-12, atol=1e-12)
    gx = K._unpack_view(cl(torch.nn.grad.conv2d_input(xp.shape, wp, gyp, 1, pad)), Ci)
    torch.testing.assert_close(gx, torch.nn.grad.conv2d_input(x.shape, w, gy, 1, pad), rtol=1e-12, atol=1e-12)
    gw = K._fold_packed_wgrad(torch.nn.grad.conv2d_weight(xp, wp.shape, gyp, 1, pad), Co, Ci, S, pad)
    torch.testing.assert_close(gw, torch.nn.grad.conv2d_weight(x, w.shape, gy, 1, pad), rtol=1e-12, atol=1e-12)


def test_packed_weight_cache_follows_versions():
    w = cl(torch.randn(4, 4, 3, 3))
    a = K.packed_weight(w, 1)
    assert K.packed_weight(w, 1) is a
    w.mul_(2.0)                                   # version bump -> recomputed
    b = K.packed_weight(w, 1)
    assert b is not a and torch.equal(b, K.packed_weight(w, 1))
    K.weights_updated()                           # raw-pointer updates (the fused Adam) drop everything
    assert K.packed_weight(w, 1) is not b


def test_channel_padding_helpers():
    x = cl(torch.randn(2, 3, 4, 4))
    xp = K._pad_channels(x, 32)
    assert tuple(xp.shape) == (2, 32, 4, 4) and xp.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(xp[:, :3], x) and float(xp[:, 3:].abs().max()) == 0.0
    w = cl(torch.randn(3, 8, 3, 3))
    wp = K._pad_dim0(w, 32)
    assert tuple(wp.shape) == (32, 8, 3, 3) and torch.equal(wp[:3], w) and float(wp[3:].abs().max()) == 0.0
    # padded convolution == original on the real channels
    y = F.conv2d(torch.randn(2, 8, 4, 4), wp, None, 1, 1)
    assert float(y[:, 3:].abs().max()) == 0.0
