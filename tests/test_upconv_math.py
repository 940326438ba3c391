"""The index algebra behind glb_upconv_* (csrc/conv_tc.cu), restated in torch on the CPU and checked against the literal
composition upsample2x -> conv3x3 (oracle/kernel_contracts.py): phase weights, phase offsets of the forward, the (phase, flipped
tap) order of the data gradient and the fold of the 16 phase/tap weight gradients back to 3 x 3.  The CUDA kernels themselves are
checked on the GPU (tests/test_gpu_parity.py::test_upconv_family_vs_contract)."""
import torch
import torch.nn.functional as F

import oracle.kernel_contracts as KC


def _taps(d, a):                       # taps of the 3-tap axis summed into tap a of phase d (up_taps in conv_tc.cu)
    return {(0, 0): (0,), (0, 1): (1, 2), (1, 0): (0, 1), (1, 1): (2,)}[(d, a)]


def phase_weights(w):                  # wp[ph][co][a][b][ci]
    Co, Ci = w.shape[:2]
    wp = torch.zeros(4, Co, 2, 2, Ci, dtype=w.dtype)
    for ph in range(4):
        dy, dx = ph >> 1, ph & 1
        for a in range(2):
            for b in range(2):
                for r in _taps(dy, a):
                    for s in _taps(dx, b):
                        wp[ph, :, a, b, :] += w[:, :, r, s]
    return wp


def shifted(x, dh, dw):                # x[n, c, i + dh, j + dw] with zero fill
    N, C, H, W = x.shape
    xp = F.pad(x, (1, 1, 1, 1))
    return xp[:, :, 1 + dh:1 + dh + H, 1 + dw:1 + dw + W]


def test_phase_forward_dgrad_wgrad_equal_the_composition():
    torch.manual_seed(0)
    N, Ci, Co, H, W = 2, 5, 4, 6, 7
    x = torch.randn(N, Ci, H, W, dtype=torch.float64)
    w = torch.randn(Co, Ci, 3, 3, dtype=torch.float64)
    gy = torch.randn(N, Co, 2 * H, 2 * W, dtype=torch.float64)
    wp = phase_weights(w)
    # forward: y[2i+dy, 2j+dx] = sum_ab x[i + a - (1-dy), j + b - (1-dx)] . wp[ph][:, a, b, :]
    y = torch.zeros(N, Co, 2 * H, 2 * W, dtype=torch.float64)
    for ph in range(4):
        dy, dx = ph >> 1, ph & 1
        for a in range(2):
            for b in range(2):
                y[:, :, dy::2, dx::2] += torch.einsum("nchw,oc->nohw", shifted(x, a - (1 - dy), b - (1 - dx)), wp[ph, :, a, b, :])
    torch.testing.assert_close(y, KC.upconv_fprop(x, w, None, 1.0, 1.0, 0, 0.2).contiguous(), rtol=1e-12, atol=1e-12)
    # data gradient: K index (ph, a', b') with wt[ci][ph*4 + a'*2 + b'][co] = wp[ph][co][1-a'][1-b'][ci], offset (a' - dy, b' - dx)
    gx = torch.zeros(N, Ci, H, W, dtype=torch.float64)
    for t16 in range(16):
        ph, a_, b_ = t16 >> 2, (t16 >> 1) & 1, t16 & 1
        dy, dx = ph >> 1, ph & 1
        wt = wp[ph, :, 1 - a_, 1 - b_, :]                                   # [co][ci]
        gx += torch.einsum("nohw,oc->nchw", shifted(gy[:, :, dy::2, dx::2], a_ - dy, b_ - dx), wt)
    torch.testing.assert_close(gx, KC.upconv_dgrad(gy, w, 1.0).contiguous(), rtol=1e-12, atol=1e-12)
    # weight gradient: gwp[co][ph*4 + a*2 + b][ci] = sum gy_ph . x shifted by (a + dy - 1, b + dx - 1); fold to 3 x 3
    gwp = torch.zeros(Co, 16, Ci, dtype=torch.float64)
    for t16 in range(16):
        ph, a, b = t16 >> 2, (t16 >> 1) & 1, t16 & 1
        dy, dx = ph >> 1, ph & 1
        gwp[:, t16, :] = torch.einsum("nohw,nchw->oc", gy[:, :, dy::2, dx::2], shifted(x, a + dy - 1, b + dx - 1))
    gw = torch.zeros(Co, Ci, 3, 3, dtype=torch.float64)
    for r in range(3):
        for s in range(3):
            ay = (0 if r == 0 else 1, 1 if r == 2 else 0)
            ax = (0 if s == 0 else 1, 1 if s == 2 else 0)
            for dy in range(2):
                for dx in range(2):
                    gw[:, :, r, s] += gwp[:, (dy * 2 + dx) * 4 + ay[dy] * 2 + ax[dx], :]
    torch.testing.assert_close(gw, KC.upconv_wgrad(x, gy, 1.0).contiguous(), rtol=1e-12, atol=1e-12)


def test_downconv_is_the_adjoint_structure_of_upconv():
    """glb_downconv_* runs on the upconv kernels: avgpool(conv(x, w)) = 0.25 * upconv_dgrad(x; w'), its data gradient =
    0.25 * upconv_fprop(gy; w'), its weight gradient = 0.25 * flip/transpose(upconv_wgrad(gy, x)), w' = flipped, transposed w."""
    torch.manual_seed(1)
    N, Ci, Co, H, W = 2, 5, 4, 3, 4
    x = torch.randn(N, Ci, 2 * H, 2 * W, dtype=torch.float64)
    w = torch.randn(Co, Ci, 3, 3, dtype=torch.float64)
    gy = torch.randn(N, Co, H, W, dtype=torch.float64)
    wprime = w.flip(2, 3).permute(1, 0, 2, 3).contiguous()
    kw = dict(rtol=1e-12, atol=1e-12)
    torch.testing.assert_close(KC.downconv_fprop(x, w, None, 1.0, 1.0, 0, 0.2).contiguous(),
                               0.25 * KC.upconv_dgrad(x, wprime, 1.0).contiguous(), **kw)
    torch.testing.assert_close(KC.downconv_dgrad(gy, w, 1.0).contiguous(),
                               0.25 * KC.upconv_fprop(gy, wprime, None, 1.0, 1.0, 0, 0.2).contiguous(), **kw)
    torch.testing.assert_close(KC.downconv_wgrad(x, gy, 1.0).contiguous(),
                               (0.25 * KC.upconv_wgrad(gy, x, 1.0)).flip(2, 3).permute(1, 0, 2, 3).contiguous(), **kw)
