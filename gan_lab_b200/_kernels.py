"""Kernel launchers: torch tensors in, torch tensors out, raw pointers across the C-ABI.

Each launcher checks device/dtype/layout, allocates the outputs with torch (the caller owns all
memory, SURVEY.md section 8b "Ownership") and enqueues the kernel on torch's current stream.
Feature maps are logical NCHW tensors in torch.channels_last memory format (= NHWC in memory);
RGB images are plain NCHW-contiguous; conv weights are logical OIHW in channels_last (= KRSC).
No CPU path: a non-CUDA tensor raises.
"""
from __future__ import annotations

import os

import torch

from ._lib import LIB, GlbError

ACT_NONE, ACT_LRELU = 0, 1
IMPL_FP32, IMPL_TF32 = 0, 1

_state = {"conv_impl": "fp32", "launches": 0}


def set_conv_impl(mode: str):
    """'fp32' = exact FFMA implicit GEMM everywhere; 'tf32' = tcgen05 tensor-core kernels wherever the
    shape is covered (TF32 operands, fp32 accumulation), FFMA kernels for the rest; 'bf16' = tcgen05 kind::f16 on bf16
    copies of the two GEMM operands (round to nearest; fp32 accumulation, bias, activation and storage) wherever THAT is
    covered (channel counts multiples of 64), the 'tf32' choice elsewhere."""
    assert mode in ("fp32", "tf32", "bf16")
    _state["conv_impl"] = mode
    weights_updated()


def get_conv_impl() -> str:
    return _state["conv_impl"]


def launch_count() -> int:
    return _state["launches"]


def _call(name, *args):
    _state["launches"] += 1
    LIB.call(name, *args)


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _chk(*ts):
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise GlbError("gan_lab_b200 kernels need CUDA tensors (there is no CPU fallback)")
        if t.dtype != torch.float32:
            raise GlbError(f"gan_lab_b200 kernels are fp32; got {t.dtype}")


def _p(t):
    return None if t is None else t.data_ptr()


def nhwc(x: torch.Tensor) -> torch.Tensor:
    """4-D tensor -> same logical tensor, channels_last memory (no copy if already so)."""
    assert x.dim() == 4
    if x.is_contiguous(memory_format=torch.channels_last):
        return x
    return x.contiguous(memory_format=torch.channels_last)


def _new_nhwc(n, c, h, w, like):
    return torch.empty((n, c, h, w), device=like.device, dtype=torch.float32, memory_format=torch.channels_last)


def _flat(t):
    return None if t is None else t.detach().reshape(-1).contiguous()


# --------------------------------------------------------------------------- zeroed accumulators
# Per-channel gradient accumulators (bias / noise-weight / RGB-weight gradients, penalty sums) are filled by atomics and need a
# zeroed buffer: ~80 of them per StyleGAN iteration, each a `torch.zeros` = one fill kernel.  Inside a learner step they are
# carved from ONE arena per step kind that is cleared by a single memset when the next step of that kind begins (a captured
# step replays that memset); outside a step (tests, metrics) `_zeros` is plain `torch.zeros`.
class _ZeroArena(object):
    CAP = 1 << 19                     # floats (2 MB): a cfg2 step uses ~60 K; cleared as a whole (~1 us), whatever was used

    def __init__(self):
        self.buf, self.off = None, 0

    def begin(self, device):
        device = torch.device(device)
        if device.index is None:
            device = torch.device(device.type, torch.cuda.current_device())
        if self.buf is None or self.buf.device != device:
            if torch.cuda.is_current_stream_capturing():
                self.buf = None       # never allocate the arena inside a capture (it must outlive the graph's pool)
                return
            self.buf = torch.zeros(self.CAP, device=device, dtype=torch.float32)
        else:
            self.buf.zero_()
        self.off = 0

    def take(self, n):
        m = (n + 3) & ~3              # 16-byte aligned slices
        if self.buf is None or self.off + m > self.CAP:
            return None
        out = self.buf[self.off:self.off + n]
        self.off += m
        return out


_arenas = {}
_active_arena = [None]
_ARENA_OFF = bool(__import__("os").environ.get("GLB_NO_ARENA"))       # A/B measurements


def begin_step(tag, device):
    """Called by the learners at the start of a discriminator / generator step (one arena per `tag`: the gradients of one
    step kind stay valid until the next step of the SAME kind begins)."""
    if _ARENA_OFF:
        return
    a = _arenas.get(tag)
    if a is None:
        a = _arenas[tag] = _ZeroArena()
    a.begin(device)
    _active_arena[0] = a


def end_step():
    _active_arena[0] = None


def _zeros(shape, like):
    n = 1
    for d in (shape if isinstance(shape, (tuple, list, torch.Size)) else (shape,)):
        n *= int(d)
    a = _active_arena[0]
    if a is not None and a.buf is not None and a.buf.device == like.device:
        t = a.take(n)
        if t is not None:
            return t.view(shape)
    return torch.zeros(shape, device=like.device, dtype=torch.float32)


# --------------------------------------------------------------------------- conv family
_KIND = {"fprop": 0, "dgrad": 1, "wgrad": 2}


def tc_covers(kind: str, N, H, W, Ci, Co, R, S, pad) -> bool:
    """Does the tcgen05 kernel of `kind` cover this shape?  Answered by the library itself (csrc/conv_tc.cu)."""
    return bool(LIB.fn("glb_conv2d_tc_covers")(_KIND[kind], N, H, W, Ci, Co, R, S, pad))


def _tc_mode() -> bool:
    return _state["conv_impl"] in ("tf32", "bf16")


def _impl_for(kind, N, H, W, Ci, Co, R, S, pad):
    if _tc_mode() and tc_covers(kind, N, H, W, Ci, Co, R, S, pad):
        return IMPL_TF32
    return IMPL_FP32


def bf16_covers(kind: str, N, H, W, Ci, Co, R, S, pad) -> bool:
    return bool(LIB.fn("glb_conv2d_bf16_covers")(_KIND[kind], N, H, W, Ci, Co, R, S, pad))


def _use_bf16(kind, N, H, W, Ci, Co, R, S, pad) -> bool:
    return _state["conv_impl"] == "bf16" and bf16_covers(kind, N, H, W, Ci, Co, R, S, pad)


_wt_cache = {}


def weights_updated():
    """Called by the optimiser after it has rewritten parameters through raw pointers (no torch version bump)."""
    _wt_cache.clear()
    _wp_cache.clear()
    _pk_cache.clear()
    _b16_cache.clear()
    _up_cache.clear()


# ---- bf16 operand copies (conv_impl "bf16") -----------------------------------------------------------------------------
# One conversion pass per tensor and use site group: an activation is converted when its forward convolution runs and found
# again by the weight-gradient kernel of the backward pass, an incoming gradient is shared by dgrad and wgrad; weights (and their
# flipped / transposed copies) are converted once per optimiser step.  Entries keep their source tensor alive, so its address
# cannot be recycled while the entry exists; everything is dropped at every optimiser step (weights_updated).
_b16_cache = {}


def cvt_bf16(t):
    """fp32 -> bf16 (round to nearest even), same logical shape and memory format; one streaming pass (6 bytes per element)."""
    _chk(t)
    if not (t.is_contiguous() or (t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last))):
        raise GlbError("cvt_bf16: the source must be dense in channels_last or contiguous layout")
    out = torch.empty_strided(t.shape, t.stride(), device=t.device, dtype=torch.bfloat16)    # same element order as the source
    _call("glb_cvt_f32_bf16", _p(t), _p(out), t.numel(), _stream())
    return out


def _to_bf16(t, tag=""):
    key = (t.data_ptr(), tuple(t.shape), tag)
    hit = _cache_get(_b16_cache, key, t._version)
    if hit is not None:
        return hit
    out = cvt_bf16(t)
    if len(_b16_cache) >= 512:
        _b16_cache.clear()
    return _cache_put(_b16_cache, key, t._version, out, t)


def bf16_operand(t):
    """bf16 copy of an NHWC activation / gradient (cached per tensor version)."""
    return _to_bf16(nhwc(t))


def bf16_weight(w, transposed=False):
    """bf16 copy of the KRSC weights, or of their flipped / transposed copy (what dgrad multiplies by)."""
    return _to_bf16(transposed_weight(w), "wt") if transposed else _to_bf16(nhwc(w), "w")


def _cache_put(cache, key, version, tensor, w):
    """Cache entries remember the stream that produced them and an event recorded behind the producing kernels: a hit from
    ANOTHER stream (the learners run independent discriminator passes on two streams) waits for that event first."""
    ev = None
    first = next(t for t in tensor if t is not None) if isinstance(tensor, tuple) else tensor
    on_gpu = first.is_cuda
    if on_gpu:
        ev = torch.cuda.Event()
        ev.record()
    cache[key] = (version, tensor, w, ev, torch.cuda.current_stream().cuda_stream if on_gpu else None)
    return tensor


def _cache_get(cache, key, version):
    hit = cache.get(key)
    if hit is None or hit[0] != version:
        return None
    if hit[3] is not None and hit[4] != torch.cuda.current_stream().cuda_stream:
        torch.cuda.current_stream().wait_event(hit[3])
    return hit[1]


def transposed_weight(w):
    """wt[Ci][R][S][Co] with flipped taps (what the tensor-core dgrad multiplies by).  Cached per weight tensor and
    version so the re-layout runs once per optimiser step however many backward passes reuse the weight; the entry
    keeps `w` alive, so its address cannot be recycled for another tensor while the entry exists."""
    key = (w.data_ptr(), tuple(w.shape))
    hit = _cache_get(_wt_cache, key, w._version)
    if hit is not None:
        return hit
    Co, Ci, R, S = w.shape
    wt = torch.empty((Ci, Co, R, S), device=w.device, dtype=torch.float32, memory_format=torch.channels_last)
    _call("glb_conv2d_weight_transpose", _p(w), _p(wt), Co, R, S, Ci, _stream())
    if len(_wt_cache) >= 96:
        _wt_cache.clear()
    return _cache_put(_wt_cache, key, w._version, wt, w)


_wp_cache = {}


def _pad_ci(kind, N, H, W, Ci, Co, R, S, pad):
    """Ragged input-channel counts (the 513 channels behind the minibatch-stddev concat, progan/architectures.py:219-226)
    miss the tensor-core kernels' Ci % 32 (fprop / wgrad) or Ci % 128 (dgrad output tile) rule and would drop to the FFMA
    implicit GEMM (58 us per launch at cfg2).  Zero-padding Ci up to a multiple of 128 is exact -- the extra input channels
    are zeros, the extra weight columns / gradient columns are dropped -- and ~4x faster.  Returns the padded Ci or None."""
    if not _tc_mode() or tc_covers(kind, N, H, W, Ci, Co, R, S, pad):
        return None
    if Ci < 32:
        cip = 32            # 3-channel image convolutions of the ResNet nets (FFMA kernel: 2.7 ms per wgrad launch at cfg5)
    elif Ci >= 128 and Ci % 128 != 0:
        cip = (Ci + 127) // 128 * 128
    else:
        return None
    return cip if tc_covers(kind, N, H, W, cip, Co, R, S, pad) else None


def _pad_co(kind, N, H, W, Ci, Co, R, S, pad):
    """Same for a narrow OUTPUT side (the 64 -> 3 convolution in front of the ResNet generator's Tanh): Co < 32 is
    zero-padded to 32 (zero weight rows / zero gy channels; the extra output channels are dropped).  Returns 32 or None."""
    if not _tc_mode() or Co >= 32 or tc_covers(kind, N, H, W, Ci, Co, R, S, pad):
        return None
    return 32 if tc_covers(kind, N, H, W, Ci, 32, R, S, pad) else None


def _pad_dim0(t, n):
    out = torch.empty((n,) + tuple(t.shape[1:]), device=t.device, dtype=t.dtype,
                      memory_format=torch.channels_last if t.dim() == 4 else torch.contiguous_format).zero_()
    out[:t.shape[0]].copy_(t)
    return out


def padded_weight_co(w, cop):
    key = (w.data_ptr(), tuple(w.shape), -cop)
    hit = _cache_get(_wp_cache, key, w._version)
    if hit is not None:
        return hit
    if len(_wp_cache) >= 16:
        _wp_cache.clear()
    return _cache_put(_wp_cache, key, w._version, _pad_dim0(w, cop), w)


def _pad_channels(t, cip):
    """[N,C,H,W] channels-last -> [N,cip,H,W] channels-last, zero filled beyond C."""
    out = torch.empty((t.shape[0], cip, t.shape[2], t.shape[3]), device=t.device, dtype=t.dtype,
                      memory_format=torch.channels_last).zero_()
    out[:, :t.shape[1]].copy_(t)
    return out


def padded_weight(w, cip):
    """w [Co,Ci,R,S] zero-padded along Ci; cached per weight version like transposed_weight()."""
    key = (w.data_ptr(), tuple(w.shape), cip)
    hit = _cache_get(_wp_cache, key, w._version)
    if hit is not None:
        return hit
    wp = _pad_channels(w, cip)
    if len(_wp_cache) >= 16:
        _wp_cache.clear()
    return _cache_put(_wp_cache, key, w._version, wp, w)


# ---- pixel-pair packing for the narrow layers (16 / 32 channels at 512^2 / 1024^2, SURVEY.md section 8d "low-intensity layers")
# The tcgen05 kernels move 32-channel (128-byte) rows; a 16-channel NHWC map [N][H][W][16] IS a 32-channel map [N][H][W/2][32]
# in memory (two neighbouring pixels per row).  A stride-1 "same" convolution on the original map equals a convolution of
# the packed map with an expanded weight W'[(po,co)][r][s'][(pi,ci)] = W[co][r][s][ci] where 2*s' + pi = po + s - pad (zero
# elsewhere): same taps count, twice the channels on both sides -- twice the MMA work, which these HBM-bound layers have
# to spare -- and NO copy of any activation.  fprop / dgrad run on the packed views with W'; wgrad produces gW' and folds
# it back.  (Without it these layers ran on the FFMA kernels: 5 ms per wgrad launch at 1024^2.)
_pk_cache = {}


def _pack_ok(kind, N, H, W, Ci, Co, R, S, pad):
    if not _tc_mode() or tc_covers(kind, N, H, W, Ci, Co, R, S, pad):
        return False
    if W % 2 != 0 or R != S or S not in (1, 3) or pad != (S - 1) // 2 or Ci % 4 != 0 or Co % 4 != 0:
        return False
    return tc_covers(kind, N, H, W // 2, 2 * Ci, 2 * Co, R, S, pad)


def _pack_view(t):
    """[N,C,H,W] channels-last -> the same memory as [N,2C,H,W/2] channels-last."""
    n, c, h, w = t.shape
    return t.permute(0, 2, 3, 1).reshape(n, h, w // 2, 2 * c).permute(0, 3, 1, 2)


def _unpack_view(t, c):
    n, c2, h, w2 = t.shape
    return t.permute(0, 2, 3, 1).reshape(n, h, 2 * w2, c).permute(0, 3, 1, 2)


def _pair_taps(S, pad):
    """(po, s) -> (s', pi): tap s of output pixel po of a pair reads pixel pi of the pair at offset s' (in pairs)."""
    out = []
    for po in (0, 1):
        for s_ in range(S):
            t = po + s_ - pad
            out.append((po, s_, t // 2 + pad, t % 2))       # python floor division / modulo: t = -1 -> (-1, 1)
    return out


def packed_weight(w, pad):
    """Expanded weight of the pixel-pair packed convolution, cached per weight version."""
    key = (w.data_ptr(), tuple(w.shape), pad)
    hit = _cache_get(_pk_cache, key, w._version)
    if hit is not None:
        return hit
    Co, Ci, R, S = w.shape
    wp = torch.empty((2 * Co, 2 * Ci, R, S), device=w.device, dtype=w.dtype, memory_format=torch.channels_last).zero_()
    for po, s_, sp, pi in _pair_taps(S, pad):
        wp[po * Co:(po + 1) * Co, pi * Ci:(pi + 1) * Ci, :, sp].copy_(w[:, :, :, s_])
    if len(_pk_cache) >= 64:
        _pk_cache.clear()
    return _cache_put(_pk_cache, key, w._version, wp, w)


def _fold_packed_wgrad(gwp, Co, Ci, S, pad):
    gw = torch.empty((Co, Ci, gwp.shape[2], S), device=gwp.device, dtype=gwp.dtype, memory_format=torch.channels_last).zero_()
    for po, s_, sp, pi in _pair_taps(S, pad):
        gw[:, :, :, s_].add_(gwp[po * Co:(po + 1) * Co, pi * Ci:(pi + 1) * Ci, :, sp])
    return gw


def conv_fprop(x, w, bias, pad, alpha, bias_scale, act, slope):
    _chk(x, w, bias)
    x, w = nhwc(x), nhwc(w)
    N, Ci, H, W = x.shape
    Co, Ci2, R, S = w.shape
    if Ci != Ci2:
        raise GlbError(f"conv_fprop: channel mismatch {Ci} vs {Ci2}")
    if _use_bf16("fprop", N, H, W, Ci, Co, R, S, pad):
        y = _new_nhwc(N, Co, H + 2 * pad - R + 1, W + 2 * pad - S + 1, x)
        _call("glb_conv2d_fprop_bf16", _p(bf16_operand(x)), _p(bf16_weight(w)), _p(_flat(bias)), _p(y), N, H, W, Ci, Co, R, S, pad,
              float(alpha), float(bias_scale), int(act), float(slope), _stream())
        return y
    if _pack_ok("fprop", N, H, W, Ci, Co, R, S, pad):
        b2 = None if bias is None else _flat(bias).repeat(2)
        yp = conv_fprop(_pack_view(x), packed_weight(w, pad), b2, pad, alpha, bias_scale, act, slope)
        return _unpack_view(yp, Co)
    cip = _pad_ci("fprop", N, H, W, Ci, Co, R, S, pad)
    if cip is not None:
        x, w, Ci = _pad_channels(x, cip), padded_weight(w, cip), cip
    cop = _pad_co("fprop", N, H, W, Ci, Co, R, S, pad) if cip is None else None
    co_out = Co
    b = _flat(bias)
    if cop is not None:
        w, Co = padded_weight_co(w, cop), cop
        b = None if b is None else _pad_dim0(b, cop)
    Ho, Wo = H + 2 * pad - R + 1, W + 2 * pad - S + 1
    y = _new_nhwc(N, Co, Ho, Wo, x)
    _call("glb_conv2d_fprop", _p(x), _p(w), _p(b), _p(y), N, H, W, Ci, Co, R, S, pad, float(alpha), float(bias_scale),
          int(act), float(slope), _impl_for("fprop", N, H, W, Ci, Co, R, S, pad), _stream())
    return y if cop is None else y[:, :co_out].contiguous(memory_format=torch.channels_last)


def conv_dgrad(gy, w, x_hw, pad, alpha):
    _chk(gy, w)
    gy, w = nhwc(gy), nhwc(w)
    N, Co, Ho, Wo = gy.shape
    Co2, Ci, R, S = w.shape
    H, W = x_hw
    if Co != Co2 or Ho != H + 2 * pad - R + 1 or Wo != W + 2 * pad - S + 1:
        raise GlbError("conv_dgrad: shape mismatch")
    if _use_bf16("dgrad", N, H, W, Ci, Co, R, S, pad):
        gx = _new_nhwc(N, Ci, H, W, gy)
        _call("glb_conv2d_dgrad_bf16", _p(bf16_operand(gy)), _p(bf16_weight(w, transposed=True)), _p(gx), N, H, W, Ci, Co, R, S, pad,
              float(alpha), _stream())
        return gx
    if _pack_ok("dgrad", N, H, W, Ci, Co, R, S, pad):
        gxp = conv_dgrad(_pack_view(gy), packed_weight(w, pad), (H, W // 2), pad, alpha)
        return _unpack_view(gxp, Ci)
    cip = _pad_ci("dgrad", N, H, W, Ci, Co, R, S, pad)
    ci_out = Ci
    if cip is not None:
        w, Ci = padded_weight(w, cip), cip
    elif _pad_co("dgrad", N, H, W, Ci, Co, R, S, pad) is not None:
        cop = _pad_co("dgrad", N, H, W, Ci, Co, R, S, pad)
        gy, w, Co = _pad_channels(gy, cop), padded_weight_co(w, cop), cop
    gx = _new_nhwc(N, Ci, H, W, gy)
    impl = _impl_for("dgrad", N, H, W, Ci, Co, R, S, pad)
    wt = transposed_weight(w) if impl == IMPL_TF32 else None
    _call("glb_conv2d_dgrad", _p(gy), _p(w), _p(wt), _p(gx), N, H, W, Ci, Co, R, S, pad, float(alpha), impl, _stream())
    return gx if cip is None else gx[:, :ci_out].contiguous(memory_format=torch.channels_last)


def conv_wgrad(x, gy, rs, pad, alpha):
    _chk(x, gy)
    x, gy = nhwc(x), nhwc(gy)
    N, Ci, H, W = x.shape
    N2, Co, Ho, Wo = gy.shape
    R, S = rs
    if N != N2 or Ho != H + 2 * pad - R + 1 or Wo != W + 2 * pad - S + 1:
        raise GlbError("conv_wgrad: shape mismatch")
    if _use_bf16("wgrad", N, H, W, Ci, Co, R, S, pad):
        gw = _new_nhwc(Co, Ci, R, S, x)
        _call("glb_conv2d_wgrad_bf16", _p(bf16_operand(x)), _p(bf16_operand(gy)), _p(gw), N, H, W, Ci, Co, R, S, pad, float(alpha),
              _stream())
        return gw
    if _pack_ok("wgrad", N, H, W, Ci, Co, R, S, pad):
        gwp = conv_wgrad(_pack_view(x), _pack_view(gy), (R, S), pad, alpha)
        return _fold_packed_wgrad(gwp, Co, Ci, S, pad)
    cip = _pad_ci("wgrad", N, H, W, Ci, Co, R, S, pad)
    ci_out, co_out = Ci, Co
    cop = None
    if cip is not None:
        x, Ci = _pad_channels(x, cip), cip
    else:
        cop = _pad_co("wgrad", N, H, W, Ci, Co, R, S, pad)
        if cop is not None:
            gy, Co = _pad_channels(gy, cop), cop
    gw = _new_nhwc(Co, Ci, R, S, x)
    _call("glb_conv2d_wgrad", _p(x), _p(gy), _p(gw), N, H, W, Ci, Co, R, S, pad, float(alpha),
          _impl_for("wgrad", N, H, W, Ci, Co, R, S, pad), _stream())
    if cip is not None:
        return gw[:, :ci_out].contiguous(memory_format=torch.channels_last)
    if cop is not None:
        return gw[:co_out].contiguous(memory_format=torch.channels_last)
    return gw


# ---- nearest-neighbour 2x upsample folded into the 3x3 convolution behind it (glb_upconv_*, TF32 tensor-core path) ---------
# conv3x3(upsample2x(x)) as four 2x2 convolutions of the LOW-resolution x with pre-summed taps: 4/9 of the multiply-adds, no
# upsampled copy.  GLB_UPCONV=0 switches back to the two-kernel sequence (A/B runs).
_up_cache = {}


def upconv_covers(kind: str, N, H, W, Ci, Co) -> bool:
    """H, W: LOW-resolution map.  Tensor-core modes only (the exact FFMA mode keeps the two-kernel sequence); in "bf16" mode a
    layer the bf16 kernels do not cover (channel counts not multiples of 64) falls back to the two-kernel sequence as well."""
    if not _tc_mode() or os.environ.get("GLB_UPCONV", "1") == "0":
        return False
    name = "glb_upconv_bf16_covers" if _state["conv_impl"] == "bf16" else "glb_upconv_covers"
    return bool(LIB.fn(name)(_KIND[kind], N, H, W, Ci, Co))


def _fold16(t):
    """bf16 copy of a re-laid-out weight operand (contiguous), cached with the operand itself"""
    return _to_bf16(t, "fold")


def _folded_weights(w, down, want, both=False):
    """The two operand re-layouts of a folded convolution, cached per weight version: (forward operand, data-gradient operand).
    One launch writes both when the caller knows the other one will be needed (`both`: the forward of a layer whose input
    requires grad), else only the one asked for; a missing one is added on demand."""
    key = (w.data_ptr(), tuple(w.shape), "down" if down else "up")
    hit = _cache_get(_up_cache, key, w._version)
    fwd, bwd = hit if hit is not None else (None, None)
    need_f = fwd is None and (want == "fwd" or both)
    need_b = bwd is None and (want == "bwd" or both)
    if (want == "fwd" and fwd is not None) or (want == "bwd" and bwd is not None):
        return fwd, bwd
    Co, Ci, R, S = w.shape
    if down:        # forward: wt [Co,16,Ci]; data gradient: wp [4*Ci,2,2,Co]
        if need_f:
            fwd = torch.empty((Co, 16, Ci), device=w.device, dtype=torch.float32)
        if need_b:
            bwd = torch.empty((4 * Ci, 2, 2, Co), device=w.device, dtype=torch.float32)
        _call("glb_downconv_weights", _p(w), _p(bwd if need_b else None), _p(fwd if need_f else None), Co, Ci, _stream())
    else:           # forward: wp [4*Co,2,2,Ci]; data gradient: wt [Ci,16,Co]
        if need_f:
            fwd = torch.empty((4 * Co, 2, 2, Ci), device=w.device, dtype=torch.float32)
        if need_b:
            bwd = torch.empty((Ci, 16, Co), device=w.device, dtype=torch.float32)
        _call("glb_upconv_weights", _p(w), _p(fwd if need_f else None), _p(bwd if need_b else None), Co, Ci, _stream())
    if len(_up_cache) >= 64:
        _up_cache.clear()
    _cache_put(_up_cache, key, w._version, (fwd, bwd), w)
    return fwd, bwd


def upconv_weights(w, want="fwd", both=False):
    """(wp [4*Co,2,2,Ci] for the forward, wt [Ci,16,Co] for the data gradient) of w [Co,Ci,3,3]."""
    return _folded_weights(w, False, want, both)


def upconv_fprop(x, w, bias, alpha, bias_scale, act, slope, need_dgrad=False):
    """need_dgrad: the data-gradient operand will be needed as well (written by the same re-layout launch)"""
    _chk(x, w, bias)
    x, w = nhwc(x), nhwc(w)
    N, Ci, H, W = x.shape
    Co, Ci2, R, S = w.shape
    if Ci != Ci2 or R != 3 or S != 3:
        raise GlbError("upconv_fprop: needs a 3x3 weight with matching channels")
    wp, _ = upconv_weights(w, "fwd", need_dgrad)
    y = _new_nhwc(N, Co, 2 * H, 2 * W, x)
    if _state["conv_impl"] == "bf16":
        _call("glb_upconv_fprop_bf16", _p(bf16_operand(x)), _p(_fold16(wp)), _p(_flat(bias)), _p(y), N, H, W, Ci, Co, float(alpha),
              float(bias_scale), int(act), float(slope), _stream())
        return y
    _call("glb_upconv_fprop", _p(x), _p(wp), _p(_flat(bias)), _p(y), N, H, W, Ci, Co, float(alpha), float(bias_scale), int(act),
          float(slope), _stream())
    return y


def upconv_dgrad(gy, w, alpha):
    _chk(gy, w)
    gy, w = nhwc(gy), nhwc(w)
    N, Co, H2, W2 = gy.shape
    Co2, Ci, R, S = w.shape
    if Co != Co2 or H2 % 2 or W2 % 2:
        raise GlbError("upconv_dgrad: shape mismatch")
    _, wt = upconv_weights(w, "bwd")
    gx = _new_nhwc(N, Ci, H2 // 2, W2 // 2, gy)
    if _state["conv_impl"] == "bf16":
        _call("glb_upconv_dgrad_bf16", _p(bf16_operand(gy)), _p(_fold16(wt)), _p(gx), N, H2 // 2, W2 // 2, Ci, Co, float(alpha), _stream())
        return gx
    _call("glb_upconv_dgrad", _p(gy), _p(wt), _p(gx), N, H2 // 2, W2 // 2, Ci, Co, float(alpha), _stream())
    return gx


def upconv_wgrad(x, gy, alpha):
    _chk(x, gy)
    x, gy = nhwc(x), nhwc(gy)
    N, Ci, H, W = x.shape
    N2, Co, H2, W2 = gy.shape
    if N != N2 or H2 != 2 * H or W2 != 2 * W:
        raise GlbError("upconv_wgrad: shape mismatch")
    gwp = torch.empty((Co, 16, Ci), device=x.device, dtype=torch.float32)
    gw = _new_nhwc(Co, Ci, 3, 3, x)
    if _state["conv_impl"] == "bf16":
        _call("glb_upconv_wgrad_bf16", _p(bf16_operand(x)), _p(bf16_operand(gy)), _p(gwp), _p(gw), N, H, W, Ci, Co, float(alpha), _stream())
        return gw
    _call("glb_upconv_wgrad", _p(x), _p(gy), _p(gwp), _p(gw), N, H, W, Ci, Co, float(alpha), _stream())
    return gw


# ---- 2x2 average pool folded into the 3x3 convolution in front of it (glb_downconv_*: the adjoint structure of upconv) ------
def downconv_covers(N, H, W, Ci, Co) -> bool:
    """H, W: LOW-resolution (output) map.  All three kernels must cover the pair (the family is closed under differentiation:
    the R1 double backward runs fprop / dgrad / wgrad of the same layer)."""
    if not _tc_mode() or os.environ.get("GLB_DOWNCONV", "1") == "0":
        return False
    f = LIB.fn("glb_downconv_bf16_covers" if _state["conv_impl"] == "bf16" else "glb_downconv_covers")
    return all(bool(f(k, N, H, W, Ci, Co)) for k in (0, 1, 2))


def downconv_weights(w, want="fwd", both=False):
    """(wt [Co,16,Ci] for the forward, wp [4*Ci,2,2,Co] for the data gradient) of w [Co,Ci,3,3]."""
    return _folded_weights(w, True, want, both)


def downconv_fprop(x, w, bias, alpha, bias_scale, act, slope, need_dgrad=False):
    """act(alpha * avgpool2x2(conv3x3_same(x, w)) + bias_scale * bias) -> [N, Co, H/2, W/2]"""
    _chk(x, w, bias)
    x, w = nhwc(x), nhwc(w)
    N, Ci, H2, W2 = x.shape
    Co, Ci2, R, S = w.shape
    if Ci != Ci2 or R != 3 or S != 3 or H2 % 2 or W2 % 2:
        raise GlbError("downconv_fprop: needs a 3x3 weight with matching channels and an even map")
    wt, _ = downconv_weights(w, "fwd", need_dgrad)
    y = _new_nhwc(N, Co, H2 // 2, W2 // 2, x)
    if _state["conv_impl"] == "bf16":
        _call("glb_downconv_fprop_bf16", _p(bf16_operand(x)), _p(_fold16(wt)), _p(_flat(bias)), _p(y), N, H2 // 2, W2 // 2, Ci, Co,
              float(alpha), float(bias_scale), int(act), float(slope), _stream())
        return y
    _call("glb_downconv_fprop", _p(x), _p(wt), _p(_flat(bias)), _p(y), N, H2 // 2, W2 // 2, Ci, Co, float(alpha), float(bias_scale),
          int(act), float(slope), _stream())
    return y


def downconv_dgrad(gy, w, alpha):
    _chk(gy, w)
    gy, w = nhwc(gy), nhwc(w)
    N, Co, H, W = gy.shape
    Co2, Ci, R, S = w.shape
    if Co != Co2:
        raise GlbError("downconv_dgrad: shape mismatch")
    _, wp = downconv_weights(w, "bwd")
    gx = _new_nhwc(N, Ci, 2 * H, 2 * W, gy)
    if _state["conv_impl"] == "bf16":
        _call("glb_downconv_dgrad_bf16", _p(bf16_operand(gy)), _p(_fold16(wp)), _p(gx), N, H, W, Ci, Co, float(alpha), _stream())
        return gx
    _call("glb_downconv_dgrad", _p(gy), _p(wp), _p(gx), N, H, W, Ci, Co, float(alpha), _stream())
    return gx


def downconv_wgrad(x, gy, alpha):
    _chk(x, gy)
    x, gy = nhwc(x), nhwc(gy)
    N, Ci, H2, W2 = x.shape
    N2, Co, H, W = gy.shape
    if N != N2 or H2 != 2 * H or W2 != 2 * W:
        raise GlbError("downconv_wgrad: shape mismatch")
    gwp = torch.empty((Ci, 16, Co), device=x.device, dtype=torch.float32)
    gw = _new_nhwc(Co, Ci, 3, 3, x)
    if _state["conv_impl"] == "bf16":
        _call("glb_downconv_wgrad_bf16", _p(bf16_operand(x)), _p(bf16_operand(gy)), _p(gwp), _p(gw), N, H, W, Ci, Co, float(alpha), _stream())
        return gw
    _call("glb_downconv_wgrad", _p(x), _p(gy), _p(gwp), _p(gw), N, H, W, Ci, Co, float(alpha), _stream())
    return gw


# --------------------------------------------------------------------------- linear
def linear_fwd(x, w, bias, alpha, bias_scale, act, slope):
    _chk(x, w, bias)
    x, w = x.contiguous(), w.contiguous()
    M, K = x.shape
    Nout, K2 = w.shape
    if K != K2:
        raise GlbError(f"linear_fwd: feature mismatch {K} vs {K2}")
    y = torch.empty((M, Nout), device=x.device, dtype=torch.float32)
    _call("glb_linear_fwd", _p(x), _p(w), _p(_flat(bias)), _p(y), M, K, Nout, float(alpha), float(bias_scale), int(act),
          float(slope), _stream())
    return y


def linear_dgrad(gy, w, alpha):
    _chk(gy, w)
    gy, w = gy.contiguous(), w.contiguous()
    M, Nout = gy.shape
    K = w.shape[1]
    gx = torch.empty((M, K), device=gy.device, dtype=torch.float32)
    _call("glb_linear_dgrad", _p(gy), _p(w), _p(gx), M, K, Nout, float(alpha), _stream())
    return gx


def linear_wgrad(x, gy, alpha):
    _chk(x, gy)
    x, gy = x.contiguous(), gy.contiguous()
    M, K = x.shape
    Nout = gy.shape[1]
    gw = torch.empty((Nout, K), device=x.device, dtype=torch.float32)
    _call("glb_linear_wgrad", _p(x), _p(gy), _p(gw), M, K, Nout, float(alpha), _stream())
    return gw


# --------------------------------------------------------------------------- grouped small linears
class GroupedLinearTable(object):
    """Device table for glb_glinear_* over a fixed list of (weight [nout,K], bias [nout] or None, alpha, bias_scale) layers.
    Rebuilt (one small H2D copy, never inside a CUDA-graph capture) only when a parameter's address changes."""

    def __init__(self):
        self.key = None
        self.dev = None
        self.nouts, self.offs, self.G = [], [], 0
        self.blocks = (0, 0, 0)

    def update(self, layers, device):
        import struct
        key = tuple((w.data_ptr(), 0 if b is None else b.data_ptr(), float(a), float(bs)) for w, b, a, bs in layers)
        if key == self.key and self.dev is not None and self.dev.device == device:
            return self
        if torch.cuda.is_available() and torch.cuda.is_current_stream_capturing():
            raise GlbError("GroupedLinearTable must be built before a CUDA-graph capture (run one eager iteration first)")
        rows = [LIB.fn("glb_glinear_rows")(i) for i in range(3)]
        buf = bytearray()
        off = 0
        blk = [0, 0, 0]
        self.nouts, self.offs = [], []
        for w, b, a, bs in layers:
            nout = w.shape[0]
            buf += struct.pack("<qqqqqqqff", w.data_ptr(), 0 if b is None else b.data_ptr(), nout, off, blk[0], blk[1], blk[2],
                               float(a), float(bs))
            self.nouts.append(nout); self.offs.append(off)
            off += nout
            for i in range(3):
                blk[i] += (nout + rows[i] - 1) // rows[i]
        self.G = off
        self.blocks = tuple(blk)
        self.dev = torch.frombuffer(buf, dtype=torch.int64).clone().to(device)
        self.key = key
        return self


def glinear_fwd(ws, tab):
    """ws [L,M,K] -> flat buffer y; layer l's output is y[M*off_l : M*(off_l+nout_l)].view(M, nout_l)."""
    _chk(ws)
    ws = ws.contiguous()
    L, M, K = ws.shape
    y = torch.empty(M * tab.G, device=ws.device, dtype=torch.float32)
    _call("glb_glinear_fwd", _p(ws), _p(tab.dev), _p(y), L, M, K, tab.blocks[0], _stream())
    return y


def glinear_dgrad(g_all, tab, L, M, K):
    _chk(g_all)
    g_all = g_all.contiguous()
    g_ws = torch.empty((L, M, K), device=g_all.device, dtype=torch.float32)
    _call("glb_glinear_dgrad", _p(g_all), tab.G, _p(tab.dev), _p(g_ws), L, M, K, tab.blocks[1], _stream())
    return g_ws


def glinear_wgrad(ws, g_all, tab):
    _chk(ws, g_all)
    ws, g_all = ws.contiguous(), g_all.contiguous()
    L, M, K = ws.shape
    gw = torch.empty((tab.G, K), device=ws.device, dtype=torch.float32)
    gb = torch.empty(tab.G, device=ws.device, dtype=torch.float32)
    _call("glb_glinear_wgrad", _p(ws), _p(g_all), tab.G, _p(tab.dev), _p(gw), _p(gb), L, M, K, tab.blocks[2], _stream())
    return gw, gb


# --------------------------------------------------------------------------- rows helpers
def _rows(x):
    """-> (tensor in kernel layout, P, C): 4-D channels_last or 2-D row-major, channel axis contiguous."""
    if x.dim() == 4:
        x = nhwc(x)
        n, c, h, w = x.shape
        return x, n * h * w, c
    x = x.contiguous()
    return x, x.shape[0], x.shape[1]


def bias_act_fwd(x, bias, bias_scale, act, slope):
    _chk(x, bias)
    x, P, C = _rows(x)
    y = torch.empty_like(x)
    _call("glb_bias_act_fwd", _p(x), _p(_flat(bias)), _p(y), P, C, float(bias_scale), int(act), float(slope), _stream())
    return y


def act_bwd(gy, y, want_bias, bias_scale, act, slope):
    """-> (gx, gbias or None); gx = gy * act'(y)."""
    _chk(gy, y)
    y, P, C = _rows(y)
    gy, _, _ = _rows(gy)
    gb = _zeros(C, gy) if want_bias else None
    if C % 4 == 0 and ((C // 4) > 256 or 256 % (C // 4) == 0):
        gx = torch.empty_like(gy)
        _call("glb_act_bwd", _p(gy), _p(y), _p(gx), _p(gb), P, C, float(bias_scale), int(act), float(slope), _stream())
        return gx, gb
    raise GlbError(f"act_bwd: unsupported channel count {C}")


def colsum(x, scale):
    _chk(x)
    x, P, C = _rows(x)
    out = _zeros(C, x)
    _call("glb_colsum", _p(x), _p(out), P, C, float(scale), _stream())
    return out


class DeviceAlpha(object):
    """The fade-in alpha kept ON THE DEVICE as `coef` = [alpha, 1 - alpha] (fp32), so that the blends of a fade-in phase -- whose
    alpha moves every iteration (reference progan/learner.py:951-952) -- can live in a captured CUDA graph.  `set()` is called
    by the learner before every iteration (outside any capture); `.alpha` / `.one_minus` are the handles the launchers accept
    wherever they accept a float coefficient."""

    def __init__(self, device):
        self.coef = torch.zeros(2, device=device, dtype=torch.float32)
        self.value = None
        self.alpha, self.one_minus = Coef(self, 0), Coef(self, 1)

    def set(self, alpha):
        alpha = float(alpha)
        if alpha != self.value:
            self.coef[0:1].fill_(alpha)
            self.coef[1:2].fill_(1. - alpha)
            self.value = alpha


class Coef(object):
    """One entry of a DeviceAlpha vector; float(c) is the host-side value last set."""

    def __init__(self, ref, idx):
        self.ref, self.idx = ref, idx

    def __float__(self):
        v = self.ref.value
        return v if self.idx == 0 else 1. - v


def coef_or_float(v):
    return v if isinstance(v, Coef) else float(v)


def axpby(a, b, alpha, beta):
    _chk(a, b)
    if a.dim() == 4 and a.shape[1] > 3:
        a = nhwc(a)
        b = nhwc(b) if b is not None else None
    else:
        a = a.contiguous()
        b = b.contiguous() if b is not None else None
    y = torch.empty_like(a)
    if isinstance(alpha, Coef):
        ib = alpha.idx
        if b is not None:
            if not (isinstance(beta, Coef) and beta.ref is alpha.ref):
                raise GlbError("axpby: both coefficients must come from the same DeviceAlpha")
            ib = beta.idx
        _call("glb_axpby_dev", _p(a), _p(b), _p(y), a.numel(), alpha.ref.coef.data_ptr(), alpha.idx, ib, _stream())
        return y
    _call("glb_axpby", _p(a), _p(b), _p(y), a.numel(), float(alpha), float(beta), _stream())
    return y


def scale_by(x, s_dev, scale):
    """y = x * scale * s_dev (device scalar tensor)."""
    _chk(x, s_dev)
    if not (x.is_contiguous() or (x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last))):
        x = x.contiguous()
    y = torch.empty_like(x)
    _call("glb_scale_by", _p(x), _p(s_dev.reshape(-1)), _p(y), x.numel(), float(scale), _stream())
    return y


def sumsq(x, scale):
    _chk(x)
    if not (x.is_contiguous() or (x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last))):
        x = x.contiguous()
    out = _zeros((), x)
    _call("glb_sumsq", _p(x), _p(out), x.numel(), float(scale), _stream())
    return out


def gp_norm_fwd(g, gamma, scale):
    _chk(g)
    g = g.contiguous()
    N, C = g.shape[0], g.shape[1]
    out = _zeros((), g)
    _call("glb_gp_norm_fwd", _p(g), _p(out), N, C, g.numel() // (N * C), float(gamma), float(scale), _stream())
    return out


def gp_norm_bwd(g, s_dev, gamma, scale):
    _chk(g, s_dev)
    g = g.contiguous()
    N, C = g.shape[0], g.shape[1]
    gg = torch.empty_like(g)
    _call("glb_gp_norm_bwd", _p(g), _p(s_dev.reshape(-1)), _p(gg), N, C, g.numel() // (N * C), float(gamma), float(scale), _stream())
    return gg


def interp_rows(a, b, eps):
    _chk(a, b, eps)
    a, b, eps = a.contiguous(), b.contiguous(), eps.reshape(-1).contiguous()
    out = torch.empty_like(a)
    _call("glb_interp_rows", _p(a), _p(b), _p(eps), _p(out), a.shape[0], a.numel() // a.shape[0], _stream())
    return out


# --------------------------------------------------------------------------- norms / stencils / resampling
def pixelnorm_fwd(x, eps):
    _chk(x)
    x, P, C = _rows(x)
    y = torch.empty_like(x)
    _call("glb_pixelnorm_fwd", _p(x), _p(y), P, C, float(eps), _stream())
    return y


def pixelnorm_bwd(gy, x, eps):
    _chk(gy, x)
    x, P, C = _rows(x)
    gy, _, _ = _rows(gy)
    gx = torch.empty_like(x)
    _call("glb_pixelnorm_bwd", _p(gy), _p(x), _p(gx), P, C, float(eps), _stream())
    return gx


def blur3x3(x):
    _chk(x)
    x = nhwc(x)
    N, C, H, W = x.shape
    y = torch.empty_like(x)
    _call("glb_blur3x3", _p(x), _p(y), N, H, W, C, _stream())
    return y


def blur_act_bwd(gz, y, want_bias, bias_scale, act, slope):
    """g = blur3x3(gz) * act'(y) (+ bias gradient): backward of `blur(act(conv + b))` in one pass -> (g, gbias or None)."""
    _chk(gz, y)
    gz, y = nhwc(gz), nhwc(y)
    N, C, H, W = y.shape
    if C % 4 != 0 or 256 % (C // 4) != 0:
        g0 = blur3x3(gz)
        return act_bwd(g0, y, want_bias, bias_scale, act, slope)
    g = torch.empty_like(y)
    gb = _zeros(C, y) if want_bias else None
    _call("glb_blur_act_bwd", _p(gz), _p(y), _p(g), _p(gb), N, H, W, C, float(bias_scale), int(act), float(slope), _stream())
    return g, gb


def upsample2x_fwd(x):
    _chk(x)
    x = nhwc(x)
    N, C, H, W = x.shape
    y = _new_nhwc(N, C, 2 * H, 2 * W, x)
    _call("glb_upsample2x_fwd", _p(x), _p(y), N, H, W, C, _stream())
    return y


def upsample2x_bwd(gy):
    _chk(gy)
    gy = nhwc(gy)
    N, C, H2, W2 = gy.shape
    gx = _new_nhwc(N, C, H2 // 2, W2 // 2, gy)
    _call("glb_upsample2x_bwd", _p(gy), _p(gx), N, H2 // 2, W2 // 2, C, _stream())
    return gx


def pool_bias_act_fwd(x, bias, bias_scale, act, slope):
    _chk(x, bias)
    x = nhwc(x)
    N, C, H, W = x.shape
    y = _new_nhwc(N, C, H // 2, W // 2, x)
    _call("glb_pool_bias_act_fwd", _p(x), _p(_flat(bias)), _p(y), N, H, W, C, float(bias_scale), int(act), float(slope), _stream())
    return y


def pool_bias_act_bwd(gy, y, want_bias, bias_scale, act, slope):
    """gy [N,C,H/2,W/2] (+ saved output y or None) -> gx [N,C,H,W] = 0.25 * gy*act'(y) broadcast, gbias."""
    _chk(gy, y)
    gy = nhwc(gy)
    y = nhwc(y) if y is not None else None
    N, C, Hh, Wh = gy.shape
    gx = _new_nhwc(N, C, 2 * Hh, 2 * Wh, gy)
    gb = _zeros(C, gy) if want_bias else None
    _call("glb_pool_bias_act_bwd", _p(gy), _p(y), _p(gx), _p(gb), N, 2 * Hh, 2 * Wh, C, float(bias_scale), int(act), float(slope),
          _stream())
    return gx, gb


# --------------------------------------------------------------------------- StyleGAN G-layer epilogue
def style_epilogue_fwd(x, noise, noise_weight, bias, style, slope, eps):
    _chk(x, noise, noise_weight, bias, style)
    x = nhwc(x)
    N, C, H, W = x.shape
    noise_c = noise.contiguous() if noise is not None else None
    style = style.contiguous()
    out = torch.empty_like(x)
    stats = torch.empty(N * 2 * C, device=x.device, dtype=torch.float32)
    work = torch.empty(LIB.fn("glb_style_epilogue_work_floats")(N, H, W, C), device=x.device, dtype=torch.float32)
    _call("glb_style_epilogue_fwd", _p(x), _p(noise_c), _p(_flat(noise_weight)), _p(_flat(bias)), _p(style), _p(out), _p(stats),
          _p(work), N, H, W, C, float(slope), float(eps), _stream())
    return out, stats


def style_epilogue_bwd(gout, x, noise, noise_weight, bias, style, stats, slope):
    _chk(gout, x, noise, noise_weight, bias, style, stats)
    x, gout = nhwc(x), nhwc(gout)
    N, C, H, W = x.shape
    noise_c = noise.contiguous() if noise is not None else None
    style = style.contiguous()
    gx = torch.empty_like(x)
    gstyle = torch.empty_like(style)
    g_nw = _zeros(C, x) if noise_weight is not None else None
    g_b = _zeros(C, x) if bias is not None else None
    work = torch.empty(LIB.fn("glb_style_epilogue_work_floats")(N, H, W, C), device=x.device, dtype=torch.float32)
    _call("glb_style_epilogue_bwd", _p(gout), _p(x), _p(noise_c), _p(_flat(noise_weight)), _p(_flat(bias)), _p(style), _p(stats),
          _p(gx), _p(gstyle), _p(g_nw), _p(g_b), _p(work), N, H, W, C, float(slope), _stream())
    return gx, gstyle, g_nw, g_b


# --------------------------------------------------------------------------- ResNet-GAN norms (csrc/norm.cu)
def _hwc(t):
    """(C,H,W)-shaped affine map -> its (H,W,C)-ordered memory (zero-copy when the parameter is stored that way)."""
    return None if t is None else t.detach().permute(1, 2, 0).contiguous()


def _new_chw_as_hwc(C, H, W, like):
    return torch.empty((H, W, C), device=like.device, dtype=torch.float32).permute(2, 0, 1)


def _ws(n_doubles, like):
    return torch.empty(int(n_doubles), device=like.device, dtype=torch.float64)


def layernorm_fwd(x, gamma, beta, eps, act, slope):
    """x [N,C,H,W] (NHWC memory), gamma/beta [C,H,W] -> (y, stats[N,2] = mean, rstd)."""
    _chk(x, gamma, beta)
    x = nhwc(x)
    N, C, H, W = x.shape
    y = torch.empty_like(x)
    stats = torch.empty((N, 2), device=x.device, dtype=torch.float32)
    ws = _ws(LIB.fn("glb_layernorm_ws_doubles")(N), x)
    g_hwc, b_hwc = _hwc(gamma), _hwc(beta)          # keep the (possibly copied) maps alive across the launch
    _call("glb_layernorm_fwd", _p(x), _p(g_hwc), _p(b_hwc), _p(y), _p(stats), _p(ws), N, C * H * W, float(eps),
          int(act), float(slope), _stream())
    return y, stats


def layernorm_bwd(gy, y, x, gamma, stats, act, slope, want_gx=True, want_params=True):
    """-> (gx, ggamma [C,H,W], gbeta [C,H,W]); entries not asked for are None."""
    _chk(gy, y, x, gamma, stats)
    gy, y, x = nhwc(gy), nhwc(y), nhwc(x)
    N, C, H, W = x.shape
    gx = torch.empty_like(x) if want_gx else None
    gg = _new_chw_as_hwc(C, H, W, x) if want_params else None
    gb = _new_chw_as_hwc(C, H, W, x) if want_params else None
    ws = _ws(LIB.fn("glb_layernorm_ws_doubles")(N), x)
    g_hwc = _hwc(gamma)
    _call("glb_layernorm_bwd", _p(gy), _p(y), _p(x), _p(g_hwc), _p(stats), _p(gx), _p(gg), _p(gb), _p(ws), N, C * H * W,
          int(act), float(slope), _stream())
    return gx, gg, gb


def layernorm_bwdbwd(u, gy, y, x, gamma, stats, act, slope, want_gy=True, want_x=True, want_gamma=True):
    """Second order: cotangent u of gx -> (g_gy, g_x, g_gamma [C,H,W])."""
    _chk(u, gy, y, x, gamma, stats)
    u, gy, y, x = nhwc(u), nhwc(gy), nhwc(y), nhwc(x)
    N, C, H, W = x.shape
    g_gy = torch.empty_like(x) if want_gy else None
    g_x = torch.empty_like(x) if want_x else None
    g_gamma = _new_chw_as_hwc(C, H, W, x) if want_gamma else None
    ws = _ws(LIB.fn("glb_layernorm_ws_doubles")(N), x)
    g_hwc = _hwc(gamma)
    _call("glb_layernorm_bwdbwd", _p(u), _p(gy), _p(y), _p(x), _p(g_hwc), _p(stats), _p(g_gy), _p(g_x), _p(g_gamma), _p(ws),
          N, C * H * W, int(act), float(slope), _stream())
    return g_gy, g_x, g_gamma


def batchnorm_fwd(x, gamma, beta, running_mean, running_var, num_batches_tracked, eps, momentum, act, slope):
    """Training-mode BatchNorm2d (+act) on x [N,C,H,W]; updates the running buffers in place -> (y, stats[2,C])."""
    _chk(x, gamma, beta, running_mean, running_var)
    x, P, C = _rows(x)
    y = torch.empty_like(x)
    stats = torch.empty((2, C), device=x.device, dtype=torch.float32)
    ws = _ws(LIB.fn("glb_batchnorm_ws_doubles")(C), x)
    if num_batches_tracked is not None and num_batches_tracked.dtype != torch.int64:
        raise GlbError("num_batches_tracked must be int64")
    _call("glb_batchnorm_fwd", _p(x), _p(_flat(gamma)), _p(_flat(beta)), _p(y), _p(stats), _p(running_mean), _p(running_var),
          _p(num_batches_tracked), _p(ws), P, C, float(eps), float(momentum), int(act), float(slope), _stream())
    return y, stats


def batchnorm_bwd(gy, y, x, gamma, stats, act, slope):
    """-> (gx, ggamma [C], gbeta [C])."""
    _chk(gy, y, x, gamma, stats)
    x, P, C = _rows(x)
    gy, _, _ = _rows(gy)
    y, _, _ = _rows(y)
    gx = torch.empty_like(x)
    gg = torch.empty(C, device=x.device, dtype=torch.float32)
    gb = torch.empty(C, device=x.device, dtype=torch.float32)
    ws = _ws(LIB.fn("glb_batchnorm_ws_doubles")(C), x)
    _call("glb_batchnorm_bwd", _p(gy), _p(y), _p(x), _p(_flat(gamma)), _p(stats), _p(gx), _p(gg), _p(gb), _p(ws), P, C, int(act),
          float(slope), _stream())
    return gx, gg, gb


def tanh_fwd(x):
    _chk(x)
    x = x if x.is_contiguous() or (x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last)) else x.contiguous()
    y = torch.empty_like(x)
    _call("glb_tanh_fwd", _p(x), _p(y), x.numel(), _stream())
    return y


def tanh_bwd(gy, y):
    _chk(gy, y)
    if gy.stride() != y.stride():
        gy = gy.contiguous(memory_format=torch.channels_last) if (y.dim() == 4 and not y.is_contiguous()) else gy.contiguous()
        if gy.stride() != y.stride():
            y = y.contiguous(); gy = gy.contiguous()
    gx = torch.empty_like(y)
    _call("glb_tanh_bwd", _p(gy), _p(y), _p(gx), y.numel(), _stream())
    return gx


# --------------------------------------------------------------------------- minibatch stddev
def mbstd_fwd(x, group):
    _chk(x)
    x = nhwc(x)
    N, C, H, W = x.shape
    y = _new_nhwc(N, C + 1, H, W, x)
    _call("glb_mbstd_fwd", _p(x), _p(y), N, H, W, C, group, _stream())
    return y


def mbstd_bwd(gy, x, group):
    _chk(gy, x)
    x, gy = nhwc(x), nhwc(gy)
    N, C, H, W = x.shape
    gx = torch.empty_like(x)
    _call("glb_mbstd_bwd", _p(gy), _p(x), _p(gx), N, H, W, C, group, _stream())
    return gx


def mbstd_bwdbwd(v, gy, x, group):
    _chk(v, gy, x)
    x, gy, v = nhwc(x), nhwc(gy), nhwc(v)
    N, C, H, W = x.shape
    ggx = torch.empty_like(x)
    ggy = torch.empty_like(gy)
    _call("glb_mbstd_bwdbwd", _p(v), _p(gy), _p(x), _p(ggx), _p(ggy), N, H, W, C, group, _stream())
    return ggx, ggy


# --------------------------------------------------------------------------- RGB 1x1 convs
def rgb_expand(img, w, ws_j, ws_c, C, bias, pool, alpha, bias_scale, act, slope):
    """img NCHW [N,3,IH,IW] -> NHWC features [N,C,IH/(1+pool),IW/(1+pool)]."""
    _chk(img, w, bias)
    img = img.contiguous()
    w = w.contiguous()
    N, three, IH, IW = img.shape
    if three != 3:
        raise GlbError("rgb_expand: image must have 3 channels")
    H, W = (IH // 2, IW // 2) if pool else (IH, IW)
    y = _new_nhwc(N, C, H, W, img)
    _call("glb_rgb_expand", _p(img), _p(w), ws_j, ws_c, _p(_flat(bias)), _p(y), N, H, W, C, int(pool), float(alpha),
          float(bias_scale), int(act), float(slope), _stream())
    return y


def rgb_contract(x, w, ws_j, ws_c, bias, pool, alpha, bias_scale):
    """NHWC features [N,C,H,W] -> img NCHW [N,3,H*(1+pool),W*(1+pool)]."""
    _chk(x, w, bias)
    x = nhwc(x)
    w = w.contiguous()
    N, C, H, W = x.shape
    s = 2 if pool else 1
    img = torch.empty((N, 3, H * s, W * s), device=x.device, dtype=torch.float32)
    _call("glb_rgb_contract", _p(x), _p(w), ws_j, ws_c, _p(_flat(bias)), _p(img), N, H, W, C, int(pool), float(alpha),
          float(bias_scale), _stream())
    return img


def rgb_wgrad(img, g, w_shape, ws_j, ws_c, pool, alpha):
    _chk(img, g)
    img = img.contiguous()
    g = nhwc(g)
    N, C, H, W = g.shape
    gw = _zeros(tuple(w_shape), g)
    _call("glb_rgb_wgrad", _p(img), _p(g), _p(gw), ws_j, ws_c, N, H, W, C, int(pool), float(alpha), _stream())
    return gw


def plane_sum(img, scale):
    _chk(img)
    img = img.contiguous()
    N, C = img.shape[0], img.shape[1]
    out = _zeros(C, img)
    _call("glb_plane_sum", _p(img), _p(out), N, C, img.numel() // (N * C), float(scale), _stream())
    return out


# --------------------------------------------------------------------------- image-space fade-in
def fade_up_blend(lo, hi, alpha):
    _chk(lo, hi)
    lo, hi = lo.contiguous(), hi.contiguous()
    N, C, H, W = hi.shape
    out = torch.empty_like(hi)
    if isinstance(alpha, Coef):
        assert alpha.idx == 0
        _call("glb_fade_up_blend_dev", _p(lo), _p(hi), _p(out), N, C, H, W, alpha.ref.coef.data_ptr(), _stream())
        return out
    _call("glb_fade_up_blend", _p(lo), _p(hi), _p(out), N, C, H, W, float(alpha), _stream())
    return out


def fade_up_blend_bwd(gout, alpha):
    _chk(gout)
    gout = gout.contiguous()
    N, C, H, W = gout.shape
    glo = torch.empty((N, C, H // 2, W // 2), device=gout.device, dtype=torch.float32)
    ghi = torch.empty_like(gout)
    if isinstance(alpha, Coef):
        assert alpha.idx == 0
        _call("glb_fade_up_blend_bwd_dev", _p(gout), _p(glo), _p(ghi), N, C, H, W, alpha.ref.coef.data_ptr(), _stream())
        return glo, ghi
    _call("glb_fade_up_blend_bwd", _p(gout), _p(glo), _p(ghi), N, C, H, W, float(alpha), _stream())
    return glo, ghi


def fade_real(x, alpha):
    _chk(x)
    x = x.contiguous()
    N, C, H, W = x.shape
    out = torch.empty_like(x)
    if isinstance(alpha, Coef):
        assert alpha.idx == 0
        _call("glb_fade_real_dev", _p(x), _p(out), N, C, H, W, alpha.ref.coef.data_ptr(), _stream())
        return out
    _call("glb_fade_real", _p(x), _p(out), N, C, H, W, float(alpha), _stream())
    return out


# --------------------------------------------------------------------------- losses / misc
LOSS_KIND = {"wgan": 0, "nonsaturating": 1, "minimax": 2}


def d_logit_loss(d_gen, d_real, kind, eps_drift):
    _chk(d_gen, d_real)
    d_gen, d_real = d_gen.contiguous(), d_real.contiguous()
    n = d_gen.numel()
    loss = torch.empty((), device=d_gen.device, dtype=torch.float32)
    gg, gr = torch.empty_like(d_gen), torch.empty_like(d_real)
    _call("glb_d_logit_loss", _p(d_gen), _p(d_real), _p(loss), _p(gg), _p(gr), n, LOSS_KIND[kind], float(eps_drift), _stream())
    return loss, gg, gr


def g_logit_loss(d_out, kind):
    _chk(d_out)
    d_out = d_out.contiguous()
    loss = torch.empty((), device=d_out.device, dtype=torch.float32)
    g = torch.empty_like(d_out)
    _call("glb_g_logit_loss", _p(d_out), _p(loss), _p(g), d_out.numel(), LOSS_KIND[kind], _stream())
    return loss, g


def w_ewma_update(w, ewma, beta):
    """In place on `ewma` (beta == 0 initialises it with the batch mean)."""
    _chk(w, ewma)
    w = w.contiguous()
    _call("glb_w_ewma", _p(w), _p(ewma), w.shape[0], w.shape[1], float(beta), _stream())
    return ewma


def adam_hyper_advance(hyper, beta1, beta2):
    _call("glb_adam_hyper_advance", _p(hyper), float(beta1), float(beta2), _stream())


def adam_ewma_multi(ptr_table, sizes, T, max_size, hyper, beta1, beta2, eps, wd, ewma_beta, ewma_mode):
    _call("glb_adam_ewma_multi", _p(ptr_table), _p(sizes), T, max_size, _p(hyper), float(beta1), float(beta2), float(eps),
          float(wd), float(ewma_beta), int(ewma_mode), _stream())


# --------------------------------------------------------------------------- real-image input pipeline
_box_tables_cache = {}


def box_resize_tables(in_size: int, out_size: int):
    """Pillow's BOX-filter coefficient tables for one axis, computed by the library's host helper (double arithmetic as in
    Pillow's Resample.c) -> (bounds int32 [out, 2], kk int32 [out, ksize]) as CPU tensors."""
    import ctypes
    ks = LIB.fn("glb_box_resize_ksize")(int(in_size), int(out_size))
    if ks <= 0:
        raise GlbError(f"glb_box_resize_ksize({in_size}, {out_size}) failed")
    bounds = torch.empty((out_size, 2), dtype=torch.int32)
    kk = torch.empty((out_size, ks), dtype=torch.int32)
    LIB.call("glb_box_resize_tables", int(in_size), int(out_size), ctypes.c_void_p(bounds.data_ptr()),
             ctypes.c_void_p(kk.data_ptr()))
    return bounds, kk


def _box_tables_on(device, in_size, out_size):
    key = (str(device), int(in_size), int(out_size))
    ent = _box_tables_cache.get(key)
    if ent is None:
        b, k = box_resize_tables(in_size, out_size)
        ent = (b.to(device), k.to(device))
        _box_tables_cache[key] = ent
    return ent


def u8_box_resize_normalize(src, index, out_hw, mean, std, flip=None):
    """uint8 images [M, Hs, Ws, 3] (decoded RGB, HWC, on the device) -> normalised fp32 batch [N, 3, Ho, Wo]: Pillow's BOX
    resize + ToTensor + Normalize in one kernel, bit-exact (csrc/input.cu).  index: int64 [N] sample picks or None (all M
    in order); flip: uint8/bool [N] or None; mean / std: 3 floats each."""
    import ctypes
    if not src.is_cuda:
        raise GlbError("gan_lab_b200 kernels need CUDA tensors (there is no CPU fallback)")
    if src.dtype != torch.uint8 or src.dim() != 4 or src.shape[3] != 3 or not src.is_contiguous():
        raise GlbError("u8_box_resize_normalize: src must be a contiguous uint8 [M, H, W, 3] tensor")
    M, Hs, Ws, _ = src.shape
    Ho, Wo = int(out_hw[0]), int(out_hw[1])
    if index is not None:
        index = index.to(device=src.device, dtype=torch.int64).contiguous()
        N = index.numel()
    else:
        N = M
    if flip is not None:
        flip = flip.to(device=src.device, dtype=torch.uint8).contiguous()
        assert flip.numel() == N
    xb, xk = _box_tables_on(src.device, Ws, Wo)
    yb, yk = _box_tables_on(src.device, Hs, Ho)
    out = torch.empty((N, 3, Ho, Wo), device=src.device, dtype=torch.float32)
    m3 = (ctypes.c_float * 3)(*[float(v) for v in mean])
    s3 = (ctypes.c_float * 3)(*[float(v) for v in std])
    _call("glb_u8_box_resize_normalize", src.data_ptr(), M, _p(index), _p(flip), out.data_ptr(), N, Hs, Ws, Ho, Wo,
          xb.data_ptr(), xk.data_ptr(), xk.shape[1], yb.data_ptr(), yk.data_ptr(), yk.shape[1],
          ctypes.cast(m3, ctypes.c_void_p), ctypes.cast(s3, ctypes.c_void_p), _stream())
    return out
