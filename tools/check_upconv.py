"""Per-layer timing of the upsample-folded convolution (glb_upconv_*) against the two-kernel sequence it replaces
(upsample2x + 3x3 conv / dgrad + upsample2x_bwd / upsample2x + wgrad), cfg2 generator shapes, TF32 path.  CUDA events,
20 launches each after 3 warm-ups."""
import sys
import torch
sys.path.insert(0, ".")
import gan_lab_b200._kernels as K

DEV = "cuda"
SHAPES = [(8, 4, 4, 512, 512), (8, 8, 8, 512, 512), (8, 16, 16, 512, 512), (8, 32, 32, 512, 256), (8, 64, 64, 256, 128)]


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def main():
    K.set_conv_impl("tf32")
    cl = lambda t: t.contiguous(memory_format=torch.channels_last)
    print(f"{'shape':28s} {'kind':6s} {'fused us':>9s} {'TF/s(exec)':>10s} {'two-kernel us':>14s} {'speed-up':>8s}")
    for N, H, W, Ci, Co in SHAPES:
        x = cl(torch.randn(N, Ci, H, W, device=DEV))
        w = cl(torch.randn(Co, Ci, 3, 3, device=DEV))
        gy = cl(torch.randn(N, Co, 2 * H, 2 * W, device=DEV))
        K.upconv_weights(w)
        K.transposed_weight(w)
        xu = K.upsample2x_fwd(x)
        fl = 2.0 * N * H * W * 16 * Ci * Co
        rows = [
            ("fprop", lambda: K.upconv_fprop(x, w, None, 1.0, 1.0, K.ACT_NONE, 0.2),
             lambda: K.conv_fprop(K.upsample2x_fwd(x), w, None, 1, 1.0, 1.0, K.ACT_NONE, 0.2)),
            ("dgrad", lambda: K.upconv_dgrad(gy, w, 1.0), lambda: K.upsample2x_bwd(K.conv_dgrad(gy, w, (2 * H, 2 * W), 1, 1.0))),
            ("wgrad", lambda: K.upconv_wgrad(x, gy, 1.0), lambda: K.conv_wgrad(xu, gy, (3, 3), 1, 1.0)),
        ]
        def prep(down):
            K._up_cache.clear()
            K._folded_weights(w, down, "fwd", both=True)   # both operands in one launch
        print(f"    operand re-layout (both operands, one launch): upconv {timed(lambda: prep(False)):6.1f} us   downconv {timed(lambda: prep(True)):6.1f} us"
              f"   ({(9 + 32) * Ci * Co * 4 / 1e6:.1f} MB)")
        for kind, fused, two in rows:
            tf, tt = timed(fused), timed(two)
            print(f"N{N} {H}x{W}->{2*H}x{2*W} {Ci}->{Co:<6d} {kind:6s} {tf:9.1f} {fl / tf / 1e6:10.1f} {tt:14.1f} {tt / tf:8.2f}")


if __name__ == "__main__":
    main()
