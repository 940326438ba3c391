// Bandwidth-bound "glue" kernels of the G/D training step: fused, 128-bit vectorised, coalesced (NHWC: the
// channel axis is the contiguous one), warp-shuffle / shared-memory reductions.  Each kernel names the ATen
// call sites of the reference it replaces; roofline = HBM bandwidth.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "tc_common.cuh"

namespace cg = cooperative_groups;

namespace glb {
namespace {

constexpr int TPB = 256;

// A/B switch for kernel variants (tools/glue_bw.py): GLB_GLUE_OLD is a bit mask that selects the PREVIOUS implementation of
// bit 0 blur3x3 (9 taps per thread), bit 1 act_bwd (one row per loop iteration).  Unset = current kernels.
inline int glue_variant() {
  static const int v = [] {
    const char* e = getenv("GLB_GLUE_OLD");
    return e ? atoi(e) : 0;
  }();
  return v;
}

// ------------------------------------------------------------------------------------------------
// bias + activation (Conv2dBias + LeakyReLU: reference utils/custom_layers.py:222-226 + shared nn.LeakyReLU)
// ------------------------------------------------------------------------------------------------
__global__ void bias_act_fwd_kernel(const float4* __restrict__ x, const float4* __restrict__ bias, float4* __restrict__ y,
                                    int64_t n4, int C4, float bias_scale, int act, float slope) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = ldg_stream(x + i);
    if (bias != nullptr) {
      const float4 b = __ldg(bias + (i % C4));
      v.x += bias_scale * b.x; v.y += bias_scale * b.y; v.z += bias_scale * b.z; v.w += bias_scale * b.w;
    }
    v.x = act_apply(v.x, act, slope); v.y = act_apply(v.y, act, slope);
    v.z = act_apply(v.z, act, slope); v.w = act_apply(v.w, act, slope);
    stg_stream(y + i, v);
  }
}

// gx = gy * act'(y); optional per-channel reduction of gx into gbias (atomics after block reduce).
// Block layout: threads cover (rows x C4) with C4 quads contiguous; requires blockDim % C4 == 0 or C4 % blockDim == 0.
__global__ void act_bwd_kernel(const float4* __restrict__ gy, const float4* __restrict__ y, float4* __restrict__ gx,
                               float* __restrict__ gbias, int64_t P, int C4, float bias_scale, int act, float slope) {
  extern __shared__ float4 red[];
  const int tid = threadIdx.x;
  // column-quads handled by this thread: if C4 <= blockDim: one quad (tid % C4), rows strided by blockDim / C4
  // else: quads tid, tid + blockDim, ... and one row at a time.
  if (C4 <= (int)blockDim.x) {
    const int q = tid % C4, rl = tid / C4, rows_per_it = blockDim.x / C4;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (rl < rows_per_it) {
      for (int64_t p = (int64_t)blockIdx.x * rows_per_it + rl; p < P; p += (int64_t)gridDim.x * rows_per_it) {
        const int64_t i = p * C4 + q;
        const float4 g = ldg_stream(gy + i);
        float4 o;
        if (y != nullptr) {
          const float4 yy = ldg_stream(y + i);
          o.x = g.x * act_grad(yy.x, act, slope); o.y = g.y * act_grad(yy.y, act, slope);
          o.z = g.z * act_grad(yy.z, act, slope); o.w = g.w * act_grad(yy.w, act, slope);
        } else {
          o = g;
        }
        if (gx != nullptr) stg_stream(gx + i, o);
        s.x += o.x; s.y += o.y; s.z += o.z; s.w += o.w;
      }
    }
    if (gbias != nullptr) {
      red[tid] = s;
      __syncthreads();
      if (tid < C4) {
        float4 t = red[tid];
        for (int r = 1; r < rows_per_it; ++r) {
          const float4 u = red[r * C4 + tid];
          t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
        }
        atomicAdd(gbias + 4 * tid + 0, bias_scale * t.x); atomicAdd(gbias + 4 * tid + 1, bias_scale * t.y);
        atomicAdd(gbias + 4 * tid + 2, bias_scale * t.z); atomicAdd(gbias + 4 * tid + 3, bias_scale * t.w);
      }
    }
  } else {
    for (int q = tid; q < C4; q += blockDim.x) {
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int64_t p = blockIdx.x; p < P; p += gridDim.x) {
        const int64_t i = p * C4 + q;
        const float4 g = ldg_stream(gy + i);
        float4 o;
        if (y != nullptr) {
          const float4 yy = ldg_stream(y + i);
          o.x = g.x * act_grad(yy.x, act, slope); o.y = g.y * act_grad(yy.y, act, slope);
          o.z = g.z * act_grad(yy.z, act, slope); o.w = g.w * act_grad(yy.w, act, slope);
        } else {
          o = g;
        }
        if (gx != nullptr) stg_stream(gx + i, o);
        s.x += o.x; s.y += o.y; s.z += o.z; s.w += o.w;
      }
      if (gbias != nullptr) {
        atomicAdd(gbias + 4 * q + 0, bias_scale * s.x); atomicAdd(gbias + 4 * q + 1, bias_scale * s.y);
        atomicAdd(gbias + 4 * q + 2, bias_scale * s.z); atomicAdd(gbias + 4 * q + 3, bias_scale * s.w);
      }
    }
  }
}

// Same contract for C4 <= blockDim, four rows per loop iteration: the eight 128-bit loads of an iteration are issued before
// any of them is used, so a thread keeps 128 B in flight instead of 32 B (a streaming kernel with two inputs needs that
// memory-level parallelism to approach the HBM roofline).
__global__ void __launch_bounds__(TPB) act_bwd_u4_kernel(const float4* __restrict__ gy, const float4* __restrict__ y,
                                                         float4* __restrict__ gx, float* __restrict__ gbias, int64_t P, int C4,
                                                         float bias_scale, int act, float slope) {
  extern __shared__ float4 red[];
  constexpr int U = 4;
  const int tid = threadIdx.x;
  const int q = tid % C4, rl = tid / C4, rows_per_it = blockDim.x / C4;
  const int64_t step = (int64_t)gridDim.x * rows_per_it;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t p0 = (int64_t)blockIdx.x * rows_per_it + rl; p0 < P; p0 += U * step) {
    float4 g[U], yy[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t p = p0 + u * step;
      if (p < P) {
        g[u] = ldg_stream(gy + p * C4 + q);
        if (y != nullptr) yy[u] = ldg_stream(y + p * C4 + q);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t p = p0 + u * step;
      if (p < P) {
        float4 o = g[u];
        if (y != nullptr) {
          o.x *= act_grad(yy[u].x, act, slope); o.y *= act_grad(yy[u].y, act, slope);
          o.z *= act_grad(yy[u].z, act, slope); o.w *= act_grad(yy[u].w, act, slope);
        }
        if (gx != nullptr) stg_stream(gx + p * C4 + q, o);
        s.x += o.x; s.y += o.y; s.z += o.z; s.w += o.w;
      }
    }
  }
  if (gbias != nullptr) {
    red[tid] = s;
    __syncthreads();
    if (tid < C4) {
      float4 t = red[tid];
      for (int r = 1; r < rows_per_it; ++r) {
        const float4 v = red[r * C4 + tid];
        t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
      }
      atomicAdd(gbias + 4 * tid + 0, bias_scale * t.x); atomicAdd(gbias + 4 * tid + 1, bias_scale * t.y);
      atomicAdd(gbias + 4 * tid + 2, bias_scale * t.z); atomicAdd(gbias + 4 * tid + 3, bias_scale * t.w);
    }
  }
}

// scalar fallback of the column sum for C % 4 != 0 (e.g. 3-channel or 1-channel rows)
__global__ void colsum_scalar_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t P, int C, float scale) {
  for (int c = blockIdx.y; c < C; c += gridDim.y) {
    float s = 0.f;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (int64_t)gridDim.x * blockDim.x) s += x[p * C + c];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) atomicAdd(out + c, scale * s);
  }
}

// coef != nullptr: alpha = coef[ia], beta = coef[ib] are read on the device (a fade-in alpha that moves every iteration
// must not be baked into a captured CUDA graph)
__global__ void axpby_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y, int64_t n,
                             float alpha, float beta, const float* __restrict__ coef = nullptr, int ia = 0, int ib = 0) {
  if (coef != nullptr) { alpha = __ldg(coef + ia); beta = __ldg(coef + ib); }
  const int64_t n4 = n >> 2;
  const float4* a4 = reinterpret_cast<const float4*>(a);
  const float4* b4 = reinterpret_cast<const float4*>(b);
  float4* y4 = reinterpret_cast<float4*>(y);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, t0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  for (int64_t i = t0; i < n4; i += stride) {
    float4 u = ldg_stream(a4 + i), o;
    if (b != nullptr) {
      const float4 v = ldg_stream(b4 + i);
      o.x = alpha * u.x + beta * v.x; o.y = alpha * u.y + beta * v.y; o.z = alpha * u.z + beta * v.z; o.w = alpha * u.w + beta * v.w;
    } else {
      o.x = alpha * u.x; o.y = alpha * u.y; o.z = alpha * u.z; o.w = alpha * u.w;
    }
    stg_stream(y4 + i, o);
  }
  for (int64_t i = (n4 << 2) + t0; i < n; i += stride) y[i] = alpha * a[i] + (b != nullptr ? beta * b[i] : 0.f);
}

__global__ void sumsq_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t n, float scale) {
  __shared__ float red[TPB / 32];
  float s = 0.f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, t0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t n4 = n >> 2;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  for (int64_t i = t0; i < n4; i += stride) {
    const float4 v = ldg_stream(x4 + i);
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  for (int64_t i = (n4 << 2) + t0; i < n; i += stride) s += x[i] * x[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = (threadIdx.x < TPB / 32) ? red[threadIdx.x] : 0.f;
    s = warp_sum(s);
    if (threadIdx.x == 0) atomicAdd(out, scale * s);
  }
}

// ------------------------------------------------------------------------------------------------
// PixelNorm (reference utils/custom_layers.py:85-86: pow, mean, add, rsqrt, mul = 5 ATen kernels -> 1)
// one warp per pixel row; C up to 1024 kept in registers (C4 <= 8 quads per lane), else re-read.
// ------------------------------------------------------------------------------------------------
template <int QPL>  // quads per lane
__global__ void pixelnorm_fwd_kernel(const float4* __restrict__ x, float4* __restrict__ y, int64_t P, int C4, float invC, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t p = warp; p < P; p += nwarps) {
    float4 v[QPL];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < QPL; ++i) {
      const int q = lane + 32 * i;
      if (q < C4) {
        v[i] = ldg_stream(x + p * C4 + q);
        s += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
      }
    }
    s = warp_sum(s);
    const float r = rsqrtf(s * invC + eps);
#pragma unroll
    for (int i = 0; i < QPL; ++i) {
      const int q = lane + 32 * i;
      if (q < C4) stg_stream(y + p * C4 + q, make_float4(v[i].x * r, v[i].y * r, v[i].z * r, v[i].w * r));
    }
  }
}

// y = x*r, r = (mean(x^2)+eps)^-1/2 ;  gx = r*gy - x * r^3 * mean(gy*x)
template <int QPL>
__global__ void pixelnorm_bwd_kernel(const float4* __restrict__ gy, const float4* __restrict__ x, float4* __restrict__ gx,
                                     int64_t P, int C4, float invC, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t p = warp; p < P; p += nwarps) {
    float4 v[QPL], g[QPL];
    float s = 0.f, d = 0.f;
#pragma unroll
    for (int i = 0; i < QPL; ++i) {
      const int q = lane + 32 * i;
      if (q < C4) {
        v[i] = ldg_stream(x + p * C4 + q);
        g[i] = ldg_stream(gy + p * C4 + q);
        s += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
        d += v[i].x * g[i].x + v[i].y * g[i].y + v[i].z * g[i].z + v[i].w * g[i].w;
      }
    }
    s = warp_sum(s); d = warp_sum(d);
    const float r = rsqrtf(s * invC + eps);
    const float k = r * r * r * d * invC;
#pragma unroll
    for (int i = 0; i < QPL; ++i) {
      const int q = lane + 32 * i;
      if (q < C4)
        stg_stream(gx + p * C4 + q, make_float4(r * g[i].x - k * v[i].x, r * g[i].y - k * v[i].y,
                                                r * g[i].z - k * v[i].z, r * g[i].w - k * v[i].w));
    }
  }
}

// ------------------------------------------------------------------------------------------------
// 3x3 binomial blur, NHWC (reference get_blur_op: depthwise F.conv2d groups=C, utils/custom_layers.py:50-51).
// Each thread produces one channel-quad of one pixel from its 3x3 neighbourhood; neighbouring threads
// share rows through L1 (the 9 taps of adjacent pixels overlap 6/9), so DRAM traffic stays ~8 B/elem.
// ------------------------------------------------------------------------------------------------
__global__ void blur3x3_tap9_kernel(const float4* __restrict__ x, float4* __restrict__ y, int N, int H, int W, int C4) {
  const int64_t total = (int64_t)N * H * W * C4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int q = (int)(i % C4);
    int64_t t = i / C4;
    const int w = (int)(t % W); t /= W;
    const int h = (int)(t % H);
    const int n = (int)(t / H);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
      const int hh = h + dy;
      if (hh < 0 || hh >= H) continue;
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const int ww = w + dx;
        if (ww < 0 || ww >= W) continue;
        const float k = (float)((2 - (dy != 0)) * (2 - (dx != 0))) * 0.0625f;
        const float4 v = __ldg(x + (((int64_t)n * H + hh) * W + ww) * C4 + q);
        acc.x = fmaf(k, v.x, acc.x); acc.y = fmaf(k, v.y, acc.y); acc.z = fmaf(k, v.z, acc.z); acc.w = fmaf(k, v.w, acc.w);
      }
    }
    stg_stream(y + i, acc);
  }
}

// Sliding-window version (default): a thread owns one channel quad of PW adjacent pixels and walks down TH rows keeping the
// horizontal [1 2 1] sums of the previous two rows in registers, so every output costs (PW+2)/PW * (TH+2)/TH loads
// (2.5 at PW=2, TH=8) instead of 9; the shared taps of neighbouring threads hit L1, halo rows of neighbouring tiles hit L2.
// MASK = true: the backward of "activation then blur" in one pass, g = blur(x) * act'(mask), plus the bias gradient (column
// sums of g; needs blockDim % C4 == 0 so that a thread's channel quad is loop invariant): saves the separate act_bwd pass
// (12 B/element) for 4 B/element of extra mask reads.
template <int PW, bool MASK>
__global__ void blur3x3_kernel(const float4* __restrict__ x, float4* __restrict__ y, int N, int H, int W, int C4, int TH,
                               const float4* __restrict__ mask = nullptr, float* __restrict__ gbias = nullptr,
                               float bias_scale = 1.f, int act = 0, float slope = 0.f) {
  extern __shared__ float4 red[];
  const int WG = (W + PW - 1) / PW, HT = (H + TH - 1) / TH;
  const int64_t total = (int64_t)N * HT * WG * C4;
  float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int q = (int)(i % C4);
    int64_t t = i / C4;
    const int w0 = (int)(t % WG) * PW; t /= WG;
    const int h0 = (int)(t % HT) * TH;
    const int n = (int)(t / HT);
    const int h1 = min(H, h0 + TH);
    const float4* xn = x + (int64_t)n * H * W * C4 + q;
    float4* yn = y + (int64_t)n * H * W * C4 + q;
    float4 rm[PW], rc[PW], rn[PW];
    auto hrow = [&](int hh, float4* r) {
      if (hh < 0 || hh >= H) {
#pragma unroll
        for (int j = 0; j < PW; ++j) r[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
      }
      const float4* row = xn + (int64_t)hh * W * C4;
      float4 v[PW + 2];
#pragma unroll
      for (int k = 0; k < PW + 2; ++k) {
        const int ww = w0 - 1 + k;
        v[k] = (ww >= 0 && ww < W) ? __ldg(row + (int64_t)ww * C4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int j = 0; j < PW; ++j) {
        r[j].x = v[j].x + 2.f * v[j + 1].x + v[j + 2].x; r[j].y = v[j].y + 2.f * v[j + 1].y + v[j + 2].y;
        r[j].z = v[j].z + 2.f * v[j + 1].z + v[j + 2].z; r[j].w = v[j].w + 2.f * v[j + 1].w + v[j + 2].w;
      }
    };
    hrow(h0 - 1, rm);
    hrow(h0, rc);
    for (int h = h0; h < h1; ++h) {
      hrow(h + 1, rn);
#pragma unroll
      for (int j = 0; j < PW; ++j) {
        if (w0 + j < W) {
          float4 o;
          o.x = 0.0625f * (rm[j].x + 2.f * rc[j].x + rn[j].x); o.y = 0.0625f * (rm[j].y + 2.f * rc[j].y + rn[j].y);
          o.z = 0.0625f * (rm[j].z + 2.f * rc[j].z + rn[j].z); o.w = 0.0625f * (rm[j].w + 2.f * rc[j].w + rn[j].w);
          if (MASK) {
            const float4 mk = ldg_stream(mask + (int64_t)n * H * W * C4 + q + ((int64_t)h * W + w0 + j) * C4);
            o.x *= act_grad(mk.x, act, slope); o.y *= act_grad(mk.y, act, slope);
            o.z *= act_grad(mk.z, act, slope); o.w *= act_grad(mk.w, act, slope);
            bsum.x += o.x; bsum.y += o.y; bsum.z += o.z; bsum.w += o.w;
          }
          stg_stream(yn + ((int64_t)h * W + w0 + j) * C4, o);
        }
        rm[j] = rc[j];
        rc[j] = rn[j];
      }
    }
  }
  if (MASK && gbias != nullptr) {      // threads tid, tid + C4, ... of the block own the same channel quad
    const int tid = threadIdx.x;
    red[tid] = bsum;
    __syncthreads();
    if (tid < C4) {
      float4 t = red[tid];
      for (int r = tid + C4; r < (int)blockDim.x; r += C4) {
        const float4 v = red[r];
        t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
      }
      atomicAdd(gbias + 4 * tid + 0, bias_scale * t.x); atomicAdd(gbias + 4 * tid + 1, bias_scale * t.y);
      atomicAdd(gbias + 4 * tid + 2, bias_scale * t.z); atomicAdd(gbias + 4 * tid + 3, bias_scale * t.w);
    }
  }
}


// ------------------------------------------------------------------------------------------------ blur, shared-memory tiles
// The same filter with the input staged by TMA: a 16 x 16-pixel x 32-channel output tile needs an 18 x 18 x 32 box (40.5 KB, one
// cp.async.bulk.tensor with zero fill outside the map = the zero padding), double buffered, so the HBM / L2 reads are few large
// asynchronous transactions instead of 4-5 dependent 16-byte loads per thread and row.  A thread owns one (column, channel quad)
// of half the tile and slides down 8 rows with the horizontal [1 2 1] sums of two rows in registers: 30 conflict-free LDS.128 per
// 8 outputs.  Blocks own contiguous tile ranges ordered channel-block-major, so the MASK variant (blur + activation mask + bias
// gradient, see blur3x3_kernel) flushes its column sums only when the channel block changes.
constexpr int BT = 16;                                   // tile edge (pixels)
constexpr int BTB = BT + 2;                              // box edge
constexpr int BT_TX = BTB * BTB * 128;                   // bytes per box
constexpr int BT_STAGE = ((BT_TX + 1023) / 1024) * 1024;

template <bool MASK>
__global__ void __launch_bounds__(256, 2)
blur_tile_kernel(const __grid_constant__ CUtensorMap tmX, float4* __restrict__ y, int N, int H, int W, int C4, int tiles_w, int tiles_h,
                 int total, const float4* __restrict__ mask, float* __restrict__ gbias, float bias_scale, int act, float slope) {
  using namespace tc;
  extern __shared__ uint8_t bt_raw[];
  __shared__ uint64_t full[2];
  __shared__ float4 red[256];
  const uint32_t raw = smem_u32(bt_raw);
  const uint32_t base = (raw + 127u) & ~127u;
  const float4* tile0 = reinterpret_cast<const float4*>(bt_raw + (base - raw));
  const int tid = threadIdx.x, q = tid & 7, col = (tid >> 3) & 15, half = tid >> 7;
  const int per_cb = N * tiles_h * tiles_w;
  const int t_begin = (int)(((long long)total * blockIdx.x) / gridDim.x), t_end = (int)(((long long)total * (blockIdx.x + 1)) / gridDim.x);
  if (tid == 0) {
    mbar_init(smem_u32(&full[0]), 1);
    mbar_init(smem_u32(&full[1]), 1);
    fence_barrier_init();
    prefetch_tmap(&tmX);
  }
  __syncthreads();
  auto issue = [&](int t, int stage) {
    const int cb = t / per_cb;
    int r = t - cb * per_cb;
    const int tw = r % tiles_w; r /= tiles_w;
    const int th = r % tiles_h;
    const int n = r / tiles_h;
    mbar_expect_tx(smem_u32(&full[stage]), BT_TX);
    tma_load_4d(base + stage * BT_STAGE, &tmX, smem_u32(&full[stage]), cb * 32, tw * BT - 1, th * BT - 1, n);
  };
  if (tid == 0 && t_begin < t_end) issue(t_begin, 0);
  float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);
  int cur_cb = -1;
  auto flush = [&]() {                                    // column sums of the finished channel block -> gbias
    red[tid] = bsum;
    __syncthreads();
    if (tid < 8) {
      float4 s = red[tid];
      for (int r = tid + 8; r < 256; r += 8) { const float4 v = red[r]; s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w; }
      float* gb = gbias + (cur_cb * 8 + tid) * 4;
      atomicAdd(gb + 0, bias_scale * s.x); atomicAdd(gb + 1, bias_scale * s.y);
      atomicAdd(gb + 2, bias_scale * s.z); atomicAdd(gb + 3, bias_scale * s.w);
    }
    __syncthreads();
    bsum = make_float4(0.f, 0.f, 0.f, 0.f);
  };
  for (int t = t_begin, it = 0; t < t_end; ++t, ++it) {
    const int stage = it & 1;
    if (tid == 0 && t + 1 < t_end) issue(t + 1, stage ^ 1);           // that stage was last read in iteration it - 1
    const int cb = t / per_cb;
    int r = t - cb * per_cb;
    const int tw = r % tiles_w; r /= tiles_w;
    const int th = r % tiles_h;
    const int n = r / tiles_h;
    if (MASK && gbias != nullptr && cb != cur_cb) {
      if (cur_cb >= 0) flush();
      cur_cb = cb;
    }
    const int r0 = half * 8, w = tw * BT + col;
    float4 mk[8];
    if (MASK) {                                           // the eight mask loads of this thread go out before the tile is awaited
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int h = th * BT + r0 + j;
        mk[j] = (h < H && w < W) ? ldg_stream(mask + (((int64_t)n * H + h) * W + w) * C4 + cb * 8 + q) : make_float4(1.f, 1.f, 1.f, 1.f);
      }
    }
    mbar_wait(smem_u32(&full[stage]), (it >> 1) & 1);
    const float4* tl = tile0 + stage * (BT_STAGE / 16) + q;
    auto hsum = [&](int br) {
      const float4 a = tl[(br * BTB + col) * 8], b = tl[(br * BTB + col + 1) * 8], c = tl[(br * BTB + col + 2) * 8];
      return make_float4(a.x + 2.f * b.x + c.x, a.y + 2.f * b.y + c.y, a.z + 2.f * b.z + c.z, a.w + 2.f * b.w + c.w);
    };
    float4 rm = hsum(r0), rc = hsum(r0 + 1);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 rn = hsum(r0 + j + 2);
      const int h = th * BT + r0 + j;
      if (h < H && w < W) {
        float4 o = make_float4(0.0625f * (rm.x + 2.f * rc.x + rn.x), 0.0625f * (rm.y + 2.f * rc.y + rn.y),
                               0.0625f * (rm.z + 2.f * rc.z + rn.z), 0.0625f * (rm.w + 2.f * rc.w + rn.w));
        const int64_t idx = (((int64_t)n * H + h) * W + w) * C4 + cb * 8 + q;
        if (MASK) {
          o.x *= act_grad(mk[j].x, act, slope); o.y *= act_grad(mk[j].y, act, slope);
          o.z *= act_grad(mk[j].z, act, slope); o.w *= act_grad(mk[j].w, act, slope);
          bsum.x += o.x; bsum.y += o.y; bsum.z += o.z; bsum.w += o.w;
        }
        stg_stream(y + idx, o);
      }
      rm = rc; rc = rn;
    }
    __syncthreads();                                                   // everyone is done with this stage before it is refilled
  }
  if (MASK && gbias != nullptr && cur_cb >= 0) flush();
}

// -> GLB_OK if launched, GLB_ERR_UNSUPPORTED if the shape is left to the register-window kernels
template <bool MASK>
int blur_tile_launch(const float* x, float* y, int N, int H, int W, int C, const float* mask, float* gbias, float bias_scale, int act,
                     float slope, cudaStream_t st) {
  if (C % 32 != 0 || H < 32 || W < 32 || (glue_variant() & 4)) return GLB_ERR_UNSUPPORTED;
  CUtensorMap tm;
  const uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)N};
  const uint64_t strides[3] = {(uint64_t)C * 4, (uint64_t)W * C * 4, (uint64_t)H * W * C * 4};
  const uint32_t box[4] = {32u, (uint32_t)BTB, (uint32_t)BTB, 1u};
  if (int rc = tc::make_tmap(&tm, x, 4, dims, strides, box, "blur input", false, false, true)) return rc;
  const int tiles_w = (W + BT - 1) / BT, tiles_h = (H + BT - 1) / BT;
  const int64_t total = (int64_t)(C / 32) * N * tiles_h * tiles_w;
  if (total >= (1ll << 31)) return GLB_ERR_UNSUPPORTED;
  const int smem = 2 * BT_STAGE + 128;
  static bool configured = false;
  if (!configured) {
    GLB_CUDA(cudaFuncSetAttribute(blur_tile_kernel<MASK>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  const int grid = total < 2 * kNumSMs ? (int)total : 2 * kNumSMs;
  blur_tile_kernel<MASK><<<grid, 256, smem, st>>>(tm, (float4*)y, N, H, W, C / 4, tiles_w, tiles_h, (int)total, (const float4*)mask, gbias,
                                                 bias_scale, act, slope);
  GLB_CHECK_LAUNCH("blur_tile_kernel");
  return GLB_OK;
}

// ------------------------------------------------------------------------------------------------
// nearest x2 upsample and its adjoint; 2x2 avg-pool (+bias+act) and its backward
// ------------------------------------------------------------------------------------------------
__global__ void upsample2x_fwd_kernel(const float4* __restrict__ x, float4* __restrict__ y, int N, int H, int W, int C4) {
  // one thread per INPUT quad: read once, write 4 output pixels
  const int64_t total = (int64_t)N * H * W * C4;
  const int OW = 2 * W;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int q = (int)(i % C4);
    int64_t t = i / C4;
    const int w = (int)(t % W); t /= W;
    const int h = (int)(t % H);
    const int64_t n = t / H;
    const float4 v = ldg_stream(x + i);
    float4* o = y + ((n * 2 * H + 2 * h) * OW + 2 * w) * C4 + q;
    stg_stream(o, v); stg_stream(o + C4, v);
    stg_stream(o + (int64_t)OW * C4, v); stg_stream(o + (int64_t)OW * C4 + C4, v);
  }
}

__global__ void pool2x2_kernel(const float4* __restrict__ x, const float4* __restrict__ bias, float4* __restrict__ y,
                               int N, int H, int W, int C4, float mult, float bias_scale, int act, float slope) {
  // x [N,H,W,C] -> y [N,H/2,W/2,C]: y = act(mult * sum_2x2(x) + bias_scale*bias); mult = .25 (avg) or 1 (upsample adjoint)
  const int OH = H / 2, OW = W / 2;
  const int64_t total = (int64_t)N * OH * OW * C4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int q = (int)(i % C4);
    int64_t t = i / C4;
    const int w = (int)(t % OW); t /= OW;
    const int h = (int)(t % OH);
    const int64_t n = t / OH;
    const float4* p = x + ((n * H + 2 * h) * W + 2 * w) * C4 + q;
    const float4 a = ldg_stream(p), b = ldg_stream(p + C4), c = ldg_stream(p + (int64_t)W * C4), d = ldg_stream(p + (int64_t)W * C4 + C4);
    float4 v = make_float4(mult * (a.x + b.x + c.x + d.x), mult * (a.y + b.y + c.y + d.y),
                           mult * (a.z + b.z + c.z + d.z), mult * (a.w + b.w + c.w + d.w));
    if (bias != nullptr) {
      const float4 bb = __ldg(bias + q);
      v.x += bias_scale * bb.x; v.y += bias_scale * bb.y; v.z += bias_scale * bb.z; v.w += bias_scale * bb.w;
    }
    v.x = act_apply(v.x, act, slope); v.y = act_apply(v.y, act, slope);
    v.z = act_apply(v.z, act, slope); v.w = act_apply(v.w, act, slope);
    stg_stream(y + i, v);
  }
}

// gx [N,H,W,C] = 0.25 * gy * act'(y) broadcast to the 2x2 window; gbias += bias_scale * sum(gy*act'(y))
__global__ void pool_bias_act_bwd_kernel(const float4* __restrict__ gy, const float4* __restrict__ y, float4* __restrict__ gx,
                                         float* __restrict__ gbias, int N, int H, int W, int C4, float bias_scale, int act,
                                         float slope) {
  extern __shared__ float4 red[];
  const int OH = H / 2, OW = W / 2;
  const int64_t P = (int64_t)N * OH * OW;
  const int tid = threadIdx.x;
  const int lanes = min(C4, (int)blockDim.x);
  const int rows_per_it = blockDim.x / lanes;
  const int rl = tid / lanes;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (rl < rows_per_it) {
    for (int q = tid % lanes; q < C4; q += lanes) {
      s = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int64_t p = (int64_t)blockIdx.x * rows_per_it + rl; p < P; p += (int64_t)gridDim.x * rows_per_it) {
        const int64_t i = p * C4 + q;
        const float4 g = ldg_stream(gy + i);
        float4 o = g;
        if (y != nullptr) {
          const float4 yy = ldg_stream(y + i);
          o.x = g.x * act_grad(yy.x, act, slope); o.y = g.y * act_grad(yy.y, act, slope);
          o.z = g.z * act_grad(yy.z, act, slope); o.w = g.w * act_grad(yy.w, act, slope);
        }
        s.x += o.x; s.y += o.y; s.z += o.z; s.w += o.w;
        const int w = (int)(p % OW);
        const int64_t t = p / OW;
        const int h = (int)(t % OH);
        const int64_t n = t / OH;
        const float4 v = make_float4(.25f * o.x, .25f * o.y, .25f * o.z, .25f * o.w);
        float4* dst = gx + ((n * H + 2 * h) * W + 2 * w) * C4 + q;
        stg_stream(dst, v); stg_stream(dst + C4, v);
        stg_stream(dst + (int64_t)W * C4, v); stg_stream(dst + (int64_t)W * C4 + C4, v);
      }
      if (gbias != nullptr && C4 > lanes) {  // wide-channel case: one row per block iteration, direct atomics
        atomicAdd(gbias + 4 * q + 0, bias_scale * s.x); atomicAdd(gbias + 4 * q + 1, bias_scale * s.y);
        atomicAdd(gbias + 4 * q + 2, bias_scale * s.z); atomicAdd(gbias + 4 * q + 3, bias_scale * s.w);
      }
    }
  }
  if (gbias != nullptr && C4 <= lanes) {
    red[tid] = (rl < rows_per_it) ? s : make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    if (tid < C4) {
      float4 t = red[tid];
      for (int r = 1; r < rows_per_it; ++r) {
        const float4 u = red[r * C4 + tid];
        t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
      }
      atomicAdd(gbias + 4 * tid + 0, bias_scale * t.x); atomicAdd(gbias + 4 * tid + 1, bias_scale * t.y);
      atomicAdd(gbias + 4 * tid + 2, bias_scale * t.z); atomicAdd(gbias + 4 * tid + 3, bias_scale * t.w);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// image-space fade-in helpers (NCHW)
// ------------------------------------------------------------------------------------------------
__global__ void fade_up_blend_kernel(const float* __restrict__ lo, const float* __restrict__ hi, float* __restrict__ out,
                                     int64_t planes, int H, int W, float alpha, const float* __restrict__ coef = nullptr) {
  if (coef != nullptr) alpha = __ldg(coef);
  const int64_t total = planes * H * W;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int w = (int)(i % W);
    int64_t t = i / W;
    const int h = (int)(t % H);
    const int64_t pl = t / H;
    const float l = __ldg(lo + (pl * (H / 2) + h / 2) * (W / 2) + w / 2);
    out[i] = (1.f - alpha) * l + alpha * hi[i];
  }
}

__global__ void fade_up_blend_bwd_kernel(const float* __restrict__ gout, float* __restrict__ glo, float* __restrict__ ghi,
                                         int64_t planes, int H, int W, float alpha, const float* __restrict__ coef = nullptr) {
  if (coef != nullptr) alpha = __ldg(coef);
  // one thread per LOW-res pixel
  const int LH = H / 2, LW = W / 2;
  const int64_t total = planes * LH * LW;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int w = (int)(i % LW);
    int64_t t = i / LW;
    const int h = (int)(t % LH);
    const int64_t pl = t / LH;
    const int64_t o = (pl * H + 2 * h) * W + 2 * w;
    const float a = gout[o], b = gout[o + 1], c = gout[o + W], d = gout[o + W + 1];
    glo[i] = (1.f - alpha) * (a + b + c + d);
    ghi[o] = alpha * a; ghi[o + 1] = alpha * b; ghi[o + W] = alpha * c; ghi[o + W + 1] = alpha * d;
  }
}

__global__ void fade_real_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t planes, int H, int W, float alpha,
                                 const float* __restrict__ coef = nullptr) {
  if (coef != nullptr) alpha = __ldg(coef);
  const int LH = H / 2, LW = W / 2;
  const int64_t total = planes * LH * LW;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int w = (int)(i % LW);
    int64_t t = i / LW;
    const int h = (int)(t % LH);
    const int64_t pl = t / LH;
    const int64_t o = (pl * H + 2 * h) * W + 2 * w;
    const float a = x[o], b = x[o + 1], c = x[o + W], d = x[o + W + 1];
    const float m = (1.f - alpha) * (0.25f * (a + b + c + d));
    out[o] = m + alpha * a; out[o + 1] = m + alpha * b; out[o + W] = m + alpha * c; out[o + W + 1] = m + alpha * d;
  }
}

// ------------------------------------------------------------------------------------------------
// logit losses (single block; n = batch size)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float softplus_f(float x) { return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x))); }
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void d_logit_loss_kernel(const float* __restrict__ dg, const float* __restrict__ dr, float* __restrict__ loss,
                                    float* __restrict__ gg, float* __restrict__ gr, int n, int kind, float eps_drift) {
  __shared__ float red[32];
  float s = 0.f;
  const float inv = 1.f / n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float a = dg[i], b = dr[i];
    float l, ga, gb;
    if (kind == 0) { l = a - b; ga = 1.f; gb = -1.f; }           // wgan: (d_gen - d_real).mean()
    else { l = softplus_f(a) + softplus_f(-b); ga = sigmoid_f(a); gb = -sigmoid_f(-b); }  // BCE(gen,0)+BCE(real,1)
    l += eps_drift * b * b; gb += 2.f * eps_drift * b;
    s += l;
    gg[i] = ga * inv; gr[i] = gb * inv;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
    s = warp_sum(s);
    if (threadIdx.x == 0) loss[0] = s * inv;
  }
}

__global__ void g_logit_loss_kernel(const float* __restrict__ d, float* __restrict__ loss, float* __restrict__ g, int n, int kind) {
  __shared__ float red[32];
  float s = 0.f;
  const float inv = 1.f / n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float a = d[i];
    float l, ga;
    if (kind == 0) { l = -a; ga = -1.f; }                                  // wgan
    else if (kind == 1) { l = softplus_f(-a); ga = -sigmoid_f(-a); }       // nonsaturating: BCE(out, 1)
    else { l = -softplus_f(a); ga = -sigmoid_f(a); }                       // minimax: -BCE(out, 0)
    s += l;
    g[i] = ga * inv;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
    s = warp_sum(s);
    if (threadIdx.x == 0) loss[0] = s * inv;
  }
}

__global__ void w_ewma_kernel(const float* __restrict__ w, float* __restrict__ ewma, int M, int K, float beta) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  float s = 0.f;
  for (int m = 0; m < M; ++m) s += w[(int64_t)m * K + k];
  s /= (float)M;
  ewma[k] = (beta == 0.f) ? s : s * (1.f - beta) + ewma[k] * beta;
}

// ------------------------------------------------------------------------------------------------
// minibatch stddev (reference utils/custom_layers.py:117-140).  One thread-block CLUSTER of MB_CL CTAs per group of `G`
// samples: each CTA owns a slice of the D = H*W*C positions (NHWC order) of the group, the group-wide scalars (the stddev
// average, the second-order term T) are folded across the cluster through distributed shared memory, then every CTA
// writes its slice of the output.  (One CTA per group -- two CTAs for the whole cfg2 batch -- took 30-45 us per launch.)
// ------------------------------------------------------------------------------------------------
constexpr int MB_MAXG = 32;
constexpr int MB_CL = 8;          // portable cluster size

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
  if (threadIdx.x < 32) t = warp_sum(t);
  if (threadIdx.x == 0) red[0] = t;
  __syncthreads();
  t = red[0];
  return t;
}

// sum of this CTA's value over the cluster (every CTA gets the total); `slot` is a shared-memory float of the caller
__device__ __forceinline__ float cluster_sum(float v, float* slot) {
  cg::cluster_group cluster = cg::this_cluster();
  if (threadIdx.x == 0) *slot = v;
  cluster.sync();
  float tot = 0.f;
  for (unsigned r = 0; r < cluster.num_blocks(); ++r) tot += *cluster.map_shared_rank(slot, r);
  cluster.sync();                      // nobody leaves (or overwrites its slot) while a peer may still read it
  return tot;
}

__global__ void __cluster_dims__(MB_CL, 1, 1) mbstd_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int HW, int C,
                                                               int G) {
  __shared__ float red[32];
  __shared__ float slot;
  const int g = blockIdx.x / MB_CL, rank = blockIdx.x % MB_CL;
  const int64_t D = (int64_t)HW * C;
  const float* xg = x + (int64_t)g * G * D;
  float s = 0.f;
  if (G > 1) {
    for (int64_t l = (int64_t)rank * blockDim.x + threadIdx.x; l < D; l += (int64_t)MB_CL * blockDim.x) {
      float mu = 0.f;
      for (int j = 0; j < G; ++j) mu += xg[j * D + l];
      mu /= (float)G;
      float ss = 0.f;
      for (int j = 0; j < G; ++j) { const float d = xg[j * D + l] - mu; ss += d * d; }
      s += sqrtf(ss / (float)(G - 1) + 1e-8f);
    }
    s = block_sum(s, red);
  }
  s = cluster_sum(s, &slot) / (float)D;
  float* yg = y + (int64_t)g * G * HW * (C + 1);
  const int64_t tot = (int64_t)G * HW * (C + 1);
  for (int64_t i = (int64_t)rank * blockDim.x + threadIdx.x; i < tot; i += (int64_t)MB_CL * blockDim.x) {
    const int c = (int)(i % (C + 1));
    const int64_t px = i / (C + 1);
    yg[i] = (c < C) ? xg[px * C + c] : s;
  }
}

// no group-wide reduction beyond gsd = sum of the G*HW stddev-channel cotangents, which every CTA recomputes: plain grid
__global__ void mbstd_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ x, float* __restrict__ gx, int HW, int C,
                                 int G) {
  __shared__ float red[32];
  const int g = blockIdx.x / MB_CL, rank = blockIdx.x % MB_CL;
  const int64_t D = (int64_t)HW * C;
  const float* xg = x + (int64_t)g * G * D;
  const float* gyg = gy + (int64_t)g * G * HW * (C + 1);
  float* gxg = gx + (int64_t)g * G * D;
  float gsd = 0.f;
  if (G > 1) {
    for (int64_t px = threadIdx.x; px < (int64_t)G * HW; px += blockDim.x) gsd += gyg[px * (C + 1) + C];
    gsd = block_sum(gsd, red);
  }
  const float a = (G > 1) ? gsd / ((float)(G - 1) * (float)D) : 0.f;
  for (int64_t l = (int64_t)rank * blockDim.x + threadIdx.x; l < D; l += (int64_t)MB_CL * blockDim.x) {
    const int64_t px = l / C;
    const int c = (int)(l % C);
    float xv[MB_MAXG];
    float mu = 0.f;
    for (int j = 0; j < G; ++j) { xv[j] = xg[j * D + l]; mu += xv[j]; }
    mu /= (float)G;
    float ss = 0.f;
    for (int j = 0; j < G; ++j) { xv[j] -= mu; ss += xv[j] * xv[j]; }
    const float inv_sigma = (G > 1) ? rsqrtf(ss / (float)(G - 1) + 1e-8f) : 0.f;
    for (int j = 0; j < G; ++j)
      gxg[j * D + l] = gyg[((int64_t)j * HW + px) * (C + 1) + c] + a * xv[j] * inv_sigma;
  }
}

// double backward: inputs v (cotangent of gx), gy, x  ->  ggx (wrt x), ggy (wrt gy)
__global__ void __cluster_dims__(MB_CL, 1, 1) mbstd_bwdbwd_kernel(const float* __restrict__ v, const float* __restrict__ gy,
                                                                  const float* __restrict__ x, float* __restrict__ ggx,
                                                                  float* __restrict__ ggy, int HW, int C, int G) {
  __shared__ float red[32];
  __shared__ float slot;
  const int g = blockIdx.x / MB_CL, rank = blockIdx.x % MB_CL;
  const int64_t D = (int64_t)HW * C;
  const float* xg = x + (int64_t)g * G * D;
  const float* vg = v + (int64_t)g * G * D;
  const float* gyg = gy + (int64_t)g * G * HW * (C + 1);
  float* ggxg = ggx + (int64_t)g * G * D;
  float* ggyg = ggy + (int64_t)g * G * HW * (C + 1);
  float gsd = 0.f, T = 0.f;
  const float aD = (G > 1) ? 1.f / ((float)(G - 1) * (float)D) : 0.f;
  if (G > 1) {
    for (int64_t px = threadIdx.x; px < (int64_t)G * HW; px += blockDim.x) gsd += gyg[px * (C + 1) + C];
    gsd = block_sum(gsd, red);
  }
  for (int64_t l = (int64_t)rank * blockDim.x + threadIdx.x; l < D; l += (int64_t)MB_CL * blockDim.x) {
    float d[MB_MAXG], vv[MB_MAXG];
    float mu = 0.f, vbar = 0.f;
    for (int j = 0; j < G; ++j) { d[j] = xg[j * D + l]; vv[j] = vg[j * D + l]; mu += d[j]; vbar += vv[j]; }
    mu /= (float)G; vbar /= (float)G;
    float ss = 0.f, vd = 0.f;
    for (int j = 0; j < G; ++j) { d[j] -= mu; ss += d[j] * d[j]; vd += vv[j] * d[j]; }
    if (G > 1) {
      const float is = rsqrtf(ss / (float)(G - 1) + 1e-8f);
      const float is3 = is * is * is / (float)(G - 1);
      T += vd * is * aD;
      for (int j = 0; j < G; ++j) ggxg[j * D + l] = gsd * aD * ((vv[j] - vbar) * is - vd * d[j] * is3);
    } else {
      for (int j = 0; j < G; ++j) ggxg[j * D + l] = 0.f;
    }
  }
  T = cluster_sum(block_sum(T, red), &slot);
  const int64_t tot = (int64_t)G * HW * (C + 1);
  for (int64_t i = (int64_t)rank * blockDim.x + threadIdx.x; i < tot; i += (int64_t)MB_CL * blockDim.x) {
    const int c = (int)(i % (C + 1));
    const int64_t px = i / (C + 1);
    ggyg[i] = (c < C) ? vg[px * C + c] : T;
  }
}

}  // namespace
}  // namespace glb

using namespace glb;

#define REQ(cond, msg) do { if (!(cond)) return glb::shape_fail(msg); } while (0)

extern "C" int glb_bias_act_fwd(const float* x, const float* bias, float* y, int64_t P, int C, float bias_scale, int act,
                                float slope, glb_stream_t stream) {
  REQ(C % 4 == 0 && P > 0, "bias_act_fwd: C % 4 != 0");
  const int64_t n4 = P * (C / 4);
  bias_act_fwd_kernel<<<grid_for(n4, TPB), TPB, 0, (cudaStream_t)stream>>>(
      (const float4*)x, (const float4*)bias, (float4*)y, n4, C / 4, bias_scale, act, slope);
  GLB_CHECK_LAUNCH("bias_act_fwd");
  return GLB_OK;
}

extern "C" int glb_act_bwd(const float* gy, const float* y, float* gx, float* gbias, int64_t P, int C, float bias_scale,
                           int act, float slope, glb_stream_t stream) {
  REQ(C % 4 == 0 && P > 0, "act_bwd: C % 4 != 0");
  const int C4 = C / 4;
  REQ(C4 > TPB || TPB % C4 == 0, "act_bwd: C/4 must divide 256 or exceed it");
  const int rows_per_it = C4 <= TPB ? TPB / C4 : 1;
  int blocks = (int)((P + rows_per_it - 1) / rows_per_it);
  if (C4 <= TPB && !(glue_variant() & 2)) {
    blocks = (blocks + 3) / 4;                       // four rows per thread and iteration
    // with a bias gradient every block ends in C atomics on the same C addresses: 2 blocks per SM keep that tail short
    // (1184 blocks x 512 channels = 0.6 M serialised atomics took longer than the streaming pass itself)
    const int cap = gbias != nullptr ? kNumSMs * 2 : kNumSMs * 8;
    if (blocks > cap) blocks = cap;
    act_bwd_u4_kernel<<<blocks, TPB, TPB * sizeof(float4), (cudaStream_t)stream>>>(
        (const float4*)gy, (const float4*)y, (float4*)gx, gbias, P, C4, bias_scale, act, slope);
    GLB_CHECK_LAUNCH("act_bwd_u4");
    return GLB_OK;
  }
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  act_bwd_kernel<<<blocks, TPB, TPB * sizeof(float4), (cudaStream_t)stream>>>(
      (const float4*)gy, (const float4*)y, (float4*)gx, gbias, P, C4, bias_scale, act, slope);
  GLB_CHECK_LAUNCH("act_bwd");
  return GLB_OK;
}

extern "C" int glb_colsum(const float* x, float* out, int64_t P, int C, float scale, glb_stream_t stream) {
  REQ(P > 0 && C > 0, "colsum");
  if (C % 4 == 0 && (C / 4 > TPB || TPB % (C / 4) == 0))
    return glb_act_bwd(x, nullptr, nullptr, out, P, C, scale, GLB_ACT_NONE, 0.f, stream);
  dim3 grid(grid_for(P, TPB, kNumSMs * 2), C < 64 ? C : 64);
  colsum_scalar_kernel<<<grid, TPB, 0, (cudaStream_t)stream>>>(x, out, P, C, scale);
  GLB_CHECK_LAUNCH("colsum");
  return GLB_OK;
}

extern "C" int glb_axpby(const float* a, const float* b, float* y, int64_t n, float alpha, float beta, glb_stream_t stream) {
  REQ(n > 0, "axpby: n");
  axpby_kernel<<<grid_for((n + 3) / 4, TPB), TPB, 0, (cudaStream_t)stream>>>(a, b, y, n, alpha, beta);
  GLB_CHECK_LAUNCH("axpby");
  return GLB_OK;
}

extern "C" int glb_scale(const float* x, float* y, int64_t n, float scale, glb_stream_t stream) {
  return glb_axpby(x, nullptr, y, n, scale, 0.f, stream);
}

extern "C" int glb_sumsq(const float* x, float* out, int64_t n, float scale, glb_stream_t stream) {
  REQ(n > 0, "sumsq: n");
  sumsq_kernel<<<grid_for((n + 3) / 4, TPB, kNumSMs * 4), TPB, 0, (cudaStream_t)stream>>>(x, out, n, scale);
  GLB_CHECK_LAUNCH("sumsq");
  return GLB_OK;
}

template <int QPL>
static int pixelnorm_launch(bool fwd, const float* gy, const float* x, float* out, int64_t P, int C, float eps, cudaStream_t st) {
  const int64_t warps = P;
  int blocks = (int)((warps * 32 + TPB - 1) / TPB);
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  if (fwd)
    pixelnorm_fwd_kernel<QPL><<<blocks, TPB, 0, st>>>((const float4*)x, (float4*)out, P, C / 4, 1.f / C, eps);
  else
    pixelnorm_bwd_kernel<QPL><<<blocks, TPB, 0, st>>>((const float4*)gy, (const float4*)x, (float4*)out, P, C / 4, 1.f / C, eps);
  GLB_CHECK_LAUNCH("pixelnorm");
  return GLB_OK;
}

static int pixelnorm_dispatch(bool fwd, const float* gy, const float* x, float* out, int64_t P, int C, float eps, cudaStream_t st) {
  REQ(C % 4 == 0 && C <= 1024 && P > 0, "pixelnorm: C % 4 != 0 or C > 1024");
  const int qpl = (C / 4 + 31) / 32;
  if (qpl <= 1) return pixelnorm_launch<1>(fwd, gy, x, out, P, C, eps, st);
  if (qpl <= 2) return pixelnorm_launch<2>(fwd, gy, x, out, P, C, eps, st);
  if (qpl <= 4) return pixelnorm_launch<4>(fwd, gy, x, out, P, C, eps, st);
  return pixelnorm_launch<8>(fwd, gy, x, out, P, C, eps, st);
}

extern "C" int glb_pixelnorm_fwd(const float* x, float* y, int64_t P, int C, float eps, glb_stream_t stream) {
  return pixelnorm_dispatch(true, nullptr, x, y, P, C, eps, (cudaStream_t)stream);
}
extern "C" int glb_pixelnorm_bwd(const float* gy, const float* x, float* gx, int64_t P, int C, float eps, glb_stream_t stream) {
  return pixelnorm_dispatch(false, gy, x, gx, P, C, eps, (cudaStream_t)stream);
}

extern "C" int glb_blur3x3(const float* x, float* y, int N, int H, int W, int C, glb_stream_t stream) {
  REQ(C % 4 == 0, "blur3x3: C % 4 != 0");
  const int C4 = C / 4;
  if (glue_variant() & 1) {
    const int64_t total = (int64_t)N * H * W * C4;
    blur3x3_tap9_kernel<<<grid_for(total, TPB), TPB, 0, (cudaStream_t)stream>>>((const float4*)x, (float4*)y, N, H, W, C4);
    GLB_CHECK_LAUNCH("blur3x3");
    return GLB_OK;
  }
  {
    const int rc = blur_tile_launch<false>(x, y, N, H, W, C, nullptr, nullptr, 1.f, 0, 0.f, (cudaStream_t)stream);
    if (rc != GLB_ERR_UNSUPPORTED) return rc;
  }
  constexpr int PW = 2;
  int TH = 8;   // rows per thread: fewer for small maps so that the grid still fills the machine
  while (TH > 1 && (int64_t)N * ((H + TH - 1) / TH) * ((W + PW - 1) / PW) * C4 < (int64_t)kNumSMs * 4 * TPB) TH >>= 1;
  const int64_t threads = (int64_t)N * ((H + TH - 1) / TH) * ((W + PW - 1) / PW) * C4;
  blur3x3_kernel<PW, false><<<grid_for(threads, TPB), TPB, 0, (cudaStream_t)stream>>>((const float4*)x, (float4*)y, N, H, W, C4, TH);
  GLB_CHECK_LAUNCH("blur3x3");
  return GLB_OK;
}

extern "C" int glb_blur_act_bwd(const float* gz, const float* ymask, float* g, float* gbias, int N, int H, int W, int C,
                                float bias_scale, int act, float slope, glb_stream_t stream) {
  REQ(C % 4 == 0 && TPB % (C / 4) == 0, "blur_act_bwd: C/4 must divide 256");
  {
    const int rc = blur_tile_launch<true>(gz, g, N, H, W, C, ymask, gbias, bias_scale, act, slope, (cudaStream_t)stream);
    if (rc != GLB_ERR_UNSUPPORTED) return rc;
  }
  const int C4 = C / 4;
  constexpr int PW = 2;
  int TH = 8;
  while (TH > 1 && (int64_t)N * ((H + TH - 1) / TH) * ((W + PW - 1) / PW) * C4 < (int64_t)kNumSMs * 4 * TPB) TH >>= 1;
  const int64_t threads = (int64_t)N * ((H + TH - 1) / TH) * ((W + PW - 1) / PW) * C4;
  // with a bias gradient every block ends in C atomics on the same addresses: <= 4 blocks per SM (grid-stride loop)
  const int blocks = grid_for(threads, TPB, gbias != nullptr ? kNumSMs * 4 : kNumSMs * 16);
  blur3x3_kernel<PW, true><<<blocks, TPB, TPB * sizeof(float4), (cudaStream_t)stream>>>(
      (const float4*)gz, (float4*)g, N, H, W, C4, TH, (const float4*)ymask, gbias, bias_scale, act, slope);
  GLB_CHECK_LAUNCH("blur_act_bwd");
  return GLB_OK;
}

extern "C" int glb_upsample2x_fwd(const float* x, float* y, int N, int H, int W, int C, glb_stream_t stream) {
  REQ(C % 4 == 0, "upsample2x: C % 4 != 0");
  const int64_t total = (int64_t)N * H * W * (C / 4);
  upsample2x_fwd_kernel<<<grid_for(total, TPB), TPB, 0, (cudaStream_t)stream>>>((const float4*)x, (float4*)y, N, H, W, C / 4);
  GLB_CHECK_LAUNCH("upsample2x_fwd");
  return GLB_OK;
}

extern "C" int glb_upsample2x_bwd(const float* gy, float* gx, int N, int H, int W, int C, glb_stream_t stream) {
  REQ(C % 4 == 0, "upsample2x_bwd: C % 4 != 0");
  const int64_t total = (int64_t)N * H * W * (C / 4);
  pool2x2_kernel<<<grid_for(total, TPB), TPB, 0, (cudaStream_t)stream>>>((const float4*)gy, nullptr, (float4*)gx, N, 2 * H, 2 * W,
                                                                        C / 4, 1.f, 0.f, GLB_ACT_NONE, 0.f);
  GLB_CHECK_LAUNCH("upsample2x_bwd");
  return GLB_OK;
}

extern "C" int glb_pool_bias_act_fwd(const float* x, const float* bias, float* y, int N, int H, int W, int C, float bias_scale,
                                     int act, float slope, glb_stream_t stream) {
  REQ(C % 4 == 0 && H % 2 == 0 && W % 2 == 0, "pool_bias_act_fwd: C % 4 or odd H/W");
  const int64_t total = (int64_t)N * (H / 2) * (W / 2) * (C / 4);
  pool2x2_kernel<<<grid_for(total, TPB), TPB, 0, (cudaStream_t)stream>>>((const float4*)x, (const float4*)bias, (float4*)y, N, H, W,
                                                                        C / 4, 0.25f, bias_scale, act, slope);
  GLB_CHECK_LAUNCH("pool_bias_act_fwd");
  return GLB_OK;
}

extern "C" int glb_pool_bias_act_bwd(const float* gy, const float* y, float* gx, float* gbias, int N, int H, int W, int C,
                                     float bias_scale, int act, float slope, glb_stream_t stream) {
  REQ(C % 4 == 0 && H % 2 == 0 && W % 2 == 0, "pool_bias_act_bwd: C % 4 or odd H/W");
  const int C4 = C / 4;
  REQ(C4 > TPB || TPB % C4 == 0, "pool_bias_act_bwd: C/4 must divide 256 or exceed it");
  const int lanes = C4 < TPB ? C4 : TPB;
  const int rows_per_it = TPB / lanes;
  const int64_t P = (int64_t)N * (H / 2) * (W / 2);
  int blocks = (int)((P + rows_per_it - 1) / rows_per_it);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  pool_bias_act_bwd_kernel<<<blocks, TPB, TPB * sizeof(float4), (cudaStream_t)stream>>>(
      (const float4*)gy, (const float4*)y, (float4*)gx, gbias, N, H, W, C4, bias_scale, act, slope);
  GLB_CHECK_LAUNCH("pool_bias_act_bwd");
  return GLB_OK;
}

extern "C" int glb_fade_up_blend(const float* lo, const float* hi, float* out, int N, int C, int H, int W, float alpha,
                                 glb_stream_t stream) {
  REQ(H % 2 == 0 && W % 2 == 0, "fade_up_blend: odd H/W");
  const int64_t total = (int64_t)N * C * H * W;
  fade_up_blend_kernel<<<grid_for(total, TPB), TPB, 0, (cudaStream_t)stream>>>(lo, hi, out, (int64_t)N * C, H, W, alpha);
  GLB_CHECK_LAUNCH("fade_up_blend");
  return GLB_OK;
}

extern "C" int glb_fade_up_blend_bwd(const float* gout, float* glo, float* ghi, int N, int C, int H, int W, float alpha,
                                     glb_stream_t stream) {
  REQ(H % 2 == 0 && W % 2 == 0, "fade_up_blend_bwd: odd H/W");
  const int64_t total = (int64_t)N * C * (H / 2) * (W / 2);
  fade_up_blend_bwd_kernel<<<grid_for(total, TPB), TPB, 0, (cudaStream_t)stream>>>(gout, glo, ghi, (int64_t)N * C, H, W, alpha);
  GLB_CHECK_LAUNCH("fade_up_blend_bwd");
  return GLB_OK;
}

extern "C" int glb_fade_real(const float* x, float* out, int N, int C, int H, int W, float alpha, glb_stream_t stream) {
  REQ(H % 2 == 0 && W % 2 == 0, "fade_real: odd H/W");
  const int64_t total = (int64_t)N * C * (H / 2) * (W / 2);
  fade_real_kernel<<<grid_for(total, TPB), TPB, 0, (cudaStream_t)stream>>>(x, out, (int64_t)N * C, H, W, alpha);
  GLB_CHECK_LAUNCH("fade_real");
  return GLB_OK;
}

// ---- the same four with alpha / beta read from a device vector `coef` = [alpha, 1 - alpha] (CUDA-graph replayable fade-in)
extern "C" int glb_axpby_dev(const float* a, const float* b, float* y, int64_t n, const float* coef, int ia, int ib,
                             glb_stream_t stream) {
  REQ(n > 0 && coef != nullptr && ia >= 0 && ib >= 0, "axpby_dev");
  axpby_kernel<<<grid_for((n + 3) / 4, TPB), TPB, 0, (cudaStream_t)stream>>>(a, b, y, n, 0.f, 0.f, coef, ia, ib);
  GLB_CHECK_LAUNCH("axpby_dev");
  return GLB_OK;
}

extern "C" int glb_fade_up_blend_dev(const float* lo, const float* hi, float* out, int N, int C, int H, int W, const float* coef,
                                     glb_stream_t stream) {
  REQ(H % 2 == 0 && W % 2 == 0 && coef != nullptr, "fade_up_blend_dev");
  const int64_t total = (int64_t)N * C * H * W;
  fade_up_blend_kernel<<<grid_for(total, TPB), TPB, 0, (cudaStream_t)stream>>>(lo, hi, out, (int64_t)N * C, H, W, 0.f, coef);
  GLB_CHECK_LAUNCH("fade_up_blend_dev");
  return GLB_OK;
}

extern "C" int glb_fade_up_blend_bwd_dev(const float* gout, float* glo, float* ghi, int N, int C, int H, int W, const float* coef,
                                         glb_stream_t stream) {
  REQ(H % 2 == 0 && W % 2 == 0 && coef != nullptr, "fade_up_blend_bwd_dev");
  const int64_t total = (int64_t)N * C * (H / 2) * (W / 2);
  fade_up_blend_bwd_kernel<<<grid_for(total, TPB), TPB, 0, (cudaStream_t)stream>>>(gout, glo, ghi, (int64_t)N * C, H, W, 0.f, coef);
  GLB_CHECK_LAUNCH("fade_up_blend_bwd_dev");
  return GLB_OK;
}

extern "C" int glb_fade_real_dev(const float* x, float* out, int N, int C, int H, int W, const float* coef, glb_stream_t stream) {
  REQ(H % 2 == 0 && W % 2 == 0 && coef != nullptr, "fade_real_dev");
  const int64_t total = (int64_t)N * C * (H / 2) * (W / 2);
  fade_real_kernel<<<grid_for(total, TPB), TPB, 0, (cudaStream_t)stream>>>(x, out, (int64_t)N * C, H, W, 0.f, coef);
  GLB_CHECK_LAUNCH("fade_real_dev");
  return GLB_OK;
}

extern "C" int glb_d_logit_loss(const float* d_gen, const float* d_real, float* loss, float* g_gen, float* g_real, int n, int kind,
                                float eps_drift, glb_stream_t stream) {
  REQ(n > 0 && kind >= 0 && kind <= 2, "d_logit_loss");
  d_logit_loss_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(d_gen, d_real, loss, g_gen, g_real, n, kind, eps_drift);
  GLB_CHECK_LAUNCH("d_logit_loss");
  return GLB_OK;
}

extern "C" int glb_g_logit_loss(const float* d_out, float* loss, float* g_out, int n, int kind, glb_stream_t stream) {
  REQ(n > 0 && kind >= 0 && kind <= 2, "g_logit_loss");
  g_logit_loss_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(d_out, loss, g_out, n, kind);
  GLB_CHECK_LAUNCH("g_logit_loss");
  return GLB_OK;
}

extern "C" int glb_w_ewma(const float* w, float* ewma, int M, int K, float beta, glb_stream_t stream) {
  REQ(M > 0 && K > 0, "w_ewma");
  w_ewma_kernel<<<(K + 127) / 128, 128, 0, (cudaStream_t)stream>>>(w, ewma, M, K, beta);
  GLB_CHECK_LAUNCH("w_ewma");
  return GLB_OK;
}

extern "C" int glb_mbstd_fwd(const float* x, float* y, int N, int H, int W, int C, int group, glb_stream_t stream) {
  REQ(group >= 1 && group <= MB_MAXG && N % group == 0, "mbstd: group must divide N and be <= 32");
  mbstd_fwd_kernel<<<(N / group) * MB_CL, 256, 0, (cudaStream_t)stream>>>(x, y, H * W, C, group);
  GLB_CHECK_LAUNCH("mbstd_fwd");
  return GLB_OK;
}

extern "C" int glb_mbstd_bwd(const float* gy, const float* x, float* gx, int N, int H, int W, int C, int group, glb_stream_t stream) {
  REQ(group >= 1 && group <= MB_MAXG && N % group == 0, "mbstd: group must divide N and be <= 32");
  mbstd_bwd_kernel<<<(N / group) * MB_CL, 256, 0, (cudaStream_t)stream>>>(gy, x, gx, H * W, C, group);
  GLB_CHECK_LAUNCH("mbstd_bwd");
  return GLB_OK;
}

extern "C" int glb_mbstd_bwdbwd(const float* v, const float* gy, const float* x, float* ggx, float* ggy, int N, int H, int W, int C,
                                int group, glb_stream_t stream) {
  REQ(group >= 1 && group <= MB_MAXG && N % group == 0, "mbstd: group must divide N and be <= 32");
  mbstd_bwdbwd_kernel<<<(N / group) * MB_CL, 256, 0, (cudaStream_t)stream>>>(v, gy, x, ggx, ggy, H * W, C, group);
  GLB_CHECK_LAUNCH("mbstd_bwdbwd");
  return GLB_OK;
}

// ---- scalar-gradient helpers for the penalties ---------------------------------------------------------
namespace glb {
namespace {
__global__ void scale_by_kernel(const float* __restrict__ x, const float* __restrict__ s_dev, float* __restrict__ y, int64_t n,
                                float scale) {
  const float s = scale * __ldg(s_dev);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) y[i] = x[i] * s;
}

__global__ void gp_norm_fwd_kernel(const float* __restrict__ g, float* __restrict__ out, int N, int C, int64_t HW, float gamma,
                                   float scale) {
  __shared__ float red[TPB / 32];
  float s = 0.f;
  const int64_t total = (int64_t)N * HW;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i / HW, p = i % HW;
    float q = 0.f;
    for (int c = 0; c < C; ++c) { const float v = g[(n * C + c) * HW + p]; q = fmaf(v, v, q); }
    const float d = sqrtf(q) - gamma;
    s += d * d;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = (threadIdx.x < TPB / 32) ? red[threadIdx.x] : 0.f;
    s = warp_sum(s);
    if (threadIdx.x == 0) atomicAdd(out, scale * s);
  }
}

__global__ void gp_norm_bwd_kernel(const float* __restrict__ g, const float* __restrict__ s_dev, float* __restrict__ gg, int N, int C,
                                   int64_t HW, float gamma, float scale) {
  const float up = 2.f * scale * __ldg(s_dev);
  const int64_t total = (int64_t)N * HW;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i / HW, p = i % HW;
    float q = 0.f;
    for (int c = 0; c < C; ++c) { const float v = g[(n * C + c) * HW + p]; q = fmaf(v, v, q); }
    const float nrm = sqrtf(q);
    const float k = nrm > 0.f ? up * (nrm - gamma) / nrm : 0.f;
    for (int c = 0; c < C; ++c) gg[(n * C + c) * HW + p] = k * g[(n * C + c) * HW + p];
  }
}

__global__ void interp_rows_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ eps,
                                   float* __restrict__ out, int N, int64_t per_row) {
  const int64_t total = (int64_t)N * per_row;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const float e = __ldg(eps + i / per_row);
    out[i] = e * a[i] + (1.f - e) * b[i];
  }
}
}  // namespace
}  // namespace glb

extern "C" int glb_scale_by(const float* x, const float* s_dev, float* y, int64_t n, float scale, glb_stream_t stream) {
  REQ(n > 0, "scale_by: n");
  scale_by_kernel<<<grid_for(n, TPB), TPB, 0, (cudaStream_t)stream>>>(x, s_dev, y, n, scale);
  GLB_CHECK_LAUNCH("scale_by");
  return GLB_OK;
}
extern "C" int glb_gp_norm_fwd(const float* g, float* out, int N, int C, int64_t HW, float gamma, float scale, glb_stream_t stream) {
  REQ(N > 0 && C > 0 && HW > 0, "gp_norm_fwd");
  gp_norm_fwd_kernel<<<grid_for((int64_t)N * HW, TPB, kNumSMs * 4), TPB, 0, (cudaStream_t)stream>>>(g, out, N, C, HW, gamma, scale);
  GLB_CHECK_LAUNCH("gp_norm_fwd");
  return GLB_OK;
}
extern "C" int glb_gp_norm_bwd(const float* g, const float* s_dev, float* gg, int N, int C, int64_t HW, float gamma, float scale,
                               glb_stream_t stream) {
  REQ(N > 0 && C > 0 && HW > 0, "gp_norm_bwd");
  gp_norm_bwd_kernel<<<grid_for((int64_t)N * HW, TPB), TPB, 0, (cudaStream_t)stream>>>(g, s_dev, gg, N, C, HW, gamma, scale);
  GLB_CHECK_LAUNCH("gp_norm_bwd");
  return GLB_OK;
}
extern "C" int glb_interp_rows(const float* a, const float* b, const float* eps, float* out, int N, int64_t per_row,
                               glb_stream_t stream) {
  REQ(N > 0 && per_row > 0, "interp_rows");
  interp_rows_kernel<<<grid_for((int64_t)N * per_row, TPB), TPB, 0, (cudaStream_t)stream>>>(a, b, eps, out, N, per_row);
  GLB_CHECK_LAUNCH("interp_rows");
  return GLB_OK;
}

// ---- fp32 -> bf16 (round to nearest even): operand copies for the bf16 tensor-core path ---------------------------------
#include <cuda_bf16.h>
namespace glb {
namespace {
__global__ void __launch_bounds__(256) cvt_f32_bf16_kernel(const float4* __restrict__ x, uint4* __restrict__ y, int64_t n8,
                                                           const float* __restrict__ xt, __nv_bfloat16* __restrict__ yt, int tail) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n8; i += stride) {
    const float4 a = ldg_stream(x + 2 * i), b = ldg_stream(x + 2 * i + 1);
    __nv_bfloat162 p0 = __floats2bfloat162_rn(a.x, a.y), p1 = __floats2bfloat162_rn(a.z, a.w);
    __nv_bfloat162 p2 = __floats2bfloat162_rn(b.x, b.y), p3 = __floats2bfloat162_rn(b.z, b.w);
    uint4 o;
    o.x = *reinterpret_cast<uint32_t*>(&p0); o.y = *reinterpret_cast<uint32_t*>(&p1);
    o.z = *reinterpret_cast<uint32_t*>(&p2); o.w = *reinterpret_cast<uint32_t*>(&p3);
    y[i] = o;
  }
  if (blockIdx.x == 0 && threadIdx.x < tail) yt[threadIdx.x] = __float2bfloat16_rn(xt[threadIdx.x]);
}
}  // namespace
}  // namespace glb

extern "C" int glb_cvt_f32_bf16(const float* x, void* y, int64_t n, glb_stream_t stream) {
  if (n <= 0) return GLB_OK;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(y) & 15)) return glb::shape_fail("cvt_f32_bf16: 16-byte alignment");
  const int64_t n8 = n / 8;
  const int tail = (int)(n - 8 * n8);
  glb::cvt_f32_bf16_kernel<<<glb::grid_for(n8 > 0 ? n8 : 1, 256, glb::kNumSMs * 8), 256, 0, (cudaStream_t)stream>>>(
      (const float4*)x, (uint4*)y, n8, x + 8 * n8, reinterpret_cast<__nv_bfloat16*>(y) + 8 * n8, tail);
  GLB_CHECK_LAUNCH("cvt_f32_bf16_kernel");
  return GLB_OK;
}
