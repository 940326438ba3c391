// StyleGAN generator-layer epilogue, fused:  noise injection + bias + leaky-ReLU + InstanceNorm + AdaIN.
//
// Reference call sites replaced (paths relative to gan_lab/): StyleAddNoise.forward stylegan/architectures.py:112-119
// (randn*weight + add), Conv2dBias utils/custom_layers.py:222-226, the shared nn.LeakyReLU, nn.InstanceNorm2d(eps=1e-8)
// utils/custom_layers.py:99 (native_batch_norm on a reshaped tensor) and the AdaIN modulation
// stylegan/architectures.py:460-462 / 524-526 (view, select, contiguous, add, mul, add): ~8 ATen kernels per layer.
//
//   u = x + nw[c]*noise[n,h,w] + b[c];  t = lrelu(u);  xhat = (t - mu[n,c]) * rstd[n,c];  out = xhat*(ys+1) + yb
//
// Plane statistics need all H*W pixels of an (n,c) plane, so forward = stats pass + apply pass (x is re-read
// from L2 for the layers that fit it); backward likewise = reduction pass + apply pass.  Partial sums are
// produced per 256-pixel chunk in fp32 and combined in fp64 (deterministic, no atomics on the statistics).
#include <cooperative_groups.h>
#include <stdlib.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace glb {
namespace {

constexpr int TPB = 256;
constexpr int MAX_CHUNK = 256;  // pixels per block, upper bound
constexpr int UNR = 4;          // rows a thread keeps in flight per loop iteration (memory-level parallelism)

struct SE {
  int N, HW, C4, chunks, chunk;  // chunk = pixels per block (runtime: sized so that ~2 blocks land on every SM)
  float slope;
};

__device__ __forceinline__ float4 lrelu4(float4 u, float slope) {
  return make_float4(u.x > 0.f ? u.x : u.x * slope, u.y > 0.f ? u.y : u.y * slope, u.z > 0.f ? u.z : u.z * slope,
                     u.w > 0.f ? u.w : u.w * slope);
}

__device__ __forceinline__ float4 pre_act(const float4 x, float nz, const float4 nw, const float4 b) {
  return make_float4(x.x + nw.x * nz + b.x, x.y + nw.y * nz + b.y, x.z + nw.z * nz + b.z, x.w + nw.w * nz + b.w);
}

// reduce (a, b) float4 pairs across the rows of the block that share a channel quad; result valid for tid < C4
__device__ __forceinline__ void block_rows_reduce(float4& a, float4& b, float4* red, int C4, int rows) {
  const int tid = threadIdx.x;
  red[tid] = a;
  red[TPB + tid] = b;
  __syncthreads();
  if (tid < C4) {
    for (int r = 1; r < rows; ++r) {
      const float4 u = red[r * C4 + tid], v = red[TPB + r * C4 + tid];
      a.x += u.x; a.y += u.y; a.z += u.z; a.w += u.w;
      b.x += v.x; b.y += v.y; b.z += v.z; b.w += v.w;
    }
  }
}

// ---- forward pass 1: per-chunk sum / sum of squares of t ------------------------------------------
// -> s, ss (valid for tid < C4): sums of (t - k4), (t - k4)^2 over the chunk; k4 = t at pixel `shift_pixel` of the sample
// (< 0: the chunk's own first pixel)
__device__ __forceinline__ void se_fwd_stats_sums(const float4* __restrict__ x, const float* __restrict__ noise,
                                                  const float4* __restrict__ nw, const float4* __restrict__ bias, const SE& g,
                                                  float4* red, int n, int chunk, int shift_pixel, float4& s, float4& ss, float4& k4) {
  const int tid = threadIdx.x, q = tid % g.C4, rl = tid / g.C4, rows = TPB / g.C4;
  const int p0 = chunk * g.chunk, p1 = min(g.HW, p0 + g.chunk);
  const float4 w4 = nw ? __ldg(nw + q) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 b4 = bias ? __ldg(bias + q) : make_float4(0.f, 0.f, 0.f, 0.f);
  s = make_float4(0.f, 0.f, 0.f, 0.f); ss = s;
  // shifted sums: k4 = a sample of the plane, so sum((t-k)^2) - sum(t-k)^2/n does not cancel when |mean| >> std
  // (e.g. the constant-input layer, whose planes are 1 + small noise)
  {
    const int64_t px = (int64_t)n * g.HW + (shift_pixel < 0 ? p0 : shift_pixel);
    k4 = lrelu4(pre_act(__ldg(x + px * g.C4 + q), noise ? __ldg(noise + px) : 0.f, w4, b4), g.slope);
  }
  for (int pb = p0 + rl; pb < p1; pb += UNR * rows) {      // UNR rows in flight per thread
    float4 xv[UNR];
    float nz[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int p = pb + u * rows;
      if (p < p1) {
        const int64_t px = (int64_t)n * g.HW + p;
        xv[u] = __ldg(x + px * g.C4 + q);
        nz[u] = noise ? __ldg(noise + px) : 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      if (pb + u * rows < p1) {
        float4 t = lrelu4(pre_act(xv[u], nz[u], w4, b4), g.slope);
        t.x -= k4.x; t.y -= k4.y; t.z -= k4.z; t.w -= k4.w;
        s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
        ss.x += t.x * t.x; ss.y += t.y * t.y; ss.z += t.z * t.z; ss.w += t.w * t.w;
      }
    }
  }
  block_rows_reduce(s, ss, red, g.C4, rows);
}

__global__ void __launch_bounds__(TPB) se_fwd_stats_kernel(const float4* __restrict__ x, const float* __restrict__ noise,
                                                           const float4* __restrict__ nw, const float4* __restrict__ bias,
                                                           float4* __restrict__ part, SE g) {
  __shared__ float4 red[2 * TPB];
  float4 s, ss, k4;
  const int n = blockIdx.y, chunk = blockIdx.x, q = threadIdx.x % g.C4;
  se_fwd_stats_sums(x, noise, nw, bias, g, red, n, chunk, -1, s, ss, k4);
  if (threadIdx.x < g.C4) {
    float4* o = part + ((int64_t)(n * g.chunks + chunk) * 3) * g.C4;
    o[q] = s;
    o[g.C4 + q] = ss;
    o[2 * g.C4 + q] = k4;
  }
}

// stats layout: [N][2][C] = (mu plane, rstd plane)
// block = 32 channels x 32 chunk lanes.  Every partial (sum, sum of squares about its own shift k_j) is re-based in fp64 to
// the shift of chunk 0 -- sum(t-K) = s_j + n_j d, sum((t-K)^2) = ss_j + 2 d s_j + n_j d^2 with d = k_j - K -- so the lanes
// only ADD (no divisions in the loop), then fold through shared memory.  K is a sample of the plane itself, so the final
// E[(t-K)^2] - E[t-K]^2 does not cancel (same argument as for the per-chunk shifts).
constexpr int FIN_CH = 32, FIN_KY = 32;

__global__ void __launch_bounds__(FIN_CH * FIN_KY) se_fwd_finalize_kernel(const float* __restrict__ part, float* __restrict__ stats,
                                                                          int N, int C, int chunks, int HW, int CHUNK, float eps) {
  __shared__ double sh[2][FIN_KY][FIN_CH + 1];
  const int cx = threadIdx.x % FIN_CH, ky = threadIdx.x / FIN_CH;
  const int c = blockIdx.x * FIN_CH + cx, n = blockIdx.y;
  double S = 0.0, SS = 0.0, K0 = 0.0;
  if (c < C) {
    K0 = (double)part[((int64_t)(n * chunks) * 3 + 2) * C + c];
    for (int k = ky; k < chunks; k += FIN_KY) {
      const float* p = part + ((int64_t)(n * chunks + k) * 3) * C;
      const int rem = HW - k * CHUNK;
      const double nk = rem < CHUNK ? rem : CHUNK;
      const double sk = (double)p[c], ssk = (double)p[C + c], d = (double)p[2 * C + c] - K0;
      S += sk + nk * d;
      SS += ssk + 2.0 * d * sk + nk * d * d;
    }
  }
  sh[0][ky][cx] = S; sh[1][ky][cx] = SS;
  __syncthreads();
  for (int o = FIN_KY / 2; o > 0; o >>= 1) {
    if (ky < o) { sh[0][ky][cx] += sh[0][ky + o][cx]; sh[1][ky][cx] += sh[1][ky + o][cx]; }
    __syncthreads();
  }
  if (ky == 0 && c < C) {
    const double m1 = sh[0][0][cx] / HW;
    double var = sh[1][0][cx] / HW - m1 * m1;
    if (var < 0.0) var = 0.0;
    stats[((int64_t)n * 2) * C + c] = (float)(K0 + m1);
    stats[((int64_t)n * 2 + 1) * C + c] = (float)(1.0 / sqrt(var + (double)eps));
  }
}

// ---- forward pass 2: normalise + modulate ------------------------------------------------------------
__device__ __forceinline__ void se_fwd_apply_body(const float4* __restrict__ x, const float* __restrict__ noise,
                                                  const float4* __restrict__ nw, const float4* __restrict__ bias,
                                                  const float4* __restrict__ style, const float4 mu, const float4 rs,
                                                  float4* __restrict__ out, const SE& g, int n, int chunk) {
  const int tid = threadIdx.x, q = tid % g.C4, rl = tid / g.C4, rows = TPB / g.C4;
  const int p0 = chunk * g.chunk, p1 = min(g.HW, p0 + g.chunk);
  const float4 w4 = nw ? __ldg(nw + q) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 b4 = bias ? __ldg(bias + q) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 ys = __ldg(style + (int64_t)n * 2 * g.C4 + q), yb = __ldg(style + ((int64_t)n * 2 + 1) * g.C4 + q);
  const float4 sc = make_float4(rs.x * (ys.x + 1.f), rs.y * (ys.y + 1.f), rs.z * (ys.z + 1.f), rs.w * (ys.w + 1.f));
  for (int pb = p0 + rl; pb < p1; pb += UNR * rows) {
    float4 xv[UNR];
    float nz[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int p = pb + u * rows;
      if (p < p1) {
        const int64_t px = (int64_t)n * g.HW + p;
        xv[u] = ldg_stream(x + px * g.C4 + q);
        nz[u] = noise ? __ldg(noise + px) : 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int p = pb + u * rows;
      if (p < p1) {
        const int64_t px = (int64_t)n * g.HW + p;
        const float4 t = lrelu4(pre_act(xv[u], nz[u], w4, b4), g.slope);
        stg_stream(out + px * g.C4 + q, make_float4((t.x - mu.x) * sc.x + yb.x, (t.y - mu.y) * sc.y + yb.y,
                                                    (t.z - mu.z) * sc.z + yb.z, (t.w - mu.w) * sc.w + yb.w));
      }
    }
  }
}

__global__ void __launch_bounds__(TPB) se_fwd_apply_kernel(const float4* __restrict__ x, const float* __restrict__ noise,
                                                           const float4* __restrict__ nw, const float4* __restrict__ bias,
                                                           const float4* __restrict__ style, const float4* __restrict__ stats,
                                                           float4* __restrict__ out, SE g) {
  const int n = blockIdx.y, q = threadIdx.x % g.C4;
  se_fwd_apply_body(x, noise, nw, bias, style, __ldg(stats + (int64_t)n * 2 * g.C4 + q), __ldg(stats + ((int64_t)n * 2 + 1) * g.C4 + q),
                    out, g, n, blockIdx.x);
}

// ---- backward pass 1: s1 = sum g, s2 = sum g*xhat per (n,c) chunk ------------------------------------------
// -> s1, s2 (valid for tid < C4)
__device__ __forceinline__ void se_bwd_stats_sums(const float4* __restrict__ gout, const float4* __restrict__ x,
                                                  const float* __restrict__ noise, const float4* __restrict__ nw,
                                                  const float4* __restrict__ bias, const float4* __restrict__ stats, const SE& g,
                                                  float4* red, int n, int chunk, float4& s1, float4& s2) {
  const int tid = threadIdx.x, q = tid % g.C4, rl = tid / g.C4, rows = TPB / g.C4;
  const int p0 = chunk * g.chunk, p1 = min(g.HW, p0 + g.chunk);
  const float4 w4 = nw ? __ldg(nw + q) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 b4 = bias ? __ldg(bias + q) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 mu = __ldg(stats + (int64_t)n * 2 * g.C4 + q), rs = __ldg(stats + ((int64_t)n * 2 + 1) * g.C4 + q);
  s1 = make_float4(0.f, 0.f, 0.f, 0.f); s2 = s1;
  for (int pb = p0 + rl; pb < p1; pb += UNR * rows) {
    float4 xv[UNR], gv[UNR];
    float nz[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int p = pb + u * rows;
      if (p < p1) {
        const int64_t px = (int64_t)n * g.HW + p;
        xv[u] = __ldg(x + px * g.C4 + q);
        gv[u] = __ldg(gout + px * g.C4 + q);
        nz[u] = noise ? __ldg(noise + px) : 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      if (pb + u * rows < p1) {
        const float4 t = lrelu4(pre_act(xv[u], nz[u], w4, b4), g.slope);
        const float4 go = gv[u];
        s1.x += go.x; s1.y += go.y; s1.z += go.z; s1.w += go.w;
        s2.x += go.x * (t.x - mu.x) * rs.x; s2.y += go.y * (t.y - mu.y) * rs.y;
        s2.z += go.z * (t.z - mu.z) * rs.z; s2.w += go.w * (t.w - mu.w) * rs.w;
      }
    }
  }
  block_rows_reduce(s1, s2, red, g.C4, rows);
}

__global__ void __launch_bounds__(TPB) se_bwd_stats_kernel(const float4* __restrict__ gout, const float4* __restrict__ x,
                                                           const float* __restrict__ noise, const float4* __restrict__ nw,
                                                           const float4* __restrict__ bias, const float4* __restrict__ stats,
                                                           float4* __restrict__ part, SE g) {
  __shared__ float4 red[2 * TPB];
  float4 s1, s2;
  const int n = blockIdx.y, chunk = blockIdx.x, q = threadIdx.x % g.C4;
  se_bwd_stats_sums(gout, x, noise, nw, bias, stats, g, red, n, chunk, s1, s2);
  if (threadIdx.x < g.C4) {
    float4* o = part + ((int64_t)(n * g.chunks + chunk) * 2) * g.C4;
    o[q] = s1;
    o[g.C4 + q] = s2;
  }
}

// gstyle[n][c] = s2 (d/d ys), gstyle[n][C+c] = s1 (d/d yb);  means[n][2][C] = ((ys+1)*s1/HW, (ys+1)*s2/HW)
__global__ void __launch_bounds__(FIN_CH * FIN_KY) se_bwd_finalize_kernel(const float* __restrict__ part, const float* __restrict__ style,
                                                                          float* __restrict__ gstyle, float* __restrict__ means, int N,
                                                                          int C, int chunks, int HW) {
  __shared__ double sh[2][FIN_KY][FIN_CH + 1];
  const int cx = threadIdx.x % FIN_CH, ky = threadIdx.x / FIN_CH;
  const int c = blockIdx.x * FIN_CH + cx, n = blockIdx.y;
  double s1 = 0.0, s2 = 0.0;
  if (c < C) {
    for (int k = ky; k < chunks; k += FIN_KY) {
      const float* p = part + ((int64_t)(n * chunks + k) * 2) * C;
      s1 += (double)p[c];
      s2 += (double)p[C + c];
    }
  }
  sh[0][ky][cx] = s1; sh[1][ky][cx] = s2;
  __syncthreads();
  for (int o = FIN_KY / 2; o > 0; o >>= 1) {
    if (ky < o) { sh[0][ky][cx] += sh[0][ky + o][cx]; sh[1][ky][cx] += sh[1][ky + o][cx]; }
    __syncthreads();
  }
  if (ky == 0 && c < C) {
    s1 = sh[0][0][cx]; s2 = sh[1][0][cx];
    gstyle[(int64_t)n * 2 * C + c] = (float)s2;
    gstyle[(int64_t)n * 2 * C + C + c] = (float)s1;
    const double sc = (double)style[(int64_t)n * 2 * C + c] + 1.0;
    means[(int64_t)n * 2 * C + c] = (float)(sc * s1 / HW);
    means[(int64_t)n * 2 * C + C + c] = (float)(sc * s2 / HW);
  }
}

// ---- backward pass 2: gx, and per-channel g_bias / g_noise_weight ------------------------------------------
__device__ __forceinline__ void se_bwd_apply_body(const float4* __restrict__ gout, const float4* __restrict__ x,
                                                  const float* __restrict__ noise, const float4* __restrict__ nw,
                                                  const float4* __restrict__ bias, const float4* __restrict__ style,
                                                  const float4* __restrict__ stats, const float4 m1, const float4 m2,
                                                  float4* __restrict__ gx, float* __restrict__ g_nw,
                                                  float* __restrict__ g_bias, const SE& g, float4* red, int n, int chunk) {
  const int tid = threadIdx.x, q = tid % g.C4, rl = tid / g.C4, rows = TPB / g.C4;
  const int p0 = chunk * g.chunk, p1 = min(g.HW, p0 + g.chunk);
  const float4 w4 = nw ? __ldg(nw + q) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 b4 = bias ? __ldg(bias + q) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 mu = __ldg(stats + (int64_t)n * 2 * g.C4 + q), rs = __ldg(stats + ((int64_t)n * 2 + 1) * g.C4 + q);
  const float4 ys = __ldg(style + (int64_t)n * 2 * g.C4 + q);
  const float4 sc = make_float4(ys.x + 1.f, ys.y + 1.f, ys.z + 1.f, ys.w + 1.f);
  float4 sb = make_float4(0.f, 0.f, 0.f, 0.f), sn = sb;
  for (int pb = p0 + rl; pb < p1; pb += UNR * rows) {
    float4 xv[UNR], gv[UNR];
    float nzv[UNR];
#pragma unroll
    for (int k = 0; k < UNR; ++k) {
      const int p = pb + k * rows;
      if (p < p1) {
        const int64_t px = (int64_t)n * g.HW + p;
        xv[k] = ldg_stream(x + px * g.C4 + q);
        gv[k] = ldg_stream(gout + px * g.C4 + q);
        nzv[k] = noise ? __ldg(noise + px) : 0.f;
      }
    }
#pragma unroll
    for (int k = 0; k < UNR; ++k) {
      const int p = pb + k * rows;
      if (p < p1) {
        const int64_t px = (int64_t)n * g.HW + p;
        const float nz = nzv[k];
        const float4 u = pre_act(xv[k], nz, w4, b4);
        const float4 t = lrelu4(u, g.slope);
        const float4 go = gv[k];
        float4 gu;
        gu.x = rs.x * (sc.x * go.x - m1.x - (t.x - mu.x) * rs.x * m2.x) * (u.x > 0.f ? 1.f : g.slope);
        gu.y = rs.y * (sc.y * go.y - m1.y - (t.y - mu.y) * rs.y * m2.y) * (u.y > 0.f ? 1.f : g.slope);
        gu.z = rs.z * (sc.z * go.z - m1.z - (t.z - mu.z) * rs.z * m2.z) * (u.z > 0.f ? 1.f : g.slope);
        gu.w = rs.w * (sc.w * go.w - m1.w - (t.w - mu.w) * rs.w * m2.w) * (u.w > 0.f ? 1.f : g.slope);
        stg_stream(gx + px * g.C4 + q, gu);
        sb.x += gu.x; sb.y += gu.y; sb.z += gu.z; sb.w += gu.w;
        sn.x += gu.x * nz; sn.y += gu.y * nz; sn.z += gu.z * nz; sn.w += gu.w * nz;
      }
    }
  }
  block_rows_reduce(sb, sn, red, g.C4, rows);
  if (tid < g.C4) {
    if (g_bias) {
      atomicAdd(g_bias + 4 * q + 0, sb.x); atomicAdd(g_bias + 4 * q + 1, sb.y);
      atomicAdd(g_bias + 4 * q + 2, sb.z); atomicAdd(g_bias + 4 * q + 3, sb.w);
    }
    if (g_nw) {
      atomicAdd(g_nw + 4 * q + 0, sn.x); atomicAdd(g_nw + 4 * q + 1, sn.y);
      atomicAdd(g_nw + 4 * q + 2, sn.z); atomicAdd(g_nw + 4 * q + 3, sn.w);
    }
  }
}

__global__ void __launch_bounds__(TPB) se_bwd_apply_kernel(const float4* __restrict__ gout, const float4* __restrict__ x,
                                                           const float* __restrict__ noise, const float4* __restrict__ nw,
                                                           const float4* __restrict__ bias, const float4* __restrict__ style,
                                                           const float4* __restrict__ stats, const float4* __restrict__ means,
                                                           float4* __restrict__ gx, float* __restrict__ g_nw,
                                                           float* __restrict__ g_bias, SE g) {
  __shared__ float4 red[2 * TPB];
  const int n = blockIdx.y, q = threadIdx.x % g.C4;
  se_bwd_apply_body(gout, x, noise, nw, bias, style, stats, __ldg(means + (int64_t)n * 2 * g.C4 + q),
                    __ldg(means + ((int64_t)n * 2 + 1) * g.C4 + q), gx, g_nw, g_bias, g, red, n, blockIdx.x);
}

// ================================================================================================ one launch, sample by sample
// The passes as ONE kernel whose blocks are ordered  [stats(n=0) | apply(n=0) | stats(n=1) | apply(n=1) | ...]:
// every statistics block adds its partial sums (fp64 atomics; all blocks of a sample use the same shift = the sample's first
// pixel) to the sample's accumulators and bumps the sample's counter; the apply blocks of that sample wait until the counter
// shows all of them and derive mean / rstd (backward: the two means) themselves.  Blocks are dispatched in index order, so
// every statistics block of a sample is resident or done before any of its apply blocks holds an SM slot -- no deadlock -- and
// the apply pass re-reads a sample's 1-16 MB a few microseconds after the statistics pass touched them: from L2, not from HBM
// (launched as whole-tensor passes the re-read came ~70 MB later and missed).  HBM traffic = the algorithmic 8E / 12E.
__device__ __forceinline__ void se_signal(int* counter) {
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) atomicAdd(counter, 1);
}

__device__ __forceinline__ void se_wait(int* counter, int target) {
  if (threadIdx.x == 0) {
    while (atomicAdd(counter, 0) < target) __nanosleep(32);
    __threadfence();
  }
  __syncthreads();
}

constexpr int SEK = 8;   // accumulator copies per sample: ~150 blocks hitting ONE address serialise (~90 ns per atomic at L2)

__device__ __forceinline__ void atomic_add4(double* p, const float4& v) {
  atomicAdd(p + 0, (double)v.x); atomicAdd(p + 1, (double)v.y); atomicAdd(p + 2, (double)v.z); atomicAdd(p + 3, (double)v.w);
}

// acc: [N][SEK][2][C] doubles, zeroed before the launch; counter: [N] ints, zeroed
__global__ void __launch_bounds__(TPB) se_fwd_fused_kernel(const float4* __restrict__ x, const float* __restrict__ noise,
                                                           const float4* __restrict__ nw, const float4* __restrict__ bias,
                                                           const float4* __restrict__ style, double* acc, float* __restrict__ stats,
                                                           float4* __restrict__ out, int* counter, SE g, float eps) {
  __shared__ float4 red[2 * TPB];
  const int per = 2 * g.chunks;
  const int n = blockIdx.x / per, r = blockIdx.x - n * per;
  const int C = 4 * g.C4, q = threadIdx.x % g.C4;
  double* an = acc + (int64_t)n * SEK * 2 * C;
  if (r < g.chunks) {
    float4 s, ss, k4;
    se_fwd_stats_sums(x, noise, nw, bias, g, red, n, r, 0, s, ss, k4);
    if (threadIdx.x < g.C4) {
      double* ak = an + (int64_t)(r % SEK) * 2 * C;
      atomic_add4(ak + 4 * q, s);
      atomic_add4(ak + C + 4 * q, ss);
    }
    se_signal(counter + n);
  } else {
    // the shift every statistics block of this sample used: t at the sample's first pixel
    const float4 w4 = nw ? __ldg(nw + q) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 b4 = bias ? __ldg(bias + q) : make_float4(0.f, 0.f, 0.f, 0.f);
    const int64_t px = (int64_t)n * g.HW;
    const float4 k4 = lrelu4(pre_act(__ldg(x + px * g.C4 + q), noise ? __ldg(noise + px) : 0.f, w4, b4), g.slope);
    se_wait(counter + n, g.chunks);
    const double kk[4] = {k4.x, k4.y, k4.z, k4.w};
    float mu[4], rs[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      double S = 0.0, SS = 0.0;
#pragma unroll
      for (int k = 0; k < SEK; ++k) {
        S += __ldcg(an + (int64_t)k * 2 * C + 4 * q + e);
        SS += __ldcg(an + (int64_t)k * 2 * C + C + 4 * q + e);
      }
      const double m1 = S / g.HW;
      double var = SS / g.HW - m1 * m1;
      if (var < 0.0) var = 0.0;
      mu[e] = (float)(kk[e] + m1);
      rs[e] = (float)(1.0 / sqrt(var + (double)eps));
    }
    if (r == g.chunks && threadIdx.x < g.C4) {          // one block per sample publishes (mean, rstd) for the backward pass
      *reinterpret_cast<float4*>(stats + (int64_t)n * 2 * C + 4 * q) = make_float4(mu[0], mu[1], mu[2], mu[3]);
      *reinterpret_cast<float4*>(stats + (int64_t)n * 2 * C + C + 4 * q) = make_float4(rs[0], rs[1], rs[2], rs[3]);
    }
    se_fwd_apply_body(x, noise, nw, bias, style, make_float4(mu[0], mu[1], mu[2], mu[3]), make_float4(rs[0], rs[1], rs[2], rs[3]), out,
                      g, n, r - g.chunks);
  }
}

__global__ void __launch_bounds__(TPB) se_bwd_fused_kernel(const float4* __restrict__ gout, const float4* __restrict__ x,
                                                           const float* __restrict__ noise, const float4* __restrict__ nw,
                                                           const float4* __restrict__ bias, const float* __restrict__ style,
                                                           const float4* __restrict__ stats, double* acc, float4* __restrict__ gx,
                                                           float* __restrict__ gstyle, float* __restrict__ g_nw,
                                                           float* __restrict__ g_bias, int* counter, SE g) {
  __shared__ float4 red[2 * TPB];
  const int per = 2 * g.chunks;
  const int n = blockIdx.x / per, r = blockIdx.x - n * per;
  const int C = 4 * g.C4, q = threadIdx.x % g.C4;
  double* an = acc + (int64_t)n * SEK * 2 * C;
  if (r < g.chunks) {
    float4 s1, s2;
    se_bwd_stats_sums(gout, x, noise, nw, bias, stats, g, red, n, r, s1, s2);
    if (threadIdx.x < g.C4) {
      double* ak = an + (int64_t)(r % SEK) * 2 * C;
      atomic_add4(ak + 4 * q, s1);
      atomic_add4(ak + C + 4 * q, s2);
    }
    se_signal(counter + n);
  } else {
    const float* sty = style + (int64_t)n * 2 * C + 4 * q;
    const double sc[4] = {(double)__ldg(sty) + 1.0, (double)__ldg(sty + 1) + 1.0, (double)__ldg(sty + 2) + 1.0, (double)__ldg(sty + 3) + 1.0};
    se_wait(counter + n, g.chunks);
    float m1[4], m2[4], S1[4], S2[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      double a1 = 0.0, a2 = 0.0;
#pragma unroll
      for (int k = 0; k < SEK; ++k) {
        a1 += __ldcg(an + (int64_t)k * 2 * C + 4 * q + e);
        a2 += __ldcg(an + (int64_t)k * 2 * C + C + 4 * q + e);
      }
      S1[e] = (float)a1; S2[e] = (float)a2;
      m1[e] = (float)(sc[e] * a1 / g.HW);
      m2[e] = (float)(sc[e] * a2 / g.HW);
    }
    if (r == g.chunks && threadIdx.x < g.C4) {          // gstyle[n][c] = s2 (d/d ys), gstyle[n][C+c] = s1 (d/d yb)
      *reinterpret_cast<float4*>(gstyle + (int64_t)n * 2 * C + 4 * q) = make_float4(S2[0], S2[1], S2[2], S2[3]);
      *reinterpret_cast<float4*>(gstyle + (int64_t)n * 2 * C + C + 4 * q) = make_float4(S1[0], S1[1], S1[2], S1[3]);
    }
    se_bwd_apply_body(gout, x, noise, nw, bias, (const float4*)style, stats, make_float4(m1[0], m1[1], m1[2], m1[3]),
                      make_float4(m2[0], m2[1], m2[2], m2[3]), gx, g_nw, g_bias, g, red, n, r - g.chunks);
  }
}

// ================================================================================================ single-read cluster kernels
// One thread-block CLUSTER per (sample, 16-channel slab): its CTAs split the H*W pixels, every CTA keeps its part of the slab
// in shared memory, the plane statistics are combined through distributed shared memory, and the normalisation runs on the
// copy in shared memory -- the activation is read from HBM ONCE (forward 8E, backward 12E bytes: the algorithmic minimum;
// the three-kernel path above moves 12E / 20E).  16 channels = 64-byte pixel segments (two full 32-byte sectors);
// 512 threads, 8 independent 16-byte loads in flight per thread (64 KB per SM) cover the HBM latency.
// Backward keeps the incoming gradient in shared memory and reads x a second time -- a few tens of microseconds after the
// first read, with < 40 MB streamed in between chip-wide, i.e. from the 126 MB L2.
constexpr int CTPB = 512;          // threads per CTA
constexpr int CS = 16;             // channels per slab
constexpr int CQ = CS / 4;         // float4 per pixel of a slab
constexpr int CRPP = CTPB / CQ;    // pixels per pass of the block
constexpr int CUNR = 8;

struct SEC {
  int N, HW, C, P;                 // P = pixels per CTA = HW / cluster size
  float slope, eps;
};

// per-channel totals of the block: thread (row, q) holds float4 a, b for channels 4q..4q+3 -> tot[0][c], tot[1][c] (double)
__device__ __forceinline__ void cta_channel_sums(const float4& a, const float4& b, double (*wred)[2][CS], double (*tot)[CS]) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);   // lanes with the same q = lane % 4
  }
  if (lane < CQ) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      wred[warp][0][lane * 4 + e] = v[e];
      wred[warp][1][lane * 4 + e] = v[4 + e];
    }
  }
  __syncthreads();
  if (tid < 2 * CS) {
    const int which = tid / CS, c = tid % CS;
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < CTPB / 32; ++w) t += wred[w][which][c];
    tot[which][c] = t;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(CTPB) se_fwd_cluster_kernel(const float4* __restrict__ x, const float* __restrict__ noise,
                                                              const float4* __restrict__ nw, const float4* __restrict__ bias,
                                                              const float* __restrict__ style, float4* __restrict__ out,
                                                              float* __restrict__ stats, SEC g) {
  extern __shared__ float4 ts[];                       // [P][CQ]: post-activation values of this CTA's part of the slab
  __shared__ double wred[CTPB / 32][2][CS];
  __shared__ double part[3][CS];                       // sum(t - k), sum((t - k)^2), k per channel; read by the cluster
  __shared__ float mu_s[CS], rs_s[CS];
  cg::cluster_group cluster = cg::this_cluster();
  const int CL = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
  const int tid = threadIdx.x, q = tid % CQ, row = tid / CQ;
  const int slab = blockIdx.x / CL, n = blockIdx.y;
  const int C4 = g.C / 4, p0 = rank * g.P;
  const float4* xg = x + ((int64_t)n * g.HW + p0) * C4 + slab * CQ + q;
  const float* nzg = noise ? noise + (int64_t)n * g.HW + p0 : nullptr;
  const float4 w4 = nw ? __ldg(nw + slab * CQ + q) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 b4 = bias ? __ldg(bias + slab * CQ + q) : make_float4(0.f, 0.f, 0.f, 0.f);
  // shift = this CTA's first pixel (sums about a sample of the plane do not cancel when |mean| >> std)
  const float4 k4 = lrelu4(pre_act(__ldg(xg), nzg ? __ldg(nzg) : 0.f, w4, b4), g.slope);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), ss = s;
  for (int pb = row; pb < g.P; pb += CUNR * CRPP) {
    float4 xv[CUNR];
    float nz[CUNR];
#pragma unroll
    for (int u = 0; u < CUNR; ++u) {
      const int p = pb + u * CRPP;
      if (p < g.P) {
        xv[u] = ldg_stream(xg + (int64_t)p * C4);
        nz[u] = nzg ? __ldg(nzg + p) : 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < CUNR; ++u) {
      const int p = pb + u * CRPP;
      if (p < g.P) {
        const float4 t = lrelu4(pre_act(xv[u], nz[u], w4, b4), g.slope);
        ts[p * CQ + q] = t;
        const float4 d = make_float4(t.x - k4.x, t.y - k4.y, t.z - k4.z, t.w - k4.w);
        s.x += d.x; s.y += d.y; s.z += d.z; s.w += d.w;
        ss.x += d.x * d.x; ss.y += d.y * d.y; ss.z += d.z * d.z; ss.w += d.w * d.w;
      }
    }
  }
  cta_channel_sums(s, ss, wred, part);
  if (row == 0) { part[2][4 * q + 0] = k4.x; part[2][4 * q + 1] = k4.y; part[2][4 * q + 2] = k4.z; part[2][4 * q + 3] = k4.w; }
  cluster.sync();
  if (tid < CS) {
    const double* p0r = cluster.map_shared_rank(&part[0][0], 0);
    const double K0 = p0r[2 * CS + tid];
    double S = 0.0, SS = 0.0;
    for (int r = 0; r < CL; ++r) {
      const double* pr = cluster.map_shared_rank(&part[0][0], r);
      const double sk = pr[tid], ssk = pr[CS + tid], d = pr[2 * CS + tid] - K0;
      S += sk + (double)g.P * d;
      SS += ssk + 2.0 * d * sk + (double)g.P * d * d;
    }
    const double m1 = S / g.HW;
    double var = SS / g.HW - m1 * m1;
    if (var < 0.0) var = 0.0;
    const float mu = (float)(K0 + m1), rs = (float)(1.0 / sqrt(var + (double)g.eps));
    mu_s[tid] = mu; rs_s[tid] = rs;
    if (rank == 0) {
      stats[((int64_t)n * 2) * g.C + slab * CS + tid] = mu;
      stats[((int64_t)n * 2 + 1) * g.C + slab * CS + tid] = rs;
    }
  }
  __syncthreads();
  const float* st = style + (int64_t)n * 2 * g.C + slab * CS + 4 * q;
  float4 mu4, sc4, yb4;
  mu4 = make_float4(mu_s[4 * q], mu_s[4 * q + 1], mu_s[4 * q + 2], mu_s[4 * q + 3]);
  sc4 = make_float4(rs_s[4 * q] * (__ldg(st) + 1.f), rs_s[4 * q + 1] * (__ldg(st + 1) + 1.f), rs_s[4 * q + 2] * (__ldg(st + 2) + 1.f),
                    rs_s[4 * q + 3] * (__ldg(st + 3) + 1.f));
  yb4 = make_float4(__ldg(st + g.C), __ldg(st + g.C + 1), __ldg(st + g.C + 2), __ldg(st + g.C + 3));
  float4* og = out + ((int64_t)n * g.HW + p0) * C4 + slab * CQ + q;
#pragma unroll 4
  for (int p = row; p < g.P; p += CRPP) {
    const float4 t = ts[p * CQ + q];
    stg_stream(og + (int64_t)p * C4, make_float4((t.x - mu4.x) * sc4.x + yb4.x, (t.y - mu4.y) * sc4.y + yb4.y,
                                                 (t.z - mu4.z) * sc4.z + yb4.z, (t.w - mu4.w) * sc4.w + yb4.w));
  }
  cluster.sync();                                      // peers may still be reading this CTA's partial sums
}

__global__ void __launch_bounds__(CTPB) se_bwd_cluster_kernel(const float4* __restrict__ gout, const float4* __restrict__ x,
                                                              const float* __restrict__ noise, const float4* __restrict__ nw,
                                                              const float4* __restrict__ bias, const float* __restrict__ style,
                                                              const float* __restrict__ stats, float4* __restrict__ gx,
                                                              float* __restrict__ gstyle, float* __restrict__ g_nw,
                                                              float* __restrict__ g_bias, SEC g) {
  extern __shared__ float4 gs[];                       // [P][CQ]: incoming gradient of this CTA's part of the slab
  __shared__ double wred[CTPB / 32][2][CS];
  __shared__ double part[2][CS];                       // sum g, sum g*xhat per channel; read by the cluster
  __shared__ float m1_s[CS], m2_s[CS];
  cg::cluster_group cluster = cg::this_cluster();
  const int CL = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
  const int tid = threadIdx.x, q = tid % CQ, row = tid / CQ;
  const int slab = blockIdx.x / CL, n = blockIdx.y;
  const int C4 = g.C / 4, p0 = rank * g.P;
  const int64_t off = ((int64_t)n * g.HW + p0) * C4 + slab * CQ + q;
  const float4* xg = x + off;
  const float4* gg = gout + off;
  const float* nzg = noise ? noise + (int64_t)n * g.HW + p0 : nullptr;
  const float4 w4 = nw ? __ldg(nw + slab * CQ + q) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 b4 = bias ? __ldg(bias + slab * CQ + q) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float* stp = stats + (int64_t)n * 2 * g.C + slab * CS + 4 * q;
  const float4 mu = make_float4(__ldg(stp), __ldg(stp + 1), __ldg(stp + 2), __ldg(stp + 3));
  const float4 rs = make_float4(__ldg(stp + g.C), __ldg(stp + g.C + 1), __ldg(stp + g.C + 2), __ldg(stp + g.C + 3));
  const float* sty = style + (int64_t)n * 2 * g.C + slab * CS + 4 * q;
  const float4 sc = make_float4(__ldg(sty) + 1.f, __ldg(sty + 1) + 1.f, __ldg(sty + 2) + 1.f, __ldg(sty + 3) + 1.f);
  constexpr int U1 = CUNR / 2;                         // two streams in pass 1: 4 + 4 loads in flight per thread
  float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
  for (int pb = row; pb < g.P; pb += U1 * CRPP) {
    float4 xv[U1], gv[U1];
    float nz[U1];
#pragma unroll
    for (int u = 0; u < U1; ++u) {
      const int p = pb + u * CRPP;
      if (p < g.P) {
        xv[u] = __ldg(xg + (int64_t)p * C4);           // read again in pass 2: keep it cacheable
        gv[u] = ldg_stream(gg + (int64_t)p * C4);
        nz[u] = nzg ? __ldg(nzg + p) : 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < U1; ++u) {
      const int p = pb + u * CRPP;
      if (p < g.P) {
        const float4 t = lrelu4(pre_act(xv[u], nz[u], w4, b4), g.slope);
        const float4 go = gv[u];
        gs[p * CQ + q] = go;
        s1.x += go.x; s1.y += go.y; s1.z += go.z; s1.w += go.w;
        s2.x += go.x * (t.x - mu.x) * rs.x; s2.y += go.y * (t.y - mu.y) * rs.y;
        s2.z += go.z * (t.z - mu.z) * rs.z; s2.w += go.w * (t.w - mu.w) * rs.w;
      }
    }
  }
  cta_channel_sums(s1, s2, wred, part);
  cluster.sync();
  if (tid < CS) {
    double S1 = 0.0, S2 = 0.0;
    for (int r = 0; r < CL; ++r) {
      const double* pr = cluster.map_shared_rank(&part[0][0], r);
      S1 += pr[tid]; S2 += pr[CS + tid];
    }
    const int c = slab * CS + tid;
    if (rank == 0) {
      gstyle[(int64_t)n * 2 * g.C + c] = (float)S2;
      gstyle[(int64_t)n * 2 * g.C + g.C + c] = (float)S1;
    }
    const double scd = (double)__ldg(style + (int64_t)n * 2 * g.C + c) + 1.0;
    m1_s[tid] = (float)(scd * S1 / g.HW);
    m2_s[tid] = (float)(scd * S2 / g.HW);
  }
  __syncthreads();
  const float4 m1 = make_float4(m1_s[4 * q], m1_s[4 * q + 1], m1_s[4 * q + 2], m1_s[4 * q + 3]);
  const float4 m2 = make_float4(m2_s[4 * q], m2_s[4 * q + 1], m2_s[4 * q + 2], m2_s[4 * q + 3]);
  float4* gxg = gx + off;
  float4 sb = make_float4(0.f, 0.f, 0.f, 0.f), sn = sb;
  for (int pb = row; pb < g.P; pb += CUNR * CRPP) {
    float4 xv[CUNR];
    float nzv[CUNR];
#pragma unroll
    for (int u = 0; u < CUNR; ++u) {
      const int p = pb + u * CRPP;
      if (p < g.P) {
        xv[u] = ldg_stream(xg + (int64_t)p * C4);
        nzv[u] = nzg ? __ldg(nzg + p) : 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < CUNR; ++u) {
      const int p = pb + u * CRPP;
      if (p < g.P) {
        const float nz = nzv[u];
        const float4 uu = pre_act(xv[u], nz, w4, b4);
        const float4 t = lrelu4(uu, g.slope);
        const float4 go = gs[p * CQ + q];
        float4 gu;
        gu.x = rs.x * (sc.x * go.x - m1.x - (t.x - mu.x) * rs.x * m2.x) * (uu.x > 0.f ? 1.f : g.slope);
        gu.y = rs.y * (sc.y * go.y - m1.y - (t.y - mu.y) * rs.y * m2.y) * (uu.y > 0.f ? 1.f : g.slope);
        gu.z = rs.z * (sc.z * go.z - m1.z - (t.z - mu.z) * rs.z * m2.z) * (uu.z > 0.f ? 1.f : g.slope);
        gu.w = rs.w * (sc.w * go.w - m1.w - (t.w - mu.w) * rs.w * m2.w) * (uu.w > 0.f ? 1.f : g.slope);
        stg_stream(gxg + (int64_t)p * C4, gu);
        sb.x += gu.x; sb.y += gu.y; sb.z += gu.z; sb.w += gu.w;
        sn.x += gu.x * nz; sn.y += gu.y * nz; sn.z += gu.z * nz; sn.w += gu.w * nz;
      }
    }
  }
  if (g_bias != nullptr || g_nw != nullptr) {
    cta_channel_sums(sb, sn, wred, part);              // (part is no longer read by the cluster: all peers are past their reads
    if (tid < CS) {                                    //  only after the final cluster.sync -- so use a private copy instead)
      if (g_bias) atomicAdd(g_bias + slab * CS + tid, (float)part[0][tid]);
      if (g_nw) atomicAdd(g_nw + slab * CS + tid, (float)part[1][tid]);
    }
  }
  cluster.sync();
}

// cluster size for an H*W plane: the smallest power of two <= 8 that brings a CTA's part of a 16-channel slab to <= 64 KB
// (two CTAs per SM), else 8 with up to 200 KB; 0 = the plane does not fit (three-kernel path)
int sec_cluster(int HW) {
  int max_cl = 8;
  if (const char* e = getenv("GLB_SE_CL16")) if (atoi(e) != 0) max_cl = 16;   // non-portable cluster size (A/B experiment)
  for (int cl = 1; cl <= max_cl; cl <<= 1)
    if (HW % cl == 0 && (int64_t)(HW / cl) * CS * 4 <= 64 * 1024) return cl;
  if (HW % 8 == 0 && (int64_t)(HW / 8) * CS * 4 <= 200 * 1024) return 8;
  return 0;
}

template <typename Kern, typename... Args>
int sec_launch(Kern kern, const char* name, int N, int HW, int C, int CL, cudaStream_t st, Args... args) {
  const size_t smem = (size_t)(HW / CL) * CS * 4;
  static size_t configured = 0;                        // per kernel instantiation (Kern differs)
  if (smem > configured) {
    GLB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    GLB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    configured = 200 * 1024;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((C / CS) * CL, N, 1);
  cfg.blockDim = dim3(CTPB, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, args...);
  if (e != cudaSuccess) return cuda_fail(e, name);
  return GLB_OK;
}

// pixels per block: aim at ~8 blocks of 256 threads per SM over the whole (N, HW) range (full occupancy; with UNR rows in
// flight per thread that is what a two-input streaming pass needs to approach the HBM roofline), a multiple of the rows a
// block covers per iteration, at most MAX_CHUNK (low-resolution layers otherwise run on 8 blocks)
int se_chunk(int N, int HW, int C) {
  const int rows = TPB / (C / 4) > 0 ? TPB / (C / 4) : 1;
  int64_t chunk = ((int64_t)N * HW + 8 * kNumSMs - 1) / (8 * kNumSMs);
  chunk = ((chunk + rows - 1) / rows) * rows;
  if (chunk > MAX_CHUNK) chunk = MAX_CHUNK;
  if (chunk < rows) chunk = rows;
  return (int)chunk;
}

// GLB_SE_MODE: 0 = three kernels, 1 = one sample-ordered kernel, 2 = cluster kernels wherever the plane fits, 3 (default) = per
// size, whichever measured faster on B200 (tools/glue_bw.py, N = 8): cluster kernels for planes up to 32 x 32 forward / 64 x 64
// backward, three kernels above.  The parity tests run modes 0-2 on every shape.
int se_mode() {
  if (const char* e = getenv("GLB_SE_MODE")) return atoi(e);
  return 3;
}
bool se_use_cluster(int HW, bool backward) {
  const int m = se_mode();
  if (m == 2) return true;
  if (m != 3) return false;
  if (const char* e = getenv("GLB_SE_CL16")) if (atoi(e) != 0) return true;     // A/B: 16-CTA clusters make every cfg2 plane fit 64 KB
  return HW <= (backward ? 4096 : 1024);
}

int se_geom(SE& g, int N, int H, int W, int C, float slope) {
  if (C % 4 != 0 || C / 4 > TPB || TPB % (C / 4) != 0) return shape_fail("style_epilogue: C/4 must divide 256");
  g.N = N; g.HW = H * W; g.C4 = C / 4; g.chunk = se_chunk(N, g.HW, C); g.chunks = (g.HW + g.chunk - 1) / g.chunk; g.slope = slope;
  if (N > 65535) return shape_fail("style_epilogue: N > 65535");
  return GLB_OK;
}

}  // namespace
}  // namespace glb

using namespace glb;

namespace glb {
namespace {
// floats of the partial sums + means of the three-kernel path; the sample-ordered kernel uses the SAME buffer as
// [N][2][C] doubles + [N] ints (8-byte aligned from the buffer's start)
int64_t se_part_floats(int N, int H, int W, int C) {
  const int chunk = se_chunk(N, H * W, C);
  const int64_t chunks = ((int64_t)H * W + chunk - 1) / chunk;
  return (int64_t)N * chunks * 3 * C + (int64_t)N * 2 * C;
}
double* se_fused_acc(float* work) {
  return reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(work) + 7) & ~(uintptr_t)7);
}
}  // namespace
}  // namespace glb

extern "C" int64_t glb_style_epilogue_work_floats(int N, int H, int W, int C) {
  if (C < 4) return 0;
  const int64_t fused = 4 * (int64_t)N * C * 8 + (int64_t)N + 16;  // [N][SEK = 8][2][C] doubles + [N] ints (+ alignment slack), in floats
  const int64_t three = glb::se_part_floats(N, H, W, C);
  return three > fused ? three : fused;
}

extern "C" int glb_style_epilogue_fwd(const float* x, const float* noise, const float* noise_weight, const float* bias,
                                      const float* style, float* out, float* stats, float* work, int N, int H, int W, int C,
                                      float slope, float eps, glb_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (const int CL = (C % CS == 0 && N <= 65535 && se_use_cluster(H * W, false)) ? sec_cluster(H * W) : 0) {
    SEC c{N, H * W, C, H * W / CL, slope, eps};
    return sec_launch(se_fwd_cluster_kernel, "se_fwd_cluster_kernel", N, H * W, C, CL, st, (const float4*)x, noise,
                      (const float4*)noise_weight, (const float4*)bias, style, (float4*)out, stats, c);
  }
  SE g;
  if (int rc = se_geom(g, N, H, W, C, slope)) return rc;
  if (se_mode() == 1) {
    double* acc = se_fused_acc(work);
    int* counter = reinterpret_cast<int*>(acc + (int64_t)N * SEK * 2 * C);
    GLB_CUDA(cudaMemsetAsync(acc, 0, sizeof(double) * N * SEK * 2 * C + sizeof(int) * N, st));
    se_fwd_fused_kernel<<<N * 2 * g.chunks, TPB, 0, st>>>((const float4*)x, noise, (const float4*)noise_weight, (const float4*)bias,
                                                         (const float4*)style, acc, stats, (float4*)out, counter, g, eps);
    GLB_CHECK_LAUNCH("se_fwd_fused");
    return GLB_OK;
  }
  dim3 grid(g.chunks, N);
  se_fwd_stats_kernel<<<grid, TPB, 0, st>>>((const float4*)x, noise, (const float4*)noise_weight, (const float4*)bias, (float4*)work, g);
  GLB_CHECK_LAUNCH("se_fwd_stats");
  se_fwd_finalize_kernel<<<dim3((C + FIN_CH - 1) / FIN_CH, N), FIN_CH * FIN_KY, 0, st>>>(work, stats, N, C, g.chunks, g.HW, g.chunk, eps);
  GLB_CHECK_LAUNCH("se_fwd_finalize");
  se_fwd_apply_kernel<<<grid, TPB, 0, st>>>((const float4*)x, noise, (const float4*)noise_weight, (const float4*)bias,
                                            (const float4*)style, (const float4*)stats, (float4*)out, g);
  GLB_CHECK_LAUNCH("se_fwd_apply");
  return GLB_OK;
}

extern "C" int glb_style_epilogue_bwd(const float* gout, const float* x, const float* noise, const float* noise_weight,
                                      const float* bias, const float* style, const float* stats, float* gx, float* gstyle,
                                      float* g_noise_weight, float* g_bias, float* work, int N, int H, int W, int C, float slope,
                                      glb_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (const int CL = (C % CS == 0 && N <= 65535 && se_use_cluster(H * W, true)) ? sec_cluster(H * W) : 0) {
    SEC c{N, H * W, C, H * W / CL, slope, 0.f};
    return sec_launch(se_bwd_cluster_kernel, "se_bwd_cluster_kernel", N, H * W, C, CL, st, (const float4*)gout, (const float4*)x,
                      noise, (const float4*)noise_weight, (const float4*)bias, style, stats, (float4*)gx, gstyle, g_noise_weight,
                      g_bias, c);
  }
  SE g;
  if (int rc = se_geom(g, N, H, W, C, slope)) return rc;
  dim3 grid(g.chunks, N);
  float* means = work + (int64_t)N * g.chunks * 2 * C;
  if (se_mode() == 1) {
    double* acc = se_fused_acc(work);
    int* counter = reinterpret_cast<int*>(acc + (int64_t)N * SEK * 2 * C);
    GLB_CUDA(cudaMemsetAsync(acc, 0, sizeof(double) * N * SEK * 2 * C + sizeof(int) * N, st));
    se_bwd_fused_kernel<<<N * 2 * g.chunks, TPB, 0, st>>>((const float4*)gout, (const float4*)x, noise, (const float4*)noise_weight,
                                                         (const float4*)bias, style, (const float4*)stats, acc, (float4*)gx, gstyle,
                                                         g_noise_weight, g_bias, counter, g);
    GLB_CHECK_LAUNCH("se_bwd_fused");
    return GLB_OK;
  }
  se_bwd_stats_kernel<<<grid, TPB, 0, st>>>((const float4*)gout, (const float4*)x, noise, (const float4*)noise_weight,
                                            (const float4*)bias, (const float4*)stats, (float4*)work, g);
  GLB_CHECK_LAUNCH("se_bwd_stats");
  se_bwd_finalize_kernel<<<dim3((C + FIN_CH - 1) / FIN_CH, N), FIN_CH * FIN_KY, 0, st>>>(work, style, gstyle, means, N, C, g.chunks, g.HW);
  GLB_CHECK_LAUNCH("se_bwd_finalize");
  se_bwd_apply_kernel<<<grid, TPB, 0, st>>>((const float4*)gout, (const float4*)x, noise, (const float4*)noise_weight,
                                            (const float4*)bias, (const float4*)style, (const float4*)stats, (const float4*)means,
                                            (float4*)gx, g_noise_weight, g_bias, g);
  GLB_CHECK_LAUNCH("se_bwd_apply");
  return GLB_OK;
}
