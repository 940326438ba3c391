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
#include "common.cuh"

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
__global__ void __launch_bounds__(TPB) se_fwd_stats_kernel(const float4* __restrict__ x, const float* __restrict__ noise,
                                                           const float4* __restrict__ nw, const float4* __restrict__ bias,
                                                           float4* __restrict__ part, SE g) {
  __shared__ float4 red[2 * TPB];
  const int tid = threadIdx.x, q = tid % g.C4, rl = tid / g.C4, rows = TPB / g.C4;
  const int n = blockIdx.y, chunk = blockIdx.x;
  const int p0 = chunk * g.chunk, p1 = min(g.HW, p0 + g.chunk);
  const float4 w4 = nw ? __ldg(nw + q) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 b4 = bias ? __ldg(bias + q) : make_float4(0.f, 0.f, 0.f, 0.f);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), ss = s;
  // shifted sums: k4 = the chunk's first pixel, so sum((t-k)^2) - sum(t-k)^2/n does not cancel when |mean| >> std
  // (e.g. the constant-input layer, whose planes are 1 + small noise)
  float4 k4;
  {
    const int64_t px = (int64_t)n * g.HW + p0;
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
  if (tid < g.C4) {
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
__global__ void __launch_bounds__(TPB) se_fwd_apply_kernel(const float4* __restrict__ x, const float* __restrict__ noise,
                                                           const float4* __restrict__ nw, const float4* __restrict__ bias,
                                                           const float4* __restrict__ style, const float4* __restrict__ stats,
                                                           float4* __restrict__ out, SE g) {
  const int tid = threadIdx.x, q = tid % g.C4, rl = tid / g.C4, rows = TPB / g.C4;
  const int n = blockIdx.y, chunk = blockIdx.x;
  const int p0 = chunk * g.chunk, p1 = min(g.HW, p0 + g.chunk);
  const float4 w4 = nw ? __ldg(nw + q) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 b4 = bias ? __ldg(bias + q) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 mu = __ldg(stats + (int64_t)n * 2 * g.C4 + q), rs = __ldg(stats + ((int64_t)n * 2 + 1) * g.C4 + q);
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

// ---- backward pass 1: s1 = sum g, s2 = sum g*xhat per (n,c) chunk ------------------------------------------
__global__ void __launch_bounds__(TPB) se_bwd_stats_kernel(const float4* __restrict__ gout, const float4* __restrict__ x,
                                                           const float* __restrict__ noise, const float4* __restrict__ nw,
                                                           const float4* __restrict__ bias, const float4* __restrict__ stats,
                                                           float4* __restrict__ part, SE g) {
  __shared__ float4 red[2 * TPB];
  const int tid = threadIdx.x, q = tid % g.C4, rl = tid / g.C4, rows = TPB / g.C4;
  const int n = blockIdx.y, chunk = blockIdx.x;
  const int p0 = chunk * g.chunk, p1 = min(g.HW, p0 + g.chunk);
  const float4 w4 = nw ? __ldg(nw + q) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 b4 = bias ? __ldg(bias + q) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 mu = __ldg(stats + (int64_t)n * 2 * g.C4 + q), rs = __ldg(stats + ((int64_t)n * 2 + 1) * g.C4 + q);
  float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
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
  if (tid < g.C4) {
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
__global__ void __launch_bounds__(TPB) se_bwd_apply_kernel(const float4* __restrict__ gout, const float4* __restrict__ x,
                                                           const float* __restrict__ noise, const float4* __restrict__ nw,
                                                           const float4* __restrict__ bias, const float4* __restrict__ style,
                                                           const float4* __restrict__ stats, const float4* __restrict__ means,
                                                           float4* __restrict__ gx, float* __restrict__ g_nw,
                                                           float* __restrict__ g_bias, SE g) {
  __shared__ float4 red[2 * TPB];
  const int tid = threadIdx.x, q = tid % g.C4, rl = tid / g.C4, rows = TPB / g.C4;
  const int n = blockIdx.y, chunk = blockIdx.x;
  const int p0 = chunk * g.chunk, p1 = min(g.HW, p0 + g.chunk);
  const float4 w4 = nw ? __ldg(nw + q) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 b4 = bias ? __ldg(bias + q) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 mu = __ldg(stats + (int64_t)n * 2 * g.C4 + q), rs = __ldg(stats + ((int64_t)n * 2 + 1) * g.C4 + q);
  const float4 ys = __ldg(style + (int64_t)n * 2 * g.C4 + q);
  const float4 m1 = __ldg(means + (int64_t)n * 2 * g.C4 + q), m2 = __ldg(means + ((int64_t)n * 2 + 1) * g.C4 + q);
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

int se_geom(SE& g, int N, int H, int W, int C, float slope) {
  if (C % 4 != 0 || C / 4 > TPB || TPB % (C / 4) != 0) return shape_fail("style_epilogue: C/4 must divide 256");
  g.N = N; g.HW = H * W; g.C4 = C / 4; g.chunk = se_chunk(N, g.HW, C); g.chunks = (g.HW + g.chunk - 1) / g.chunk; g.slope = slope;
  if (N > 65535) return shape_fail("style_epilogue: N > 65535");
  return GLB_OK;
}

}  // namespace
}  // namespace glb

using namespace glb;

extern "C" int64_t glb_style_epilogue_work_floats(int N, int H, int W, int C) {
  if (C < 4) return 0;
  const int chunk = glb::se_chunk(N, H * W, C);
  const int64_t chunks = ((int64_t)H * W + chunk - 1) / chunk;
  return (int64_t)N * chunks * 3 * C + (int64_t)N * 2 * C;
}

extern "C" int glb_style_epilogue_fwd(const float* x, const float* noise, const float* noise_weight, const float* bias,
                                      const float* style, float* out, float* stats, float* work, int N, int H, int W, int C,
                                      float slope, float eps, glb_stream_t stream) {
  SE g;
  if (int rc = se_geom(g, N, H, W, C, slope)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
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
  SE g;
  if (int rc = se_geom(g, N, H, W, C, slope)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(g.chunks, N);
  float* means = work + (int64_t)N * g.chunks * 2 * C;
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
