// Normalisation kernels of the ResNet-GAN path, NHWC activations:
//   * BatchNorm2d with batch statistics (generator blocks, reference resnetgan/resblocks.py:43-46 via
//     utils/custom_layers.py:100-102) + the ReLU that always follows it       -> column statistics over [P = N*H*W][C]
//   * LayerNorm([C,H,W]) with elementwise affine (discriminator blocks, custom_layers.py:103-106) + ReLU
//                                                                           -> row statistics over [N][L = H*W*C]
//     including the second-order term the WGAN-GP penalty sends back through it (resnetgan/learner.py:811-825)
//   * Tanh on the generator output (resnetgan/architectures.py:58, 96)
// All bandwidth bound: 128-bit coalesced accesses, statistics as shifted sums (shift = first element of the row /
// column, so that E[(x-s)^2] - E[x-s]^2 does not cancel), block partials combined with fp64 atomics into a workspace.
#include "common.cuh"

namespace glb {
namespace {

constexpr int TPB = 256;

__device__ __forceinline__ float block_sum(float v, float* sh) {   // sh: >= 32 floats; result valid in every thread
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  float t = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.f;
  if (w == 0) {
    t = warp_sum(t);
    if (l == 0) sh[0] = t;
  }
  __syncthreads();
  return sh[0];
}

__device__ __forceinline__ void red_add4(float* addr, const float4& v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ float4 f4_mask(const float4& g, const float4& y, int act, float slope) {
  return make_float4(g.x * act_grad(y.x, act, slope), g.y * act_grad(y.y, act, slope), g.z * act_grad(y.z, act, slope),
                     g.w * act_grad(y.w, act, slope));
}

// ------------------------------------------------------------------------------------------------ LayerNorm
// ws[n*K + k] (double) accumulators, zeroed by the host wrapper before the stats kernels.
__global__ void ln_stats_kernel(const float4* __restrict__ x, double* __restrict__ ws, int64_t L4) {
  __shared__ float sh[32];
  const int n = blockIdx.y;
  const float4* xr = x + (int64_t)n * L4;
  const float shift = __ldg(reinterpret_cast<const float*>(xr));
  float s1 = 0.f, s2 = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < L4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = ldg_stream(xr + i);
    const float a = v.x - shift, b = v.y - shift, c = v.z - shift, d = v.w - shift;
    s1 += (a + b) + (c + d);
    s2 += (a * a + b * b) + (c * c + d * d);
  }
  s1 = block_sum(s1, sh);
  s2 = block_sum(s2, sh);
  if (threadIdx.x == 0) {
    atomicAdd(ws + 2 * n + 0, (double)s1);
    atomicAdd(ws + 2 * n + 1, (double)s2);
  }
}

__device__ __forceinline__ void ln_mean_rstd(const float4* xr, const double* ws, int n, int64_t L, float eps, float& mean,
                                             float& rstd) {
  const double shift = (double)__ldg(reinterpret_cast<const float*>(xr));
  const double m1 = ws[2 * n] / (double)L, m2 = ws[2 * n + 1] / (double)L;
  double var = m2 - m1 * m1;
  if (var < 0.0) var = 0.0;
  mean = (float)(shift + m1);
  rstd = (float)(1.0 / sqrt(var + (double)eps));
}

__global__ void ln_apply_kernel(const float4* __restrict__ x, const float4* __restrict__ gamma, const float4* __restrict__ beta,
                                float4* __restrict__ y, float* __restrict__ stats, const double* __restrict__ ws, int64_t L4,
                                float eps, int act, float slope) {
  const int n = blockIdx.y;
  const float4* xr = x + (int64_t)n * L4;
  float mean, rstd;
  ln_mean_rstd(xr, ws, n, L4 * 4, eps, mean, rstd);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    stats[2 * n] = mean;
    stats[2 * n + 1] = rstd;
  }
  float4* yr = y + (int64_t)n * L4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < L4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = ldg_stream(xr + i);
    const float4 g = __ldg(gamma + i), b = __ldg(beta + i);
    float4 o;
    o.x = act_apply((v.x - mean) * rstd * g.x + b.x, act, slope);
    o.y = act_apply((v.y - mean) * rstd * g.y + b.y, act, slope);
    o.z = act_apply((v.z - mean) * rstd * g.z + b.z, act, slope);
    o.w = act_apply((v.w - mean) * rstd * g.w + b.w, act, slope);
    stg_stream(yr + i, o);
  }
}

// per-sample sums for the backward family.  K = 2: {sum g, sum g*xhat};  K = 5 (u != null): + {sum u, sum u*xhat, sum u*g}
// with g = gy * act'(y) * gamma.
__global__ void ln_bwd_stats_kernel(const float4* __restrict__ gy, const float4* __restrict__ y, const float4* __restrict__ x,
                                    const float4* __restrict__ gamma, const float4* __restrict__ u,
                                    const float* __restrict__ stats, double* __restrict__ ws, int64_t L4, int act, float slope) {
  __shared__ float sh[32];
  const int n = blockIdx.y;
  const int K = (u != nullptr) ? 5 : 2;
  const float mean = stats[2 * n], rstd = stats[2 * n + 1];
  const int64_t base = (int64_t)n * L4;
  float s[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < L4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 g = ldg_stream(gy + base + i);
    if (y != nullptr) g = f4_mask(g, ldg_stream(y + base + i), act, slope);
    const float4 gm = __ldg(gamma + i);
    g.x *= gm.x; g.y *= gm.y; g.z *= gm.z; g.w *= gm.w;
    const float4 v = ldg_stream(x + base + i);
    const float4 xh = make_float4((v.x - mean) * rstd, (v.y - mean) * rstd, (v.z - mean) * rstd, (v.w - mean) * rstd);
    s[0] += (g.x + g.y) + (g.z + g.w);
    s[1] += (g.x * xh.x + g.y * xh.y) + (g.z * xh.z + g.w * xh.w);
    if (u != nullptr) {
      const float4 uu = ldg_stream(u + base + i);
      s[2] += (uu.x + uu.y) + (uu.z + uu.w);
      s[3] += (uu.x * xh.x + uu.y * xh.y) + (uu.z * xh.z + uu.w * xh.w);
      s[4] += (uu.x * g.x + uu.y * g.y) + (uu.z * g.z + uu.w * g.w);
    }
  }
  for (int k = 0; k < K; ++k) {
    const float t = block_sum(s[k], sh);
    if (threadIdx.x == 0) atomicAdd(ws + (int64_t)K * n + k, (double)t);
  }
}

// gx = rstd * (g - mean(g) - xhat * mean(g*xhat))
__global__ void ln_bwd_apply_kernel(const float4* __restrict__ gy, const float4* __restrict__ y, const float4* __restrict__ x,
                                    const float4* __restrict__ gamma, const float* __restrict__ stats,
                                    const double* __restrict__ ws, float4* __restrict__ gx, int64_t L4, int act, float slope) {
  const int n = blockIdx.y;
  const float mean = stats[2 * n], rstd = stats[2 * n + 1];
  const double invL = 1.0 / (double)(L4 * 4);
  const float a = (float)(ws[2 * n] * invL), b = (float)(ws[2 * n + 1] * invL);
  const int64_t base = (int64_t)n * L4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < L4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 g = ldg_stream(gy + base + i);
    if (y != nullptr) g = f4_mask(g, ldg_stream(y + base + i), act, slope);
    const float4 gm = __ldg(gamma + i);
    const float4 v = ldg_stream(x + base + i);
    float4 o;
    o.x = rstd * (g.x * gm.x - a - (v.x - mean) * rstd * b);
    o.y = rstd * (g.y * gm.y - a - (v.y - mean) * rstd * b);
    o.z = rstd * (g.z * gm.z - a - (v.z - mean) * rstd * b);
    o.w = rstd * (g.w * gm.w - a - (v.w - mean) * rstd * b);
    stg_stream(gx + base + i, o);
  }
}

// parameter gradients: one thread per float4 of the [L] affine maps, looping over the samples (coalesced across threads)
//   ggamma[l] = sum_n gym[n,l] * xhat[n,l],  gbeta[l] = sum_n gym[n,l],   gym = gy * act'(y)
__global__ void ln_param_grad_kernel(const float4* __restrict__ gy, const float4* __restrict__ y, const float4* __restrict__ x,
                                     const float* __restrict__ stats, float4* __restrict__ ggamma, float4* __restrict__ gbeta,
                                     int N, int64_t L4, int act, float slope) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= L4) return;
  float4 sg = make_float4(0.f, 0.f, 0.f, 0.f), sb = sg;
  // blockIdx.y owns a slice of the samples (the outputs are zeroed by the host and combined with red.add when sliced)
  const int per = (N + gridDim.y - 1) / gridDim.y, n_begin = blockIdx.y * per, n_end = min(N, n_begin + per);
  for (int n = n_begin; n < n_end; ++n) {
    const int64_t j = (int64_t)n * L4 + i;
    float4 g = ldg_stream(gy + j);
    if (y != nullptr) g = f4_mask(g, ldg_stream(y + j), act, slope);
    const float mean = stats[2 * n], rstd = stats[2 * n + 1];
    const float4 v = ldg_stream(x + j);
    sg.x += g.x * (v.x - mean) * rstd; sg.y += g.y * (v.y - mean) * rstd;
    sg.z += g.z * (v.z - mean) * rstd; sg.w += g.w * (v.w - mean) * rstd;
    sb.x += g.x; sb.y += g.y; sb.z += g.z; sb.w += g.w;
  }
  if (gridDim.y == 1) {
    ggamma[i] = sg;
    gbeta[i] = sb;
  } else {
    red_add4(reinterpret_cast<float*>(ggamma + i), sg);
    red_add4(reinterpret_cast<float*>(gbeta + i), sb);
  }
}

// Second order.  With u the cotangent of gx = B(gy; x, gamma), xh = xhat, r = rstd, g = gy*m*gamma,
//   a = mean g, b = mean g*xh, mu = mean u, mux = mean u*xh, mug = mean u*g,  P(v) = v - mean v - xh * mean(v*xh):
//   d/dgy    : m * gamma * r * P(u)
//   d/dx     : -r^2 * ( (mug - a*mu - b*mux) * xh + b * P(u) + mux * P(g) )
//   d/dgamma : sum_n gy*m * r_n * P(u)        (ln_bwdbwd_gamma_kernel)
__global__ void ln_bwdbwd_apply_kernel(const float4* __restrict__ u, const float4* __restrict__ gy, const float4* __restrict__ y,
                                       const float4* __restrict__ x, const float4* __restrict__ gamma,
                                       const float* __restrict__ stats, const double* __restrict__ ws, float4* __restrict__ g_gy,
                                       float4* __restrict__ g_x, int64_t L4, int act, float slope) {
  const int n = blockIdx.y;
  const float mean = stats[2 * n], r = stats[2 * n + 1];
  const double invL = 1.0 / (double)(L4 * 4);
  const float a = (float)(ws[5 * n] * invL), b = (float)(ws[5 * n + 1] * invL), mu = (float)(ws[5 * n + 2] * invL),
              mux = (float)(ws[5 * n + 3] * invL), mug = (float)(ws[5 * n + 4] * invL);
  const float c1 = mug - a * mu - b * mux;
  const float r2 = r * r;
  const int64_t base = (int64_t)n * L4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < L4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 gyv = ldg_stream(gy + base + i);
    float4 m = make_float4(1.f, 1.f, 1.f, 1.f);
    if (y != nullptr) m = f4_mask(m, ldg_stream(y + base + i), act, slope);
    const float4 gm = __ldg(gamma + i);
    const float4 v = ldg_stream(x + base + i);
    const float4 uu = ldg_stream(u + base + i);
    float4 ogy, ox;
#define GLB_LN2(c)                                                            \
    {                                                                         \
      const float xh = (v.c - mean) * r;                                      \
      const float g = gyv.c * m.c * gm.c;                                     \
      const float pu = uu.c - mu - xh * mux;                                  \
      const float pg = g - a - xh * b;                                        \
      ogy.c = m.c * gm.c * r * pu;                                            \
      ox.c = -r2 * (c1 * xh + b * pu + mux * pg);                             \
    }
    GLB_LN2(x) GLB_LN2(y) GLB_LN2(z) GLB_LN2(w)
#undef GLB_LN2
    if (g_gy != nullptr) stg_stream(g_gy + base + i, ogy);
    if (g_x != nullptr) stg_stream(g_x + base + i, ox);
  }
}

__global__ void ln_bwdbwd_gamma_kernel(const float4* __restrict__ u, const float4* __restrict__ gy, const float4* __restrict__ y,
                                       const float4* __restrict__ x, const float* __restrict__ stats,
                                       const double* __restrict__ ws, float4* __restrict__ g_gamma, int N, int64_t L4, int act,
                                       float slope) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= L4) return;
  const double invL = 1.0 / (double)(L4 * 4);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  const int per = (N + gridDim.y - 1) / gridDim.y, n_begin = blockIdx.y * per, n_end = min(N, n_begin + per);
  for (int n = n_begin; n < n_end; ++n) {
    const int64_t j = (int64_t)n * L4 + i;
    float4 g = ldg_stream(gy + j);
    if (y != nullptr) g = f4_mask(g, ldg_stream(y + j), act, slope);
    const float mean = stats[2 * n], r = stats[2 * n + 1];
    const float mu = (float)(ws[5 * n + 2] * invL), mux = (float)(ws[5 * n + 3] * invL);
    const float4 v = ldg_stream(x + j);
    const float4 uu = ldg_stream(u + j);
    s.x += g.x * r * (uu.x - mu - (v.x - mean) * r * mux);
    s.y += g.y * r * (uu.y - mu - (v.y - mean) * r * mux);
    s.z += g.z * r * (uu.z - mu - (v.z - mean) * r * mux);
    s.w += g.w * r * (uu.w - mu - (v.w - mean) * r * mux);
  }
  if (gridDim.y == 1) g_gamma[i] = s;
  else red_add4(reinterpret_cast<float*>(g_gamma + i), s);
}

// ------------------------------------------------------------------------------------------------ BatchNorm
// Column sums over [P][C] rows.  Thread layout: quad q = tid % C4 (C4 <= blockDim) and row lane tid / C4; block partials are
// combined in shared memory, then one fp64 atomic per column and block.
//   mode 0: {sum (x-shift_c), sum (x-shift_c)^2},  shift_c = x[0][c]
//   mode 1: {sum g, sum g*xhat},  g = gy * act'(y), xhat from stats (mean [C], rstd [C])
__global__ void bn_colsum_kernel(const float4* __restrict__ x, const float4* __restrict__ gy, const float4* __restrict__ y,
                                 const float* __restrict__ stats, double* __restrict__ ws, int64_t P, int C4, int mode, int act,
                                 float slope) {
  extern __shared__ float4 red[];   // [2][blockDim]
  const int tid = threadIdx.x;
  const int rows_per_it = blockDim.x / C4;
  const int q = tid % C4, rl = tid / C4;
  float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
  if (rl < rows_per_it) {
    float4 c0, c1;
    if (mode == 0) {
      c0 = __ldg(x + q);                    // shift
      c1 = c0;
    } else {
      c0 = __ldg(reinterpret_cast<const float4*>(stats) + q);            // mean
      c1 = __ldg(reinterpret_cast<const float4*>(stats) + C4 + q);       // rstd
    }
    for (int64_t p = (int64_t)blockIdx.x * rows_per_it + rl; p < P; p += (int64_t)gridDim.x * rows_per_it) {
      const int64_t i = p * C4 + q;
      const float4 v = ldg_stream(x + i);
      if (mode == 0) {
        const float a = v.x - c0.x, b = v.y - c0.y, c = v.z - c0.z, d = v.w - c0.w;
        s1.x += a; s1.y += b; s1.z += c; s1.w += d;
        s2.x += a * a; s2.y += b * b; s2.z += c * c; s2.w += d * d;
      } else {
        float4 g = ldg_stream(gy + i);
        if (y != nullptr) g = f4_mask(g, ldg_stream(y + i), act, slope);
        s1.x += g.x; s1.y += g.y; s1.z += g.z; s1.w += g.w;
        s2.x += g.x * (v.x - c0.x) * c1.x; s2.y += g.y * (v.y - c0.y) * c1.y;
        s2.z += g.z * (v.z - c0.z) * c1.z; s2.w += g.w * (v.w - c0.w) * c1.w;
      }
    }
  }
  red[tid] = s1;
  red[blockDim.x + tid] = s2;
  __syncthreads();
  if (tid < C4) {
    float4 t1 = red[tid], t2 = red[blockDim.x + tid];
    for (int r = 1; r < rows_per_it; ++r) {
      const float4 a = red[r * C4 + tid], b = red[blockDim.x + r * C4 + tid];
      t1.x += a.x; t1.y += a.y; t1.z += a.z; t1.w += a.w;
      t2.x += b.x; t2.y += b.y; t2.z += b.z; t2.w += b.w;
    }
    const int C = 4 * C4;
    atomicAdd(ws + 4 * tid + 0, (double)t1.x); atomicAdd(ws + 4 * tid + 1, (double)t1.y);
    atomicAdd(ws + 4 * tid + 2, (double)t1.z); atomicAdd(ws + 4 * tid + 3, (double)t1.w);
    atomicAdd(ws + C + 4 * tid + 0, (double)t2.x); atomicAdd(ws + C + 4 * tid + 1, (double)t2.y);
    atomicAdd(ws + C + 4 * tid + 2, (double)t2.z); atomicAdd(ws + C + 4 * tid + 3, (double)t2.w);
  }
}

// stats[c] = mean, stats[C + c] = rstd; running statistics as nn.BatchNorm2d keeps them (momentum update, unbiased variance)
__global__ void bn_finalize_kernel(const float* __restrict__ x, const double* __restrict__ ws, float* __restrict__ stats,
                                   float* __restrict__ running_mean, float* __restrict__ running_var, int64_t* __restrict__ nbt,
                                   int64_t P, int C, float eps, float momentum) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && nbt != nullptr) *nbt += 1;
  if (c >= C) return;
  const double shift = (double)x[c];
  const double m1 = ws[c] / (double)P, m2 = ws[C + c] / (double)P;
  double var = m2 - m1 * m1;
  if (var < 0.0) var = 0.0;
  const double mean = shift + m1;
  stats[c] = (float)mean;
  stats[C + c] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean != nullptr) {
    const double unb = (P > 1) ? var * (double)P / (double)(P - 1) : var;
    running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * mean);
    running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unb);
  }
}

__global__ void bn_apply_kernel(const float4* __restrict__ x, const float4* __restrict__ gamma, const float4* __restrict__ beta,
                                const float* __restrict__ stats, float4* __restrict__ y, int64_t n4, int C4, int act,
                                float slope) {
  const float4* mean4 = reinterpret_cast<const float4*>(stats);
  const float4* rstd4 = mean4 + C4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const int q = (int)(i % C4);
    const float4 v = ldg_stream(x + i);
    const float4 m = __ldg(mean4 + q), r = __ldg(rstd4 + q), g = __ldg(gamma + q), b = __ldg(beta + q);
    float4 o;
    o.x = act_apply((v.x - m.x) * r.x * g.x + b.x, act, slope);
    o.y = act_apply((v.y - m.y) * r.y * g.y + b.y, act, slope);
    o.z = act_apply((v.z - m.z) * r.z * g.z + b.z, act, slope);
    o.w = act_apply((v.w - m.w) * r.w * g.w + b.w, act, slope);
    stg_stream(y + i, o);
  }
}

// ggamma = sum g*xhat, gbeta = sum g (straight from the workspace)
__global__ void bn_bwd_finalize_kernel(const double* __restrict__ ws, float* __restrict__ ggamma, float* __restrict__ gbeta, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (gbeta != nullptr) gbeta[c] = (float)ws[c];
  if (ggamma != nullptr) ggamma[c] = (float)ws[C + c];
}

// gx = gamma * rstd * (g - mean_p g - xhat * mean_p(g*xhat))
__global__ void bn_bwd_apply_kernel(const float4* __restrict__ gy, const float4* __restrict__ y, const float4* __restrict__ x,
                                    const float4* __restrict__ gamma, const float* __restrict__ stats,
                                    const double* __restrict__ ws, float4* __restrict__ gx, int64_t n4, int C4, int64_t P, int act,
                                    float slope) {
  const float4* mean4 = reinterpret_cast<const float4*>(stats);
  const float4* rstd4 = mean4 + C4;
  const int C = 4 * C4;
  const double invP = 1.0 / (double)P;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const int q = (int)(i % C4);
    float4 g = ldg_stream(gy + i);
    if (y != nullptr) g = f4_mask(g, ldg_stream(y + i), act, slope);
    const float4 v = ldg_stream(x + i);
    const float4 m = __ldg(mean4 + q), r = __ldg(rstd4 + q), gm = __ldg(gamma + q);
    float4 o;
#define GLB_BN1(c, k)                                                          \
    {                                                                          \
      const float a = (float)(ws[4 * q + k] * invP), b = (float)(ws[C + 4 * q + k] * invP); \
      o.c = gm.c * r.c * (g.c - a - (v.c - m.c) * r.c * b);                    \
    }
    GLB_BN1(x, 0) GLB_BN1(y, 1) GLB_BN1(z, 2) GLB_BN1(w, 3)
#undef GLB_BN1
    stg_stream(gx + i, o);
  }
}

// ------------------------------------------------------------------------------------------------ tanh
__global__ void tanh_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) y[i] = tanhf(x[i]);
}
__global__ void tanh_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ y, float* __restrict__ gx, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float t = y[i];
    gx[i] = gy[i] * (1.f - t * t);
  }
}

inline dim3 ln_grid(int N, int64_t L4) {
  int chunks = (int)((L4 + TPB * 8 - 1) / (TPB * 8));     // ~8 float4 per thread
  const int cap = (kNumSMs * 16 + N - 1) / N;
  if (chunks > cap) chunks = cap;
  if (chunks < 1) chunks = 1;
  return dim3(chunks, N);
}

// sample slices (grid.y) of the LayerNorm parameter-gradient kernels: enough blocks for ~8 per SM, at least 4 samples each
inline int ln_sample_slices(int N, int64_t L4) {
  const int64_t bx = (L4 + 127) / 128;
  int ny = (int)((8 * (int64_t)kNumSMs + bx - 1) / bx);
  if (ny > N / 4) ny = N / 4;
  if (ny > 16) ny = 16;
  return ny < 1 ? 1 : ny;
}

inline int bn_threads(int C4) { return ((TPB + C4 - 1) / C4) * C4 > 1024 ? C4 : ((TPB + C4 - 1) / C4) * C4; }

}  // namespace
}  // namespace glb

using namespace glb;

extern "C" int64_t glb_layernorm_ws_doubles(int N) { return 5 * (int64_t)N; }

extern "C" int glb_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* y, float* stats, void* ws, int N,
                                 int64_t L, float eps, int act, float slope, glb_stream_t stream) {
  if (N <= 0 || L <= 0 || (L % 4) != 0) return shape_fail("layernorm_fwd (L must be a multiple of 4)");
  cudaStream_t st = (cudaStream_t)stream;
  GLB_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * 2 * N, st));
  const dim3 grid = ln_grid(N, L / 4);
  ln_stats_kernel<<<grid, TPB, 0, st>>>((const float4*)x, (double*)ws, L / 4);
  GLB_CHECK_LAUNCH("ln_stats_kernel");
  ln_apply_kernel<<<grid, TPB, 0, st>>>((const float4*)x, (const float4*)gamma, (const float4*)beta, (float4*)y, stats,
                                        (const double*)ws, L / 4, eps, act, slope);
  GLB_CHECK_LAUNCH("ln_apply_kernel");
  return GLB_OK;
}

extern "C" int glb_layernorm_bwd(const float* gy, const float* y, const float* x, const float* gamma, const float* stats, float* gx,
                                 float* ggamma, float* gbeta, void* ws, int N, int64_t L, int act, float slope,
                                 glb_stream_t stream) {
  if (N <= 0 || L <= 0 || (L % 4) != 0) return shape_fail("layernorm_bwd");
  cudaStream_t st = (cudaStream_t)stream;
  const float4* y4 = (act == GLB_ACT_NONE) ? nullptr : (const float4*)y;
  const dim3 grid = ln_grid(N, L / 4);
  if (gx != nullptr) {
    GLB_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * 2 * N, st));
    ln_bwd_stats_kernel<<<grid, TPB, 0, st>>>((const float4*)gy, y4, (const float4*)x, (const float4*)gamma, nullptr, stats,
                                              (double*)ws, L / 4, act, slope);
    GLB_CHECK_LAUNCH("ln_bwd_stats_kernel");
    ln_bwd_apply_kernel<<<grid, TPB, 0, st>>>((const float4*)gy, y4, (const float4*)x, (const float4*)gamma, stats,
                                              (const double*)ws, (float4*)gx, L / 4, act, slope);
    GLB_CHECK_LAUNCH("ln_bwd_apply_kernel");
  }
  if (ggamma != nullptr && gbeta != nullptr) {
    const int ny = ln_sample_slices(N, L / 4);
    if (ny > 1) {
      GLB_CUDA(cudaMemsetAsync(ggamma, 0, sizeof(float) * L, st));
      GLB_CUDA(cudaMemsetAsync(gbeta, 0, sizeof(float) * L, st));
    }
    ln_param_grad_kernel<<<dim3((unsigned)((L / 4 + 127) / 128), ny), 128, 0, st>>>((const float4*)gy, y4, (const float4*)x, stats,
                                                                         (float4*)ggamma, (float4*)gbeta, N, L / 4, act, slope);
    GLB_CHECK_LAUNCH("ln_param_grad_kernel");
  }
  return GLB_OK;
}

extern "C" int glb_layernorm_bwdbwd(const float* u, const float* gy, const float* y, const float* x, const float* gamma,
                                    const float* stats, float* g_gy, float* g_x, float* g_gamma, void* ws, int N, int64_t L,
                                    int act, float slope, glb_stream_t stream) {
  if (N <= 0 || L <= 0 || (L % 4) != 0) return shape_fail("layernorm_bwdbwd");
  cudaStream_t st = (cudaStream_t)stream;
  const float4* y4 = (act == GLB_ACT_NONE) ? nullptr : (const float4*)y;
  const dim3 grid = ln_grid(N, L / 4);
  GLB_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * 5 * N, st));
  ln_bwd_stats_kernel<<<grid, TPB, 0, st>>>((const float4*)gy, y4, (const float4*)x, (const float4*)gamma, (const float4*)u,
                                            stats, (double*)ws, L / 4, act, slope);
  GLB_CHECK_LAUNCH("ln_bwd_stats_kernel");
  if (g_gy != nullptr || g_x != nullptr) {
    ln_bwdbwd_apply_kernel<<<grid, TPB, 0, st>>>((const float4*)u, (const float4*)gy, y4, (const float4*)x,
                                                 (const float4*)gamma, stats, (const double*)ws, (float4*)g_gy, (float4*)g_x,
                                                 L / 4, act, slope);
    GLB_CHECK_LAUNCH("ln_bwdbwd_apply_kernel");
  }
  if (g_gamma != nullptr) {
    const int ny = ln_sample_slices(N, L / 4);
    if (ny > 1) GLB_CUDA(cudaMemsetAsync(g_gamma, 0, sizeof(float) * L, st));
    ln_bwdbwd_gamma_kernel<<<dim3((unsigned)((L / 4 + 127) / 128), ny), 128, 0, st>>>((const float4*)u, (const float4*)gy, y4,
                                                                           (const float4*)x, stats, (const double*)ws,
                                                                           (float4*)g_gamma, N, L / 4, act, slope);
    GLB_CHECK_LAUNCH("ln_bwdbwd_gamma_kernel");
  }
  return GLB_OK;
}

extern "C" int64_t glb_batchnorm_ws_doubles(int C) { return 2 * (int64_t)C; }

extern "C" int glb_batchnorm_fwd(const float* x, const float* gamma, const float* beta, float* y, float* stats,
                                 float* running_mean, float* running_var, int64_t* num_batches_tracked, void* ws, int64_t P,
                                 int C, float eps, float momentum, int act, float slope, glb_stream_t stream) {
  if (P <= 0 || C <= 0 || (C % 4) != 0 || C > 4096) return shape_fail("batchnorm_fwd (C must be a multiple of 4, <= 4096)");
  cudaStream_t st = (cudaStream_t)stream;
  const int C4 = C / 4, threads = bn_threads(C4);
  GLB_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * 2 * C, st));
  const int rows_per_it = threads / C4;
  const int blocks = grid_for((P + rows_per_it - 1) / rows_per_it, 8, kNumSMs * 4);
  bn_colsum_kernel<<<blocks, threads, 2 * threads * sizeof(float4), st>>>((const float4*)x, nullptr, nullptr, nullptr,
                                                                          (double*)ws, P, C4, 0, act, slope);
  GLB_CHECK_LAUNCH("bn_colsum_kernel");
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, st>>>(x, (const double*)ws, stats, running_mean, running_var,
                                                      num_batches_tracked, P, C, eps, momentum);
  GLB_CHECK_LAUNCH("bn_finalize_kernel");
  const int64_t n4 = P * C4;
  bn_apply_kernel<<<grid_for(n4, TPB * 4), TPB, 0, st>>>((const float4*)x, (const float4*)gamma, (const float4*)beta, stats,
                                                         (float4*)y, n4, C4, act, slope);
  GLB_CHECK_LAUNCH("bn_apply_kernel");
  return GLB_OK;
}

extern "C" int glb_batchnorm_bwd(const float* gy, const float* y, const float* x, const float* gamma, const float* stats,
                                 float* gx, float* ggamma, float* gbeta, void* ws, int64_t P, int C, int act, float slope,
                                 glb_stream_t stream) {
  if (P <= 0 || C <= 0 || (C % 4) != 0 || C > 4096) return shape_fail("batchnorm_bwd");
  cudaStream_t st = (cudaStream_t)stream;
  const int C4 = C / 4, threads = bn_threads(C4);
  const float4* y4 = (act == GLB_ACT_NONE) ? nullptr : (const float4*)y;
  GLB_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * 2 * C, st));
  const int rows_per_it = threads / C4;
  const int blocks = grid_for((P + rows_per_it - 1) / rows_per_it, 8, kNumSMs * 4);
  bn_colsum_kernel<<<blocks, threads, 2 * threads * sizeof(float4), st>>>((const float4*)x, (const float4*)gy, y4, stats,
                                                                          (double*)ws, P, C4, 1, act, slope);
  GLB_CHECK_LAUNCH("bn_colsum_kernel");
  if (ggamma != nullptr || gbeta != nullptr) {
    bn_bwd_finalize_kernel<<<(C + 127) / 128, 128, 0, st>>>((const double*)ws, ggamma, gbeta, C);
    GLB_CHECK_LAUNCH("bn_bwd_finalize_kernel");
  }
  if (gx != nullptr) {
    const int64_t n4 = P * C4;
    bn_bwd_apply_kernel<<<grid_for(n4, TPB * 4), TPB, 0, st>>>((const float4*)gy, y4, (const float4*)x, (const float4*)gamma,
                                                               stats, (const double*)ws, (float4*)gx, n4, C4, P, act, slope);
    GLB_CHECK_LAUNCH("bn_bwd_apply_kernel");
  }
  return GLB_OK;
}

extern "C" int glb_tanh_fwd(const float* x, float* y, int64_t n, glb_stream_t stream) {
  if (n <= 0) return shape_fail("tanh_fwd");
  tanh_fwd_kernel<<<grid_for(n, TPB * 4), TPB, 0, (cudaStream_t)stream>>>(x, y, n);
  GLB_CHECK_LAUNCH("tanh_fwd_kernel");
  return GLB_OK;
}

extern "C" int glb_tanh_bwd(const float* gy, const float* y, float* gx, int64_t n, glb_stream_t stream) {
  if (n <= 0) return shape_fail("tanh_bwd");
  tanh_bwd_kernel<<<grid_for(n, TPB * 4), TPB, 0, (cudaStream_t)stream>>>(gy, y, gx, n);
  GLB_CHECK_LAUNCH("tanh_bwd_kernel");
  return GLB_OK;
}
