// Small-batch equalized-LR linear layers (LinearEx, reference utils/custom_layers.py:230-291):
//   y[M,Nout] = act(alpha * x[M,K] . w[Nout,K]^T + bias_scale * bias)          (aten::addmm + mul)
// with M = the per-GPU batch (8 .. 16).  These GEMMs move the whole weight matrix for a handful of rows: they are
// weight-bandwidth / launch bound, not tensor-core work, so they are streaming kernels:
// every weight element is read exactly once, coalesced, and the M activation rows sit in shared memory.
// Larger M falls through to the implicit-GEMM kernels (conv_simt.cu).
#include "common.cuh"

namespace glb {

int conv_fprop_simt(const float*, const float*, const float*, float*, int, int, int, int, int, int, int, int, float, float, int,
                    float, cudaStream_t);
int conv_dgrad_simt(const float*, const float*, float*, int, int, int, int, int, int, int, int, float, cudaStream_t);
int conv_wgrad_simt(const float*, const float*, float*, int, int, int, int, int, int, int, int, float, cudaStream_t);

namespace {

constexpr int kMaxM = 16;     // rows handled by the streaming kernels
constexpr int kMaxXs = 8192;  // floats of shared memory for the activation rows (M*K <= 8192)

// ---- forward: one warp per output feature, lanes split K (float4), warp-shuffle reduction ----------------------
template <int MB>
__global__ void __launch_bounds__(256) linear_fwd_small_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                               const float* __restrict__ bias, float* __restrict__ y, int M, int K,
                                                               int Nout, float alpha, float bias_scale, int act, float slope) {
  extern __shared__ __align__(16) float xs[];  // [M][K]
  for (int i = threadIdx.x * 4; i < M * K; i += blockDim.x * 4)
    *reinterpret_cast<float4*>(xs + i) = __ldg(reinterpret_cast<const float4*>(x + i));
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * (blockDim.x >> 5) + warp;
  if (n >= Nout) return;
  float acc[MB];
#pragma unroll
  for (int m = 0; m < MB; ++m) acc[m] = 0.f;
  const float* wr = w + (int64_t)n * K;
  for (int k0 = lane * 4; k0 < K; k0 += 4 * 128) {       // four weight quads in flight per lane
    float4 wv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (k0 + u * 128 < K) wv[u] = ldg_stream(reinterpret_cast<const float4*>(wr + k0 + u * 128));
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int k = k0 + u * 128;
      if (k < K) {
#pragma unroll
        for (int m = 0; m < MB; ++m) {
          if (m < M) {
            const float4 xv = *reinterpret_cast<const float4*>(xs + m * K + k);
            acc[m] = fmaf(xv.x, wv[u].x, fmaf(xv.y, wv[u].y, fmaf(xv.z, wv[u].z, fmaf(xv.w, wv[u].w, acc[m]))));
          }
        }
      }
    }
  }
#pragma unroll
  for (int m = 0; m < MB; ++m) acc[m] = warp_sum(acc[m]);
  const float b = bias ? bias_scale * __ldg(bias + n) : 0.f;
#pragma unroll
  for (int m = 0; m < MB; ++m)
    if (lane == m && m < M) y[(int64_t)m * Nout + n] = act_apply(alpha * acc[m] + b, act, slope);
}

// ---- dgrad: gx[m][k] = alpha * sum_n gy[m][n] w[n][k].  Thread = one k quad; block = one slice of n; slices combine
// with red.add into the zero-initialised gx (a few dozen adds per element).
template <int MB>
__global__ void __launch_bounds__(128) linear_dgrad_small_kernel(const float* __restrict__ gy, const float* __restrict__ w,
                                                                 float* __restrict__ gx, int M, int K, int Nout, int n_per_block,
                                                                 float alpha) {
  extern __shared__ __align__(16) float gs[];  // [n_per_block][M]
  const int n0 = blockIdx.y * n_per_block;
  const int nn = min(n_per_block, Nout - n0);
  for (int i = threadIdx.x; i < nn * M; i += blockDim.x) {
    const int j = i / M, m = i - j * M;
    gs[i] = __ldg(gy + (int64_t)m * Nout + n0 + j);
  }
  __syncthreads();
  const int k = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (k >= K) return;
  float4 acc[MB];
#pragma unroll
  for (int m = 0; m < MB; ++m) acc[m] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int j0 = 0; j0 < nn; j0 += 8) {                    // eight weight rows in flight per thread
    float4 wv[8];
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (j0 + u < nn) wv[u] = ldg_stream(reinterpret_cast<const float4*>(w + (int64_t)(n0 + j0 + u) * K + k));
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (j0 + u < nn) {
#pragma unroll
        for (int m = 0; m < MB; ++m) {
          if (m < M) {
            const float g = gs[(j0 + u) * M + m];
            acc[m].x = fmaf(g, wv[u].x, acc[m].x); acc[m].y = fmaf(g, wv[u].y, acc[m].y);
            acc[m].z = fmaf(g, wv[u].z, acc[m].z); acc[m].w = fmaf(g, wv[u].w, acc[m].w);
          }
        }
      }
    }
  }
#pragma unroll
  for (int m = 0; m < MB; ++m) {
    if (m < M) {
      float* o = gx + (int64_t)m * K + k;
      if (gridDim.y == 1) {
        *reinterpret_cast<float4*>(o) = make_float4(alpha * acc[m].x, alpha * acc[m].y, alpha * acc[m].z, alpha * acc[m].w);
      } else {
        atomicAdd(o + 0, alpha * acc[m].x); atomicAdd(o + 1, alpha * acc[m].y);
        atomicAdd(o + 2, alpha * acc[m].z); atomicAdd(o + 3, alpha * acc[m].w);
      }
    }
  }
}

// ---- wgrad: gw[n][k] = alpha * sum_m gy[m][n] x[m][k]  (rank-M outer product; pure weight-gradient write stream) ----
template <int MB>
__global__ void __launch_bounds__(256) linear_wgrad_small_kernel(const float* __restrict__ x, const float* __restrict__ gy,
                                                                 float* __restrict__ gw, int M, int K, int Nout, int rows_per_block,
                                                                 float alpha) {
  extern __shared__ __align__(16) float xs[];  // [M][K]
  for (int i = threadIdx.x * 4; i < M * K; i += blockDim.x * 4)
    *reinterpret_cast<float4*>(xs + i) = __ldg(reinterpret_cast<const float4*>(x + i));
  __syncthreads();
  const int n0 = blockIdx.x * rows_per_block;
  const int K4 = K >> 2;
  for (int i = threadIdx.x; i < rows_per_block * K4; i += blockDim.x) {
    const int r = i / K4, k = (i - r * K4) * 4;
    const int n = n0 + r;
    if (n >= Nout) break;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int m = 0; m < MB; ++m) {
      if (m < M) {
        const float g = __ldg(gy + (int64_t)m * Nout + n);
        const float4 xv = *reinterpret_cast<const float4*>(xs + m * K + k);
        acc.x = fmaf(g, xv.x, acc.x); acc.y = fmaf(g, xv.y, acc.y); acc.z = fmaf(g, xv.z, acc.z); acc.w = fmaf(g, xv.w, acc.w);
      }
    }
    stg_stream(reinterpret_cast<float4*>(gw + (int64_t)n * K + k), make_float4(alpha * acc.x, alpha * acc.y, alpha * acc.z, alpha * acc.w));
  }
}

// ---- grouped small linears: L layers with their own weights, all fed by rows of ws[L][M][K], in ONE launch each for forward,
// dgrad and wgrad (the 12 style affines of a StyleGAN generator pass, stylegan/architectures.py:460 / 524: 12 + 12 + 12
// launch-bound kernels plus the select/accumulate kernels autograd adds around them).
// Table per layer (int64[8]): weight ptr, bias ptr, nout, column offset, first block (fwd), first block (dgrad),
// first block (wgrad), alpha / bias_scale packed as two floats.
struct GLayer {
  const float* w;
  const float* b;
  long long nout, off, blk_fwd, blk_dgrad, blk_wgrad;
  float alpha, bscale;
};
static_assert(sizeof(GLayer) == 64, "table row = 8 x int64");

__device__ __forceinline__ int glayer_of(const GLayer* tab, int L, long long GLayer::*start, int blk) {
  int l = 0;
  while (l + 1 < L && (long long)blk >= tab[l + 1].*start) ++l;
  return l;
}

constexpr int kGWarps = 4, kGDgradRows = 16, kGWgradRows = 8;

// y chunk of layer l = [M][nout_l] contiguous at y + M * off_l
template <int MB>
__global__ void __launch_bounds__(kGWarps * 32) glinear_fwd_kernel(const float* __restrict__ ws, const GLayer* __restrict__ tab,
                                                                   float* __restrict__ y, int L, int M, int K) {
  extern __shared__ __align__(16) float xs[];
  const int l = glayer_of(tab, L, &GLayer::blk_fwd, blockIdx.x);
  const GLayer g = tab[l];
  const float* x = ws + (int64_t)l * M * K;
  for (int i = threadIdx.x * 4; i < M * K; i += blockDim.x * 4)
    *reinterpret_cast<float4*>(xs + i) = __ldg(reinterpret_cast<const float4*>(x + i));
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = (blockIdx.x - (int)g.blk_fwd) * kGWarps + warp;
  if (n >= g.nout) return;
  float acc[MB];
#pragma unroll
  for (int m = 0; m < MB; ++m) acc[m] = 0.f;
  const float* wr = g.w + (int64_t)n * K;
  for (int k0 = lane * 4; k0 < K; k0 += 4 * 128) {
    float4 wv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (k0 + u * 128 < K) wv[u] = ldg_stream(reinterpret_cast<const float4*>(wr + k0 + u * 128));
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int k = k0 + u * 128;
      if (k < K) {
#pragma unroll
        for (int m = 0; m < MB; ++m)
          if (m < M) {
            const float4 xv = *reinterpret_cast<const float4*>(xs + m * K + k);
            acc[m] = fmaf(xv.x, wv[u].x, fmaf(xv.y, wv[u].y, fmaf(xv.z, wv[u].z, fmaf(xv.w, wv[u].w, acc[m]))));
          }
      }
    }
  }
#pragma unroll
  for (int m = 0; m < MB; ++m) acc[m] = warp_sum(acc[m]);
  const float b = g.b ? g.bscale * __ldg(g.b + n) : 0.f;
  float* yl = y + (int64_t)M * g.off;
#pragma unroll
  for (int m = 0; m < MB; ++m)
    if (lane == m && m < M) yl[(int64_t)m * g.nout + n] = g.alpha * acc[m] + b;
}

// g_all [M][G] (row stride G, layer l at columns off_l ..); g_ws [L][M][K] zeroed by the host, slices combine with atomics
template <int MB>
__global__ void __launch_bounds__(128) glinear_dgrad_kernel(const float* __restrict__ g_all, int G, const GLayer* __restrict__ tab,
                                                            float* __restrict__ g_ws, int L, int M, int K) {
  __shared__ float gs[kGDgradRows * 16];
  const int l = glayer_of(tab, L, &GLayer::blk_dgrad, blockIdx.x);
  const GLayer g = tab[l];
  const int n0 = (blockIdx.x - (int)g.blk_dgrad) * kGDgradRows;
  const int nn = min(kGDgradRows, (int)g.nout - n0);
  for (int i = threadIdx.x; i < nn * M; i += blockDim.x) {
    const int j = i / M, m = i - j * M;
    gs[i] = __ldg(g_all + (int64_t)m * G + g.off + n0 + j);
  }
  __syncthreads();
  for (int k = threadIdx.x * 4; k < K; k += blockDim.x * 4) {
    float4 acc[MB];
#pragma unroll
    for (int m = 0; m < MB; ++m) acc[m] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j0 = 0; j0 < nn; j0 += 8) {
      float4 wv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (j0 + u < nn) wv[u] = ldg_stream(reinterpret_cast<const float4*>(g.w + (int64_t)(n0 + j0 + u) * K + k));
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (j0 + u < nn) {
#pragma unroll
          for (int m = 0; m < MB; ++m)
            if (m < M) {
              const float gv = gs[(j0 + u) * M + m];
              acc[m].x = fmaf(gv, wv[u].x, acc[m].x); acc[m].y = fmaf(gv, wv[u].y, acc[m].y);
              acc[m].z = fmaf(gv, wv[u].z, acc[m].z); acc[m].w = fmaf(gv, wv[u].w, acc[m].w);
            }
        }
    }
#pragma unroll
    for (int m = 0; m < MB; ++m)
      if (m < M) {
        float* o = g_ws + ((int64_t)l * M + m) * K + k;
        atomicAdd(o + 0, g.alpha * acc[m].x); atomicAdd(o + 1, g.alpha * acc[m].y);
        atomicAdd(o + 2, g.alpha * acc[m].z); atomicAdd(o + 3, g.alpha * acc[m].w);
      }
  }
}

// gw_all [G][K] (layer l = rows off_l ..), gb_all [G]
template <int MB>
__global__ void __launch_bounds__(256) glinear_wgrad_kernel(const float* __restrict__ ws, const float* __restrict__ g_all, int G,
                                                            const GLayer* __restrict__ tab, float* __restrict__ gw_all,
                                                            float* __restrict__ gb_all, int L, int M, int K) {
  extern __shared__ __align__(16) float xs[];
  const int l = glayer_of(tab, L, &GLayer::blk_wgrad, blockIdx.x);
  const GLayer g = tab[l];
  const float* x = ws + (int64_t)l * M * K;
  for (int i = threadIdx.x * 4; i < M * K; i += blockDim.x * 4)
    *reinterpret_cast<float4*>(xs + i) = __ldg(reinterpret_cast<const float4*>(x + i));
  __syncthreads();
  const int n0 = (blockIdx.x - (int)g.blk_wgrad) * kGWgradRows;
  const int K4 = K >> 2;
  for (int i = threadIdx.x; i < kGWgradRows * K4; i += blockDim.x) {
    const int r = i / K4, k = (i - r * K4) * 4;
    const int n = n0 + r;
    if (n >= g.nout) break;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float gsum = 0.f;
#pragma unroll
    for (int m = 0; m < MB; ++m)
      if (m < M) {
        const float gv = __ldg(g_all + (int64_t)m * G + g.off + n);
        const float4 xv = *reinterpret_cast<const float4*>(xs + m * K + k);
        acc.x = fmaf(gv, xv.x, acc.x); acc.y = fmaf(gv, xv.y, acc.y); acc.z = fmaf(gv, xv.z, acc.z); acc.w = fmaf(gv, xv.w, acc.w);
        gsum += gv;
      }
    stg_stream(reinterpret_cast<float4*>(gw_all + (int64_t)(g.off + n) * K + k),
               make_float4(g.alpha * acc.x, g.alpha * acc.y, g.alpha * acc.z, g.alpha * acc.w));
    if (k == 0 && gb_all != nullptr) gb_all[g.off + n] = g.bscale * gsum;
  }
}

inline bool small_ok(int M, int K) { return M <= kMaxM && K % 4 == 0 && (int64_t)M * K <= kMaxXs; }

}  // namespace
}  // namespace glb

using namespace glb;

extern "C" int glb_linear_fwd(const float* x, const float* w, const float* bias, float* y, int M, int K, int Nout, float alpha,
                              float bias_scale, int act, float slope, glb_stream_t stream) {
  if (M <= 0 || K <= 0 || Nout <= 0) return shape_fail("linear_fwd");
  cudaStream_t st = (cudaStream_t)stream;
  if (!small_ok(M, K)) return conv_fprop_simt(x, w, bias, y, M, 1, 1, K, Nout, 1, 1, 0, alpha, bias_scale, act, slope, st);
  const int warps = Nout >= 4096 ? 8 : 4;           // >= 128 blocks for the 512-wide layers (mapping net, style affines)
  const dim3 grid((Nout + warps - 1) / warps);
  const size_t smem = sizeof(float) * (size_t)M * K;
  if (M <= 8)
    linear_fwd_small_kernel<8><<<grid, warps * 32, smem, st>>>(x, w, bias, y, M, K, Nout, alpha, bias_scale, act, slope);
  else
    linear_fwd_small_kernel<16><<<grid, warps * 32, smem, st>>>(x, w, bias, y, M, K, Nout, alpha, bias_scale, act, slope);
  GLB_CHECK_LAUNCH("linear_fwd_small_kernel");
  return GLB_OK;
}

extern "C" int glb_linear_dgrad(const float* gy, const float* w, float* gx, int M, int K, int Nout, float alpha,
                                glb_stream_t stream) {
  if (M <= 0 || K <= 0 || Nout <= 0) return shape_fail("linear_dgrad");
  cudaStream_t st = (cudaStream_t)stream;
  if (!small_ok(M, K)) return conv_dgrad_simt(gy, w, gx, M, 1, 1, K, Nout, 1, 1, 0, alpha, st);
  const int kblocks = (K / 4 + 127) / 128;
  int nsplit = (2 * kNumSMs) / kblocks;                       // aim at ~2 blocks per SM
  if (nsplit > (Nout + 15) / 16) nsplit = (Nout + 15) / 16;   // >= 16 weight rows per block
  if (nsplit < 1) nsplit = 1;
  const int n_per_block = (Nout + nsplit - 1) / nsplit;
  nsplit = (Nout + n_per_block - 1) / n_per_block;
  if (nsplit > 1) GLB_CUDA(cudaMemsetAsync(gx, 0, sizeof(float) * (size_t)M * K, st));
  const dim3 grid(kblocks, nsplit);
  const size_t smem = sizeof(float) * (size_t)n_per_block * M;
  if (M <= 8)
    linear_dgrad_small_kernel<8><<<grid, 128, smem, st>>>(gy, w, gx, M, K, Nout, n_per_block, alpha);
  else
    linear_dgrad_small_kernel<16><<<grid, 128, smem, st>>>(gy, w, gx, M, K, Nout, n_per_block, alpha);
  GLB_CHECK_LAUNCH("linear_dgrad_small_kernel");
  return GLB_OK;
}

extern "C" int glb_linear_wgrad(const float* x, const float* gy, float* gw, int M, int K, int Nout, float alpha,
                                glb_stream_t stream) {
  if (M <= 0 || K <= 0 || Nout <= 0) return shape_fail("linear_wgrad");
  cudaStream_t st = (cudaStream_t)stream;
  if (!small_ok(M, K)) return conv_wgrad_simt(x, gy, gw, M, 1, 1, K, Nout, 1, 1, 0, alpha, st);
  int rows_per_block = (int)(((int64_t)Nout + 2 * kNumSMs - 1) / (2 * kNumSMs));
  const int min_rows = (1024 + K / 4 - 1) / (K / 4);  // >= 4 float4 stores per thread
  if (rows_per_block < min_rows) rows_per_block = min_rows;
  const dim3 grid((Nout + rows_per_block - 1) / rows_per_block);
  const size_t smem = sizeof(float) * (size_t)M * K;
  if (M <= 8)
    linear_wgrad_small_kernel<8><<<grid, 256, smem, st>>>(x, gy, gw, M, K, Nout, rows_per_block, alpha);
  else
    linear_wgrad_small_kernel<16><<<grid, 256, smem, st>>>(x, gy, gw, M, K, Nout, rows_per_block, alpha);
  GLB_CHECK_LAUNCH("linear_wgrad_small_kernel");
  return GLB_OK;
}

extern "C" int glb_glinear_fwd(const float* ws, const void* tab, float* y, int L, int M, int K, int blocks, glb_stream_t stream) {
  if (L <= 0 || M <= 0 || M > 16 || K % 4 != 0 || (int64_t)M * K > 8192 || blocks <= 0) return shape_fail("glinear_fwd");
  const size_t smem = sizeof(float) * (size_t)M * K;
  if (M <= 8) glinear_fwd_kernel<8><<<blocks, kGWarps * 32, smem, (cudaStream_t)stream>>>(ws, (const GLayer*)tab, y, L, M, K);
  else glinear_fwd_kernel<16><<<blocks, kGWarps * 32, smem, (cudaStream_t)stream>>>(ws, (const GLayer*)tab, y, L, M, K);
  GLB_CHECK_LAUNCH("glinear_fwd_kernel");
  return GLB_OK;
}

extern "C" int glb_glinear_dgrad(const float* g_all, int G, const void* tab, float* g_ws, int L, int M, int K, int blocks,
                                 glb_stream_t stream) {
  if (L <= 0 || M <= 0 || M > 16 || K % 4 != 0 || blocks <= 0) return shape_fail("glinear_dgrad");
  cudaStream_t st = (cudaStream_t)stream;
  GLB_CUDA(cudaMemsetAsync(g_ws, 0, sizeof(float) * (size_t)L * M * K, st));
  if (M <= 8) glinear_dgrad_kernel<8><<<blocks, 128, 0, st>>>(g_all, G, (const GLayer*)tab, g_ws, L, M, K);
  else glinear_dgrad_kernel<16><<<blocks, 128, 0, st>>>(g_all, G, (const GLayer*)tab, g_ws, L, M, K);
  GLB_CHECK_LAUNCH("glinear_dgrad_kernel");
  return GLB_OK;
}

extern "C" int glb_glinear_wgrad(const float* ws, const float* g_all, int G, const void* tab, float* gw_all, float* gb_all, int L,
                                 int M, int K, int blocks, glb_stream_t stream) {
  if (L <= 0 || M <= 0 || M > 16 || K % 4 != 0 || (int64_t)M * K > 8192 || blocks <= 0) return shape_fail("glinear_wgrad");
  const size_t smem = sizeof(float) * (size_t)M * K;
  cudaStream_t st = (cudaStream_t)stream;
  if (M <= 8) glinear_wgrad_kernel<8><<<blocks, 256, smem, st>>>(ws, g_all, G, (const GLayer*)tab, gw_all, gb_all, L, M, K);
  else glinear_wgrad_kernel<16><<<blocks, 256, smem, st>>>(ws, g_all, G, (const GLayer*)tab, gw_all, gb_all, L, M, K);
  GLB_CHECK_LAUNCH("glinear_wgrad_kernel");
  return GLB_OK;
}

extern "C" int glb_glinear_rows(int which) { return which == 0 ? kGWarps : (which == 1 ? kGDgradRows : kGWgradRows); }
