// Exact-fp32 implicit-GEMM convolution family (FFMA, shared-memory tiled).
//
// This is the fp32-accurate implementation behind GLB_IMPL_FP32 (SURVEY.md section 7 "TF32 vs the 1e-4
// bar") and the path for shapes the tcgen05 kernels do not cover (Ci = 513 after the minibatch-stddev
// concat, tiny channel counts).  Replaces aten::convolution / convolution_backward as dispatched from
// Conv2dEx (reference utils/custom_layers.py:166-169, 202-211) and nn.Linear in LinearEx (:249, 282-291).
//
// Data layout: activations NHWC, weights KRSC ([Co][R][S][Ci]).
//   fprop: out[m=(n,oh,ow)][co] = sum_{r,s,ci} in[n,oh+r-pad,ow+s-pad,ci] * w[co][r][s][ci]
//   dgrad: out[m=(n,h,w)][ci]   = sum_{r,s,co} gy[n,h+pad-r,w+pad-s,co]   * w[co][r][s][ci]
//   wgrad: gw[co][r][s][ci]     = sum_{p=(n,oh,ow)} gy[p][co] * x[n,oh+r-pad,ow+s-pad,ci]
#include "common.cuh"

namespace glb {
namespace {

constexpr int BM = 128, BN = 128, BK = 16, NT = 256;

struct IGemm {
  // input tensor [N, IH, IW, IC], output [N, OH, OW, OC]
  int N, IH, IW, IC, OH, OW, OC, R, S;
  int dh0, dw0, dsign;  // input row = oh + dh0 + dsign*r
  int M, K;             // M = N*OH*OW, K = R*S*IC
  int kper;             // k range per split-K slice (multiple of BK)
};

// MODE 0 = fprop (weight element (oc, r, s, ic) at ((oc*R+r)*S+s)*IC+ic  -> contiguous along k)
// MODE 1 = dgrad (weight element (ic=co, r, s, oc=ci) at ((ic*R+r)*S+s)*OC+oc -> contiguous along n)
template <int MODE, bool VEC>
__global__ void __launch_bounds__(NT) igemm_kernel(const float* __restrict__ in, const float* __restrict__ wgt,
                                                   const float* __restrict__ bias, float* __restrict__ out,
                                                   IGemm g, float alpha, float bias_scale, int act, float slope) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

  // A loader: thread owns row (tid % BM) and k sub-range [(tid / BM) * 8, +8)
  const int a_m = tid % BM, a_kq = (tid / BM) * 8;
  const int gm = m0 + a_m;
  int a_n = 0, a_oh = 0, a_ow = 0;
  const bool a_valid = gm < g.M;
  if (a_valid) {
    a_ow = gm % g.OW;
    int t = gm / g.OW;
    a_oh = t % g.OH;
    a_n = t / g.OH;
  }
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const int tx = tid % 16, ty = tid / 16;

  // split-K: slice z of gridDim.z owns k in [k_begin, k_end); partial tiles are combined with atomicAdd
  const int k_begin = blockIdx.z * g.kper;
  const int k_end = min(g.K, k_begin + g.kper);
  for (int k0 = k_begin; k0 < k_end; k0 += BK) {
    // ---- A tile ----
    float av[8];
    if (VEC) {
      // IC % 16 == 0: the 16-wide k block lies inside one tap
      const int tap = k0 / g.IC, ic = k0 - tap * g.IC + a_kq;
      const int r = tap / g.S, s = tap - r * g.S;
      const int ih = a_oh + g.dh0 + g.dsign * r, iw = a_ow + g.dw0 + g.dsign * s;
      if (a_valid && ih >= 0 && ih < g.IH && iw >= 0 && iw < g.IW) {
        const float4* p = reinterpret_cast<const float4*>(in + (((int64_t)a_n * g.IH + ih) * g.IW + iw) * g.IC + ic);
        float4 v0 = __ldg(p), v1 = __ldg(p + 1);
        av[0] = v0.x; av[1] = v0.y; av[2] = v0.z; av[3] = v0.w;
        av[4] = v1.x; av[5] = v1.y; av[6] = v1.z; av[7] = v1.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) av[j] = 0.f;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = k0 + a_kq + j;
        float v = 0.f;
        if (a_valid && k < k_end) {
          const int tap = k / g.IC, ic = k - tap * g.IC;
          const int r = tap / g.S, s = tap - r * g.S;
          const int ih = a_oh + g.dh0 + g.dsign * r, iw = a_ow + g.dw0 + g.dsign * s;
          if (ih >= 0 && ih < g.IH && iw >= 0 && iw < g.IW)
            v = __ldg(in + (((int64_t)a_n * g.IH + ih) * g.IW + iw) * g.IC + ic);
        }
        av[j] = v;
      }
    }
    // ---- B tile ----
    float bv[8];
    int b_k, b_n;  // where this thread's 8 values go: MODE 0 -> Bs[b_k + j][b_n]; MODE 1 -> Bs[b_k][b_n + j]
    if (MODE == 0) {
      b_n = tid % BN; b_k = (tid / BN) * 8;
      const int oc = n0 + b_n;
      if (VEC) {
        if (oc < g.OC) {
          const float4* p = reinterpret_cast<const float4*>(wgt + (int64_t)oc * g.K + k0 + b_k);
          float4 v0 = __ldg(p), v1 = __ldg(p + 1);
          bv[0] = v0.x; bv[1] = v0.y; bv[2] = v0.z; bv[3] = v0.w;
          bv[4] = v1.x; bv[5] = v1.y; bv[6] = v1.z; bv[7] = v1.w;
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) bv[j] = 0.f;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int k = k0 + b_k + j;
          bv[j] = (oc < g.OC && k < k_end) ? __ldg(wgt + (int64_t)oc * g.K + k) : 0.f;
        }
      }
    } else {
      b_k = tid / 16; b_n = (tid % 16) * 8;
      const int k = k0 + b_k;
      if (k < k_end) {
        const int tap = k / g.IC, ic = k - tap * g.IC;  // ic indexes the ORIGINAL Co
        const int64_t base = ((int64_t)ic * g.R * g.S + tap) * g.OC + n0 + b_n;
        if (VEC && n0 + b_n + 8 <= g.OC) {
          const float4* p = reinterpret_cast<const float4*>(wgt + base);
          float4 v0 = __ldg(p), v1 = __ldg(p + 1);
          bv[0] = v0.x; bv[1] = v0.y; bv[2] = v0.z; bv[3] = v0.w;
          bv[4] = v1.x; bv[5] = v1.y; bv[6] = v1.z; bv[7] = v1.w;
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) bv[j] = (n0 + b_n + j < g.OC) ? __ldg(wgt + base + j) : 0.f;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) bv[j] = 0.f;
      }
    }
    __syncthreads();  // previous iteration's compute is done
#pragma unroll
    for (int j = 0; j < 8; ++j) As[a_kq + j][a_m] = av[j];
    if (MODE == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) Bs[b_k + j][b_n] = bv[j];
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) Bs[b_k][b_n + j] = bv[j];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 8]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][tx * 8 + 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
  // ---- epilogue ----
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + ty * 8 + i;
    if (m >= g.M) continue;
    float* orow = out + (int64_t)m * g.OC;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + tx * 8 + j;
      if (n < g.OC) {
        float v = alpha * acc[i][j];
        if (gridDim.z > 1) {
          atomicAdd(orow + n, v);  // bias / activation are applied by a second pass (see launch_igemm)
        } else {
          if (bias != nullptr) v += bias_scale * __ldg(bias + n);
          orow[n] = act_apply(v, act, slope);
        }
      }
    }
  }
}

// wgrad: gw[co][tap][ci] += alpha * sum_{p in split} gy[p][co] * x[p + tap][ci]
struct WGeom {
  int N, H, W, Ci, Co, R, S, pad, OH, OW;
  int P;        // N*OH*OW
  int psplit;   // pixels per split (multiple of BK)
};

__global__ void __launch_bounds__(NT) wgrad_kernel(const float* __restrict__ x, const float* __restrict__ gy,
                                                   float* __restrict__ gw, WGeom g, float alpha) {
  __shared__ __align__(16) float As[BK][BM + 4];  // [pixel][co]
  __shared__ __align__(16) float Bs[BK][BN + 4];  // [pixel][ci]
  const int tid = threadIdx.x;
  const int co0 = blockIdx.x * BM, ci0 = blockIdx.y * BN;
  const int taps = g.R * g.S;
  const int tap = blockIdx.z % taps, split = blockIdx.z / taps;
  const int r = tap / g.S, s = tap - r * g.S;
  const int p_begin = split * g.psplit;
  const int p_end = min(g.P, p_begin + g.psplit);

  const int l_k = tid / 16, l_q = (tid % 16) * 8;  // this thread loads pixel row l_k, columns [l_q, l_q+8)
  const bool vecA = (g.Co % 4 == 0), vecB = (g.Ci % 4 == 0);
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  const int tx = tid % 16, ty = tid / 16;

  for (int p0 = p_begin; p0 < p_end; p0 += BK) {
    const int p = p0 + l_k;
    float av[8], bv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { av[j] = 0.f; bv[j] = 0.f; }
    if (p < p_end) {
      const float* ap = gy + (int64_t)p * g.Co + co0 + l_q;
      if (vecA && co0 + l_q + 8 <= g.Co) {
        float4 v0 = __ldg(reinterpret_cast<const float4*>(ap)), v1 = __ldg(reinterpret_cast<const float4*>(ap) + 1);
        av[0] = v0.x; av[1] = v0.y; av[2] = v0.z; av[3] = v0.w; av[4] = v1.x; av[5] = v1.y; av[6] = v1.z; av[7] = v1.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) if (co0 + l_q + j < g.Co) av[j] = __ldg(ap + j);
      }
      const int ow = p % g.OW;
      const int t = p / g.OW;
      const int oh = t % g.OH, n = t / g.OH;
      const int ih = oh + r - g.pad, iw = ow + s - g.pad;
      if (ih >= 0 && ih < g.H && iw >= 0 && iw < g.W) {
        const float* bp = x + (((int64_t)n * g.H + ih) * g.W + iw) * g.Ci + ci0 + l_q;
        if (vecB && ci0 + l_q + 8 <= g.Ci) {
          float4 v0 = __ldg(reinterpret_cast<const float4*>(bp)), v1 = __ldg(reinterpret_cast<const float4*>(bp) + 1);
          bv[0] = v0.x; bv[1] = v0.y; bv[2] = v0.z; bv[3] = v0.w; bv[4] = v1.x; bv[5] = v1.y; bv[6] = v1.z; bv[7] = v1.w;
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) if (ci0 + l_q + j < g.Ci) bv[j] = __ldg(bp + j);
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) { As[l_k][l_q + j] = av[j]; Bs[l_k][l_q + j] = bv[j]; }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 8]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][tx * 8 + 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
  const bool single = (gridDim.z == taps);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int co = co0 + ty * 8 + i;
    if (co >= g.Co) continue;
    float* orow = gw + ((int64_t)co * taps + tap) * g.Ci;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int ci = ci0 + tx * 8 + j;
      if (ci < g.Ci) {
        if (single) orow[ci] = alpha * acc[i][j];
        else atomicAdd(orow + ci, alpha * acc[i][j]);
      }
    }
  }
}

__global__ void weight_transpose_kernel(const float* __restrict__ w, float* __restrict__ wt, int Co, int R, int S, int Ci) {
  // wt[ci][R-1-r][S-1-s][co] = w[co][r][s][ci];   32x32 smem tile transpose over (co, ci) for each tap
  __shared__ float tile[32][33];
  const int tap = blockIdx.z, r = tap / S, s = tap - r * S;
  const int ftap = (R - 1 - r) * S + (S - 1 - s);
  const int co0 = blockIdx.y * 32, ci0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int co = co0 + i, ci = ci0 + threadIdx.x;
    tile[i][threadIdx.x] = (co < Co && ci < Ci) ? w[((int64_t)co * R * S + tap) * Ci + ci] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int ci = ci0 + i, co = co0 + threadIdx.x;
    if (co < Co && ci < Ci) wt[((int64_t)ci * R * S + ftap) * Co + co] = tile[threadIdx.x][i];
  }
}

}  // namespace

// Few output tiles and a long reduction (low-resolution layers, e.g. the 513-channel conv after the minibatch-stddev
// concat: 128 pixels x K = 4617): split K across gridDim.z so the SMs are busy; slices are summed with atomicAdd into
// the zeroed output and bias / activation run as a second (tiny) elementwise pass.
template <int MODE>
int launch_igemm(const float* in, const float* w, const float* bias, float* out, IGemm g, int n_out, bool vec, float alpha,
                 float bias_scale, int act, float slope, cudaStream_t st) {
  dim3 grid((g.M + BM - 1) / BM, (n_out + BN - 1) / BN, 1);
  const int tiles = grid.x * grid.y;
  int splits = 1;
  const bool epilogue_pass = (bias != nullptr || act != GLB_ACT_NONE);
  if (tiles < kNumSMs / 2 && g.K >= 16 * BK && (!epilogue_pass || n_out % 4 == 0)) {
    splits = (2 * kNumSMs + tiles - 1) / tiles;
    const int max_splits = g.K / (4 * BK);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
  }
  g.kper = (((g.K + splits - 1) / splits + BK - 1) / BK) * BK;
  splits = (g.K + g.kper - 1) / g.kper;
  grid.z = splits;
  if (splits > 1) GLB_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)g.M * n_out, st));
  if (vec)
    igemm_kernel<MODE, true><<<grid, NT, 0, st>>>(in, w, bias, out, g, alpha, bias_scale, act, slope);
  else
    igemm_kernel<MODE, false><<<grid, NT, 0, st>>>(in, w, bias, out, g, alpha, bias_scale, act, slope);
  GLB_CHECK_LAUNCH("igemm_kernel");
  if (splits > 1 && epilogue_pass)
    return glb_bias_act_fwd(out, bias, out, g.M, n_out, bias_scale, act, slope, (glb_stream_t)st);
  return GLB_OK;
}

int conv_fprop_simt(const float* x, const float* w, const float* bias, float* y, int N, int H, int W, int Ci, int Co,
                    int R, int S, int pad, float alpha, float bias_scale, int act, float slope, cudaStream_t st) {
  IGemm g;
  g.N = N; g.IH = H; g.IW = W; g.IC = Ci; g.OH = H + 2 * pad - R + 1; g.OW = W + 2 * pad - S + 1; g.OC = Co;
  g.R = R; g.S = S; g.dh0 = -pad; g.dw0 = -pad; g.dsign = 1;
  if (g.OH <= 0 || g.OW <= 0) return shape_fail("conv output size <= 0");
  g.M = N * g.OH * g.OW; g.K = R * S * Ci;
  return launch_igemm<0>(x, w, bias, y, g, Co, Ci % 16 == 0, alpha, bias_scale, act, slope, st);
}

int conv_dgrad_simt(const float* gy, const float* w, float* gx, int N, int H, int W, int Ci, int Co, int R, int S,
                    int pad, float alpha, cudaStream_t st) {
  IGemm g;
  const int OH = H + 2 * pad - R + 1, OW = W + 2 * pad - S + 1;
  if (OH <= 0 || OW <= 0) return shape_fail("conv output size <= 0");
  g.N = N; g.IH = OH; g.IW = OW; g.IC = Co; g.OH = H; g.OW = W; g.OC = Ci;
  g.R = R; g.S = S; g.dh0 = pad; g.dw0 = pad; g.dsign = -1;
  g.M = N * H * W; g.K = R * S * Co;
  return launch_igemm<1>(gy, w, nullptr, gx, g, Ci, Co % 16 == 0 && Ci % 4 == 0, alpha, 0.f, GLB_ACT_NONE, 0.f, st);
}

int conv_wgrad_simt(const float* x, const float* gy, float* gw, int N, int H, int W, int Ci, int Co, int R, int S,
                    int pad, float alpha, cudaStream_t st) {
  WGeom g;
  g.N = N; g.H = H; g.W = W; g.Ci = Ci; g.Co = Co; g.R = R; g.S = S; g.pad = pad;
  g.OH = H + 2 * pad - R + 1; g.OW = W + 2 * pad - S + 1;
  if (g.OH <= 0 || g.OW <= 0) return shape_fail("conv output size <= 0");
  g.P = N * g.OH * g.OW;
  const int tiles = ((Co + BM - 1) / BM) * ((Ci + BN - 1) / BN) * R * S;
  int splits = (2 * kNumSMs + tiles - 1) / tiles;
  const int max_splits = (g.P + BK - 1) / BK;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int psplit = (g.P + splits - 1) / splits;
  psplit = ((psplit + BK - 1) / BK) * BK;
  splits = (g.P + psplit - 1) / psplit;
  g.psplit = psplit;
  if (splits > 1) GLB_CUDA(cudaMemsetAsync(gw, 0, sizeof(float) * (size_t)Co * R * S * Ci, st));
  dim3 grid((Co + BM - 1) / BM, (Ci + BN - 1) / BN, R * S * splits);
  wgrad_kernel<<<grid, NT, 0, st>>>(x, gy, gw, g, alpha);
  GLB_CHECK_LAUNCH("wgrad_kernel");
  return GLB_OK;
}

}  // namespace glb

extern "C" int glb_conv2d_weight_transpose(const float* w, float* wt, int Co, int R, int S, int Ci, glb_stream_t stream) {
  dim3 grid((Ci + 31) / 32, (Co + 31) / 32, R * S), block(32, 8);
  glb::weight_transpose_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(w, wt, Co, R, S, Ci);
  GLB_CHECK_LAUNCH("weight_transpose_kernel");
  return GLB_OK;
}

