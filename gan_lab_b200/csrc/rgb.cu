// RGB <-> feature 1x1 convolutions (3 channels on one side): pure bandwidth, never a GEMM.
// fromRGB: reference progan/architectures.py:286-292 (Conv2dEx 1x1 3->C + LeakyReLU); toRGB: stylegan/architectures.py:338-341
// and progan/architectures.py:150-153 (Conv2dEx 1x1 C->3, gain_sq_base=1).  Images are NCHW (the reference's layout at the
// model boundary), features NHWC; the fade-in skip branches (avg-pool of the image before prev_fromrgb,
// progan/architectures.py:312) are folded in through `pool`.
#include "common.cuh"

namespace glb {
namespace {

constexpr int TPB = 256;

// y[n,h,w,c] = act(alpha * sum_j img_j * w(j,c) + bias_scale*bias[c]);  H, W = OUTPUT size; img is [N,3,H*(1+pool),W*(1+pool)]
__global__ void rgb_expand_kernel(const float* __restrict__ img, const float* __restrict__ w, int ws_j, int ws_c,
                                  const float* __restrict__ bias, float4* __restrict__ y, int N, int H, int W, int C4, int pool,
                                  float alpha, float bias_scale, int act, float slope) {
  // blockDim is a multiple of C4 (host), so a thread's channel quad is loop invariant: its 12 weights (alpha folded in) and 4
  // biases live in registers; per output quad the loop then costs 3 broadcast image loads + 12 FMAs + one 128-bit store.
  const int q = threadIdx.x % C4;
  float wr[3][4], br[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = 4 * q + k;
#pragma unroll
    for (int j = 0; j < 3; ++j) wr[j][k] = alpha * __ldg(w + j * ws_j + c * ws_c);
    br[k] = bias != nullptr ? bias_scale * __ldg(bias + c) : 0.f;
  }
  const int ppb = blockDim.x / C4;                 // pixels per block and iteration
  const int64_t P = (int64_t)N * H * W;
  const int IH = pool ? 2 * H : H, IW = pool ? 2 * W : W;
  const int64_t plane = (int64_t)IH * IW, HW = (int64_t)H * W;
  constexpr int EU = 4;                                 // pixels per thread and iteration: their image loads are independent
  const int64_t stride = (int64_t)gridDim.x * ppb;
  for (int64_t p0 = (int64_t)blockIdx.x * ppb + threadIdx.x / C4; p0 < P; p0 += EU * stride) {
    float v[EU][3];
#pragma unroll
    for (int u = 0; u < EU; ++u) {
      const int64_t p = p0 + u * stride;
      if (p < P) {
        const int64_t n = p / HW, r = p - n * HW;
        if (pool) {
          const int hh = (int)(r / W), ww = (int)(r - (int64_t)hh * W);
          const int64_t o = (int64_t)(2 * hh) * IW + 2 * ww;
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const float* ip = img + (n * 3 + j) * plane;
            v[u][j] = 0.25f * (__ldg(ip + o) + __ldg(ip + o + 1) + __ldg(ip + o + IW) + __ldg(ip + o + IW + 1));
          }
        } else {
#pragma unroll
          for (int j = 0; j < 3; ++j) v[u][j] = __ldg(img + (n * 3 + j) * plane + r);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < EU; ++u) {
      const int64_t p = p0 + u * stride;
      if (p < P) {
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
          o[k] = act_apply(fmaf(v[u][0], wr[0][k], fmaf(v[u][1], wr[1][k], fmaf(v[u][2], wr[2][k], br[k]))), act, slope);
        stg_stream(y + p * C4 + q, make_float4(o[0], o[1], o[2], o[3]));
      }
    }
  }
}

// img[n,j,h,w] = alpha * sum_c x[n,h,w,c] * w(j,c) + bias_scale*bias[j]; GS lanes cooperate on one pixel.
// pool: scatter 0.25*value into the 2x2 window of an image of twice the size (adjoint of the avg-pool).
template <int GS, int QPL>   // QPL = quads per lane held in registers (0: read the weights through L1 every time)
__global__ void rgb_contract_kernel(const float4* __restrict__ x, const float* __restrict__ w, int ws_j, int ws_c,
                                    const float* __restrict__ bias, float* __restrict__ img, int N, int H, int W, int C4, int pool,
                                    float alpha, float bias_scale) {
  // PX pixels per group and iteration: PX * QPL independent 16-byte loads in flight per lane (one pixel per iteration left the
  // kernel latency bound at ~2 TB/s)
  constexpr int PX = (QPL > 0 && QPL <= 2) ? 4 : (QPL == 4 ? 2 : 1);
  const int64_t P = (int64_t)N * H * W;
  const int lane = threadIdx.x % GS;
  const int64_t grp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / GS, ngrp = ((int64_t)gridDim.x * blockDim.x) / GS;
  // the pixel range is padded up so every lane of a warp takes part in the shuffles
  const int64_t per = (32 / GS) * PX;
  const int64_t Ppad = ((P + per - 1) / per) * per;
  float wr[QPL > 0 ? QPL : 1][3][4];
  if (QPL > 0) {
#pragma unroll
    for (int i = 0; i < QPL; ++i)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int c = 4 * (lane + i * GS) + k;
#pragma unroll
        for (int j = 0; j < 3; ++j) wr[i][j][k] = (lane + i * GS < C4) ? __ldg(w + j * ws_j + c * ws_c) : 0.f;
      }
  }
  for (int64_t pbase = grp * PX; pbase < Ppad; pbase += ngrp * PX) {
    float acc[PX][3];
#pragma unroll
    for (int u = 0; u < PX; ++u) acc[u][0] = acc[u][1] = acc[u][2] = 0.f;
    if (QPL > 0) {
      float4 v[PX][QPL > 0 ? QPL : 1];
#pragma unroll
      for (int u = 0; u < PX; ++u)
#pragma unroll
        for (int i = 0; i < QPL; ++i)
          v[u][i] = (pbase + u < P && lane + i * GS < C4) ? ldg_stream(x + (pbase + u) * C4 + lane + i * GS) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < PX; ++u)
#pragma unroll
        for (int i = 0; i < QPL; ++i) {
          const float vv[4] = {v[u][i].x, v[u][i].y, v[u][i].z, v[u][i].w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            acc[u][0] = fmaf(vv[k], wr[i][0][k], acc[u][0]);
            acc[u][1] = fmaf(vv[k], wr[i][1][k], acc[u][1]);
            acc[u][2] = fmaf(vv[k], wr[i][2][k], acc[u][2]);
          }
        }
    } else if (pbase < P) {
      for (int q = lane; q < C4; q += GS) {
        const float4 v = ldg_stream(x + pbase * C4 + q);
        const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int c = 4 * q + k;
          acc[0][0] = fmaf(vv[k], __ldg(w + 0 * ws_j + c * ws_c), acc[0][0]);
          acc[0][1] = fmaf(vv[k], __ldg(w + 1 * ws_j + c * ws_c), acc[0][1]);
          acc[0][2] = fmaf(vv[k], __ldg(w + 2 * ws_j + c * ws_c), acc[0][2]);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < PX; ++u) {
#pragma unroll
      for (int o = GS / 2; o > 0; o >>= 1) {
        acc[u][0] += __shfl_xor_sync(0xffffffffu, acc[u][0], o);
        acc[u][1] += __shfl_xor_sync(0xffffffffu, acc[u][1], o);
        acc[u][2] += __shfl_xor_sync(0xffffffffu, acc[u][2], o);
      }
      const int64_t p = pbase + u;
      if (p < P && lane < 3) {
        float a = lane == 0 ? acc[u][0] : (lane == 1 ? acc[u][1] : acc[u][2]);
        a *= alpha;
        if (bias != nullptr) a += bias_scale * __ldg(bias + lane);
        const int ww = (int)(p % W);
        int64_t t = p / W;
        const int hh = (int)(t % H);
        const int64_t n = t / H;
        if (pool) {
          const int OW = 2 * W;
          float* o = img + ((n * 3 + lane) * (2 * H) + 2 * hh) * (int64_t)OW + 2 * ww;
          a *= 0.25f;
          o[0] = a; o[1] = a; o[OW] = a; o[OW + 1] = a;
        } else {
          img[((n * 3 + lane) * H + hh) * (int64_t)W + ww] = a;
        }
      }
    }
  }
}

// gw(j,c) += alpha * sum_p img_j[p] * g[p,c]; thread owns a channel quad, rows strided; block reduce then atomics.
__global__ void rgb_wgrad_kernel(const float* __restrict__ img, const float4* __restrict__ g, float* __restrict__ gw, int ws_j,
                                 int ws_c, int N, int H, int W, int C4, int pool, float alpha) {
  extern __shared__ float4 red[];  // 3 * blockDim
  const int tid = threadIdx.x;
  const int lanes = min(C4, (int)blockDim.x), rows = blockDim.x / lanes, rl = tid / lanes;
  const int IH = pool ? 2 * H : H, IW = pool ? 2 * W : W;
  const int64_t plane = (int64_t)IH * IW;
  const int HW = H * W;
  for (int q = tid % lanes; q < C4; q += lanes) {
    float4 s[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) s[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    // blockIdx.y = sample, blockIdx.x strides the pixels of its plane: 32-bit index math only in the inner loop
    for (int n = blockIdx.y; n < N; n += gridDim.y) {
      const float4* gn = g + (int64_t)n * HW * C4 + q;
      const float* in = img + (int64_t)n * 3 * plane;
      constexpr int WU = 4;                             // rows in flight per thread (one row per iteration: 1.3 TB/s, latency bound)
      const int stride = gridDim.x * rows;
      for (int r0 = blockIdx.x * rows + rl; r0 < HW; r0 += WU * stride) {
        float4 gv[WU];
        float v[WU][3];
#pragma unroll
        for (int u = 0; u < WU; ++u) {
          const int r = r0 + u * stride;
          if (r < HW) {
            gv[u] = ldg_stream(gn + (int64_t)r * C4);
            if (pool) {
              const int hh = r / W, ww = r - hh * W;
              const int o = (2 * hh) * IW + 2 * ww;
#pragma unroll
              for (int j = 0; j < 3; ++j) {
                const float* ip = in + j * plane;
                v[u][j] = 0.25f * (__ldg(ip + o) + __ldg(ip + o + 1) + __ldg(ip + o + IW) + __ldg(ip + o + IW + 1));
              }
            } else {
#pragma unroll
              for (int j = 0; j < 3; ++j) v[u][j] = __ldg(in + j * plane + r);
            }
          } else {
            gv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            v[u][0] = v[u][1] = v[u][2] = 0.f;
          }
        }
#pragma unroll
        for (int u = 0; u < WU; ++u)
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            s[j].x = fmaf(v[u][j], gv[u].x, s[j].x); s[j].y = fmaf(v[u][j], gv[u].y, s[j].y);
            s[j].z = fmaf(v[u][j], gv[u].z, s[j].z); s[j].w = fmaf(v[u][j], gv[u].w, s[j].w);
          }
      }
    }
    if (C4 > lanes) {  // wide: direct atomics
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        atomicAdd(gw + j * ws_j + (4 * q + 0) * ws_c, alpha * s[j].x); atomicAdd(gw + j * ws_j + (4 * q + 1) * ws_c, alpha * s[j].y);
        atomicAdd(gw + j * ws_j + (4 * q + 2) * ws_c, alpha * s[j].z); atomicAdd(gw + j * ws_j + (4 * q + 3) * ws_c, alpha * s[j].w);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 3; ++j) red[j * blockDim.x + tid] = s[j];
      __syncthreads();
      if (tid < C4) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          float4 t = red[j * blockDim.x + tid];
          for (int r = 1; r < rows; ++r) {
            const float4 u = red[j * blockDim.x + r * C4 + tid];
            t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
          }
          atomicAdd(gw + j * ws_j + (4 * tid + 0) * ws_c, alpha * t.x); atomicAdd(gw + j * ws_j + (4 * tid + 1) * ws_c, alpha * t.y);
          atomicAdd(gw + j * ws_j + (4 * tid + 2) * ws_c, alpha * t.z); atomicAdd(gw + j * ws_j + (4 * tid + 3) * ws_c, alpha * t.w);
        }
      }
    }
  }
}

// out[c] += scale * sum_{n,hw} img[n,c,hw]  (bias gradient of toRGB; NCHW planes)
__global__ void plane_sum_kernel(const float* __restrict__ img, float* __restrict__ out, int N, int C, int64_t HW, float scale) {
  __shared__ float red[32];
  const int c = blockIdx.y;
  float s = 0.f;
  for (int n = 0; n < N; ++n) {
    const float* p = img + ((int64_t)n * C + c) * HW;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < HW; i += (int64_t)gridDim.x * blockDim.x) s += p[i];
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
    s = warp_sum(s);
    if (threadIdx.x == 0) atomicAdd(out + c, scale * s);
  }
}

}  // namespace
}  // namespace glb

using namespace glb;
#define REQ(cond, msg) do { if (!(cond)) return glb::shape_fail(msg); } while (0)

extern "C" int glb_rgb_expand(const float* img, const float* w, int ws_j, int ws_c, const float* bias, float* y, int N, int H,
                              int W, int C, int pool, float alpha, float bias_scale, int act, float slope, glb_stream_t stream) {
  REQ(C % 4 == 0, "rgb_expand: C % 4 != 0");
  const int C4 = C / 4;
  REQ(C4 <= 1024, "rgb_expand: C > 4096");
  const int threads = C4 <= TPB ? (TPB / C4) * C4 : C4;            // a multiple of C4: the channel quad of a thread is fixed
  const int64_t P = (int64_t)N * H * W;
  const int ppb = threads / C4;
  rgb_expand_kernel<<<grid_for((P + ppb - 1) / ppb, 1), threads, 0, (cudaStream_t)stream>>>(img, w, ws_j, ws_c, bias, (float4*)y, N, H, W,
                                                                                          C4, pool, alpha, bias_scale, act, slope);
  GLB_CHECK_LAUNCH("rgb_expand");
  return GLB_OK;
}

extern "C" int glb_rgb_contract(const float* x, const float* w, int ws_j, int ws_c, const float* bias, float* img, int N, int H,
                                int W, int C, int pool, float alpha, float bias_scale, glb_stream_t stream) {
  REQ(C % 4 == 0, "rgb_contract: C % 4 != 0");
  const int C4 = C / 4;
  const int64_t P = (int64_t)N * H * W;
  cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH(GS, QPL)                                                                                                 \
  rgb_contract_kernel<GS, QPL><<<grid_for(P * GS, TPB), TPB, 0, st>>>((const float4*)x, w, ws_j, ws_c, bias, img, N, H, W, C4,  \
                                                                      pool, alpha, bias_scale)
  if (C4 == 32) LAUNCH(32, 1);
  else if (C4 == 64) LAUNCH(32, 2);
  else if (C4 == 128) LAUNCH(32, 4);
  else if (C4 > 32) LAUNCH(32, 0);
  else if (C4 == 16) LAUNCH(16, 1);
  else if (C4 >= 16) LAUNCH(16, 0);
  else if (C4 == 8) LAUNCH(8, 1);
  else if (C4 >= 8) LAUNCH(8, 0);
  else if (C4 == 4) LAUNCH(4, 1);
  else LAUNCH(4, 0);
#undef LAUNCH
  GLB_CHECK_LAUNCH("rgb_contract");
  return GLB_OK;
}

extern "C" int glb_rgb_wgrad(const float* img, const float* g, float* gw, int ws_j, int ws_c, int N, int H, int W, int C,
                             int pool, float alpha, glb_stream_t stream) {
  REQ(C % 4 == 0, "rgb_wgrad: C % 4 != 0");
  const int C4 = C / 4;
  REQ(C4 > TPB || TPB % C4 == 0, "rgb_wgrad: C/4 must divide 256 or exceed it");
  const int lanes = C4 < TPB ? C4 : TPB, rows = TPB / lanes;
  REQ((int64_t)H * W * 4 < (1ll << 31), "rgb_wgrad: plane too large");
  // every block ends in 3*C atomics on the same 3*C addresses (80 us of pure contention with 6 blocks per SM at C = 512): at most
  // ~2 blocks per SM, fewer when there is little to read (>= 8 rows of 4 per thread), spread as (plane chunks) x (samples)
  int by = N < 16 ? N : 16;
  int64_t want = ((int64_t)N * H * W * C4 + (int64_t)TPB * 32 - 1) / ((int64_t)TPB * 32);
  if (want > 2 * kNumSMs) want = 2 * kNumSMs;
  if (want < by) want = by;
  int bx = (int)((want + by - 1) / by);
  const int max_bx = (H * W + rows - 1) / rows;
  if (bx > max_bx) bx = max_bx;
  rgb_wgrad_kernel<<<dim3(bx, by), TPB, 3 * TPB * sizeof(float4), (cudaStream_t)stream>>>(img, (const float4*)g, gw, ws_j, ws_c, N, H, W,
                                                                                         C4, pool, alpha);
  GLB_CHECK_LAUNCH("rgb_wgrad");
  return GLB_OK;
}

extern "C" int glb_plane_sum(const float* img, float* out, int N, int C, int64_t HW, float scale, glb_stream_t stream) {
  REQ(N > 0 && C > 0 && HW > 0, "plane_sum");
  dim3 grid(grid_for(HW, TPB, 64), C);
  plane_sum_kernel<<<grid, TPB, 0, (cudaStream_t)stream>>>(img, out, N, C, HW, scale);
  GLB_CHECK_LAUNCH("plane_sum");
  return GLB_OK;
}
