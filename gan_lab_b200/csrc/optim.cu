// Fused multi-tensor Adam with the EWMA ("lagged") generator update folded into the same pass.
// Reference: torch.optim.Adam as configured by utils/backprop_utils.py:109-120 (foreach lerp_/addcmul_/sqrt/addcdiv_
// kernels per step) and the per-parameter EWMA loop progan/learner.py:909-916 (3 allocating ATen kernels per
// parameter per step).  HBM-bound: 28 B/param for Adam (+8 B/param when the EWMA copy rides along).
#include "common.cuh"

namespace glb {
namespace {

// ptrs: [T][5] = p, g, m, v, lagged ; hyper: lr, bc1 = 1-b1^t, bc2 = 1-b2^t
// ewma_mode: 0 = off, 1 = lagged = p*(1-b)+lagged*b, 2 = first step: lagged = p*(1-b)+p*b (progan/learner.py:472 aliasing)
__global__ void adam_ewma_kernel(const void* const* __restrict__ ptrs, const int64_t* __restrict__ sizes,
                                 const float* __restrict__ hyper, float b1, float b2, float eps, float wd, float eb, int ewma_mode) {
  const int t = blockIdx.y;
  const int64_t n = sizes[t];
  float* __restrict__ p = (float*)ptrs[5 * t + 0];
  const float* __restrict__ g = (const float*)ptrs[5 * t + 1];
  float* __restrict__ m = (float*)ptrs[5 * t + 2];
  float* __restrict__ v = (float*)ptrs[5 * t + 3];
  float* __restrict__ lag = (float*)ptrs[5 * t + 4];
  const float lr = hyper[0], bc1 = hyper[1], bc2 = hyper[2];
  const float step_size = lr / bc1, inv_bc2_sqrt = rsqrtf(bc2);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (g == nullptr) {  // EWMA-only entry (parameter outside the optimiser or without a gradient this step)
    if (ewma_mode != 0 && lag != nullptr)
      for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
        const float pi = p[i];
        lag[i] = pi * (1.f - eb) + ((ewma_mode == 2) ? pi : lag[i]) * eb;
      }
    return;
  }
  auto upd = [&](float gi, float& pi, float& mi, float& vi) {
    if (wd != 0.f) gi = fmaf(wd, pi, gi);
    mi = (1.f - b1 >= 0.5f) ? gi - (gi - mi) * b1 : mi + (gi - mi) * (1.f - b1);   // torch lerp_ formula
    vi = vi * b2 + (1.f - b2) * gi * gi;
    const float denom = sqrtf(vi) * inv_bc2_sqrt + eps;
    pi = pi - step_size * (mi / denom);
  };
  const bool with_lag = ewma_mode != 0 && lag != nullptr;
  // 128-bit path when every stream of this tensor is 16-byte aligned (always the case for torch allocations)
  const bool vec = ((((uintptr_t)p) | ((uintptr_t)g) | ((uintptr_t)m) | ((uintptr_t)v) | ((uintptr_t)lag)) & 15) == 0;
  int64_t done = 0;
  if (vec) {
    const int64_t n4 = n >> 2;
    float4* p4 = (float4*)p; const float4* g4 = (const float4*)g; float4* m4 = (float4*)m; float4* v4 = (float4*)v;
    float4* l4 = (float4*)lag;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
      const float4 gv = ldg_stream(g4 + i);
      float4 pv = p4[i], mv = m4[i], vv = v4[i];
      float4 lv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (with_lag && ewma_mode != 2) lv = l4[i];
      upd(gv.x, pv.x, mv.x, vv.x); upd(gv.y, pv.y, mv.y, vv.y); upd(gv.z, pv.z, mv.z, vv.z); upd(gv.w, pv.w, mv.w, vv.w);
      p4[i] = pv; m4[i] = mv; v4[i] = vv;
      if (with_lag) {
        if (ewma_mode == 2) lv = pv;
        l4[i] = make_float4(pv.x * (1.f - eb) + lv.x * eb, pv.y * (1.f - eb) + lv.y * eb, pv.z * (1.f - eb) + lv.z * eb,
                            pv.w * (1.f - eb) + lv.w * eb);
      }
    }
    done = n4 << 2;
  }
  for (int64_t i = done + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
    float pi = p[i], mi = m[i], vi = v[i];
    upd(g[i], pi, mi, vi);
    p[i] = pi; m[i] = mi; v[i] = vi;
    if (with_lag) {
      const float prev = (ewma_mode == 2) ? pi : lag[i];
      lag[i] = pi * (1.f - eb) + prev * eb;
    }
  }
}

// hyper = [lr, 1-b1^t, 1-b2^t, t]: advance t by one and refresh the bias corrections on the device, so that a captured
// (CUDA graph) optimiser step replays with the right step count (torch.optim.Adam computes these on the host).
__global__ void adam_hyper_advance_kernel(float* __restrict__ hyper, float b1, float b2) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const float t = hyper[3] + 1.f;
    hyper[3] = t;
    hyper[1] = (float)(1.0 - pow((double)b1, (double)t));
    hyper[2] = (float)(1.0 - pow((double)b2, (double)t));
  }
}

}  // namespace
}  // namespace glb

extern "C" int glb_adam_hyper_advance(float* hyper, float beta1, float beta2, glb_stream_t stream) {
  glb::adam_hyper_advance_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(hyper, beta1, beta2);
  GLB_CHECK_LAUNCH("adam_hyper_advance_kernel");
  return GLB_OK;
}

extern "C" int glb_adam_ewma_multi(const void* const* ptrs, const int64_t* sizes, int T, int64_t max_size, const float* hyper,
                                   float beta1, float beta2, float eps, float wd, float ewma_beta, int ewma_mode,
                                   glb_stream_t stream) {
  if (T <= 0) return GLB_OK;
  if (T > 65535) return glb::shape_fail("adam: more than 65535 tensors");
  int bx = (int)((max_size + 256 * 16 - 1) / (256 * 16));   // >= 4 float4 per thread for the largest tensor
  if (bx < 1) bx = 1;
  if (bx > 96) bx = 96;
  dim3 grid(bx, T);
  glb::adam_ewma_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(ptrs, sizes, hyper, beta1, beta2, eps, wd, ewma_beta, ewma_mode);
  GLB_CHECK_LAUNCH("adam_ewma_kernel");
  return GLB_OK;
}
