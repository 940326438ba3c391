// Shared helpers for the ganlab_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/ganlab_b200.h"

namespace glb {

void set_error(const std::string& msg);

inline int cuda_fail(cudaError_t e, const char* what) {
  set_error(std::string(what) + ": " + cudaGetErrorString(e));
  return GLB_ERR_CUDA;
}

inline int shape_fail(const char* what) {
  set_error(std::string("bad shape/argument: ") + what);
  return GLB_ERR_SHAPE;
}

#define GLB_CHECK_LAUNCH(name)                                  \
  do {                                                          \
    cudaError_t e__ = cudaGetLastError();                       \
    if (e__ != cudaSuccess) return glb::cuda_fail(e__, name);   \
  } while (0)

#define GLB_CUDA(call)                                          \
  do {                                                          \
    cudaError_t e__ = (call);                                   \
    if (e__ != cudaSuccess) return glb::cuda_fail(e__, #call);  \
  } while (0)

constexpr int kNumSMs = 148;  // B200

__device__ __forceinline__ float act_apply(float v, int act, float slope) {
  return (act == GLB_ACT_LRELU) ? (v > 0.f ? v : v * slope) : v;
}
// derivative selected by the sign of the saved OUTPUT (slope > 0 keeps the sign; for ReLU y>0 <=> pre>0)
__device__ __forceinline__ float act_grad(float y, int act, float slope) {
  return (act == GLB_ACT_LRELU) ? (y > 0.f ? 1.f : slope) : 1.f;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// streaming 128-bit accesses (read-once / write-once activations: keep them out of L1)
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

inline int grid_for(int64_t work_items, int threads, int max_blocks = kNumSMs * 16) {
  int64_t b = (work_items + threads - 1) / threads;
  if (b < 1) b = 1;
  if (b > max_blocks) b = max_blocks;
  return (int)b;
}

}  // namespace glb
