// Measured tcgen05 peak of the chip: every SM issues back-to-back tcgen05.mma (cta_group::1, M = 128, N = 256) from fixed
// shared-memory operands; CUDA events around the launch -> TFLOP/s for kind::tf32 (K = 8) and kind::f16 with bf16 operands
// (K = 16), as a burst (~30 ms) and sustained (~3 s, under the power cap).  This is the denominator a TF32 / bf16 implicit
// GEMM can be held against on THIS chip (MEASURED_PEAKS.json only has the cuBLAS bf16 figure).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I gan_lab_b200/csrc -I include -o tools/micro/mma_peak tools/micro/mma_peak.cu
#include <cstdio>
#include <string>
#include "tc_common.cuh"
namespace glb { void set_error(const std::string&) {} }
using namespace glb::tc;

__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// kind::f16 instruction descriptor: fp32 accumulate, bf16 A and B, K-major
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int BF16>
__global__ void __launch_bounds__(128, 1) peak_kernel(int iters) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t* f = reinterpret_cast<uint32_t*>(smem_raw + (base - raw));
  // finite, non-trivial operand bits in either format (fp32 1.0 .. 1.9 / bf16 pairs around 1.0)
  for (int i = threadIdx.x; i < (48 * 1024) / 4; i += blockDim.x) f[i] = BF16 ? (0x3F803F80u + ((i & 7) << 16) + (i & 3)) : (0x3F800000u + ((i & 7) << 20));
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&slot), 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 1 && lane == 0) {
    constexpr uint32_t idesc = BF16 ? make_idesc_bf16(128, 256) : make_idesc_tf32(128, 256, 0, 0);
    const uint32_t a_addr = base, b_addr = base + 16384;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const uint64_t ad = make_smem_desc(a_addr + kk * 32, 16, 1024);
        const uint64_t bd = make_smem_desc(b_addr + kk * 32, 16, 1024);
        if (BF16) mma_f16(tmem + (it & 1) * 256, ad, bd, idesc, 1u);
        else mma_tf32(tmem + (it & 1) * 256, ad, bd, idesc, 1u);
      }
    }
    mma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int BF16>
double run(int iters, int sms) {
  const int smem = 48 * 1024 + 2048;
  cudaFuncSetAttribute(peak_kernel<BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  peak_kernel<BF16><<<sms, 128, smem>>>(1000);     // warm-up
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  peak_kernel<BF16><<<sms, 128, smem>>>(iters);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double flop = 2.0 * 128 * 256 * (BF16 ? 16 : 8) * 4.0 * iters * sms;
  const double tf = flop / (ms * 1e-3) / 1e12;
  printf("%-5s iters %9d: %9.2f ms  %8.1f TFLOP/s  (%.1f cycles/MMA at 1965 MHz)  [%s]\n", BF16 ? "bf16" : "tf32", iters, ms, tf,
         ms * 1e-3 * 1.965e9 / (4.0 * iters), cudaGetErrorString(e));
  return tf;
}

int main() {
  cudaDeviceProp pr;
  cudaGetDeviceProperties(&pr, 0);
  const int sms = pr.multiProcessorCount;
  printf("device %s, %d SMs\n", pr.name, sms);
  double r[4];
  r[0] = run<0>(100000, sms);     // ~30 ms burst
  r[1] = run<1>(100000, sms);
  r[2] = run<0>(10000000, sms);   // ~3 s sustained
  r[3] = run<1>(10000000, sms);
  printf("{\"tf32_tflops_burst\": %.1f, \"bf16_mma_tflops_burst\": %.1f, \"tf32_tflops_sustained\": %.1f, \"bf16_mma_tflops_sustained\": %.1f}\n",
         r[0], r[1], r[2], r[3]);
  return 0;
}
