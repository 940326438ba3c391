// Microbenchmark: cycles per tcgen05.mma (kind::tf32, cta_group::1, M=128) issued back to back from fixed shared-memory
// operands, for K-major (128B swizzle) vs MN-major (128B swizzle, 32B atoms) operands and N = 64/128/256.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I gan_lab_b200/csrc -o tools/micro/mma_rate tools/micro/mma_rate.cu
#include <cstdio>
#include "tc_common.cuh"
namespace glb { void set_error(const std::string&) {} }
using namespace glb::tc;

template <int N, int MN>
__global__ void __launch_bounds__(128, 1) rate_kernel(long long* out, int iters, int ctas_active) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // fill operands with something finite
  float* f = reinterpret_cast<float*>(smem_raw + (base - raw));
  for (int i = threadIdx.x; i < (64 * 1024) / 4; i += blockDim.x) f[i] = 1.0f + (i & 7) * 0.125f;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&slot), 256);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 1 && lane == 0) {
    constexpr uint32_t idesc = make_idesc_tf32(128, N, MN, MN);
    const uint32_t a_addr = base, b_addr = base + 16384;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        uint64_t ad, bd;
        if (MN) {
          ad = make_smem_desc(a_addr + kk * 1024, 4096, 512, kLayoutSw128Base32);
          bd = make_smem_desc(b_addr + kk * 1024, 4096, 512, kLayoutSw128Base32);
        } else {
          ad = make_smem_desc(a_addr + kk * 32, 16, 1024);
          bd = make_smem_desc(b_addr + kk * 32, 16, 1024);
        }
        mma_tf32(tmem, ad, bd, idesc, 1u);
      }
    }
    mma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

// Same MMA stream, but with the real kernels' per-stage protocol around every group of GROUP MMAs: a parity wait on an
// mbarrier (here: the commit of the group issued RING groups ago, i.e. normally already complete) + fence + tcgen05.commit.
template <int N, int GROUP>
__global__ void __launch_bounds__(128, 1) proto_kernel(long long* out, int iters) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  constexpr int RING = 6;
  __shared__ uint64_t bars[RING];
  __shared__ uint64_t done;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* f = reinterpret_cast<float*>(smem_raw + (base - raw));
  for (int i = threadIdx.x; i < (64 * 1024) / 4; i += blockDim.x) f[i] = 1.0f + (i & 7) * 0.125f;
  if (threadIdx.x == 0) {
    for (int i = 0; i < RING; ++i) mbar_init(smem_u32(&bars[i]), 1);
    mbar_init(smem_u32(&done), 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&slot), 256);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 1 && lane == 0) {
    constexpr uint32_t idesc = make_idesc_tf32(128, N, 0, 0);
    const uint32_t a_addr = base, b_addr = base + 16384;
    int stage = 0; uint32_t phase = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      mbar_wait(smem_u32(&bars[stage]), phase ^ 1);
      tc_fence_after();
#pragma unroll
      for (int kk = 0; kk < GROUP; ++kk) {
        const uint64_t ad = make_smem_desc(a_addr + (kk & 3) * 32, 16, 1024);
        const uint64_t bd = make_smem_desc(b_addr + (kk & 3) * 32, 16, 1024);
        mma_tf32(tmem, ad, bd, idesc, 1u);
      }
      mma_commit(smem_u32(&bars[stage]));
      if (++stage == RING) { stage = 0; phase ^= 1; }
    }
    mma_commit(smem_u32(&done));
    mbar_wait(smem_u32(&done), 0);
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

template <int N, int GROUP>
void run_proto(const char* name) {
  long long* d; cudaMalloc(&d, sizeof(long long) * 148);
  const int iters = 2000, smem = 64 * 1024 + 2048;
  cudaFuncSetAttribute(proto_kernel<N, GROUP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  proto_kernel<N, GROUP><<<148, 128, smem>>>(d, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, d, sizeof(long long), cudaMemcpyDeviceToHost);
  double cyc = (double)h / (iters * GROUP);
  printf("%-40s %7.1f cycles / MMA (ideal %d)  [%s]\n", name, cyc, N / 2, cudaGetErrorString(e));
  cudaFree(d);
}

template <int N, int MN>
void run(const char* name, int grid) {
  long long* d; cudaMalloc(&d, sizeof(long long) * grid);
  const int iters = 2000, smem = 64 * 1024 + 2048;
  cudaFuncSetAttribute(rate_kernel<N, MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  rate_kernel<N, MN><<<grid, 128, smem>>>(d, iters, grid);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148] = {0};
  cudaMemcpy(h, d, sizeof(long long) * (grid < 148 ? grid : 148), cudaMemcpyDeviceToHost);
  double cyc = (double)h[0] / (iters * 4);
  printf("%-28s grid %3d: %7.1f cycles / MMA (128x%dx8)  -> %6.0f MAC/cycle/SM  ideal-at-2048 = %d cyc  [%s]\n", name, grid, cyc, N,
         128.0 * N * 8 / cyc, N / 2, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int grid : {148}) {
    run<256, 0>("K-major  N=256", grid);
    run<128, 0>("K-major  N=128", grid);
    run<64, 0>("K-major  N=64", grid);
    run<256, 1>("MN-major N=256", grid);
    run<128, 1>("MN-major N=128", grid);
    run<64, 1>("MN-major N=64", grid);
  }
  run_proto<256, 4>("wait+commit per 4 MMAs, N=256");
  run_proto<128, 4>("wait+commit per 4 MMAs, N=128");
  run_proto<128, 8>("wait+commit per 8 MMAs, N=128");
  run_proto<64, 4>("wait+commit per 4 MMAs, N=64");
  run_proto<64, 8>("wait+commit per 8 MMAs, N=64");
  return 0;
}
