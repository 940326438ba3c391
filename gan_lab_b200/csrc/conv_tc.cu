// tcgen05 / TMEM / TMA implicit-GEMM convolution family (TF32).  Placeholder until the kernels land.
#include "common.cuh"
namespace glb {
int conv_fprop_tc(const float*, const float*, const float*, float*, int, int, int, int, int, int, int, int, float, float, int,
                  float, cudaStream_t) { set_error("tcgen05 fprop: shape not covered"); return GLB_ERR_UNSUPPORTED; }
int conv_dgrad_tc(const float*, const float*, float*, int, int, int, int, int, int, int, int, float, cudaStream_t) {
  set_error("tcgen05 dgrad: shape not covered"); return GLB_ERR_UNSUPPORTED; }
int conv_wgrad_tc(const float*, const float*, float*, int, int, int, int, int, int, int, int, float, cudaStream_t) {
  set_error("tcgen05 wgrad: shape not covered"); return GLB_ERR_UNSUPPORTED; }
}  // namespace glb
