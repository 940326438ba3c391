// C-ABI plumbing: error string, version, implementation dispatch of the convolution family.
#include "common.cuh"

namespace glb {
static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }

int conv_fprop_simt(const float*, const float*, const float*, float*, int, int, int, int, int, int, int, int, float, float, int,
                    float, cudaStream_t);
int conv_dgrad_simt(const float*, const float*, float*, int, int, int, int, int, int, int, int, float, cudaStream_t);
int conv_wgrad_simt(const float*, const float*, float*, int, int, int, int, int, int, int, int, float, cudaStream_t);
int conv_fprop_tc(const float*, const float*, const float*, float*, int, int, int, int, int, int, int, int, float, float, int,
                  float, cudaStream_t);
int conv_dgrad_tc(const float*, const float*, float*, int, int, int, int, int, int, int, int, float, cudaStream_t);
int conv_wgrad_tc(const float*, const float*, float*, int, int, int, int, int, int, int, int, float, cudaStream_t);
bool conv_fprop_tc_covers(int, int, int, int, int, int, int, int);
bool conv_dgrad_tc_covers(int, int, int, int, int, int, int, int);
bool conv_wgrad_tc_covers(int, int, int, int, int, int, int, int);
}  // namespace glb

extern "C" const char* glb_last_error(void) { return glb::g_last_error.c_str(); }
extern "C" int glb_version(void) { return 100; }

extern "C" int glb_tc_available(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}

extern "C" int glb_conv2d_fprop(const float* x, const float* w, const float* bias, float* y, int N, int H, int W, int Ci, int Co,
                                int R, int S, int pad, float alpha, float bias_scale, int act, float slope, int impl,
                                glb_stream_t stream) {
  if (N <= 0 || H <= 0 || W <= 0 || Ci <= 0 || Co <= 0 || R <= 0 || S <= 0 || pad < 0) return glb::shape_fail("conv2d_fprop");
  if (impl == GLB_IMPL_TF32)
    return glb::conv_fprop_tc(x, w, bias, y, N, H, W, Ci, Co, R, S, pad, alpha, bias_scale, act, slope, (cudaStream_t)stream);
  return glb::conv_fprop_simt(x, w, bias, y, N, H, W, Ci, Co, R, S, pad, alpha, bias_scale, act, slope, (cudaStream_t)stream);
}

extern "C" int glb_conv2d_tc_covers(int kind, int N, int H, int W, int Ci, int Co, int R, int S, int pad) {
  if (N <= 0 || H <= 0 || W <= 0 || Ci <= 0 || Co <= 0 || R <= 0 || S <= 0 || pad < 0) return 0;
  switch (kind) {
    case 0: return glb::conv_fprop_tc_covers(N, H, W, Ci, Co, R, S, pad) ? 1 : 0;
    case 1: return glb::conv_dgrad_tc_covers(N, H, W, Ci, Co, R, S, pad) ? 1 : 0;
    case 2: return glb::conv_wgrad_tc_covers(N, H, W, Ci, Co, R, S, pad) ? 1 : 0;
  }
  return 0;
}

extern "C" int glb_conv2d_dgrad(const float* gy, const float* w, const float* wt, float* gx, int N, int H, int W, int Ci, int Co,
                                int R, int S, int pad, float alpha, int impl, glb_stream_t stream) {
  if (N <= 0 || H <= 0 || W <= 0 || Ci <= 0 || Co <= 0 || R <= 0 || S <= 0 || pad < 0) return glb::shape_fail("conv2d_dgrad");
  if (impl == GLB_IMPL_TF32) return glb::conv_dgrad_tc(gy, wt, gx, N, H, W, Ci, Co, R, S, pad, alpha, (cudaStream_t)stream);
  return glb::conv_dgrad_simt(gy, w, gx, N, H, W, Ci, Co, R, S, pad, alpha, (cudaStream_t)stream);
}

extern "C" int glb_conv2d_wgrad(const float* x, const float* gy, float* gw, int N, int H, int W, int Ci, int Co, int R, int S,
                                int pad, float alpha, int impl, glb_stream_t stream) {
  if (N <= 0 || H <= 0 || W <= 0 || Ci <= 0 || Co <= 0 || R <= 0 || S <= 0 || pad < 0) return glb::shape_fail("conv2d_wgrad");
  if (impl == GLB_IMPL_TF32) return glb::conv_wgrad_tc(x, gy, gw, N, H, W, Ci, Co, R, S, pad, alpha, (cudaStream_t)stream);
  return glb::conv_wgrad_simt(x, gy, gw, N, H, W, Ci, Co, R, S, pad, alpha, (cudaStream_t)stream);
}

// ---- bf16 operand path (GLB_IMPL_BF16): tcgen05 kind::f16, bf16 x bf16 -> fp32 ------------------------------------------
namespace glb {
bool conv_fprop_bf16_covers(int N, int H, int W, int Ci, int Co, int R, int S, int pad);
bool conv_dgrad_bf16_covers(int N, int H, int W, int Ci, int Co, int R, int S, int pad);
bool conv_wgrad_bf16_covers(int N, int H, int W, int Ci, int Co, int R, int S, int pad);
int conv_fprop_bf16(const void* x, const void* w, const float* bias, float* y, int N, int H, int W, int Ci, int Co, int R, int S,
                    int pad, float alpha, float bias_scale, int act, float slope, cudaStream_t st);
int conv_dgrad_bf16(const void* gy, const void* wt, float* gx, int N, int H, int W, int Ci, int Co, int R, int S, int pad,
                    float alpha, cudaStream_t st);
int conv_wgrad_bf16(const void* x, const void* gy, float* gw, int N, int H, int W, int Ci, int Co, int R, int S, int pad,
                    float alpha, cudaStream_t st);
}  // namespace glb

extern "C" int glb_conv2d_bf16_covers(int kind, int N, int H, int W, int Ci, int Co, int R, int S, int pad) {
  if (N <= 0 || H <= 0 || W <= 0 || Ci <= 0 || Co <= 0 || R <= 0 || S <= 0 || pad < 0) return 0;
  switch (kind) {
    case 0: return glb::conv_fprop_bf16_covers(N, H, W, Ci, Co, R, S, pad) ? 1 : 0;
    case 1: return glb::conv_dgrad_bf16_covers(N, H, W, Ci, Co, R, S, pad) ? 1 : 0;
    case 2: return glb::conv_wgrad_bf16_covers(N, H, W, Ci, Co, R, S, pad) ? 1 : 0;
  }
  return 0;
}

extern "C" int glb_conv2d_fprop_bf16(const void* x, const void* w, const float* bias, float* y, int N, int H, int W, int Ci, int Co,
                                     int R, int S, int pad, float alpha, float bias_scale, int act, float slope,
                                     glb_stream_t stream) {
  if (N <= 0 || H <= 0 || W <= 0 || Ci <= 0 || Co <= 0 || R <= 0 || S <= 0 || pad < 0) return glb::shape_fail("conv2d_fprop_bf16");
  return glb::conv_fprop_bf16(x, w, bias, y, N, H, W, Ci, Co, R, S, pad, alpha, bias_scale, act, slope, (cudaStream_t)stream);
}

extern "C" int glb_conv2d_dgrad_bf16(const void* gy, const void* wt, float* gx, int N, int H, int W, int Ci, int Co, int R, int S,
                                     int pad, float alpha, glb_stream_t stream) {
  if (N <= 0 || H <= 0 || W <= 0 || Ci <= 0 || Co <= 0 || R <= 0 || S <= 0 || pad < 0) return glb::shape_fail("conv2d_dgrad_bf16");
  return glb::conv_dgrad_bf16(gy, wt, gx, N, H, W, Ci, Co, R, S, pad, alpha, (cudaStream_t)stream);
}

extern "C" int glb_conv2d_wgrad_bf16(const void* x, const void* gy, float* gw, int N, int H, int W, int Ci, int Co, int R, int S,
                                     int pad, float alpha, glb_stream_t stream) {
  if (N <= 0 || H <= 0 || W <= 0 || Ci <= 0 || Co <= 0 || R <= 0 || S <= 0 || pad < 0) return glb::shape_fail("conv2d_wgrad_bf16");
  return glb::conv_wgrad_bf16(x, gy, gw, N, H, W, Ci, Co, R, S, pad, alpha, (cudaStream_t)stream);
}

// ---- nearest-neighbour 2x upsample folded into the following 3x3 convolution (csrc/conv_tc.cu) -----------------------------
namespace glb {
bool conv_upconv_covers(int kind, int N, int H, int W, int Ci, int Co);
int conv_upconv_weights(const float* w, float* wp, float* wt, int Co, int Ci, cudaStream_t st);
int conv_upconv_fprop_tc(const float* x, const float* wp, const float* bias, float* y, int N, int H, int W, int Ci, int Co, float alpha,
                         float bias_scale, int act, float slope, cudaStream_t st);
int conv_upconv_dgrad_tc(const float* gy, const float* wt, float* gx, int N, int H, int W, int Ci, int Co, float alpha, cudaStream_t st);
int conv_upconv_wgrad_tc(const float* x, const float* gy, float* gwp, float* gw, int N, int H, int W, int Ci, int Co, float alpha,
                         cudaStream_t st);
}  // namespace glb

extern "C" int glb_upconv_covers(int kind, int N, int H, int W, int Ci, int Co) {
  if (N <= 0 || H <= 0 || W <= 0 || Ci <= 0 || Co <= 0) return 0;
  return glb::conv_upconv_covers(kind, N, H, W, Ci, Co) ? 1 : 0;
}
extern "C" int glb_upconv_weights(const float* w, float* wp, float* wt, int Co, int Ci, glb_stream_t stream) {
  if (Co <= 0 || Ci <= 0) return glb::shape_fail("upconv_weights");
  return glb::conv_upconv_weights(w, wp, wt, Co, Ci, (cudaStream_t)stream);
}
extern "C" int glb_upconv_fprop(const float* x, const float* wp, const float* bias, float* y, int N, int H, int W, int Ci, int Co,
                                float alpha, float bias_scale, int act, float slope, glb_stream_t stream) {
  if (N <= 0 || H <= 0 || W <= 0 || Ci <= 0 || Co <= 0) return glb::shape_fail("upconv_fprop");
  return glb::conv_upconv_fprop_tc(x, wp, bias, y, N, H, W, Ci, Co, alpha, bias_scale, act, slope, (cudaStream_t)stream);
}
extern "C" int glb_upconv_dgrad(const float* gy, const float* wt, float* gx, int N, int H, int W, int Ci, int Co, float alpha,
                                glb_stream_t stream) {
  if (N <= 0 || H <= 0 || W <= 0 || Ci <= 0 || Co <= 0) return glb::shape_fail("upconv_dgrad");
  return glb::conv_upconv_dgrad_tc(gy, wt, gx, N, H, W, Ci, Co, alpha, (cudaStream_t)stream);
}
extern "C" int glb_upconv_wgrad(const float* x, const float* gy, float* gwp, float* gw, int N, int H, int W, int Ci, int Co,
                                float alpha, glb_stream_t stream) {
  if (N <= 0 || H <= 0 || W <= 0 || Ci <= 0 || Co <= 0) return glb::shape_fail("upconv_wgrad");
  return glb::conv_upconv_wgrad_tc(x, gy, gwp, gw, N, H, W, Ci, Co, alpha, (cudaStream_t)stream);
}

// ---- 2x2 average pool folded into the 3x3 convolution in front of it (csrc/conv_tc.cu) ---------------------------------------
namespace glb {
bool conv_downconv_covers(int kind, int N, int H, int W, int Ci, int Co);
int conv_downconv_weights(const float* w, float* wp, float* wt, int Co, int Ci, cudaStream_t st);
int conv_downconv_fprop_tc(const float* x, const float* wt, const float* bias, float* y, int N, int H, int W, int Ci, int Co, float alpha,
                           float bias_scale, int act, float slope, cudaStream_t st);
int conv_downconv_dgrad_tc(const float* gy, const float* wp, float* gx, int N, int H, int W, int Ci, int Co, float alpha, cudaStream_t st);
int conv_downconv_wgrad_tc(const float* x, const float* gy, float* gwp, float* gw, int N, int H, int W, int Ci, int Co, float alpha,
                           cudaStream_t st);
}  // namespace glb

extern "C" int glb_downconv_covers(int kind, int N, int H, int W, int Ci, int Co) {
  if (N <= 0 || H <= 0 || W <= 0 || Ci <= 0 || Co <= 0) return 0;
  return glb::conv_downconv_covers(kind, N, H, W, Ci, Co) ? 1 : 0;
}
extern "C" int glb_downconv_weights(const float* w, float* wp, float* wt, int Co, int Ci, glb_stream_t stream) {
  if (Co <= 0 || Ci <= 0) return glb::shape_fail("downconv_weights");
  return glb::conv_downconv_weights(w, wp, wt, Co, Ci, (cudaStream_t)stream);
}
extern "C" int glb_downconv_fprop(const float* x, const float* wt, const float* bias, float* y, int N, int H, int W, int Ci, int Co,
                                  float alpha, float bias_scale, int act, float slope, glb_stream_t stream) {
  if (N <= 0 || H <= 0 || W <= 0 || Ci <= 0 || Co <= 0) return glb::shape_fail("downconv_fprop");
  return glb::conv_downconv_fprop_tc(x, wt, bias, y, N, H, W, Ci, Co, alpha, bias_scale, act, slope, (cudaStream_t)stream);
}
extern "C" int glb_downconv_dgrad(const float* gy, const float* wp, float* gx, int N, int H, int W, int Ci, int Co, float alpha,
                                  glb_stream_t stream) {
  if (N <= 0 || H <= 0 || W <= 0 || Ci <= 0 || Co <= 0) return glb::shape_fail("downconv_dgrad");
  return glb::conv_downconv_dgrad_tc(gy, wp, gx, N, H, W, Ci, Co, alpha, (cudaStream_t)stream);
}
extern "C" int glb_downconv_wgrad(const float* x, const float* gy, float* gwp, float* gw, int N, int H, int W, int Ci, int Co,
                                  float alpha, glb_stream_t stream) {
  if (N <= 0 || H <= 0 || W <= 0 || Ci <= 0 || Co <= 0) return glb::shape_fail("downconv_wgrad");
  return glb::conv_downconv_wgrad_tc(x, gy, gwp, gw, N, H, W, Ci, Co, alpha, (cudaStream_t)stream);
}

// ---- bf16-operand variants of the folded convolutions (opt-in: set_conv_impl("bf16")) ------------------------------------------
namespace glb {
bool conv_upconv_covers(int kind, int N, int H, int W, int Ci, int Co, int chunk);
bool conv_downconv_covers(int kind, int N, int H, int W, int Ci, int Co, int chunk);
int conv_upconv_fprop_bf16(const void* x, const void* wp, const float* bias, float* y, int N, int H, int W, int Ci, int Co, float alpha,
                           float bias_scale, int act, float slope, cudaStream_t st);
int conv_upconv_dgrad_bf16(const void* gy, const void* wt, float* gx, int N, int H, int W, int Ci, int Co, float alpha, cudaStream_t st);
int conv_upconv_wgrad_bf16(const void* x, const void* gy, float* gwp, float* gw, int N, int H, int W, int Ci, int Co, float alpha,
                           cudaStream_t st);
int conv_downconv_fprop_bf16(const void* x, const void* wt, const float* bias, float* y, int N, int H, int W, int Ci, int Co, float alpha,
                             float bias_scale, int act, float slope, cudaStream_t st);
int conv_downconv_dgrad_bf16(const void* gy, const void* wp, float* gx, int N, int H, int W, int Ci, int Co, float alpha, cudaStream_t st);
int conv_downconv_wgrad_bf16(const void* x, const void* gy, float* gwp, float* gw, int N, int H, int W, int Ci, int Co, float alpha,
                             cudaStream_t st);
}  // namespace glb

extern "C" int glb_upconv_bf16_covers(int kind, int N, int H, int W, int Ci, int Co) {
  if (N <= 0 || H <= 0 || W <= 0 || Ci <= 0 || Co <= 0) return 0;
  return glb::conv_upconv_covers(kind, N, H, W, Ci, Co, 64) ? 1 : 0;
}
extern "C" int glb_downconv_bf16_covers(int kind, int N, int H, int W, int Ci, int Co) {
  if (N <= 0 || H <= 0 || W <= 0 || Ci <= 0 || Co <= 0) return 0;
  return glb::conv_downconv_covers(kind, N, H, W, Ci, Co, 64) ? 1 : 0;
}
extern "C" int glb_upconv_fprop_bf16(const void* x_bf16, const void* wp_bf16, const float* bias, float* y, int N, int H, int W, int Ci,
                                     int Co, float alpha, float bias_scale, int act, float slope, glb_stream_t stream) {
  if (N <= 0 || H <= 0 || W <= 0 || Ci <= 0 || Co <= 0) return glb::shape_fail("upconv_fprop_bf16");
  return glb::conv_upconv_fprop_bf16(x_bf16, wp_bf16, bias, y, N, H, W, Ci, Co, alpha, bias_scale, act, slope, (cudaStream_t)stream);
}
extern "C" int glb_upconv_dgrad_bf16(const void* gy_bf16, const void* wt_bf16, float* gx, int N, int H, int W, int Ci, int Co, float alpha,
                                     glb_stream_t stream) {
  if (N <= 0 || H <= 0 || W <= 0 || Ci <= 0 || Co <= 0) return glb::shape_fail("upconv_dgrad_bf16");
  return glb::conv_upconv_dgrad_bf16(gy_bf16, wt_bf16, gx, N, H, W, Ci, Co, alpha, (cudaStream_t)stream);
}
extern "C" int glb_upconv_wgrad_bf16(const void* x_bf16, const void* gy_bf16, float* gwp, float* gw, int N, int H, int W, int Ci, int Co,
                                     float alpha, glb_stream_t stream) {
  if (N <= 0 || H <= 0 || W <= 0 || Ci <= 0 || Co <= 0) return glb::shape_fail("upconv_wgrad_bf16");
  return glb::conv_upconv_wgrad_bf16(x_bf16, gy_bf16, gwp, gw, N, H, W, Ci, Co, alpha, (cudaStream_t)stream);
}
extern "C" int glb_downconv_fprop_bf16(const void* x_bf16, const void* wt_bf16, const float* bias, float* y, int N, int H, int W, int Ci,
                                       int Co, float alpha, float bias_scale, int act, float slope, glb_stream_t stream) {
  if (N <= 0 || H <= 0 || W <= 0 || Ci <= 0 || Co <= 0) return glb::shape_fail("downconv_fprop_bf16");
  return glb::conv_downconv_fprop_bf16(x_bf16, wt_bf16, bias, y, N, H, W, Ci, Co, alpha, bias_scale, act, slope, (cudaStream_t)stream);
}
extern "C" int glb_downconv_dgrad_bf16(const void* gy_bf16, const void* wp_bf16, float* gx, int N, int H, int W, int Ci, int Co,
                                       float alpha, glb_stream_t stream) {
  if (N <= 0 || H <= 0 || W <= 0 || Ci <= 0 || Co <= 0) return glb::shape_fail("downconv_dgrad_bf16");
  return glb::conv_downconv_dgrad_bf16(gy_bf16, wp_bf16, gx, N, H, W, Ci, Co, alpha, (cudaStream_t)stream);
}
extern "C" int glb_downconv_wgrad_bf16(const void* x_bf16, const void* gy_bf16, float* gwp, float* gw, int N, int H, int W, int Ci,
                                       int Co, float alpha, glb_stream_t stream) {
  if (N <= 0 || H <= 0 || W <= 0 || Ci <= 0 || Co <= 0) return glb::shape_fail("downconv_wgrad_bf16");
  return glb::conv_downconv_wgrad_bf16(x_bf16, gy_bf16, gwp, gw, N, H, W, Ci, Co, alpha, (cudaStream_t)stream);
}
