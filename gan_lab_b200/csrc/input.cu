// Real-image input pipeline on the device.
//
// Reference: every real sample goes through torchvision on the host -- Resize(curr_res, interpolation=PIL BOX) ->
// ToTensor() -> Normalize(mean, std) (data_config.py:312-342; the Resize is rewritten at every resolution increase,
// progan/learner.py:1099-1112) -- i.e. Pillow's two-pass 8-bit resampler (Pillow 12.2 src/libImaging/Resample.c:
// precompute_coeffs / normalize_coeffs_8bpc / ImagingResampleHorizontal_8bpc / ImagingResampleVertical_8bpc) followed by
// u8 -> float, /255, -mean, /std in fp32.  Here the decoded uint8 images stay resident in HBM (or arrive as uint8: 4x less
// H2D than fp32) and ONE kernel produces the normalised fp32 NCHW batch at the current resolution, bit-exact with Pillow:
//   * the coefficient tables are Pillow's: double arithmetic on the host (glb_box_resize_tables), 22-bit fixed point;
//   * horizontal pass first, rounded and clipped to uint8 per input row, then the vertical pass over those uint8 values
//     (the intermediate image never leaves the SM: it lives in registers, one value per thread, channel and input row);
//   * the float tail uses IEEE division / subtraction in torchvision's order.
// HBM-bound byte work: algorithmic traffic = the source rows read once (3 B per input pixel) + 12 B per output pixel.
#include <math.h>

#include <vector>

#include "common.cuh"

namespace glb {
namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;  // Resample.c PRECISION_BITS
constexpr int kTileX = 128;                 // output columns per work item = threads per block
constexpr int kSmemBytes = 40 * 1024;       // staged source rows per chunk

__device__ __forceinline__ int clip8(int v) {
  v >>= kPrecisionBits;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

struct Params {
  const uint8_t* src;    // [M][Hs][Ws][3]
  const int64_t* index;  // [N] sample indices into src, or nullptr (sample n = image n)
  const uint8_t* flip;   // [N] 1 = mirror the sample horizontally (RandomHorizontalFlip after the resize), or nullptr
  float* dst;            // [N][3][Ho][Wo]
  const int* xb;         // [Wo][2] first source column, count
  const int* xk;         // [Wo][xks] fixed-point weights
  const int* yb;         // [Ho][2]
  const int* yk;         // [Ho][yks]
  int N, Hs, Ws, Ho, Wo, xks, yks;
  int64_t src_bytes;     // bytes addressable behind src (vector loads never cross it)
  float mean[3], stdv[3];
  int64_t items;         // N * Ho * ceil(Wo / kTileX)
  int xtiles;
};

__global__ void __launch_bounds__(kTileX) box_resize_normalize_kernel(const Params p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int t = threadIdx.x;
  for (int64_t item = blockIdx.x; item < p.items; item += gridDim.x) {
    const int xt = (int)(item % p.xtiles);
    const int yy = (int)((item / p.xtiles) % p.Ho);
    const int n = (int)(item / ((int64_t)p.xtiles * p.Ho));
    const int xx0 = xt * kTileX;
    const int xx1 = min(xx0 + kTileX, p.Wo);
    const int xx = xx0 + t;
    const bool live = xx < xx1;
    // source span of this tile (bounds are monotone in xx) and of this output row
    const int col_lo = p.xb[2 * xx0];
    const int col_hi = p.xb[2 * (xx1 - 1)] + p.xb[2 * (xx1 - 1) + 1];
    const int ymin = p.yb[2 * yy], ycnt = p.yb[2 * yy + 1];
    const int64_t img = p.index ? p.index[n] : n;
    const int64_t img_off = img * (int64_t)p.Hs * p.Ws * 3;
    const int span = (col_hi - col_lo) * 3;           // bytes of one source row this tile needs
    const int pitch = (span + 15 + 15) & ~15;         // smem bytes per staged row (room for the alignment skew)
    const int rows_per_chunk = max(1, min(ycnt, kSmemBytes / pitch));
    int xmin = 0, xcnt = 0;
    if (live) { xmin = p.xb[2 * xx]; xcnt = p.xb[2 * xx + 1]; }
    const int* kx = p.xk + (int64_t)xx * p.xks;
    const int* ky = p.yk + (int64_t)yy * p.yks;
    int acc0 = 1 << (kPrecisionBits - 1), acc1 = acc0, acc2 = acc0;   // vertical accumulators (Resample.c ss0..ss2)

    for (int r0 = 0; r0 < ycnt; r0 += rows_per_chunk) {
      const int rows = min(rows_per_chunk, ycnt - r0);
      __syncthreads();                                // previous chunk fully consumed
      // ---- stage `rows` source-row spans: 16-byte cp.async where the vector lies inside the buffer, bytes otherwise
      for (int r = 0; r < rows; ++r) {
        const int64_t g0 = img_off + ((int64_t)(ymin + r0 + r) * p.Ws + col_lo) * 3;   // first byte needed
        const int skew = (int)(((uintptr_t)(p.src + g0)) & 15);
        const int64_t a0 = g0 - skew;                                              // 16-byte aligned start
        uint8_t* srow = smem + r * pitch;                                          // srow[skew + b] = byte b of the span
        const int nvec = (skew + span + 15) >> 4;
        for (int v = t; v < nvec; v += kTileX) {
          const int64_t a = a0 + 16 * (int64_t)v;
          if (a >= 0 && a + 16 <= p.src_bytes) {
            cp_async16(srow + 16 * v, p.src + a);
          } else {
            for (int q = 0; q < 16; ++q) {
              const int64_t b = a + q;
              srow[16 * v + q] = (b >= 0 && b < p.src_bytes) ? p.src[b] : (uint8_t)0;
            }
          }
        }
      }
      cp_async_wait_all();
      __syncthreads();
      // ---- horizontal pass per staged row (rounded to uint8 like Pillow's intermediate image), vertical accumulate
      if (live) {
        for (int r = 0; r < rows; ++r) {
          const int64_t g0 = img_off + ((int64_t)(ymin + r0 + r) * p.Ws + col_lo) * 3;
          const int skew = (int)(((uintptr_t)(p.src + g0)) & 15);
          const uint8_t* px = smem + r * pitch + skew + (xmin - col_lo) * 3;
          int h0 = 1 << (kPrecisionBits - 1), h1 = h0, h2 = h0;
          for (int x = 0; x < xcnt; ++x) {
            const int k = __ldg(kx + x);
            h0 += (int)px[3 * x + 0] * k;
            h1 += (int)px[3 * x + 1] * k;
            h2 += (int)px[3 * x + 2] * k;
          }
          const int kv = __ldg(ky + r0 + r);
          acc0 += clip8(h0) * kv;
          acc1 += clip8(h1) * kv;
          acc2 += clip8(h2) * kv;
        }
      }
    }
    if (live) {
      const int xo = (p.flip && p.flip[n]) ? (p.Wo - 1 - xx) : xx;
      const int64_t plane = (int64_t)p.Ho * p.Wo;
      float* o = p.dst + ((int64_t)n * 3 * p.Ho + yy) * p.Wo + xo;
      // ToTensor: u8 -> float, / 255 ; Normalize: (x - mean) / std   (torchvision order, IEEE ops)
      o[0] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)clip8(acc0), 255.f), p.mean[0]), p.stdv[0]);
      o[plane] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)clip8(acc1), 255.f), p.mean[1]), p.stdv[1]);
      o[2 * plane] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)clip8(acc2), 255.f), p.mean[2]), p.stdv[2]);
    }
  }
}

}  // namespace

// Resample.c precompute_coeffs + normalize_coeffs_8bpc for the BOX filter (support 0.5), in double like Pillow.
static int box_ksize(int in_size, int out_size) {
  double filterscale = (double)in_size / (double)out_size;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 0.5 * filterscale;
  return (int)ceil(support) * 2 + 1;
}

}  // namespace glb

extern "C" int glb_box_resize_ksize(int in_size, int out_size) {
  if (in_size <= 0 || out_size <= 0) return -1;
  return glb::box_ksize(in_size, out_size);
}

extern "C" int glb_box_resize_tables(int in_size, int out_size, int32_t* bounds, int32_t* kk) {
  if (in_size <= 0 || out_size <= 0 || !bounds || !kk) return glb::shape_fail("box_resize_tables");
  const double scale = (double)in_size / (double)out_size;
  double filterscale = scale;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 0.5 * filterscale;
  const int ksize = (int)ceil(support) * 2 + 1;
  const double ss = 1.0 / filterscale;
  std::vector<double> w((size_t)ksize);
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    double ww = 0.0;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    int x = 0;
    for (; x < xmax; ++x) {
      const double a = (x + xmin - center + 0.5) * ss;
      w[x] = (a > -0.5 && a <= 0.5) ? 1.0 : 0.0;   // box_filter
      ww += w[x];
    }
    for (x = 0; x < xmax; ++x)
      if (ww != 0.0) w[x] /= ww;
    for (; x < ksize; ++x) w[x] = 0.0;
    for (x = 0; x < ksize; ++x)
      kk[(int64_t)xx * ksize + x] = (w[x] < 0) ? (int)(-0.5 + w[x] * (1 << glb::kPrecisionBits))
                                               : (int)(0.5 + w[x] * (1 << glb::kPrecisionBits));
    bounds[2 * xx + 0] = xmin;
    bounds[2 * xx + 1] = xmax;
  }
  return GLB_OK;
}

extern "C" int glb_u8_box_resize_normalize(const uint8_t* src, int64_t src_images, const int64_t* index, const uint8_t* flip,
                                           float* dst, int N, int Hs, int Ws, int Ho, int Wo, const int32_t* xb,
                                           const int32_t* xk, int xks, const int32_t* yb, const int32_t* yk, int yks,
                                           const float* mean3, const float* std3, glb_stream_t stream) {
  if (N <= 0 || Hs <= 0 || Ws <= 0 || Ho <= 0 || Wo <= 0 || xks <= 0 || yks <= 0 || src_images <= 0 || !mean3 || !std3)
    return glb::shape_fail("u8_box_resize_normalize");
  if (xks != glb::box_ksize(Ws, Wo) || yks != glb::box_ksize(Hs, Ho))
    return glb::shape_fail("u8_box_resize_normalize: coefficient tables do not belong to these sizes");
  glb::Params p;
  p.src = src; p.index = index; p.flip = flip; p.dst = dst;
  p.xb = xb; p.xk = xk; p.yb = yb; p.yk = yk;
  p.N = N; p.Hs = Hs; p.Ws = Ws; p.Ho = Ho; p.Wo = Wo; p.xks = xks; p.yks = yks;
  p.src_bytes = src_images * (int64_t)Hs * Ws * 3;
  for (int c = 0; c < 3; ++c) { p.mean[c] = mean3[c]; p.stdv[c] = std3[c]; }
  p.xtiles = (Wo + glb::kTileX - 1) / glb::kTileX;
  p.items = (int64_t)N * Ho * p.xtiles;
  // one staged row must fit: a tile spans at most Ws source columns
  if ((int64_t)Ws * 3 + 32 > glb::kSmemBytes) return glb::shape_fail("u8_box_resize_normalize: source rows wider than 13 K pixels");
  const int grid = (int)(p.items < (int64_t)glb::kNumSMs * 8 ? p.items : (int64_t)glb::kNumSMs * 8);
  glb::box_resize_normalize_kernel<<<grid, glb::kTileX, glb::kSmemBytes, (cudaStream_t)stream>>>(p);
  GLB_CHECK_LAUNCH("box_resize_normalize_kernel");
  return GLB_OK;
}
