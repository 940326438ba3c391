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

#include <algorithm>
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "tc_common.cuh"

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

// ------------------------------------------------------------------------------------------------ grouped / pipelined variant
// Same arithmetic, restructured for bandwidth (the kernel above is bound by its per-item latency chain -- table loads, staging,
// two block barriers for ONE output row -- and by single-byte shared-memory loads with 6-way bank conflicts: 0.08-0.24 of HBM):
//   * a work item is a GROUP of G output rows x 128 output columns; its source rows are staged once (G chosen by the host so
//     that the rows of the largest group fit one buffer);
//   * two buffers: the cp.async staging of the next item runs under the arithmetic of the current one;
//   * 512 threads = 128 columns x 4 row lanes.  Pass 1 (Pillow's horizontal pass): every (staged source row, output column)
//     pair is resampled once, rounded to uint8 and kept in shared memory -- the staged bytes are read as 32-bit words
//     (funnel-shifted to the window's byte alignment), four pixels (12 bytes = 3 words) per step; coefficients beyond the
//     window are zero, so whole steps need no tail handling.  Pass 2 (vertical pass) runs over those uint8 triples.
constexpr int kBufBytes = 40 * 1024;        // per staging buffer (+ slack behind it for the word reads of the last window)
constexpr int kBufSlack = 64;
constexpr int kMaxRows = 40;                // staged source rows per item (intermediate image: kMaxRows x 128 x 4 B = 20 KB)
constexpr int kThreads2 = 512;
constexpr int kRowLanes = kThreads2 / kTileX;

struct Params2 {
  Params b;
  int G;            // output rows per item
  int ygroups;      // ceil(Ho / G)
  int ytab_in_smem; // the row bounds / weights ([Ho][2 + yks] ints) live in shared memory
};

__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_only(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

// stage the source rows [row_lo, row_hi) x columns [col_lo, col_hi) of image `img`: buf[r * pitch + skew[r] + b] = byte b of row r;
// completion is signalled on `bar` (one phase per call).  Interior items: warp 0 issues ONE bulk copy (TMA, 1-D) per row from
// the 16-byte aligned start below the row's first byte; items that touch the first / last bytes of the source buffer (where an
// aligned vector could leave it) take the element-wise path.  Block-uniform: every thread calls it.
__device__ __forceinline__ void stage_item(const Params& p, int64_t img, int col_lo, int col_hi, int row_lo, int row_hi, uint8_t* buf,
                                           int* skews, int pitch, uint32_t bar) {
  const int span = (col_hi - col_lo) * 3;
  const int64_t base = (img * (int64_t)p.Hs + row_lo) * p.Ws * 3 + (int64_t)col_lo * 3;     // first byte of the first staged row
  const int row_bytes = p.Ws * 3;
  const int nrows = row_hi - row_lo;
  const bool interior = base >= 16 && base + (int64_t)(nrows - 1) * row_bytes + span + 32 <= p.src_bytes;
  if (interior) {
    if (threadIdx.x < 32) {
      for (int r = threadIdx.x; r < nrows; r += 32) {
        const uint8_t* g = p.src + base + (int64_t)r * row_bytes;
        const int skew = (int)(((uintptr_t)g) & 15);
        const uint32_t bytes = (uint32_t)((skew + span + 15) & ~15);
        skews[r] = skew;
        mbar_expect_tx_only(bar, bytes);
        bulk_copy_g2s(tc::smem_u32(buf + r * pitch), g - skew, bytes, bar);
      }
      __syncwarp();
      if (threadIdx.x == 0) tc::mbar_arrive(bar);          // after every expect_tx of this phase
    }
    return;
  }
  const int tx = threadIdx.x & (kTileX - 1), ty = threadIdx.x / kTileX;
  for (int r = ty; r < nrows; r += kRowLanes) {
    const int64_t g0 = base + (int64_t)r * row_bytes;
    const int skew = (int)(((uintptr_t)(p.src + g0)) & 15);
    const int64_t a0 = g0 - skew;
    uint8_t* srow = buf + r * pitch;
    if (tx == 0) skews[r] = skew;
    const int nvec = (skew + span + 15) >> 4;
    for (int v = tx; v < nvec; v += kTileX) {
      const int64_t a = a0 + 16 * (int64_t)v;
      for (int qb = 0; qb < 16; ++qb) {
        const int64_t b = a + qb;
        srow[16 * v + qb] = (b >= 0 && b < p.src_bytes) ? p.src[b] : (uint8_t)0;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) tc::mbar_arrive(bar);
}


// one step of the horizontal pass: 4 pixels = 12 bytes (b0, b1, b2 = the window's next three words) times k0..k3
__device__ __forceinline__ void hstep(uint32_t b0, uint32_t b1, uint32_t b2, int k0, int k1, int k2, int k3, int& h0, int& h1, int& h2) {
  h0 += (int)(b0 & 0xffu) * k0; h1 += (int)((b0 >> 8) & 0xffu) * k0; h2 += (int)((b0 >> 16) & 0xffu) * k0;
  h0 += (int)(b0 >> 24) * k1; h1 += (int)(b1 & 0xffu) * k1; h2 += (int)((b1 >> 8) & 0xffu) * k1;
  h0 += (int)((b1 >> 16) & 0xffu) * k2; h1 += (int)(b1 >> 24) * k2; h2 += (int)(b2 & 0xffu) * k2;
  h0 += (int)((b2 >> 8) & 0xffu) * k3; h1 += (int)((b2 >> 16) & 0xffu) * k3; h2 += (int)(b2 >> 24) * k3;
}

template <bool YS>     // YS: the row bounds / weights live in shared memory (q.ytab_in_smem)
__global__ void __launch_bounds__(kThreads2, 2) box_resize_normalize_grouped_kernel(const Params2 q) {
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ float lut[3][256];        // ToTensor + Normalize of every uint8 value, in torchvision's order with IEEE ops
  __shared__ uint64_t bars[2];         // "staging buffer b has landed"
  const Params& p = q.b;
  uint8_t* const buf0 = smem;
  uint8_t* const buf1 = smem + kBufBytes + kBufSlack;
  uint32_t* const hbuf = reinterpret_cast<uint32_t*>(smem + 2 * (kBufBytes + kBufSlack));            // [kMaxRows][128] packed u8 triples
  int* const skew_tab = reinterpret_cast<int*>(smem + 2 * (kBufBytes + kBufSlack) + kMaxRows * kTileX * 4);   // [2][kMaxRows]
  int* const ytab = skew_tab + 2 * kMaxRows;                                                          // [Ho][2 + yks] when q.ytab_in_smem
  const int tx = threadIdx.x & (kTileX - 1), ty = threadIdx.x / kTileX;
  for (int i = threadIdx.x; i < 3 * 256; i += kThreads2) {
    const int c = i >> 8, v = i & 255;
    lut[c][v] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)v, 255.f), p.mean[c]), p.stdv[c]);
  }
  if (threadIdx.x == 0) {
    tc::mbar_init(tc::smem_u32(&bars[0]), 1);
    tc::mbar_init(tc::smem_u32(&bars[1]), 1);
    tc::fence_barrier_init();
  }
  const int ystride = 2 + p.yks;
  if (YS) {
    for (int i = threadIdx.x; i < p.Ho * ystride; i += kThreads2) {
      const int yy = i / ystride, j = i - yy * ystride;
      ytab[i] = j < 2 ? __ldg(p.yb + 2 * yy + j) : __ldg(p.yk + (int64_t)yy * p.yks + (j - 2));
    }
  }
  __syncthreads();
  auto ylo = [&](int yy) { return YS ? ytab[yy * ystride] : __ldg(p.yb + 2 * yy); };
  auto ycn = [&](int yy) { return YS ? ytab[yy * ystride + 1] : __ldg(p.yb + 2 * yy + 1); };
  auto ykk = [&](int yy, int r) { return YS ? ytab[yy * ystride + 2 + r] : __ldg(p.yk + (int64_t)yy * p.yks + r); };
  // the grid is a multiple of xtiles: a block keeps its column tile, so everything about the columns is loop invariant
  const int xt = (int)(blockIdx.x % p.xtiles);
  const int xx0 = xt * kTileX, xx1 = min(xx0 + kTileX, p.Wo);
  const int col_lo = __ldg(p.xb + 2 * xx0);
  const int col_hi = __ldg(p.xb + 2 * (xx1 - 1)) + __ldg(p.xb + 2 * (xx1 - 1) + 1);
  const int pitch = ((col_hi - col_lo) * 3 + 15 + 31) & ~15;
  const int xx = xx0 + tx;
  const bool live = xx < xx1;
  int xcnt = 0, col_off = 0;
  int kreg[12];
  const int* kx = p.xk + (int64_t)(live ? xx : xx0) * p.xks;
#pragma unroll
  for (int i = 0; i < 12; ++i) kreg[i] = (live && i < p.xks) ? __ldg(kx + i) : 0;
  if (live) { xcnt = __ldg(p.xb + 2 * xx + 1); col_off = (__ldg(p.xb + 2 * xx) - col_lo) * 3; }
  const bool in_regs = p.xks <= 12;
  // BOX windows have ONE weight for all their taps (integer ratios; otherwise a zero may sit at a window edge): then
  // sum(byte * k) = k * sum(byte), and the byte sums are three dp4a per word instead of twelve unpack + multiply-add pairs
  bool uniform = in_regs;
#pragma unroll
  for (int i = 1; i < 12; ++i) uniform = uniform && (i >= xcnt || kreg[i] == kreg[0]);
  // item -> (sample n, row group yg); items advance by gridDim.x = du whole (n, yg) steps, tracked without divisions
  auto rows_of = [&](int yg, int& yy0, int& yy1, int& row_lo, int& row_hi) {
    yy0 = yg * q.G; yy1 = min(yy0 + q.G, p.Ho);
    row_lo = ylo(yy0); row_hi = ylo(yy1 - 1) + ycn(yy1 - 1);
  };
  int item = blockIdx.x;
  const int items = (int)p.items;                 // < 2^30 (checked by the host)
  const int du = (int)gridDim.x / p.xtiles, dq = du / q.ygroups, dr = du - dq * q.ygroups;
  int n = (item / p.xtiles) / q.ygroups, yg = (item / p.xtiles) - n * q.ygroups;
  int yy0 = 0, yy1 = 0, row_lo = 0, row_hi = 0;
  int which = 0;
  uint32_t phases = 0u;                // bit b = parity to wait for on bars[b]
  if (item < items) {
    rows_of(yg, yy0, yy1, row_lo, row_hi);
    stage_item(p, p.index ? p.index[n] : n, col_lo, col_hi, row_lo, row_hi, buf0, skew_tab, pitch, tc::smem_u32(&bars[0]));
  }
  int n2 = n, yg2 = yg, yy0b = 0, yy1b = 0, row_lo2 = 0, row_hi2 = 0;
  for (; item < items; item += gridDim.x, which ^= 1) {
    {  // prefetch the next item of this block into the other buffer
      n2 = n + dq; yg2 = yg + dr;
      if (yg2 >= q.ygroups) { yg2 -= q.ygroups; ++n2; }
      if (item + (int)gridDim.x < items) {
        rows_of(yg2, yy0b, yy1b, row_lo2, row_hi2);
        stage_item(p, p.index ? p.index[n2] : n2, col_lo, col_hi, row_lo2, row_hi2, which ? buf0 : buf1,
                   skew_tab + (which ^ 1) * kMaxRows, pitch, tc::smem_u32(&bars[which ^ 1]));
      }
    }
    tc::mbar_wait(tc::smem_u32(&bars[which]), (phases >> which) & 1u);     // this item's rows have landed
    phases ^= 1u << which;
    __syncthreads();                                             // (the skew table was written with ordinary stores)
    const uint8_t* buf = which ? buf1 : buf0;
    const int* skews = skew_tab + which * kMaxRows;
    // ---- pass 1: horizontal, one (source row, output column) per step of this thread
    if (live) {
      for (int rr = ty; rr < row_hi - row_lo; rr += kRowLanes) {
        const uint32_t byte_addr = (uint32_t)(rr * pitch + skews[rr] + col_off);
        const uint32_t* wp = reinterpret_cast<const uint32_t*>(buf + (byte_addr & ~3u));
        const uint32_t sh = (byte_addr & 3u) * 8u;
        int h0 = 1 << (kPrecisionBits - 1), h1 = h0, h2 = h0;
        uint32_t w0 = wp[0];
        if (uniform && xcnt > 1) {
          uint32_t s0 = 0u, s1 = 0u, s2 = 0u;
#pragma unroll
          for (int x4 = 0; x4 < 12; x4 += 4) {
            if (x4 < xcnt) {
              const uint32_t w1 = wp[1], w2 = wp[2], w3 = wp[3];
              const uint32_t b0 = __funnelshift_r(w0, w1, sh), b1 = __funnelshift_r(w1, w2, sh), b2 = __funnelshift_r(w2, w3, sh);
              w0 = w3; wp += 3;
              const int nr = xcnt - x4;                   // pixels of this step inside the window (>= 1)
              const uint32_t e0 = nr >= 2 ? 0xffffffffu : 0x00ffffffu;                             // word 0: px0 bytes 0-2, px1 byte 3
              const uint32_t e1 = nr >= 3 ? 0xffffffffu : (nr == 2 ? 0x0000ffffu : 0u);           // word 1: px1 bytes 0-1, px2 bytes 2-3
              const uint32_t e2 = nr >= 4 ? 0xffffffffu : (nr == 3 ? 0x000000ffu : 0u);           // word 2: px2 byte 0, px3 bytes 1-3
              s0 = __dp4a(b0, 0x01000001u & e0, s0); s1 = __dp4a(b0, 0x00000100u & e0, s1); s2 = __dp4a(b0, 0x00010000u & e0, s2);
              s0 = __dp4a(b1, 0x00010000u & e1, s0); s1 = __dp4a(b1, 0x01000001u & e1, s1); s2 = __dp4a(b1, 0x00000100u & e1, s2);
              s0 = __dp4a(b2, 0x00000100u & e2, s0); s1 = __dp4a(b2, 0x00010000u & e2, s1); s2 = __dp4a(b2, 0x01000001u & e2, s2);
            }
          }
          h0 += (int)s0 * kreg[0]; h1 += (int)s1 * kreg[0]; h2 += (int)s2 * kreg[0];
        } else if (xcnt == 1) {                           // resolution kept (or enlarged): one tap
          const uint32_t b0 = __funnelshift_r(w0, wp[1], sh);
          h0 += (int)(b0 & 0xffu) * kreg[0]; h1 += (int)((b0 >> 8) & 0xffu) * kreg[0]; h2 += (int)((b0 >> 16) & 0xffu) * kreg[0];
        } else if (in_regs) {
#pragma unroll
          for (int x4 = 0; x4 < 12; x4 += 4) {
            if (x4 < xcnt) {
              const uint32_t w1 = wp[1], w2 = wp[2], w3 = wp[3];
              hstep(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh), kreg[x4], kreg[x4 + 1],
                    kreg[x4 + 2], kreg[x4 + 3], h0, h1, h2);
              w0 = w3; wp += 3;
            }
          }
        } else {
          for (int x4 = 0; x4 < xcnt; x4 += 4) {
            const uint32_t w1 = wp[1], w2 = wp[2], w3 = wp[3];
            const int k0 = __ldg(kx + x4);
            const int k1 = (x4 + 1 < p.xks) ? __ldg(kx + x4 + 1) : 0;
            const int k2 = (x4 + 2 < p.xks) ? __ldg(kx + x4 + 2) : 0;
            const int k3 = (x4 + 3 < p.xks) ? __ldg(kx + x4 + 3) : 0;
            hstep(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh), k0, k1, k2, k3, h0, h1, h2);
            w0 = w3; wp += 3;
          }
        }
        hbuf[rr * kTileX + tx] = (uint32_t)clip8(h0) | ((uint32_t)clip8(h1) << 8) | ((uint32_t)clip8(h2) << 16);
      }
    }
    __syncthreads();
    // ---- pass 2: vertical over the uint8 intermediate, one (output row, output column) per step
    if (live) {
      const int xo = (p.flip && p.flip[n]) ? (p.Wo - 1 - xx) : xx;
      const int plane = p.Ho * p.Wo;                                   // < 2^31 / 3 (checked by the host)
      float* const o_n = p.dst + (int64_t)n * 3 * plane + xo;
      for (int yy = yy0 + ty; yy < yy1; yy += kRowLanes) {
        const int ymin = ylo(yy), ycnt = ycn(yy);
        int acc0 = 1 << (kPrecisionBits - 1), acc1 = acc0, acc2 = acc0;
        for (int r = 0; r < ycnt; ++r) {
          const uint32_t v = hbuf[(ymin + r - row_lo) * kTileX + tx];
          const int kv = ykk(yy, r);
          acc0 += (int)(v & 0xffu) * kv;
          acc1 += (int)((v >> 8) & 0xffu) * kv;
          acc2 += (int)((v >> 16) & 0xffu) * kv;
        }
        float* o = o_n + yy * p.Wo;
        o[0] = lut[0][clip8(acc0)];
        o[plane] = lut[1][clip8(acc1)];
        o[2 * plane] = lut[2][clip8(acc2)];
      }
    }
    __syncthreads();                   // staging buffer and intermediate image are rewritten by the next iteration
    n = n2; yg = yg2; yy0 = yy0b; yy1 = yy1b; row_lo = row_lo2; row_hi = row_hi2;
  }
}

}  // namespace

// Resample.c precompute_coeffs + normalize_coeffs_8bpc for the BOX filter (support 0.5), in double like Pillow.
// [lo, hi) source index range of every output index: the bounds of glb_box_resize_tables without the weights
static void box_bounds(int in_size, int out_size, int* lo, int* hi) {
  const double scale = (double)in_size / (double)out_size;
  double filterscale = scale;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 0.5 * filterscale;
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    lo[xx] = xmin; hi[xx] = xmax;
  }
}

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
  // grouped / pipelined kernel: G output rows per item, the largest G <= 32 whose worst group fits one staging buffer
  if (getenv("GLB_INPUT_V1") == nullptr) {
    std::vector<int> ylo((size_t)Ho), yhi((size_t)Ho), xlo((size_t)Wo), xhi((size_t)Wo);
    glb::box_bounds(Hs, Ho, ylo.data(), yhi.data());
    glb::box_bounds(Ws, Wo, xlo.data(), xhi.data());
    int span_max = 0;
    for (int x0 = 0; x0 < Wo; x0 += glb::kTileX) {
      const int x1 = (x0 + glb::kTileX < Wo ? x0 + glb::kTileX : Wo) - 1;
      span_max = std::max(span_max, xhi[x1] - xlo[x0]);
    }
    const int pitch_max = (span_max * 3 + 15 + 31) & ~15;
    int G = 0;
    for (int g = 32; g >= 1; --g) {
      int rows_max = 0;
      for (int y0 = 0; y0 < Ho; y0 += g) {
        const int y1 = (y0 + g < Ho ? y0 + g : Ho) - 1;
        rows_max = std::max(rows_max, yhi[y1] - ylo[y0]);
      }
      if (rows_max <= glb::kMaxRows && (int64_t)rows_max * pitch_max <= glb::kBufBytes) { G = g; break; }
    }
    if (G > 0 && (int64_t)N * ((Ho + G - 1) / G) * p.xtiles < (int64_t)1 << 30 && (int64_t)Ho * Wo * 3 < (int64_t)1 << 31) {
      glb::Params2 q;
      q.G = G; q.ygroups = (Ho + G - 1) / G;
      p.items = (int64_t)N * q.ygroups * p.xtiles;
      q.b = p;
      const int64_t ytab_bytes = (int64_t)Ho * (2 + yks) * 4;
      q.ytab_in_smem = ytab_bytes <= 8 * 1024 ? 1 : 0;       // two blocks per SM must still fit
      const int smem_fixed = 2 * (glb::kBufBytes + glb::kBufSlack) + glb::kMaxRows * glb::kTileX * 4 + 2 * glb::kMaxRows * (int)sizeof(int);
      const int smem = smem_fixed + (q.ytab_in_smem ? (int)ytab_bytes : 0);
      static bool configured = false;
      if (!configured) {
        GLB_CUDA(cudaFuncSetAttribute(glb::box_resize_normalize_grouped_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      smem_fixed + 8 * 1024));
        GLB_CUDA(cudaFuncSetAttribute(glb::box_resize_normalize_grouped_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      smem_fixed));
        configured = true;
      }
      // a multiple of xtiles (every block keeps its column tile), at most two blocks per SM
      int64_t cap = (int64_t)glb::kNumSMs * 2 / p.xtiles * p.xtiles;
      if (cap < p.xtiles) cap = p.xtiles;
      const int grid = (int)(p.items < cap ? p.items : cap);
      if (q.ytab_in_smem) glb::box_resize_normalize_grouped_kernel<true><<<grid, glb::kThreads2, smem, (cudaStream_t)stream>>>(q);
      else glb::box_resize_normalize_grouped_kernel<false><<<grid, glb::kThreads2, smem, (cudaStream_t)stream>>>(q);
      GLB_CHECK_LAUNCH("box_resize_normalize_grouped_kernel");
      return GLB_OK;
    }
  }
  p.items = (int64_t)N * Ho * p.xtiles;
  // one staged row must fit: a tile spans at most Ws source columns
  if ((int64_t)Ws * 3 + 32 > glb::kSmemBytes) return glb::shape_fail("u8_box_resize_normalize: source rows wider than 13 K pixels");
  const int grid = (int)(p.items < (int64_t)glb::kNumSMs * 8 ? p.items : (int64_t)glb::kNumSMs * 8);
  glb::box_resize_normalize_kernel<<<grid, glb::kTileX, glb::kSmemBytes, (cudaStream_t)stream>>>(p);
  GLB_CHECK_LAUNCH("box_resize_normalize_kernel");
  return GLB_OK;
}
