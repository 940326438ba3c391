// tcgen05 / TMEM / TMA implicit-GEMM convolution family (TF32 operands, fp32 accumulate) for sm_100a.
//
// Replaces aten::convolution / convolution_backward as dispatched from Conv2dEx (reference
// utils/custom_layers.py:166-169, 202-211) on the tensor cores.  Layouts as in include/ganlab_b200.h:
// activations NHWC, weights KRSC.
//
// fprop (and dgrad, which is fprop over gy with the flipped/transposed weights of
// glb_conv2d_weight_transpose):
//   GEMM  D[m = output pixel][n = co] = sum_{tap=(r,s)} sum_{ci} X[pixel shifted by (r-pad, s-pad)][ci] * W[co][tap][ci]
//   * M tile = 128 output pixels = a (bn images x bh rows x bw cols) box; for every tap and 32-channel
//     chunk ONE 4-D TMA box load of the shifted input window lands the A tile in shared memory in the
//     K-major 128B-swizzled layout tcgen05 wants; TMA's out-of-bounds zero fill implements the padding.
//   * N tile = BN output channels; the B tile is a 3-D TMA box (32 ci, 1 tap, BN co) of the KRSC weights.
//   * persistent CTAs (one per SM), warp specialised: warp 0 = TMA producer, warp 1 = MMA issuer (one
//     thread), warp 2 = TMEM allocator, warps 4-7 = epilogue (TMEM -> registers -> alpha/bias/act -> HBM).
//     smem ring (full/empty mbarriers) between TMA and MMA; two TMEM accumulator stages (tmem_full/empty
//     mbarriers) so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Kernel map (all in this file):
//   conv_fprop_tc_kernel<BN,KC,BF>        single CTA, split-K; the low-resolution layers and every uncovered special case
//   conv_fprop_tc2_kernel<BN,BF>          CTA pair (cta_group::2, M = 256)
//   conv_fprop_tc_halo_kernel / conv_fprop_tc2_halo_kernel   3x3 tap reuse (three column-shifted boxes per chunk serve nine taps)
//   conv_fprop_tc2_fold_kernel<BN,BF,UP>  tap reuse for the 2x2 phase taps of the upsample- / pool-folded layers
//   conv_fprop_tc_m2_kernel, conv_fprop_tc2_h1_kernel          measured experiments (A/B switches, off by default)
//   conv_wgrad_tc_kernel<BN,PIX,BF>, conv_wgrad_tc3_kernel<BN>, conv_wgrad_fold2_kernel<BN> (A/B)    weight gradients (MN-major operands)
//   fold_weights_kernel<DOWN>, upconv_wgrad_fold_kernel, fold_wgrad_transposed_kernel    operand re-layouts of the folded layers
// The folded layers (glb_upconv_* / glb_downconv_*, FpropParams::up, WgradParams::up) are documented above upconv_launch().
#include <stdlib.h>

#include "tc_common.cuh"

namespace glb {
namespace {
using namespace tc;

struct FpropParams {
  float* y;
  const float* bias;
  int N, Ho, Wo, Co;
  int Ci, R, S, pad;
  int bw, bh, bn;  // M-tile box, bw*bh*bn == 128
  int tiles_w, tiles_h, tiles_n, tiles_co;
  int num_tiles;   // work items = output tiles x ksplit
  long long* trace;   // tuning experiments only: per-CTA clock64 timeline (8 slots) or NULL
  int dbg;            // tuning experiments only: 1 = skip the A loads of taps > 0, 2 = skip the B loads of all but the first k step
  int ksplit, k_per;  // split-K: slice j of an output tile owns k iterations [j*k_per, min(k_iters, (j+1)*k_per))
  float alpha, bias_scale;
  int act;
  float slope;
  // nearest-neighbour 2x upsample folded into a 3x3 "same" convolution (conv_fprop_tc_kernel / conv_fprop_tc2_kernel only):
  //   up = 1: forward -- output phase (dy, dx) of pixel (2i+dy, 2j+dx) is a 2x2 convolution of the LOW-resolution input with
  //           pre-summed taps (4/9 of the MMA work, no upsampled copy): the phase is part of the M index, pad = 1 - d per axis,
  //           weights [4*Co][2][2][Ci] phase-major, (Ho, Wo) = the low-resolution grid, (OH, OW) = 2x that
  //   up = 2: its data gradient -- K runs over (phase, tap, Co chunk): A comes from four strided views of the high-resolution
  //           gradient (one tensor map per phase), weights [Ci][16][Co]
  int up;
  int OH, OW;         // dims of the output tensor (address computation); = (Ho, Wo) unless up == 1
};

struct alignas(64) TMapSet { CUtensorMap m[4]; };   // A-operand maps: m[0] only, or one per output phase (up == 2)

// tcgen05 kind::tf32 TRUNCATES its fp32 operands to 10 mantissa bits (cuDNN's TF32 kernels round to nearest).  Truncation is
// biased: for a log-uniform mantissa the expected relative error of an operand is 0.5 * 2^-10 * E[1/m] = 3.52e-4, i.e. every
// product comes out 7.04e-4 too small on average -- a coherent scale error that accumulates over the ~26 convolutions a
// discriminator gradient passes through (measured at cfg2's size: R1 penalty value -2.3 % against the reference's CPU run, its
// own cuDNN-TF32 run -0.08 %).  The kernels fold the reciprocal of the expected shrinkage into the epilogue scale `alpha`
// (mean-compensated truncation): the bias goes, the zero-mean part (same variance as round-to-nearest) stays.
constexpr float kTf32TruncComp = 1.0f / (1.0f - 7.04e-4f);

constexpr int kBM = 128;           // MMA M (output pixels per tile)
constexpr int kChunk = 32;         // channels per K chunk = 128 bytes = one swizzle row
constexpr int kABytes = kBM * 128; // 16 KB

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// KC = 32-channel K chunks per pipeline stage.  The single MMA-issuing thread pays ~225 cycles of wait / fence / commit per
// stage (tools/micro/mma_rate.cu): with 4 MMAs of N <= 128 (<= 64 cycles each) per stage that protocol, not the tensor
// pipe, sets the pace (measured 576-642 cycles per stage at BN = 128), so narrow tiles use 2 chunks = 8 MMAs per stage.
template <int BN, int KC>
struct FpropCfg {
  static constexpr int kBBytes = BN * 128;
  static constexpr int kChunkBytes = kABytes + kBBytes;
  static constexpr int kStageBytes = KC * kChunkBytes;
  static constexpr int kStages = (192 * 1024 / kStageBytes) > 8 ? 8 : (192 * 1024 / kStageBytes);
  static constexpr int kTmemCols = (2 * BN < 64) ? 64 : 2 * BN;  // power of two >= 32; the epilogue reads 32-column groups
  static constexpr int kEpiWarps = (BN >= 64) ? 8 : 4;           // 8: two warps per TMEM lane quarter, half the columns each
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

constexpr int kH1Default = 0;     // conv_fprop_tc2_h1_kernel: off until verified on hardware (GLB_FPROP_H1=1 / 2)
constexpr int kHaloDefault = 12;  // tap-reuse kernels used by default: bits 0/1 = single CTA at BN 128/256, bits 2/3 = CTA pair at BN 128/256
constexpr int kFpropThreads = 384;  // warps 0-3: TMA producer, MMA issuer, TMEM allocator, spare; warps 4-11: epilogue

template <int BN, int KC, bool BF>
__global__ void __launch_bounds__(kFpropThreads, 1)
conv_fprop_tc_kernel(const __grid_constant__ TMapSet tmAs, const __grid_constant__ CUtensorMap tmB, const FpropParams p) {
  using Cfg = FpropCfg<BN, KC>;
  constexpr int STAGES = Cfg::kStages;
  constexpr int kChunk = BF ? 64 : 32;              // channels per 128-byte operand row (shadows the fp32 constant)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;  // 128B swizzle atoms need 1024-byte alignment
  uint8_t* smem = smem_raw + (base - raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::kStageBytes);
  const uint32_t bar0 = base + STAGES * Cfg::kStageBytes;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * STAGES + 2 + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k_iters = p.R * p.S * (p.Ci / kChunk);
  const int chunks = p.Ci / kChunk;
  long long* tr = p.trace ? p.trace + 8 * blockIdx.x : nullptr;
  if (tr && threadIdx.x == 0) tr[0] = clock64();

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmAs.m[0]);
    prefetch_tmap(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), Cfg::kEpiWarps);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(tmem_slot), Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tr && threadIdx.x == 0) tr[1] = clock64();   // setup done

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < p.num_tiles; item += gridDim.x) {
        const int ks = item % p.ksplit;
        int t = item / p.ksplit;
        const int tco = t % p.tiles_co; t /= p.tiles_co;
        int ph = 0;
        if (p.up == 1) { ph = t & 3; t >>= 2; }
        const int tw = t % p.tiles_w; t /= p.tiles_w;
        const int th = t % p.tiles_h; t /= p.tiles_h;
        const int tn = t;
        const int w0 = tw * p.bw, h0 = th * p.bh, n0 = tn * p.bn, co0 = tco * BN + ph * p.Co;
        const int pad_h = p.up == 1 ? 1 - (ph >> 1) : p.pad, pad_w = p.up == 1 ? 1 - (ph & 1) : p.pad;
        const int k0 = ks * p.k_per, k1 = min(k_iters, k0 + p.k_per);
        int tap = k0 / chunks, ch = k0 - tap * chunks;
        for (int k = k0; k < k1; k += KC) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          mbar_expect_tx(full_bar(stage), Cfg::kStageBytes);
#pragma unroll
          for (int c = 0; c < KC; ++c) {
            int mi = 0, dh, dw;
            if (p.up == 2) {                              // tap = phase * 4 + a * 2 + b (taps already flipped in the weights)
              mi = tap >> 2; dh = ((tap >> 1) & 1) - (mi >> 1); dw = (tap & 1) - (mi & 1);
            } else {
              const int r = tap / p.S, s = tap - r * p.S;
              dh = r - pad_h; dw = s - pad_w;
            }
            const uint32_t a_dst = base + stage * Cfg::kStageBytes + c * Cfg::kChunkBytes;
            tma_load_4d(a_dst, &tmAs.m[mi], full_bar(stage), ch * kChunk, w0 + dw, h0 + dh, n0);
            tma_load_3d(a_dst + kABytes, &tmB, full_bar(stage), ch * kChunk, tap, co0);
            if (++ch == chunks) { ch = 0; ++tap; }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (single thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc<BF>(kBM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int item = blockIdx.x; item < p.num_tiles; item += gridDim.x) {
        mbar_wait(tempty_bar(as), aphase ^ 1);  // epilogue has drained this accumulator stage
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        const int ks = item % p.ksplit;
        const int k_cnt = min(k_iters, (ks + 1) * p.k_per) - ks * p.k_per;
        for (int k = 0; k < k_cnt; k += KC) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          if (tr && k == 0 && item == (int)blockIdx.x) tr[2] = clock64();   // first operands landed
#pragma unroll
          for (int c = 0; c < KC; ++c) {
            const uint32_t a_addr = base + stage * Cfg::kStageBytes + c * Cfg::kChunkBytes;
            const uint32_t b_addr = a_addr + kABytes;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {  // 4 x (K = 8 tf32 = 32 bytes) per 128-byte swizzle row
              const uint64_t ad = make_smem_desc(a_addr + kk * 32, 16, 1024);
              const uint64_t bd = make_smem_desc(b_addr + kk * 32, 16, 1024);
              mma_ss<BF>(d_tmem, ad, bd, idesc, (k | c | kk) ? 1u : 0u);
            }
          }
          mma_commit(empty_bar(stage));  // frees the smem stage once these MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        mma_commit(tfull_bar(as));  // accumulator complete -> epilogue
        if (tr && item == (int)blockIdx.x) tr[3] = clock64();                // first tile: all MMAs issued
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
      if (tr) tr[4] = clock64();                                            // all tiles: all MMAs issued
    }
  } else if (warp >= 4 && warp < 4 + Cfg::kEpiWarps) {
    // ===================== epilogue: TMEM -> regs -> alpha/bias/act -> global (NHWC) =====================
    // A thread owns one output pixel (TMEM lane) and writes its 32-column fragments with 16-byte stores.  Measured
    // alternatives (clock64 trace, tools/trace_fprop.py): a shared-memory-staged TMA tile store and a per-warp staged,
    // fully coalesced store were both SLOWER (16.8k / 21k vs 10.7k cycles per 128x256 tile when all CTAs drain at once):
    // the burst of 128 KB per CTA from every SM at the same moment is bound by the chip's write path (~3 TB/s), not by
    // the store instruction pattern.
    const int q = warp & 3;  // TMEM lane quarter this warp may access (warp id % 4)
    constexpr int CW = (Cfg::kEpiWarps == 8) ? BN / 2 : BN;   // columns per epilogue warp
    const int c_begin = ((warp - 4) >> 2) * CW;
    const int row = q * 32 + lane;
    const int rw = row % p.bw, rh = (row / p.bw) % p.bh, rn = row / (p.bw * p.bh);
    int as = 0;
    uint32_t aphase = 0;
    for (int item = blockIdx.x; item < p.num_tiles; item += gridDim.x) {
      int t = item / p.ksplit;
      const int tco = t % p.tiles_co; t /= p.tiles_co;
      int ph = 0;
      if (p.up == 1) { ph = t & 3; t >>= 2; }
      const int tw = t % p.tiles_w; t /= p.tiles_w;
      const int th = t % p.tiles_h; t /= p.tiles_h;
      const int tn = t;
      const int w = tw * p.bw + rw, h = th * p.bh + rh, n = tn * p.bn + rn, co0 = tco * BN;
      const bool valid = (w < p.Wo) && (h < p.Ho) && (n < p.N);
      const int oh = p.up == 1 ? 2 * h + (ph >> 1) : h, ow = p.up == 1 ? 2 * w + (ph & 1) : w;
      float* out = p.y + (((int64_t)n * p.OH + oh) * p.OW + ow) * p.Co + co0;
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      if (tr && warp == 4 && lane == 0 && item == (int)blockIdx.x) tr[5] = clock64();   // first accumulator complete
      constexpr int EC = (BN < 32) ? BN : 32;
#pragma unroll 1
      for (int c = c_begin; c < c_begin + CW; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + as * BN + c, v);
        tmem_ld_wait();
        if (valid) {
          if (p.ksplit > 1) {  // split-K partial: summed into the zeroed output; bias / activation run as a second pass
#pragma unroll
            for (int j = 0; j < EC; j += 4)
              red_add_v4(out + c + j, p.alpha * __uint_as_float(v[j]), p.alpha * __uint_as_float(v[j + 1]),
                         p.alpha * __uint_as_float(v[j + 2]), p.alpha * __uint_as_float(v[j + 3]));
          } else {
#pragma unroll
            for (int j = 0; j < EC; j += 4) {
              float4 o;
              float* oo = reinterpret_cast<float*>(&o);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float a = p.alpha * __uint_as_float(v[j + e]);
                if (p.bias) a += p.bias_scale * __ldg(p.bias + co0 + c + j + e);
                oo[e] = act_apply(a, p.act, p.slope);
              }
              *reinterpret_cast<float4*>(out + c + j) = o;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(as));
      if (tr && warp == 4 && lane == 0 && item == (int)blockIdx.x) tr[6] = clock64();   // first epilogue done
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
    if (tr && warp == 4 && lane == 0) tr[7] = clock64();                                 // last epilogue done
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------ two M tiles per item
// Same implicit GEMM with TWO 128-pixel M tiles per work item sharing one B tile (for N tiles of 128 channels).  With a
// single 128 x 128 tile every 32-channel K step reads 16 KB of A and 16 KB of B from shared memory for 64 "MMA cycles" of
// work -- the SS-mode MMA is then bound by shared-memory operand bandwidth and by L2 -> SM traffic (measured 445-525
// TFLOP/s against 630-745 for the 256-wide tiles).  Two accumulators per B tile bring bytes per flop down to what the
// BN = 256 configuration has (48 KB per 2 x 64 cycles): stage = [A0 | A1 | B], 8 MMAs per stage, TMEM = 2 stages x 2 tiles x
// BN columns.  Used only without split-K and with an even number of M tiles (mt1 = mt0 + 1 in (n, h, w) tile order).
template <int BN>
struct FpropM2Cfg {
  static constexpr int kBBytes = BN * 128;
  static constexpr int kStageBytes = 2 * kABytes + kBBytes;
  static constexpr int kStages = (192 * 1024 / kStageBytes) > 8 ? 8 : (192 * 1024 / kStageBytes);
  static constexpr int kTmemCols = 4 * BN;   // 2 accumulator stages x 2 tiles
  static constexpr int kEpiWarps = 8;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
};

template <int BN>
__global__ void __launch_bounds__(kFpropThreads, 1)
conv_fprop_tc_m2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const FpropParams p) {
  using Cfg = FpropM2Cfg<BN>;
  static_assert(Cfg::kTmemCols <= 512, "TMEM has 512 columns");
  constexpr int STAGES = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::kStageBytes);
  const uint32_t bar0 = base + STAGES * Cfg::kStageBytes;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * STAGES + 2 + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunks = p.Ci / kChunk;
  const int k_iters = p.R * p.S * chunks;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), Cfg::kEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(tmem_slot), Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // item -> (pair of M tiles, N tile); M tile index mt = (tn * tiles_h + th) * tiles_w + tw
  auto tile_origin = [&](int mt, int& w0, int& h0, int& n0) {
    const int tw = mt % p.tiles_w;
    int t = mt / p.tiles_w;
    const int th = t % p.tiles_h;
    const int tn = t / p.tiles_h;
    w0 = tw * p.bw; h0 = th * p.bh; n0 = tn * p.bn;
  };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < p.num_tiles; item += gridDim.x) {
        const int tco = item % p.tiles_co, mt0 = 2 * (item / p.tiles_co);
        int w0[2], h0[2], n0[2];
        tile_origin(mt0, w0[0], h0[0], n0[0]);
        tile_origin(mt0 + 1, w0[1], h0[1], n0[1]);
        const int co0 = tco * BN;
        int tap = 0, ch = 0;
        for (int k = 0; k < k_iters; ++k) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          mbar_expect_tx(full_bar(stage), Cfg::kStageBytes);
          const int r = tap / p.S, s = tap - r * p.S;
          const uint32_t dst = base + stage * Cfg::kStageBytes;
          tma_load_4d(dst, &tmA, full_bar(stage), ch * kChunk, w0[0] + s - p.pad, h0[0] + r - p.pad, n0[0]);
          tma_load_4d(dst + kABytes, &tmA, full_bar(stage), ch * kChunk, w0[1] + s - p.pad, h0[1] + r - p.pad, n0[1]);
          tma_load_3d(dst + 2 * kABytes, &tmB, full_bar(stage), ch * kChunk, tap, co0);
          if (++ch == chunks) { ch = 0; ++tap; }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(kBM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int item = blockIdx.x; item < p.num_tiles; item += gridDim.x) {
        mbar_wait(tempty_bar(as), aphase ^ 1);
        tc_fence_after();
        const uint32_t d0 = tmem_base + as * (2 * BN), d1 = d0 + BN;
        for (int k = 0; k < k_iters; ++k) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t a0 = base + stage * Cfg::kStageBytes, a1 = a0 + kABytes, b = a0 + 2 * kABytes;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t bd = make_smem_desc(b + kk * 32, 16, 1024);
            mma_tf32(d0, make_smem_desc(a0 + kk * 32, 16, 1024), bd, idesc, (k | kk) ? 1u : 0u);
            mma_tf32(d1, make_smem_desc(a1 + kk * 32, 16, 1024), bd, idesc, (k | kk) ? 1u : 0u);
          }
          mma_commit(empty_bar(stage));
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        mma_commit(tfull_bar(as));
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else if (warp >= 4 && warp < 4 + Cfg::kEpiWarps) {
    const int q = warp & 3;
    constexpr int CW = BN / 2;                         // columns per epilogue warp (two warps per TMEM lane quarter)
    const int c_begin = ((warp - 4) >> 2) * CW;
    const int row = q * 32 + lane;
    const int rw = row % p.bw, rh = (row / p.bw) % p.bh, rn = row / (p.bw * p.bh);
    int as = 0;
    uint32_t aphase = 0;
    for (int item = blockIdx.x; item < p.num_tiles; item += gridDim.x) {
      const int tco = item % p.tiles_co, mt0 = 2 * (item / p.tiles_co);
      const int co0 = tco * BN;
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
#pragma unroll 1
      for (int mt = 0; mt < 2; ++mt) {
        int w0, h0, n0;
        tile_origin(mt0 + mt, w0, h0, n0);
        const int w = w0 + rw, h = h0 + rh, n = n0 + rn;
        const bool valid = (w < p.Wo) && (h < p.Ho) && (n < p.N);
        float* out = p.y + (((int64_t)n * p.Ho + h) * p.Wo + w) * p.Co + co0;
#pragma unroll 1
        for (int c = c_begin; c < c_begin + CW; c += 32) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + as * (2 * BN) + mt * BN + c, v);
          tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 o;
              float* oo = reinterpret_cast<float*>(&o);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float a = p.alpha * __uint_as_float(v[j + e]);
                if (p.bias) a += p.bias_scale * __ldg(p.bias + co0 + c + j + e);
                oo[e] = act_apply(a, p.act, p.slope);
              }
              *reinterpret_cast<float4*>(out + c + j) = o;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(as));
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------ tap reuse ("halo") fprop
// The kernels above re-read the input window once per filter tap: R*S = 9 TMA loads of 16 KB per 32-channel chunk and tile.
// Measured (ncu, clock64 traces): every large layer then moves ~16 KB of A per ~660 cycles and SM whatever the N tile
// (128 wide, 256 wide as a CTA pair, two M tiles per B tile) -- the L2 -> SM path for the per-SM-unique A data is the
// bound (~7 TB/s chip wide), not the tensor pipe (256 / 512 MMA cycles per chunk step) and not shared memory.
// Here the M tile is 16 rows x 8 columns of one image, so that one image row of the tile is exactly one 1024-byte
// swizzle atom (8 pixels x 128 B).  Per chunk S boxes are loaded, one per COLUMN shift s: (32 ch, 8 wide, 16+R-1 tall) =
// 18 KB each; the ROW shift r of a tap is then just a start-address offset of r atoms (r * 1024 B) into box s -- every
// descriptor stays canonical (SBO = 1024 B, 1024-byte aligned starts).  A traffic per chunk and tile: S * 18 KB = 54 KB
// instead of 144 KB.  Two rings: A (one stage = the S boxes of a chunk, released after its R*S taps) and B (one stage
// = one tap of one chunk).
// ONEBOX: ONE box per chunk, (32 ch, 8 + S - 1 wide, 16 + R - 1 tall), covering every tap: the column shift s becomes a start-address
// offset of s pixels (s * 128 B) as well, with SBO = (8 + S - 1) * 128 B between the 8-pixel row groups of the MMA tile.  The
// start is then no longer 1024-byte aligned; the 128B swizzle is a function of the shared-memory ADDRESS bits (chunk ^= addr[7:9],
// what TMA applies on the way in), so a 128-byte-aligned start inside the pattern addresses the same bytes.  A traffic per chunk
// and tile: 22.5 KB instead of 54 KB (the loop is bound by L2 -> SM bytes: ncu lts2xbar at 82 % of its peak).
template <int BN, int R_, int S_, bool ONEBOX = false>
struct FpropHaloCfg {
  static constexpr int kBoxRows = 16 + R_ - 1;
  static constexpr int kBoxCols = ONEBOX ? 8 + S_ - 1 : 8;
  static constexpr int kBoxBytes = ONEBOX ? ((kBoxRows * kBoxCols * 128 + 1023) / 1024) * 1024 : kBoxRows * 1024;
  static constexpr int kAStageBytes = (ONEBOX ? 1 : S_) * kBoxBytes;
  static constexpr int kTxBytes = ONEBOX ? kBoxRows * kBoxCols * 128 : kAStageBytes;   // bytes TMA actually writes per A stage
  static constexpr int kRowPitch = kBoxCols * 128;                                     // SBO: one image row of the box
  // taps per B stage: the MMA-issuing thread pays a fixed wait / fence / commit cost per stage, so the 64-cycle MMAs of a
  // 128-wide N tile get 8 per stage (cf. FpropCfg's KC)
  static constexpr int kTPS = 1;                 // must divide R*S (stages never straddle a chunk; everything unrolls)
  static constexpr int kTapBytes = BN * 128;
  static constexpr int kBStageBytes = kTPS * kTapBytes;
  static constexpr int kAStages = 2;
  static constexpr int kBStages = (222 * 1024 - kAStages * kAStageBytes) / kBStageBytes > 8
                                      ? 8 : (222 * 1024 - kAStages * kAStageBytes) / kBStageBytes;
  static constexpr int kTmemCols = 2 * BN;
  static constexpr int kEpiWarps = 8;
  static constexpr int kSmemBytes = kAStages * kAStageBytes + kBStages * kBStageBytes + 1024 + 256;
};

template <int BN, int R_, int S_, bool ONEBOX>
__global__ void __launch_bounds__(kFpropThreads, 1)
conv_fprop_tc_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const FpropParams p) {
  using Cfg = FpropHaloCfg<BN, R_, S_, ONEBOX>;
  static_assert(Cfg::kBStages >= 3 || Cfg::kTPS > 1, "B ring too shallow");
  constexpr int NA = Cfg::kAStages, NB = Cfg::kBStages, TAPS = R_ * S_, TPS = Cfg::kTPS;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t a_base = base, b_base = base + NA * Cfg::kAStageBytes;
  const uint32_t bar0 = b_base + NB * Cfg::kBStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NA * Cfg::kAStageBytes + NB * Cfg::kBStageBytes);
  auto afull = [&](int s) { return bar0 + 8u * s; };
  auto aempty = [&](int s) { return bar0 + 8u * (NA + s); };
  auto bfull = [&](int s) { return bar0 + 8u * (2 * NA + s); };
  auto bempty = [&](int s) { return bar0 + 8u * (2 * NA + NB + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * NA + 2 * NB + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * NA + 2 * NB + 2 + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NA + 2 * NB + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunks = p.Ci / kChunk;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NA; ++s) { mbar_init(afull(s), 1); mbar_init(aempty(s), 1); }
    for (int s = 0; s < NB; ++s) { mbar_init(bfull(s), 1); mbar_init(bempty(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), Cfg::kEpiWarps); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(tmem_slot), Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      for (int item = blockIdx.x; item < p.num_tiles; item += gridDim.x) {
        int t = item;
        const int tco = t % p.tiles_co; t /= p.tiles_co;
        const int tw = t % p.tiles_w; t /= p.tiles_w;
        const int th = t % p.tiles_h; t /= p.tiles_h;
        const int n0 = t;
        const int w0 = tw * 8, h0 = th * 16, co0 = tco * BN;
        for (int ch = 0; ch < chunks; ++ch) {
          mbar_wait(aempty(sa), pa ^ 1);
          mbar_expect_tx(afull(sa), Cfg::kTxBytes);
          if (ONEBOX) {
            tma_load_4d(a_base + sa * Cfg::kAStageBytes, &tmA, afull(sa), ch * kChunk, w0 - p.pad, h0 - p.pad, n0);
          } else {
#pragma unroll
            for (int s = 0; s < S_; ++s)
              tma_load_4d(a_base + sa * Cfg::kAStageBytes + s * Cfg::kBoxBytes, &tmA, afull(sa), ch * kChunk, w0 + s - p.pad,
                          h0 - p.pad, n0);
          }
          if (++sa == NA) { sa = 0; pa ^= 1; }
#pragma unroll
          for (int g = 0; g < TAPS / TPS; ++g) {              // one B stage = TPS taps of this chunk
            mbar_wait(bempty(sb), pb ^ 1);
            mbar_expect_tx(bfull(sb), Cfg::kBStageBytes);
#pragma unroll
            for (int u = 0; u < TPS; ++u)
              tma_load_3d(b_base + sb * Cfg::kBStageBytes + u * Cfg::kTapBytes, &tmB, bfull(sb), ch * kChunk, g * TPS + u, co0);
            if (++sb == NB) { sb = 0; pb ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (single thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(kBM, BN, 0, 0);
      int sa = 0, sb = 0, as = 0;
      uint32_t pa = 0, pb = 0, aphase = 0;
      for (int item = blockIdx.x; item < p.num_tiles; item += gridDim.x) {
        mbar_wait(tempty_bar(as), aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        // The issuing thread is the pacing item for 64-cycle MMAs: taps are unrolled, so every descriptor is the stage's base
        // descriptor plus a compile-time constant in its 14-bit address field ((offset) >> 4; shared memory < 256 KB: no carry).
        for (int ch = 0; ch < chunks; ++ch) {
          mbar_wait(afull(sa), pa);
          tc_fence_after();
          const uint64_t ad0 = make_smem_desc(a_base + sa * Cfg::kAStageBytes, 16, Cfg::kRowPitch);
#pragma unroll
          for (int g = 0; g < TAPS / TPS; ++g) {
            mbar_wait(bfull(sb), pb);
            tc_fence_after();
            const uint64_t bd0 = make_smem_desc(b_base + sb * Cfg::kBStageBytes, 16, 1024);
#pragma unroll
            for (int u = 0; u < TPS; ++u) {
              const int tap = g * TPS + u, r = tap / S_, s = tap - r * S_;     // row shift = whole box rows
              const int a_off = ONEBOX ? (r * Cfg::kBoxCols + s) * 128 : s * Cfg::kBoxBytes + r * 1024;
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                mma_tf32(d_tmem, ad0 + (uint64_t)((a_off + kk * 32) >> 4),
                         bd0 + (uint64_t)((u * Cfg::kTapBytes + kk * 32) >> 4), idesc, (ch | tap | kk) ? 1u : 0u);
            }
            mma_commit(bempty(sb));
            if (++sb == NB) { sb = 0; pb ^= 1; }
          }
          mma_commit(aempty(sa));      // the boxes of this chunk are free once all R*S taps have read them
          if (++sa == NA) { sa = 0; pa ^= 1; }
        }
        mma_commit(tfull_bar(as));
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else if (warp >= 4 && warp < 4 + Cfg::kEpiWarps) {
    // ===================== epilogue (TMEM lane = pixel (row / 8, row % 8) of the 16 x 8 tile) =====================
    const int q = warp & 3;
    constexpr int CW = BN / 2;
    const int c_begin = ((warp - 4) >> 2) * CW;
    const int row = q * 32 + lane;
    const int rw = row & 7, rh = row >> 3;
    int as = 0;
    uint32_t aphase = 0;
    for (int item = blockIdx.x; item < p.num_tiles; item += gridDim.x) {
      int t = item;
      const int tco = t % p.tiles_co; t /= p.tiles_co;
      const int tw = t % p.tiles_w; t /= p.tiles_w;
      const int th = t % p.tiles_h; t /= p.tiles_h;
      const int n = t;
      const int w = tw * 8 + rw, h = th * 16 + rh, co0 = tco * BN;
      const bool valid = (w < p.Wo) && (h < p.Ho);
      float* out = p.y + (((int64_t)n * p.Ho + h) * p.Wo + w) * p.Co + co0;
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
#pragma unroll 1
      for (int c = c_begin; c < c_begin + CW; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + as * BN + c, v);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 o;
            float* oo = reinterpret_cast<float*>(&o);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float a = p.alpha * __uint_as_float(v[j + e]);
              if (p.bias) a += p.bias_scale * __ldg(p.bias + co0 + c + j + e);
              oo[e] = act_apply(a, p.act, p.slope);
            }
            *reinterpret_cast<float4*>(out + c + j) = o;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(as));
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------ CTA-pair fprop
// Same implicit GEMM with tcgen05 cta_group::2: a cluster of two CTAs (one TPC) computes a 256-pixel x BN tile.  Each CTA
// stages its own 128-pixel A tile and HALF of the B (weight) tile; one MMA issued by the leader reads A from both CTAs
// and the two B halves (M = 256, N = BN).  Per SM and K step that is (16 KB + BN*64 B) of TMA fill and the same of operand
// reads instead of (16 KB + BN*128 B) -- the single-CTA kernel is bound by exactly that shared-memory traffic (ncu: tensor
// pipe 68-71 % at BN = 256, 46 % at BN = 128, while a bare tcgen05.mma loop reaches 100 %, tools/micro/mma_rate.cu).
// Barriers: full[] lives in the leader (count 1 = its arrive.expect_tx for BOTH CTAs' bytes; both CTAs' TMA complete_tx on
// it), empty[] / tmem_full[] exist in both CTAs and are arrived by the leader's multicast tcgen05.commit, tmem_empty[] lives
// in the leader and collects the 8 epilogue warps of the pair.
template <int BN>
struct Fprop2Cfg {
  static constexpr int kBBytes = (BN / 2) * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN >= 256) ? 6 : 8;
  static constexpr int kTmemCols = 2 * BN;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
};

template <int BN, bool BF>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kFpropThreads, 1)
conv_fprop_tc2_kernel(const __grid_constant__ TMapSet tmAs, const __grid_constant__ CUtensorMap tmB, const FpropParams p) {
  using Cfg = Fprop2Cfg<BN>;
  constexpr int STAGES = Cfg::kStages;
  constexpr int kChunk = BF ? 64 : 32;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::kStageBytes);
  const uint32_t bar0 = base + STAGES * Cfg::kStageBytes;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * STAGES + 2 + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = (rank == 0);
  const int chunks = p.Ci / kChunk;
  const int k_iters = p.R * p.S * chunks;
  const int pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmAs.m[0]);
    prefetch_tmap(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 16);  // 8 epilogue warps in each CTA of the pair
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc2(smem_u32(tmem_slot), Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();          // (the cluster barrier below orders the CTA as well; compute-sanitizer racecheck only models this one)
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = pair_id; item < p.num_tiles; item += num_pairs) {
        const int tco = item % p.tiles_co;
        int u = item / p.tiles_co, ph = 0;
        if (p.up == 1) { ph = u & 3; u >>= 2; }       // both M tiles of a pair belong to one output phase (they share B)
        int t = u * 2 + (int)rank;                    // this CTA's M tile
        const int tw = t % p.tiles_w; t /= p.tiles_w;
        const int th = t % p.tiles_h; t /= p.tiles_h;
        const int tn = t;
        const int w0 = tw * p.bw, h0 = th * p.bh, n0 = tn * p.bn;
        const int co0 = tco * BN + (int)rank * (BN / 2) + ph * p.Co;  // this CTA's half of the weight tile
        const int pad_h = p.up == 1 ? 1 - (ph >> 1) : p.pad, pad_w = p.up == 1 ? 1 - (ph & 1) : p.pad;
        int tap = 0, ch = 0;
        for (int k = 0; k < k_iters; ++k) {
          int mi = 0, dh, dw;
          if (p.up == 2) {
            mi = tap >> 2; dh = ((tap >> 1) & 1) - (mi >> 1); dw = (tap & 1) - (mi & 1);
          } else {
            const int r = tap / p.S, s = tap - r * p.S;
            dh = r - pad_h; dw = s - pad_w;
          }
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t a_dst = base + stage * Cfg::kStageBytes;
          const uint32_t b_dst = a_dst + kABytes;
          if (leader) mbar_expect_tx(full_bar(stage), 2 * Cfg::kStageBytes);
          const uint32_t lfull = mapa(full_bar(stage), 0);
          tma2_load_4d(a_dst, &tmAs.m[mi], lfull, ch * kChunk, w0 + dw, h0 + dh, n0);
          tma2_load_3d(b_dst, &tmB, lfull, ch * kChunk, tap, co0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
          if (++ch == chunks) { ch = 0; ++tap; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA, one thread) =====================
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc<BF>(2 * kBM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int item = pair_id; item < p.num_tiles; item += num_pairs) {
        mbar_wait(tempty_bar(as), aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int k = 0; k < k_iters; ++k) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t a_addr = base + stage * Cfg::kStageBytes;
          const uint32_t b_addr = a_addr + kABytes;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t ad = make_smem_desc(a_addr + kk * 32, 16, 1024);
            const uint64_t bd = make_smem_desc(b_addr + kk * 32, 16, 1024);
            mma2_ss<BF>(d_tmem, ad, bd, idesc, (k | kk) ? 1u : 0u);
          }
          mma2_commit_mc(empty_bar(stage), 3);  // frees this stage in both CTAs
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        mma2_commit_mc(tfull_bar(as), 3);  // accumulators complete -> both epilogues
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs): own 128 rows of the pair's accumulator =====================
    const int q = warp & 3;
    const int c_begin = ((warp - 4) >> 2) * (BN / 2);  // two warps per TMEM lane quarter, half the columns each
    const int row = q * 32 + lane;
    const int rw = row % p.bw, rh = (row / p.bw) % p.bh, rn = row / (p.bw * p.bh);
    int as = 0;
    uint32_t aphase = 0;
    for (int item = pair_id; item < p.num_tiles; item += num_pairs) {
      const int tco = item % p.tiles_co;
      int u = item / p.tiles_co, ph = 0;
      if (p.up == 1) { ph = u & 3; u >>= 2; }
      int t = u * 2 + (int)rank;
      const int tw = t % p.tiles_w; t /= p.tiles_w;
      const int th = t % p.tiles_h; t /= p.tiles_h;
      const int tn = t;
      const int w = tw * p.bw + rw, h = th * p.bh + rh, n = tn * p.bn + rn, co0 = tco * BN;
      const bool valid = (w < p.Wo) && (h < p.Ho) && (n < p.N);
      const int oh = p.up == 1 ? 2 * h + (ph >> 1) : h, ow = p.up == 1 ? 2 * w + (ph & 1) : w;
      float* out = p.y + (((int64_t)n * p.OH + oh) * p.OW + ow) * p.Co + co0;
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
#pragma unroll 1
      for (int c = c_begin; c < c_begin + BN / 2; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + as * BN + c, v);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 o;
            float* oo = reinterpret_cast<float*>(&o);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float a = p.alpha * __uint_as_float(v[j + e]);
              if (p.bias) a += p.bias_scale * __ldg(p.bias + co0 + c + j + e);
              oo[e] = act_apply(a, p.act, p.slope);
            }
            *reinterpret_cast<float4*>(out + c + j) = o;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa(tempty_bar(as), 0));
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, Cfg::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------ CTA-pair + tap reuse
// The two ideas combined: a CTA pair (cta_group::2, M = 256 = two 16 x 8 pixel tiles, each CTA stages its own tile and HALF
// of the weight tile) with the tap-reuse A staging of conv_fprop_tc_halo_kernel.  Per SM and tap step: 6 KB of A
// (54 KB per chunk / 9 taps) + BN*64 B of B instead of 16 KB + BN*64 B (pair) or 16 KB + BN*128 B (single CTA).
// Rings / barriers: A ring (stage = the S boxes of one chunk) and B ring (stage = one tap of one chunk) in both CTAs;
// afull[] / bfull[] live in the leader and count BOTH CTAs' bytes; aempty[] / bempty[] / tmem_full[] exist in both CTAs and
// are arrived by the leader's multicast tcgen05.commit; tmem_empty[] lives in the leader (16 epilogue warps of the pair).
template <int BN>
struct Fprop2HaloCfg {
  static constexpr int kBoxBytes = 18 * 1024;
  static constexpr int kAStageBytes = 3 * kBoxBytes;
  static constexpr int kTPS = (BN <= 128) ? 3 : 1;          // taps per B stage (12 x 64-cycle MMAs per commit at BN = 128)
  static constexpr int kTapBytes = (BN / 2) * 128;
  static constexpr int kBStageBytes = kTPS * kTapBytes;
  static constexpr int kAStages = 2;
  static constexpr int kBStages = (BN >= 256) ? 6 : 4;
  static constexpr int kTmemCols = 2 * BN;
  static constexpr int kSmemBytes = kAStages * kAStageBytes + kBStages * kBStageBytes + 1024 + 256;
};

template <int BN, bool BF>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kFpropThreads, 1)
conv_fprop_tc2_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const FpropParams p) {
  using Cfg = Fprop2HaloCfg<BN>;
  constexpr int NA = Cfg::kAStages, NB = Cfg::kBStages, S_ = 3, TAPS = 9, TPS = Cfg::kTPS;
  constexpr int kChunk = BF ? 64 : 32;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t a_base = base, b_base = base + NA * Cfg::kAStageBytes;
  const uint32_t bar0 = b_base + NB * Cfg::kBStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NA * Cfg::kAStageBytes + NB * Cfg::kBStageBytes);
  auto afull = [&](int s) { return bar0 + 8u * s; };
  auto aempty = [&](int s) { return bar0 + 8u * (NA + s); };
  auto bfull = [&](int s) { return bar0 + 8u * (2 * NA + s); };
  auto bempty = [&](int s) { return bar0 + 8u * (2 * NA + NB + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * NA + 2 * NB + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * NA + 2 * NB + 2 + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NA + 2 * NB + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = (rank == 0);
  const int chunks = p.Ci / kChunk;
  const int pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NA; ++s) { mbar_init(afull(s), 1); mbar_init(aempty(s), 1); }
    for (int s = 0; s < NB; ++s) { mbar_init(bfull(s), 1); mbar_init(bempty(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 16); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc2(smem_u32(tmem_slot), Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();          // (the cluster barrier below orders the CTA as well; compute-sanitizer racecheck only models this one)
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // this CTA's M tile of pair item `item`: tiles ordered (n, th, tw), tw fastest; 16 x 8 pixels each
  auto my_tile = [&](int item, int& w0, int& h0, int& n0) {
    int t = (item / p.tiles_co) * 2 + (int)rank;
    const int tw = t % p.tiles_w; t /= p.tiles_w;
    const int th = t % p.tiles_h; t /= p.tiles_h;
    w0 = tw * 8; h0 = th * 16; n0 = t;
  };

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      for (int item = pair_id; item < p.num_tiles; item += num_pairs) {
        int w0, h0, n0;
        my_tile(item, w0, h0, n0);
        const int co0 = (item % p.tiles_co) * BN + (int)rank * (BN / 2);   // this CTA's half of the weight tile
        for (int ch = 0; ch < chunks; ++ch) {
          mbar_wait(aempty(sa), pa ^ 1);
          if (leader) mbar_expect_tx(afull(sa), 2 * Cfg::kAStageBytes);
          const uint32_t lafull = mapa(afull(sa), 0);
#pragma unroll
          for (int s = 0; s < S_; ++s)
            tma2_load_4d(a_base + sa * Cfg::kAStageBytes + s * Cfg::kBoxBytes, &tmA, lafull, ch * kChunk, w0 + s - p.pad, h0 - p.pad, n0);
          if (++sa == NA) { sa = 0; pa ^= 1; }
#pragma unroll
          for (int g = 0; g < TAPS / TPS; ++g) {
            mbar_wait(bempty(sb), pb ^ 1);
            if (leader) mbar_expect_tx(bfull(sb), 2 * Cfg::kBStageBytes);
            const uint32_t lbfull = mapa(bfull(sb), 0);
#pragma unroll
            for (int u = 0; u < TPS; ++u)
              tma2_load_3d(b_base + sb * Cfg::kBStageBytes + u * Cfg::kTapBytes, &tmB, lbfull, ch * kChunk, g * TPS + u, co0);
            if (++sb == NB) { sb = 0; pb ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA, one thread) =====================
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc<BF>(2 * kBM, BN, 0, 0);
      int sa = 0, sb = 0, as = 0;
      uint32_t pa = 0, pb = 0, aphase = 0;
      for (int item = pair_id; item < p.num_tiles; item += num_pairs) {
        mbar_wait(tempty_bar(as), aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int ch = 0; ch < chunks; ++ch) {
          mbar_wait(afull(sa), pa);
          tc_fence_after();
          const uint64_t ad0 = make_smem_desc(a_base + sa * Cfg::kAStageBytes, 16, 1024);
#pragma unroll
          for (int g = 0; g < TAPS / TPS; ++g) {
            mbar_wait(bfull(sb), pb);
            tc_fence_after();
            const uint64_t bd0 = make_smem_desc(b_base + sb * Cfg::kBStageBytes, 16, 1024);
#pragma unroll
            for (int u = 0; u < TPS; ++u) {
              const int tap = g * TPS + u, r = tap / S_, s = tap - r * S_;
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                mma2_ss<BF>(d_tmem, ad0 + (uint64_t)((s * Cfg::kBoxBytes + r * 1024 + kk * 32) >> 4),
                            bd0 + (uint64_t)((u * Cfg::kTapBytes + kk * 32) >> 4), idesc, (ch | tap | kk) ? 1u : 0u);
            }
            mma2_commit_mc(bempty(sb), 3);
            if (++sb == NB) { sb = 0; pb ^= 1; }
          }
          mma2_commit_mc(aempty(sa), 3);
          if (++sa == NA) { sa = 0; pa ^= 1; }
        }
        mma2_commit_mc(tfull_bar(as), 3);
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs): own 128 rows of the pair's accumulator =====================
    const int q = warp & 3;
    const int c_begin = ((warp - 4) >> 2) * (BN / 2);
    const int row = q * 32 + lane;
    const int rw = row & 7, rh = row >> 3;
    int as = 0;
    uint32_t aphase = 0;
    for (int item = pair_id; item < p.num_tiles; item += num_pairs) {
      int w0, h0, n;
      my_tile(item, w0, h0, n);
      const int w = w0 + rw, h = h0 + rh, co0 = (item % p.tiles_co) * BN;
      const bool valid = (w < p.Wo) && (h < p.Ho);
      float* out = p.y + (((int64_t)n * p.Ho + h) * p.Wo + w) * p.Co + co0;
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
#pragma unroll 1
      for (int c = c_begin; c < c_begin + BN / 2; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + as * BN + c, v);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 o;
            float* oo = reinterpret_cast<float*>(&o);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float a = p.alpha * __uint_as_float(v[j + e]);
              if (p.bias) a += p.bias_scale * __ldg(p.bias + co0 + c + j + e);
              oo[e] = act_apply(a, p.act, p.slope);
            }
            *reinterpret_cast<float4*>(out + c + j) = o;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa(tempty_bar(as), 0));
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, Cfg::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------ CTA pair + tap reuse for the folded layers
// conv_fprop_tc2_halo_kernel's staging for the 2x2 phase taps of the upsample- / pool-folded convolutions (FpropParams::up):
// per 32-channel chunk (and, for up = 2, input phase) a CTA loads TWO column-shifted boxes of 17 rows x 8 pixels once and the
// four taps read them at row offsets of 1024 B -- 8.5 KB of A per tap instead of the 16 KB the generic pair kernel re-loads.
// up = 1: the output phase is part of the M index (both tiles of a pair share it), pad = 1 - d per axis, strided output;
// up = 2: K runs over (chunk, input phase, tap): A comes from the phase's strided view (TMapSet), offsets (a' - dy, b' - dx).
template <int BN>
struct Fprop2FoldCfg {
  static constexpr int kBoxBytes = 17 * 1024;
  static constexpr int kAStageBytes = 2 * kBoxBytes;
  static constexpr int kTPS = (BN <= 128) ? 4 : 2;          // taps per B stage
  static constexpr int kTapBytes = (BN / 2) * 128;
  static constexpr int kBStageBytes = kTPS * kTapBytes;     // 32 KB either way
  static constexpr int kAStages = 3;
  static constexpr int kBStages = 3;
  static constexpr int kTmemCols = 2 * BN;
  static constexpr int kSmemBytes = kAStages * kAStageBytes + kBStages * kBStageBytes + 1024 + 256;
};

template <int BN, bool BF, int UP>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kFpropThreads, 1)
conv_fprop_tc2_fold_kernel(const __grid_constant__ TMapSet tmAs, const __grid_constant__ CUtensorMap tmB, const FpropParams p) {
  using Cfg = Fprop2FoldCfg<BN>;
  constexpr int NA = Cfg::kAStages, NB = Cfg::kBStages, TAPS = 4, TPS = Cfg::kTPS;
  constexpr int kChunk = BF ? 64 : 32;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t a_base = base, b_base = base + NA * Cfg::kAStageBytes;
  const uint32_t bar0 = b_base + NB * Cfg::kBStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NA * Cfg::kAStageBytes + NB * Cfg::kBStageBytes);
  auto afull = [&](int s) { return bar0 + 8u * s; };
  auto aempty = [&](int s) { return bar0 + 8u * (NA + s); };
  auto bfull = [&](int s) { return bar0 + 8u * (2 * NA + s); };
  auto bempty = [&](int s) { return bar0 + 8u * (2 * NA + NB + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * NA + 2 * NB + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * NA + 2 * NB + 2 + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NA + 2 * NB + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = (rank == 0);
  const int chunks = p.Ci / kChunk;
  const int groups = UP == 2 ? 4 * chunks : chunks;      // A stages per item: (chunk) or (chunk, input phase)
  const int pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmAs.m[0]);
    prefetch_tmap(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NA; ++s) { mbar_init(afull(s), 1); mbar_init(aempty(s), 1); }
    for (int s = 0; s < NB; ++s) { mbar_init(bfull(s), 1); mbar_init(bempty(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 16); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc2(smem_u32(tmem_slot), Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // this CTA's M tile of pair item `item` (16 x 8 pixels of the LOW-resolution grid) and, for up = 1, the output phase
  auto my_tile = [&](int item, int& w0, int& h0, int& n0, int& ph) {
    int u = item / p.tiles_co;
    ph = 0;
    if (UP == 1) { ph = u & 3; u >>= 2; }
    int t = u * 2 + (int)rank;
    const int tw = t % p.tiles_w; t /= p.tiles_w;
    const int th = t % p.tiles_h; t /= p.tiles_h;
    w0 = tw * 8; h0 = th * 16; n0 = t;
  };

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      for (int item = pair_id; item < p.num_tiles; item += num_pairs) {
        int w0, h0, n0, ph;
        my_tile(item, w0, h0, n0, ph);
        const int co0 = (item % p.tiles_co) * BN + (int)rank * (BN / 2) + ph * p.Co;   // this CTA's half of the weight tile
        for (int g = 0; g < groups; ++g) {
          const int ch = UP == 2 ? (g >> 2) : g, kph = UP == 2 ? (g & 3) : 0;
          // first row / column of the boxes: up = 1: (h0 - pad_h, w0 + b - pad_w) with pad = 1 - d of the OUTPUT phase;
          //                                  up = 2: (h0 - dy, w0 + b' - dx) with (dy, dx) of the INPUT phase kph
          const int oy = UP == 1 ? (ph >> 1) - 1 : -(kph >> 1), ox = UP == 1 ? (ph & 1) - 1 : -(kph & 1);
          mbar_wait(aempty(sa), pa ^ 1);
          if (leader) mbar_expect_tx(afull(sa), 2 * Cfg::kAStageBytes);
          const uint32_t lafull = mapa(afull(sa), 0);
#pragma unroll
          for (int b = 0; b < 2; ++b)
            tma2_load_4d(a_base + sa * Cfg::kAStageBytes + b * Cfg::kBoxBytes, &tmAs.m[kph], lafull, ch * kChunk, w0 + b + ox, h0 + oy, n0);
          if (++sa == NA) { sa = 0; pa ^= 1; }
#pragma unroll
          for (int gq = 0; gq < TAPS / TPS; ++gq) {
            mbar_wait(bempty(sb), pb ^ 1);
            if (leader) mbar_expect_tx(bfull(sb), 2 * Cfg::kBStageBytes);
            const uint32_t lbfull = mapa(bfull(sb), 0);
#pragma unroll
            for (int u = 0; u < TPS; ++u)
              tma2_load_3d(b_base + sb * Cfg::kBStageBytes + u * Cfg::kTapBytes, &tmB, lbfull, ch * kChunk, kph * 4 + gq * TPS + u, co0);
            if (++sb == NB) { sb = 0; pb ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA, one thread) =====================
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc<BF>(2 * kBM, BN, 0, 0);
      int sa = 0, sb = 0, as = 0;
      uint32_t pa = 0, pb = 0, aphase = 0;
      for (int item = pair_id; item < p.num_tiles; item += num_pairs) {
        mbar_wait(tempty_bar(as), aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int g = 0; g < groups; ++g) {
          mbar_wait(afull(sa), pa);
          tc_fence_after();
          const uint64_t ad0 = make_smem_desc(a_base + sa * Cfg::kAStageBytes, 16, 1024);
#pragma unroll
          for (int gq = 0; gq < TAPS / TPS; ++gq) {
            mbar_wait(bfull(sb), pb);
            tc_fence_after();
            const uint64_t bd0 = make_smem_desc(b_base + sb * Cfg::kBStageBytes, 16, 1024);
#pragma unroll
            for (int u = 0; u < TPS; ++u) {
              const int tap = gq * TPS + u, a = tap >> 1, b = tap & 1;       // row offset a (1024 B), column-shifted box b
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                mma2_ss<BF>(d_tmem, ad0 + (uint64_t)((b * Cfg::kBoxBytes + a * 1024 + kk * 32) >> 4),
                            bd0 + (uint64_t)((u * Cfg::kTapBytes + kk * 32) >> 4), idesc, (g | tap | kk) ? 1u : 0u);
            }
            mma2_commit_mc(bempty(sb), 3);
            if (++sb == NB) { sb = 0; pb ^= 1; }
          }
          mma2_commit_mc(aempty(sa), 3);
          if (++sa == NA) { sa = 0; pa ^= 1; }
        }
        mma2_commit_mc(tfull_bar(as), 3);
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs): own 128 rows of the pair's accumulator =====================
    const int q = warp & 3;
    const int c_begin = ((warp - 4) >> 2) * (BN / 2);
    const int row = q * 32 + lane;
    const int rw = row & 7, rh = row >> 3;
    int as = 0;
    uint32_t aphase = 0;
    for (int item = pair_id; item < p.num_tiles; item += num_pairs) {
      int w0, h0, n, ph;
      my_tile(item, w0, h0, n, ph);
      const int w = w0 + rw, h = h0 + rh, co0 = (item % p.tiles_co) * BN;
      const bool valid = (w < p.Wo) && (h < p.Ho);
      const int oh = UP == 1 ? 2 * h + (ph >> 1) : h, ow = UP == 1 ? 2 * w + (ph & 1) : w;
      float* out = p.y + (((int64_t)n * p.OH + oh) * p.OW + ow) * p.Co + co0;
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
#pragma unroll 1
      for (int c = c_begin; c < c_begin + BN / 2; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + as * BN + c, v);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 o;
            float* oo = reinterpret_cast<float*>(&o);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float a = p.alpha * __uint_as_float(v[j + e]);
              if (p.bias) a += p.bias_scale * __ldg(p.bias + co0 + c + j + e);
              oo[e] = act_apply(a, p.act, p.slope);
            }
            *reinterpret_cast<float4*>(out + c + j) = o;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa(tempty_bar(as), 0));
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, Cfg::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------ CTA pair, one-box tap reuse, MT tiles per B tile
// What bounds the kernels above on the large layers is the L2 -> SM path (ncu: lts2xbar at 82 % of its peak while the tensor
// pipe is 76 % active): per 32-channel chunk a CTA of conv_fprop_tc2_halo_kernel<256> pulls 54 KB of A (three column-shifted
// boxes) + 147 KB of B for 4608 MMA cycles = 44 B / cycle / SM against a chip-wide ~42 B / cycle / SM L2 output cap.  Here
//   * ONE box per M tile and chunk -- (32 ch, 10 wide, 18 tall) = 22.5 KB -- serves all nine taps: tap (r, s) is the start-address
//     offset (r * 10 + s) * 128 B into it, with SBO = 10 * 128 B between the 8-pixel row groups (see FpropHaloCfg ONEBOX);
//   * every B (weight) tile is used for MT = 2 M tiles per CTA (four per pair): two accumulators per TMEM stage (MT * BN * 2
//     stages = 512 columns at BN = 128), B bytes per flop halved.
// 2 * 22.5 + 72 = 117 KB per chunk and CTA for the same 4608 MMA cycles = 26 B / cycle / SM.
// Work is cut into UNITS = (one 16 x 8 pixel tile per CTA of the pair, one 128-channel N tile), N tile slowest; every pair owns a
// contiguous range of units (sizes differ by at most one: no wave quantisation) and processes them two at a time where the
// two share the N tile.
template <int BN, int MT>
struct Fprop2H1Cfg {
  static constexpr int kBoxCols = 10, kBoxRows = 18;
  static constexpr int kTileTx = kBoxRows * kBoxCols * 128;                  // bytes TMA writes per tile and chunk
  static constexpr int kTileBytes = ((kTileTx + 1023) / 1024) * 1024;       // swizzle pattern restarts on 1024-byte boundaries
  static constexpr int kAStageBytes = MT * kTileBytes;
  static constexpr int kRowPitch = kBoxCols * 128;
  static constexpr int kTPS = 3;                                              // taps per B stage
  static constexpr int kTapBytes = (BN / 2) * 128;
  static constexpr int kBStageBytes = kTPS * kTapBytes;
  static constexpr int kAStages = 2;
  static constexpr int kBStages = (BN >= 256) ? 3 : 4;
  static constexpr int kTmemCols = 2 * MT * BN;
  static constexpr int kSmemBytes = kAStages * kAStageBytes + kBStages * kBStageBytes + 1024 + 256;
};

struct H1Range { int begin, end; };
__device__ __forceinline__ H1Range h1_range(int units, int pair_id, int num_pairs) {
  H1Range r;
  r.begin = (int)(((long long)units * pair_id) / num_pairs);
  r.end = (int)(((long long)units * (pair_id + 1)) / num_pairs);
  return r;
}

template <int BN, int MT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kFpropThreads, 1)
conv_fprop_tc2_h1_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const FpropParams p) {
  using Cfg = Fprop2H1Cfg<BN, MT>;
  static_assert(Cfg::kTmemCols <= 512, "TMEM has 512 columns");
  constexpr int NA = Cfg::kAStages, NB = Cfg::kBStages, S_ = 3, TAPS = 9, TPS = Cfg::kTPS;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t a_base = base, b_base = base + NA * Cfg::kAStageBytes;
  const uint32_t bar0 = b_base + NB * Cfg::kBStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NA * Cfg::kAStageBytes + NB * Cfg::kBStageBytes);
  auto afull = [&](int s) { return bar0 + 8u * s; };
  auto aempty = [&](int s) { return bar0 + 8u * (NA + s); };
  auto bfull = [&](int s) { return bar0 + 8u * (2 * NA + s); };
  auto bempty = [&](int s) { return bar0 + 8u * (2 * NA + NB + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * NA + 2 * NB + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * NA + 2 * NB + 2 + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NA + 2 * NB + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = (rank == 0);
  const int chunks = p.Ci / kChunk;
  const int pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  // p.num_tiles = units; p.tiles_n = M-tile pairs per N tile (units = tiles_co * m_pairs, N tile slowest)
  const int m_pairs = p.num_tiles / p.tiles_co;
  const H1Range rg = h1_range(p.num_tiles, pair_id, num_pairs);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NA; ++s) { mbar_init(afull(s), 1); mbar_init(aempty(s), 1); }
    for (int s = 0; s < NB; ++s) { mbar_init(bfull(s), 1); mbar_init(bempty(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 16); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc2(smem_u32(tmem_slot), Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();          // (the cluster barrier below orders the CTA as well; compute-sanitizer racecheck only models this one)
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // this CTA's tile of unit u: M tiles ordered (n, th, tw), tw fastest; 16 x 8 pixels each
  auto unit_tile = [&](int u, int& w0, int& h0, int& n0) {
    int t = (u % m_pairs) * 2 + (int)rank;
    const int tw = t % p.tiles_w; t /= p.tiles_w;
    const int th = t % p.tiles_h; t /= p.tiles_h;
    w0 = tw * 8; h0 = th * 16; n0 = t;
  };
  // units of the item that starts at u: MT when the next unit exists in this pair's range and shares the N tile
  auto item_units = [&](int u) { return (MT > 1 && u + 1 < rg.end && (u + 1) / m_pairs == u / m_pairs) ? 2 : 1; };

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      for (int u = rg.begin; u < rg.end;) {
        const int mt = item_units(u);
        int w0[MT], h0[MT], n0[MT];
#pragma unroll
        for (int j = 0; j < MT; ++j)
          if (j < mt) unit_tile(u + j, w0[j], h0[j], n0[j]);
        const int co0 = (u / m_pairs) * BN + (int)rank * (BN / 2);   // this CTA's half of the weight tile
        for (int ch = 0; ch < chunks; ++ch) {
          mbar_wait(aempty(sa), pa ^ 1);
          if (leader) mbar_expect_tx(afull(sa), 2 * mt * Cfg::kTileTx);
          const uint32_t lafull = mapa(afull(sa), 0);
#pragma unroll
          for (int j = 0; j < MT; ++j)
            if (j < mt)
              tma2_load_4d(a_base + sa * Cfg::kAStageBytes + j * Cfg::kTileBytes, &tmA, lafull, ch * kChunk, w0[j] - p.pad,
                           h0[j] - p.pad, n0[j]);
          if (++sa == NA) { sa = 0; pa ^= 1; }
#pragma unroll
          for (int g = 0; g < TAPS / TPS; ++g) {
            mbar_wait(bempty(sb), pb ^ 1);
            if (leader) mbar_expect_tx(bfull(sb), 2 * Cfg::kBStageBytes);
            const uint32_t lbfull = mapa(bfull(sb), 0);
#pragma unroll
            for (int t = 0; t < TPS; ++t)
              tma2_load_3d(b_base + sb * Cfg::kBStageBytes + t * Cfg::kTapBytes, &tmB, lbfull, ch * kChunk, g * TPS + t, co0);
            if (++sb == NB) { sb = 0; pb ^= 1; }
          }
        }
        u += mt;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA, one thread) =====================
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(2 * kBM, BN, 0, 0);
      int sa = 0, sb = 0, as = 0;
      uint32_t pa = 0, pb = 0, aphase = 0;
      for (int u = rg.begin; u < rg.end;) {
        const int mt = item_units(u);
        mbar_wait(tempty_bar(as), aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * (MT * BN);
        for (int ch = 0; ch < chunks; ++ch) {
          mbar_wait(afull(sa), pa);
          tc_fence_after();
          const uint64_t ad0 = make_smem_desc(a_base + sa * Cfg::kAStageBytes, 16, Cfg::kRowPitch);
#pragma unroll
          for (int g = 0; g < TAPS / TPS; ++g) {
            mbar_wait(bfull(sb), pb);
            tc_fence_after();
            const uint64_t bd0 = make_smem_desc(b_base + sb * Cfg::kBStageBytes, 16, 1024);
#pragma unroll
            for (int t = 0; t < TPS; ++t) {
              const int tap = g * TPS + t, r = tap / S_, s = tap - r * S_;
              const int a_off = (r * Cfg::kBoxCols + s) * 128;
#pragma unroll
              for (int j = 0; j < MT; ++j) {
                if (j < mt) {
#pragma unroll
                  for (int kk = 0; kk < 4; ++kk)
                    mma2_tf32(d_tmem + j * BN, ad0 + (uint64_t)((j * Cfg::kTileBytes + a_off + kk * 32) >> 4),
                              bd0 + (uint64_t)((t * Cfg::kTapBytes + kk * 32) >> 4), idesc, (ch | tap | kk) ? 1u : 0u);
                }
              }
            }
            mma2_commit_mc(bempty(sb), 3);
            if (++sb == NB) { sb = 0; pb ^= 1; }
          }
          mma2_commit_mc(aempty(sa), 3);
          if (++sa == NA) { sa = 0; pa ^= 1; }
        }
        mma2_commit_mc(tfull_bar(as), 3);
        if (++as == 2) { as = 0; aphase ^= 1; }
        u += mt;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs): own 128 rows of each of the pair's accumulators =====================
    const int q = warp & 3;
    const int c_begin = ((warp - 4) >> 2) * (BN / 2);
    const int row = q * 32 + lane;
    const int rw = row & 7, rh = row >> 3;
    int as = 0;
    uint32_t aphase = 0;
    for (int u = rg.begin; u < rg.end;) {
      const int mt = item_units(u);
      const int co0 = (u / m_pairs) * BN;
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
#pragma unroll 1
      for (int j = 0; j < mt; ++j) {
        int w0, h0, n;
        unit_tile(u + j, w0, h0, n);
        const int w = w0 + rw, h = h0 + rh;
        const bool valid = (w < p.Wo) && (h < p.Ho);
        float* out = p.y + (((int64_t)n * p.Ho + h) * p.Wo + w) * p.Co + co0;
#pragma unroll 1
        for (int c = c_begin; c < c_begin + BN / 2; c += 32) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + as * (MT * BN) + j * BN + c, v);
          tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int e4 = 0; e4 < 32; e4 += 4) {
              float4 o;
              float* oo = reinterpret_cast<float*>(&o);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float a = p.alpha * __uint_as_float(v[e4 + e]);
                if (p.bias) a += p.bias_scale * __ldg(p.bias + co0 + c + e4 + e);
                oo[e] = act_apply(a, p.act, p.slope);
              }
              *reinterpret_cast<float4*>(out + c + e4) = o;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa(tempty_bar(as), 0));
      if (++as == 2) { as = 0; aphase ^= 1; }
      u += mt;
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, Cfg::kTmemCols);
  }
}

template <int BN, int MT>
int launch_fprop2_h1(const CUtensorMap& tmA, const CUtensorMap& tmB, const FpropParams& p, cudaStream_t st) {
  using Cfg = Fprop2H1Cfg<BN, MT>;
  static bool configured = false;
  if (!configured) {
    GLB_CUDA(cudaFuncSetAttribute(conv_fprop_tc2_h1_kernel<BN, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured = true;
  }
  const int items = (p.num_tiles + MT - 1) / MT;
  const int pairs = items < kNumSMs / 2 ? items : kNumSMs / 2;
  conv_fprop_tc2_h1_kernel<BN, MT><<<2 * pairs, kFpropThreads, Cfg::kSmemBytes, st>>>(tmA, tmB, p);
  GLB_CHECK_LAUNCH("conv_fprop_tc2_h1_kernel");
  return GLB_OK;
}

template <int BN, bool BF = false>
int launch_fprop2_halo(const CUtensorMap& tmA, const CUtensorMap& tmB, const FpropParams& p, cudaStream_t st) {
  using Cfg = Fprop2HaloCfg<BN>;
  static bool configured = false;
  if (!configured) {
    GLB_CUDA(cudaFuncSetAttribute(conv_fprop_tc2_halo_kernel<BN, BF>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured = true;
  }
  const int pairs = p.num_tiles < kNumSMs / 2 ? p.num_tiles : kNumSMs / 2;
  conv_fprop_tc2_halo_kernel<BN, BF><<<2 * pairs, kFpropThreads, Cfg::kSmemBytes, st>>>(tmA, tmB, p);
  GLB_CHECK_LAUNCH("conv_fprop_tc2_halo_kernel");
  return GLB_OK;
}

template <int BN, bool BF, int UP>
int launch_fprop2_fold(const TMapSet& tmA, const CUtensorMap& tmB, const FpropParams& p, cudaStream_t st) {
  using Cfg = Fprop2FoldCfg<BN>;
  static bool configured = false;
  if (!configured) {
    GLB_CUDA(cudaFuncSetAttribute(conv_fprop_tc2_fold_kernel<BN, BF, UP>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured = true;
  }
  const int pairs = p.num_tiles < kNumSMs / 2 ? p.num_tiles : kNumSMs / 2;
  conv_fprop_tc2_fold_kernel<BN, BF, UP><<<2 * pairs, kFpropThreads, Cfg::kSmemBytes, st>>>(tmA, tmB, p);
  GLB_CHECK_LAUNCH("conv_fprop_tc2_fold_kernel");
  return GLB_OK;
}

template <int BN, bool BF = false>
int launch_fprop2(const TMapSet& tmA, const CUtensorMap& tmB, const FpropParams& p, cudaStream_t st) {
  using Cfg = Fprop2Cfg<BN>;
  static bool configured = false;
  if (!configured) {
    GLB_CUDA(cudaFuncSetAttribute(conv_fprop_tc2_kernel<BN, BF>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured = true;
  }
  const int pairs = p.num_tiles < kNumSMs / 2 ? p.num_tiles : kNumSMs / 2;
  conv_fprop_tc2_kernel<BN, BF><<<2 * pairs, kFpropThreads, Cfg::kSmemBytes, st>>>(tmA, tmB, p);
  GLB_CHECK_LAUNCH("conv_fprop_tc2_kernel");
  return GLB_OK;
}

inline int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

template <int BN, bool ONEBOX>
int launch_fprop_halo(const CUtensorMap& tmA, const CUtensorMap& tmB, const FpropParams& p, cudaStream_t st) {
  using Cfg = FpropHaloCfg<BN, 3, 3, ONEBOX>;
  static bool configured = false;
  if (!configured) {
    GLB_CUDA(cudaFuncSetAttribute(conv_fprop_tc_halo_kernel<BN, 3, 3, ONEBOX>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured = true;
  }
  const int grid = p.num_tiles < kNumSMs ? p.num_tiles : kNumSMs;
  conv_fprop_tc_halo_kernel<BN, 3, 3, ONEBOX><<<grid, kFpropThreads, Cfg::kSmemBytes, st>>>(tmA, tmB, p);
  GLB_CHECK_LAUNCH("conv_fprop_tc_halo_kernel");
  return GLB_OK;
}

template <int BN>
int launch_fprop_m2(const CUtensorMap& tmA, const CUtensorMap& tmB, const FpropParams& p, cudaStream_t st) {
  using Cfg = FpropM2Cfg<BN>;
  static bool configured = false;
  if (!configured) {
    GLB_CUDA(cudaFuncSetAttribute(conv_fprop_tc_m2_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured = true;
  }
  const int grid = p.num_tiles < kNumSMs ? p.num_tiles : kNumSMs;
  conv_fprop_tc_m2_kernel<BN><<<grid, kFpropThreads, Cfg::kSmemBytes, st>>>(tmA, tmB, p);
  GLB_CHECK_LAUNCH("conv_fprop_tc_m2_kernel");
  return GLB_OK;
}

template <int BN, int KC, bool BF = false>
int launch_fprop(const TMapSet& tmA, const CUtensorMap& tmB, const FpropParams& p, cudaStream_t st) {
  using Cfg = FpropCfg<BN, KC>;
  static bool configured = false;  // per-process, per-instantiation; attribute is sticky for the function
  if (!configured) {
    GLB_CUDA(cudaFuncSetAttribute(conv_fprop_tc_kernel<BN, KC, BF>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured = true;
  }
  const int grid = p.num_tiles < kNumSMs ? p.num_tiles : kNumSMs;
  conv_fprop_tc_kernel<BN, KC, BF><<<grid, kFpropThreads, Cfg::kSmemBytes, st>>>(tmA, tmB, p);
  GLB_CHECK_LAUNCH("conv_fprop_tc_kernel");
  return GLB_OK;
}

// Low-resolution layers: a handful of output tiles behind a long K loop is latency bound (one CTA streams K at ~0.3 us per
// 32-channel step).  Pick the N tile and a split of K over CTAs (partials reduced with red.add) that minimise a simple cost
// model: waves x (k steps x 0.3 us + epilogue).  Leaves (BN, ksplit) alone when the layer has at least half a wave of tiles.
void pick_tile_and_split(int m_tiles, int Co, int k_iters, int& BN, int& ksplit) {
  if (m_tiles * (Co / BN) >= kNumSMs / 2) return;
  double best = 1e30;
  for (int bn = 256; bn >= 32; bn >>= 1) {
    if (Co % bn != 0) continue;
    const int tiles = m_tiles * (Co / bn);
    for (int sp = 1; sp <= 16; ++sp) {
      const int kper = (k_iters + sp - 1) / sp;
      if (sp > 1 && kper < 8) break;
      const int items = tiles * ((k_iters + kper - 1) / kper);
      const int waves = (items + kNumSMs - 1) / kNumSMs;
      const double cost = waves * (kper * (0.28 + 0.0004 * bn) + (sp > 1 ? 0.06 : 0.012) * bn + 3.0);
      if (cost < best) { best = cost; BN = bn; ksplit = sp; }
    }
  }
}

}  // namespace

// Shapes the tensor-core fprop covers: stride-1 RxS with `pad`, Ci % 32 == 0, Co in {16, 32, 64} or a multiple of 128.
// (bf16 operands: 64 channels per 128-byte row -> Ci % 64 == 0)
static bool fprop_covers(int N, int H, int W, int Ci, int Co, int R, int S, int pad, int chunk) {
  if (Ci % chunk != 0 || Ci < chunk) return false;
  if (!(Co == 16 || Co == 32 || Co == 64 || Co % 128 == 0)) return false;
  const int Ho = H + 2 * pad - R + 1, Wo = W + 2 * pad - S + 1;
  if (Ho <= 0 || Wo <= 0) return false;
  if (R * S > 25) return false;
  return true;
}
bool conv_fprop_tc_covers(int N, int H, int W, int Ci, int Co, int R, int S, int pad) {
  return fprop_covers(N, H, W, Ci, Co, R, S, pad, 32);
}
bool conv_fprop_bf16_covers(int N, int H, int W, int Ci, int Co, int R, int S, int pad) {
  return fprop_covers(N, H, W, Ci, Co, R, S, pad, 64);
}

// BF = false: x, w are fp32 (kind::tf32 truncates them); BF = true: x, w are bf16 copies (round-to-nearest, glb_cvt_f32_bf16)
template <bool BF>
static int conv_fprop_tc_impl(const void* x, const void* w, const float* bias, float* y, int N, int H, int W, int Ci, int Co, int R,
                              int S, int pad, float alpha, float bias_scale, int act, float slope, cudaStream_t st) {
  constexpr int kChunk = BF ? 64 : 32;             // channels per 128-byte operand row
  constexpr uint64_t ES = BF ? 2 : 4;              // operand element size
  if (!fprop_covers(N, H, W, Ci, Co, R, S, pad, kChunk)) {
    set_error("tcgen05 fprop: shape not covered (need Ci % 32 == 0 (bf16: 64) and Co in {16,32,64} or Co % 128 == 0)");
    return GLB_ERR_UNSUPPORTED;
  }
  FpropParams p;
  p.y = y; p.bias = bias;
  p.N = N; p.Ho = H + 2 * pad - R + 1; p.Wo = W + 2 * pad - S + 1; p.Co = Co;
  p.Ci = Ci; p.R = R; p.S = S; p.pad = pad;
  p.bw = next_pow2(p.Wo) < 16 ? next_pow2(p.Wo) : 16;
  p.bh = next_pow2(p.Ho) < kBM / p.bw ? next_pow2(p.Ho) : kBM / p.bw;
  p.bn = kBM / (p.bw * p.bh);
  p.tiles_w = (p.Wo + p.bw - 1) / p.bw;
  p.tiles_h = (p.Ho + p.bh - 1) / p.bh;
  p.tiles_n = (N + p.bn - 1) / p.bn;
  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  const int k_iters = R * S * (Ci / kChunk);
  int BN = Co % 256 == 0 ? 256 : (Co % 128 == 0 ? 128 : Co);
  int ksplit = 1;
  pick_tile_and_split(m_tiles, Co, k_iters, BN, ksplit);
  p.k_per = (k_iters + ksplit - 1) / ksplit;
  if (ksplit > 1 && (p.k_per & 1)) ++p.k_per;   // whole 2-chunk stages per slice
  p.ksplit = (k_iters + p.k_per - 1) / p.k_per;
  p.tiles_co = Co / BN;
  p.num_tiles = m_tiles * p.tiles_co * p.ksplit;
  p.alpha = BF ? alpha : alpha * kTf32TruncComp; p.bias_scale = bias_scale; p.act = act; p.slope = slope;
  p.up = 0; p.OH = p.Ho; p.OW = p.Wo;
  p.dbg = 0;
  p.trace = nullptr;
  if (const char* e = getenv("GLB_FPROP_DBG")) p.dbg = atoi(e);
  if (const char* e = getenv("GLB_FPROP_TRACE")) p.trace = (long long*)strtoull(e, nullptr, 0);
  const bool post_pass = p.ksplit > 1 && (bias != nullptr || act != GLB_ACT_NONE);
  if (p.ksplit > 1) GLB_CUDA(cudaMemsetAsync(y, 0, sizeof(float) * (size_t)N * p.Ho * p.Wo * Co, st));

  // CTA pair + one-box tap reuse + two M tiles per B tile (conv_fprop_tc2_h1_kernel): 3x3 "same" convolutions on maps that tile
  // into 16 x 8 pixel tiles, an even number of them, Co a multiple of 128, at least one unit pair per CTA pair
  {
    int h1 = 0;                                   // GLB_FPROP_H1: 0 = off, 1 = <128, 2> (default once measured), 2 = <256, 1>
    if (const char* e = getenv("GLB_FPROP_H1")) h1 = atoi(e); else h1 = kH1Default;
    if (BF) h1 = 0;
    const int bnh = (h1 == 2) ? 256 : 128;
    const int m_h = (p.Ho / 16) * (p.Wo / 8) * N;
    const bool ok = h1 != 0 && p.ksplit == 1 && R == 3 && S == 3 && pad == 1 && p.Ho % 16 == 0 && p.Wo % 8 == 0 && Co % bnh == 0 &&
                    m_h % 2 == 0 && (m_h / 2) * (Co / bnh) >= kNumSMs;
    if (ok) {
      FpropParams q = p;
      q.bw = 8; q.bh = 16; q.bn = 1;
      q.tiles_w = p.Wo / 8; q.tiles_h = p.Ho / 16; q.tiles_n = N; q.tiles_co = Co / bnh;
      q.num_tiles = (m_h / 2) * q.tiles_co;          // units
      CUtensorMap hA, hB;
      const uint64_t dimsA[4] = {(uint64_t)Ci, (uint64_t)W, (uint64_t)H, (uint64_t)N};
      const uint64_t stridesA[3] = {(uint64_t)Ci * ES, (uint64_t)W * Ci * ES, (uint64_t)H * W * Ci * ES};
      const uint32_t boxA[4] = {(uint32_t)kChunk, 10u, 18u, 1u};
      int rc = make_tmap(&hA, x, 4, dimsA, stridesA, boxA, "conv input (10 x 18 box)", false, BF);
      if (rc) return rc;
      const uint64_t dimsB[3] = {(uint64_t)Ci, (uint64_t)(R * S), (uint64_t)Co};
      const uint64_t stridesB[2] = {(uint64_t)Ci * ES, (uint64_t)R * S * Ci * ES};
      const uint32_t boxB[3] = {(uint32_t)kChunk, 1u, (uint32_t)(bnh / 2)};
      rc = make_tmap(&hB, w, 3, dimsB, stridesB, boxB, "conv weight (half tile)", false, BF);
      if (rc) return rc;
      if (!BF) return bnh == 128 ? launch_fprop2_h1<128, 2>(hA, hB, q, st) : launch_fprop2_h1<256, 1>(hA, hB, q, st);
    }
  }

  // tap-reuse kernel: 3x3 "same" convolutions on maps that tile exactly into 16 x 8 pixel tiles, >= one wave of tiles
  {
    int halo = 0;   // GLB_FPROP_HALO: bit 0 = use for BN 128, bit 1 = use for BN 256 (default: see below)
    if (const char* e = getenv("GLB_FPROP_HALO")) halo = atoi(e); else halo = kHaloDefault;
    const bool shape_ok = p.ksplit == 1 && R == 3 && S == 3 && pad == 1 && p.Ho % 16 == 0 && p.Wo % 8 == 0 &&
                          (BN == 128 || BN == 256);
    const int h_tiles = (p.Ho / 16) * (p.Wo / 8) * N * (Co / BN);
    if (!BF && shape_ok && h_tiles >= kNumSMs && ((BN == 128 && (halo & 1)) || (BN == 256 && (halo & 2)))) {
      FpropParams q = p;
      q.bw = 8; q.bh = 16; q.bn = 1;
      q.tiles_w = p.Wo / 8; q.tiles_h = p.Ho / 16; q.tiles_n = N; q.tiles_co = Co / BN;
      q.num_tiles = h_tiles;
      CUtensorMap hA, hB;
      const uint64_t dimsA[4] = {(uint64_t)Ci, (uint64_t)W, (uint64_t)H, (uint64_t)N};
      const uint64_t stridesA[3] = {(uint64_t)Ci * ES, (uint64_t)W * Ci * ES, (uint64_t)H * W * Ci * ES};
      const bool onebox = getenv("GLB_FPROP_ONEBOX") != nullptr && atoi(getenv("GLB_FPROP_ONEBOX")) != 0;   // A/B experiment
      const uint32_t boxA[4] = {(uint32_t)kChunk, onebox ? 10u : 8u, 18u, 1u};
      int rc = make_tmap(&hA, x, 4, dimsA, stridesA, boxA, "conv input (halo boxes)", false, BF);
      if (rc) return rc;
      const uint64_t dimsB[3] = {(uint64_t)Ci, (uint64_t)(R * S), (uint64_t)Co};
      const uint64_t stridesB[2] = {(uint64_t)Ci * ES, (uint64_t)R * S * Ci * ES};
      const uint32_t boxB[3] = {(uint32_t)kChunk, 1u, (uint32_t)BN};
      rc = make_tmap(&hB, w, 3, dimsB, stridesB, boxB, "conv weight", false, BF);
      if (rc) return rc;
      if (onebox) return BN == 128 ? launch_fprop_halo<128, true>(hA, hB, q, st) : launch_fprop_halo<256, true>(hA, hB, q, st);
      return BN == 128 ? launch_fprop_halo<128, false>(hA, hB, q, st) : launch_fprop_halo<256, false>(hA, hB, q, st);
    }
    // bits 2 / 3: CTA pair + tap reuse for BN 128 / 256 (M = 256 per pair: needs an even number of 16 x 8 tiles)
    const int m_h = (p.Ho / 16) * (p.Wo / 8) * N;
    if (shape_ok && m_h % 2 == 0 && (m_h / 2) * (Co / BN) >= kNumSMs / 2 &&
        ((BN == 128 && (halo & 4)) || (BN == 256 && (halo & 8)))) {
      FpropParams q = p;
      q.bw = 8; q.bh = 16; q.bn = 1;
      q.tiles_w = p.Wo / 8; q.tiles_h = p.Ho / 16; q.tiles_n = N; q.tiles_co = Co / BN;
      q.num_tiles = (m_h / 2) * q.tiles_co;          // pair items
      CUtensorMap hA, hB;
      const uint64_t dimsA[4] = {(uint64_t)Ci, (uint64_t)W, (uint64_t)H, (uint64_t)N};
      const uint64_t stridesA[3] = {(uint64_t)Ci * ES, (uint64_t)W * Ci * ES, (uint64_t)H * W * Ci * ES};
      const uint32_t boxA[4] = {(uint32_t)kChunk, 8u, 18u, 1u};
      int rc = make_tmap(&hA, x, 4, dimsA, stridesA, boxA, "conv input (halo boxes)", false, BF);
      if (rc) return rc;
      const uint64_t dimsB[3] = {(uint64_t)Ci, (uint64_t)(R * S), (uint64_t)Co};
      const uint64_t stridesB[2] = {(uint64_t)Ci * ES, (uint64_t)R * S * Ci * ES};
      const uint32_t boxB[3] = {(uint32_t)kChunk, 1u, (uint32_t)(BN / 2)};
      rc = make_tmap(&hB, w, 3, dimsB, stridesB, boxB, "conv weight (half tile)", false, BF);
      if (rc) return rc;
      return BN == 128 ? launch_fprop2_halo<128, BF>(hA, hB, q, st) : launch_fprop2_halo<256, BF>(hA, hB, q, st);
    }
  }

  // CTA pairs (cta_group::2) for the layers with at least one wave of pair tiles
  // (measured: +10 % at BN = 256, no gain at BN = 128, where the loop is bound by L2 -> SM traffic rather than shared memory)
  bool use_pair = p.ksplit == 1 && BN == 256 && m_tiles % 2 == 0 && (m_tiles / 2) * p.tiles_co >= kNumSMs / 4;
  if (const char* e = getenv("GLB_FPROP_PAIR")) use_pair = use_pair && atoi(e) != 0;  // tuning experiments only
  if (use_pair) p.num_tiles = (m_tiles / 2) * p.tiles_co;   // pair items

  TMapSet tmA;
  CUtensorMap tmB;
  {
    const uint64_t dims[4] = {(uint64_t)Ci, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    const uint64_t strides[3] = {(uint64_t)Ci * ES, (uint64_t)W * Ci * ES, (uint64_t)H * W * Ci * ES};
    const uint32_t box[4] = {(uint32_t)kChunk, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn};
    int rc = make_tmap(&tmA.m[0], x, 4, dims, strides, box, "conv input", false, BF);
    if (rc) return rc;
    tmA.m[1] = tmA.m[2] = tmA.m[3] = tmA.m[0];
  }
  {
    const uint64_t dims[3] = {(uint64_t)Ci, (uint64_t)(R * S), (uint64_t)Co};
    const uint64_t strides[2] = {(uint64_t)Ci * ES, (uint64_t)R * S * Ci * ES};
    const uint32_t box[3] = {(uint32_t)kChunk, 1u, (uint32_t)(use_pair ? BN / 2 : BN)};
    int rc = make_tmap(&tmB, w, 3, dims, strides, box, "conv weight", false, BF);
    if (rc) return rc;
  }
  if (use_pair) return BN == 256 ? launch_fprop2<256, BF>(tmA, tmB, p, st) : launch_fprop2<128, BF>(tmA, tmB, p, st);
  // two M tiles per item for the 128-wide N tiles of the high-resolution layers (>= one wave of pair items)
  bool use_m2 = !BF && BN == 128 && p.ksplit == 1 && m_tiles % 2 == 0 && (m_tiles / 2) * p.tiles_co >= kNumSMs;
  if (const char* e = getenv("GLB_FPROP_M2")) use_m2 = use_m2 && atoi(e) != 0;   // tuning experiments only
  if (use_m2) {
    p.num_tiles = (m_tiles / 2) * p.tiles_co;
    return launch_fprop_m2<128>(tmA.m[0], tmB, p, st);
  }
  int rc = GLB_ERR_UNSUPPORTED;
  // two K chunks per stage for the narrow tiles (see FpropCfg); needs whole stages per tap row and per split-K slice
  const bool kc2 = BN <= 128 && (Ci / kChunk) % 2 == 0 && p.k_per % 2 == 0 && getenv("GLB_FPROP_KC1") == nullptr;
  switch (BN) {
    case 256: rc = launch_fprop<256, 1, BF>(tmA, tmB, p, st); break;
    case 128: rc = kc2 ? launch_fprop<128, 2, BF>(tmA, tmB, p, st) : launch_fprop<128, 1, BF>(tmA, tmB, p, st); break;
    case 64: rc = kc2 ? launch_fprop<64, 2, BF>(tmA, tmB, p, st) : launch_fprop<64, 1, BF>(tmA, tmB, p, st); break;
    case 32: rc = kc2 ? launch_fprop<32, 2, BF>(tmA, tmB, p, st) : launch_fprop<32, 1, BF>(tmA, tmB, p, st); break;
    case 16: rc = launch_fprop<16, 1, BF>(tmA, tmB, p, st); break;
    default: set_error("tcgen05 fprop: no kernel for this N tile");
  }
  if (rc == GLB_OK && post_pass)
    rc = glb_bias_act_fwd(y, bias, y, (int64_t)N * p.Ho * p.Wo, Co, bias_scale, act, slope, (glb_stream_t)st);
  return rc;
}

int conv_fprop_tc(const float* x, const float* w, const float* bias, float* y, int N, int H, int W, int Ci, int Co, int R, int S,
                  int pad, float alpha, float bias_scale, int act, float slope, cudaStream_t st) {
  return conv_fprop_tc_impl<false>(x, w, bias, y, N, H, W, Ci, Co, R, S, pad, alpha, bias_scale, act, slope, st);
}

// bf16 operands (x [N,H,W,Ci] and w [Co,R,S,Ci] as bf16), fp32 accumulate / bias / activation / output
int conv_fprop_bf16(const void* x, const void* w, const float* bias, float* y, int N, int H, int W, int Ci, int Co, int R, int S,
                    int pad, float alpha, float bias_scale, int act, float slope, cudaStream_t st) {
  return conv_fprop_tc_impl<true>(x, w, bias, y, N, H, W, Ci, Co, R, S, pad, alpha, bias_scale, act, slope, st);
}

int conv_dgrad_bf16(const void* gy, const void* wt, float* gx, int N, int H, int W, int Ci, int Co, int R, int S, int pad,
                    float alpha, cudaStream_t st) {
  if (wt == nullptr || R != S) {
    set_error("bf16 dgrad needs the transposed bf16 weights and a square filter");
    return GLB_ERR_UNSUPPORTED;
  }
  const int Ho = H + 2 * pad - R + 1, Wo = W + 2 * pad - S + 1;
  return conv_fprop_tc_impl<true>(gy, wt, nullptr, gx, N, Ho, Wo, Co, Ci, R, S, R - 1 - pad, alpha, 0.f, GLB_ACT_NONE, 0.f, st);
}

bool conv_dgrad_bf16_covers(int N, int H, int W, int Ci, int Co, int R, int S, int pad) {
  if (R != S || R - 1 - pad < 0) return false;
  const int Ho = H + 2 * pad - R + 1, Wo = W + 2 * pad - S + 1;
  return conv_fprop_bf16_covers(N, Ho, Wo, Co, Ci, R, S, R - 1 - pad);
}

// dgrad = fprop over gy with the flipped / transposed weights wt[Ci][R][S][Co] (glb_conv2d_weight_transpose):
//   gx[n,h,w,ci] = alpha * sum_{r',s',co} gy[n, h + r' - (R-1-pad), w + s' - (S-1-pad), co] * wt[ci][r'][s'][co]
int conv_dgrad_tc(const float* gy, const float* wt, float* gx, int N, int H, int W, int Ci, int Co, int R, int S, int pad,
                  float alpha, cudaStream_t st) {
  if (wt == nullptr) {
    set_error("tcgen05 dgrad needs the transposed weights (glb_conv2d_weight_transpose) in `wt`");
    return GLB_ERR_UNSUPPORTED;
  }
  if (R != S) {
    set_error("tcgen05 dgrad: square filters only");
    return GLB_ERR_UNSUPPORTED;
  }
  const int Ho = H + 2 * pad - R + 1, Wo = W + 2 * pad - S + 1;
  // input of this "fprop" is gy [N,Ho,Wo,Co]; output is gx [N,H,W,Ci]; H = Ho + 2*pad' - R + 1 with pad' = R-1-pad
  return conv_fprop_tc(gy, wt, nullptr, gx, N, Ho, Wo, Co, Ci, R, S, R - 1 - pad, alpha, 0.f, GLB_ACT_NONE, 0.f, st);
}

bool conv_dgrad_tc_covers(int N, int H, int W, int Ci, int Co, int R, int S, int pad) {
  if (R != S || R - 1 - pad < 0) return false;
  const int Ho = H + 2 * pad - R + 1, Wo = W + 2 * pad - S + 1;
  return conv_fprop_tc_covers(N, Ho, Wo, Co, Ci, R, S, R - 1 - pad);
}

// ------------------------------------------------------------------------------------------------ upsample folded into the conv
// y = conv3x3_same(upsample2x_nearest(x), w)  (reference stylegan/architectures.py:155-156 / progan/architectures.py: the
// `nn.Upsample` in front of a Conv2dEx) WITHOUT the upsampled copy and with 4/9 of the multiply-adds: output phase (dy, dx),
//   y[n, 2i+dy, 2j+dx, co] = sum_{a,b in {0,1}} sum_ci x[n, i + a - (1-dy), j + b - (1-dx), ci] * wp[dy*2+dx][co][a][b][ci],
// where along each axis  d = 0: tap a=0 <- w[0], a=1 <- w[1] + w[2];   d = 1: a=0 <- w[0] + w[1], a=1 <- w[2]
// (rows 2i-1 | 2i, 2i+1 of the upsampled map are rows i-1 | i, i of x, and so on; the zero padding of the upsampled map is
// exactly the out-of-bounds zero fill on x).  glb_upconv_weights writes wp and, for the data gradient, wt[ci][ph*4+a'*2+b'][co] =
// wp[ph][co][1-a'][1-b'][ci].
// chunk = channels per 128-byte operand row: 32 (fp32 operands, kind::tf32) or 64 (bf16 operands, kind::f16)
bool conv_upconv_covers(int kind, int N, int H, int W, int Ci, int Co, int chunk) {
  if (N <= 0 || H <= 0 || W <= 0) return false;
  const bool n_ok_co = (Co == 32 || Co == 64 || Co % 128 == 0), n_ok_ci = (Ci == 32 || Ci == 64 || Ci % 128 == 0);
  switch (kind) {
    case 0: return Ci % chunk == 0 && n_ok_co;               // fprop: K = Ci, GEMM N = Co
    case 1: return Co % chunk == 0 && n_ok_ci;               // dgrad: K = Co, GEMM N = Ci
    case 2: return Ci % chunk == 0 && Co % chunk == 0 && (chunk == 32 || Ci % 64 == 0);          // wgrad
  }
  return false;
}
bool conv_upconv_covers(int kind, int N, int H, int W, int Ci, int Co) { return conv_upconv_covers(kind, N, H, W, Ci, Co, 32); }

// shared by the forward (up = 1) and the data gradient (up = 2): `kc` = channels of the reduction, `nc` = channels of the output
template <bool BF>
static int upconv_launch(int up, const void* in, const void* wmat, const float* bias, float* out, int N, int H, int W, int kc, int nc,
                         float alpha, float bias_scale, int act, float slope, cudaStream_t st) {
  constexpr int kChunk = BF ? 64 : 32;              // channels per 128-byte operand row (shadows the fp32 constant)
  constexpr uint64_t ES = BF ? 2 : 4;               // operand element size
  FpropParams p;
  p.y = out; p.bias = bias;
  p.N = N; p.Ho = H; p.Wo = W; p.Co = nc;
  p.Ci = kc; p.R = up == 1 ? 2 : 4; p.S = p.R; p.pad = 0;
  p.up = up; p.OH = up == 1 ? 2 * H : H; p.OW = up == 1 ? 2 * W : W;
  p.bw = next_pow2(W) < 16 ? next_pow2(W) : 16;
  p.bh = next_pow2(H) < kBM / p.bw ? next_pow2(H) : kBM / p.bw;
  p.bn = kBM / (p.bw * p.bh);
  p.tiles_w = (W + p.bw - 1) / p.bw;
  p.tiles_h = (H + p.bh - 1) / p.bh;
  p.tiles_n = (N + p.bn - 1) / p.bn;
  const int m_one = p.tiles_w * p.tiles_h * p.tiles_n;         // M tiles of one phase
  const int m_tiles = up == 1 ? 4 * m_one : m_one;
  const int k_iters = p.R * p.S * (kc / kChunk);
  int BN = nc % 256 == 0 ? 256 : (nc % 128 == 0 ? 128 : nc);
  int ksplit = 1;
  pick_tile_and_split(m_tiles, nc, k_iters, BN, ksplit);
  p.k_per = (k_iters + ksplit - 1) / ksplit;
  if (ksplit > 1 && (p.k_per & 1)) ++p.k_per;
  p.ksplit = (k_iters + p.k_per - 1) / p.k_per;
  p.tiles_co = nc / BN;
  p.num_tiles = m_tiles * p.tiles_co * p.ksplit;
  p.alpha = BF ? alpha : alpha * kTf32TruncComp; p.bias_scale = bias_scale; p.act = act; p.slope = slope;
  p.dbg = 0; p.trace = nullptr;
  const int64_t out_rows = (int64_t)N * p.OH * p.OW;
  const bool post_pass = p.ksplit > 1 && (bias != nullptr || act != GLB_ACT_NONE);
  if (p.ksplit > 1) GLB_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)out_rows * nc, st));
  // CTA pair + tap reuse (conv_fprop_tc2_fold_kernel): maps that tile into 16 x 8 pixel tiles, an even number of them per phase
  const int m_fold = (H / 16) * (W / 8) * N;
  bool use_fold = p.ksplit == 1 && (BN == 128 || BN == 256) && H % 16 == 0 && W % 8 == 0 && m_fold % 2 == 0 &&
                  ((up == 1 ? 4 : 1) * (m_fold / 2)) * (nc / BN) >= kNumSMs / 8;
  if (const char* e = getenv("GLB_FOLD_HALO")) use_fold = use_fold && atoi(e) != 0;      // A/B runs
  if (use_fold) {
    p.bw = 8; p.bh = 16; p.bn = 1;
    p.tiles_w = W / 8; p.tiles_h = H / 16; p.tiles_n = N;
    p.num_tiles = ((up == 1 ? 4 : 1) * (m_fold / 2)) * p.tiles_co;
  }
  const bool use_pair = !use_fold && p.ksplit == 1 && BN == 256 && m_one % 2 == 0 && (m_tiles / 2) * p.tiles_co >= kNumSMs / 4;
  if (use_pair) p.num_tiles = (m_tiles / 2) * p.tiles_co;

  TMapSet tmA;
  CUtensorMap tmB;
  const uint32_t boxA[4] = {(uint32_t)kChunk, (uint32_t)(use_fold ? 8 : p.bw), (uint32_t)(use_fold ? 17 : p.bh), (uint32_t)(use_fold ? 1 : p.bn)};
  if (up == 1) {
    const uint64_t dims[4] = {(uint64_t)kc, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    const uint64_t strides[3] = {(uint64_t)kc * ES, (uint64_t)W * kc * ES, (uint64_t)H * W * kc * ES};
    int rc = make_tmap(&tmA.m[0], in, 4, dims, strides, boxA, "upconv input", false, BF);
    if (rc) return rc;
    tmA.m[1] = tmA.m[2] = tmA.m[3] = tmA.m[0];
  } else {
    // phase (dy, dx) of the high-resolution gradient [N][2H][2W][kc] as a strided [N][H][W][kc] view
    const uint64_t dims[4] = {(uint64_t)kc, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    const uint64_t strides[3] = {(uint64_t)2 * kc * ES, (uint64_t)2 * (2 * W) * kc * ES, (uint64_t)(2 * H) * (2 * W) * kc * ES};
    for (int ph = 0; ph < 4; ++ph) {
      const char* base = (const char*)in + ((int64_t)(ph >> 1) * (2 * W) + (ph & 1)) * kc * ES;
      int rc = make_tmap(&tmA.m[ph], base, 4, dims, strides, boxA, "upconv gradient (phase view)", false, BF);
      if (rc) return rc;
    }
  }
  {
    const int taps = p.R * p.S;
    const uint64_t rows = up == 1 ? (uint64_t)4 * nc : (uint64_t)nc;
    const uint64_t dims[3] = {(uint64_t)kc, (uint64_t)taps, rows};
    const uint64_t strides[2] = {(uint64_t)kc * ES, (uint64_t)taps * kc * ES};
    const uint32_t box[3] = {(uint32_t)kChunk, 1u, (uint32_t)((use_pair || use_fold) ? BN / 2 : BN)};
    int rc = make_tmap(&tmB, wmat, 3, dims, strides, box, "upconv weight", false, BF);
    if (rc) return rc;
  }
  if (use_fold) {
    if (up == 1) return BN == 256 ? launch_fprop2_fold<256, BF, 1>(tmA, tmB, p, st) : launch_fprop2_fold<128, BF, 1>(tmA, tmB, p, st);
    return BN == 256 ? launch_fprop2_fold<256, BF, 2>(tmA, tmB, p, st) : launch_fprop2_fold<128, BF, 2>(tmA, tmB, p, st);
  }
  if (use_pair) return launch_fprop2<256, BF>(tmA, tmB, p, st);
  int rc = GLB_ERR_UNSUPPORTED;
  const bool kc2 = BN <= 128 && (kc / kChunk) % 2 == 0 && p.k_per % 2 == 0;
  switch (BN) {
    case 256: rc = launch_fprop<256, 1, BF>(tmA, tmB, p, st); break;
    case 128: rc = kc2 ? launch_fprop<128, 2, BF>(tmA, tmB, p, st) : launch_fprop<128, 1, BF>(tmA, tmB, p, st); break;
    case 64: rc = kc2 ? launch_fprop<64, 2, BF>(tmA, tmB, p, st) : launch_fprop<64, 1, BF>(tmA, tmB, p, st); break;
    case 32: rc = kc2 ? launch_fprop<32, 2, BF>(tmA, tmB, p, st) : launch_fprop<32, 1, BF>(tmA, tmB, p, st); break;
    default: set_error("upconv: no kernel for this N tile");
  }
  if (rc == GLB_OK && post_pass) rc = glb_bias_act_fwd(out, bias, out, out_rows, nc, bias_scale, act, slope, (glb_stream_t)st);
  return rc;
}

int conv_upconv_fprop_tc(const float* x, const float* wp, const float* bias, float* y, int N, int H, int W, int Ci, int Co, float alpha,
                         float bias_scale, int act, float slope, cudaStream_t st) {
  if (!conv_upconv_covers(0, N, H, W, Ci, Co)) {
    set_error("upconv fprop: shape not covered (Ci % 32 == 0, Co in {32, 64} or a multiple of 128)");
    return GLB_ERR_UNSUPPORTED;
  }
  return upconv_launch<false>(1, x, wp, bias, y, N, H, W, Ci, Co, alpha, bias_scale, act, slope, st);
}
int conv_upconv_fprop_bf16(const void* x, const void* wp, const float* bias, float* y, int N, int H, int W, int Ci, int Co, float alpha,
                           float bias_scale, int act, float slope, cudaStream_t st) {
  if (!conv_upconv_covers(0, N, H, W, Ci, Co, 64)) {
    set_error("upconv fprop (bf16): shape not covered (Ci % 64 == 0, Co in {32, 64} or a multiple of 128)");
    return GLB_ERR_UNSUPPORTED;
  }
  return upconv_launch<true>(1, x, wp, bias, y, N, H, W, Ci, Co, alpha, bias_scale, act, slope, st);
}

// gx[n,i,j,ci] = alpha * sum_{ph,a',b',co} gy[n, 2(i + a' - dy) + dy, 2(j + b' - dx) + dx, co] * wt[ci][ph*4 + a'*2 + b'][co]
int conv_upconv_dgrad_tc(const float* gy, const float* wt, float* gx, int N, int H, int W, int Ci, int Co, float alpha, cudaStream_t st) {
  if (!conv_upconv_covers(1, N, H, W, Ci, Co)) {
    set_error("upconv dgrad: shape not covered (Co % 32 == 0, Ci in {32, 64} or a multiple of 128)");
    return GLB_ERR_UNSUPPORTED;
  }
  return upconv_launch<false>(2, gy, wt, nullptr, gx, N, H, W, Co, Ci, alpha, 0.f, GLB_ACT_NONE, 0.f, st);
}
int conv_upconv_dgrad_bf16(const void* gy, const void* wt, float* gx, int N, int H, int W, int Ci, int Co, float alpha, cudaStream_t st) {
  if (!conv_upconv_covers(1, N, H, W, Ci, Co, 64)) {
    set_error("upconv dgrad (bf16): shape not covered (Co % 64 == 0, Ci in {32, 64} or a multiple of 128)");
    return GLB_ERR_UNSUPPORTED;
  }
  return upconv_launch<true>(2, gy, wt, nullptr, gx, N, H, W, Co, Ci, alpha, 0.f, GLB_ACT_NONE, 0.f, st);
}

namespace {
using namespace tc;
// ------------------------------------------------------------------------------------------------ wgrad
// gw[co][tap][ci] = alpha * sum_{p = (n,ho,wo)} gy[p][co] * x[p shifted by (r-pad, s-pad)][ci]
//   GEMM  D[m = co][n = ci] over K = output pixels, one accumulator per (128 co, BN ci, tap) tile, split-K over pixel
//   blocks across CTAs (partials reduced with red.global.add.v4.f32 into the zero-initialised gw).
//   Both operands have the channel axis contiguous and K (pixels) strided -> "MN-major" tcgen05 operands: a TMA box
//   (32 channels, 32 pixels) is one [32 px][128 B] block of the canonical MN-major layout "128B swizzle with 32-byte
//   atoms" -- the only one tcgen05 accepts for MN-major 32-bit operands (TMA mode SWIZZLE_128B_ATOM_32B; 4-pixel swizzle
//   period) -- with LBO = 4 KB between 32-channel blocks and SBO = 512 B between groups of 4 pixels.  Channel blocks beyond Co are
//   zero-filled by TMA (M padded to 128), padding taps likewise.
struct WgradParams {
  float* gw;
  int Co, Ci, RS, S, pad;
  int N, Ho, Wo;
  int bw, bh, bn;  // K block = PIX pixels = bn x bh x bw box
  int tiles_w, tiles_h, tiles_n, num_pb;
  int splits, pb_per_split;
  int tiles_co, tiles_ci;
  float alpha;
  int atomic;
  int up;          // 1: weight gradient of the upsample-folded convolution (see FpropParams::up): 16 "taps" = phase*4 + a*2 + b,
                   //    gy comes from the strided view of its phase, (Ho, Wo) = the low-resolution grid, pad = 1
};

// PIX = pixels (K) per pipeline stage: 32 (4 MMAs) for BN = 256, 64 (8 MMAs) for the narrower tiles, whose 64-cycle MMAs would
// otherwise leave the issuing thread's per-stage wait / commit protocol as the pacing item (see FpropCfg).
// The tensor maps are 5-D -- (32 ch, W, H, N, C/32) -- so ONE TMA instruction per operand and stage lands all its
// [channel block][pixel][32 ch] blocks (12 separate 4 KB loads per stage held the main loop at ~35 % tensor-pipe activity).
template <int BN, int PIX, bool BF = false>
struct WgradCfg {
  static constexpr int kCH = BF ? 64 : 32;              // channels per 128-byte row
  static constexpr int kBlkBytes = PIX * 128;           // one (kCH ch x PIX px) block
  static constexpr int kABytes = (128 / kCH) * kBlkBytes;
  static constexpr int kBBytes = (BN / kCH) * kBlkBytes;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (192 * 1024 / kStageBytes) > 8 ? 8 : (192 * 1024 / kStageBytes);
  static constexpr int kTmemCols = BN < 32 ? 32 : BN;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
};

// BF: bf16 operands.  16-bit MN-major operands use the PLAIN 128-byte swizzle: one atom = 8 pixels x 64 channels (1024 B),
// SBO = 1024 B between 8-pixel groups, LBO = one (64 ch x PIX px) block between 64-channel blocks, K = 16 pixels per MMA.
template <int BN, int PIX, bool BF>
__global__ void __launch_bounds__(256, 1)
conv_wgrad_tc_kernel(const __grid_constant__ TMapSet tmGys, const __grid_constant__ CUtensorMap tmX, const WgradParams p) {
  using Cfg = WgradCfg<BN, PIX, BF>;
  constexpr int STAGES = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::kStageBytes);
  const uint32_t bar0 = base + STAGES * Cfg::kStageBytes;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
  const uint32_t tfull_bar = bar0 + 8u * (2 * STAGES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int t = blockIdx.x;
  const int split = t % p.splits; t /= p.splits;
  const int tci = t % p.tiles_ci; t /= p.tiles_ci;
  const int tco = t % p.tiles_co; t /= p.tiles_co;
  const int tap = t;
  int r = tap / p.S, s = tap - r * p.S, mi = 0;
  if (p.up) { mi = tap >> 2; r = ((tap >> 1) & 1) + (mi >> 1); s = (tap & 1) + (mi & 1); }   // shift a - (1 - dy) = r - pad with pad = 1
  const CUtensorMap& tmGy = tmGys.m[mi];
  const int co0 = tco * 128, ci0 = tci * BN;
  const int pb_begin = split * p.pb_per_split;
  const int pb_end = min(p.num_pb, pb_begin + p.pb_per_split);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmGy);
    prefetch_tmap(&tmX);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(full_bar(i), 1);
      mbar_init(empty_bar(i), 1);
    }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(tmem_slot), Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int pb = pb_begin; pb < pb_end; ++pb) {
        int u = pb;
        const int tw = u % p.tiles_w; u /= p.tiles_w;
        const int th = u % p.tiles_h; u /= p.tiles_h;
        const int w0 = tw * p.bw, h0 = th * p.bh, n0 = u * p.bn;
        mbar_wait(empty_bar(stage), phase ^ 1);
        const uint32_t a_dst = base + stage * Cfg::kStageBytes;
        const uint32_t b_dst = a_dst + Cfg::kABytes;
        mbar_expect_tx(full_bar(stage), Cfg::kStageBytes);
        tma_load_5d(a_dst, &tmGy, full_bar(stage), 0, w0, h0, n0, co0 / Cfg::kCH);
        tma_load_5d(b_dst, &tmX, full_bar(stage), 0, w0 + s - p.pad, h0 + r - p.pad, n0, ci0 / Cfg::kCH);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc<BF>(128, BN, 1, 1);
      int stage = 0;
      uint32_t phase = 0;
      for (int pb = pb_begin; pb < pb_end; ++pb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t a_addr = base + stage * Cfg::kStageBytes;
        const uint32_t b_addr = a_addr + Cfg::kABytes;
        if (BF) {
#pragma unroll
          for (int kg = 0; kg < PIX / 16; ++kg) {  // K = 16 pixels per MMA = two 1024 B swizzle atoms per 64-channel block
            const uint64_t ad = make_smem_desc(a_addr + kg * 2048, Cfg::kBlkBytes, 1024, kLayoutSw128);
            const uint64_t bd = make_smem_desc(b_addr + kg * 2048, Cfg::kBlkBytes, 1024, kLayoutSw128);
            mma_bf16(tmem_base, ad, bd, idesc, (pb > pb_begin || kg > 0) ? 1u : 0u);
          }
        } else {
#pragma unroll
          for (int kg = 0; kg < PIX / 8; ++kg) {  // K = 8 pixels per MMA = two 512 B swizzle atoms per channel block
            const uint64_t ad = make_smem_desc(a_addr + kg * 1024, Cfg::kBlkBytes, 512, kLayoutSw128Base32);
            const uint64_t bd = make_smem_desc(b_addr + kg * 1024, Cfg::kBlkBytes, 512, kLayoutSw128Base32);
            mma_tf32(tmem_base, ad, bd, idesc, (pb > pb_begin || kg > 0) ? 1u : 0u);
          }
        }
        mma_commit(empty_bar(stage));
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      mma_commit(tfull_bar);
    }
  } else if (warp >= 4) {
    const int q = warp - 4;
    const int co = co0 + q * 32 + lane;
    const bool valid = co < p.Co;
    float* out = p.gw + ((int64_t)co * p.RS + tap) * p.Ci + ci0;
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + c, v);
      tmem_ld_wait();
      if (valid) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float a0 = p.alpha * __uint_as_float(v[j]), a1 = p.alpha * __uint_as_float(v[j + 1]);
          const float a2 = p.alpha * __uint_as_float(v[j + 2]), a3 = p.alpha * __uint_as_float(v[j + 3]);
          if (p.atomic) red_add_v4(out + c + j, a0, a1, a2, a3);
          else *reinterpret_cast<float4*>(out + c + j) = make_float4(a0, a1, a2, a3);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------ wgrad with tap reuse
// One CTA per (filter COLUMN s, 128 co, BN ci, split) computes the three taps (r = 0..2, s) at once: per K block (4 rows x 8
// columns = 32 output pixels of one image) it loads the gy block once and ONE x box of (4 + 2) rows x 8 columns shifted by
// s; the row shift r of a tap is an offset of r*8 pixels = r*1024 B into that box (MN-major operands: K = pixels advances
// in whole 1024-byte steps anyway).  Three accumulators (3 x BN TMEM columns).  Per MMA that is 2.4x less TMA / L2 traffic
// than one tap per CTA (40 KB per 12 MMAs instead of 64 KB per 8 at BN = 128).
template <int BN>
struct Wgrad3Cfg {
  static constexpr int kPix = 32, kBoxPix = 48;              // 4 x 8 output pixels; (4 + 2) x 8 input pixels
  static constexpr int kABlk = kPix * 128, kBBlk = kBoxPix * 128;
  static constexpr int kABytes = 4 * kABlk;
  static constexpr int kBBytes = (BN / 32) * kBBlk;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (200 * 1024 / kStageBytes) > 8 ? 8 : (200 * 1024 / kStageBytes);
  static constexpr int kTmemCols = (3 * BN <= 128) ? 128 : ((3 * BN <= 256) ? 256 : 512);
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
};

template <int BN>
__global__ void __launch_bounds__(256, 1)
conv_wgrad_tc3_kernel(const __grid_constant__ CUtensorMap tmGy, const __grid_constant__ CUtensorMap tmX, const WgradParams p) {
  using Cfg = Wgrad3Cfg<BN>;
  static_assert(3 * BN <= 512, "three accumulators must fit TMEM");
  constexpr int STAGES = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::kStageBytes);
  const uint32_t bar0 = base + STAGES * Cfg::kStageBytes;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
  const uint32_t tfull_bar = bar0 + 8u * (2 * STAGES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int t = blockIdx.x;
  const int split = t % p.splits; t /= p.splits;
  const int tci = t % p.tiles_ci; t /= p.tiles_ci;
  const int tco = t % p.tiles_co; t /= p.tiles_co;
  const int s = t;                                   // filter column of this CTA; rows r = 0..2 are its three accumulators
  const int co0 = tco * 128, ci0 = tci * BN;
  const int pb_begin = split * p.pb_per_split;
  const int pb_end = min(p.num_pb, pb_begin + p.pb_per_split);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmGy);
    prefetch_tmap(&tmX);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(full_bar(i), 1);
      mbar_init(empty_bar(i), 1);
    }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(tmem_slot), Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int pb = pb_begin; pb < pb_end; ++pb) {
        int u = pb;
        const int tw = u % p.tiles_w; u /= p.tiles_w;
        const int th = u % p.tiles_h; u /= p.tiles_h;
        const int w0 = tw * 8, h0 = th * 4, n0 = u;
        mbar_wait(empty_bar(stage), phase ^ 1);
        const uint32_t a_dst = base + stage * Cfg::kStageBytes;
        mbar_expect_tx(full_bar(stage), Cfg::kStageBytes);
        tma_load_5d(a_dst, &tmGy, full_bar(stage), 0, w0, h0, n0, co0 / 32);
        tma_load_5d(a_dst + Cfg::kABytes, &tmX, full_bar(stage), 0, w0 + s - p.pad, h0 - p.pad, n0, ci0 / 32);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(128, BN, 1, 1);
      int stage = 0;
      uint32_t phase = 0;
      for (int pb = pb_begin; pb < pb_end; ++pb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t a_addr = base + stage * Cfg::kStageBytes;
        const uint64_t ad0 = make_smem_desc(a_addr, Cfg::kABlk, 512, kLayoutSw128Base32);
        const uint64_t bd0 = make_smem_desc(a_addr + Cfg::kABytes, Cfg::kBBlk, 512, kLayoutSw128Base32);
        const uint32_t acc = (pb > pb_begin) ? 1u : 0u;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
#pragma unroll
          for (int kg = 0; kg < Cfg::kPix / 8; ++kg)   // K = 8 pixels per MMA = 1024 B; tap row r = +8 pixels in the x box
            mma_tf32(tmem_base + r * BN, ad0 + (uint64_t)((kg * 1024) >> 4), bd0 + (uint64_t)(((r + kg) * 1024) >> 4), idesc,
                     kg > 0 ? 1u : acc);
        }
        mma_commit(empty_bar(stage));
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      mma_commit(tfull_bar);
    }
  } else if (warp >= 4) {
    const int q = warp - 4;
    const int co = co0 + q * 32 + lane;
    const bool valid = co < p.Co;
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int r = 0; r < 3; ++r) {
      float* out = p.gw + ((int64_t)co * p.RS + (r * p.S + s)) * p.Ci + ci0;
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + r * BN + c, v);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float a0 = p.alpha * __uint_as_float(v[j]), a1 = p.alpha * __uint_as_float(v[j + 1]);
            const float a2 = p.alpha * __uint_as_float(v[j + 2]), a3 = p.alpha * __uint_as_float(v[j + 3]);
            if (p.atomic) red_add_v4(out + c + j, a0, a1, a2, a3);
            else *reinterpret_cast<float4*>(out + c + j) = make_float4(a0, a1, a2, a3);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------ wgrad of the folded layers, two taps per CTA
// conv_wgrad_tc3_kernel's idea for the 2x2 phase taps of the folded layers (WgradParams::up): one CTA per (phase, tap COLUMN b,
// 128 co, BN ci, split) computes the two taps a = 0, 1 of that column at once: per K block (4 rows x 8 columns = 32 pixels of the
// phase's low-resolution grid) it loads the gradient block once and ONE x box of (4 + 1) rows x 8 columns; tap row a is an offset
// of 8 pixels = 1024 B into the box.  Two accumulators (2 x BN TMEM columns).  Per MMA 7 KB of TMA / L2 traffic instead of 12 KB.
template <int BN>
struct WgradFold2Cfg {
  static constexpr int kPix = 32, kBoxPix = 40;              // 4 x 8 output pixels; (4 + 1) x 8 input pixels
  static constexpr int kABlk = kPix * 128, kBBlk = kBoxPix * 128;
  static constexpr int kABytes = 4 * kABlk;
  static constexpr int kBBytes = (BN / 32) * kBBlk;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (200 * 1024 / kStageBytes) > 8 ? 8 : (200 * 1024 / kStageBytes);
  static constexpr int kTmemCols = (2 * BN < 64) ? 64 : 2 * BN;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
};

template <int BN>
__global__ void __launch_bounds__(256, 1)
conv_wgrad_fold2_kernel(const __grid_constant__ TMapSet tmGys, const __grid_constant__ CUtensorMap tmX, const WgradParams p) {
  using Cfg = WgradFold2Cfg<BN>;
  static_assert(2 * BN <= 512, "two accumulators must fit TMEM");
  constexpr int STAGES = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::kStageBytes);
  const uint32_t bar0 = base + STAGES * Cfg::kStageBytes;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
  const uint32_t tfull_bar = bar0 + 8u * (2 * STAGES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int t = blockIdx.x;
  const int split = t % p.splits; t /= p.splits;
  const int tci = t % p.tiles_ci; t /= p.tiles_ci;
  const int tco = t % p.tiles_co; t /= p.tiles_co;
  const int ph = t >> 1, b = t & 1;                  // phase (dy, dx) and tap column of this CTA; tap rows a = 0, 1 are its accumulators
  const int dy = ph >> 1, dx = ph & 1;
  const CUtensorMap& tmGy = tmGys.m[ph];
  const int co0 = tco * 128, ci0 = tci * BN;
  const int pb_begin = split * p.pb_per_split;
  const int pb_end = min(p.num_pb, pb_begin + p.pb_per_split);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmGy);
    prefetch_tmap(&tmX);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(full_bar(i), 1);
      mbar_init(empty_bar(i), 1);
    }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(tmem_slot), Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int pb = pb_begin; pb < pb_end; ++pb) {
        int u = pb;
        const int tw = u % p.tiles_w; u /= p.tiles_w;
        const int th = u % p.tiles_h; u /= p.tiles_h;
        const int w0 = tw * 8, h0 = th * 4, n0 = u;
        mbar_wait(empty_bar(stage), phase ^ 1);
        const uint32_t a_dst = base + stage * Cfg::kStageBytes;
        mbar_expect_tx(full_bar(stage), Cfg::kStageBytes);
        tma_load_5d(a_dst, &tmGy, full_bar(stage), 0, w0, h0, n0, co0 / 32);
        // x shift of tap (a, b) in phase (dy, dx): (a + dy - 1, b + dx - 1); the box starts at a = 0
        tma_load_5d(a_dst + Cfg::kABytes, &tmX, full_bar(stage), 0, w0 + b + dx - 1, h0 + dy - 1, n0, ci0 / 32);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(128, BN, 1, 1);
      int stage = 0;
      uint32_t phase = 0;
      for (int pb = pb_begin; pb < pb_end; ++pb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t a_addr = base + stage * Cfg::kStageBytes;
        const uint64_t ad0 = make_smem_desc(a_addr, Cfg::kABlk, 512, kLayoutSw128Base32);
        const uint64_t bd0 = make_smem_desc(a_addr + Cfg::kABytes, Cfg::kBBlk, 512, kLayoutSw128Base32);
        const uint32_t acc = (pb > pb_begin) ? 1u : 0u;
#pragma unroll
        for (int a = 0; a < 2; ++a) {
#pragma unroll
          for (int kg = 0; kg < Cfg::kPix / 8; ++kg)   // K = 8 pixels per MMA = 1024 B; tap row a = +8 pixels in the x box
            mma_tf32(tmem_base + a * BN, ad0 + (uint64_t)((kg * 1024) >> 4), bd0 + (uint64_t)(((a + kg) * 1024) >> 4), idesc,
                     kg > 0 ? 1u : acc);
        }
        mma_commit(empty_bar(stage));
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      mma_commit(tfull_bar);
    }
  } else if (warp >= 4) {
    const int q = warp - 4;
    const int co = co0 + q * 32 + lane;
    const bool valid = co < p.Co;
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int a = 0; a < 2; ++a) {
      float* out = p.gw + ((int64_t)co * 16 + (ph * 4 + a * 2 + b)) * p.Ci + ci0;
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + a * BN + c, v);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float a0 = p.alpha * __uint_as_float(v[j]), a1 = p.alpha * __uint_as_float(v[j + 1]);
            const float a2 = p.alpha * __uint_as_float(v[j + 2]), a3 = p.alpha * __uint_as_float(v[j + 3]);
            if (p.atomic) red_add_v4(out + c + j, a0, a1, a2, a3);
            else *reinterpret_cast<float4*>(out + c + j) = make_float4(a0, a1, a2, a3);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

template <int BN>
int launch_wgrad_fold2(const TMapSet& tmGy, const CUtensorMap& tmX, const WgradParams& p, int grid, cudaStream_t st) {
  using Cfg = WgradFold2Cfg<BN>;
  static bool configured = false;
  if (!configured) {
    GLB_CUDA(cudaFuncSetAttribute(conv_wgrad_fold2_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured = true;
  }
  conv_wgrad_fold2_kernel<BN><<<grid, 256, Cfg::kSmemBytes, st>>>(tmGy, tmX, p);
  GLB_CHECK_LAUNCH("conv_wgrad_fold2_kernel");
  return GLB_OK;
}

template <int BN>
int launch_wgrad3(const CUtensorMap& tmGy, const CUtensorMap& tmX, const WgradParams& p, int grid, cudaStream_t st) {
  using Cfg = Wgrad3Cfg<BN>;
  static bool configured = false;
  if (!configured) {
    GLB_CUDA(cudaFuncSetAttribute(conv_wgrad_tc3_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured = true;
  }
  conv_wgrad_tc3_kernel<BN><<<grid, 256, Cfg::kSmemBytes, st>>>(tmGy, tmX, p);
  GLB_CHECK_LAUNCH("conv_wgrad_tc3_kernel");
  return GLB_OK;
}

template <int BN, int PIX, bool BF = false>
int launch_wgrad(const TMapSet& tmGy, const CUtensorMap& tmX, const WgradParams& p, int grid, cudaStream_t st) {
  using Cfg = WgradCfg<BN, PIX, BF>;
  static bool configured = false;
  if (!configured) {
    GLB_CUDA(cudaFuncSetAttribute(conv_wgrad_tc_kernel<BN, PIX, BF>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured = true;
  }
  conv_wgrad_tc_kernel<BN, PIX, BF><<<grid, 256, Cfg::kSmemBytes, st>>>(tmGy, tmX, p);
  GLB_CHECK_LAUNCH("conv_wgrad_tc_kernel");
  return GLB_OK;
}

bool conv_wgrad_tc_covers_impl(int N, int H, int W, int Ci, int Co, int R, int S, int pad, int chunk = 32) {
  if (Ci % chunk != 0 || Co % chunk != 0) return false;
  const int Ho = H + 2 * pad - R + 1, Wo = W + 2 * pad - S + 1;
  if (Ho <= 0 || Wo <= 0 || R * S > 25) return false;
  return true;
}

}  // namespace

bool conv_wgrad_tc_covers(int N, int H, int W, int Ci, int Co, int R, int S, int pad) {
  return conv_wgrad_tc_covers_impl(N, H, W, Ci, Co, R, S, pad);
}

bool conv_wgrad_bf16_covers(int N, int H, int W, int Ci, int Co, int R, int S, int pad) {
  return conv_wgrad_tc_covers_impl(N, H, W, Ci, Co, R, S, pad, 64);
}

template <bool BF>
static int conv_wgrad_tc_impl(const void* x, const void* gy, float* gw, int N, int H, int W, int Ci, int Co, int R, int S, int pad,
                              float alpha, cudaStream_t st) {
  constexpr int CH = BF ? 64 : 32;
  constexpr uint64_t ES = BF ? 2 : 4;
  if (!conv_wgrad_tc_covers_impl(N, H, W, Ci, Co, R, S, pad, CH)) {
    set_error("tcgen05 wgrad: shape not covered (need Ci and Co multiples of 32; bf16: of 64)");
    return GLB_ERR_UNSUPPORTED;
  }
  WgradParams p;
  p.gw = gw; p.Co = Co; p.Ci = Ci; p.RS = R * S; p.S = S; p.pad = pad; p.up = 0;
  p.N = N; p.Ho = H + 2 * pad - R + 1; p.Wo = W + 2 * pad - S + 1;
  {
    // tap-reuse kernel: 3x3 "same" convolutions whose maps tile into 4 x 8 pixel K blocks, enough of them for one wave
    bool tap3 = !BF && R == 3 && S == 3 && pad == 1 && p.Wo % 8 == 0 && p.Ho % 4 == 0 && (Ci % 128 == 0 || Ci == 64 || Ci == 32);
    if (const char* e = getenv("GLB_WGRAD_TAP3")) tap3 = tap3 && atoi(e) != 0;   // tuning experiments only
    const int bn3 = Ci % 128 == 0 ? 128 : Ci;
    p.tiles_w = p.Wo / 8; p.tiles_h = p.Ho / 4; p.tiles_n = N;
    p.num_pb = p.tiles_w * p.tiles_h * p.tiles_n;
    p.tiles_co = (Co + 127) / 128;
    p.tiles_ci = Ci / bn3;
    const int tiles3 = p.tiles_co * p.tiles_ci * 3;
    // measured (tools/check_tc.py): +5..18 % on the 128x128 layers, -7 % at 32x32 / 64x64 where the one-tap kernel needs
    // no or little split-K (three accumulators per CTA triple the red.add epilogue) -> high-resolution layers only
    if (tap3 && p.num_pb >= 2048 && tiles3 <= kNumSMs) {
      int splits = kNumSMs / tiles3;
      if (splits > p.num_pb / 8) splits = p.num_pb / 8;
      if (splits < 1) splits = 1;
      p.pb_per_split = (p.num_pb + splits - 1) / splits;
      p.splits = (p.num_pb + p.pb_per_split - 1) / p.pb_per_split;
      p.alpha = alpha * kTf32TruncComp;
      p.atomic = p.splits > 1 ? 1 : 0;
      p.bw = 8; p.bh = 4; p.bn = 1;
      if (p.atomic) GLB_CUDA(cudaMemsetAsync(gw, 0, sizeof(float) * (size_t)Co * R * S * Ci, st));
      CUtensorMap tmGy, tmX;
      const uint64_t dg[5] = {32u, (uint64_t)p.Wo, (uint64_t)p.Ho, (uint64_t)N, (uint64_t)(Co / 32)};
      const uint64_t sg[4] = {(uint64_t)Co * 4, (uint64_t)p.Wo * Co * 4, (uint64_t)p.Ho * p.Wo * Co * 4, 128u};
      const uint32_t bg[5] = {32u, 8u, 4u, 1u, 4u};
      int rc = make_tmap_f32(&tmGy, gy, 5, dg, sg, bg, "wgrad gy (4x8 blocks)", true);
      if (rc) return rc;
      const uint64_t dx[5] = {32u, (uint64_t)W, (uint64_t)H, (uint64_t)N, (uint64_t)(Ci / 32)};
      const uint64_t sx[4] = {(uint64_t)Ci * 4, (uint64_t)W * Ci * 4, (uint64_t)H * W * Ci * 4, 128u};
      const uint32_t bx[5] = {32u, 8u, 6u, 1u, (uint32_t)(bn3 / 32)};
      rc = make_tmap_f32(&tmX, x, 5, dx, sx, bx, "wgrad x (6x8 boxes)", true);
      if (rc) return rc;
      const int grid = tiles3 * p.splits;
      switch (bn3) {
        case 128: return launch_wgrad3<128>(tmGy, tmX, p, grid, st);
        case 64: return launch_wgrad3<64>(tmGy, tmX, p, grid, st);
        case 32: return launch_wgrad3<32>(tmGy, tmX, p, grid, st);
      }
    }
  }
  const int BN = Ci % 256 == 0 ? 256 : (Ci % 128 == 0 ? 128 : (Ci % 64 == 0 ? 64 : 32));
  const int PIX = (BN >= 256 ? 32 : 64) * (BF ? 2 : 1);      // same bytes and MMAs per stage for either operand type
  p.bw = next_pow2(p.Wo) < PIX ? next_pow2(p.Wo) : PIX;
  p.bh = next_pow2(p.Ho) < PIX / p.bw ? next_pow2(p.Ho) : PIX / p.bw;
  p.bn = PIX / (p.bw * p.bh);
  p.tiles_w = (p.Wo + p.bw - 1) / p.bw;
  p.tiles_h = (p.Ho + p.bh - 1) / p.bh;
  p.tiles_n = (N + p.bn - 1) / p.bn;
  p.num_pb = p.tiles_w * p.tiles_h * p.tiles_n;
  p.tiles_co = (Co + 127) / 128;
  p.tiles_ci = Ci / BN;
  const int tiles = p.tiles_co * p.tiles_ci * p.RS;
  // one wave: every CTA pays the 128 x BN red.add epilogue once, so fewer, longer work items win (ncu: with ~2.4 waves
  // of short items the atomics epilogue cost more than the MMA main loop)
  int splits = kNumSMs / tiles;
  if (splits > p.num_pb / 4) splits = p.num_pb / 4;         // >= 4 pixel blocks of K per work item
  if (splits > p.num_pb) splits = p.num_pb;
  if (const char* e = getenv("GLB_WGRAD_SPLITS")) splits = atoi(e);  // tuning experiments only
  if (splits < 1) splits = 1;
  p.pb_per_split = (p.num_pb + splits - 1) / splits;
  p.splits = (p.num_pb + p.pb_per_split - 1) / p.pb_per_split;  // every split owns >= 1 pixel block
  p.alpha = BF ? alpha : alpha * kTf32TruncComp;
  p.atomic = p.splits > 1 ? 1 : 0;
  if (p.atomic) GLB_CUDA(cudaMemsetAsync(gw, 0, sizeof(float) * (size_t)Co * R * S * Ci, st));

  TMapSet tmGy;
  CUtensorMap tmX;
  {  // (32 ch, Wo, Ho, N, Co/32): box = PIX pixels x 4 channel blocks (blocks beyond Co/32 are zero-filled: M padded to 128)
    const uint64_t dims[5] = {(uint64_t)CH, (uint64_t)p.Wo, (uint64_t)p.Ho, (uint64_t)N, (uint64_t)(Co / CH)};
    const uint64_t strides[4] = {(uint64_t)Co * ES, (uint64_t)p.Wo * Co * ES, (uint64_t)p.Ho * p.Wo * Co * ES, 128u};
    const uint32_t box[5] = {(uint32_t)CH, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn, (uint32_t)(128 / CH)};
    int rc = make_tmap(&tmGy.m[0], gy, 5, dims, strides, box, "wgrad gy", !BF, BF);
    if (rc) return rc;
    tmGy.m[1] = tmGy.m[2] = tmGy.m[3] = tmGy.m[0];
  }
  {
    const uint64_t dims[5] = {(uint64_t)CH, (uint64_t)W, (uint64_t)H, (uint64_t)N, (uint64_t)(Ci / CH)};
    const uint64_t strides[4] = {(uint64_t)Ci * ES, (uint64_t)W * Ci * ES, (uint64_t)H * W * Ci * ES, 128u};
    const uint32_t box[5] = {(uint32_t)CH, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn, (uint32_t)(BN / CH)};
    int rc = make_tmap(&tmX, x, 5, dims, strides, box, "wgrad x", !BF, BF);
    if (rc) return rc;
  }
  const int grid = tiles * p.splits;
  if (BF) {
    switch (BN) {
      case 256: return launch_wgrad<256, 64, true>(tmGy, tmX, p, grid, st);
      case 128: return launch_wgrad<128, 128, true>(tmGy, tmX, p, grid, st);
      case 64: return launch_wgrad<64, 128, true>(tmGy, tmX, p, grid, st);
    }
    return GLB_ERR_UNSUPPORTED;
  }
  switch (BN) {
    case 256: return launch_wgrad<256, 32>(tmGy, tmX, p, grid, st);
    case 128: return launch_wgrad<128, 64>(tmGy, tmX, p, grid, st);
    case 64: return launch_wgrad<64, 64>(tmGy, tmX, p, grid, st);
    case 32: return launch_wgrad<32, 64>(tmGy, tmX, p, grid, st);
  }
  return GLB_ERR_UNSUPPORTED;
}

int conv_wgrad_tc(const float* x, const float* gy, float* gw, int N, int H, int W, int Ci, int Co, int R, int S, int pad,
                  float alpha, cudaStream_t st) {
  return conv_wgrad_tc_impl<false>(x, gy, gw, N, H, W, Ci, Co, R, S, pad, alpha, st);
}

int conv_wgrad_bf16(const void* x, const void* gy, float* gw, int N, int H, int W, int Ci, int Co, int R, int S, int pad,
                    float alpha, cudaStream_t st) {
  return conv_wgrad_tc_impl<true>(x, gy, gw, N, H, W, Ci, Co, R, S, pad, alpha, st);
}

// ------------------------------------------------------------------------------------------------ upconv: weights and wgrad
namespace {
// taps of the 3-tap axis that land on tap `a` of phase `d`:  d=0: a=0 -> {0}, a=1 -> {1,2};  d=1: a=0 -> {0,1}, a=1 -> {2}
__device__ __forceinline__ void up_taps(int d, int a, int& lo, int& hi) {
  lo = d == 0 ? (a == 0 ? 0 : 1) : (a == 0 ? 0 : 2);
  hi = d == 0 ? (a == 0 ? 0 : 2) : (a == 0 ? 1 : 2);
}

// Operand re-layout of the folded convolutions, ONE pass over w [Co][3][3][Ci] per layer and optimiser step.  With
//   S[ph][a][b][co][ci] = sum_{r in T(dy,a), s in T(dx,b)} wsrc[co][r][s][ci],   wsrc = w (upconv) or w with flipped taps (downconv)
// the four operands are
//   upconv   forward  wp [ph][co][a][b][ci]          = S[ph][a][b]        (ci innermost: "A" output)
//            dgrad    wt [ci][ph*4 + a'*2 + b'][co]  = S[ph][1-a'][1-b']  (co innermost: "B" output, transposed)
//   downconv forward  wt'[co][ph*4 + a'*2 + b'][ci]  = S[ph][1-a'][1-b']  ("A")
//            dgrad    wp'[ph][ci][a][b][co]          = S[ph][a][b]        ("B")
// A block stages the nine 32 x 32 (co, ci) tap tiles in shared memory and writes all 16 slots of both outputs from there;
// 1024 threads, one (co, ci) element of either orientation per thread (with 256 threads and four rows per thread the
// serial instruction stream of a block took 22 us whatever the layer size).
template <bool DOWN>
__global__ void __launch_bounds__(1024) fold_weights_kernel(const float* __restrict__ w, float* __restrict__ outA, float* __restrict__ outB,
                                                           int Co, int Ci) {
  __shared__ float t[9][32][33];
  const int x = threadIdx.x & 31, y = threadIdx.x >> 5;
  const int co0 = blockIdx.y * 32, ci0 = blockIdx.x * 32;
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const int src = DOWN ? 8 - tap : tap;
    const int co = co0 + y, ci = ci0 + x;
    t[tap][y][x] = (co < Co && ci < Ci) ? __ldg(w + ((int64_t)co * 9 + src) * Ci + ci) : 0.f;
  }
  __syncthreads();
  // per (row i): the nine taps of element (co0+i, ci0+x) ["A" orientation] and of (co0+x, ci0+i) ["B"] go to registers, the 16
  // slot sums are compile-time tap subsets (everything below unrolls), stores are 128-byte rows of either output
  {
    const int i = y;
    float va[9], vb[9];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) { va[tap] = t[tap][i][x]; vb[tap] = t[tap][x][i]; }
    const int coA = co0 + i, ciA = ci0 + x, coB = co0 + x, ciB = ci0 + i;
    const bool okA = outA != nullptr && coA < Co && ciA < Ci, okB = outB != nullptr && coB < Co && ciB < Ci;
#pragma unroll
    for (int slot = 0; slot < 16; ++slot) {
      const int ph = slot >> 2, a = (slot >> 1) & 1, b = slot & 1;
      const int flipped = ph * 4 + (1 - a) * 2 + (1 - b);
      int r0, r1, s0, s1;
      up_taps(ph >> 1, a, r0, r1);
      up_taps(ph & 1, b, s0, s1);
      float sa = 0.f, sb = 0.f;
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int q = 0; q < 3; ++q)
          if (r >= r0 && r <= r1 && q >= s0 && q <= s1) { sa += va[r * 3 + q]; sb += vb[r * 3 + q]; }
      if (okA) {
        if (DOWN) outA[((int64_t)coA * 16 + flipped) * Ci + ciA] = sa;
        else outA[(((int64_t)ph * Co + coA) * 4 + a * 2 + b) * Ci + ciA] = sa;
      }
      if (okB) {
        if (DOWN) outB[(((int64_t)ph * Ci + ciB) * 4 + a * 2 + b) * Co + coB] = sb;
        else outB[((int64_t)ciB * 16 + flipped) * Co + coB] = sb;
      }
    }
  }
}

// downconv weight gradient: gw[co][r][s][ci] = gw'[ci][2-r][2-s][co] with gw' the 3x3 fold (below) of gwp [ci][16][co]: fold
// and transpose in one pass (16 ci x 32 co tiles of all 16 slots staged in shared memory)
__global__ void __launch_bounds__(512) fold_wgrad_transposed_kernel(const float* __restrict__ gwp, float* __restrict__ gw, int Co, int Ci) {
  __shared__ float t[16][16][33];
  const int x = threadIdx.x & 31, y = threadIdx.x >> 5;
  const int co0 = blockIdx.y * 32, ci0 = blockIdx.x * 16;
#pragma unroll
  for (int slot = 0; slot < 16; ++slot) {
    const int ci = ci0 + y, co = co0 + x;                  // 512 threads: y = 0..15
    t[slot][y][x] = (ci < Ci && co < Co) ? __ldg(gwp + ((int64_t)ci * 16 + slot) * Co + co) : 0.f;
  }
  __syncthreads();
  const int xi = threadIdx.x & 15, j = threadIdx.x >> 4;
#pragma unroll
  for (int rs = 0; rs < 9; ++rs) {
    const int r = 2 - rs / 3, q = 2 - rs % 3;              // tap of gw' behind tap rs of gw
    const int ay[2] = {r == 0 ? 0 : 1, r == 2 ? 1 : 0}, ax[2] = {q == 0 ? 0 : 1, q == 2 ? 1 : 0};
    {
      float acc = 0.f;
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) acc += t[(dy * 2 + dx) * 4 + ay[dy] * 2 + ax[dx]][xi][j];
      const int co = co0 + j, ci = ci0 + xi;
      if (co < Co && ci < Ci) gw[((int64_t)co * 9 + rs) * Ci + ci] = acc;
    }
  }
}

// gw[co][r][s][ci] = sum over the (phase, tap) pairs whose pre-summed tap contains (r, s) of gwp[co][ph*4 + a*2 + b][ci]
__global__ void upconv_wgrad_fold_kernel(const float4* __restrict__ gwp, float4* __restrict__ gw, int Co, int Ci4) {
  const int64_t total = (int64_t)Co * 9 * Ci4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % Ci4);
    int64_t t = i / Ci4;
    const int rs = (int)(t % 9), co = (int)(t / 9);
    const int r = rs / 3, s = rs % 3;
    // axis sources (d, a): r=0: (0,0),(1,0)   r=1: (0,1),(1,0)   r=2: (0,1),(1,1)
    const int ay[2] = {r == 0 ? 0 : 1, r == 2 ? 1 : 0}, ax[2] = {s == 0 ? 0 : 1, s == 2 ? 1 : 0};
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int t16 = (dy * 2 + dx) * 4 + ay[dy] * 2 + ax[dx];
        const float4 v = __ldg(gwp + ((int64_t)co * 16 + t16) * Ci4 + c4);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    gw[i] = acc;
  }
}
}  // namespace

int conv_upconv_weights(const float* w, float* wp, float* wt, int Co, int Ci, cudaStream_t st) {
  if (wp == nullptr && wt == nullptr) return GLB_OK;
  dim3 grid((Ci + 31) / 32, (Co + 31) / 32);
  fold_weights_kernel<false><<<grid, 1024, 0, st>>>(w, wp, wt, Co, Ci);
  GLB_CHECK_LAUNCH("fold_weights_kernel");
  return GLB_OK;
}

// gw [Co][3][3][Ci] of conv3x3(upsample2x(x)); gwp = [Co][16][Ci] scratch (phase/tap gradients, folded at the end)
template <bool BF>
static int upconv_wgrad_impl(const void* x, const void* gy, float* gwp, float* gw, int N, int H, int W, int Ci, int Co, float alpha,
                             bool transposed_out, cudaStream_t st) {
  constexpr int CH = BF ? 64 : 32;
  constexpr uint64_t ES = BF ? 2 : 4;
  if (!conv_upconv_covers(2, N, H, W, Ci, Co, CH)) {
    set_error("upconv wgrad: shape not covered (Ci and Co multiples of 32)");
    return GLB_ERR_UNSUPPORTED;
  }
  WgradParams p;
  p.gw = gwp; p.Co = Co; p.Ci = Ci; p.RS = 16; p.S = 4; p.pad = 1; p.up = 1;
  p.N = N; p.Ho = H; p.Wo = W;
  // two taps per CTA (conv_wgrad_fold2_kernel): maps whose low-resolution grid tiles into 4 x 8 pixel K blocks, enough of them
  {
    // measured on B200 (cfg2 layers, batch 8): SLOWER than one tap per CTA -- 47 vs 37 us at 16^2 -> 32^2, 65 vs 61 us at
    // 64^2 -> 128^2, whole step 809 vs 815 img/s: half as many CTAs with twice the red.add epilogue and three pipeline stages
    // instead of four outweigh the saved operand traffic (the same finding as conv_wgrad_tc3_kernel below 128^2).  Kept for
    // A/B runs only: GLB_WGRAD_FOLD2=1.
    bool fold2 = false;
    if (const char* e = getenv("GLB_WGRAD_FOLD2"))
      fold2 = atoi(e) != 0 && !BF && W % 8 == 0 && H % 4 == 0 && (Ci % 128 == 0 || Ci == 64 || Ci == 32);
    const int bn2 = Ci % 256 == 0 ? 256 : (Ci % 128 == 0 ? 128 : Ci);
    const int num_pb = (W / 8) * (H / 4) * N;
    if (fold2 && num_pb >= 16) {
      p.tiles_w = W / 8; p.tiles_h = H / 4; p.tiles_n = N;
      p.num_pb = num_pb;
      p.tiles_co = (Co + 127) / 128;
      p.tiles_ci = Ci / bn2;
      const int tiles2 = p.tiles_co * p.tiles_ci * 8;
      int splits = kNumSMs / tiles2;
      if (splits > num_pb / 8) splits = num_pb / 8;
      if (splits < 1) splits = 1;
      p.pb_per_split = (num_pb + splits - 1) / splits;
      p.splits = (num_pb + p.pb_per_split - 1) / p.pb_per_split;
      p.alpha = alpha * kTf32TruncComp;
      p.atomic = p.splits > 1 ? 1 : 0;
      p.bw = 8; p.bh = 4; p.bn = 1;
      if (p.atomic) GLB_CUDA(cudaMemsetAsync(gwp, 0, sizeof(float) * (size_t)Co * 16 * Ci, st));
      TMapSet tmGy;
      CUtensorMap tmX;
      const uint64_t dg[5] = {32u, (uint64_t)W, (uint64_t)H, (uint64_t)N, (uint64_t)(Co / 32)};
      const uint64_t sg[4] = {(uint64_t)2 * Co * 4, (uint64_t)2 * (2 * W) * Co * 4, (uint64_t)(2 * H) * (2 * W) * Co * 4, 128u};
      const uint32_t bg[5] = {32u, 8u, 4u, 1u, 4u};
      for (int ph = 0; ph < 4; ++ph) {
        const char* gbase = (const char*)gy + ((int64_t)(ph >> 1) * (2 * W) + (ph & 1)) * Co * 4;
        int rc = make_tmap_f32(&tmGy.m[ph], gbase, 5, dg, sg, bg, "upconv wgrad gy (phase view, 4x8 blocks)", true);
        if (rc) return rc;
      }
      const uint64_t dx[5] = {32u, (uint64_t)W, (uint64_t)H, (uint64_t)N, (uint64_t)(Ci / 32)};
      const uint64_t sx[4] = {(uint64_t)Ci * 4, (uint64_t)W * Ci * 4, (uint64_t)H * W * Ci * 4, 128u};
      const uint32_t bx[5] = {32u, 8u, 5u, 1u, (uint32_t)(bn2 / 32)};
      int rc = make_tmap_f32(&tmX, x, 5, dx, sx, bx, "upconv wgrad x (5x8 boxes)", true);
      if (rc) return rc;
      const int grid2 = tiles2 * p.splits;
      switch (bn2) {
        case 256: rc = launch_wgrad_fold2<256>(tmGy, tmX, p, grid2, st); break;
        case 128: rc = launch_wgrad_fold2<128>(tmGy, tmX, p, grid2, st); break;
        case 64: rc = launch_wgrad_fold2<64>(tmGy, tmX, p, grid2, st); break;
        case 32: rc = launch_wgrad_fold2<32>(tmGy, tmX, p, grid2, st); break;
        default: rc = GLB_ERR_UNSUPPORTED;
      }
      if (rc != GLB_OK) return rc;
      if (transposed_out) {
        dim3 tgrid((Co + 15) / 16, (Ci + 31) / 32);
        fold_wgrad_transposed_kernel<<<tgrid, 512, 0, st>>>(gwp, gw, /*Co_layer =*/Ci, /*Ci_layer =*/Co);
        GLB_CHECK_LAUNCH("fold_wgrad_transposed_kernel");
        return GLB_OK;
      }
      const int64_t total = (int64_t)Co * 9 * (Ci / 4);
      const int fgrid = (int)((total + 255) / 256 < 4 * kNumSMs ? (total + 255) / 256 : 4 * kNumSMs);
      upconv_wgrad_fold_kernel<<<fgrid, 256, 0, st>>>((const float4*)gwp, (float4*)gw, Co, Ci / 4);
      GLB_CHECK_LAUNCH("upconv_wgrad_fold_kernel");
      return GLB_OK;
    }
  }
  const int BN = Ci % 256 == 0 ? 256 : (Ci % 128 == 0 ? 128 : (Ci % 64 == 0 ? 64 : 32));
  const int PIX = (BN >= 256 ? 32 : 64) * (BF ? 2 : 1);
  p.bw = next_pow2(W) < PIX ? next_pow2(W) : PIX;
  p.bh = next_pow2(H) < PIX / p.bw ? next_pow2(H) : PIX / p.bw;
  p.bn = PIX / (p.bw * p.bh);
  p.tiles_w = (W + p.bw - 1) / p.bw;
  p.tiles_h = (H + p.bh - 1) / p.bh;
  p.tiles_n = (N + p.bn - 1) / p.bn;
  p.num_pb = p.tiles_w * p.tiles_h * p.tiles_n;
  p.tiles_co = (Co + 127) / 128;
  p.tiles_ci = Ci / BN;
  const int tiles = p.tiles_co * p.tiles_ci * 16;
  int splits = kNumSMs / tiles;
  if (splits > p.num_pb / 4) splits = p.num_pb / 4;
  if (splits > p.num_pb) splits = p.num_pb;
  if (splits < 1) splits = 1;
  p.pb_per_split = (p.num_pb + splits - 1) / splits;
  p.splits = (p.num_pb + p.pb_per_split - 1) / p.pb_per_split;
  p.alpha = BF ? alpha : alpha * kTf32TruncComp;
  p.atomic = p.splits > 1 ? 1 : 0;
  if (p.atomic) GLB_CUDA(cudaMemsetAsync(gwp, 0, sizeof(float) * (size_t)Co * 16 * Ci, st));
  TMapSet tmGy;
  CUtensorMap tmX;
  {
    const uint64_t dims[5] = {(uint64_t)CH, (uint64_t)W, (uint64_t)H, (uint64_t)N, (uint64_t)(Co / CH)};
    const uint64_t strides[4] = {(uint64_t)2 * Co * ES, (uint64_t)2 * (2 * W) * Co * ES, (uint64_t)(2 * H) * (2 * W) * Co * ES, 128u};
    const uint32_t box[5] = {(uint32_t)CH, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn, (uint32_t)(128 / CH)};
    for (int ph = 0; ph < 4; ++ph) {
      const char* base = (const char*)gy + ((int64_t)(ph >> 1) * (2 * W) + (ph & 1)) * Co * ES;
      int rc = make_tmap(&tmGy.m[ph], base, 5, dims, strides, box, "upconv wgrad gy (phase view)", !BF, BF);
      if (rc) return rc;
    }
  }
  {
    const uint64_t dims[5] = {(uint64_t)CH, (uint64_t)W, (uint64_t)H, (uint64_t)N, (uint64_t)(Ci / CH)};
    const uint64_t strides[4] = {(uint64_t)Ci * ES, (uint64_t)W * Ci * ES, (uint64_t)H * W * Ci * ES, 128u};
    const uint32_t box[5] = {(uint32_t)CH, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn, (uint32_t)(BN / CH)};
    int rc = make_tmap(&tmX, x, 5, dims, strides, box, "upconv wgrad x", !BF, BF);
    if (rc) return rc;
  }
  const int grid = tiles * p.splits;
  int rc = GLB_ERR_UNSUPPORTED;
  if (BF) {
    switch (BN) {
      case 256: rc = launch_wgrad<256, 64, true>(tmGy, tmX, p, grid, st); break;
      case 128: rc = launch_wgrad<128, 128, true>(tmGy, tmX, p, grid, st); break;
      case 64: rc = launch_wgrad<64, 128, true>(tmGy, tmX, p, grid, st); break;
    }
  } else {
    switch (BN) {
      case 256: rc = launch_wgrad<256, 32>(tmGy, tmX, p, grid, st); break;
      case 128: rc = launch_wgrad<128, 64>(tmGy, tmX, p, grid, st); break;
      case 64: rc = launch_wgrad<64, 64>(tmGy, tmX, p, grid, st); break;
      case 32: rc = launch_wgrad<32, 64>(tmGy, tmX, p, grid, st); break;
    }
  }
  if (rc != GLB_OK) return rc;
  if (transposed_out) {     // downconv: this problem's (Co, Ci) are the layer's (Ci, Co); gw is the layer's [Co_layer][3][3][Ci_layer]
    dim3 tgrid((Co + 15) / 16, (Ci + 31) / 32);
    fold_wgrad_transposed_kernel<<<tgrid, 512, 0, st>>>(gwp, gw, /*Co_layer =*/Ci, /*Ci_layer =*/Co);
    GLB_CHECK_LAUNCH("fold_wgrad_transposed_kernel");
    return GLB_OK;
  }
  const int64_t total = (int64_t)Co * 9 * (Ci / 4);
  const int fgrid = (int)((total + 255) / 256 < 4 * kNumSMs ? (total + 255) / 256 : 4 * kNumSMs);
  upconv_wgrad_fold_kernel<<<fgrid, 256, 0, st>>>((const float4*)gwp, (float4*)gw, Co, Ci / 4);
  GLB_CHECK_LAUNCH("upconv_wgrad_fold_kernel");
  return GLB_OK;
}

int conv_upconv_wgrad_tc(const float* x, const float* gy, float* gwp, float* gw, int N, int H, int W, int Ci, int Co, float alpha,
                         cudaStream_t st) {
  return upconv_wgrad_impl<false>(x, gy, gwp, gw, N, H, W, Ci, Co, alpha, false, st);
}
int conv_upconv_wgrad_bf16(const void* x, const void* gy, float* gwp, float* gw, int N, int H, int W, int Ci, int Co, float alpha,
                           cudaStream_t st) {
  return upconv_wgrad_impl<true>(x, gy, gwp, gw, N, H, W, Ci, Co, alpha, false, st);
}

// ------------------------------------------------------------------------------------------------ 2x2 average pool folded into the conv
// y = avgpool2x2(conv3x3_same(x, w))  (reference progan/architectures.py:267-284: Conv2dEx followed by nn.AvgPool2d in the
// discriminator blocks) = 0.25 * U^T C_w x with U the nearest 2x upsample: the ADJOINT structure of the upsample-folded
// convolution above.  With w' = the flipped / transposed weights (C_w = C_w'^T):
//   fprop  y  = 0.25 * upconv_dgrad(x;  w')      (K over (input phase, 2x2 tap, Ci): a stride-2 4x4 convolution, 4/9 of the MMAs,
//                                                 the full-resolution conv output never exists)
//   dgrad  gx = 0.25 * upconv_fprop(gy; w')
//   wgrad  gw = 0.25 * transpose_flip(upconv_wgrad(x_lo := gy, gy_hi := x))
// so the same three kernels serve it; H, W below are the LOW-resolution (output) dims, x / gx are [N,2H,2W,Ci].
// glb_downconv_weights: wp [4*Ci][2][2][Co] (dgrad) and wt [Co][16][Ci] (fprop) = glb_upconv_weights of w' (fold_weights_kernel<true>).
bool conv_downconv_covers(int kind, int N, int H, int W, int Ci, int Co, int chunk) {
  switch (kind) {
    case 0: return conv_upconv_covers(1, N, H, W, Co, Ci, chunk);
    case 1: return conv_upconv_covers(0, N, H, W, Co, Ci, chunk);
    case 2: return conv_upconv_covers(2, N, H, W, Co, Ci, chunk);
  }
  return false;
}
bool conv_downconv_covers(int kind, int N, int H, int W, int Ci, int Co) { return conv_downconv_covers(kind, N, H, W, Ci, Co, 32); }

int conv_downconv_weights(const float* w, float* wp, float* wt, int Co, int Ci, cudaStream_t st) {
  if (wp == nullptr && wt == nullptr) return GLB_OK;
  dim3 grid((Ci + 31) / 32, (Co + 31) / 32);
  fold_weights_kernel<true><<<grid, 1024, 0, st>>>(w, /*A: forward operand*/ wt, /*B: dgrad operand*/ wp, Co, Ci);
  GLB_CHECK_LAUNCH("fold_weights_kernel");
  return GLB_OK;
}

int conv_downconv_fprop_tc(const float* x, const float* wt, const float* bias, float* y, int N, int H, int W, int Ci, int Co, float alpha,
                           float bias_scale, int act, float slope, cudaStream_t st) {
  if (!conv_downconv_covers(0, N, H, W, Ci, Co)) {
    set_error("downconv fprop: shape not covered (Ci % 32 == 0, Co in {32, 64} or a multiple of 128)");
    return GLB_ERR_UNSUPPORTED;
  }
  return upconv_launch<false>(2, x, wt, bias, y, N, H, W, Ci, Co, 0.25f * alpha, bias_scale, act, slope, st);
}
int conv_downconv_fprop_bf16(const void* x, const void* wt, const float* bias, float* y, int N, int H, int W, int Ci, int Co, float alpha,
                             float bias_scale, int act, float slope, cudaStream_t st) {
  if (!conv_downconv_covers(0, N, H, W, Ci, Co, 64)) {
    set_error("downconv fprop (bf16): shape not covered");
    return GLB_ERR_UNSUPPORTED;
  }
  return upconv_launch<true>(2, x, wt, bias, y, N, H, W, Ci, Co, 0.25f * alpha, bias_scale, act, slope, st);
}

int conv_downconv_dgrad_tc(const float* gy, const float* wp, float* gx, int N, int H, int W, int Ci, int Co, float alpha, cudaStream_t st) {
  if (!conv_downconv_covers(1, N, H, W, Ci, Co)) {
    set_error("downconv dgrad: shape not covered (Co % 32 == 0, Ci in {32, 64} or a multiple of 128)");
    return GLB_ERR_UNSUPPORTED;
  }
  return upconv_launch<false>(1, gy, wp, nullptr, gx, N, H, W, Co, Ci, 0.25f * alpha, 0.f, GLB_ACT_NONE, 0.f, st);
}
int conv_downconv_dgrad_bf16(const void* gy, const void* wp, float* gx, int N, int H, int W, int Ci, int Co, float alpha, cudaStream_t st) {
  if (!conv_downconv_covers(1, N, H, W, Ci, Co, 64)) {
    set_error("downconv dgrad (bf16): shape not covered");
    return GLB_ERR_UNSUPPORTED;
  }
  return upconv_launch<true>(1, gy, wp, nullptr, gx, N, H, W, Co, Ci, 0.25f * alpha, 0.f, GLB_ACT_NONE, 0.f, st);
}

// gwp [Ci][16][Co] is scratch
int conv_downconv_wgrad_tc(const float* x, const float* gy, float* gwp, float* gw, int N, int H, int W, int Ci, int Co, float alpha,
                           cudaStream_t st) {
  if (!conv_downconv_covers(2, N, H, W, Ci, Co)) {
    set_error("downconv wgrad: shape not covered (Ci and Co multiples of 32)");
    return GLB_ERR_UNSUPPORTED;
  }
  return upconv_wgrad_impl<false>(gy, x, gwp, gw, N, H, W, /*Ci' =*/Co, /*Co' =*/Ci, 0.25f * alpha, true, st);
}
int conv_downconv_wgrad_bf16(const void* x, const void* gy, float* gwp, float* gw, int N, int H, int W, int Ci, int Co, float alpha,
                             cudaStream_t st) {
  if (!conv_downconv_covers(2, N, H, W, Ci, Co, 64)) {
    set_error("downconv wgrad (bf16): shape not covered");
    return GLB_ERR_UNSUPPORTED;
  }
  return upconv_wgrad_impl<true>(gy, x, gwp, gw, N, H, W, /*Ci' =*/Co, /*Co' =*/Ci, 0.25f * alpha, true, st);
}

}  // namespace glb
