/*
 * ganlab_b200.h -- C-ABI of the B200-native (sm_100a) kernels behind gan-lab's G/D training step.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  The
 * reference has NO native code; every entry point below replaces one or more ATen call sites of the
 * reference's Python layers (cited per function, paths relative to the reference's gan_lab/).
 *
 * Conventions
 *   - all tensors are fp32 and live in device memory owned by the caller (torch allocates; kernels
 *     never allocate or free);
 *   - feature maps are NHWC ("channels-last": [N][H][W][C], C fastest); RGB images are NCHW exactly as
 *     the reference passes them ([N][3][H][W]);
 *   - conv weights are KRSC ([Cout][R][S][Cin] = the OIHW parameter stored channels-last); linear
 *     weights are [out][in] row-major exactly as nn.Linear stores them;
 *   - every function enqueues on `stream` (a cudaStream_t), never synchronises the device, is
 *     re-entrant and safe under CUDA-graph capture;
 *   - return value: 0 = ok, non-zero = error (message via glb_last_error(), thread-local);
 *   - `impl`: GLB_IMPL_FP32 = exact fp32 FFMA implicit GEMM; GLB_IMPL_TF32 = tcgen05/TMEM/TMA
 *     tensor-core implicit GEMM (TF32 operands, fp32 accumulate).  A shape the tensor-core kernel
 *     does not cover returns GLB_ERR_UNSUPPORTED (the host then picks GLB_IMPL_FP32 explicitly --
 *     both are CUDA kernels; there is no CPU path anywhere).
 */
#ifndef GANLAB_B200_H_
#define GANLAB_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GLB_OK 0
#define GLB_ERR_CUDA 1
#define GLB_ERR_SHAPE 2
#define GLB_ERR_UNSUPPORTED 3

#define GLB_IMPL_FP32 0
#define GLB_IMPL_TF32 1
#define GLB_IMPL_BF16 2   /* the *_bf16 entry points below; not a value of the `impl` argument of the fp32-pointer calls */

#define GLB_ACT_NONE 0
#define GLB_ACT_LRELU 1   /* leaky ReLU with `slope` (slope 0 = ReLU) */

typedef void* glb_stream_t; /* cudaStream_t */

const char* glb_last_error(void);
int glb_version(void);
/* 1 if the tcgen05 path can run on the current device (compute capability 10.x), else 0. */
int glb_tc_available(void);

/* ---- dense convolution family (stride 1) ------------------------------------------------------ *
 * Replaces nn.Conv2d inside Conv2dEx (utils/custom_layers.py:166-169, 202-211) with the equalized-LR
 * runtime scale folded in:   y = act( alpha * conv(x, w) + bias_scale * bias )
 *   alpha = wscale * lrmul, bias_scale = lrmul      (custom_layers.py:204-209)
 * fprop : x [N,H,W,Ci], w [Co,R,S,Ci] -> y [N,Ho,Wo,Co],  Ho = H + 2*pad - R + 1
 * dgrad : gy [N,Ho,Wo,Co], w -> gx [N,H,W,Ci] = alpha * conv_transpose(gy, w)   (aten convolution_backward, input)
 * wgrad : x, gy -> gw [Co,R,S,Ci] = alpha * sum_{n,ho,wo} gy (x) x              (aten convolution_backward, weight)
 * {fprop, dgrad, wgrad} is closed under differentiation (conv is bilinear), which is what the R1 /
 * WGAN-GP double backward needs (aten _convolution_double_backward; resnetgan/learner.py:811-825).   */
int glb_conv2d_fprop(const float* x, const float* w, const float* bias, float* y,
                     int N, int H, int W, int Ci, int Co, int R, int S, int pad,
                     float alpha, float bias_scale, int act, float slope, int impl, glb_stream_t stream);
/* `wt` = the flipped/transposed weights written by glb_conv2d_weight_transpose: required by GLB_IMPL_TF32 (the
 * tensor-core dgrad is the fprop kernel run over gy with wt), ignored (may be NULL) by GLB_IMPL_FP32, which reads `w`. */
int glb_conv2d_dgrad(const float* gy, const float* w, const float* wt, float* gx,
                     int N, int H, int W, int Ci, int Co, int R, int S, int pad,
                     float alpha, int impl, glb_stream_t stream);
int glb_conv2d_wgrad(const float* x, const float* gy, float* gw,
                     int N, int H, int W, int Ci, int Co, int R, int S, int pad,
                     float alpha, int impl, glb_stream_t stream);
/* 1 if the GLB_IMPL_TF32 kernel of `kind` (0 fprop, 1 dgrad, 2 wgrad) covers this shape (arguments as in the calls above). */
int glb_conv2d_tc_covers(int kind, int N, int H, int W, int Ci, int Co, int R, int S, int pad);
/* workspace-free weight re-layout used by the tensor-core dgrad: wt[Ci][R][S][Co] (taps flipped). */
int glb_conv2d_weight_transpose(const float* w, float* wt, int Co, int R, int S, int Ci, glb_stream_t stream);

/* ---- bf16-operand variant of the convolution family (opt-in: set_conv_impl("bf16")) ---------------------------- *
 * Same contract as above with the two GEMM operands given as bf16 copies (round to nearest even, made by
 * glb_cvt_f32_bf16 from the fp32 tensors; the weight copies once per optimiser step), fp32 accumulation, bias,
 * activation and output: tcgen05 kind::f16 runs at twice the kind::tf32 rate.  Replaces the same ATen call sites as
 * glb_conv2d_fprop / _dgrad / _wgrad (utils/custom_layers.py:202-211) under a looser stated tolerance.
 * dgrad takes the bf16 copy of the flipped / transposed weights (glb_conv2d_weight_transpose, then glb_cvt_f32_bf16). */
int glb_cvt_f32_bf16(const float* x, void* y_bf16, int64_t n, glb_stream_t stream);
int glb_conv2d_bf16_covers(int kind, int N, int H, int W, int Ci, int Co, int R, int S, int pad);
int glb_conv2d_fprop_bf16(const void* x_bf16, const void* w_bf16, const float* bias, float* y,
                          int N, int H, int W, int Ci, int Co, int R, int S, int pad,
                          float alpha, float bias_scale, int act, float slope, glb_stream_t stream);
int glb_conv2d_dgrad_bf16(const void* gy_bf16, const void* wt_bf16, float* gx,
                          int N, int H, int W, int Ci, int Co, int R, int S, int pad, float alpha, glb_stream_t stream);
int glb_conv2d_wgrad_bf16(const void* x_bf16, const void* gy_bf16, float* gw,
                          int N, int H, int W, int Ci, int Co, int R, int S, int pad, float alpha, glb_stream_t stream);

/* ---- nearest-neighbour 2x upsample folded into the 3x3 convolution that follows it (TF32 tensor-core path) ------ *
 * Replaces `nn.Upsample(scale_factor=2)` + `Conv2dEx(ks=3, padding=1)` of the generator blocks
 * (stylegan/architectures.py:155-156, 331; progan/architectures.py:209-224; resnetgan/learner.py:154-158) by ONE
 * launch on the LOW-resolution map: each of the four output phases (dy, dx) of y[2i+dy, 2j+dx] is a 2x2 convolution of
 * x with pre-summed taps -- 4/9 of the multiply-adds and no upsampled copy in HBM.  Same for the data gradient
 * (= upsample2x_bwd(dgrad)) and the weight gradient (= wgrad(upsample2x(x), gy)).
 * x [N,H,W,Ci], y / gy [N,2H,2W,Co], w / gw [Co,3,3,Ci];
 * wp [4*Co,2,2,Ci] (forward) and wt [Ci,16,Co] (data gradient) are written by glb_upconv_weights (either may be NULL);
 * gwp [Co,16,Ci] is scratch of glb_upconv_wgrad.  kind: 0 fprop, 1 dgrad, 2 wgrad. */
int glb_upconv_covers(int kind, int N, int H, int W, int Ci, int Co);
int glb_upconv_weights(const float* w, float* wp, float* wt, int Co, int Ci, glb_stream_t stream);
int glb_upconv_fprop(const float* x, const float* wp, const float* bias, float* y, int N, int H, int W, int Ci, int Co,
                     float alpha, float bias_scale, int act, float slope, glb_stream_t stream);
int glb_upconv_dgrad(const float* gy, const float* wt, float* gx, int N, int H, int W, int Ci, int Co,
                     float alpha, glb_stream_t stream);
int glb_upconv_wgrad(const float* x, const float* gy, float* gwp, float* gw, int N, int H, int W, int Ci, int Co,
                     float alpha, glb_stream_t stream);

/* ---- 2x2 average pool folded into the 3x3 convolution in front of it (TF32 tensor-core path) ---------------------- *
 * Replaces `Conv2dEx(ks=3, padding=1)` + `nn.AvgPool2d(2, 2)` (+ Conv2dBias + LeakyReLU) of the discriminator blocks
 * (progan/architectures.py:267-284; resnetgan/learner.py:160-164) by ONE launch that writes the LOW-resolution map:
 * avgpool(conv3x3(x)) is a stride-2 4x4 convolution = 0.25 * U^T C_w, the adjoint structure of glb_upconv_* -- four
 * strided input phases x 2x2 pre-summed taps, 4/9 of the multiply-adds, no full-resolution intermediate -- and runs on
 * the same kernels with the flipped / transposed weights.
 * x / gx [N,2H,2W,Ci], y / gy [N,H,W,Co] (H, W = output dims), w / gw [Co,3,3,Ci];
 * y = act(0.25 * alpha * pooled conv + bias_scale * bias);  wt [Co,16,Ci] (forward) and wp [4*Ci,2,2,Co] (data gradient)
 * come from glb_downconv_weights (either may be NULL); gwp [Ci,16,Co] is scratch of the wgrad. */
int glb_downconv_covers(int kind, int N, int H, int W, int Ci, int Co);
int glb_downconv_weights(const float* w, float* wp, float* wt, int Co, int Ci, glb_stream_t stream);
int glb_downconv_fprop(const float* x, const float* wt, const float* bias, float* y, int N, int H, int W, int Ci, int Co,
                       float alpha, float bias_scale, int act, float slope, glb_stream_t stream);
int glb_downconv_dgrad(const float* gy, const float* wp, float* gx, int N, int H, int W, int Ci, int Co,
                       float alpha, glb_stream_t stream);
int glb_downconv_wgrad(const float* x, const float* gy, float* gwp, float* gw, int N, int H, int W, int Ci, int Co,
                       float alpha, glb_stream_t stream);

/* bf16-operand variants of glb_upconv_* / glb_downconv_* (opt-in: set_conv_impl("bf16")): activations / gradients and the
 * re-laid-out weights (wp / wt of glb_upconv_weights / glb_downconv_weights) given as bf16 copies (glb_cvt_f32_bf16), fp32
 * accumulation, bias, activation and outputs; channel counts multiples of 64. */
int glb_upconv_bf16_covers(int kind, int N, int H, int W, int Ci, int Co);
int glb_upconv_fprop_bf16(const void* x_bf16, const void* wp_bf16, const float* bias, float* y, int N, int H, int W, int Ci, int Co,
                          float alpha, float bias_scale, int act, float slope, glb_stream_t stream);
int glb_upconv_dgrad_bf16(const void* gy_bf16, const void* wt_bf16, float* gx, int N, int H, int W, int Ci, int Co,
                          float alpha, glb_stream_t stream);
int glb_upconv_wgrad_bf16(const void* x_bf16, const void* gy_bf16, float* gwp, float* gw, int N, int H, int W, int Ci, int Co,
                          float alpha, glb_stream_t stream);
int glb_downconv_bf16_covers(int kind, int N, int H, int W, int Ci, int Co);
int glb_downconv_fprop_bf16(const void* x_bf16, const void* wt_bf16, const float* bias, float* y, int N, int H, int W, int Ci, int Co,
                            float alpha, float bias_scale, int act, float slope, glb_stream_t stream);
int glb_downconv_dgrad_bf16(const void* gy_bf16, const void* wp_bf16, float* gx, int N, int H, int W, int Ci, int Co,
                            float alpha, glb_stream_t stream);
int glb_downconv_wgrad_bf16(const void* x_bf16, const void* gy_bf16, float* gwp, float* gw, int N, int H, int W, int Ci, int Co,
                            float alpha, glb_stream_t stream);

/* ---- RGB 1x1 convolutions (3 <-> C channels; pure bandwidth) ---------------------------------- *
 * fromRGB (progan/architectures.py:286-292) and toRGB (stylegan/architectures.py:338-341).
 * In all three the weight element (j = rgb channel, c = feature channel) lives at w[j*ws_j + c*ws_c], so the
 * same kernels serve fromRGB ([C,3]: ws_j=1, ws_c=3) and toRGB ([3,C]: ws_j=C, ws_c=1) and their gradients.
 * expand  : img NCHW [N,3,H*(1+pool),W*(1+pool)] (pool=1: 2x2 average-pooled first = the fade-in skip branch,
 *           progan/architectures.py:312) -> y NHWC [N,H,W,C] = act(alpha*img.w + bias_scale*bias)
 * contract: x NHWC [N,H,W,C], w (element (j,c) at w[j*ws_j + c*ws_c]) -> img NCHW [N,3,H,W]
 *           = alpha * sum_c x*w + bias_scale*bias ; if pool != 0 the result is instead scattered as the
 *           adjoint of the 2x2 average pool into an image of [N,3,2H,2W]
 * wgrad   : img NCHW (pooled if pool), g NHWC -> gw (element (j,c) at gw[j*ws_j + c*ws_c]) = alpha * sum_p img_j g_c */
int glb_rgb_expand(const float* img, const float* w, int ws_j, int ws_c, const float* bias, float* y,
                   int N, int H, int W, int C, int pool, float alpha, float bias_scale, int act, float slope,
                   glb_stream_t stream);
int glb_rgb_contract(const float* x, const float* w, int ws_j, int ws_c, const float* bias, float* img,
                     int N, int H, int W, int C, int pool, float alpha, float bias_scale, glb_stream_t stream);
int glb_rgb_wgrad(const float* img, const float* g, float* gw, int ws_j, int ws_c,
                  int N, int H, int W, int C, int pool, float alpha, glb_stream_t stream);   /* gw zeroed by caller */
/* out[c] += scale * sum_{n,h,w} img[n,c,h,w] (NCHW; toRGB bias gradient); out zeroed by caller. */
int glb_plane_sum(const float* img, float* out, int N, int C, int64_t HW, float scale, glb_stream_t stream);

/* ---- small-M linear (LinearEx, utils/custom_layers.py:230-291) -------------------------------- *
 * y[M,Nout] = act( alpha * x[M,K] . w[Nout,K]^T + bias_scale * bias )                               */
int glb_linear_fwd(const float* x, const float* w, const float* bias, float* y, int M, int K, int Nout,
                   float alpha, float bias_scale, int act, float slope, glb_stream_t stream);
int glb_linear_dgrad(const float* gy, const float* w, float* gx, int M, int K, int Nout, float alpha, glb_stream_t stream);
int glb_linear_wgrad(const float* x, const float* gy, float* gw, int M, int K, int Nout, float alpha, glb_stream_t stream);

/* ---- elementwise / reduction glue ------------------------------------------------------------- */
/* y = act(x + bias_scale*bias[c]) over [P,C] rows (Conv2dBias + LeakyReLU, custom_layers.py:213-226). */
int glb_bias_act_fwd(const float* x, const float* bias, float* y, int64_t P, int C, float bias_scale, int act, float slope, glb_stream_t stream);
/* gx = gy * act'(y) (sign of the saved OUTPUT y); if gbias != NULL also gbias[c] += bias_scale * sum_p gx.
 * gbias must be zeroed by the caller.  aten leaky_relu_backward (+ bias reduction). */
int glb_act_bwd(const float* gy, const float* y, float* gx, float* gbias, int64_t P, int C, float bias_scale, int act, float slope, glb_stream_t stream);
/* out[c] += scale * sum_p x[p,c]   (bias gradient without activation); out zeroed by caller. */
int glb_colsum(const float* x, float* out, int64_t P, int C, float scale, glb_stream_t stream);
/* y = a*alpha + b*beta (elementwise, n elements): fade-in blend (progan/architectures.py:312-313). */
int glb_axpby(const float* a, const float* b, float* y, int64_t n, float alpha, float beta, glb_stream_t stream);
/* y = x*scale */
int glb_scale(const float* x, float* y, int64_t n, float scale, glb_stream_t stream);
/* y = x * scale * s_dev[0]   (s_dev: device scalar, e.g. the upstream gradient of a scalar loss) */
int glb_scale_by(const float* x, const float* s_dev, float* y, int64_t n, float scale, glb_stream_t stream);
/* out[0] += scale * sum(x*x) ; out zeroed by caller (R1 / R2 penalty reduction, resnetgan/learner.py:825). */
int glb_sumsq(const float* x, float* out, int64_t n, float scale, glb_stream_t stream);
/* WGAN-GP (resnetgan/learner.py:817-823): g NCHW [N,C,HW]; per pixel nrm = sqrt(sum_c g^2);
 * fwd: out[0] += scale * sum_{n,p} (nrm - gamma)^2 ; out zeroed by caller
 * bwd: gg = g * (2*scale*(nrm-gamma)/nrm) * s_dev[0]   (0 where nrm == 0) */
int glb_gp_norm_fwd(const float* g, float* out, int N, int C, int64_t HW, float gamma, float scale, glb_stream_t stream);
int glb_gp_norm_bwd(const float* g, const float* s_dev, float* gg, int N, int C, int64_t HW, float gamma, float scale, glb_stream_t stream);
/* out[n,:] = eps[n]*a[n,:] + (1-eps[n])*b[n,:]  (WGAN-GP interpolation, resnetgan/learner.py:794-796) */
int glb_interp_rows(const float* a, const float* b, const float* eps, float* out, int N, int64_t per_row, glb_stream_t stream);

/* PixelNorm2d (utils/custom_layers.py:81-86) over the channel axis of [P,C] rows. */
int glb_pixelnorm_fwd(const float* x, float* y, int64_t P, int C, float eps, glb_stream_t stream);
int glb_pixelnorm_bwd(const float* gy, const float* x, float* gx, int64_t P, int C, float eps, glb_stream_t stream);

/* 3x3 binomial FIR [1 2 1]x[1 2 1]/16, zero pad 1, depthwise, NHWC (get_blur_op, custom_layers.py:36-53).
 * Symmetric + zero padded => self-adjoint: forward, backward and double-backward are this same kernel. */
int glb_blur3x3(const float* x, float* y, int N, int H, int W, int C, glb_stream_t stream);

/* nearest x2 upsample NHWC (nn.Upsample, resnetgan/learner.py:154-158) and its adjoint (2x2 sum). */
int glb_upsample2x_fwd(const float* x, float* y, int N, int H, int W, int C, glb_stream_t stream);   /* x [N,H,W,C] -> y [N,2H,2W,C] */
int glb_upsample2x_bwd(const float* gy, float* gx, int N, int H, int W, int C, glb_stream_t stream); /* gy [N,2H,2W,C] -> gx [N,H,W,C] */
/* 2x2 average pool (+ bias + act): D downsampling tail, progan/architectures.py:267-284.
 * fwd: x [N,H,W,C] -> y [N,H/2,W/2,C] = act(avg(x) + bias_scale*bias); bias may be NULL (plain pool, act none).
 * bwd: gx [N,H,W,C] = 0.25 * (gy * act'(y)) broadcast;  gbias (zeroed by caller, may be NULL) += bias_scale*sum */
int glb_pool_bias_act_fwd(const float* x, const float* bias, float* y, int N, int H, int W, int C, float bias_scale, int act, float slope, glb_stream_t stream);
int glb_pool_bias_act_bwd(const float* gy, const float* y, float* gx, float* gbias, int N, int H, int W, int C, float bias_scale, int act, float slope, glb_stream_t stream);

/* StyleGAN G-layer epilogue (stylegan/architectures.py:112-119, 255/263/333, 460-462/524-526):
 *   t = lrelu(x + noise_weight[c]*noise[n,h,w] + bias[c]);  out = (t-mu)*rstd * (ys+1) + yb
 * with per-(n,c) mean / biased variance over H*W (nn.InstanceNorm2d eps=1e-8), style [N,2C] = [ys;yb].
 * stats (2*N*C floats: mu plane, rstd plane per sample) is written by fwd and consumed by bwd.  noise/noise_weight may be NULL together.
 * bwd writes gx, gstyle [N,2C] and accumulates g_noise_weight[C], g_bias[C] (zeroed by caller).
 * `work` is caller-owned scratch: style_epilogue_work_floats(N,H,W,C) floats. */
int64_t glb_style_epilogue_work_floats(int N, int H, int W, int C);
int glb_style_epilogue_fwd(const float* x, const float* noise, const float* noise_weight, const float* bias,
                           const float* style, float* out, float* stats, float* work,
                           int N, int H, int W, int C, float slope, float eps, glb_stream_t stream);
int glb_style_epilogue_bwd(const float* gout, const float* x, const float* noise, const float* noise_weight,
                           const float* bias, const float* style, const float* stats,
                           float* gx, float* gstyle, float* g_noise_weight, float* g_bias, float* work,
                           int N, int H, int W, int C, float slope, glb_stream_t stream);

/* minibatch-stddev concat (concat_mbstd_layer, utils/custom_layers.py:117-140): x [N,H,W,C] -> y [N,H,W,C+1].
 * G groups of `group` consecutive samples (N = G*group); sd [G] = mean_{c,h,w} sqrt(var_unbiased + 1e-8).
 * bwd   : gx = gy[..., :C] + d(sd)/dx * sum_{n in group,h,w} gy[..., C]
 * bwdbwd: given v (cotangent of gx) -> ggx (grad wrt x) and ggy (grad wrt gy), the double backward that R1
 *         sends back into the forward graph (SURVEY.md section 7 'Hard parts'). */
int glb_mbstd_fwd(const float* x, float* y, int N, int H, int W, int C, int group, glb_stream_t stream);
int glb_mbstd_bwd(const float* gy, const float* x, float* gx, int N, int H, int W, int C, int group, glb_stream_t stream);
int glb_mbstd_bwdbwd(const float* v, const float* gy, const float* x, float* ggx, float* ggy,
                     int N, int H, int W, int C, int group, glb_stream_t stream);

/* image-space fade-in helpers (NCHW, 3 channels):
 * out = (1-alpha) * up2x(lo) + alpha * hi      (stylegan/architectures.py:482-487; lo [N,3,H/2,W/2], hi [N,3,H,W])
 * real-image fade (progan/learner.py:770-779): out = (1-alpha) * up2x(avgpool2(x)) + alpha * x, done on device. */
int glb_fade_up_blend(const float* lo, const float* hi, float* out, int N, int C, int H, int W, float alpha, glb_stream_t stream);
int glb_fade_up_blend_bwd(const float* gout, float* glo, float* ghi, int N, int C, int H, int W, float alpha, glb_stream_t stream);
int glb_fade_real(const float* x, float* out, int N, int C, int H, int W, float alpha, glb_stream_t stream);
/* The same blends with the coefficients read ON THE DEVICE from `coef` = float[2] {alpha, 1 - alpha}: alpha moves every
 * iteration of a fade-in phase (progan/learner.py:951-952), and a value baked into a captured CUDA graph could not.
 * glb_axpby_dev: y = a*coef[ia] + b*coef[ib] (b may be NULL: y = a*coef[ia]). */
int glb_axpby_dev(const float* a, const float* b, float* y, int64_t n, const float* coef, int ia, int ib, glb_stream_t stream);
int glb_fade_up_blend_dev(const float* lo, const float* hi, float* out, int N, int C, int H, int W, const float* coef, glb_stream_t stream);
int glb_fade_up_blend_bwd_dev(const float* gout, float* glo, float* ghi, int N, int C, int H, int W, const float* coef, glb_stream_t stream);
int glb_fade_real_dev(const float* x, float* out, int N, int C, int H, int W, const float* coef, glb_stream_t stream);

/* ---- real-image input pipeline ------------------------------------------------------------- *
 * Replaces the host-side torchvision chain every real sample goes through in the reference -- Resize(curr_res, PIL BOX)
 * -> ToTensor -> Normalize(mean, std) (data_config.py:312-342; Resize rewritten per resolution, progan/learner.py:1099-1112)
 * -- bit-exactly (Pillow 12.2 src/libImaging/Resample.c: two 8-bit passes with 22-bit fixed-point weights, the horizontal
 * one rounded to uint8 first; then u8 -> float, /255, -mean, /std in fp32).
 * glb_box_resize_ksize / _tables: HOST helpers, Pillow's precompute_coeffs + normalize_coeffs_8bpc for one axis (BOX filter):
 *   bounds int32 [out][2] = first source index, count ; kk int32 [out][ksize] fixed-point weights.  The caller uploads them.
 * glb_u8_box_resize_normalize: src uint8 [src_images][Hs][Ws][3] (decoded RGB, HWC) -> dst fp32 [N][3][Ho][Wo];
 *   index int64 [N] picks the samples (NULL: the first N), flip uint8 [N] mirrors a sample horizontally after the resize
 *   (RandomHorizontalFlip, data_config.py:331-332; NULL: none); xb/xk/yb/yk = device copies of the tables of the two axes;
 *   mean3/std3 are HOST pointers to 3 floats. */
int glb_box_resize_ksize(int in_size, int out_size);
int glb_box_resize_tables(int in_size, int out_size, int32_t* bounds, int32_t* kk);
int glb_u8_box_resize_normalize(const uint8_t* src, int64_t src_images, const int64_t* index, const uint8_t* flip, float* dst,
                                int N, int Hs, int Ws, int Ho, int Wo, const int32_t* xb, const int32_t* xk, int xks,
                                const int32_t* yb, const int32_t* yk, int yks, const float* mean3, const float* std3,
                                glb_stream_t stream);

/* logit losses (progan/learner.py:791-812, 883-896).  kind: 0 = wgan, 1 = nonsaturating, 2 = minimax.
 * d-loss: loss[0] = L(d_gen, d_real) + eps_drift*mean(d_real^2); also writes dL/d d_gen, dL/d d_real.
 * g-loss: loss[0] = L(d_out); writes dL/d d_out. */
int glb_d_logit_loss(const float* d_gen, const float* d_real, float* loss, float* g_gen, float* g_real,
                     int n, int kind, float eps_drift, glb_stream_t stream);
int glb_g_logit_loss(const float* d_out, float* loss, float* g_out, int n, int kind, glb_stream_t stream);

/* w_ewma (stylegan/architectures.py:427-437): ewma[k] = mean_m w[m,k]*(1-beta) + ewma[k]*beta (init: beta=0). */
int glb_w_ewma(const float* w, float* ewma, int M, int K, float beta, glb_stream_t stream);

/* ---- fused multi-tensor Adam (+ EWMA generator) ----------------------------------------------- *
 * torch.optim.Adam semantics (utils/backprop_utils.py:109-120) on a flat table of tensors; optionally
 * also lagged = p*(1-ewma_beta) + lagged*ewma_beta in the same pass (progan/learner.py:909-916).
 * ptrs: device array of 5*T pointers [p, g, m, v, lagged(or NULL)] per tensor (g == NULL: EWMA only, no Adam);
 * sizes: device int64[T]; ewma_mode 0 = off, 1 = running, 2 = first step (lagged aliases p, progan/learner.py:472);
 * bc1 = 1-beta1^step, bc2 = 1-beta2^step are read from the device vector `hyper` (float[4]: lr, bc1, bc2, step) so
 * the launch is CUDA-graph replayable with a changing step count; glb_adam_hyper_advance does step += 1 and
 * refreshes bc1/bc2 on the device (what torch.optim.Adam computes on the host each step). */
int glb_adam_hyper_advance(float* hyper, float beta1, float beta2, glb_stream_t stream);
int glb_adam_ewma_multi(const void* const* ptrs, const int64_t* sizes, int T, int64_t max_size,
                        const float* hyper, float beta1, float beta2, float eps, float wd,
                        float ewma_beta, int ewma_mode, glb_stream_t stream);

/* ---- ResNet-GAN normalisation layers (resnetgan/resblocks.py:43-46 through NormalizeLayer, utils/custom_layers.py:100-106) -- *
 * Both are fused with the ReLU that always follows them in the blocks (act / slope as everywhere else; y is then the
 * POST-activation output and doubles as the mask for the backward).  `ws` is a caller-owned scratch of
 * glb_*_ws_doubles() doubles (block partial sums are combined in fp64).
 *
 * LayerNorm([C,H,W]) with elementwise affine on x [N][L], L = H*W*C in NHWC order; gamma/beta [L] in the SAME (H,W,C)
 * order (the module keeps its (C,H,W)-shaped parameter in that memory order); eps inside the rsqrt, biased variance.
 *   fwd    : y = act((x-mean_n)*rstd_n*gamma + beta); stats[n] = {mean, rstd}
 *   bwd    : g = gy*act'(y)*gamma; gx = rstd*(g - mean(g) - xhat*mean(g*xhat)); ggamma = sum_n gy*act'*xhat; gbeta = sum_n gy*act'
 *            (gx or the parameter pair may be NULL)
 *   bwdbwd : cotangent u of gx -> gradients w.r.t. gy, x and gamma: the second-order term WGAN-GP sends back through the
 *            discriminator's LayerNorms (resnetgan/learner.py:811-825; aten native_layer_norm_backward's own backward). */
int64_t glb_layernorm_ws_doubles(int N);
int glb_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* y, float* stats, void* ws,
                      int N, int64_t L, float eps, int act, float slope, glb_stream_t stream);
int glb_layernorm_bwd(const float* gy, const float* y, const float* x, const float* gamma, const float* stats,
                      float* gx, float* ggamma, float* gbeta, void* ws, int N, int64_t L, int act, float slope,
                      glb_stream_t stream);
int glb_layernorm_bwdbwd(const float* u, const float* gy, const float* y, const float* x, const float* gamma,
                         const float* stats, float* g_gy, float* g_x, float* g_gamma, void* ws, int N, int64_t L,
                         int act, float slope, glb_stream_t stream);
/* BatchNorm2d in training mode (batch statistics over the P = N*H*W rows of x [P][C], biased variance for the
 * normalisation, momentum update of running_mean / running_var with the unbiased variance, num_batches_tracked += 1;
 * any of the three running pointers may be NULL).  stats = {mean[C], rstd[C]}.  Generator side only -> first order. */
int64_t glb_batchnorm_ws_doubles(int C);
int glb_batchnorm_fwd(const float* x, const float* gamma, const float* beta, float* y, float* stats,
                      float* running_mean, float* running_var, int64_t* num_batches_tracked, void* ws,
                      int64_t P, int C, float eps, float momentum, int act, float slope, glb_stream_t stream);
int glb_batchnorm_bwd(const float* gy, const float* y, const float* x, const float* gamma, const float* stats,
                      float* gx, float* ggamma, float* gbeta, void* ws, int64_t P, int C, int act, float slope,
                      glb_stream_t stream);
/* nn.Tanh on the generator output (resnetgan/architectures.py:58, 96): y = tanh(x); gx = gy*(1-y^2). */
int glb_tanh_fwd(const float* x, float* y, int64_t n, glb_stream_t stream);
int glb_tanh_bwd(const float* gy, const float* y, float* gx, int64_t n, glb_stream_t stream);

/* grouped small linears: L layers with their own weights, layer l fed by ws[l] ([L][M][K]), one launch per direction (the 12
 * style affines of a StyleGAN generator pass, stylegan/architectures.py:460 / 524).  `tab` = device array of L rows of
 * 8 x int64: {weight ptr [nout][K], bias ptr or 0, nout, column offset off, first block fwd, first block dgrad, first block
 * wgrad, (alpha, bias_scale) as two packed floats}; blocks per layer = ceil(nout / glb_glinear_rows(which)).
 *   fwd   : y chunk of layer l = [M][nout_l] at y + M*off_l :  alpha_l * ws[l] . W_l^T + bias_scale_l * b_l
 *   dgrad : g_all [M][G] (layer l at columns off_l..) -> g_ws [L][M][K]
 *   wgrad : -> gw_all [G][K] (layer l = rows off_l..), gb_all [G] */
int glb_glinear_rows(int which);
int glb_glinear_fwd(const float* ws, const void* tab, float* y, int L, int M, int K, int blocks, glb_stream_t stream);
int glb_glinear_dgrad(const float* g_all, int G, const void* tab, float* g_ws, int L, int M, int K, int blocks, glb_stream_t stream);
int glb_glinear_wgrad(const float* ws, const float* g_all, int G, const void* tab, float* gw_all, float* gb_all,
                      int L, int M, int K, int blocks, glb_stream_t stream);

/* backward of "activation, then blur" (the D block's conv + bias + lrelu followed by the pre-blur of the next conv,
 * progan/architectures.py:267-284) in one pass: g = blur3x3(gz) * act'(ymask); gbias (zero-initialised by the caller, may
 * be NULL) += bias_scale * column sums of g.  Replaces glb_blur3x3 + glb_act_bwd. */
int glb_blur_act_bwd(const float* gz, const float* ymask, float* g, float* gbias, int N, int H, int W, int C,
                     float bias_scale, int act, float slope, glb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GANLAB_B200_H_ */
