/* pvg_b200.h - C ABI of the B200-native CADDY hot path (libpvg_b200.so).
 *
 * The reference (willi-menapace/PlayableVideoGeneration) has no native/FFI layer: every op on its hot path is a
 * torch.nn / torch.nn.functional call.  Each entry point below therefore replaces a *library op
 * call site* of the reference; the file:line next to it is that call site.  The Python host side
 * (playablevideogeneration_b200/ops.py) binds these with ctypes and wraps them in torch.autograd.Function.
 *
 * Conventions
 *   - all pointers are DEVICE pointers into caller-owned storage (the callee never allocates, frees or retains);
 *   - activations are NHWC fp32 ("channels_last" physical layout of a logical NCHW tensor);
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on it and re-entrant per stream;
 *   - return 0 on success, negative on error; pvg_last_error() returns the message of the last failure
 *     on the calling thread;
 *   - "lo" companions: in the fp32-equivalent 3xTF32 mode (nprod == 3) a tensor-core operand X is consumed as
 *     X_hi + X_lo with X_lo = X - tf32(X); pvg_split_tf32 produces X_lo.
 *   - nprod == 2 is the same split with the two correction products (2^-11 of the result) evaluated as 16-bit MMAs
 *     (kind::f16, K = 16): the "lo" argument is then a PAIR of 16-bit planes [2][numel] =
 *     { f16((X - trunc_tf32(X)) * 2^12), f16(X) } for activations (pvg_split_16 / pvg_act_bwd_split_16) and
 *     { f16(W_lo * 2^12), f16(W_hi) } for packed weights (pvg_pack_16x2), f16 = bf16 or fp16 (pvg_conv_desc.corr_fmt; both
 *     operands of one conv use the same format).  fp16 holds the tf32 mantissa of a weight exactly (no error that is
 *     coherent over the batch) and suits O(1) activations; bf16 has fp32's exponent range (gradients).
 */
#ifndef PVG_B200_H_
#define PVG_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PVG_ACT_NONE    0
#define PVG_ACT_LRELU   1   /* leaky_relu(slope)  - model/layers/residual_block.py:30, same_block.py:33, up_block.py:40 */
#define PVG_ACT_RELU    2   /* torchvision VGG19 features, model/layers/vgg.py:16 */
#define PVG_ACT_TANH    3   /* model/layers/final_block.py:27 */
#define PVG_ACT_SIGMOID 4   /* model/main_model/representation_network.py:55 */
#define PVG_ACT_LSTM    5   /* internal: the fused ConvLSTM cell epilogue of pvg_convlstm_step */

#define PVG_CORR_BF16 0
#define PVG_CORR_FP16 1
#define PVG_CORR_FP16_ALL 2 /* fp16 planes { f16((X - f16(X)) * 2^12), f16(X) }: the pair alone carries X to 22 bits, so ALL three
                               products of the split run as kind::f16 MMAs on the planes and the convolution never reads the
                               fp32 tensor (x / w may be NULL).  For O(1) operands (forward activations, weights); gradients keep
                               the TF32 main product with bf16 corrections (fp16 has no range for 1e-6..1e-12). */

#define PVG_ALGO_AUTO 0
#define PVG_ALGO_SIMT 1     /* fp32 CUDA-core implicit GEMM (any shape) */
#define PVG_ALGO_UMMA 2     /* tcgen05 / TMEM / TMA implicit GEMM (Cin % 4 == 0, % 8 with 16-bit correction planes) */
#define PVG_ALGO_UMMA_PERSISTENT 3   /* force the persistent-tile variant of the split-product kernels (the default for 64/128-wide
                                        tiles; PVG_PERSISTENT=0 in the environment selects one tile per CTA instead) */

typedef struct pvg_conv_desc {
  int32_t N, H, W;          /* output == input spatial size (all convs on the path are stride 1, "same" padding) */
  int32_t Cin;              /* physical channels of x (tensor-core path: any multiple of 4, of 8 with 16-bit planes; the K
                               loop runs over Cin rounded up to 32, TMA zero-fills, and w is packed with that count) */
  int32_t Cout;             /* physical channels of y */
  int32_t R, S, pad;        /* kernel height/width and zero padding (3,3,1 | 1,1,0 | 7,7,3) */
  int32_t act;              /* PVG_ACT_* fused into the epilogue (after bias) */
  float   slope;            /* leaky slope for PVG_ACT_LRELU */
  int32_t algo;             /* PVG_ALGO_* */
  int32_t nprod;            /* 1 = single TF32 product, 3 = 3xTF32 (fp32-equivalent), 2 = TF32 + 2 bf16 corrections
                               (fp32-equivalent, see above); SIMT ignores it */
  int32_t corr_fmt;         /* PVG_CORR_BF16 / PVG_CORR_FP16 / PVG_CORR_FP16_ALL: format of the 16-bit planes (nprod == 2) */
} pvg_conv_desc;

const char* pvg_last_error(void);
int pvg_version(void);
/* 1 if the current device is sm_100 (tcgen05 path usable) */
int pvg_has_umma(void);

/* ---- convolution: replaces nn.Conv2d forward (cuDNN) at residual_block.py:52,57,63, same_block.py:38,
 *      up_block.py:37, final_block.py:26, convolutional_lstm_cell.py:92-95, representation_network.py:41,
 *      model.py:413, vgg.py:48-52 ------------------------------------------------------------------------------ */
/* w: [Cout][R][S][Cin] (K-major pack from pvg_pack_conv_weight); w_lo may be NULL when nprod == 1 / SIMT.
 * y[n,h,w,co] = act(bias[co] + sum_{r,s,ci} x[n,h+r-pad,w+s-pad,ci] * w[co,r,s,ci]).   bias may be NULL. */
int pvg_conv2d_fwd(const pvg_conv_desc* d, const float* x, const void* x_lo, const float* w, const void* w_lo,
                   const float* bias, float* y, void* stream);
/* Image-facing 3 -> 64 channel 3x3 convolution (VGG19 conv1_1, vgg.py:48-52) in fp32 on the CUDA cores, output-bound: coalesced
 * stores of y and, when y_planes != NULL, of the PVG_CORR_FP16_ALL plane pair of y that conv1_2 consumes.
 * x: [N,H,W,3]; w: [64][3][3][3] (the K-major pack with CinK = 3); d->act / d->slope as for pvg_conv2d_fwd. */
int pvg_conv2d_stem_planes(const pvg_conv_desc* d, const float* x, const float* w, const float* bias, float* y, void* y_planes,
                           void* stream);
/* The same convolution with x and w given ONLY as fp16 plane pairs (PVG_CORR_FP16_ALL: pvg_split_16 / pvg_pack_16x2 with that
 * format, or the y_planes of a previous call): all three products of the split run as kind::f16 MMAs; 3x3 convolutions reuse
 * one haloed shared-memory tile for their 9 taps; persistent CTAs (CTA pairs on large problems).  d->Cin % 8 == 0.
 * y_planes (optional, Cout % 8 == 0): fp16 [2][N*H*W*Cout], the plane pair of y, written by the epilogue so that the next
 * convolution needs no separate split pass (vgg.py:48-52 conv -> relu -> conv chains, residual_block.py:52-57). */
int pvg_conv2d_fwd_planes(const pvg_conv_desc* d, const void* x_planes, const void* w_planes, const float* bias, float* y,
                          void* y_planes, const float* out_scale, double* bn_sums, int bn_groups, uint32_t* amax_out, void* stream);
/* amax_out (optional, one zero-initialised uint32): receives the bit pattern of max|y| - when y is a gradient, the next backward
 * kernel up the chain takes its power-of-two scale from it and needs no pvg_amax pass over y. */
/* bn_sums (optional; double[bn_groups][2][Cout], zero-initialised by the caller; d->act == PVG_ACT_NONE): the epilogue also
 * accumulates the per-channel sum and sum of squares of y per batch group - the statistics of the BatchNorm that follows the
 * convolution (residual_block.py:52-58, up_block.py:37-38), i.e. pvg_bn_stats without its pass over y. */
/* out_scale (optional, one float in DEVICE memory): the accumulator is multiplied by it before bias / activation - the 1 / S of
 * a scaled gradient operand (pvg_split_16_scaled) when the call computes a data gradient (w_planes = the flipped pack). */
/* Weight gradient from plane pairs only: x_planes = PVG_CORR_FP16_ALL planes of x [N,H,W,d->Cin] (the forward operand, reused),
 * g_planes = scaled planes of dY [N,H,W,d->Cout] (d->Cout % 8 == 0), out_scale = its 1 / S.  Three kind::f16 MMAs per product,
 * MN-major operands, split-K; scratch / dw_oihw / accumulate as for pvg_conv2d_wgrad_umma. */
int pvg_conv2d_wgrad_planes(const pvg_conv_desc* d, int Cin_logical, const void* x_planes, const void* g_planes,
                            const float* out_scale, float* scratch, float* dw_oihw, int accumulate, void* stream);
/* dw_oihw == NULL defers the unpack: a weight that is used at every time step (autograd sums over its uses,
 * training/trainer.py:584-587) lets all its uses accumulate their split-K partials in ONE scratch and unpacks once: */
/* Weight gradient of a 3x3 convolution with 16 / 32 channels on both sides, in fp32 on the CUDA cores (the GEMM is tiny, K is every
 * pixel: the first encoder stage, residual_block.py:52-58).  x: [N,H,W,d->Cin], g = dY: [N,H,W,d->Cout], both fp32; scratch /
 * dw_oihw / accumulate exactly as for pvg_conv2d_wgrad_planes (same packed scratch, so uses of one weight may mix both). */
int pvg_conv2d_wgrad_small(const pvg_conv_desc* d, int Cin_logical, const float* x, const float* g, float* scratch, float* dw_oihw,
                           int accumulate, void* stream);
int pvg_unpack_dw(const float* scratch, int Cout, int Cin_logical, int R, int S, int CinPhys, float* dw_oihw, int accumulate,
                  void* stream);
/* OIHW [Cout][Cin][R][S] -> forward pack [Cout][R][S][CinK] and data-gradient pack [CinRows][R][S][CoutK] with flipped
 * taps (dgrad = pvg_conv2d_fwd(dy, bwd pack)).  CinRows >= Cin: physical channels of the activation (zero-padded concat
 * buffers); CinK >= CinRows, CoutK >= Cout: K-side channel counts of the consumer (tensor-core kernels: rounded up to 32,
 * zero filled; SIMT: the unpadded counts).  *_lo = w - tf32(w); *_hi = tf32(w) when round_hi != 0 (tensor-core
 * consumers) else w (SIMT consumers); any output pointer may be NULL. */
int pvg_pack_conv_weight(const float* w_oihw, int Cout, int Cin, int R, int S, int CinRows, int CinK, int CoutK, int round_hi,
                         float* fwd_hi, float* fwd_lo, float* bwd_hi, float* bwd_lo, void* stream);
/* the same with CoutRows >= Cout rows in the forward pack (zero rows): the convolution then writes an output physically padded
 * to CoutRows channels */
int pvg_pack_conv_weight_ex(const float* w_oihw, int Cout, int Cin, int R, int S, int CinRows, int CinK, int CoutK, int CoutRows,
                            int round_hi, float* fwd_hi, float* fwd_lo, float* bwd_hi, float* bwd_lo, void* stream);
/* weight gradient (cudnnConvolutionBackwardFilter): dw_oihw[co][ci][r][s] += sum_pixels dy * x ; x has CinP
 * physical channels of which the first Cin are real.  dw must be zero-initialised by the caller. */
int pvg_conv2d_wgrad(const pvg_conv_desc* d, int Cin_logical, const float* x, const float* dy, float* dw_oihw,
                     void* stream);
/* Tensor-core weight gradient (tcgen05, MN-major tf32 operands, split-K over pixel patches).  x: [N,H,W,d->Cin]
 * (d->Cin % 32 == 0), g = dY: [N,H,W,d->Cout] (d->Cout % 4 == 0), *_lo their 3xTF32 residual planes (nprod == 3, else
 * NULL).  scratch: float[Cout*R*S*roundup(Cin,32)], zero-initialised by the caller.
 * dw_oihw[co][ci<Cin_logical][r][s] = result (+ its previous content when accumulate != 0). */
int pvg_conv2d_wgrad_umma(const pvg_conv_desc* d, int Cin_logical, const float* x, const void* x_lo, const float* g,
                          const void* g_lo, float* scratch, float* dw_oihw, int accumulate, void* stream);
/* out[c] = sum over M rows of x[M][C]  (bias gradient); scratch: double[C] */
int pvg_channel_sum(const float* x, int64_t M, int C, double* scratch, float* out, void* stream);
/* hi != NULL: hi = rna_tf32(x), lo = x - hi.  hi == NULL: lo = x - trunc_tf32(x) (x itself then serves as the hi
 * operand; valid when the tensor core truncates raw fp32 inputs - probed at start-up by the host side). */
int pvg_split_tf32(const float* x, float* hi, float* lo, int64_t n, void* stream);
/* planes: 16-bit [2][n] = { f16((x - trunc_tf32(x)) * 2^12), f16(x) }, fmt = PVG_CORR_BF16 | PVG_CORR_FP16 (saturating):
 * the correction operands of nprod == 2 */
int pvg_split_16(const float* x, void* planes, int64_t n, int fmt, void* stream);
/* the same for a packed weight given its tf32 hi / residual lo planes: planes = { f16(lo * 2^12), f16(hi) } */
int pvg_pack_16x2(const float* hi, const float* lo, void* planes, int64_t n, int fmt, void* stream);
/* g = dy * act'(y) for an activation fused in a conv epilogue (y is the activation OUTPUT) */
int pvg_act_bwd(const float* dy, const float* y, int act, float slope, float* g, int64_t n, void* stream);
/* the same, also emitting the 3xTF32 planes of g in the same pass (hi may be NULL: truncation mode, see pvg_split_tf32) */
int pvg_act_bwd_split(const float* dy, const float* y, int act, float slope, float* g, float* hi, float* lo, int64_t n,
                      void* stream);
/* ---- gradients as SCALED fp16 plane pairs (all-fp16 data / weight gradient kernels): fp16 has no range for raw gradients, so
 *      the tensor is multiplied by a power of two S with max|g| * S in [2^13, 2^14) before the PVG_CORR_FP16_ALL split; the
 *      consuming kernel multiplies its result by 1 / S (inv_scale, one float in device memory). --------------------------- */
/* *amax_bits (uint32, zero-initialised by the caller) = max(*amax_bits, bits of max|x|) */
int pvg_amax(const float* x, int64_t n, uint32_t* amax_bits, void* stream);
/* planes = PVG_CORR_FP16_ALL planes of x * S, *inv_scale = 1 / S, with S chosen from *amax_bits */
int pvg_split_16_scaled(const float* x, void* planes, int64_t n, const uint32_t* amax_bits, float* inv_scale, void* stream);
/* g = dy * act'(y) (g may be NULL) and the scaled planes of g; *amax_bits = max|dy| bounds max|g| (|act'| <= 1) */
int pvg_act_bwd_split_16_scaled(const float* dy, const float* y, int act, float slope, float* g, void* planes, int64_t n,
                                const uint32_t* amax_bits, float* inv_scale, void* stream);
/* The same two with a feature-matching L1 term folded in (vgg.py:41-56 taps + losses.py:465): the gradient that is pushed through
 * the activation is dy + tap_gout[n] / count * sign(y - target), n = the sample ([N][count] layout) - i.e. what
 * pvg_absdiff_mean_bwd(target, y, tap_gout) and an addition would have produced, without their passes over the feature map.
 * The scaled variant takes its scale from the bound *amax_bits + max|tap_gout| / count. */
int pvg_act_bwd_tap(const float* dy, const float* y, int act, float slope, float* g, int N, int64_t count, const float* target,
                    const float* tap_gout, void* stream);
int pvg_act_bwd_tap_split_16_scaled(const float* dy, const float* y, int act, float slope, float* g, void* planes, int N,
                                    int64_t count, const uint32_t* amax_bits, float* inv_scale, const float* target,
                                    const float* tap_gout, void* stream);
/* g = dy * act'(y) and the 16-bit plane pair of g (pvg_split_16) in one pass */
int pvg_act_bwd_split_16(const float* dy, const float* y, int act, float slope, float* g, void* planes, int64_t n, int fmt,
                         void* stream);

/* ---- BatchNorm2d (training: per-call batch statistics; eval: running statistics), optionally preceded by
 *      avg_pool2d(2) and followed by (+residual) and an activation.  Replaces F.avg_pool2d + nn.BatchNorm2d +
 *      LeakyReLU at residual_block.py:53-55,58,66-68, same_block.py:40-45, up_block.py:38-40,
 *      representation_network.py:42-44, conv_dynamics_network.py:41-45.
 *      `groups`: the batch dim is split in `groups` equal chunks with independent statistics (one reference
 *      BatchNorm call per chunk) - lets time steps be batched without changing the reference's semantics. ------- */
/* sums: double[groups][2][C], zero-initialised by caller: sum and sum of squares per channel */
int pvg_bn_stats(const float* x, int N, int HW, int C, int groups, double* sums, void* stream);
/* y = avg_pool2d(x, 2) (H, W even) and the statistics of y */
int pvg_pool2_stats(const float* x, int N, int H, int W, int C, float* y, int groups, double* sums, void* stream);
/* mean/invstd: float[groups][C]; running stats updated group after group with `momentum` (unbiased variance),
 * exactly as `groups` successive nn.BatchNorm2d training calls would. count = elements per channel per group. */
int pvg_bn_finalize(const double* sums, int64_t count, int groups, int C, float eps, float momentum,
                    float* running_mean, float* running_var, float* mean, float* invstd, void* stream);
/* pvg_bn_finalize + pvg_bn_apply in one launch (training mode; needs groups * C * 8 bytes <= 48 KB of shared memory) */
int pvg_bn_finalize_apply(const float* x, int N, int HW, int C, int groups, const double* sums, int64_t count, float eps,
                          float momentum, float* running_mean, float* running_var, float* mean, float* invstd,
                          const float* weight, const float* bias, const float* residual, int act, float slope, float* y,
                          void* stream);
/* eval mode: mean = running_mean, invstd = rsqrt(running_var + eps) */
int pvg_bn_eval_prepare(const float* running_mean, const float* running_var, int C, float eps,
                        float* mean, float* invstd, void* stream);
/* y = act((x - mean) * invstd * weight + bias (+ residual)) ; weight/bias may be NULL (affine=False) */
int pvg_bn_apply(const float* x, int N, int HW, int C, int groups, const float* mean, const float* invstd,
                 const float* weight, const float* bias, const float* residual, int act, float slope,
                 float* y, void* stream);
/* *_ex: the same pass also writes up to two 16-bit plane pairs of y (planes_a / planes_b: [2][numel(y)], formats fmt_a / fmt_b =
 * PVG_CORR_*; NULL = none; C % 8 == 0): the operands of the convolution that consumes y (forward: PVG_CORR_FP16_ALL; its
 * weight gradient: PVG_CORR_BF16), so that no separate pvg_split_16 pass re-reads y. */
int pvg_bn_apply_ex(const float* x, int N, int HW, int C, int groups, const float* mean, const float* invstd,
                    const float* weight, const float* bias, const float* residual, int act, float slope,
                    float* y, void* planes_a, int fmt_a, void* planes_b, int fmt_b, int Cparams, void* stream);
/* Cparams (0 = C): number of channels that HAVE parameters / running statistics - a 65-channel BatchNorm
 * (representation_network.py:28) applied to a tensor physically padded to 72 channels so that the convolutions around it run on
 * the tensor cores; padding channels are zero in, zero out, and are skipped by the running-statistics update. */
int pvg_bn_finalize_apply_ex(const float* x, int N, int HW, int C, int groups, const double* sums, int64_t count, float eps,
                             float momentum, float* running_mean, float* running_var, float* mean, float* invstd,
                             const float* weight, const float* bias, const float* residual, int act, float slope, float* y,
                             void* planes_a, int fmt_a, void* planes_b, int fmt_b, int Cparams, void* stream);
/* backward pass 1: with g = dy * act'(y): sums2 (double[groups][2][C], zeroed) += (sum g, sum g * xhat) */
int pvg_bn_bwd_reduce(const float* dy, const float* y, const float* x, int N, int HW, int C, int groups,
                      const float* mean, const float* invstd, int act, float slope, double* sums2, void* stream);
/* backward pass 2: dx = weight*invstd*(g - sum_g/M - xhat*sum_gx/M).  g_out (optional) receives g (gradient of the
 * residual branch).  unpool != 0: x/dy are at (H/2,W/2) and dx is written at (H,W) as 0.25*dx (avg_pool2d backward).
 * eval != 0: statistics are constants: dx = weight*invstd*g. */
int pvg_bn_bwd_apply(const float* dy, const float* y, const float* x, int N, int H, int W, int C, int groups,
                     const float* mean, const float* invstd, const float* weight, int act, float slope,
                     const double* sums2, int eval, int unpool, float* dx, float* g_out, float* dweight, float* dbias,
                     void* stream);   /* dweight / dbias (optional): pvg_bn_bwd_params folded into the same launch */
int pvg_bn_bwd_apply_ex(const float* dy, const float* y, const float* x, int N, int H, int W, int C, int groups,
                        const float* mean, const float* invstd, const float* weight, int act, float slope,
                        const double* sums2, int eval, int unpool, float* dx, float* g_out, float* dweight, float* dbias,
                        int Cparams, uint32_t* amax_out, void* stream);   /* Cparams: see pvg_bn_apply_ex; amax_out (optional): max|dx| bits */
/* dweight[c] = sum_groups sum_gx ; dbias[c] = sum_groups sum_g */
int pvg_bn_bwd_params(const double* sums2, int groups, int C, float* dweight, float* dbias, void* stream);

/* ---- input pipeline: PIL crop + transforms.ToTensor + transforms.Normalize(0.5, 0.5) of dataset/transforms.py:15-32,
 *      90-108 for frames that already have the target size (every shipped config).  src: [N][Hs][Ws][3] uint8 RGB,
 *      dst: [N][H][W][3] fp32 = ((u8 / 255) - mean) / std of the box at (left, top); bit-identical to the CPU transform. */
int pvg_frames_u8_to_nhwc(const uint8_t* src, int N, int Hs, int Ws, int left, int top, int H, int W, float mean, float std,
                          float* dst, void* stream);
/* One pass of PIL's Image.resize(size, BILINEAR) on 8-bit RGB frames (dataset/transforms.py:28: the resize applied when the
 * cropped frame does not have the model's input size; antialiased, 22-bit fixed point): src [N,Hs,Ws,3], input box
 * (left, top, Hin, Win), resampled along the width (vertical == 0: dst [N,Hin,out_size,3]) or the height (vertical != 0: dst
 * [N,out_size,Win,3]).  bounds [out_size][2] = (first input position, count), kk [out_size][ksize] = Pillow's integer
 * coefficients (device arrays; the host side computes them as precompute_coeffs + normalize_coeffs_8bpc do).  Horizontal pass
 * first, then vertical, as in ImagingResample: bit-identical to Pillow. */
int pvg_resample_u8(const uint8_t* src, int N, int Hs, int Ws, int left, int top, int Hin, int Win, int vertical, int out_size,
                    const int* bounds, const int* kk, int ksize, uint8_t* dst, void* stream);

/* ---- resampling: F.interpolate(scale_factor=2, mode='bilinear', align_corners=False) at up_block.py:35,43;
 *      F.interpolate(size, 'bilinear') of the ground truth at losses.py:92,450; nn.MaxPool2d(2) of VGG19 -------- */
int pvg_upsample2x_fwd(const float* x, int N, int H, int W, int C, float* y, void* stream);
int pvg_upsample2x_bwd(const float* dy, int N, int H, int W, int C, float* dx, void* stream);   /* H,W = input size */
int pvg_resize_bilinear(const float* x, int N, int H, int W, int C, float* y, int OH, int OW, void* stream);
int pvg_maxpool2_fwd(const float* x, int N, int H, int W, int C, float* y, void* stream);
int pvg_maxpool2_fwd_ex(const float* x, int N, int H, int W, int C, float* y, void* planes_a, int fmt_a, void* stream);
int pvg_resize_bilinear_ex(const float* x, int N, int H, int W, int C, float* y, int OH, int OW, void* planes_a, int fmt_a,
                           void* planes_b, int fmt_b, void* stream);
/* dx = (x is the first arg-max of its 2x2 window) ? dy : 0, times relu'(x) when relu_mask != 0 */
int pvg_maxpool2_bwd(const float* dy, const float* x, const float* y, int N, int H, int W, int C, int relu_mask,
                     float* dx, void* stream);

/* ---- channel concat feeding the recurrent / non-recurrent blocks of the dynamics network: torch.cat of maps and of (N, C) vectors
 *      repeated over H x W (conv_dynamics_network.py:64-109, convolutional_lstm_cell.py:88-89), zero-padded to Cpad physical
 *      channels (a tensor-core-legal K), optionally with the 16-bit plane pairs of the result. ------------------------- */
#define PVG_CONCAT_MAX_PARTS 6
typedef struct pvg_concat_desc {
  int32_t N, H, W, Cpad;
  int32_t nparts;
  int32_t c[PVG_CONCAT_MAX_PARTS];        /* channels of each part, in output order */
  int32_t is_vec[PVG_CONCAT_MAX_PARTS];   /* 0: NHWC map [N][H][W][c]; 1: vector [N][c] broadcast over H x W */
  int64_t bstride[PVG_CONCAT_MAX_PARTS];  /* elements between consecutive samples of the part (>= H*W*c for maps, >= c for vectors) */
  const void* src[PVG_CONCAT_MAX_PARTS];
} pvg_concat_desc;
int pvg_concat_pad(const pvg_concat_desc* d, float* y, void* planes_a, int fmt_a, void* planes_b, int fmt_b, void* stream);

/* ---- ConvLSTM cell point-wise part, convolutional_lstm_cell.py:92-101.  gates: [M][4][C] pre-activations in the
 *      order input, forget, output, cell. ------------------------------------------------------------------------ */
int pvg_lstm_fwd(const float* gates, const float* c_prev, int64_t M, int C, float* c_new, float* h_new, void* stream);
/* One ConvLSTM cell step with the point-wise part FUSED into the gate convolution (convolutional_lstm_cell.py:88-101): the four
 * gate convolutions are one implicit GEMM whose output columns are interleaved (column 4c + {0,1,2,3} = input, forget, output,
 * cell gate of hidden channel c: pack the weight rows and the bias in that order), and the epilogue applies bias, sigmoid / tanh,
 * c' = f*c + i*g, h' = o*tanh(c') in registers.  z_planes: PVG_CORR_FP16_ALL planes of the padded channel concat
 * [inputs..., h] [N,H,W,d->Cin]; w_planes: planes of the packed weight (pvg_pack_conv_weight + pvg_pack_16x2), d->Cout = 4C;
 * gates_act (NULL for inference): [N,H,W,4C] ACTIVATED gates, interleaved - what pvg_lstm_bwd_act needs. */
int pvg_convlstm_step(const pvg_conv_desc* d, const void* z_planes, const void* w_planes, const float* bias,
                      const float* c_prev, float* c_new, float* h_new, float* gates_act, void* stream);
/* backward of the point-wise part from the ACTIVATED, interleaved gates: dgates (pre-activation gradients, interleaved) and
 * dc_prev; dh / dc_new may be NULL */
int pvg_lstm_bwd_act(const float* gates_act, const float* c_prev, const float* c_new, const float* dh, const float* dc_new,
                     int64_t M, int C, float* dgates, float* dc_prev, void* stream);
int pvg_lstm_bwd(const float* gates, const float* c_prev, const float* c_new, const float* dh, const float* dc_new,
                 int64_t M, int C, float* dgates, float* dc_prev, void* stream);

/* ---- losses: nn.L1Loss (losses.py:59,118), |a-b|.mean(dim=[1,2,3]) per VGG level (losses.py:465) ------------- */
/* out[n] = mean_i |a[n,i] - b[n,i]| ; out must be zeroed by the caller */
int pvg_absdiff_mean_fwd(const float* a, const float* b, int N, int64_t count, double* out, void* stream);
/* db[n,i] = -sign(a-b) * gout[n] / count */
int pvg_absdiff_mean_bwd(const float* a, const float* b, const float* gout, int N, int64_t count, float* db,
                         void* stream);

/* ---- evaluator: per-frame squared-error reductions of evaluation/metrics/{mse,psnr,motion_masked_mse}.py and the
 *      uint8 frame conversion of evaluation/evaluation_dataset_builder.py:66-68,142-154 --------------------------------- */
/* out[n] (double, zeroed by the caller) = mean over (C, P) of (a - b)^2, n = b * T + t; use_mask: each pixel weighted by
 * sum_c |a_t - a_(t-1)| / C of the reference frames a (0 for t = 0).  channels_last: element (c, p) at c + p * C, else c * P + p. */
int pvg_sqdiff_mean(const float* a, const float* b, int N, int T, int C, int64_t P, int channels_last, int use_mask, double* out,
                    void* stream);
/* out[i] = (uint8)(v * 255), v = x in [0, 1], or (x + 1) / 2 when any element of x is negative; min_scratch: one int */
int pvg_frames_to_u8(const float* x, int64_t n, int* min_scratch, uint8_t* out, void* stream);

/* ---- optimiser: torch.optim.Adam(lr, weight_decay) as configured at training/trainer.py:36 -------------------- */
int pvg_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                  float eps, float weight_decay, int step, float grad_scale, void* stream);

/* same update with the step-dependent scalars read from DEVICE memory, so the launch can live in a CUDA graph:
 * hyper = [lr / (1 - beta1^t), sqrt(1 - beta2^t), beta1, beta2, eps, weight_decay, grad_scale] */
int pvg_adam_step_dev(float* p, const float* g, float* m, float* v, int64_t n, const float* hyper, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PVG_B200_H_ */
