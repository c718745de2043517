/* falnet_b200.h -- C ABI of libfalnet_sm100.so (hand-written sm_100a kernels for the FAL-net hot path).
 *
 * The reference (JuanLuisGonzalez/FAL_net) has NO native/FFI layer: its hot path is Python calling
 * ATen/cuDNN.  Each entry point below therefore cites the reference *Python*
 * code whose device work it replaces; the Python side of the boundary (fal_net_b200/*.py) keeps the
 * reference's names and signatures and calls these functions through ctypes with raw device
 * pointers + the current CUDA stream (INTEGRATION.md shows the binding).
 *
 * Conventions
 *   - every function is extern "C", takes plain pointers / sizes, returns int: 0 = ok, <0 = error
 *     (faln_last_error() gives the thread-local message).  No allocation, no host sync, no
 *     exceptions inside: every call only enqueues work on `stream` and is CUDA-graph capturable.
 *   - all pointers are DEVICE pointers unless named h_*.  Tensors are contiguous in the stated
 *     layout; "pitch" arguments are in ELEMENTS.
 *   - fp32 tensors are NCHW like the reference's; bf16 activations/weights of the conv family are
 *     NHWC / KRSC (channels innermost).
 */
#ifndef FALNET_B200_H_
#define FALNET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* faln_stream_t; /* cudaStream_t */

#define FALN_OK 0
#define FALN_ERR_ARG (-1)     /* bad argument (null pointer, unsupported size, misalignment) */
#define FALN_ERR_LAUNCH (-2)  /* CUDA launch error; see faln_last_error() */
#define FALN_ERR_UNSUPPORTED (-3)

int faln_version(void);
const char* faln_last_error(void);
/* Number of kernels this library has launched from the calling process since load (all threads). */
long long faln_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * MED view synthesis (the fused hot kernel).
 *
 * Replaces, in /root/reference/models/FAL_netB.py: softmax :216, disparity expectation :219-226,
 * affine_grid + 2N grid clones :231-243,258-262,270-271, the N grid_sample + O(N^2) cat :244-247,
 * softmax :248, the blend loop :279-282 and the sub-occlusion masks :264-273,290-294.
 *
 *   logits  [B,N,H,Wp] fp32   dlog0 (row pitch `logit_pitch` >= W elements, plane stride H*pitch)
 *   image   [B,3,H,W]  fp32   the input view
 *   g0x     [W]        fp32   x row of F.affine_grid(identity, align_corners=True): linspace(-1,1,W)
 *   x_of    [B,N]      fp32   normalised-grid offset of level n (reference :241)
 *   d_lvl   [B,N]      fp32   disparity of level n in pixels   (reference :225)
 * outputs (any of pan / disp / maskL+maskR may be NULL = not wanted):
 *   pan     [B,3,H,W], disp [B,1,H,W], maskL/maskR [B,1,H,W] (already clamped to <= 1)
 *   lse0    [B,1,H,W]  log-sum-exp over planes of the un-warped logits   (saved for backward)
 *   lsew    [B,1,H,W]  log-sum-exp over planes of the warped logits      (saved for backward)
 * flags: FALN_MED_FORCE_GENERIC forces the per-pixel floor path on every plane (testing).
 */
#define FALN_MED_FORCE_GENERIC 1u
#define FALN_MED_TUNE_2CTA 2u /* rows <= 1024 px: 2 CTAs/SM at 112 registers instead of 3 CTAs/SM at 96 (tuning aid) */
#define FALN_MED_ZERO_PAD 8u  /* caller's promise: logit rows are 16-byte aligned, logit_pitch % 4 == 0 and the pad columns
                               * [W, logit_pitch) hold zeros (what fal_net_b200.layout.alloc_planar produces): enables the
                               * branch-free fast kernels for W % 4 != 0 (rows with W % 4 == 0 qualify without it) */
#define FALN_MED_NO_FAST 16u  /* never take the fast kernels (validation of one path against the other) */
#define FALN_MED_NO_V3 32u    /* skip the third-generation kernels (csrc/med3.cu: barrier-free gathers, max-free softmax
                               * sums + clean-up launch) and use the second-generation ones (A/B comparison) */
#define FALN_MED_V3_GENERIC 64u /* third generation: every plane on its per-pixel generic code (testing) */
int faln_med_fwd(const float* logits, const float* image, const float* g0x, const float* x_of,
                 const float* d_lvl, float* pan, float* disp, float* maskL, float* maskR, float* lse0,
                 float* lsew, int B, int N, int H, int W, long long logit_pitch, unsigned flags,
                 faln_stream_t stream);

/* Backward of the MED section w.r.t. the logits (autograd of reference :216-282; the image, the level
 * tables and the masks carry no gradient, :223,237,256).  Single sweep over the planes: uses the
 * forward's saved pan / disp / lse0 / lsew (dot(x) = <g_pan(x), pan(x)>), gather-only, no atomics.
 *   g_pan [B,3,H,W] or NULL, g_disp [B,1,H,W] or NULL  ->  g_logits [B,N,H,Wp] (pitch = g_pitch)
 */
int faln_med_bwd(const float* logits, const float* image, const float* g0x, const float* x_of,
                 const float* d_lvl, const float* pan, const float* disp, const float* lse0,
                 const float* lsew, const float* g_pan, const float* g_disp, float* g_logits, int B,
                 int N, int H, int W, long long logit_pitch, long long g_pitch, unsigned flags,
                 faln_stream_t stream);

/* Disparity-only epilogue (inference path, reference :216-229): disp = sum_n d_n softmax_n(logits). */
int faln_med_disp(const float* logits, const float* d_lvl, float* disp, int B, int N, int H, int W,
                  long long logit_pitch, faln_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Losses.  All reductions are deterministic two-stage sums: each call writes per-block partials to
 * `partials` (>= faln_loss_partials_len() floats) and a last-block-done pass folds them, in fixed
 * order, into out[0] (+= semantics are NOT used: out[0] is overwritten).
 * ---------------------------------------------------------------------------------------------- */
int faln_loss_partials_len(void);

/* Masked L1 reconstruction + (optionally) the blended image fed to VGG
 * (/root/reference/loss_functions.py:52-56):
 *   out[0]   = mean(mask * |synth - label|)          over B*3*H*W elements
 *   blend    = mask*synth + (1-mask)*label            (written if blend != NULL)
 *   g_synth  = g_scale * mask * sign(synth-label)/(B*3*H*W) (+ g_blend*mask if g_blend != NULL)
 * mask is [B,1,H,W] or NULL (the integer 1 of Stage-1, Train_Stage1_K.py:247).
 * If flip_x != 0, synth is read (and g_synth written) x-reversed -- the exact index flip that replaces
 * the reference's grid_sample un-flip, Train_Stage2_K.py:283-286; label/mask/blend stay un-flipped. */
int faln_loss_rec_l1(const float* synth, const float* label, const float* mask, float* blend,
                     float* out, float* partials, int B, int H, int W, int flip_x,
                     faln_stream_t stream);
/* Backward kernels scale by g_scale * (g_dev ? *g_dev : 1): g_dev is an optional DEVICE scalar (the
 * upstream gradient of the loss value) so that no host sync is needed to chain losses. */
int faln_loss_rec_l1_bwd(const float* synth, const float* label, const float* mask, const float* g_blend,
                         float g_scale, const float* g_dev, float* g_synth, int B, int H, int W, int flip_x,
                         faln_stream_t stream);

/* Edge-aware smoothness on the column window [x_lo, x_hi) of img [B,3,H,W] / disp [B,1,H,W]
 * (/root/reference/loss_functions.py:70-109 applied to the slices of Train_Stage1_K.py:255 /
 * Train_Stage2_K.py:312-313; stencils are zero-padded at the WINDOW border like the reference's
 * conv2d on the sliced tensor).  flip_x: disp is read x-reversed. */
int faln_loss_smooth(const float* img, const float* disp, float gamma, float* out, float* partials,
                     int B, int H, int W, int x_lo, int x_hi, int flip_x, faln_stream_t stream);
int faln_loss_smooth_bwd(const float* img, const float* disp, float gamma, float g_scale, const float* g_dev,
                         float* g_disp, int accumulate, int B, int H, int W, int x_lo, int x_hi, int flip_x,
                         faln_stream_t stream);

/* Mirror loss (/root/reference/Train_Stage2_K.py:319-324) on the window [x_lo,x_hi):
 *   out[0] = mean_{b,y,x in window} (1/max_b(mdisp)) * (1 - occ) * |disp - mdisp|
 * inv_max [B] = 1 / max over the whole image of mdisp[b]. */
int faln_loss_mirror(const float* disp, const float* mdisp, const float* occ, const float* inv_max,
                     float* out, float* partials, int B, int H, int W, int x_lo, int x_hi, int flip_x,
                     faln_stream_t stream);
int faln_loss_mirror_bwd(const float* disp, const float* mdisp, const float* occ, const float* inv_max,
                         float g_scale, const float* g_dev, float* g_disp, int accumulate, int B, int H, int W,
                         int x_lo, int x_hi, int flip_x, faln_stream_t stream);

/* mean((a-b)^2) over n elements of bf16 NHWC feature maps (perceptual loss,
 * /root/reference/loss_functions.py:59-67) and its gradient w.r.t. a. */
int faln_mse_bf16(const void* a, const void* b, long long n, float* out, float* partials,
                  faln_stream_t stream);
int faln_mse_bf16_bwd(const void* a, const void* b, long long n, float g_scale, const float* g_dev, void* g_a,
                      faln_stream_t stream);

/* Per-image max of a [B,1,H,W] fp32 map -> inv_max[b] = 1/max (F.max_pool2d(kernel=(H,W)),
 * /root/reference/Train_Stage2_K.py:319-320). */
int faln_inv_rowmax(const float* x, float* inv_max, int B, long long n_per_image, faln_stream_t stream);

/* Occlusion masks of Stage-2 (/root/reference/Train_Stage2_K.py:295-299):
 *   O = flip_a?(a) * flip_b?(b);  O[..., one_lo:one_hi] = 1     (flips are exact x index reversals) */
int faln_occ_mask(const float* a, const float* b, float* out, int B, int H, int W, int flip_a, int flip_b,
                  int one_lo, int one_hi, faln_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Optimiser: fused multi-tensor Adam (torch.optim.Adam as configured at
 * /root/reference/Train_Stage1_K.py:177-181: betas (0.5, 0.999), eps 1e-8, wd 0), over one flat
 * fp32 arena: p, g, m, v are [n] device arrays.  Also refreshes the bf16 shadow copy used by the conv
 * kernels (w16 may be NULL).  grad_scale multiplies g first (1/world_size after an allreduce-SUM).
 * ---------------------------------------------------------------------------------------------- */
int faln_adam(float* p, const float* g, float* m, float* v, void* w16, long long n, float lr, float beta1,
              float beta2, float eps, float weight_decay, int step, float grad_scale,
              faln_stream_t stream);
/* CUDA-graph-capturable variant: the step counter and learning rate live in device memory, hp = float[4]
 * {lr, step, (derived) lr/(1-beta1^step), (derived) 1/sqrt(1-beta2^step)}; each call advances hp[1] by one. */
int faln_adam_dev(float* p, const float* g, float* m, float* v, void* w16, long long n, float* hp, float beta1,
                  float beta2, float eps, float weight_decay, float grad_scale, faln_stream_t stream);
/* The same update over a RANGE of the arenas (p, g, m, v, w16 point at the range's first element, n % 4 == 0): lets the
 * optimiser run bucket by bucket, right behind each bucket's gradient all-reduce, while the rest of backward is still
 * executing.  tick != 0 advances the step counter hp[1] first (exactly one call per step must tick: the first one). */
int faln_adam_dev_range(float* p, const float* g, float* m, float* v, void* w16, long long n, float* hp, float beta1,
                        float beta2, float eps, float weight_decay, float grad_scale, int tick, faln_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Layout / elementwise helpers of the conv pipeline (bf16 NHWC activations).
 * ---------------------------------------------------------------------------------------------- */
/* fp32 NCHW [B,C,H,W] -> bf16 NHWC [B,H,W,Cp] (channels C..Cp-1 zero); flip_x reverses x. */
int faln_nchw_to_nhwc_bf16(const float* src, void* dst, int B, int C, int H, int W, int Cp, int flip_x,
                           faln_stream_t stream);
/* bf16 NHWC [B,H,W,Cp] -> fp32 planar [B,C,H,pitch] (first C channels). */
int faln_nhwc_bf16_to_planar(const void* src, float* dst, int B, int C, int H, int W, int Cp,
                             long long pitch, faln_stream_t stream);
int faln_planar_to_nhwc_bf16(const float* src, void* dst, int B, int C, int H, int W, int Cp,
                             long long pitch, faln_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * tcgen05 / TMEM implicit-GEMM 3x3 convolution, pad 1, stride 1|2 (bf16 in, fp32 accumulate): replaces the
 * cuDNN calls behind nn.Conv2d at /root/reference/models/FAL_netB.py:99-127,144-174 and the VGG slices at
 * /root/reference/loss_functions.py:21-29, with bias / ELU / ReLU / residual add (:47,59,79) fused.
 *   x  [B,H,W,C1] bf16 NHWC; x2 [B,H,W,C2] bf16 or NULL = second source, channel-concatenated after x
 *      (the skip connections of :153-173 -- torch.cat never materialises); C1, C2 multiples of 32
 *   w  [Cout_pad,3,3,C1+C2] bf16 (KRSC; rows >= Cout are zero), bias [Cout] fp32 or NULL
 *   ctab [16,Cout] + cscale [B] fp32 or NULL: a spatially constant extra input channel of value cscale[b] (the
 *      max_disp/100 plane, :145,208-209) folded into a per-border-class bias: class = rc*4+cc,
 *      rc = (top tap outside) | (bottom tap outside) << 1, cc likewise; ctab = sums of that channel's weights over
 *      the taps inside the image
 *   residual [B,Ho,Wo,out_c] bf16 or NULL;  act: 0 none, 1 ELU, 2 ReLU
 *   y  planar == 0: bf16 NHWC [B,Ho,Wo,out_c];  planar == 1: fp32 [B,Cout,Ho,out_pitch] (the logits layout
 *      the MED kernels stream; lets the last layer emit fp32 straight from the accumulator)
 * Ho = (H-1)/stride + 1, Wo likewise.
 * ---------------------------------------------------------------------------------------------- */
int faln_conv3x3_fwd(const void* x, const void* x2, const void* w, const float* bias, const float* ctab,
                     const float* cscale, const void* residual, void* y, int B, int H, int W, int C1, int C2, int Cout, int Cout_pad, int stride, int act,
                     int planar, long long out_pitch, int out_c, faln_stream_t stream);

/* Data gradient of the same convolution on the same tcgen05 kernel (replaces cuDNN dgrad behind autograd of
 * /root/reference/models/FAL_netB.py:35-80): g [B,Hg,Wg,Cg] bf16 NHWC (gradient w.r.t. the conv's pre-activation output),
 * wd [Cx_pad,3,3,Cg] bf16 (weights re-packed per input channel; one call per concatenated source), gx [B,H,W,gx_c] bf16.
 * Fused epilogue: gx[..., :Cx] = ([accum] gx_old + dgrad + [residual]) * act'(ysave), dact 0 none / 1 ELU / 2 ReLU.
 * stride 2 = four output-parity classes in one launch. */
int faln_conv3x3_dgrad(const void* g, const void* wd, void* gx, const void* residual, const void* ysave, int B, int H,
                       int W, int Cg, int Cx, int Cx_pad, int stride, int accum, int dact, int gx_c, int res_c,
                       int ysave_c, faln_stream_t stream);
/* bf16 shadow of the flat fp32 parameter arena (what faln_adam keeps up to date): plain conversion, n % 4 == 0. */
int faln_f32_to_bf16(const float* src, void* dst, long long n, faln_stream_t stream);
/* Re-pack every 3x3 conv weight of the bf16 shadow arena for faln_conv3x3_dgrad in ONE launch:
 * per job j, jobs[6j..] = {src_off, dst_off, Cout, Cin_tot, Cin_used, 0} (element offsets; Cout, Cin_used % 32 == 0):
 *   w16 + src_off: [Cout][3][3][Cin_tot] (KRSC)  ->  wd16 + dst_off: [Cin_used][3][3][Cout].
 * max_tiles >= max_j 9 * (Cout/32) * (Cin_used/32).  `jobs` is device memory. */
int faln_pack_dgrad_batched(const void* w16, void* wd16, const long long* jobs, int njobs, int max_tiles,
                            faln_stream_t stream);
/* The same over a flat tile list (no empty blocks, 4-byte accesses): jobs[6j+5] = index of job j's first 32x32 tile in the
 * concatenation of all jobs' tiles (9 * (Cout/32) * (Cin_used/32) tiles per job), total_tiles their number. */
int faln_pack_dgrad_flat(const void* w16, void* wd16, const long long* jobs, int njobs, int total_tiles,
                         faln_stream_t stream);
/* Inference form of the last layer (reference models/FAL_netB.py:174,215-229; Test_KITTI.py:196): the folded logits conv
 * with the softmax-expectation over the N disparity levels fused into its epilogue, disp [B,1,H,W] fp32 =
 * sum_n d_lvl[b,n] * softmax_n(conv3x3(cat(x, x2)) + bias).  The logit planes never reach HBM.  Stride 1, W >= 192,
 * N <= 64 with the weight padded to 64 rows, bias [64] padded with -inf and d_lvl [B,64] padded with 0 (both 16-byte
 * aligned); otherwise FALN_ERR_ARG (write the logits and call faln_med_disp). */
int faln_conv3x3_logits_disp(const void* x, const void* x2, const void* w, const float* bias, const float* d_lvl,
                             float* disp, int B, int H, int W, int C1, int C2, int N, int Cout_pad, faln_stream_t stream);
/* Weight gradient of the 3x3 convolution (pad 1, stride 1 or 2) on tcgen05 -- replaces the cuDNN wgrad autograd runs for
 * nn.Conv2d of /root/reference/models/FAL_netB.py:99-127 in loss.backward() (/root/reference/Train_Stage1_K.py:260).
 *   dW[co, kh, kw, ci_off + ci] += sum_{b,ho,wo} g[b,ho,wo,co] * x[b, ho*s+kh-1, wo*s+kw-1, ci]
 * g [B,(H-1)/s+1,(W-1)/s+1,Cg] bf16 NHWC (pre-activation gradient), x [B,H,W,Cxs] bf16 NHWC (one source of a concatenated
 * input per call), dW [Cout,3,3,Cin_tot] fp32 -- the KRSC memory of a torch.channels_last [Cout,Cin_tot,3,3] tensor --
 * ACCUMULATED with split-K fp32 reductions (zero it once per step).
 * Cg, Cxs: 32 or multiples of 64; Cout <= Cg and Cx <= Cxs select the channels actually written.
 * flags: 0 normally; bit 0 disables the halo-tile path (validation only). */
int faln_conv3x3_wgrad(const void* g, const void* x, float* dW, int B, int H, int W, int Cg, int Cxs, int Cout, int Cx,
                       int ci_off, int Cin_tot, int stride, unsigned flags, faln_stream_t stream);
/* The same with the layer's BIAS gradient taken along: dbias [Cout] fp32 (NULL = none) += sum over (b, ho, wo) of
 * g[b,ho,wo,co] -- replaces the sum autograd runs for nn.Conv2d(bias=True) (/root/reference/models/FAL_netB.py:35-48).  Nine
 * taps leave one operand-window slot of the kernel's last MMA row unused; it reads ones, so the sum costs no extra pass over g.
 * For a concatenated input pass dbias with ONE of the per-source calls only. */
int faln_conv3x3_wgrad_bias(const void* g, const void* x, float* dW, float* dbias, int B, int H, int W, int Cg, int Cxs,
                            int Cout, int Cx, int ci_off, int Cin_tot, int stride, unsigned flags, faln_stream_t stream);
/* Several layers' weight (+ bias) gradients at once: every job is what one faln_conv3x3_wgrad_bias call takes; jobs that map to
 * the same kernel configuration share ONE grid.  For the small-map layers (3x10 ... 12x40), whose separate launches are ~15 us
 * latency chains that hold SMs beside the data-gradient chain of loss.backward(). */
typedef struct faln_wgrad_job {
  const void* g;
  const void* x;
  float* dW;
  float* dbias; /* NULL = no bias gradient */
  int B, H, W, Cg, Cxs, Cout, Cx, ci_off, Cin_tot, stride;
  unsigned flags;
} faln_wgrad_job_t;
int faln_conv3x3_wgrad_multi(const faln_wgrad_job_t* jobs, int njobs, faln_stream_t stream);
/* Weight gradient of the reference's deconv block -- F.interpolate(scale 2, nearest) then conv3x3
 * (/root/reference/models/FAL_netB.py:51-60) -- taken straight from the LOW-resolution input, the counterpart of
 * faln_conv3x3_up2_fwd / _dgrad: sixteen quarter-resolution correlations folded into the nine taps (2.25x fewer MMAs, no
 * up-sampled tensor).  g [B,2H,2W,Cg] bf16 NHWC (pre-activation gradient on the up-sampled grid), x [B,H,W,Cxs] bf16 NHWC,
 * dW [Cout,3,3,Cin_tot] fp32 KRSC accumulated in columns [ci_off, ci_off + Cx).  Cg, Cxs multiples of 64. */
int faln_conv3x3_wgrad_up2(const void* g, const void* x, float* dW, int B, int H, int W, int Cg, int Cxs, int Cout, int Cx,
                           int ci_off, int Cin_tot, faln_stream_t stream);
/* Several folded deconv layers at once (one grid; H, W of a job = the LOW-resolution size; stride, flags, dbias ignored). */
int faln_conv3x3_wgrad_up2_multi(const faln_wgrad_job_t* jobs, int njobs, faln_stream_t stream);
/* out [B,3,3,C] fp32 += sums of g [B,H,W,Cs] (bf16 NHWC) per sample over the 3x3 border classes (first / interior / last
 * row x column): the weight gradient of a spatially constant input channel (reference :145,208-209) is a 9-term
 * combination of these.  H, W >= 2. */
int faln_border_sum_nhwc(const void* g, float* out, int B, int H, int W, int C, int Cs, faln_stream_t stream);
/* Backward of F.interpolate(mode='nearest') (reference :58) fused with the producer's activation derivative. */
int faln_upsample_nearest_bwd_nhwc(const void* g_hi, const void* y_lo, void* g_lo, int B, int Hl, int Wl, int Hh, int Wh,
                                   int C, int dact, int accum, faln_stream_t stream);
/* Backward of the VGG 2x2 max-pool (/root/reference/loss_functions.py:21-29) fused with the ReLU derivative. */
int faln_maxpool2_bwd_nhwc(const void* x, const void* g_y, void* g_x, int B, int Hi, int Wi, int C, int dact,
                           faln_stream_t stream);
/* Bias gradient: out[c] += sum over pixels of g[pix, c] (bf16 NHWC with Cstride channels per pixel, fp32 accumulate). */
int faln_channel_sum_nhwc(const void* g, float* out, long long npix, int C, int Cstride, faln_stream_t stream);

/* Stem convolution for 3-channel fp32 NCHW input (conv0.0 of :99 and VGG conv1_1): reads the image directly
 * (flip_x: x-reversed), w [Cout,3,3,3] fp32 (torch OIHW), Cout 32 or 64, writes bf16 NHWC [B,H,W,Cout]. */
int faln_stem_conv(const float* x, const float* w, const float* bias, void* y, int B, int H, int W, int Cout,
                   int act, int flip_x, faln_stream_t stream);
/* The same layer as ONE tcgen05 kernel (the default): every thread gathers the 27 taps of its pixel from the fp32 image, writes
 * them as a bf16 hi + lo pair into a swizzled K = 64 operand row in shared memory (the patch matrix never reaches HBM and the
 * image keeps 16 mantissa bits), four MMAs per 128 pixels, bias + activation epilogue out of TMEM.  Same arguments. */
int faln_stem_conv_mma(const float* x, const float* w, const float* bias, void* y, int B, int H, int W, int Cout, int act,
                       int flip_x, faln_stream_t stream);
/* Weight and bias gradient of that first layer (3 -> 32, conv0.0 of /root/reference/models/FAL_netB.py:99) from the fp32 image
 * itself -- autograd's grad_weight / grad_bias of the first nn.Conv2d: x [B,3,H,W] fp32 NCHW, g [B,H,W,32] bf16 NHWC
 * (pre-activation gradient), dW [32,3,3,3] fp32 in KRSC memory ([co][kh][kw][c]) and dbias [32] (NULL = none) accumulated into. */
int faln_stem_wgrad(const float* x, const void* g, float* dW, float* dbias, int B, int H, int W, faln_stream_t stream);
/* F.interpolate(mode='nearest') (:58) on bf16 NHWC: src index = min(floor(dst * in/out), in-1). */
int faln_upsample_nearest_nhwc(const void* src, void* dst, int B, int Hi, int Wi, int Ho, int Wo, int C,
                               faln_stream_t stream);
/* 2x2 stride-2 max pooling (VGG pools, /root/reference/loss_functions.py:21-29) on bf16 NHWC. */
int faln_maxpool2_nhwc(const void* src, void* dst, int B, int Hi, int Wi, int C, faln_stream_t stream);

/* Tensor-core form of faln_stem_conv: the 3x3x3 patch of every pixel as a 32-wide bf16 K vector (one pass over the image,
 * zero padding, optional flip) + a K = 32 GEMM on the tcgen05 tile kernel with bias / activation in the epilogue.
 * col: scratch [B,H,W,32] bf16, wpack: scratch [Cout,32] bf16 (both written by the call).  Same operands / result otherwise. */
int faln_stem_conv_tc(const float* x, const float* w, const float* bias, void* y, void* col, void* wpack, int B, int H, int W,
                      int Cout, int act, int flip_x, faln_stream_t stream);
/* Nearest 2x up-sampling folded into the 3x3 convolution that follows it (the reference's deconv block,
 * /root/reference/models/FAL_netB.py:51-60: F.interpolate(nearest) :58 + conv3x3 + ELU :59), exact-2x sizes only: four taps per
 * output parity class instead of nine, the up-sampled tensor never exists.  x [B,H,W,C1] bf16 NHWC (LOW resolution);
 * w [Cout_pad][16][C1] bf16 from faln_pack_up2_weights; y [B,2H,2W,out_c] bf16 NHWC. */
int faln_conv3x3_up2_fwd(const void* x, const void* w, const float* bias, void* y, int B, int H, int W, int C1, int Cout,
                         int Cout_pad, int act, int out_c, faln_stream_t stream);
/* Its data gradient w.r.t. the LOW-resolution input: nearest-upsample backward (2x2 sum) + 3x3 data gradient as one 4x4
 * stride-2 gather, with the producer's activation derivative fused (dact / ysave as in faln_conv3x3_dgrad).
 * g [B,2H,2W,Cg] bf16; wd [Cx][16][Cg] bf16 from faln_pack_up2_weights; gx [B,H,W,gx_c]. */
int faln_conv3x3_up2_dgrad(const void* g, const void* wd, void* gx, const void* ysave, int B, int H, int W, int Cg, int Cx,
                           int dact, int gx_c, int ysave_c, faln_stream_t stream);
/* Folded weights of such a block from the fp32 [Cout,Cin,3,3] weight (element strides so,sc,sh,sw): sums of the original taps
 * in fp32, rounded once to bf16.  fwd_pack [Cout_pad][16][Cin], dgrad_pack [Cin_pad][16][Cout_pad]. */
int faln_pack_up2_weights(const float* w, long long so, long long sc, long long sh, long long sw, void* fwd_pack,
                          void* dgrad_pack, int Cout, int Cin, int Cout_pad, int Cin_pad, faln_stream_t stream);
/* The same for several deconv layers in one launch. */
typedef struct faln_up2_pack_job {
  const float* w;
  long long so, sc, sh, sw; /* element strides of w [Cout][Cin][3][3] */
  void* fwd_pack;
  void* dgrad_pack;
  int Cout, Cin, Cout_pad, Cin_pad;
} faln_up2_pack_job_t;
int faln_pack_up2_weights_multi(const faln_up2_pack_job_t* jobs, int njobs, faln_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Small device-side helpers that keep ATen / cuBLAS launches out of a training step: csrc/small_ops.cu.
 * ---------------------------------------------------------------------------------------------- */
/* d_lvl[b,n], x_of[b,n] of /root/reference/models/FAL_netB.py:204-205,224-225,241 in ONE launch, same fp32 op order as the
 * reference's per-level ATen kernels (every operation rounded to fp32; expf / logf).  min_disp, max_disp [B]. */
int faln_level_tables(const float* min_disp, const float* max_disp, float* d_lvl, float* x_of, int B, int N, int W,
                      faln_stream_t stream);
/* Folded logits layer W' = W0 . W_iconv1 (/root/reference/models/FAL_netB.py:174,190,215: a 3x3 conv without bias or
 * activation followed by a 1x1 conv).  w0 [N,N] fp32; w_iconv1 logical [N,C,3,3] fp32 with element strides (so,sc,sh,sw);
 * fwd_pack [Np,3,3,C] bf16 (rows >= N zero), dgrad_pack [Cp,3,3,Np] bf16 (zero padded). */
int faln_fold_logit_conv(const float* w0, const float* w_iconv1, long long so, long long sc, long long sh, long long sw,
                         void* fwd_pack, void* dgrad_pack, int N, int C, int Np, int Cp, faln_stream_t stream);
/* Adjoint of the fold: gwf [N,3,3,C] fp32 (gradient of W', KRSC) -> g_w0 [N,N] += <gwf, W_iconv1>,
 * g_w_iconv1 (strided like the parameter's gradient view) += W0^T gwf. */
int faln_fold_logit_conv_bwd(const float* gwf, const float* w0, const float* w_iconv1, long long so, long long sc, long long sh,
                             long long sw, float* g_w_iconv1, long long gso, long long gsc, long long gsh, long long gsw,
                             float* g_w0, int N, int C, faln_stream_t stream);
/* Border-class sums [16,Cout] of the weights of input channel `channel` (the constant max_disp/100 plane,
 * /root/reference/models/FAL_netB.py:145,208-209) of a strided [Cout,Cin,3,3] weight. */
int faln_const_channel_table(const float* w, long long so, long long sc, long long sh, long long sw, int channel, float* ctab,
                             int Cout, faln_stream_t stream);
/* dW[:, channel, :, :] += weight gradient of that constant channel from the nine border-class sums of the output gradient
 * (border_sums [B,3,3,Cout], faln_border_sum_nhwc) and the per-sample value [B]. */
int faln_const_channel_wgrad(const float* border_sums, const float* value, float* dW, long long so, long long sc, long long sh,
                             long long sw, int channel, int B, int Cout, int last_row_clipped, int last_col_clipped,
                             faln_stream_t stream);
/* out = sum_i weights[i] * *terms[i]  (n <= 8 device scalars; `terms` and `weights` are HOST arrays): the loss arithmetic of
 * /root/reference/Train_Stage1_K.py:258 / Train_Stage2_K.py:309-327 in one launch; and its adjoint out[i] = weights[i] * *g. */
int faln_scalar_combine(const float* const* terms, const float* weights, int n, float* out, faln_stream_t stream);
int faln_scalar_scale(const float* g, const float* weights, int n, float* out, faln_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Inference post-processing and validation metrics: csrc/postproc.cu.
 * ---------------------------------------------------------------------------------------------- */
/* flip_x (optional) + F.interpolate(mode='bilinear', align_corners=True): /root/reference/Test_KITTI.py:291-292.
 * in [BC,H,W] fp32 -> out [BC,Ho,Wo] fp32. */
int faln_flip_resize_bilinear(const float* in, float* out, int BC, int H, int W, int Ho, int Wo, int flip_x,
                              faln_stream_t stream);
/* out[b] = (float)(numpy.percentile(x[b, :n], 100*q) + add), exact (radix select on the device): replaces the
 * `np.percentile(disp.detach().cpu().numpy(), 95) + 1e-6` host round trip of /root/reference/Test_KITTI.py:297.
 * x [B, stride >= n] fp32. */
int faln_percentile_rows(const float* x, int B, long long n, long long stride, double q, double add, float* out,
                         faln_stream_t stream);
/* out = (1 - norm) * disp + norm * up_mul * unflip(nearest_up(small)), norm = min(disp / p[b], 1):
 * /root/reference/Test_KITTI.py:294-300.  disp, out [B,1,H,W]; small [B,1,Hs,Ws] (flipped coordinates); p [B]. */
int faln_mspp_blend(const float* disp, const float* small, const float* p, float* out, int B, int H, int W, int Hs, int Ws,
                    float up_mul, faln_stream_t stream);
/* The seven KITTI depth errors as per-image fp64 partial sums {count, abs_rel, sq_rel, sq, log_sq, n_a1, n_a2, n_a3}
 * (sums [B,8], zeroed by the call): /root/reference/myUtils.py:196-232 applied to the depths of :234-254 (mode 0,
 * KITTI2015: gt is a disparity map, both sides fb / disp) or :256-277 (mode 1, Eigen: gt is a depth map, window =
 * [H-219,H-4) x [44,1180)).  gt, pred [B,H,W] fp32. */
int faln_kitti_errors(const float* gt, const float* pred, double* sums, int B, int H, int W, int y0, int y1, int x0, int x1,
                      int mode, double fb_gt, double fb_pred, double min_d, double max_d, faln_stream_t stream);
/* realEPE (/root/reference/loss_functions.py:124-141,170-173): sums = {sum |target - bilinear_up(output)|, count} over
 * target != 0 (sparse) or all pixels.  output [B,1,h,w], target [B,1,H,W]. */
int faln_real_epe(const float* output, const float* target, double* sums, int B, int h, int w, int H, int W, int sparse,
                  faln_stream_t stream);
/* get_rmse (/root/reference/myUtils.py:138-150): sum = sum over [B,3,H,W] of (clamp((o+mean_c)*255,0,255) - (l+mean_c)*255)^2 */
int faln_rmse255(const float* output, const float* label, double* sum, int B, int H, int W, float m0, float m1, float m2,
                 faln_stream_t stream);

/* FAL_netA's maskR only (/root/reference/models/FAL_netA.py:264): the right-from-left occlusion mask sampled with
 * grid_sample's default align_corners=False on the align_corners=True grid (a 2-D bilinear resampling), from the logits
 * and the saved log-sum-exp of the un-warped logits.  g0x [W], g0y [H]: rows of the identity affine grid. */
int faln_maskr_noalign(const float* logits, const float* lse0, const float* g0x, const float* g0y, const float* x_of,
                       float* maskR, int B, int N, int H, int W, long long logit_pitch, faln_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Training input pipeline on the device: csrc/input_pipe.cu.
 * Replaces /root/reference/data_transforms.py:46-157 + the input transform of /root/reference/Train_Stage1_K.py:124-128
 * (4 DataLoader workers running PIL / numpy per sample): decoded uint8 HWC images in, normalised fp32 NCHW crops out,
 * bit-exact with the reference (integer resampling; value chain as a 3x256 table per image).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const void* src; /* device pointer: uint8 [H,W,3] (HWC, like the numpy array the reference's loader returns) */
  int H, W;        /* source size */
  int row0, rows;  /* source rows the crop needs: [row0, row0+rows) */
  int x_tab, y_tab;/* offsets (ints) into `tabs`: bounds[2*t] then coeffs[t*ks], for the crop's columns / rows */
  int ksx, ksy;    /* taps per output pixel (faln_pil_bicubic_ksize) */
  int flip;        /* 1: mirror the crop horizontally (RandomHorizontalFlip, data_transforms.py:96-116) */
  int dst;         /* image slot in the output tensor */
  int lut;         /* offset (floats) into `luts`: 3 x 256 values, uint8 -> normalised fp32 */
  int pad_;
} faln_aug_desc;

/* HOST functions (no CUDA): Pillow's bicubic resampling tables (Resample.c precompute_coeffs + normalize_coeffs_8bpc)
 * for output pixels [out_lo, out_lo+out_n) of an in_size -> out_size resize; what `Image.resize(..., Image.BICUBIC)`
 * at data_transforms.py:67 computes internally.  bounds [2*out_n], coeffs [out_n * ksize]. */
int faln_pil_bicubic_ksize(int in_size, int out_size);
int faln_pil_bicubic_coeffs(int in_size, int out_size, int out_lo, int out_n, int* bounds, int* coeffs);
/* Two batched launches (horizontal pass into `inter`, vertical pass + value table + mirror into `out`).
 * descs [n_img], tabs, luts: device; inter: n_img * inter_stride bytes of scratch (>= max_rows*tw*3 each);
 * out [n_slots,3,th,tw] fp32. */
int faln_augment_crops_u8(const faln_aug_desc* descs, int n_img, const int* tabs, const float* luts, unsigned char* inter,
                          long long inter_stride, int max_rows, float* out, int th, int tw, faln_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* FALNET_B200_H_ */
