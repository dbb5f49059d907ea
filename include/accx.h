/* accx -- C ABI of the B200-native (sm_100a) kernels behind ACC-UNet's HANC/MLFC blocks.
 *
 * This is the drop-in boundary.  The reference ( /root/reference/ACC_UNet/ACC_UNet.py ) has no
 * native interface for this path: its five torch.nn.Module classes call ATen operators.  The
 * entry points below are what a binding for those operators has to call; each one names the
 * reference lines it replaces.  Plain pointers and sizes only (no torch types); the caller owns
 * every buffer; nothing here allocates, frees or synchronises; every launch goes to the
 * caller-supplied CUDA stream (`stream` is a cudaStream_t passed as void*), so all calls are
 * capturable in a CUDA graph.  Return value: 0 on success, negative on error
 * (accx_last_error() gives the message for the calling thread).
 *
 * Data model
 *   activations  NHWC: a dense [P, C] matrix, P = B*H*W; storage ACCX_F32 or ACCX_BF16.
 *   parameters   fp32, in the reference's own layouts (Conv2d weight [out, in, kh, kw], ...).
 *   lazy operand raw tensor x + pending per-channel affine + activation:
 *                a = act(x*scale[c] + shift[c]); act 0 none, 1 affine, 2 affine+LeakyReLU(0.01).
 *                Training-mode BatchNorm2d + LeakyReLU never materialise: producers emit raw
 *                outputs and per-channel (sum, sum of squares); consumers normalise on load.
 *   stats        float[2*C]: sum then sum of squares, ACCUMULATED (atomicAdd) -- zero them first.
 */
#ifndef ACCX_H_
#define ACCX_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ACCX_F32 0
#define ACCX_BF16 1

#define ACCX_OK 0
#define ACCX_ERR_INVALID (-1)
#define ACCX_ERR_CUDA (-2)

#define ACCX_MAX_OPERANDS 9
#define ACCX_MAX_ADDENDS 4

const char* accx_last_error(void);
int accx_version(void);
/* Launch-geometry tuning knob `index` (see KNOB_* in csrc/common.cuh) := value; 0 restores the built-in default.
 * Process-global, meant for the sweeps in tests/bench_knobs.py -- results never depend on it. */
int accx_set_knob(int index, int value);
/* Deterministic reduction mode.  By default the cross-block part of every reduction (BatchNorm statistics, SE sums,
 * weight gradients) ends in fp32 atomics: results vary at rounding level from run to run, and ACC-UNet's 220
 * stacked BatchNorm2d layers (ACC_UNet.py:34,74,244-260,311-319,388-410) amplify that.  With a caller-owned
 * workspace installed, every reducing kernel adds its per-block partial sums in a FIXED order instead (two-stage:
 * per-block slots, folded in block order by the last block to arrive; tile-shaped partials -- weight gradients,
 * the persistent tcgen05 / TMA-tiled kernels -- get exactly one contribution per address): two runs are then
 * bit-identical.  `counters` = n_counters (>= 16384) zero-initialised 32-bit words; workspace >= 16 MiB (512 MiB
 * covers every shape of the 512 x 512 configuration).  Both are divided into 16 slices, one per CUDA stream that
 * launches reducing kernels, so concurrent streams (and the branches of a captured graph) never share scratch.
 * workspace = NULL restores the default.  Process-global; slower (parity / debugging mode), same arithmetic per
 * element. */
int accx_set_deterministic(void* workspace, int64_t workspace_bytes, unsigned int* counters, int n_counters);

/* One A-operand of a pointwise contraction together with its weight slice.
 *   value(p, k) = act(data[p'*ld + k]*scale[k] + shift[k]),  p' = pixel p shifted by (dy, dx)
 *                 inside its image (zero outside -- this is conv padding);
 *   weight(n, k) = w[n*w_ld + k*w_ks]   (pointer pre-offset by the caller).
 * The strided weight view is how the reference's interleaved concat orders are honoured with
 * no data movement: HANCLayer K index c*(2k-1)+j (ACC_UNet.py:138), MLFC merge 2c+j (:492),
 * MLFC gather block order (:431-485), dense 3x3 taps (ResPath, :316-318). */
typedef struct accx_operand {
  const void* data;
  int64_t ld;
  int32_t K;
  int32_t act;
  const float* scale;
  const float* shift;
  const float* w;
  int64_t w_ld;
  int64_t w_ks;
  int32_t dy, dx;
} accx_operand_t;

/* Y[p, n] = sum_ops sum_k value_op(p, k) * weight_op(n, k) + bias[n]
 *           + sum_j add_j[coarse_j(p), n]          (nearest-upsample-add, log2 factor add_log2s[j])
 * and stats[n] += sum_p Y, stats[N+n] += sum_p Y^2 (if stats != NULL).
 * Replaces: every 1x1 Conv2d of the path (HANCBlock.conv1/conv3 ACC_UNet.py:243,259;
 * HANCLayer.cnv :72 in split form W.cat_j(up(p_j)) = sum_j up(W_j.p_j); Conv2d_batchnorm.conv1
 * :171), the dense 3x3 of ResPath (:316-318, nine shifted operands), and all their input
 * gradients (same contraction with transposed weight strides).
 * add_j are fp32 [B, H>>l, W>>l, N] tensors.  B,H,W describe the pixel grid of Y. */
int accx_pw_fwd(int dtype, int out_dtype, int B, int H, int W, int N,
                const accx_operand_t* ops, int n_ops, const float* bias,
                const float* const* add, const int* add_log2s, int n_add,
                void* y, int64_t ldy, float* stats, void* stream);

/* Same contract as accx_pw_fwd on the tcgen05 tensor cores (bf16 operands, fp32 accumulation in
 * TMEM): producer warps apply the pending BatchNorm affine + LeakyReLU while staging the
 * 128B-swizzled A tiles, weights arrive by bulk TMA from a bf16 re-pack made in `workspace`
 * (accx_pw_tc_workspace_bytes gives its size).  Requires every operand K and ld to be a multiple
 * of 8 elements and 16-byte aligned bases; otherwise use accx_pw_fwd. */
int64_t accx_pw_tc_workspace_bytes(int N, const accx_operand_t* ops, int n_ops);
int accx_pw_fwd_tc(int dtype, int out_dtype, int B, int H, int W, int N,
                   const accx_operand_t* ops, int n_ops, const float* bias,
                   const float* const* add, const int* add_log2s, int n_add,
                   void* y, int64_t ldy, float* stats, void* workspace, int64_t workspace_bytes, void* stream);
/* Diagnostics: with knob 19 (KNOB_TC_DEBUG) bit 5 set, CTA 0 of accx_pw_fwd_tc stamps %globaltimer at the role hand-offs
 * of its first 64 tiles; this copies the 10 x 64 nanosecond stamps out (event * 64 + tile; events in csrc/gemm_tc.cu). */
int accx_debug_tc_trace(unsigned long long* dst, int n);
/* The launch plan accx_pw_fwd_tc_res makes for these arguments, without touching the device (host logic only; of the
 * pointers only `ops` and `add_log2s` are read, the operands' data / weight pointers are not dereferenced):
 * plan[0..11] = {pixel folding (two pixels per row) 0/1, dense-3x3 slab mode 0/1, column tile BN, column tiles, pixel
 * tiles, pipeline stages, dynamic shared-memory bytes, weights resident in shared memory 0/1, floats of the first addend
 * staged per row (0 = direct loads), TMEM columns, k-blocks, grid}.  n_plan >= 12.  Nothing in the reference. */
int accx_pw_fwd_tc_plan(int dtype, int out_dtype, int B, int H, int W, int N, const accx_operand_t* ops, int n_ops,
                        int n_add, const int* add_log2s, int has_residual, int64_t ld_res, int64_t ldy, int has_stats,
                        int* plan, int n_plan);
/* accx_pw_fwd_tc with a residual: Y = contraction (+ bias + addends) + R, R a [P, ld_res] matrix in the OUTPUT dtype
 * (16-byte aligned base and row pitch).  R may be Y itself (in-place accumulation: every tile is read before it is
 * written).  Fuses the `x + inp` / gradient-accumulation passes that follow an input-gradient contraction
 * (HANCBlock backward ACC_UNet.py:279, ResPath backward :325-326, MLFC's per-level sums :431-485). */
int accx_pw_fwd_tc_res(int dtype, int out_dtype, int B, int H, int W, int N,
                       const accx_operand_t* ops, int n_ops, const float* bias,
                       const float* const* add, const int* add_log2s, int n_add, const void* residual, int64_t ld_res,
                       void* y, int64_t ldy, float* stats, void* workspace, int64_t workspace_bytes, void* stream);

/* dW[n*w_ld + k*w_ks] += sum_p dY[p, n] * value(p, k)  for one operand (weight gradient of the
 * contraction above; fp32 atomics).  dy is a plain [P, ldy] matrix in `dtype` storage
 * (or fp32 when dy_f32 != 0). */
int accx_pw_wgrad(int dtype, int B, int H, int W, int N, const accx_operand_t* op, float* dw,
                  const void* dy, int64_t ldy, int dy_f32, void* stream);

/* accx_pw_wgrad on the tcgen05 tensor cores: both operands are TMA-loaded as pixel-major (MN-major)
 * 128B-swizzled tiles, the activation operand is normalised in place, partial tiles are split over
 * pixels and combined with fp32 atomics.  bf16 operands; K, N, ld multiples of 8. */
int accx_pw_wgrad_tc(int B, int H, int W, int N, const accx_operand_t* op, float* dw, const void* dy, int64_t ldy,
                     void* stream);
/* Several taps of a dense k x k convolution's weight gradient (ResPath 3x3, ACC_UNet.py:316-318) in ONE pass over dY
 * and the activation: tap t reads the activation shifted by (taps_dy[t], taps_dx[t]) (zero outside the image) and
 * accumulates into dw + taps_woff[t] (floats; the [C, C, 3, 3] weight has w_ld = 9C, w_ks = 9, offset = tap index).
 * op->dy/dx are ignored.  n_taps x min(round16(K), 64) must fit the 512 TMEM columns. */
int accx_pw_wgrad_taps_tc(int B, int H, int W, int N, const accx_operand_t* op, int n_taps, const int* taps_dy,
                          const int* taps_dx, const int64_t* taps_woff, float* dw, const void* dy, int64_t ldy,
                          void* stream);
/* The launch plan of accx_pw_wgrad_tc / accx_pw_wgrad_taps_tc for these arguments, without touching the device (host logic
 * only; the operand's data / weight pointers are not dereferenced): plan[0..10] = {pixel folding 0/1, A-channel tile NB,
 * dY-channel tiles, A-channel tiles, dY blocks per stage, pixels per pipeline stage, splits over pixels, pipeline stages,
 * dynamic shared-memory bytes, TMEM columns, grid}.  n_plan >= 11.  Nothing in the reference. */
int accx_pw_wgrad_tc_plan(int B, int H, int W, int N, const accx_operand_t* op, int n_taps, const int* taps_dy,
                          const int* taps_dx, int64_t ldy, int* plan, int n_plan);


/* BatchNorm2d statistics -> affine (torch.nn.BatchNorm2d as used at ACC_UNet.py:34,74,178,
 * 244-260,311-319,388-410).  training: mean/var from stats (biased var), running buffers
 * updated with `momentum` and the unbiased variance, num_batches_tracked += 1.  eval: running
 * buffers.  Outputs scale = gamma*rstd, shift = beta - mean*scale, and mean/rstd for backward.
 * conv_bias (optional): the bias of the convolution that produced the tensor, which was NOT added to
 * it: a per-channel constant cancels in batch normalisation, so it only enters the running mean
 * (training) / the shift (eval) here and never touches the activation tensor. */
int accx_bn_finalize(int C, double count, const float* stats, const float* gamma, const float* beta,
                     const float* conv_bias, float eps, float momentum, int training, float* running_mean,
                     float* running_var, int64_t* num_batches_tracked, float* scale, float* shift, float* mean,
                     float* rstd, void* stream);

/* out = post(act(x)) with optional second affine post(u) = u*scale2 + shift2 (ResPath tail
 * BN(lrelu(BN(x))), ACC_UNet.py:328), optional residual add, optional stats of the value
 * written.  out == NULL: statistics only. */
int accx_act_apply(int dtype, int64_t P, int C, const void* x, const float* scale, const float* shift, int act,
                   const float* scale2, const float* shift2, const void* residual, void* out, float* stats,
                   void* stream);

/* BatchNorm2d backward through a lazy boundary a = act(y*scale+shift):
 *   g = da * act'(.), sums[c] += sum g, sums[C+c] += sum g*xhat          (reduce)
 *   dy = gamma*rstd*(g - sums[c]/n - xhat*sums[C+c]/n); dgamma += sums[C+c]; dbeta += sums[c] (apply)
 * dy may alias da. */
int accx_bn_bwd_reduce(int dtype, int64_t P, int C, const void* y, const float* scale, const float* shift, int act,
                       const float* mean, const float* rstd, const void* da, float* sums, void* stream);
int accx_bn_bwd_apply(int dtype, int64_t P, int C, const void* y, const float* scale, const float* shift, int act,
                      const float* mean, const float* rstd, const float* gamma, const void* da, const float* sums,
                      double count, void* dy, float* dgamma, float* dbeta, void* stream);

/* Depthwise 3x3, stride 1, zero padding 1 (HANCBlock.conv2, ACC_UNet.py:246-252) on a lazy
 * input; raw output + stats.  w is the Conv2d weight [C,1,3,3]; flip != 0 rotates the filter by
 * 180 degrees (= input gradient).  wgrad: dw[c*9 + tap] += sum dy[p] * a[p + tap]. */
int accx_dw3x3_fwd(int dtype, int B, int H, int W, int C, const void* x, const float* scale, const float* shift,
                   int act, const float* w, const float* bias, int flip, void* y, float* stats, void* stream);
int accx_dw3x3_wgrad(int dtype, int B, int H, int W, int C, const void* x, const float* scale, const float* shift,
                     int act, const void* dy, float* dw, void* stream);
/* Input gradient da = dwconv(dy, rot180(w)) fused with the BatchNorm-backward reduction of the layer in front of
 * the depthwise conv (norm1, ACC_UNet.py:244,270): y1 is that layer's raw output, g = da*act'(y1*bn_scale+bn_shift),
 * sums[c] += sum g, sums[C+c] += sum g*(y1-mean)*rstd.  Needs TMA-addressable tensors (C*elem % 16 == 0). */
int accx_dw3x3_dgrad_bnred(int dtype, int B, int H, int W, int C, const void* dy, const float* w, void* da,
                           const void* y1, const float* bn_scale, const float* bn_shift, int bn_act,
                           const float* bn_mean, const float* bn_rstd, float* sums, void* stream);

/* HANC pyramid (HANCLayer.forward, ACC_UNet.py:83-136), one 2x level per call.
 * first != 0: x is the lazy [B,H,W,C] map, else x is the previous level [B,H,W,2C] (avg | max).
 * out [B,H/2,W/2,2C]: channels [0,C) 2x2 average, [C,2C) 2x2 maximum. */
int accx_hanc_pool_fwd(int dtype, int B, int H, int W, int C, int first, const void* x, const float* scale,
                       const float* shift, int act, void* out, void* stream);
/* Gradient of the s x s (s = 1<<log2s) avg/max branches w.r.t. the activated map a:
 * da[p] (+)= davg[blk]/s^2 + (p is the FIRST row-major maximum of its block ? dmax[blk] : 0),
 * dpool = [B,H/s,W/s,2C] (avg | max), fp32. */
int accx_hanc_unpool_bwd(int dtype, int B, int H, int W, int C, int log2s, const void* x, const float* scale,
                         const float* shift, int act, const float* dpool, void* da, int accumulate, void* stream);

/* The same gradient for ALL levels of a k = 2 / 3 pyramid (levels = 1 / 2; dpool1 = level-1 gradients
 * [B,H/2,W/2,2C], dpool2 = level-2 [B,H/4,W/4,2C], fp32) accumulated into da in ONE pass, fused with the
 * BatchNorm-backward reduction of the layer in front (accx_bn_bwd_reduce on the updated da):
 * sums[c] += sum g, sums[C+c] += sum g*xhat.  bf16 storage, C % 4 == 0. */
int accx_hanc_unpool_bnred(int dtype, int B, int H, int W, int C, int levels, const void* y, const float* scale,
                           const float* shift, int act, const float* dpool1, const float* dpool2, void* da,
                           const float* mean, const float* rstd, float* sums, void* stream);

/* out[b, h, w, coff + c] = mul * sum over the s x s block of x (MLFC's chained AvgPool2d(2),
 * ACC_UNet.py:361,448-480 with mul = 1/s^2; block sums of gradients with mul = 1). */
int accx_pool_sum(int in_dtype, int out_dtype, int B, int H, int W, int C, int log2s, float mul, const void* x,
                  void* out, int64_t out_ld, void* stream);
/* dst[b,h,w,c] (+)= mul * src[b, h>>l, w>>l, c]  (nearest Upsample, ACC_UNet.py:360, and the
 * gradient of the average pool). */
int accx_upsample_add(int in_dtype, int out_dtype, int B, int H, int W, int C, int log2s, float mul,
                      const void* src, int64_t src_ld, void* dst, int accumulate, void* stream);

/* ChannelSELayer (ACC_UNet.py:37-49) in three launches:
 *   squeeze  S[0,b,c] += sum_hw a, S[1,b,c] += sum_hw a^2   (one read gives the gate input AND
 *            the batch statistics of the gated tensor: mean_c = sum_b g*S1/n, E[z^2] = sum_b g^2*S2/n)
 *   gate     fc1 -> LeakyReLU -> fc2 -> sigmoid, then the trailing BatchNorm's scale/shift
 *   apply    out = lrelu(a*gate[b,c]*scale[c] + shift[c]) (+ residual) (stats of out optional)
 * mix (device scalar, ACC_UNet_w.py:497-522): out = v*mix + residual*(1-mix). */
int accx_se_squeeze(int dtype, int B, int HW, int C, const void* x, const float* scale, const float* shift, int act,
                    float* S, void* stream);
int accx_se_gate(int B, int C, int Cr, double HW, const float* S, const float* w1, const float* b1, const float* w2,
                 const float* b2, const float* gamma, const float* beta, float eps, float momentum, int training,
                 float* running_mean, float* running_var, int64_t* num_batches_tracked, float* gate, float* hidden,
                 float* scale, float* shift, float* mean, float* rstd, unsigned int* counter, void* stream);
int accx_se_apply(int dtype, int B, int HW, int C, const void* x, const float* scale, const float* shift, int act,
                  const float* gate, const float* se_scale, const float* se_shift, const void* residual,
                  const float* mix, void* out, float* stats, void* stream);
/* backward: G[0,b,c] += sum_hw g', G[1,b,c] += sum_hw g'*a with g' = dout*lrelu'(v)  (reduce);
 * the tiny gate kernel turns G into per-(b,c) coefficients PQR and all parameter gradients;
 * apply: da (+)= P*g' + Q*a + R.  With mix: g' carries the factor mix and
 * dmix += sum dout*(v - residual).  bn_sums != NULL (apply): the lazy input's own BatchNorm-backward reduction
 * (accx_bn_bwd_reduce on the da just written) is accumulated in the same pass.
 * training = 0 (gate): the layer's BatchNorm ran on its running statistics (eval mode) -- a fixed affine, so the
 * batch-mean terms of its backward are dropped (the reference differentiates eval-mode modules the same way). */
int accx_se_bwd_reduce(int dtype, int B, int HW, int C, const void* x, const float* scale, const float* shift, int act,
                       const float* gate, const float* se_scale, const float* se_shift, const void* dout,
                       const float* mix, const void* residual, float* dmix, float* G, void* stream);
int accx_se_bwd_gate(int B, int C, int Cr, double HW, const float* S, const float* G, const float* gate,
                     const float* hidden, const float* w1, const float* w2, const float* gamma, const float* mean,
                     const float* rstd, float* dw1, float* db1, float* dw2, float* db2, float* dgamma, float* dbeta,
                     float* PQR, int training, void* stream);
int accx_se_bwd_apply(int dtype, int B, int HW, int C, const void* x, const float* scale, const float* shift, int act,
                      const float* gate, const float* se_scale, const float* se_shift, const void* dout,
                      const float* mix, const float* PQR, void* da, int accumulate, const float* bn_mean,
                      const float* bn_rstd, float* bn_sums, void* stream);

/* z = act(a) + r with stats (HANCBlock: norm(x + inp), ACC_UNet.py:279). */
int accx_add_fwd(int dtype, int64_t P, int C, const void* a, const float* scale, const float* shift, int act,
                 const void* r, void* z, float* stats, void* stream);

/* NCHW <-> NHWC with optional dtype change (model entry / exit only). */
int accx_nchw_to_nhwc(int in_dtype, int out_dtype, int B, int C, int HW, const void* src, void* dst, void* stream);
int accx_nhwc_to_nchw(int in_dtype, int out_dtype, int B, int C, int HW, const void* src, void* dst, void* stream);

/* ---- the steps either side of the blocks inside one training step ---- */

/* MaxPool2d(2) between encoder levels (ACC_UNet.py:552,608-618), NHWC.  Backward recomputes the arg-max from
 * the input and routes the gradient to the first maximum in row-major window order (ATen's tie rule);
 * it writes every element of dx (H, W even). */
int accx_maxpool2_fwd(int dtype, int B, int H, int W, int C, const void* x, void* out, void* stream);
int accx_maxpool2_bwd(int dtype, int B, int H, int W, int C, const void* x, const void* dy, void* dx, void* stream);

/* ConvTranspose2d(2, 2, stride 2) + the skip torch.cat of the decoder (ACC_UNet.py:578-599,620-631).  The
 * transposed conv itself is ONE accx_pw_fwd contraction [P, Cin] x [Cin, 4*Co] (the weight [Cin, Co, 2, 2] read
 * through a strided view; column co*4 + ky*2 + kx = output pixel (2h+ky, 2w+kx), channel co) into `temp`
 * [B, H, W, 4*Co].  forward = 1: out[b, 2h+ky, 2w+kx, co] = temp[b, h, w, co*4+ky*2+kx] + bias[co], out is the
 * [B, 2H, 2W, ld_out] concat buffer (left Co columns); forward = 0: the inverse gather of the gradient into temp,
 * and dbias[co] += sum of that gradient over pixels and taps (the bias gradient; may be NULL).
 * accx_copy_cols copies C columns between matrices of different row pitch (skip half of the concat, its gradient). */
int accx_upshuffle(int dtype, int forward, int B, int H, int W, int Co, void* temp, const float* bias, void* out,
                   int64_t ld_out, float* dbias, void* stream);
int accx_copy_cols(int dtype, int64_t P, int C, const void* src, int64_t ld_src, void* dst, int64_t ld_dst,
                   void* stream);

/* UNeXt shifted tokenized-MLP block (Experiments/nets/UNext.py:38-160; BASELINE configs[2]).  Tokens [B, N, C] are an
 * NHWC tensor: the shift (pad / chunk(5) / roll / narrow, :78-84,:97-103) is expressed as five shifted operands of
 * accx_pw_fwd / accx_pw_wgrad, DWConv (:150-160) is accx_dw3x3_fwd with a bias, and these are the two remaining passes:
 *   layernorm  nn.LayerNorm(C) over the channels of every row (shiftedBlock.norm2, :151,:156): y = (x-mean)*rstd*gamma + beta,
 *              biased variance, mean / rstd [R] (fp32) saved for the backward (either may be NULL);
 *              bwd: dx, and dgamma / dbeta ACCUMULATED (zero them first; either may be NULL).  C <= 1024.
 *   gelu       nn.GELU, exact erf form (:47,:90); bwd: dx = dy * gelu'(x). */
int accx_layernorm_fwd(int dtype, int64_t R, int C, const void* x, const float* gamma, const float* beta, float eps,
                       void* y, float* mean, float* rstd, void* stream);
int accx_layernorm_bwd(int dtype, int64_t R, int C, const void* x, const float* gamma, const float* mean,
                       const float* rstd, const void* dy, void* dx, float* dgamma, float* dbeta, void* stream);
int accx_gelu_fwd(int dtype, int64_t n, const void* x, void* y, void* stream);
int accx_gelu_bwd(int dtype, int64_t n, const void* x, const void* dy, void* dx, void* stream);

/* Per-step segmentation metrics left on the device (the reference syncs device -> host every step for them:
 * Experiments/Train_one_epoch.py:134-135).  pred = sigmoid(logit) >= 0.5, mask = truth > 0;
 * out[0] = mean over images of sklearn's binary jaccard_score (iou_on_batch, Experiments/utils.py:478-494; 0 for an
 * empty union), out[1] = WeightedDiceBCE._show_dice (utils.py:148-157, including its second sigmoid on the binarised
 * prediction).  logit [B, N] (dtype), truth [B, N] fp32; counts: uint32[4*B + 1] ZEROED by the caller
 * (per image TP, #pred, #mask as exact integers + a block counter). */
int accx_seg_metrics(int dtype, int B, int64_t N, const void* logit, const float* truth, unsigned int* counts, float* out,
                     void* stream);

/* WeightedDiceBCE(dice_weight, BCE_weight) on one-class logits (Experiments/utils.py:21-74 BCE normalised over
 * positives / negatives, :109-138 soft Dice on sigmoid(logit) with class weights [0.5, 0.5], :140-171 the sum).
 * logit [B, N] (dtype), truth [B, N] fp32.  sums: float[8*B + 8], ZEROED by the caller (per-image partial sums
 * + a block counter); loss: one float, written by the last block.  bwd: dlogit = gscale[0] * dloss/dlogit in
 * grad_dtype (gscale may be NULL = 1). */
int accx_dice_bce_fwd(int dtype, int B, int64_t N, const void* logit, const float* truth, float dice_w, float bce_w,
                      float* sums, float* loss, void* stream);
int accx_dice_bce_bwd(int dtype, int grad_dtype, int B, int64_t N, const void* logit, const float* truth,
                      const float* sums, float dice_w, float bce_w, const float* gscale, void* dlogit, void* stream);

/* torch.optim.Adam(lr) step (train_model.py:647) over ONE flat fp32 buffer holding all parameters back to
 * back (n a multiple of 4, 16-byte aligned buffers).  state[0] = step count, kept on the device and incremented
 * by the call (graph-capturable).  grad is multiplied by grad_scale first; weight_decay is torch's L2 form.
 * lr < 0: the learning rate is read from state[1] on the device (LR schedules under a captured graph). */
int accx_adam_step(int64_t n, float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float* state,
                   float lr, float beta1, float beta2, float eps, float weight_decay, float grad_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ACCX_H_ */
