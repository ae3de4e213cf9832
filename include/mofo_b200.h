/* libmofo_sm100.so — C ABI of the B200-native MOFO pretraining hot path.
 *
 * The reference (Moohnai/MOFO) has no FFI layer: the hot path is plain PyTorch reached through the Python
 * modules masking_generator / modeling_pretrain / engine_for_pretraining.  Every entry point
 * below names the reference code (file:line under the MOFO tree) whose device work it replaces; the Python
 * look-alike modules in mofo_b200/ bind them with ctypes.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; all pointers are DEVICE pointers unless stated otherwise.
 *   - every call returns 0 on success or a negative mofo status; mofo_last_error() returns a thread-local message.
 *   - every call takes the CUDA stream it launches on (cudaStream_t passed as void*); no call synchronises,
 *     allocates device memory or takes ownership of anything.  Outputs and workspaces are caller-allocated.
 *   - bf16 tensors are passed as mofo_bf16* (raw 16-bit storage); "rows" are tokens, row-major, leading
 *     dimension given in elements.
 *   - token n of a clip = t*(H'*W') + h*W' + w; feature orders are stated per call.
 */
#ifndef MOFO_B200_H_
#define MOFO_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MOFO_B200_VERSION 200

typedef uint16_t mofo_bf16;

/* status codes */
#define MOFO_STATUS_OK 0
#define MOFO_STATUS_INVALID (-1)
#define MOFO_STATUS_CUDA (-2)
#define MOFO_STATUS_UNSUPPORTED (-3)

int mofo_version(void);
const char* mofo_last_error(void);
int mofo_sm_count(void);

/* ------------------------------------------------------------------------------------------------------------
 * (1) Tube masking with the motion-box constraint.
 * Replaces TubeMaskingGenerator_BB.__call__ (masking_generator.py:43-85) for a batch of clips, bit-exact given the
 * same random draw.  bb_first[b] = the FIRST frame's box (x1,y1,x2,y2) (the reference reads bb[0] only, :46,55).
 * rng_words[b, 0..W) = the raw 32-bit MT19937 outputs numpy's legacy shuffle would consume for clip b.
 * Outputs: mask[b, T*H*Wd] (1 = masked), vis_idx[b, T*(H*Wd-n_mask)] / msk_idx[b, T*n_mask] = ascending token ids
 * (what x[~mask] / x[mask] enumerate, modeling_pretrain.py:90,261-262), words_used[b] (-1 if W was too small).
 * mofo_tube_mask_plain is the non-BB generator (masking_generator.py:17-24).
 */
int mofo_tube_mask_bb(const double* bb_first, const uint32_t* rng_words, int B, int W, int T, int H, int Wd,
                      int n_mask_per_frame, double ratio_bb, uint8_t* mask, int32_t* vis_idx, int32_t* msk_idx,
                      int32_t* words_used, void* stream);
int mofo_tube_mask_plain(const uint32_t* rng_words, int B, int W, int T, int H, int Wd, int n_mask_per_frame,
                         uint8_t* mask, int32_t* vis_idx, int32_t* msk_idx, int32_t* words_used, void* stream);

/* Index lists of a caller-supplied boolean mask [B,N] (1 byte per token, non-zero = masked): the ascending visible /
 * masked token ids that x[~mask] / x[mask] enumerate (modeling_pretrain.py:90,261-262), without the device->host sync
 * of torch.nonzero.  Every row must hold exactly n_msk masked tokens (the reference's .reshape(B,-1,C) requires equal
 * counts); bad_rows[0] is incremented for each row that does not. */
int mofo_mask_indices(const uint8_t* mask, int B, int N, int n_msk, int32_t* vis_idx, int32_t* msk_idx,
                      int32_t* bad_rows, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * (2) Tubelet gather for the patch embedding.
 * Replaces the input side of PatchEmbed.forward (Conv3d k=s=(2,16,16), modeling_finetune.py:238-248) followed by
 * x[~mask] (modeling_pretrain.py:90): only the visible tubes are read.  Writes the im2col matrix
 * A[b*n_idx + j, c*512 + p0*256 + p1*16 + p2] = video[b, c, 2t+p0, 16h+p1, 16w+p2] in bf16, where (t,h,w) is token
 * idx[b, j].  Feature order (c,p0,p1,p2) = the flattened Conv3d weight order, so the embedding is A · W^T.
 * video: f32 [B,3,frames,size,size] contiguous (NCTHW).  tubelet 2, patch 16 are fixed (modeling_pretrain.py:112).
 */
int mofo_gather_tubes(const float* video, const int32_t* idx, int B, int n_idx, int frames, int size,
                      mofo_bf16* A, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * (3) Dense layers on tcgen05 tensor cores (bf16 x bf16 -> f32 accumulate in TMEM).
 * mofo_gemm_tn:  C[M,N] = epilogue( A[M,K] · B[N,K]^T ), A and B row-major bf16 (both K-contiguous).
 *   Replaces F.linear / nn.Linear forward (modeling_finetune.py:45-50,82-85,96; modeling_pretrain.py:156,256) and,
 *   with B = W^T (bf16 copy kept by the caller), the input-gradient GEMMs of their backward.
 * Epilogues (aux pointers may be NULL when unused):
 *   BIAS_BF16       out0 bf16 = acc + bias[n]
 *   BIAS_GELU_BF16  u = bf16(acc + bias[n]) (the pre-activation F.linear returns under autocast);
 *                   out1 bf16 = gelu_erf(u)  (Mlp.forward fc1 + nn.GELU, modeling_finetune.py:45-46);
 *                   out0 bf16 = gelu_erf'(u) (kept for backward in place of u: same erf/exp evaluation, so the
 *                   backward epilogue is a single multiply)
 *   BIAS_RESID_F32  out0 f32 = acc + bias[n] + resid[m,n]  (residual add of Block.forward, :218-219); with row_scale
 *                   (f32, one value per group of group_rows rows = per clip): out0 = resid + row_scale[m / group_rows] *
 *                   (acc + bias[n]) - DropPath on the residual branch (modeling_finetune.py:20-31, 218-219)
 *   PLAIN_BF16      out0 bf16 = acc
 *   GELU_BWD_BF16   out0 bf16 = acc * aux_bf16[m,n], aux = the gelu_erf'(u) saved by BIAS_GELU_BF16 (backward through nn.GELU)
 *   BIAS_POS_F32    out0 f32 [row' , n] = acc + bias[n] + pos[row_idx[m], n]; row' = (m / group_rows) *
 *                   out_group_rows + m % group_rows.  With bias = patch-embed bias this is "+ pos_embed" then the
 *                   visible gather (modeling_pretrain.py:85-90); with bias = NULL, group_rows = N_vis,
 *                   out_group_rows = N it is encoder_to_decoder + pos_emd_vis written in place of the
 *                   torch.cat of :260-263.
 * bias/pos/resid are f32.  ld* in elements.  Requirements: K % 8 == 0, N % 8 == 0, 16-byte aligned bases.
 */
enum {
  MOFO_EPI_BIAS_BF16 = 0,
  MOFO_EPI_BIAS_GELU_BF16 = 1,
  MOFO_EPI_BIAS_RESID_F32 = 2,
  MOFO_EPI_PLAIN_BF16 = 3,
  MOFO_EPI_GELU_BWD_BF16 = 4,
  MOFO_EPI_BIAS_POS_F32 = 5
};
int mofo_gemm_tn(const mofo_bf16* A, int lda, const mofo_bf16* B, int ldb, int M, int N, int K, int epilogue,
                 const float* bias, const float* resid, int ldr, const mofo_bf16* aux_bf16, int ldaux,
                 const float* pos, const int32_t* row_idx, int group_rows, int out_group_rows, void* out0, int ldo0,
                 void* out1, int ldo1, const float* row_scale, void* stream);

/* mofo_gemm_wgrad: dW[N,K] += dY[M,N]^T · X[M,K]   (f32 accumulate into dW with red.global.add; dW must hold the
 * running sum, e.g. a zeroed slice of the gradient arena).  Replaces the weight-gradient GEMM of every nn.Linear /
 * the Conv3d patch embedding in autograd's backward (utils.py:354 scale(loss).backward()).
 * dY, X row-major bf16; reduction runs over rows, split across CTAs.  N % 8 == 0, K % 8 == 0.
 * dbias (may be NULL): f32 [N], dbias[n] += sum_m dY[m,n] (the layer's bias gradient) computed in the same pass;
 * columns n in [dbias_skip_lo, dbias_skip_hi) are left untouched (the K third of the qkv bias, which is the constant 0
 * of modeling_finetune.py:82-84 and has no parameter). */
int mofo_gemm_wgrad(const mofo_bf16* dY, int ldy, const mofo_bf16* X, int ldx, int M, int N, int K, float* dW,
                    int ldw, float* dbias, int dbias_skip_lo, int dbias_skip_hi, void* stream);

/* mofo_gemm_wgrad_grouped: n (1..4) independent mofo_gemm_wgrad problems that share the reduction length M, in ONE launch
 * (HOST arrays of n device pointers / sizes; dbias, dbias_skip_lo, dbias_skip_hi may be NULL).  The four weight gradients of
 * a transformer block (fc2, fc1, proj, qkv) are independent and off the backward critical path.  One k-tile width serves the
 * whole group: every K[i] % 192 == 0 (ViT-S / ViT-B widths), or every K[i] % 256 == 0 (ViT-L: 1024 / 4096, decoder 512). */
int mofo_gemm_wgrad_grouped(int n, const mofo_bf16* const* dY, const int* ldy, const mofo_bf16* const* X, const int* ldx, int M,
                            const int* N, const int* K, float* const* dW, const int* ldw, float* const* dbias,
                            const int* dbias_skip_lo, const int* dbias_skip_hi, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * (4) Fused multi-head attention, head_dim 64 (every registry model, modeling_pretrain.py:268-338).
 * Replaces Attention.forward lines modeling_finetune.py:85-95 (split heads, q*scale, softmax(q k^T), attn @ v,
 * merge heads) and their backward.  qkv: bf16 [B*S, 3*H*64] = output of the QKV linear (q | k | v, each H*64 wide);
 * out: bf16 [B*S, H*64]; lse: f32 [B,H,S] (log2-domain log-sum-exp kept for backward).
 * out_lo (may be NULL): bf16 [B*S, H*64] = out_exact - out, the second bf16 word of the attention output.  Backward uses
 *   it only for delta_i = dO_i . O_i, the softmax-backward row term (softmax's Jacobian needs sum_j dS_ij = 0; an error
 *   e_i in delta leaks e_i * (probability-weighted mean key) into dQ, which the q_bias gradient - a batch-wide sum of
 *   cancelling terms - accumulates).  Sequences with S <= 192 run single-pass kernels that hold the whole score row in
 *   tensor memory and form delta_i = sum_j P_ij dP_ij directly; they neither write nor read out_lo.
 * Backward: dqkv bf16 [B*S, 3*H*64]; delta: f32 workspace [B,H,S] (untouched when S <= 192).
 */
int mofo_attn_fwd(const mofo_bf16* qkv, int B, int S, int H, float scale, mofo_bf16* out, mofo_bf16* out_lo, float* lse,
                  void* stream);
int mofo_attn_bwd(const mofo_bf16* qkv, const mofo_bf16* out, const mofo_bf16* out_lo, const mofo_bf16* dout, const float* lse,
                  int B, int S, int H, float scale, mofo_bf16* dqkv, float* delta, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * (5) LayerNorm (eps 1e-6; nn.LayerNorm at modeling_finetune.py:200,206, modeling_pretrain.py:51,123).
 * Row selection: logical row m reads x row (m / group_rows) * in_group_rows + in_row_offset + m % group_rows
 * (group_rows = M, in_group_rows = M, offset 0 for a plain call; decoder.norm on x[:, -N_mask:] uses
 * group_rows = N_mask, in_group_rows = N, offset = N - N_mask; modeling_pretrain.py:156).
 * fwd: y bf16 [M,D] (dense), mean/rstd f32 [M].
 * bwd: dx = LN'(dy) (+ dres[row] if dres != NULL) written to dx_f32 / dx_bf16 at the mapped x rows (either may be
 *      NULL); dgamma/dbeta f32 [D] (16-byte aligned) are accumulated (+=) with one 16-byte vector reduction per
 *      4 columns per CTA.  bf16_row_scale (may be NULL; f32, one value per group of group_rows logical rows): the bf16
 *      copy is written as bf16(scale * dx) while dx_f32 stays unscaled - the gradient entering a DropPath-scaled
 *      residual branch (the branch's GEMMs read the bf16 copy, the residual path the f32 one).
 */
int mofo_layernorm_fwd(const float* x, const float* gamma, const float* beta, int M, int D, float eps,
                       int group_rows, int in_group_rows, int in_row_offset, mofo_bf16* y, float* mean, float* rstd,
                       void* stream);
int mofo_layernorm_bwd(const mofo_bf16* dy, const float* x, const float* gamma, const float* mean, const float* rstd,
                       const float* dres, int M, int D, int group_rows, int in_group_rows, int in_row_offset,
                       float* dx_f32, mofo_bf16* dx_bf16, float* dgamma, float* dbeta, const float* bf16_row_scale,
                       void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * (6) Decoder input assembly, masked rows (modeling_pretrain.py:260-263): for each clip b and masked slot j,
 * x_full[b, n_vis + j, :] = mask_token + pos[msk_idx[b,j], :]  (f32).  The visible rows are written by the
 * encoder_to_decoder GEMM (BIAS_POS_F32 epilogue).  bwd: dmask_token[Dd] += sum over b,j of dx_full[b, n_vis+j, :],
 * and dvis bf16 [B*n_vis, Dd] = dx_full[b, j<n_vis, :] (the gradient entering encoder_to_decoder).
 */
int mofo_decoder_assemble_fwd(const float* mask_token, const float* pos, const int32_t* msk_idx, int B, int n_vis,
                              int n_msk, int Dd, float* x_full, void* stream);
int mofo_decoder_assemble_bwd(const float* dx_full, int B, int n_vis, int n_msk, int Dd, float* dmask_token,
                              mofo_bf16* dvis, void* stream);
/* mofo_zero_rows: zeroes the first n_zero rows of every group of group_rows rows of x_f32 / x_bf16 [groups*group_rows, D]
 * (either may be NULL).  The head of the decoder sees only the masked rows (x[:, -N_mask:], modeling_pretrain.py:156), so
 * the gradient entering the last decoder block is zero on each clip's visible rows. */
int mofo_zero_rows(float* x_f32, mofo_bf16* x_bf16, int groups, int group_rows, int n_zero, int D, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * (6b) Token mean pooling of the finetuning classifier: VisionTransformer.forward_features ends with
 * fc_norm(x.mean(1)) (modeling_finetune.py:398-401).  fwd: pooled f32 [B, D] = mean over the N tokens of x f32 [B, N, D]
 * (16-byte aligned; deterministic: fixed summation order, no atomics).  bwd: dx[b, n, :] = dpooled[b, :] / N for every token, written as
 * f32 and / or bf16 [B*N, D] (either may be NULL) - the gradient entering the last transformer block.
 */
int mofo_token_mean_fwd(const float* x, const float* weights /* f32 [B,N] or NULL (= 1/N) */, int B, int N, int D, float* pooled,
                        void* stream);
int mofo_token_mean_bwd(const float* dpooled, const float* weights, int B, int N, int D, float* dx_f32, mofo_bf16* dx_bf16,
                        const float* bf16_row_scale /* f32 [B] or NULL: bf16 copy = bf16(scale[b] * dx) */, void* stream);

/* (6d) Key-masked softmax for the cross attention of VisionTransformer_BB_focused, fusing 'MCA' (modeling_finetune.py:100-160,
 * 575-583): queries = the tokens in the box, keys / values = the tokens outside it, 3 heads of embed_dim / 3 (256).  The
 * score and value products run as mofo_gemm_tn / mofo_gemm_wgrad calls on per-clip, per-head views; these two kernels sit
 * between them.  fwd: P bf16 [B * rows_per_b, Nk] = softmax over k of scale * S f32 (`q * self.scale`, :150) restricted to
 * the keys with key_allowed u8 [B, Nk] != 0, exactly 0 elsewhere.  bwd: dS = scale * P * (dP - sum_k P * dP).
 * Nk % 4 == 0, Nk <= 2048.  mofo_cast_f32_bf16: strided f32 -> bf16 copy (the f32 dK / dV the wgrad kernel accumulates -> the
 * bf16 operand of the next GEMM). */
int mofo_masked_softmax_fwd(const float* S, const uint8_t* key_allowed, int B, int rows_per_b, int Nk, float scale, mofo_bf16* P,
                            void* stream);
int mofo_masked_softmax_bwd(const mofo_bf16* P, const float* dP, int64_t rows, int Nk, float scale, mofo_bf16* dS, void* stream);
int mofo_cast_f32_bf16(const float* src, int lds, int M, int N, mofo_bf16* dst, int ldd, void* stream);

/* (6c) Box-focused pooling of VisionTransformer_BB_focused (modeling_finetune.py:589-630 and 555-585): inbox u8 [B, N] = the
 * tokens whose tube touches the per-frame box (boxes int64 [B, frames, 4] = x1,y1,x2,y2 with Python-slice semantics) - the
 * closed form of the reference's all-ones Conv3d over a painted clip - and weights f32 [B, N] (may be NULL) such that
 * sum_n weights[b,n] * x[b,n,:] is the pooled feature: mode 0 ('org') the plain mean, mode 1 ('weighted_mean')
 * (mean_in + 0.5 * mean_out) / 2, mode 2 ('soft_attn' as written, :282-303: its broadcast reduces to) mean_in + mean_out;
 * the plain mean when no token is in the box. */
int mofo_box_tokens(const int64_t* boxes, int B, int frames, int size, int mode, uint8_t* inbox, float* weights, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * (7) Target + loss (engine_for_pretraining.py:258-304): un-normalise with ImageNet mean/std (:260-265), patchify
 * 'b c (t p0) (h p1) (w p2) -> b (t h w) (p0 p1 p2) c' (:268), per-(tube,channel) mean / unbiased std + 1e-6
 * (:269-270), flatten to feature f = p*3 + c (:276), gather the masked tubes (:285-286), MSE mean (:301-304).
 * One CTA per masked tube; the labels are never materialised.
 *   pred bf16 [B*n_msk, 1536] (model output), msk_idx [B, n_msk];
 *   loss_partials f32 [B*n_msk] = per-tube sum of squared errors; loss f32[1] = mean (second tiny kernel);
 *   dpred bf16 [B*n_msk, 1536] = dloss/dpred * grad_scale = 2 (pred - label) / (B*n_msk*1536) * grad_scale
 *   (NULL to skip); labels_out f32 [B*n_msk,1536] optional debug/test output (NULL in production).
 * normalize_target = 0 selects the reference's un-normalised branch (:280).
 */
int mofo_target_mse(const float* video, const int32_t* msk_idx, const mofo_bf16* pred, int B, int n_msk, int frames,
                    int size, int normalize_target, float grad_scale, float* loss_partials, float* loss,
                    mofo_bf16* dpred, float* labels_out, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * (8) Small dense helpers used by the step.
 * mofo_cast_weight: f32 master weight W[R,C] -> bf16 W (row-major, may be NULL) and bf16 W^T [C,R] (may be NULL).
 * mofo_pack_qkv_bias: out f32[3*D] = [q_bias, 0, v_bias]   (modeling_finetune.py:82-84).
 * mofo_colsum_bf16: out f32[N] += column sums of a bf16 [M,N] matrix (bias gradients; caller zeroes out).
 * mofo_sq_norm_f32: out f32[1] += sum(x^2) over n elements (gradient-norm of the arena, utils.py:376-388).
 */
int mofo_cast_weight(const float* W, int R, int C, mofo_bf16* W_bf16, mofo_bf16* Wt_bf16, void* stream);
int mofo_pack_qkv_bias(const float* q_bias, const float* v_bias, int D, float* out, void* stream);
int mofo_colsum_bf16(const mofo_bf16* X, int ldx, int M, int N, float* out, void* stream);
int mofo_sq_norm_f32(const float* x, int64_t n, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * (8b) On-GPU input normalisation (first piece): uint8 clip [B,3,frames,size,size] (NCTHW) ->
 * f32 (x/255 - mean_c)/std_c with the ImageNet constants, bit-identical to ToTorchFormatTensor(div=True) +
 * GroupNormalize (datasets.py:44-50, transforms.py:346-382).  Lets the host ship 1 byte per sample instead of 4.
 */
int mofo_normalize_u8(const uint8_t* clip_u8, int B, int frames, int size, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * (8c) Clip preprocessing on the GPU: what the reference does per sample on the CPU in
 * GroupMultiScaleCrop_BB_no_global_union.__call__ (transforms.py:103-135: albumentations Crop + Resize(224, bilinear) with
 * the pascal_voc motion box riding along) -> Stack -> ToTorchFormatTensor(div=True) -> GroupNormalize (datasets.py:44-50).
 * frames: uint8 [B,T,H,W,3] (decoded RGB frames, HWC); crops: int32 [B,4] = (x_off, y_off, crop_w, crop_h) per clip (chosen
 * on the host as transforms.py:137-159 does); boxes_in f64 [B,T,4] (x1,y1,x2,y2 in frame pixels; may be NULL).
 * clip_out: f32 [B,3,T,out,out] (NCTHW), bit-identical to cv2.resize(INTER_LINEAR) + /255 + (x-mean)/std;
 * boxes_out: f64 [B,T,4] in output pixels, [0,0,1,1] where the crop leaves nothing of the box (transforms.py:120-123).
 * crop_w / crop_h must fit inside the frame.  The host ships 1 byte per source pixel instead of 4 per output sample.
 */
int mofo_clip_preprocess(const uint8_t* frames, int B, int T, int H, int W, const int32_t* crops, const double* boxes_in,
                         int out_size, float* clip_out, double* boxes_out, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * (9) Fused AdamW over the flat arenas.  Replaces torch.optim.AdamW.step() as created by
 * optim_factory.create_optimizer (optim_factory.py:126-127) and driven by utils.py:355-364 (clip + step), plus the
 * per-step fp32->bf16 operand casts.  params / grads / exp_avg / exp_avg_sq are f32 arenas with identical layout.
 * segs  : int64 [n_seg][6] = {arena offset, rows, cols, param-group index, w16 offset or -1, wt16 offset or -1}
 *         (offsets into w16 in elements: bf16 copy W[rows,cols] and transposed copy W^T[cols,rows]).
 * tiles : int32 [n_tiles][2] = {segment, tile}: a tile is a 32 x 128 block (row-major tile order, ceil(cols/128) tiles
 *         per tile row) when the segment has a transposed copy, else 4096 consecutive elements.
 * hyper : f32 device array {beta1, beta2, eps, 1-beta1^t, sqrt(1-beta2^t), 0, 0, 0, lr_0, wd_0, lr_1, wd_1, ...}.
 * clip_coef (may be NULL): device scalar multiplied into every gradient (gradient clipping, utils.py:358).
 * loss_guard (may be NULL): device scalar; when it is NaN/Inf the whole update is skipped (the engine then exits as
 * engine_for_pretraining.py:418-420 does, with parameters untouched).
 * sq_norm_out (may be NULL): device scalar; += sum of squares of the (unclipped) gradients of the tiles processed - the
 * gradient norm of utils.py:376-388 without a second pass over the gradient arena (only usable when no clip coefficient
 * derived from that norm is needed first).
 * Update rule = torch AdamW: p *= 1-lr*wd; m = lerp(m,g,1-b1); v = b2 v + (1-b2) g^2; p -= lr/bc1 * m/(sqrt(v)/sqrt(bc2)+eps).
 */
int mofo_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, mofo_bf16* w16,
                    const int64_t* segs, const int32_t* tiles, int n_tiles, const float* hyper, const float* clip_coef,
                    const float* loss_guard, float* sq_norm_out, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * (10) Motion-box preprocessing, pixel stages.  The reference runs these once per dataset on the CPU with
 * numpy / scipy.ndimage / cv2; both calls are bit-exact with that arithmetic (oracle/motion_oracle.py).
 *
 * mofo_motion_map: optical-flow video -> motion map.  Replaces the frame loop of create_lmdb_video_dataset_flow_mag
 * (scripts/data/motion_map_creator.py:160-205) and compute_motion_boudary / zero_boundary (scripts/motion_sts.py:5-37).
 * flows: uint8 [T,H,W,C] (C >= 2; channel 0 = u, 1 = v; the reference decodes C = 3).  For frame idx = 1..T the ws-frame
 * window of :163-172, per channel the window sum of the 3x3 motion-boundary responses (ndimage.convolve, 'reflect'),
 * float32 magnitude as cv2.cartToPolar computes it, (mag_u + mag_v) / 2, a `border`-pixel frame zeroed (the reference: 8;
 * 0 = none), cast to uint8 as ndarray.astype does (trunc mod 256), written to out[T,H,W,out_channels] with the value
 * replicated over the channels (the reference writes 3).
 *
 * mofo_motion_box_filter: the per-frame filtering in front of the contour search of json_creator
 * (scripts/data/SSV2/bounding_box_creator_SSV.py:125-166; the Epic-Kitchens variant is the same code): for every frame of
 * frames uint8 [T,H,W,3]: scipy gaussian_filter(sigma_before) over all three axes with uint8 storage after each 1-D pass,
 * zero below remove_thrd * max, zero below std_k * (std + std_eps), gaussian_filter(sigma_after), cv2 BGR2GRAY.
 * w_before[0..r_before] / w_after[0..r_after]: float64 gaussian weights by distance from the centre, as scipy's
 * _gaussian_kernel1d(sigma, 0, int(4*sigma+0.5)) gives them (device pointers).  work: 2*T*H*W*3 bytes; stats: uint64 [T,4]
 * (cleared by the call; afterwards [t] = {max, sum, sum of squares, -}).  Outputs: filtered uint8 [T,H,W,3] (the frame
 * the reference hands to cvtColor) and gray uint8 [T,H,W] (what cv2.findContours receives).  frames is not modified.
 * The contour search, ranking and temporal smoothing (:168-475) are sequential host code and stay on the host.
 */
int mofo_motion_map(const uint8_t* flows, int T, int H, int W, int C, int ws, int border, uint8_t* out, int out_channels,
                    void* stream);
int mofo_motion_box_filter(const uint8_t* frames, int T, int H, int W, const double* w_before, int r_before, const double* w_after,
                           int r_after, double remove_thrd, double std_k, double std_eps, uint8_t* work, uint64_t* stats,
                           uint8_t* filtered, uint8_t* gray, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MOFO_B200_H_ */
