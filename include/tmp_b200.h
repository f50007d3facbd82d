/* tmp_b200.h -- C ABI of libtmp_b200.so: the B200 (sm_100a) kernels under the `tri_mbt_vsltcls` training hot path.
 *
 * The reference (AITRICS/Medical_Tri_Modal_Pilot) has no native layer at all: its "operator API" is Python
 * (SURVEY.md 8b). Each entry point below names the reference code it replaces (file:line under the reference
 * root). The Python host (medical_tri_modal_pilot_b200/) binds these with ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions: every pointer is a DEVICE pointer unless stated; tensors are row-major contiguous; 16-bit buffers
 * are passed as void*. PRECISION: every 16-bit tensor on the product path is IEEE fp16 (the reference's autocast
 * dtype, trainer.py:126; tcgen05 kind::f16 needs both MMA operands in one format); gradient tensors carry a
 * caller-chosen power-of-two scale (the kernels are linear in it); accumulation, parameters and parameter
 * gradients are fp32. The GEMM entry points also accept bf16 x bf16 through their `*_fmt` arguments
 * (TMP_FMT_F16 = 0, TMP_FMT_BF16 = 1; A and B must use the same format). Every call is asynchronous on `stream` (a cudaStream_t), allocates nothing,
 * keeps no global state and never synchronises. Return 0 on success, >0 = cudaError_t, <0 = argument/driver
 * error; tmp_last_error() returns the message (thread-local). D = 256 channels, H = 4 heads of 64 everywhere
 * (control/config.py:97,99 defaults; the kernels are specialised for them).
 * CUDA GRAPHS: every launch is capturable. The only per-step scalars of a training step are read from DEVICE words
 * when the caller supplies them, so that one captured graph serves every step: `seed_dev` (uint32, added to the
 * scalar dropout `seed`; NULL = scalar only) and tmp_adamw_step_dev's `lr_dev` / `step_dev`.
 */
#ifndef TMP_B200_H
#define TMP_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define TMP_FMT_F16 0
#define TMP_FMT_BF16 1

int tmp_abi_version(void);
const char* tmp_last_error(void);

/* ---- e: SMs left to communication kernels -----------------------------------------------------------------
 * Every persistent / one-wave grid of the library is sized from (physical SMs - reserved) with equal work per CTA.
 * Data-parallel runs reserve a few SMs for NCCL's all-reduce CTAs (GradSync, trainer.py; the reference has no
 * multi-GPU path): a compute kernel that finds SMs taken would otherwise run its last CTAs as a second
 * wave. n in [0, 64); also read once from env TMP_B200_RESERVE_SMS. Call before capturing a CUDA graph of the step.
 * tmp_num_sms: the count grid sizing currently uses. */
int tmp_set_reserved_sms(int n);
int tmp_num_sms(void);

/* ---- a3/a4/a14: lengths and masks ------------------------------------------------------------------------
 * kv_len[3,B] int32 = number of attendable keys per (stream, sample) INCLUDING the 4 bottleneck keys and CLS.
 * Replaces get_attn_pad_mask/get_non_pad_mask (builder/models/src/transformer/utils.py:79-125) as called from
 * TrimodalTransformerEncoder_MBT.forward (mbt_encoder.py:703-714, 748) and the image length code in
 * tri_mbt_vsltcls.py:226-237. `missing` = trainer.py:68-84 code (0 all, 1 txt missing, 2 img missing, 3 both);
 * with skip_missing the de-selected streams of a sample get kv_len 0 (exact: SURVEY.md Appendix A). */
int tmp_build_lengths(const long long* input_lengths, const long long* txt_lengths, const float* img_time, int n_img,
                      int multiimages, const long long* missing, int skip_missing, int B, int T_v, int T_i, int T_t,
                      int32_t* kv_len, void* stream);
/* test-only: mask[b,q,k] = (k >= kv_len[b]) as uint8 [B,T,T], the tensor the reference materialises. */
int tmp_debug_materialize_mask(const int32_t* kv_len, int B, int T, uint8_t* mask, void* stream);

/* ---- a1: UMSE / TIE embedding (tri_mbt_vsltcls.py:183-190) -------------------------------------------------
 * x[n_tok,3] fp32 (time, value, feature-id-as-float) -> E[n_tok,256] (fp32 or fp16).
 * val4/tim4: HOST arrays of 4 device pointers {Linear.weight[256], Linear.bias, LayerNorm.weight, LayerNorm.bias}
 * of ie_vslt / ie_time; Wfeat = ie_feat.weight[20,256]. */
int tmp_umse_embed_fwd(const float* x, long long n_tok, const float* const* val4, const float* const* tim4,
                       const float* Wfeat, void* out, int out_is_fp16, void* stream);

/* ---- a1+a2+a5: stream prologue --------------------------------------------------------------------------
 * X0[B, 5+n, 256] fp16 = [bottlenecks(4); Dropout(LN_in([CLS; E]) (+PE))]   (mbt_encoder.py:697-699, 719-729)
 * kind 0 (vslt): E from x[B,n,3] as above.  kind 1 (img/txt): E = proj[B*n,256] (fp16) + ie_time(times[b, j / (n/n_slots)])
 * + ie_feat[feat_id]   (tri_mbt_vsltcls.py:216-224).  pe = positional_encoding.pe rows (txt only) or NULL. */
int tmp_stream_prologue_fwd(int kind, int B, int n, const float* x, const float* const* val4, const void* proj,
                            const float* times, int n_slots, int feat_id, const float* const* tim4, const float* Wfeat,
                            const float* cls, const float* bottlenecks, const float* ln_g, const float* ln_b,
                            const float* pe, float drop_p, uint32_t seed, uint32_t salt, const uint32_t* seed_dev,
                            void* X0, void* stream);
/* gradient accumulators are fp32 and ADDED to: g_val/g_tim [4,256] (dW, db, dLN.w, dLN.b), g_feat[20,256],
 * g_cls[256], g_bott[4,256], g_ln[2,256]; dX0[B,5+n,256] and dproj[B*n,256] (written, kind 1) are fp16. */
int tmp_stream_prologue_bwd(int kind, int B, int n, const float* x, const float* const* val4, const void* proj,
                            const float* times, int n_slots, int feat_id, const float* const* tim4, const float* Wfeat,
                            const float* cls, const float* bottlenecks, const float* ln_g, const float* ln_b,
                            const float* pe, float drop_p, uint32_t seed, uint32_t salt, const uint32_t* seed_dev,
                            const void* dX0, float* g_val, float* g_tim, float* g_feat, float* g_cls, float* g_bott,
                            float* g_ln, void* dproj, void* stream);

/* ---- a9: LayerNorm (module.py:130-144: unbiased std, eps on std) ------------------------------------------
 * fwd: if add != NULL: sum_out = x + add, y = LN(sum_out) (encoder.py:27-30 residual fused); else y = LN(x).
 * bwd: dx = dres + dLN(dy; x); dgamma/dbeta fp32 += ; optional dx_drop = dropout(seed,salt)(dx). rows of 256.
 * All 16-bit tensors fp16. */
int tmp_layernorm_fwd(const void* x, const void* add, const float* gamma, const float* beta, long long rows,
                      void* sum_out, void* y, void* stream);
int tmp_layernorm_bwd(const void* dy, const void* x, const void* dres, const float* gamma, long long rows, void* dx,
                      void* dx_drop, float drop_p, uint32_t seed, uint32_t salt, const uint32_t* seed_dev, float* dgamma,
                      float* dbeta, void* stream);
/* The LayerNorm backward in front of the attention backward (dx is the attention's dO): additionally writes
 * delta[B,4,T_lse] = per-head sum_d dO.O (attn_O = attention output of the forward) and zeroes the dQ columns [0,256) of
 * dQKV [rows,768], so that tmp_mma_attn_bwd can run in its fused protocol (dQ_acc == NULL). rows = B*T. */
int tmp_layernorm_bwd_attn(const void* dy, const void* x, const void* dres, const float* gamma, long long rows, void* dx,
                           float* dgamma, float* dbeta, const void* attn_O, int T, int T_lse, float* delta, void* dQKV,
                           void* stream);

/* ---- a10/a11: tcgen05 GEMMs -------------------------------------------------------------------------------
 * out[M,N] = residual + dropout( gate>0 ? act(alpha * A[M,K].B[N,K]^T + bias) : 0 ),  act = relu: 0 none, 1 ReLU, 2 GELU(erf)
 * A, B 16-bit K-major, both fp16 or both bf16 (B = nn.Linear / Conv1d(k=1) weight `[out,in]`: attention.py:68-70,
 * module.py:74-80; dgrad passes the gradient as A and the transposed weight copy as B).
 * N % 128 == 0, K % 64 == 0. Any of bias/gate/residual may be NULL; out16 (in out_fmt) and/or out_f32 get the result.
 * mask_out (optional, [M, N/32] uint32): bit c%32 of word (row, c/32) = (result before the residual > 0) -- the ReLU /
 * dropout pattern in one bit per element. A later call passes it back as `gate` with gate_fmt = 3 and ld_gate = N/32
 * (the FFN2 input gradient) instead of re-reading the 16-bit activation. gate_fmt / res_fmt 2 = fp32 tensors (fp32 mode).
 * row_live (optional, uint8 per group of rows_per_group consecutive rows): output tiles whose rows all lie in dead groups
 * are skipped (images without a consumer in the Swin feed).
 * A, B and out16 go through TMA: 16-byte aligned base addresses, leading dimensions multiples of 8 elements. */
int tmp_gemm_bias_act_fwd(const void* A, int a_fmt, int lda, const void* B, int b_fmt, int ldb, int M, int N, int K,
                          float alpha, const float* bias, int relu, const void* gate, int gate_fmt, int ld_gate,
                          const void* residual, int res_fmt, int ld_res, float drop_p, uint32_t seed, uint32_t salt,
                          const uint32_t* seed_dev, void* out16, int out_fmt, float* out_f32, int ld_out,
                          uint32_t* mask_out, const uint8_t* row_live, int rows_per_group, void* stream);
/* dW[N,K] fp32 += dY[M,N]^T . X[M,K]  (weight gradient; N,K % 128 == 0; same format for dY and X).
 * dbias (optional, may be NULL): dbias[N] fp32 += column sums of dY (the bias gradient), computed from the dY tiles
 * the kernel stages in shared memory anyway -- replaces a separate tmp_colsum pass over dY.
 * dW is contiguous [N,K] and 16-byte aligned (its tiles are added with TMA reduce operations). */
int tmp_gemm_wgrad(const void* dY, int y_fmt, int ldy, const void* X, int x_fmt, int ldx, int M, int N, int K,
                   float* dW, float* dbias, void* stream);
/* out[N] fp32 += column sums of dY[M,N] fp16 (bias gradient; the 16-bit path uses the fused form in tmp_gemm_wgrad) */
int tmp_colsum(const void* dY, int ld, long long M, int N, float* out, void* stream);

/* ---- a10: modality-aware attention (attention.py:24-49, 65-84) ----------------------------------------------
 * tmp_mma_attn_bwd protocols: dQ_acc != NULL -> stand-alone (delta computed inside, dQ accumulated in the fp32 workspace
 * dQ_acc, converted at the end); dQ_acc == NULL -> fused: delta already written and the dQ columns of dQKV already zeroed
 * by tmp_layernorm_bwd_attn, dQ tiles are reduce-added in fp16 in place.
 * qkv[B*T,768] fp16 = Q|K|V with head h at columns h*64; kv_len[B] (or NULL = unmasked);
 * O[B*T,ld_o] fp16; lse2[B,H,T_lse] fp32 (log2-domain logsumexp of the scaled scores, kept for backward). */
/* q_rows (both calls): only the leading q_rows query rows of every sample are computed / carry a gradient (rounded up to
 * whole 128-row tiles; pass T for all). The last fused layer under --mbt-only-vslt 1 only consumes the CLS row. */
int tmp_mma_attn_fwd(const void* qkv, const int32_t* kv_len, int B, int T, int H, void* O, int ld_o, float* lse2,
                     int T_lse, int q_rows, void* stream);
/* qkv, O, dO fp16; delta[B,H,T_lse] and dQ_acc[B*T,256] fp32 are workspaces; dQKV[B*T,768] fp16 receives
 * dQ|dK|dV. T_lse % 128 == 0. */
int tmp_mma_attn_bwd(const void* qkv, const void* O, const void* dO, int ld, const int32_t* kv_len, int B, int T, int H,
                     const float* lse2, int T_lse, float* delta, float* dQ_acc, void* dQKV, int q_rows, void* stream);

/* Single-query form of the attention backward: the gradient enters through ONE query row `q_row` per sample (the CLS row of
 * the last fused layer under --mbt-only-vslt 1, mbt_encoder.py:757-763). dO_row / O_row [B,256] fp16 = that row of dO / of the
 * forward output for every sample; every element of dQKV [B*T,768] is written (zeros where nothing flows). */
int tmp_attn_bwd_single_query(const void* qkv, const void* dO_row, const void* O_row, const int32_t* kv_len, int B, int T,
                              int H, int q_row, const float* lse2, int T_lse, void* dQKV, void* stream);

/* ---- a7: bottleneck exchange (mbt_encoder.py:764-776), in place on rows 0..3 of Y_m[B,T_m,256]
 * (fp16) ---------------------------------------------------------------------------------------------------- */
int tmp_bottleneck_mix_fwd(void* Yv, void* Yi, void* Yt, int Tv, int Ti, int Tt, const long long* missing, int B,
                           void* stream);
/* backward: dYd_m (each optional, all ignored when drop_p == 0) additionally receive rows 0..3 of dY_m after dropout with
 * the mask of (seed [+ *seed_dev], salt_m, element index in the stream's [B*T_m,256] matrix) -- the rows this call changes in
 * a tensor whose dropped copy tmp_layernorm_bwd(dx_drop) has already written. */
int tmp_bottleneck_mix_bwd(void* dYv, void* dYi, void* dYt, int Tv, int Ti, int Tt, int upper_has_img_txt,
                           const long long* missing, int B, void* dYd_v, void* dYd_i, void* dYd_t, float drop_p,
                           uint32_t seed, const uint32_t* seed_dev, uint32_t salt_v, uint32_t salt_i, uint32_t salt_t,
                           void* stream);

/* ---- helpers ---------------------------------------------------------------------------------------------- */
int tmp_dropout_apply(const void* in, void* out, long long n, float drop_p, uint32_t seed, uint32_t salt,
                      const uint32_t* seed_dev, void* stream);
/* fp16 gradient tensors, n % 8 == 0.
 * tmp_cast_weights: descs = device array of n_desc records {const float* src; fp16* dst; fp16* dst_t; int R; int C}
 * (32 bytes): dst[R,C] = fp16(src), dst_t[C,R] = fp16(src)^T (either may be NULL). */
int tmp_cast_weights(const void* descs, int n_desc, int max_R, int max_C, void* stream);


/* ---- a13 (optimizer tail): torch.optim.AdamW semantics (reference 2_train.py:110) over the
 * flat fp32 parameter / gradient buffers of the fused path: w,g,m,v fp32 [n], n % 4 == 0; step >= 1. ---------- */
int tmp_adamw_step(float* w, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                   float eps, float weight_decay, int step, void* stream);
/* same update with the learning rate and the step count read from DEVICE words -- the form a captured CUDA graph replays.
 * lr_dev: fp32. step_dev: int32[3] = {optimizer calls so far (>= 1, incremented by the caller before this call), calls
 * skipped, last call whose gradient was non-finite}: when step_dev[2] == step_dev[0] the call leaves w, m, v untouched and
 * increments step_dev[1] (only if count_skip != 0: a second launch of the same step over another buffer passes 0); bias
 * corrections use step_dev[0] - step_dev[1] (evaluated in the kernel). */
int tmp_adamw_step_dev(float* w, const float* g, float* m, float* v, long long n, const float* lr_dev, float beta1,
                       float beta2, float eps, float weight_decay, int32_t* step_dev, int count_skip, void* stream);
/* state[2] <- state[0] if any of g[0..n) (fp32, n % 4 == 0) is inf or NaN: the overflow guard of the 16-bit plan (gradients
 * pass through fp16 scratch with a static scale). Call it on the fully reduced gradient, before tmp_adamw_step_dev. */
int tmp_grad_nonfinite(const float* g, long long n, int32_t* state, void* stream);

/* ---- a12 / f3: classifier head in training mode (reference tri_mbt_vsltcls.py:176-177 demographic branch, :248-255 head;
 * module definitions :72-76, :152-158). One forward launch, two backward launches, all fp32, no floating-point atomics
 * (cross-CTA sums go through per-CTA partials added in a fixed order by the last CTA):
 *   logit = Linear(256,1)(ReLU(BatchNorm1d(Linear(512,256)([LayerNorm(cls) | ReLU(LayerNorm(Linear(2,256)([age, gender])))]))))
 * params / grads: 12 device pointers in the order layer_norms_after_concat.{weight,bias}, ie_demo.0.{weight[256,2],bias},
 *   ie_demo.1.{weight,bias}, fc_list.0.{weight[256,512],bias}, fc_list.1.{weight,bias}, fc_list.3.{weight[1,256],bias[1]}.
 * saved: 7 device pointers written by the forward and read by the backward: Z [B,512], XC [B,256], XD [B,256], rstd_c [B],
 *   rstd_d [B], XH [B,256], invstd [256]. run_mean / run_var [256]: BatchNorm running statistics, updated in place with
 *   `momentum` (unbiased variance), *nbt (int64 num_batches_tracked, may be NULL) += 1. 2 <= B <= 4096 (batch statistics).
 * scratch: >= max(32 * B, 64 * 7 * 256) floats; counter: one uint32, zero before the first call, left at zero.
 * tmp_head_bwd OVERWRITES the 12 gradient tensors and dcls [B,256]; DH: [B,256] scratch. */
int tmp_head_fwd(const float* cls, const float* age, const float* gen, int B, const void* const* params, float* run_mean,
                 float* run_var, long long* nbt, float momentum, float bn_eps, void* const* saved, float* scratch,
                 unsigned int* counter, float* logits, void* stream);
int tmp_head_bwd(const float* dlogit, const float* age, const float* gen, int B, const void* const* params,
                 void* const* saved, void* const* grads, float* dcls, float* DH, float* scratch, unsigned int* counter,
                 void* stream);

/* ---- image-encoder feed: the glue of the frozen Swin-T forward around the tcgen05 GEMMs
 * (reference builder/models/src/swin_transformer.py: patch embedding :541-551, SwinTransformerBlock.forward :447-450,
 * shifted_window_attention :115-214, PatchMerging :34-46,75-86). Activations fp16 [tokens, Cp] (Cp = padded channel
 * stride, pad channels zero); LayerNorm parameters fp32; n_img images; H x W token map; C real channels. -------------- */
/* img fp32 [n_img,224,224]; Wt [16,96] = conv weight transposed (pixel-major); out [n_img*56*56, Cp] */
/* live (every tmp_swin_* call): optional uint8 [n_img]; 0 = the image's features have no consumer (empty slot, img_time == 10,
 * tri_mbt_vsltcls.py:229-231, or an img-missing sample): its work is skipped. The last kernel of the encoder passes
 * zero_dead = 1 so that dead images end up as zero rows (a masked key still meets 0 * V in the attention). */
int tmp_swin_patch_embed_ln(const float* img, int n_img, const float* Wt, const float* bconv, const float* g,
                            const float* b, void* out, int Cp, const uint8_t* live, void* stream);
/* out[window order] = LayerNorm(x[natural order]) after torch.roll(-shift) and window partition (7x7 windows) */
int tmp_swin_ln_window(const void* x, const float* g, const float* b, int n_img, int H, int W, int C, int Cp, int shift,
                       void* out, const uint8_t* live, int zero_dead, void* stream);
/* qkv [tokens(window order), ld_qkv] = q|k|v (head h at h*32 inside each C-wide part); rel_bias fp32 [heads,49,49];
 * out [tokens(window order), ld_out] */
int tmp_swin_window_attn(const void* qkv, int ld_qkv, const float* rel_bias, int n_img, int H, int W, int C, int heads,
                         int shift, void* out, int ld_out, const uint8_t* live, void* stream);
/* x[natural] += y[window order] (window reverse + reverse shift); hn = LayerNorm(x) unless hn == NULL */
int tmp_swin_unwindow_add_ln(const void* y, void* x, const float* g, const float* b, int n_img, int H, int W, int C,
                             int Cp, int shift, void* hn, const uint8_t* live, void* stream);
/* out[(n,i,j), 4C] = LayerNorm(x0|x1|x2|x3) of the 2x2 neighbourhood (PatchMerging), out stride 4C */
int tmp_swin_merge_ln(const void* x, const float* g, const float* b, int n_img, int H, int W, int C, int Cp, void* out,
                      const uint8_t* live, void* stream);

/* ---- fp32 ("precise") mode: north-star's FP32 parity mode of the same path (trainer.py:126 is the only place the
 * reference drops precision; called outside autocast the reference runs in fp32). Every activation / gradient tensor is
 * fp32; the *_f32 operators are the kernels above compiled for fp32 storage (same arguments, `float*` tensors).
 * GEMMs: tmp_split_bf16x3 turns an fp32 operand into three bf16 terms laid out as six blocks along the reduction
 * dimension (A side: hi|hi|mid|hi|lo|mid, B side: hi|mid|hi|lo|hi|mid), then tmp_gemm_bias_act_fwd / tmp_gemm_wgrad run on
 * the bf16 blocks (a_fmt = b_fmt = 1) with fp32 accumulation, fp32 output (out_f32) and fp32 gate / residual (format 2):
 * A . B^T to ~2^-17 relative on the tensor cores. Attention runs on the CUDA cores in fp32 (no atomics: deterministic). */
int tmp_layernorm_fwd_f32(const float* x, const float* add, const float* gamma, const float* beta, long long rows,
                          float* sum_out, float* y, void* stream);
int tmp_layernorm_bwd_f32(const float* dy, const float* x, const float* dres, const float* gamma, long long rows, float* dx,
                          float* dx_drop, float drop_p, uint32_t seed, uint32_t salt, const uint32_t* seed_dev,
                          float* dgamma, float* dbeta, void* stream);
int tmp_bottleneck_mix_fwd_f32(float* Yv, float* Yi, float* Yt, int Tv, int Ti, int Tt, const long long* missing, int B,
                               void* stream);
int tmp_bottleneck_mix_bwd_f32(float* dYv, float* dYi, float* dYt, int Tv, int Ti, int Tt, int upper_has_img_txt,
                               const long long* missing, int B, float* dYd_v, float* dYd_i, float* dYd_t, float drop_p,
                               uint32_t seed, const uint32_t* seed_dev, uint32_t salt_v, uint32_t salt_i, uint32_t salt_t,
                               void* stream);
int tmp_dropout_apply_f32(const float* in, float* out, long long n, float drop_p, uint32_t seed, uint32_t salt,
                          const uint32_t* seed_dev, void* stream);
int tmp_colsum_f32(const float* dY, int ld, long long M, int N, float* out, void* stream);
int tmp_stream_prologue_fwd_f32(int kind, int B, int n, const float* x, const float* const* val4, const float* proj,
                                const float* times, int n_slots, int feat_id, const float* const* tim4,
                                const float* Wfeat, const float* cls, const float* bottlenecks, const float* ln_g,
                                const float* ln_b, const float* pe, float drop_p, uint32_t seed, uint32_t salt,
                                const uint32_t* seed_dev, float* X0, void* stream);
int tmp_stream_prologue_bwd_f32(int kind, int B, int n, const float* x, const float* const* val4, const float* proj,
                                const float* times, int n_slots, int feat_id, const float* const* tim4,
                                const float* Wfeat, const float* cls, const float* bottlenecks, const float* ln_g,
                                const float* ln_b, const float* pe, float drop_p, uint32_t seed, uint32_t salt,
                                const uint32_t* seed_dev, const float* dX0, float* g_val, float* g_tim, float* g_feat,
                                float* g_cls, float* g_bott, float* g_ln, float* dproj, void* stream);
/* src [R,C] fp32 (row stride ld elements, C % 4 == 0) -> dst bf16 [R,6C] (stack_rows = 0) or [6R,C] (stack_rows = 1);
 * side_b selects the B-operand term order */
int tmp_split_bf16x3(const float* src, long long ld, long long R, int C, int side_b, int stack_rows, void* dst,
                     void* stream);
/* attention.py:24-49 in fp32: qkv [B*T,768] = Q|K|V, O [B*T,ld_o], lse2 [B,H,T_lse] log2-domain logsumexp */
int tmp_attn_fwd_f32(const float* qkv, const int32_t* kv_len, int B, int T, int H, float* O, int ld_o, float* lse2,
                     int T_lse, void* stream);
/* dQKV [B*T,768] = dQ|dK|dV; delta [B,H,T_lse] workspace */
int tmp_attn_bwd_f32(const float* qkv, const float* O, const float* dO, int ld_o, const int32_t* kv_len, int B, int T,
                     int H, const float* lse2, int T_lse, float* delta, float* dQKV, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TMP_B200_H */
