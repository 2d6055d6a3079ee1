/*
 * b200_lora.h — C ABI of the B200-native LoRA-training hot path.
 *
 * This is the drop-in boundary (SURVEY.md §8b "C-ABI layer"): plain pointers and sizes, no torch
 * types, `int` status codes (0 = ok, <0 = error; text via b200_last_error()), no ownership transfer
 * (the caller owns every buffer), asynchronous on the caller's CUDA stream.  Each entry point names
 * the reference code it replaces (paths relative to ostris/ai-toolkit @ 27a03a9).
 *
 * Conventions
 *   - all matrices are row-major; "ld" = leading dimension in ELEMENTS; bf16 = __nv_bfloat16 bits
 *   - every bf16 matrix that feeds a tensor-core GEMM needs a 16-byte aligned base and ld % 8 == 0
 *   - `stream` is a cudaStream_t passed as void*
 *   - the library targets sm_100a only and has no CPU or library fallback: if the device is not
 *     compute capability 10.x every launch returns B200_ERR_ARCH.
 */
#ifndef B200_LORA_H_
#define B200_LORA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_OK 0
#define B200_ERR_INVALID (-1) /* bad argument (shape / alignment / null) */
#define B200_ERR_CUDA (-2)    /* a CUDA runtime / driver call failed      */
#define B200_ERR_ARCH (-3)    /* device is not sm_100                     */
#define B200_ERR_NCCL (-4)

/* activation / epilogue selectors for b200_gemm_desc.act */
#define B200_ACT_NONE 0
#define B200_ACT_GELU_TANH 1

/* b200_gemm_desc.config */
#define B200_GEMM_AUTO 0
#define B200_GEMM_1CTA_N256 1 /* 128x256 tile, cta_group::1                     */
#define B200_GEMM_2CTA_N256 2 /* 256x256 tile over a CTA pair, cta_group::2     */
#define B200_GEMM_1CTA_N128 3
#define B200_GEMM_1CTA_N64 4 /* skinny (rank-side) GEMM, optional split-K       */
#define B200_GEMM_SKINNY_CLUSTER 5 /* N <= 64, bf16 out: split-K over a CTA cluster, DSMEM reduce */
#define B200_GEMM_1CTA_N160 6 /* 128x160 / 128x192 tiles: AUTO picks them when 128x128 tiles would leave a nearly empty */
#define B200_GEMM_1CTA_N192 7 /* second wave (e.g. SDXL's 2048 x 1280 outputs: 160 tiles on 148 SMs); N160: B K-major only */

typedef struct b200_ctx b200_ctx;

int b200_version(void);
const char* b200_last_error(void);
int b200_ctx_create(b200_ctx** out, int device);
int b200_ctx_destroy(b200_ctx* ctx);
/* number of kernels this library has launched through `ctx` since creation (bench.py's gpu_launches) */
int64_t b200_ctx_launch_count(const b200_ctx* ctx);

/*
 * Tensor-core GEMM with a two-segment contraction and a fused epilogue (tcgen05 / TMEM / TMA):
 *
 *   acc[M,N] = alpha * row_alpha[m / rows_per_sample] * ( op(A0)[M,K0] . op(B0)[N,K0]^T + op(A1)[M,K1] . op(B1)[N,K1]^T )
 *   y   = bf16(acc + bias[n])
 *   aux_out[m,n] = y                         (optional: pre-activation kept for backward)
 *   y   = bf16(gelu_tanh(y))                 (act == B200_ACT_GELU_TANH)
 *   y   = bf16(y * gelu_tanh'(aux_in[m,n]))  (optional: backward through the activation)
 *   y   = bf16(y * gate[m / rows_per_sample, n])   (optional AdaLN-Zero gate)
 *   out[m,n] = bf16(res[m,n] + y)            (optional residual; res may alias out)
 *
 * Operand storage (row-major, ld in elements):
 *   trans_a == 0: A is [M, K] (K contiguous, "K-major");  trans_a != 0: A is stored [K, M] (M contiguous)
 *   trans_b == 0: B is [N, K] (K contiguous);             trans_b != 0: B is stored [K, N] (N contiguous)
 * so  forward  Y  = X W^T        is (trans_a 0, trans_b 0) with B0 = W[out,in]
 *     dgrad    dX = dY W         is (trans_a 0, trans_b 1) with B0 = W[out,in] as stored (no transposed copy)
 *     wgrad    dB = dY^T Z       is (trans_a 1, trans_b 1) with A0 = dY[tokens,out], B0 = Z[tokens,r]
 * Both segments share the same trans flags.
 *
 * The second segment is how the LoRA up-projection rides in the same TMEM tile as the frozen base
 * GEMM:  A1 = bf16(m_b * s * (x A^T)) padded to 64 columns, B1 = lora_up weight padded to 64 columns.
 * Replaces, per wrapped Linear, `org_forward(x) + (lora_up(lora_down(x.float())) * scale * multiplier).to(bf16)`
 * (toolkit/network_mixins.py:304-342) and, with trans_b, the dX half of its autograd backward; the
 * trans_a/trans_b form computes dA / dB (rank-r sides) without ever forming dW.
 *
 * f32_mode 1: write the fp32 accumulator (no epilogue); `splits` partial results of a split-K
 *             contraction go to out + split*M*ldo.
 * f32_mode 2: atomically ADD alpha*acc (fp32) into out[m*ldo+n] (or out[n*ldo+m] when f32_trans), for the
 *             columns n < n_store only; with `splits` > 1 the K range is split across CTAs (gradient
 *             accumulation semantics: the caller zeroes `out` once per optimizer step).
 */
typedef struct b200_gemm_desc {
  int32_t M, N, K0, K1;
  const void* A0; int32_t lda0;
  const void* B0; int32_t ldb0;
  const void* A1; int32_t lda1;
  const void* B1; int32_t ldb1;
  int32_t trans_a, trans_b;
  float alpha;                            /* accumulator scale (set 1.0f for none) */
  const void* row_alpha;                  /* fp32 [ceil(M/rows_per_sample)] per-sample scale or NULL */
  const void* bias;                       /* bf16 [N] or NULL */
  const void* res; int32_t ldres;         /* bf16 [M,N] or NULL */
  const void* gate; int32_t ldgate;       /* bf16 [ceil(M/rows_per_sample), N] or NULL */
  int32_t rows_per_sample;                /* >0 when gate / row_alpha is given */
  const void* aux_in; int32_t ldaux_in;   /* bf16 [M,N] or NULL */
  void* aux_out; int32_t ldaux_out;       /* bf16 [M,N] or NULL */
  void* out; int32_t ldo;                 /* bf16 [M,N]; fp32 per f32_mode */
  int32_t act;
  int32_t act_ncols;                      /* act and aux_out apply to columns n < act_ncols (0 = all) */
  int32_t f32_mode;
  int32_t f32_trans;
  int32_t n_store;                        /* fp32 modes: columns written (0 = N) */
  int32_t splits;                         /* split-K factor, fp32 modes only (0/1 = none) */
  int32_t config;                         /* B200_GEMM_* */
} b200_gemm_desc;

int b200_gemm_bf16(b200_ctx* ctx, const b200_gemm_desc* d, void* stream);

/*
 * Both rank-side weight gradients of ONE adapter in one launch (they contract over the same tokens):
 *   dB[out, r] += alpha * dY[tokens, out]^T . Zc[tokens, :r]        dA[r, in] += alpha * T[tokens, :r]^T . X[tokens, in]
 * fp32 atomics into the flat gradient buffer, split-K over `splits` CTAs per 128-row tile; dW [out, in] is never formed.
 * Zc / T have `zcols` (<= 64, multiple of 8) valid columns starting at the given pointer (column slices of a fused group).
 * Replaces autograd's grads of lora_up / lora_down (toolkit/network_mixins.py:304-342 backward).
 */
int b200_lora_wgrad(b200_ctx* ctx, const void* dY, int lddy, const void* Zc, int ldz, const void* X, int ldx, const void* T,
                    int ldt, void* dB, void* dA, int tokens, int out_dim, int in_dim, int r, int zcols, float alpha,
                    int splits, void* stream);


/* -------------------------------------------------------------------------------------------------
 * AdaLN-Zero modulation around LayerNorm (no affine):
 *   out[m,:] = bf16( bf16( bf16(LN(x[m,:])) * bf16(1 + scale[s,:]) ) + shift[s,:] ),  s = m / rows_per_sample
 * with the bf16 rounding points of the eager model (diffusers AdaLayerNormZero / the in-tree
 * extensions_built_in/diffusion_models/chroma/src/layers.py:471-560).  shift/scale may be NULL (plain LN).
 * mean/rstd [M] fp32 are saved for the backward.  D % 256 == 0.  Algorithmic bytes: 4 per element.
 */
int b200_ln_modulate_fwd(b200_ctx* ctx, const void* x, int ldx, const void* shift, const void* scale, int ldmod,
                         int rows_per_sample, void* out, int ldo, void* mean, void* rstd, int M, int D, float eps,
                         void* stream);
/* out = dres + dLN/dx(dy * (1 + scale));  dres may be NULL.  Algorithmic bytes: 8 per element (6 without dres). */
int b200_ln_modulate_bwd(b200_ctx* ctx, const void* dy, int lddy, const void* x, int ldx, const void* mean,
                         const void* rstd, const void* scale, int ldmod, int rows_per_sample, const void* dres,
                         int lddres, void* out, int ldo, int M, int D, void* stream);
/*
 * Per-sample column reductions for the gradients of the modulation vector (shift / scale / gate), fp32
 * atomics into [samples, ldsum]:  sum_a += sum_rows a;  sum_ab += sum_rows a * f(b) with
 * f(b) = (b - mean[m]) * rstd[m] when mean/rstd are given, else b;  optional mul_out = bf16(a * g[s,:]).
 */
int b200_col_reduce(b200_ctx* ctx, const void* a, int lda, const void* b, int ldb, const void* mean, const void* rstd,
                    const void* g, int ldg, void* mul_out, int ldmul, void* sum_a, void* sum_ab, int ldsum,
                    int rows_per_sample, int M, int D, void* stream);
/*
 * Per-head RMSNorm of q and k (weights wq/wk bf16 [128]), rotary embedding from fp32 cos/sin tables
 * [Ltot,128] (pairs interleaved), and re-layout of q/k/v rows [B*Lseg, ld] into head-major Q/K/V
 * [B, H, Ltot, 128] at sequence offset seq_off (text tokens first, then image tokens).
 * Reference arithmetic: chroma/src/layers.py:72-91 (QKNorm), chroma/src/math.py:33-51 (apply_rope).
 */
int b200_qk_norm_rope_fwd(b200_ctx* ctx, const void* q, const void* k, const void* v, int ld, const void* wq,
                          const void* wk, const void* cos_t, const void* sin_t, void* Q, void* K, void* V, int B, int Lseg,
                          int seq_off, int Ltot, int H, int head_dim, float eps, void* stream);
int b200_qk_norm_rope_bwd(b200_ctx* ctx, const void* dQ, const void* dK, const void* dV, const void* q, const void* k,
                          int ld, const void* wq, const void* wk, const void* cos_t, const void* sin_t, void* dq, void* dk,
                          void* dv, int ldd, int B, int Lseg, int seq_off, int Ltot, int H, int head_dim, float eps,
                          void* stream);
int b200_silu(b200_ctx* ctx, const void* x, void* y, int64_t n, void* stream);
/* out[b,:] = bf16([cos(t f_i) | sin(t f_i)]), t = bf16(bf16(t_in[b] / div) * mult)  (chroma/src/layers.py:30-53);
 * mult <= 0: t = t_in[b] / div in fp32 (Wan2.1 embeds the trainer's fp32 timestep directly) */
int b200_timestep_embed(b200_ctx* ctx, const void* t_in, void* out, int B, int dim, float max_period, float div,
                        float mult, void* stream);
int b200_add_bf16(b200_ctx* ctx, const void* a, const void* b, const void* c, void* y, int64_t n, void* stream);

/* -------------------------------------------------------------------------------------------------
 * Joint attention of the DiT blocks (head_dim 128, non-causal): Q/K/V head-major [B,H,L,128] bf16.
 * Forward writes O token-major: tokens l < split to o0[(b*split + l)*ld0 + h*128 ..], the rest to
 * o1[(b*(L-split) + l-split)*ld1 + h*128 ..] (text / image streams of a double block; split = 0 for a
 * single-stream block writing into its concat buffer), and lse [B,H,L] fp32 (natural log).
 * Backward: delta [B,H,L] fp32 and dOh [B,H,L,128] bf16 are caller-provided scratch; dQ/dK/dV head-major.
 * Replaces torch SDPA inside diffusers' FLUX attention processor (toolkit/stable_diffusion_model.py:2192-2205
 * call path; in-tree equivalent chroma/src/math.py:13-30) and its autograd backward.
 * Algorithmic FLOPs: forward 4 B H L^2 128, backward 2.5x that.
 */
int b200_attn_fwd(b200_ctx* ctx, const void* Q, const void* K, const void* V, void* o0, int ld0, void* o1, int ld1,
                  void* lse, int B, int H, int L, int split, float scale, void* stream);
int b200_attn_bwd(b200_ctx* ctx, const void* Q, const void* K, const void* V, const void* o0, int ld0, const void* o1,
                  int ld1, const void* do0, int ldd0, const void* do1, int ldd1, const void* lse, void* delta, void* dOh,
                  void* dQ, void* dK, void* dV, int B, int H, int L, int split, float scale, void* stream);

/* Cross attention: Q [B,H,L,128] against K / V [B,H,Lk,128] with Lk != L (Wan2.1's text cross-attention,
 * toolkit/models/wan21/wan_attn.py:70-76: F.scaled_dot_product_attention(query, key, value) with key / value from the
 * 512 text tokens).  Same kernels, outputs, scratch and layouts as above; lse / delta [B,H,L]; dK / dV [B,H,Lk,128]. */
int b200_attn_fwd_x(b200_ctx* ctx, const void* Q, const void* K, const void* V, void* o0, int ld0, void* o1, int ld1,
                    void* lse, int B, int H, int L, int Lk, int split, float scale, void* stream);
int b200_attn_bwd_x(b200_ctx* ctx, const void* Q, const void* K, const void* V, const void* o0, int ld0, const void* o1,
                    int ld1, const void* do0, int ldd0, const void* do1, int ldd1, const void* lse, void* delta, void* dOh,
                    void* dQ, void* dK, void* dV, int B, int H, int L, int Lk, int split, float scale, void* stream);

/* `_xd`: the same with `head_live` = 128, or 64 when only the first 64 of the 128 head-dim columns of Q / K / V are non-zero (heads
 * of <= 64 channels zero-padded to the kernels' 128: SDXL's 64, SD1.5's 40): every MMA then skips the zero half of its
 * contraction / output (S, dP: K-steps 0..3; P V, dV, dK, dQ: N = 64) and the forward does not load it; the padded output columns
 * are written as zeros.  Default forward (6) / backward (2, 3) variants only. */
int b200_attn_fwd_xd(b200_ctx* ctx, const void* Q, const void* K, const void* V, void* o0, int ld0, void* o1, int ld1,
                     void* lse, int B, int H, int L, int Lk, int split, float scale, int head_live, void* stream);
int b200_attn_bwd_xd(b200_ctx* ctx, const void* Q, const void* K, const void* V, const void* o0, int ld0, const void* o1,
                     int ld1, const void* do0, int ldd0, const void* do1, int ldd1, const void* lse, void* delta, void* dOh,
                     void* dQ, void* dK, void* dV, int B, int H, int L, int Lk, int split, float scale, int head_live, void* stream);

/* Attention for head dims the tcgen05 kernels do not cover (head_dim 160, 192 or 256; SD1.5's 1280-channel levels: 8 heads of
 * 160 at 64..1024 tokens), on the CUDA cores (csrc/small_attn.cu).  TOKEN-major operands as the projections write them:
 * q / out / dO / dq [B*L, ld] and k / v / dk / dv [B*Lk, ld], head h in columns [h*head_dim, (h+1)*head_dim); lse / delta [B,H,L]
 * fp32 (natural log).  softmax(scale q k^T) v and its gradients, fp32 statistics; L != Lk allowed (cross attention).
 * Other head dims: B200_ERR_INVALID (<= 128 belongs to b200_attn_fwd_xd).  Replaces F.scaled_dot_product_attention as called by
 * diffusers' AttnProcessor2_0 for the UNet (the reference's own call of it: toolkit/models/wan21/wan_attn.py:70-76). */
int b200_attn_small_fwd(b200_ctx* ctx, const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out, int ldo,
                        void* lse, int B, int H, int L, int Lk, int head_dim, float scale, void* stream);
int b200_attn_small_bwd(b200_ctx* ctx, const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, const void* o,
                        int ldo, const void* dO, int lddo, const void* lse, void* delta, void* dq, int lddq, void* dk, int lddk,
                        void* dv, int lddv, int B, int H, int L, int Lk, int head_dim, float scale, void* stream);

/* Wan2.1 attention pre-processing (toolkit/models/wan21/wan_attn.py:34-61): RMSNorm ACROSS heads (one RMS over the whole
 * inner dimension H*128, diffusers qk_norm="rms_norm_across_heads") + RoPE on interleaved pairs (cos/sin [Ltot,128] fp32, NULL:
 * no rotation = the cross-attention) + re-layout of ONE tensor x [B*Lseg, ld] to head-major out [B,H,Ltot,128].
 * mode 0: plain re-layout (values); mode 1: norm (+ rope), rstd [B*Lseg] fp32 saved for the backward. */
int b200_rms_rope_fwd(b200_ctx* ctx, const void* x, int ld, const void* weight, const void* cos_t, const void* sin_t, void* out,
                      void* rstd, int B, int Lseg, int seq_off, int Ltot, int H, float eps, int mode, void* stream);
int b200_rms_rope_bwd(b200_ctx* ctx, const void* dY, const void* x, int ld, const void* weight, const void* cos_t,
                      const void* sin_t, const void* rstd, void* dx, int ldd, int B, int Lseg, int seq_off, int Ltot, int H,
                      int mode, void* stream);

/* -------------------------------------------------------------------------------------------------
 * LoRA-wrapped Linear for Bm <= 8 rows (the AdaLN modulation projections): weight streaming at the
 * HBM roofline, fp32 master A [r,K] / B [N,r] used directly.
 *   y = bf16( bf16(x W^T + bias) + bf16( (c x A^T) B^T ) );  z = c x A^T [Bm,r] fp32 is saved.
 * r == 0: plain frozen Linear.  Replaces toolkit/network_mixins.py:304-342 for these modules.
 */
int b200_lora_gemv_fwd(b200_ctx* ctx, const void* x, int ldx, const void* W, int ldw, const void* bias, const void* A,
                       const void* Bw, int r, float c, void* y, int ldy, void* z, int Bm, int N, int K, void* stream);
/* dA += (c dy B)^T x,  dB += dy^T z   with dy fp32 [Bm, lddy];  t_ws: fp32 [Bm*r] scratch. */
int b200_lora_gemv_bwd(b200_ctx* ctx, const void* dy, int lddy, const void* x, int ldx, const void* z, const void* A,
                       const void* Bw, int r, float c, void* dA, void* dBw, void* t_ws, int Bm, int N, int K,
                       void* stream);

/* Same with a per-row coefficient: row r uses c * row_c[r] (row_c fp32 [Bm], may be NULL).  The per-sample
 * `network.multiplier` list of the trainer (SDTrainer.py:1558 -> toolkit/network_mixins.py:311-322) on modules whose
 * input is the [B, D] conditioning vector. */
int b200_lora_gemv_fwd_rows(b200_ctx* ctx, const void* x, int ldx, const void* W, int ldw, const void* bias, const void* A,
                            const void* Bw, int r, float c, const void* row_c, void* y, int ldy, void* z, int Bm, int N,
                            int K, void* stream);
int b200_lora_gemv_bwd_rows(b200_ctx* ctx, const void* dy, int lddy, const void* x, int ldx, const void* z, const void* A,
                            const void* Bw, int r, float c, const void* row_c, void* dA, void* dBw, void* t_ws, int Bm,
                            int N, int K, void* stream);

/* -------------------------------------------------------------------------------------------------
 * Flow-matching batch preparation and loss.
 * b200_flow_add_noise: out = bf16((1 - t/1000) x0 + (t/1000) noise), optionally written in FLUX's packed
 *   "b c (h 2) (w 2) -> b (h w) (c 4)" layout (toolkit/samplers/custom_flowmatch_sampler.py:91-102,
 *   toolkit/stable_diffusion_model.py:2166-2172).  t fp32 [B] in [0,1000].  8 bytes per latent element
 *   together with b200_flow_loss's reads.
 * b200_flow_loss: target = bf16(noise - x0); loss_per_sample[b] = mean((pred - target)^2);
 *   loss_total = mean_b; dpred = bf16(2 (pred - target) gscale / (C H W B)) in pred's layout
 *   (extensions_built_in/sd_trainer/SDTrainer.py:644-646, 916, 987-990, 1013).
 */
int b200_flow_add_noise(b200_ctx* ctx, const void* latents, const void* noise, const void* t, void* out, int B, int C,
                        int H, int W, int pack, void* stream);
int b200_flow_loss(b200_ctx* ctx, const void* pred, const void* latents, const void* noise, void* dpred,
                   void* loss_per_sample, void* loss_total, int B, int C, int H, int W, int pack, float gscale,
                   void* stream);

/* -------------------------------------------------------------------------------------------------
 * eps / v-prediction models (SD1.5, SDXL) and the non-default branches of `calculate_loss`.
 * b200_ddpm_add_noise: diffusers DDPMScheduler.add_noise as built by toolkit/sampler.py:31-50,120-185 and called
 *   through toolkit/stable_diffusion_model.py:1854-1876 with INTEGER timesteps: noisy = sqrt(ac[t]) x0 +
 *   sqrt(1 - ac[t]) noise in bf16 tensor arithmetic (table value, both roots and both products rounded to bf16).
 *   timesteps int64 [B]; alphas_cumprod fp32 [n_train]; per_sample = C H W (even).
 * b200_train_loss: the default 'mse' path of SDTrainer.calculate_loss (SDTrainer.py:522-1052):
 *   target = `target` if given (prior prediction :619-621, or anything precomputed)
 *          | bf16(bf16(coef_noise[b] noise) - bf16(coef_latent[b] x0))   coef NULL = 1: flow matching (:644-646);
 *            (1, 0): eps (:650); (sqrt(ac), sqrt(1-ac)): v-prediction (:623-625, DDPMScheduler.get_velocity)
 *   loss_per_sample[b] = sample_weight[b] * mean((pred - target)^2 * mask[b, (c), h, w])      (:916, :923-959, :987-1011;
 *     timestep weights, loss_multiplier and SNR-gamma weights are per-sample scalars: the caller multiplies them into
 *     sample_weight, fp32 [B] or NULL; mask fp32 [B, mask_channels in {1, C}, H, W] or NULL)
 *   loss_total = mean_b (:1013); dpred = bf16(d loss_total / d pred * gscale) in pred's layout (pack bit 0 as
 *   b200_flow_loss; pack bit 1: `target` lies in pred's layout -- a prior prediction of the same model, :1211-1339).
 */
int b200_ddpm_add_noise(b200_ctx* ctx, const void* latents, const void* noise, const void* timesteps_i64,
                        const void* alphas_cumprod_f32, int n_train, void* out, int B, int64_t per_sample, void* stream);
int b200_train_loss(b200_ctx* ctx, const void* pred, const void* latents, const void* noise, const void* target,
                    const void* coef_noise, const void* coef_latent, const void* sample_weight, const void* mask,
                    int mask_channels, void* dpred, void* loss_per_sample, void* loss_total, int B, int C, int H, int W,
                    int pack, float gscale, void* stream);

/* -------------------------------------------------------------------------------------------------
 * Conv2d LoRA (toolkit/lora_special.py:95-104: lora_down = Conv2d(in, r, k, stride, padding), lora_up = 1x1) and the
 * frozen Conv2d it wraps, as the SAME fused tcgen05 GEMM as a Linear over rows [B Ho Wo, C kh kw]:
 * b200_nchw_rows: to_rows = 1: NCHW bf16 [B, C, HW] -> rows [B HW, ld_rows] (first C columns); 0: the inverse.
 * b200_im2col: cols[(b, oy, ox), (c, ky, kx)] = x[b, c, oy sh + ky - ph, ox sw + kx - pw] (0 outside), the column order
 *   of `weight.view(out, -1)`; columns [C kh kw, ld) are zero-filled.
 * b200_col2im: the adjoint (gather form, deterministic): dx (+)= sum of the matching dcols entries.
 * b200_mask_rows: z[row, col] *= row_mask[row, col] * col_mask[sample(row), col]  (either may be NULL): the dropout /
 *   rank-dropout masks of toolkit/network_mixins.py:197-239 on the rank-side activations.
 */
int b200_nchw_rows(b200_ctx* ctx, const void* src, void* dst, int B, int C, int HW, int ld_rows, int to_rows, void* stream);
int b200_im2col(b200_ctx* ctx, const void* x, void* cols, int B, int C, int H, int W, int kh, int kw, int sh, int sw, int ph,
                int pw, int ld, void* stream);
int b200_col2im(b200_ctx* ctx, const void* dcols, void* dx, int B, int C, int H, int W, int kh, int kw, int sh, int sw, int ph,
                int pw, int ld, int accumulate, void* stream);
int b200_mask_rows(b200_ctx* ctx, void* z, int ldz, const void* row_mask, int ldm, const void* col_mask, int rows_per_sample,
                   int64_t rows, int cols, void* stream);

/* -------------------------------------------------------------------------------------------------
 * Row / layout kernels of the UNet `Transformer2DModel` blocks (SD1.5 / SDXL; the reference's default LoRA target,
 * toolkit/kohya_lora.py:750; called through toolkit/stable_diffusion_model.py:2049-2055, :2260-2265).
 * b200_ln_affine_fwd/bwd: LayerNorm with weight / bias for any D % 8 == 0 (640 / 1280 channels); bwd adds `dres`.
 * b200_groupnorm_fwd/bwd: GroupNorm(G groups, affine, optional SiLU) over NCHW bf16; mean / rstd fp32 [B*G].
 * b200_geglu_fwd/bwd: proj [M, 2F] = (hidden | gate) -> hidden * gelu_erf(gate) [M, F] (diffusers GEGLU) and its gradient.
 * b200_heads_pad: [B*L, ld] with H heads of head_dim <= 128 columns <-> head-major [B,H,L,128] zero-padded (to_heads 1 / 0),
 *   so that the head-dim-128 attention kernels serve heads of 40 / 64 / 80 (pass scale = 1/sqrt(head_dim) to them).
 */
int b200_ln_affine_fwd(b200_ctx* ctx, const void* x, int ldx, const void* weight, const void* bias, void* out, int ldo,
                       void* mean, void* rstd, int M, int D, float eps, void* stream);
int b200_ln_affine_bwd(b200_ctx* ctx, const void* dy, int lddy, const void* x, int ldx, const void* mean, const void* rstd,
                       const void* weight, const void* dres, int lddres, void* out, int ldo, int M, int D, void* stream);
int b200_groupnorm_fwd(b200_ctx* ctx, const void* x, const void* weight, const void* bias, void* out, void* mean, void* rstd,
                       int B, int C, int HW, int G, float eps, int silu, void* stream);
int b200_groupnorm_bwd(b200_ctx* ctx, const void* dy, const void* x, const void* weight, const void* bias, const void* mean,
                       const void* rstd, void* dx, int B, int C, int HW, int G, int silu, void* stream);
int b200_geglu_fwd(b200_ctx* ctx, const void* proj, int ldp, void* out, int ldo, int64_t M, int F, void* stream);
int b200_geglu_bwd(b200_ctx* ctx, const void* dy, int lddy, const void* proj, int ldp, void* dproj, int lddp, int64_t M, int F,
                   void* stream);
int b200_heads_pad(b200_ctx* ctx, const void* src, void* dst, int ld, int B, int L, int H, int head_dim, int to_heads,
                   void* stream);
/* the same for up to three tensors of one attention (q / k / v, or dQ / dK / dV) in ONE launch; tensor i has L_i tokens per sample
 * (cross attention: q has L, k / v have Lk) */
int b200_heads_pad3(b200_ctx* ctx, const void* src0, void* dst0, int ld0, int L0, const void* src1, void* dst1, int ld1, int L1,
                    const void* src2, void* dst2, int ld2, int L2, int n, int B, int H, int head_dim, int to_heads, void* stream);

/* -------------------------------------------------------------------------------------------------
 * Optimizer over the flat fp32 LoRA parameter buffer.
 * b200_grad_sumsq: *sumsq_f64 = sum g^2.
 * b200_clip_adamw: total_norm = sqrt(sumsq) * hyper[7]; g *= min(1, max_norm / (total_norm + 1e-6))
 *   (accelerator.clip_grad_norm_, SDTrainer.py:2278-2283); torch.optim.AdamW update with decoupled weight
 *   decay (toolkit/optimizer.py:78-79); optional EMA shadow -= (1 - decay_t)(shadow - p) (toolkit/ema.py:100-152)
 *   with decay_t = decay, as the trainer builds it (BaseSDTrainProcess.py:798-803), or
 *   decay_t = min(decay, (1 + n) / (10 + n)) when state int64[1] != 0 (`use_num_updates=True`).
 *   hyper fp32[8] (device): lr, beta1, beta2, eps, weight_decay, max_norm, ema_decay, grad_prescale.
 *   state (device, 64 bytes, zero-initialised): int64[0] step counter, int64[1] EMA warm-up flag, then per-step
 *   derived scalars.
 *   28 bytes per parameter (36 with EMA).
 * b200_repack_lora: bf16 padded operand copies of the fp32 masters; table = n_entries x
 *   {int64 src_off, int64 dst_off, int32 rows, cols, dst_ld, pad}.
 */
int b200_grad_sumsq(b200_ctx* ctx, const void* g, int64_t n, void* sumsq_f64, void* stream);
int b200_clip_adamw(b200_ctx* ctx, void* p, void* g, void* m, void* v, void* ema, const void* sumsq_f64,
                    const void* hyper, void* state, int64_t n, void* norm_out, void* stream);
int b200_repack_lora(b200_ctx* ctx, const void* flat_f32, void* pack_bf16, const void* table, int n_entries,
                     void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200_LORA_H_ */
