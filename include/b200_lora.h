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
 *   acc[M,N] = A0[M,K0] . B0[N,K0]^T  +  A1[M,K1] . B1[N,K1]^T          (fp32 accumulate in TMEM)
 *   y   = bf16(acc + bias[n])
 *   aux_out[m,n] = y                         (optional: pre-activation kept for backward)
 *   y   = bf16(gelu_tanh(y))                 (act == B200_ACT_GELU_TANH)
 *   y   = bf16(y * gelu_tanh'(aux_in[m,n]))  (optional: backward through the activation)
 *   y   = bf16(y * gate[m / rows_per_sample, n])   (optional AdaLN-Zero gate)
 *   out[m,n] = bf16(res[m,n] + y)            (optional residual; res may alias out)
 *
 * The second segment is how the LoRA up-projection rides in the same TMEM tile as the frozen base
 * GEMM:  A1 = bf16(m_b * s * (x A^T)) padded to 64 columns, B1 = lora_up weight padded to 64 columns.
 * Replaces, per wrapped Linear, `org_forward(x) + (lora_up(lora_down(x.float())) * scale * multiplier).to(bf16)`
 * (toolkit/network_mixins.py:304-342) and the dX half of its autograd backward.
 *
 * out_f32 != 0: write the raw fp32 accumulator instead (no epilogue), `splits` partial results of a
 * split-K contraction go to out + split*M*ldo (used by the rank-side GEMMs).
 */
typedef struct b200_gemm_desc {
  int32_t M, N, K0, K1;
  const void* A0; int32_t lda0;
  const void* B0; int32_t ldb0;
  const void* A1; int32_t lda1;
  const void* B1; int32_t ldb1;
  const void* bias;                       /* bf16 [N] or NULL */
  const void* res; int32_t ldres;         /* bf16 [M,N] or NULL */
  const void* gate; int32_t ldgate;       /* bf16 [ceil(M/rows_per_sample), N] or NULL */
  int32_t rows_per_sample;                /* >0 when gate is given */
  const void* aux_in; int32_t ldaux_in;   /* bf16 [M,N] or NULL */
  void* aux_out; int32_t ldaux_out;       /* bf16 [M,N] or NULL */
  void* out; int32_t ldo;                 /* bf16 [M,N] (or fp32 [splits,M,N] when out_f32) */
  int32_t act;
  int32_t out_f32;
  int32_t splits;                         /* split-K factor, only with out_f32 (0/1 = none) */
  int32_t config;                         /* B200_GEMM_* */
} b200_gemm_desc;

int b200_gemm_bf16(b200_ctx* ctx, const b200_gemm_desc* d, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200_LORA_H_ */
