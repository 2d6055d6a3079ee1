// Device helpers and argument blocks shared by the attention kernels (attention.cu = the validated round-1 kernels,
// attention_r2.cu = the opt-in round-2 candidates selected with B200_ATTN_FWD / B200_ATTN_BWD).
#pragma once
#include "common.cuh"
#include "ctx.h"

namespace b200 {

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 2^x on the FMA pipe (Cody-Waite split + degree-3 minimax polynomial, rel. error ~1e-4, plenty for a bf16 P):
// the softmax is MUFU-bound (16 ex2 / clk / SM = the same 1024 cycles per 128 x 128 tile as the two MMAs), so a share
// of the exponentials is computed here instead.  Valid for x in [-126, 127].
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -126.0f);
  const float xr = __fadd_rd(x, 12582912.0f);       // 1.5 * 2^23: the integer part lands in the low mantissa bits
  const float f = x - (xr - 12582912.0f);           // fractional part in [0, 1)
  float p = fmaf(f, 0.077119089663028717f, 0.227564394474029541f);
  p = fmaf(p, f, 0.695146143436431885f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(xr) << 23));
}
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
// 16-byte store of 8 bf16 into a [rows x 128 B] SWIZZLE_128B tile (row r, 16-byte chunk c of 8)
__device__ __forceinline__ void st_sw128(uint8_t* tile, int r, int c, uint4 v) {
  *reinterpret_cast<uint4*>(tile + r * 128 + ((c ^ (r & 7)) << 4)) = v;
}

#ifndef B200_ATTN_POLY
#define B200_ATTN_POLY 1
#endif
constexpr int kPolyPairsOf4 = B200_ATTN_POLY;  // of every 4 column pairs, how many use ex2_poly (0 = all MUFU)
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// Packed fp32 pairs (sm_100 FFMA2 / FADD2: two fp32 lanes per instruction = half the issue slots)
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ float2 fadd2_rm(float2 a, float2 b) {  // both lanes rounded towards -inf
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "add.rm.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "mul.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "add.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}


// ---- batched issue: ONE elect.sync for a whole K loop -------------------------------------------------------------------
// The per-instruction path above costs ~13 SASS instructions (ELECT, VOTEU, 5 x R2UR ...) per tcgen05.mma; for the M 128 x
// N 64 MMAs of the backward (32 cycles on the tensor pipe) that is as long as the MMA itself, and the diagnostics show the
// backward bound by its MMA / barrier pipeline (750 of 795 us without any softmax work).  Here the elected lane is chosen once
// and the n descriptors are base + compile-time offsets computed inside the asm block.
template <int OA1, int OA2, int OA3, int OA4, int OA5, int OA6, int OA7, int OB1, int OB2, int OB3, int OB4, int OB5, int OB6, int OB7>
__device__ __forceinline__ void umma_bf16_ss_w_x8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                   uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred e, p, t;\n\t.reg .b64 da, db;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.eq.b32 t, %3, %3;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "add.s64 da, %1, %5;\n\tadd.s64 db, %2, %12;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, t;\n\t"
      "add.s64 da, %1, %6;\n\tadd.s64 db, %2, %13;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, t;\n\t"
      "add.s64 da, %1, %7;\n\tadd.s64 db, %2, %14;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, t;\n\t"
      "add.s64 da, %1, %8;\n\tadd.s64 db, %2, %15;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, t;\n\t"
      "add.s64 da, %1, %9;\n\tadd.s64 db, %2, %16;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, t;\n\t"
      "add.s64 da, %1, %10;\n\tadd.s64 db, %2, %17;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, t;\n\t"
      "add.s64 da, %1, %11;\n\tadd.s64 db, %2, %18;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, t;\n\t"
      "}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "n"(OA1), "n"(OA2), "n"(OA3), "n"(OA4), "n"(OA5), "n"(OA6), "n"(OA7), "n"(OB1), "n"(OB2), "n"(OB3), "n"(OB4), "n"(OB5), "n"(OB6), "n"(OB7)
      : "memory");
}
template <int OA1, int OA2, int OA3, int OA4, int OA5, int OA6, int OA7, int OB1, int OB2, int OB3, int OB4, int OB5, int OB6, int OB7>
__device__ __forceinline__ void umma_bf16_ts_w_x8(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                                   uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred e, p, t;\n\t.reg .b32 ta;\n\t.reg .b64 db;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.eq.b32 t, %3, %3;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "add.u32 ta, %1, %5;\n\tadd.s64 db, %2, %12;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, %3, t;\n\t"
      "add.u32 ta, %1, %6;\n\tadd.s64 db, %2, %13;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, %3, t;\n\t"
      "add.u32 ta, %1, %7;\n\tadd.s64 db, %2, %14;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, %3, t;\n\t"
      "add.u32 ta, %1, %8;\n\tadd.s64 db, %2, %15;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, %3, t;\n\t"
      "add.u32 ta, %1, %9;\n\tadd.s64 db, %2, %16;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, %3, t;\n\t"
      "add.u32 ta, %1, %10;\n\tadd.s64 db, %2, %17;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, %3, t;\n\t"
      "add.u32 ta, %1, %11;\n\tadd.s64 db, %2, %18;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, %3, t;\n\t"
      "}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "n"(OA1), "n"(OA2), "n"(OA3), "n"(OA4), "n"(OA5), "n"(OA6), "n"(OA7), "n"(OB1), "n"(OB2), "n"(OB3), "n"(OB4), "n"(OB5), "n"(OB6), "n"(OB7)
      : "memory");
}
template <int OA1, int OA2, int OA3, int OB1, int OB2, int OB3>
__device__ __forceinline__ void umma_bf16_ss_w_x4(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                   uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred e, p, t;\n\t.reg .b64 da, db;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.eq.b32 t, %3, %3;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "add.s64 da, %1, %5;\n\tadd.s64 db, %2, %8;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, t;\n\t"
      "add.s64 da, %1, %6;\n\tadd.s64 db, %2, %9;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, t;\n\t"
      "add.s64 da, %1, %7;\n\tadd.s64 db, %2, %10;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, t;\n\t"
      "}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "n"(OA1), "n"(OA2), "n"(OA3), "n"(OB1), "n"(OB2), "n"(OB3)
      : "memory");
}
template <int OA1, int OA2, int OA3, int OB1, int OB2, int OB3>
__device__ __forceinline__ void umma_bf16_ts_w_x4(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                                   uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred e, p, t;\n\t.reg .b32 ta;\n\t.reg .b64 db;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.eq.b32 t, %3, %3;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "add.u32 ta, %1, %5;\n\tadd.s64 db, %2, %8;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, %3, t;\n\t"
      "add.u32 ta, %1, %6;\n\tadd.s64 db, %2, %9;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, %3, t;\n\t"
      "add.u32 ta, %1, %7;\n\tadd.s64 db, %2, %10;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, %3, t;\n\t"
      "}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "n"(OA1), "n"(OA2), "n"(OA3), "n"(OB1), "n"(OB2), "n"(OB3)
      : "memory");
}

struct AttnFwdArgs {
  bf16* o0;
  int ld0;
  bf16* o1;
  int ld1;
  float* lse;
  int B, H, L, split;
  float scale;
  int Lk;  // key / value length (== L for self / joint attention; != L for cross attention: fwd variants 3, 4, 5 only)
  int dlive;  // 128, or 64: only the first 64 head-dim columns of Q / K / V are non-zero (heads of <= 64 channels zero-padded to
              // 128, SDXL / SD1.5): the MMAs skip the zero half (variant 6 forward, variants 2 / 3 backward)
};


struct AttnBwdArgs {
  const float* lse;
  const float* delta;
  const bf16* r0;  // MODE_Q: Q  [B,H,L,128] (the stationary operands are read straight into TMEM)
  const bf16* r1;  // MODE_Q: dO [B,H,L,128]
  bf16* out0;  // MODE_KV: dV ; MODE_Q: dQ
  bf16* out1;  // MODE_KV: dK
  int L;   // length of the STATIONARY side (MODE_KV: keys / values; MODE_Q: queries)
  float scale;
  int Lt;  // length of the STREAMED side (MODE_KV: queries; MODE_Q: keys / values); == L for self / joint attention
  int dbg; // timing experiments only (B200_ATTN_BWD_DBG=2): skip the softmax work, hand the barriers on (results are wrong)
  int dlive;  // 128 or 64 (see AttnFwdArgs)
};


// ex2_poly for a pair (same split and polynomial, packed instructions: 10 issue slots per pair instead of 16)
__device__ __forceinline__ float2 ex2_poly2(float2 x) {
  const float2 magic = make_float2(12582912.0f, 12582912.0f);
  x.x = fmaxf(x.x, -126.0f);
  x.y = fmaxf(x.y, -126.0f);
  const float2 xr = fadd2_rm(x, magic);
  const float2 f = fadd2(x, fadd2(magic, make_float2(-xr.x, -xr.y)));  // x - (xr - magic), both roundings exact
  float2 p = ffma2(f, make_float2(0.077119089663028717f, 0.077119089663028717f),
                   make_float2(0.227564394474029541f, 0.227564394474029541f));
  p = ffma2(p, f, make_float2(0.695146143436431885f, 0.695146143436431885f));
  p = ffma2(p, f, make_float2(1.0f, 1.0f));
  return make_float2(__int_as_float(__float_as_int(p.x) + (__float_as_int(xr.x) << 23)),
                     __int_as_float(__float_as_int(p.y) + (__float_as_int(xr.y) << 23)));
}

// opt-in round-2 candidates (attention_r2.cu); variant numbers are the values of B200_ATTN_FWD / B200_ATTN_BWD
int attn_fwd_r2_launch(int variant, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnFwdArgs& a,
                       cudaStream_t stream);
// variant 5: two Q tiles per CTA in ping-pong (attention_pp.cu)
int attn_fwd_pp_launch(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnFwdArgs& a,
                       cudaStream_t stream);
int attn_fwd_pp3_launch(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnFwdArgs& a,
                        cudaStream_t stream);
int attn_fwd_pp2_launch(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnFwdArgs& a,
                        cudaStream_t stream);
int attn_bwd_r2_launch(int variant, const CUtensorMap& k128, const CUtensorMap& v128, const CUtensorMap& q64,
                       const CUtensorMap& d64, const CUtensorMap& q128, const CUtensorMap& d128, const CUtensorMap& k64,
                       const CUtensorMap& v64, const AttnBwdArgs& akv, const AttnBwdArgs& aq, int B, int H,
                       cudaStream_t stream);

}  // namespace b200
