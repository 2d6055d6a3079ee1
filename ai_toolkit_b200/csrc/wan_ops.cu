// Wan2.1 attention pre-processing: RMSNorm ACROSS heads (diffusers qk_norm="rms_norm_across_heads": one RMS over the whole
// inner dimension H * 128) + 3-D RoPE + re-layout to head-major [B, H, Ltot, 128], one tensor per launch because the
// cross-attention takes its queries and its keys / values from different token streams.
// Reference arithmetic: toolkit/models/wan21/wan_attn.py:34-61 (in-tree attention processor): `attn.norm_q(query)` BEFORE
// the head split, `apply_rotary_emb` on interleaved pairs (complex multiply in float64, one rounding to the tensor dtype);
// diffusers' RMSNorm rounds x * rsqrt(mean(x^2) + eps) to the weight dtype (bf16) and multiplies by the weight.
//   mode 0: plain re-layout (values)      mode 1: RMSNorm (+ RoPE when cos / sin are given)
// One warp per token row; lane owns elements 4 lane .. 4 lane + 3 of every head (two rotary pairs).
#include "common.cuh"
#include "ctx.h"

namespace b200 {

__device__ __forceinline__ void wld4(const bf16* p, float (&v)[4]) {
  uint2 u = *reinterpret_cast<const uint2*>(p);
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
__device__ __forceinline__ void wst4(bf16* p, const float (&v)[4]) {
  uint2 u;
  u.x = pack_bf16x2(v[0], v[1]);
  u.y = pack_bf16x2(v[2], v[3]);
  *reinterpret_cast<uint2*>(p) = u;
}

__global__ void __launch_bounds__(256) rms_rope_fwd_kernel(const bf16* __restrict__ x, int ld, const bf16* __restrict__ w,
                                                           const float* __restrict__ cos_t, const float* __restrict__ sin_t,
                                                           bf16* __restrict__ out, float* __restrict__ rstd_out, int B, int Lseg,
                                                           int seq_off, int Ltot, int H, float eps, int mode) {
  pdl_grid_sync();
  const long long tok = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (tok >= static_cast<long long>(B) * Lseg) return;
  const int l = static_cast<int>(tok % Lseg), b = static_cast<int>(tok / Lseg);
  const int pos = seq_off + l;
  const bf16* xr = x + static_cast<size_t>(tok) * ld + lane * 4;
  float rstd = 1.0f;
  if (mode == 1) {
    float ss = 0.f;
    for (int h = 0; h < H; ++h) {
      float v[4];
      wld4(xr + h * 128, v);
      ss += v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3];
    }
    rstd = rsqrtf(warp_sum(ss) / static_cast<float>(H * 128) + eps);
    if (lane == 0 && rstd_out) rstd_out[tok] = rstd;
  }
  float c[4] = {1.f, 1.f, 1.f, 1.f}, s[4] = {0.f, 0.f, 0.f, 0.f};
  if (mode == 1 && cos_t) {
    *reinterpret_cast<float4*>(c) = *reinterpret_cast<const float4*>(cos_t + static_cast<size_t>(pos) * 128 + lane * 4);
    *reinterpret_cast<float4*>(s) = *reinterpret_cast<const float4*>(sin_t + static_cast<size_t>(pos) * 128 + lane * 4);
  }
  for (int h = 0; h < H; ++h) {
    const size_t dst = ((static_cast<size_t>(b) * H + h) * Ltot + pos) * 128 + lane * 4;
    if (mode == 0) {
      *reinterpret_cast<uint2*>(out + dst) = *reinterpret_cast<const uint2*>(xr + h * 128);
      continue;
    }
    float v[4], wv[4], y[4], o[4];
    wld4(xr + h * 128, v);
    wld4(w + h * 128 + lane * 4, wv);
#pragma unroll
    for (int i = 0; i < 4; ++i) y[i] = bf16_round(bf16_round(v[i] * rstd) * wv[i]);
    o[0] = y[0] * c[0] - y[1] * s[0];
    o[1] = y[1] * c[1] + y[0] * s[1];
    o[2] = y[2] * c[2] - y[3] * s[2];
    o[3] = y[3] * c[3] + y[2] * s[3];
    wst4(out + dst, o);
  }
}

// dx = rstd (w g) - x rstd^3 mean_row(x w g),  g = RoPE^T dY  (mode 1);  dx = dY re-laid token-major (mode 0)
__global__ void __launch_bounds__(256) rms_rope_bwd_kernel(const bf16* __restrict__ dY, const bf16* __restrict__ x, int ld,
                                                           const bf16* __restrict__ w, const float* __restrict__ cos_t,
                                                           const float* __restrict__ sin_t, const float* __restrict__ rstd_in,
                                                           bf16* __restrict__ dx, int ldd, int B, int Lseg, int seq_off, int Ltot,
                                                           int H, int mode) {
  pdl_grid_sync();
  const long long tok = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (tok >= static_cast<long long>(B) * Lseg) return;
  const int l = static_cast<int>(tok % Lseg), b = static_cast<int>(tok / Lseg);
  const int pos = seq_off + l;
  bf16* dr = dx + static_cast<size_t>(tok) * ldd + lane * 4;
  if (mode == 0) {
    for (int h = 0; h < H; ++h) {
      const size_t hm = ((static_cast<size_t>(b) * H + h) * Ltot + pos) * 128 + lane * 4;
      *reinterpret_cast<uint2*>(dr + h * 128) = *reinterpret_cast<const uint2*>(dY + hm);
    }
    return;
  }
  const bf16* xr = x + static_cast<size_t>(tok) * ld + lane * 4;
  const float rstd = rstd_in[tok];
  float c[4] = {1.f, 1.f, 1.f, 1.f}, s[4] = {0.f, 0.f, 0.f, 0.f};
  if (cos_t) {
    *reinterpret_cast<float4*>(c) = *reinterpret_cast<const float4*>(cos_t + static_cast<size_t>(pos) * 128 + lane * 4);
    *reinterpret_cast<float4*>(s) = *reinterpret_cast<const float4*>(sin_t + static_cast<size_t>(pos) * 128 + lane * 4);
  }
  auto grad4 = [&](int h, float (&gw)[4], float (&xv)[4]) {
    const size_t hm = ((static_cast<size_t>(b) * H + h) * Ltot + pos) * 128 + lane * 4;
    float g[4], wv[4];
    wld4(dY + hm, g);
    wld4(w + h * 128 + lane * 4, wv);
    wld4(xr + h * 128, xv);
    gw[0] = (g[0] * c[0] + g[1] * s[1]) * wv[0];
    gw[1] = (g[1] * c[1] - g[0] * s[0]) * wv[1];
    gw[2] = (g[2] * c[2] + g[3] * s[3]) * wv[2];
    gw[3] = (g[3] * c[3] - g[2] * s[2]) * wv[3];
  };
  float dot = 0.f;
  for (int h = 0; h < H; ++h) {
    float gw[4], xv[4];
    grad4(h, gw, xv);
    dot += gw[0] * xv[0] + gw[1] * xv[1] + gw[2] * xv[2] + gw[3] * xv[3];
  }
  dot = warp_sum(dot) / static_cast<float>(H * 128);
  const float r3 = rstd * rstd * rstd * dot;
  for (int h = 0; h < H; ++h) {
    float gw[4], xv[4], o[4];
    grad4(h, gw, xv);
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = rstd * gw[i] - xv[i] * r3;
    wst4(dr + h * 128, o);
  }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_rms_rope_fwd(b200_ctx* ctx, const void* x, int ld, const void* weight, const void* cos_t, const void* sin_t,
                                 void* out, void* rstd, int B, int Lseg, int seq_off, int Ltot, int H, float eps, int mode,
                                 void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  B200_REQUIRE(x && out && B > 0 && Lseg > 0 && H > 0 && ld % 4 == 0 && (mode == 0 || mode == 1), "b200_rms_rope_fwd: bad args");
  B200_REQUIRE(mode == 0 || (weight && rstd), "b200_rms_rope_fwd: mode 1 needs the norm weight and an rstd buffer");
  B200_REQUIRE((cos_t == nullptr) == (sin_t == nullptr), "b200_rms_rope_fwd: cos and sin go together");
  B200_REQUIRE(seq_off >= 0 && seq_off + Lseg <= Ltot, "b200_rms_rope_fwd: segment out of range");
  const long long rows = static_cast<long long>(B) * Lseg;
  B200_KLAUNCH(rms_rope_fwd_kernel, static_cast<unsigned>((rows + 7) / 8), 256, 0, reinterpret_cast<cudaStream_t>(stream),
               (const bf16*)x, ld, (const bf16*)weight, (const float*)cos_t, (const float*)sin_t, (bf16*)out, (float*)rstd, B,
               Lseg, seq_off, Ltot, H, eps, mode);
  B200_CUDA_CHECK(cudaGetLastError());
  ctx->launches.fetch_add(1);
  return B200_OK;
}

extern "C" int b200_rms_rope_bwd(b200_ctx* ctx, const void* dY, const void* x, int ld, const void* weight, const void* cos_t,
                                 const void* sin_t, const void* rstd, void* dx, int ldd, int B, int Lseg, int seq_off, int Ltot,
                                 int H, int mode, void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  B200_REQUIRE(dY && dx && B > 0 && Lseg > 0 && H > 0 && ldd % 4 == 0 && (mode == 0 || mode == 1), "b200_rms_rope_bwd: bad args");
  B200_REQUIRE(mode == 0 || (x && weight && rstd && ld % 4 == 0), "b200_rms_rope_bwd: mode 1 needs x, the norm weight and rstd");
  B200_REQUIRE((cos_t == nullptr) == (sin_t == nullptr), "b200_rms_rope_bwd: cos and sin go together");
  const long long rows = static_cast<long long>(B) * Lseg;
  B200_KLAUNCH(rms_rope_bwd_kernel, static_cast<unsigned>((rows + 7) / 8), 256, 0, reinterpret_cast<cudaStream_t>(stream),
               (const bf16*)dY, (const bf16*)x, ld, (const bf16*)weight, (const float*)cos_t, (const float*)sin_t,
               (const float*)rstd, (bf16*)dx, ldd, B, Lseg, seq_off, Ltot, H, mode);
  B200_CUDA_CHECK(cudaGetLastError());
  ctx->launches.fetch_add(1);
  return B200_OK;
}
