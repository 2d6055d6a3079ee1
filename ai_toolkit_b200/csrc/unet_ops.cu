// Row / layout kernels of the UNet `Transformer2DModel` blocks (SD1.5 / SDXL: the adapter-bearing part of the UNet, the
// reference's default LoRA target `Transformer2DModel`, toolkit/kohya_lora.py:750).  All HBM-bound, fp32 math, one rounding to
// bf16 at the points where the eager bf16 model rounds (diffusers BasicTransformerBlock / GEGLU / Transformer2DModel, called
// through toolkit/stable_diffusion_model.py:2049-2055 and :2260-2265):
//   * LayerNorm with affine weight / bias for ANY width D % 8 == 0 (640 / 1280 channels are not multiples of 256, which the
//     AdaLN kernels of the DiT engines assume), forward and backward (+ residual-stream gradient)
//   * GroupNorm (32 groups, affine, optional SiLU) on NCHW, forward and backward
//   * GEGLU: hidden * gelu_erf(gate) over the two halves of the ff.net.0.proj output, forward and backward
//   * head re-layout with zero padding: [B L, H d] (d = 40 / 64 / 80 <= 128) <-> head-major [B, H, L, 128], so that the
//     head-dim-128 tcgen05 attention kernels serve the UNet heads (the padded columns contribute 0 to Q K^T and receive 0)
#include "common.cuh"
#include "ctx.h"

namespace b200 {

__device__ __forceinline__ void uld8(const bf16* p, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
__device__ __forceinline__ void ust8(bf16* p, const float (&v)[8]) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]);
  u.y = pack_bf16x2(v[2], v[3]);
  u.z = pack_bf16x2(v[4], v[5]);
  u.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}

// y = bf16( (x - mean) rstd w + b ), one warp per row, lanes stride over 8-element chunks
__global__ void __launch_bounds__(256) ln_affine_fwd_kernel(const bf16* __restrict__ x, int ldx, const bf16* __restrict__ w,
                                                            const bf16* __restrict__ b, bf16* __restrict__ out, int ldo,
                                                            float* __restrict__ mean_out, float* __restrict__ rstd_out, int M,
                                                            int D, float eps) {
  pdl_grid_sync();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const bf16* xr = x + static_cast<size_t>(row) * ldx;
  float s = 0.f;
  for (int c = lane * 8; c < D; c += 256) {
    float v[8];
    uld8(xr + c, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i];
  }
  const float mean = warp_sum(s) / static_cast<float>(D);
  float q = 0.f;
  for (int c = lane * 8; c < D; c += 256) {
    float v[8];
    uld8(xr + c, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) q += (v[i] - mean) * (v[i] - mean);
  }
  const float rstd = rsqrtf(warp_sum(q) / static_cast<float>(D) + eps);
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
  bf16* orow = out + static_cast<size_t>(row) * ldo;
  for (int c = lane * 8; c < D; c += 256) {
    float v[8], wv[8], bv[8], o[8];
    uld8(xr + c, v);
    uld8(w + c, wv);
    uld8(b + c, bv);
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = (v[i] - mean) * rstd * wv[i] + bv[i];
    ust8(orow + c, o);
  }
}

// dx = rstd (g - mean(g) - xhat mean(g xhat)),  g = dy w;  out = dres + dx
__global__ void __launch_bounds__(256) ln_affine_bwd_kernel(const bf16* __restrict__ dy, int lddy, const bf16* __restrict__ x,
                                                            int ldx, const float* __restrict__ mean_in,
                                                            const float* __restrict__ rstd_in, const bf16* __restrict__ w,
                                                            const bf16* __restrict__ dres, int lddres, bf16* __restrict__ out,
                                                            int ldo, int M, int D) {
  pdl_grid_sync();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const bf16* xr = x + static_cast<size_t>(row) * ldx;
  const bf16* gr = dy + static_cast<size_t>(row) * lddy;
  const float mean = mean_in[row], rstd = rstd_in[row];
  float s1 = 0.f, s2 = 0.f;
  for (int c = lane * 8; c < D; c += 256) {
    float v[8], g[8], wv[8];
    uld8(xr + c, v);
    uld8(gr + c, g);
    uld8(w + c, wv);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float gw = g[i] * wv[i];
      s1 += gw;
      s2 += gw * (v[i] - mean) * rstd;
    }
  }
  s1 = warp_sum(s1) / static_cast<float>(D);
  s2 = warp_sum(s2) / static_cast<float>(D);
  bf16* orow = out + static_cast<size_t>(row) * ldo;
  for (int c = lane * 8; c < D; c += 256) {
    float v[8], g[8], wv[8], r[8], o[8];
    uld8(xr + c, v);
    uld8(gr + c, g);
    uld8(w + c, wv);
    if (dres) uld8(dres + static_cast<size_t>(row) * lddres + c, r);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float xh = (v[i] - mean) * rstd;
      o[i] = rstd * (g[i] * wv[i] - s1 - xh * s2) + (dres ? r[i] : 0.f);
    }
    ust8(orow + c, o);
  }
}

// GroupNorm over NCHW: one block per (sample, group); the group's cpg channels x HW elements are contiguous
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
  if (threadIdx.x < 32) {
    t = warp_sum(t);
    if (threadIdx.x == 0) red[0] = t;
  }
  __syncthreads();
  return red[0];
}

__global__ void __launch_bounds__(512) groupnorm_fwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w,
                                                            const bf16* __restrict__ b, bf16* __restrict__ out,
                                                            float* __restrict__ mean_out, float* __restrict__ rstd_out, int C,
                                                            int HW, int G, float eps, int silu) {
  pdl_grid_sync();
  __shared__ float red[32];
  const int bg = blockIdx.x;  // b * G + g
  const int g = bg % G;
  const int cpg = C / G;
  const long long n = static_cast<long long>(cpg) * HW;
  const bf16* xp = x + (static_cast<long long>(bg / G) * C + static_cast<long long>(g) * cpg) * HW;
  bf16* op = out + (static_cast<long long>(bg / G) * C + static_cast<long long>(g) * cpg) * HW;
  float s = 0.f;
  for (long long i = threadIdx.x * 2; i < n; i += blockDim.x * 2) {
    const float2 v = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(xp + i));
    s += v.x + v.y;
  }
  const float mean = block_sum(s, red) / static_cast<float>(n);
  float q = 0.f;
  for (long long i = threadIdx.x * 2; i < n; i += blockDim.x * 2) {
    const float2 v = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(xp + i));
    q += (v.x - mean) * (v.x - mean) + (v.y - mean) * (v.y - mean);
  }
  const float rstd = rsqrtf(block_sum(q, red) / static_cast<float>(n) + eps);
  if (threadIdx.x == 0) {
    mean_out[bg] = mean;
    rstd_out[bg] = rstd;
  }
  for (long long i = threadIdx.x * 2; i < n; i += blockDim.x * 2) {  // HW is even: a pair never straddles two channels
    const int c = g * cpg + static_cast<int>(i / HW);
    const float wv = __bfloat162float(w[c]), bv = __bfloat162float(b[c]);
    const float2 v = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(xp + i));
    float y0 = (v.x - mean) * rstd * wv + bv, y1 = (v.y - mean) * rstd * wv + bv;
    if (silu) {  // the eager model rounds the normalised value to bf16 before SiLU
      y0 = silu_f(bf16_round(y0));
      y1 = silu_f(bf16_round(y1));
    }
    *reinterpret_cast<uint32_t*>(op + i) = pack_bf16x2(y0, y1);
  }
}

__global__ void __launch_bounds__(512) groupnorm_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x,
                                                            const bf16* __restrict__ w, const bf16* __restrict__ b,
                                                            const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                                                            bf16* __restrict__ dx, int C, int HW, int G, int silu) {
  pdl_grid_sync();
  __shared__ float red[32];
  const int bg = blockIdx.x;
  const int g = bg % G;
  const int cpg = C / G;
  const long long n = static_cast<long long>(cpg) * HW;
  const long long base = (static_cast<long long>(bg / G) * C + static_cast<long long>(g) * cpg) * HW;
  const float mean = mean_in[bg], rstd = rstd_in[bg];
  auto grad = [&](long long i, float xv, float dyv) -> float {  // d loss / d (normalised, affine) value times w
    const int c = g * cpg + static_cast<int>(i / HW);
    const float wv = __bfloat162float(w[c]);
    float gy = dyv;
    if (silu) {
      const float y = bf16_round((xv - mean) * rstd * wv + __bfloat162float(b[c]));
      const float sg = 1.0f / (1.0f + __expf(-y));
      gy *= sg * (1.0f + y * (1.0f - sg));
    }
    return gy * wv;
  };
  float s1 = 0.f, s2 = 0.f;
  for (long long i = threadIdx.x * 2; i < n; i += blockDim.x * 2) {
    const float2 v = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(x + base + i));
    const float2 d = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(dy + base + i));
    const float g0 = grad(i, v.x, d.x), g1 = grad(i, v.y, d.y);
    s1 += g0 + g1;
    s2 += g0 * (v.x - mean) * rstd + g1 * (v.y - mean) * rstd;
  }
  s1 = block_sum(s1, red) / static_cast<float>(n);
  s2 = block_sum(s2, red) / static_cast<float>(n);
  for (long long i = threadIdx.x * 2; i < n; i += blockDim.x * 2) {
    const float2 v = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(x + base + i));
    const float2 d = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(dy + base + i));
    const float o0 = rstd * (grad(i, v.x, d.x) - s1 - (v.x - mean) * rstd * s2);
    const float o1 = rstd * (grad(i, v.y, d.y) - s1 - (v.y - mean) * rstd * s2);
    *reinterpret_cast<uint32_t*>(dx + base + i) = pack_bf16x2(o0, o1);
  }
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.7071067811865476f)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
  return 0.5f * (1.0f + erff(x * 0.7071067811865476f)) + x * 0.3989422804014327f * __expf(-0.5f * x * x);
}

// proj [M, 2 F] = (hidden | gate) -> out [M, F] = bf16( hidden * bf16(gelu_erf(gate)) )      (diffusers GEGLU)
__global__ void __launch_bounds__(256) geglu_fwd_kernel(const bf16* __restrict__ proj, int ldp, bf16* __restrict__ out, int ldo,
                                                        long long M, int F) {
  pdl_grid_sync();
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int f8 = F / 8;
  if (idx >= M * f8) return;
  const long long r = idx / f8;
  const int c = static_cast<int>(idx % f8) * 8;
  float h[8], gt[8], o[8];
  uld8(proj + r * ldp + c, h);
  uld8(proj + r * ldp + F + c, gt);
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = h[i] * bf16_round(gelu_erf(gt[i]));
  ust8(out + r * ldo + c, o);
}
// dproj = ( dy * gelu(gate) | dy * hidden * gelu'(gate) )
__global__ void __launch_bounds__(256) geglu_bwd_kernel(const bf16* __restrict__ dy, int lddy, const bf16* __restrict__ proj,
                                                        int ldp, bf16* __restrict__ dproj, int lddp, long long M, int F) {
  pdl_grid_sync();
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int f8 = F / 8;
  if (idx >= M * f8) return;
  const long long r = idx / f8;
  const int c = static_cast<int>(idx % f8) * 8;
  float h[8], gt[8], d[8], dh[8], dg[8];
  uld8(proj + r * ldp + c, h);
  uld8(proj + r * ldp + F + c, gt);
  uld8(dy + r * lddy + c, d);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    dh[i] = d[i] * gelu_erf(gt[i]);
    dg[i] = d[i] * h[i] * gelu_erf_grad(gt[i]);
  }
  ust8(dproj + r * lddp + c, dh);
  ust8(dproj + r * lddp + F + c, dg);
}

// x [B L, ld] (H heads of d <= 128 columns) <-> head-major [B, H, L, 128] with zero padding; to_heads = 0: the inverse
// (the padded columns of the head-major tensor are dropped).  One warp per token, lane owns 4 of the 128 columns of a head.
struct HeadsPadSeg {
  const bf16* src;
  bf16* dst;
  int ld, L;
};
struct HeadsPadArgs {
  HeadsPadSeg seg[3];  // blockIdx.y selects the tensor: q / k / v (or dQ / dK / dV) of one attention in ONE launch
  int B, H, d, to_heads;
};

__global__ void __launch_bounds__(256) heads_pad_kernel(const HeadsPadArgs g) {
  pdl_grid_sync();
  const HeadsPadSeg sg = g.seg[blockIdx.y];
  const bf16* __restrict__ src = sg.src;
  bf16* __restrict__ dst = sg.dst;
  const int L = sg.L, ld = sg.ld, H = g.H, d = g.d;
  const long long tok = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (tok >= static_cast<long long>(g.B) * L) return;
  const int l = static_cast<int>(tok % L), b = static_cast<int>(tok / L);
  const bool mine = lane * 4 < d;
  for (int h = 0; h < H; ++h) {
    const size_t hm = ((static_cast<size_t>(b) * H + h) * L + l) * 128 + lane * 4;
    const size_t tm = static_cast<size_t>(tok) * ld + h * d + lane * 4;
    if (g.to_heads) {
      uint2 v = make_uint2(0u, 0u);
      if (mine) v = *reinterpret_cast<const uint2*>(src + tm);
      *reinterpret_cast<uint2*>(dst + hm) = v;
    } else if (mine) {
      *reinterpret_cast<uint2*>(dst + tm) = *reinterpret_cast<const uint2*>(src + hm);
    }
  }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_ln_affine_fwd(b200_ctx* ctx, const void* x, int ldx, const void* weight, const void* bias, void* out, int ldo,
                                  void* mean, void* rstd, int M, int D, float eps, void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  B200_REQUIRE(x && weight && bias && out && M > 0 && D > 0 && D % 8 == 0 && ldx % 8 == 0 && ldo % 8 == 0,
               "b200_ln_affine_fwd: bad args M=%d D=%d", M, D);
  B200_KLAUNCH(ln_affine_fwd_kernel, (M + 7) / 8, 256, 0, reinterpret_cast<cudaStream_t>(stream), (const bf16*)x, ldx,
               (const bf16*)weight, (const bf16*)bias, (bf16*)out, ldo, (float*)mean, (float*)rstd, M, D, eps);
  B200_CUDA_CHECK(cudaGetLastError());
  ctx->launches.fetch_add(1);
  return B200_OK;
}

extern "C" int b200_ln_affine_bwd(b200_ctx* ctx, const void* dy, int lddy, const void* x, int ldx, const void* mean,
                                  const void* rstd, const void* weight, const void* dres, int lddres, void* out, int ldo, int M,
                                  int D, void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  B200_REQUIRE(dy && x && mean && rstd && weight && out && M > 0 && D % 8 == 0 && lddy % 8 == 0 && ldx % 8 == 0 && ldo % 8 == 0,
               "b200_ln_affine_bwd: bad args M=%d D=%d", M, D);
  B200_KLAUNCH(ln_affine_bwd_kernel, (M + 7) / 8, 256, 0, reinterpret_cast<cudaStream_t>(stream), (const bf16*)dy, lddy,
               (const bf16*)x, ldx, (const float*)mean, (const float*)rstd, (const bf16*)weight, (const bf16*)dres, lddres,
               (bf16*)out, ldo, M, D);
  B200_CUDA_CHECK(cudaGetLastError());
  ctx->launches.fetch_add(1);
  return B200_OK;
}

extern "C" int b200_groupnorm_fwd(b200_ctx* ctx, const void* x, const void* weight, const void* bias, void* out, void* mean,
                                  void* rstd, int B, int C, int HW, int G, float eps, int silu, void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  B200_REQUIRE(x && weight && bias && out && mean && rstd && B > 0 && G > 0 && C % G == 0 && HW % 2 == 0,
               "b200_groupnorm_fwd: bad args C=%d G=%d HW=%d", C, G, HW);
  B200_KLAUNCH(groupnorm_fwd_kernel, B * G, 512, 0, reinterpret_cast<cudaStream_t>(stream), (const bf16*)x, (const bf16*)weight,
               (const bf16*)bias, (bf16*)out, (float*)mean, (float*)rstd, C, HW, G, eps, silu);
  B200_CUDA_CHECK(cudaGetLastError());
  ctx->launches.fetch_add(1);
  return B200_OK;
}

extern "C" int b200_groupnorm_bwd(b200_ctx* ctx, const void* dy, const void* x, const void* weight, const void* bias,
                                  const void* mean, const void* rstd, void* dx, int B, int C, int HW, int G, int silu,
                                  void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  B200_REQUIRE(dy && x && weight && bias && mean && rstd && dx && B > 0 && G > 0 && C % G == 0 && HW % 2 == 0,
               "b200_groupnorm_bwd: bad args");
  B200_KLAUNCH(groupnorm_bwd_kernel, B * G, 512, 0, reinterpret_cast<cudaStream_t>(stream), (const bf16*)dy, (const bf16*)x,
               (const bf16*)weight, (const bf16*)bias, (const float*)mean, (const float*)rstd, (bf16*)dx, C, HW, G, silu);
  B200_CUDA_CHECK(cudaGetLastError());
  ctx->launches.fetch_add(1);
  return B200_OK;
}

extern "C" int b200_geglu_fwd(b200_ctx* ctx, const void* proj, int ldp, void* out, int ldo, int64_t M, int F, void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  B200_REQUIRE(proj && out && M > 0 && F > 0 && F % 8 == 0 && ldp % 8 == 0 && ldo % 8 == 0, "b200_geglu_fwd: bad args");
  const long long n = static_cast<long long>(M) * (F / 8);
  B200_KLAUNCH(geglu_fwd_kernel, static_cast<unsigned>((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream),
               (const bf16*)proj, ldp, (bf16*)out, ldo, static_cast<long long>(M), F);
  B200_CUDA_CHECK(cudaGetLastError());
  ctx->launches.fetch_add(1);
  return B200_OK;
}

extern "C" int b200_geglu_bwd(b200_ctx* ctx, const void* dy, int lddy, const void* proj, int ldp, void* dproj, int lddp, int64_t M,
                              int F, void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  B200_REQUIRE(dy && proj && dproj && M > 0 && F % 8 == 0 && lddy % 8 == 0 && ldp % 8 == 0 && lddp % 8 == 0,
               "b200_geglu_bwd: bad args");
  const long long n = static_cast<long long>(M) * (F / 8);
  B200_KLAUNCH(geglu_bwd_kernel, static_cast<unsigned>((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream),
               (const bf16*)dy, lddy, (const bf16*)proj, ldp, (bf16*)dproj, lddp, static_cast<long long>(M), F);
  B200_CUDA_CHECK(cudaGetLastError());
  ctx->launches.fetch_add(1);
  return B200_OK;
}

extern "C" int b200_heads_pad3(b200_ctx* ctx, const void* src0, void* dst0, int ld0, int L0, const void* src1, void* dst1, int ld1,
                               int L1, const void* src2, void* dst2, int ld2, int L2, int n, int B, int H, int head_dim,
                               int to_heads, void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  B200_REQUIRE(n >= 1 && n <= 3 && B > 0 && H > 0 && head_dim > 0 && head_dim <= 128 && head_dim % 4 == 0,
               "b200_heads_pad3: bad args n=%d head_dim=%d (<= 128, multiple of 4)", n, head_dim);
  HeadsPadArgs a = {};
  const void* srcs[3] = {src0, src1, src2};
  void* dsts[3] = {dst0, dst1, dst2};
  const int lds[3] = {ld0, ld1, ld2}, Ls[3] = {L0, L1, L2};
  long long rows = 0;
  for (int i = 0; i < n; ++i) {
    B200_REQUIRE(srcs[i] && dsts[i] && Ls[i] > 0 && lds[i] % 4 == 0, "b200_heads_pad3: tensor %d: null pointer, L <= 0 or ld %% 4", i);
    a.seg[i] = HeadsPadSeg{(const bf16*)srcs[i], (bf16*)dsts[i], lds[i], Ls[i]};
    rows = max(rows, static_cast<long long>(B) * Ls[i]);
  }
  a.B = B; a.H = H; a.d = head_dim; a.to_heads = to_heads;
  dim3 grid(static_cast<unsigned>((rows + 7) / 8), n);
  B200_KLAUNCH(heads_pad_kernel, grid, 256, 0, reinterpret_cast<cudaStream_t>(stream), a);
  B200_CUDA_CHECK(cudaGetLastError());
  ctx->launches.fetch_add(1);
  return B200_OK;
}

extern "C" int b200_heads_pad(b200_ctx* ctx, const void* src, void* dst, int ld, int B, int L, int H, int head_dim, int to_heads,
                              void* stream) {
  return b200_heads_pad3(ctx, src, dst, ld, L, nullptr, nullptr, 0, 0, nullptr, nullptr, 0, 0, 1, B, H, head_dim, to_heads, stream);
}
