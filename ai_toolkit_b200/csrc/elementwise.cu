// HBM-bound row kernels of the DiT block: AdaLN-Zero modulation (LayerNorm + scale/shift) forward and
// backward, the per-sample column reductions that produce the modulation-vector gradients, QK-RMSNorm +
// RoPE + head-major re-layout (forward and backward), SiLU and the sinusoidal timestep embedding.
//
// One warp owns one row (or one (token, head) pair): 16-byte vector loads, fp32 math, warp-shuffle
// reductions; nothing is staged through shared memory because every element is touched exactly once.
// Algorithmic bytes per element are stated at each entry point (include/b200_lora.h).
//
// Reference arithmetic restated here (the model code is third-party `diffusers`; in-tree equivalents):
//   AdaLN-Zero modulation ...... extensions_built_in/diffusion_models/chroma/src/layers.py:471-560
//   QK RMSNorm ................. extensions_built_in/diffusion_models/chroma/src/layers.py:72-91
//   RoPE ....................... extensions_built_in/diffusion_models/chroma/src/math.py:33-51
//   timestep embedding ......... extensions_built_in/diffusion_models/chroma/src/layers.py:30-53
#include "common.cuh"
#include "ctx.h"

namespace b200 {

__device__ __forceinline__ void ld8(const bf16* p, float (&v)[8]) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
__device__ __forceinline__ void st8(bf16* p, const float (&v)[8]) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]);
  u.y = pack_bf16x2(v[2], v[3]);
  u.z = pack_bf16x2(v[4], v[5]);
  u.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm (no affine, eps) + modulation:  y = bf16( bf16( bf16((x-mean)*rstd) * bf16(1+scale) ) + shift )
// NCH = D / 256 (each lane owns NCH chunks of 8 consecutive elements, chunk c at column c*256 + lane*8).
// ------------------------------------------------------------------------------------------------
// Round 2: the row is STAGED IN SHARED MEMORY by cp.async (16 bytes per lane and chunk, no registers), read once for the
// statistics and once for the output -- round 1 held fp32 copies in registers (124 / 254 registers at D = 3072), i.e. ONE
// 8-warp block per SM whose warps all load, then all compute, then all store: 2.4-2.6 TB/s (profiles/r2_time_rows.log).
// Each lane reads back only the bytes it copied itself, so no barrier is needed beyond cp.async.wait_all.  Same arithmetic
// in the same order as round 1.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <int NCH>
__global__ void __launch_bounds__(256, NCH <= 12 ? 2 : 1) ln_modulate_fwd_kernel(const bf16* __restrict__ x, int ldx, const bf16* __restrict__ shift,
                                                              const bf16* __restrict__ scale, int ldmod,
                                                              int rows_per_sample, bf16* __restrict__ out, int ldo,
                                                              float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                              int M, float eps) {
  extern __shared__ __align__(16) uint8_t ln_rows[];
  pdl_grid_sync();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  constexpr int D = NCH * 256;
  bf16* sx = reinterpret_cast<bf16*>(ln_rows) + (threadIdx.x >> 5) * D;
  const bf16* xr = x + static_cast<size_t>(row) * ldx;
#pragma unroll
  for (int c = 0; c < NCH; ++c) cp_async16(sx + c * 256 + lane * 8, xr + c * 256 + lane * 8);
  cp_async_wait_all();
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    float v[8];
    ld8(sx + c * 256 + lane * 8, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i];
  }
  const float mean = warp_sum(s) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    float v[8];
    ld8(sx + c * 256 + lane * 8, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float d = v[i] - mean;
      q += d * d;
    }
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + eps);
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
  const int sample = row / rows_per_sample;
  const bf16* sh = shift ? shift + static_cast<size_t>(sample) * ldmod : nullptr;
  const bf16* sc = scale ? scale + static_cast<size_t>(sample) * ldmod : nullptr;
  bf16* orow = out + static_cast<size_t>(row) * ldo;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int col = c * 256 + lane * 8;
    float v[8], o[8];
    float a[8], b[8];
    ld8(sx + col, v);
    if (sc) ld8(sc + col, a);
    if (sh) ld8(sh + col, b);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float y = bf16_round((v[i] - mean) * rstd);
      if (sc) y = bf16_round(y * bf16_round(1.0f + a[i]));
      if (sh) y = bf16_round(y + b[i]);
      o[i] = y;
    }
    st8(orow + col, o);
  }
}

// Backward of the above w.r.t. x, fused with the residual-stream gradient:
//   dxhat = dy * (1 + scale);  dx = rstd * (dxhat - mean(dxhat) - xhat * mean(dxhat * xhat));  out = dres + dx
template <int NCH>
__global__ void __launch_bounds__(256, NCH <= 12 ? 2 : 1) ln_modulate_bwd_kernel(const bf16* __restrict__ dy, int lddy, const bf16* __restrict__ x,
                                                              int ldx, const float* __restrict__ mean_in,
                                                              const float* __restrict__ rstd_in,
                                                              const bf16* __restrict__ scale, int ldmod,
                                                              int rows_per_sample, const bf16* __restrict__ dres, int lddres,
                                                              bf16* __restrict__ out, int ldo, int M) {
  extern __shared__ __align__(16) uint8_t ln_rows[];
  pdl_grid_sync();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  constexpr int D = NCH * 256;
  bf16* sx = reinterpret_cast<bf16*>(ln_rows) + (threadIdx.x >> 5) * (2 * D);
  bf16* sg = sx + D;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int col = c * 256 + lane * 8;
    cp_async16(sx + col, x + static_cast<size_t>(row) * ldx + col);
    cp_async16(sg + col, dy + static_cast<size_t>(row) * lddy + col);
  }
  const float mean = mean_in[row], rstd = rstd_in[row];
  const int sample = row / rows_per_sample;
  const bf16* sc = scale ? scale + static_cast<size_t>(sample) * ldmod : nullptr;
  cp_async_wait_all();
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int col = c * 256 + lane * 8;
    float xv[8], gv[8], a[8];
    ld8(sx + col, xv);
    ld8(sg + col, gv);
    if (sc) ld8(sc + col, a);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float xh = (xv[i] - mean) * rstd;
      float gg = gv[i];
      if (sc) gg *= bf16_round(1.0f + a[i]);
      s1 += gg;
      s2 += gg * xh;
    }
  }
  s1 = warp_sum(s1) * (1.0f / D);
  s2 = warp_sum(s2) * (1.0f / D);
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int col = c * 256 + lane * 8;
    float xv[8], gv[8], a[8], o[8], r[8];
    if (dres) ld8(dres + static_cast<size_t>(row) * lddres + col, r);
    ld8(sx + col, xv);
    ld8(sg + col, gv);
    if (sc) ld8(sc + col, a);  // (the per-sample scale row is L1 / L2 resident)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float xh = (xv[i] - mean) * rstd;
      float gg = gv[i];
      if (sc) gg *= bf16_round(1.0f + a[i]);
      const float d = rstd * (gg - s1 - xh * s2);
      o[i] = dres ? d + r[i] : d;
    }
    st8(out + static_cast<size_t>(row) * ldo + col, o);
  }
}

// ------------------------------------------------------------------------------------------------
// Per-sample column reductions (gradients of the modulation vector), optional fused product output:
//   sum_a [s, d] += sum_rows a[m, d]
//   sum_ab[s, d] += sum_rows a[m, d] * f(b[m, d]),   f(b) = (b - mean[m]) * rstd[m]  or  b
//   mul_out[m, d] = bf16(a[m, d] * g[s, d])                                  (optional)
// Block: 256 threads = 256 column pairs (512 columns) x ROWS rows of one sample; fp32 atomics to [S, D].
// (Round 2 tried a 16-byte-per-thread variant, 128 threads x 8 columns x 32 rows with 4 rows of loads in flight: slower on the
// B200 -- 22.6 vs 15.9 us back-to-back, 38.9 vs 30.7 us L2-flushed at M = 4608, D = 3072 (profiles/r2_time_rows.log) -- removed.)
// ------------------------------------------------------------------------------------------------
constexpr int kColRows = 64;
__global__ void __launch_bounds__(256) col_reduce_kernel(const bf16* __restrict__ a, int lda, const bf16* __restrict__ b, int ldb,
                                                         const float* __restrict__ mean, const float* __restrict__ rstd,
                                                         const bf16* __restrict__ g, int ldg, bf16* __restrict__ mul_out,
                                                         int ldmul, float* __restrict__ sum_a, float* __restrict__ sum_ab,
                                                         int ldsum, int rows_per_sample, int M, int D) {
  pdl_grid_sync();
  const int col = (blockIdx.x * 256 + threadIdx.x) * 2;
  if (col >= D) return;
  const int chunks_per_sample = (rows_per_sample + kColRows - 1) / kColRows;
  const int sample = blockIdx.y / chunks_per_sample;
  const int r0 = sample * rows_per_sample + (blockIdx.y % chunks_per_sample) * kColRows;
  const int r1 = min(min(r0 + kColRows, (sample + 1) * rows_per_sample), M);
  float2 gv = make_float2(0.f, 0.f);
  if (g) gv = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(g + static_cast<size_t>(sample) * ldg + col));
  float sa0 = 0.f, sa1 = 0.f, sb0 = 0.f, sb1 = 0.f;
#pragma unroll 4
  for (int r = r0; r < r1; ++r) {
    const float2 av = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(a + static_cast<size_t>(r) * lda + col));
    sa0 += av.x;
    sa1 += av.y;
    if (b) {
      float2 bv = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(b + static_cast<size_t>(r) * ldb + col));
      if (mean) {
        const float mu = mean[r], rs = rstd[r];
        bv.x = (bv.x - mu) * rs;
        bv.y = (bv.y - mu) * rs;
      }
      sb0 += av.x * bv.x;
      sb1 += av.y * bv.y;
    }
    if (mul_out)
      *reinterpret_cast<uint32_t*>(mul_out + static_cast<size_t>(r) * ldmul + col) = pack_bf16x2(av.x * gv.x, av.y * gv.y);
  }
  if (sum_a) {
    atomicAdd(sum_a + static_cast<size_t>(sample) * ldsum + col, sa0);
    atomicAdd(sum_a + static_cast<size_t>(sample) * ldsum + col + 1, sa1);
  }
  if (sum_ab) {
    atomicAdd(sum_ab + static_cast<size_t>(sample) * ldsum + col, sb0);
    atomicAdd(sum_ab + static_cast<size_t>(sample) * ldsum + col + 1, sb1);
  }
}

// ------------------------------------------------------------------------------------------------
// QK-RMSNorm + RoPE + re-layout to head-major [B, H, Ltot, 128].  One warp per (token, head);
// lane owns elements 4*lane .. 4*lane+3 (two rotary pairs).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ld4(const bf16* p, float (&v)[4]) {
  uint2 u = *reinterpret_cast<const uint2*>(p);
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
__device__ __forceinline__ void st4(bf16* p, const float (&v)[4]) {
  uint2 u;
  u.x = pack_bf16x2(v[0], v[1]);
  u.y = pack_bf16x2(v[2], v[3]);
  *reinterpret_cast<uint2*>(p) = u;
}

__global__ void __launch_bounds__(256) qk_norm_rope_fwd_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k,
                                                               const bf16* __restrict__ v, int ld, const bf16* __restrict__ wq,
                                                               const bf16* __restrict__ wk, const float* __restrict__ cos_t,
                                                               const float* __restrict__ sin_t, bf16* __restrict__ Q,
                                                               bf16* __restrict__ K, bf16* __restrict__ V, int B, int Lseg,
                                                               int seq_off, int Ltot, int H, float eps) {
  pdl_grid_sync();
  const long long w = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const long long total = static_cast<long long>(B) * Lseg * H;
  if (w >= total) return;
  const int h = static_cast<int>(w % H);
  const long long tok = w / H;  // b * Lseg + l
  const int l = static_cast<int>(tok % Lseg);
  const int b = static_cast<int>(tok / Lseg);
  const int pos = seq_off + l;
  const size_t src = static_cast<size_t>(tok) * ld + h * 128 + lane * 4;
  const size_t dst = ((static_cast<size_t>(b) * H + h) * Ltot + pos) * 128 + lane * 4;
  float c[4], s[4];
  *reinterpret_cast<float4*>(c) = *reinterpret_cast<const float4*>(cos_t + static_cast<size_t>(pos) * 128 + lane * 4);
  *reinterpret_cast<float4*>(s) = *reinterpret_cast<const float4*>(sin_t + static_cast<size_t>(pos) * 128 + lane * 4);
#pragma unroll
  for (int which = 0; which < 2; ++which) {
    const bf16* in = which == 0 ? q : k;
    const bf16* wgt = which == 0 ? wq : wk;
    bf16* out = which == 0 ? Q : K;
    float x[4], wv[4];
    ld4(in + src, x);
    ld4(wgt + lane * 4, wv);
    float ss = x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3];
    const float rstd = rsqrtf(warp_sum(ss) * (1.0f / 128.0f) + eps);
    float y[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) y[i] = bf16_round(bf16_round(x[i] * rstd) * wv[i]);
    float o[4];
    o[0] = y[0] * c[0] - y[1] * s[0];
    o[1] = y[1] * c[1] + y[0] * s[1];
    o[2] = y[2] * c[2] - y[3] * s[2];
    o[3] = y[3] * c[3] + y[2] * s[3];
    st4(out + dst, o);
  }
  *reinterpret_cast<uint2*>(V + dst) = *reinterpret_cast<const uint2*>(v + src);
}

__global__ void __launch_bounds__(256) qk_norm_rope_bwd_kernel(const bf16* __restrict__ dQ, const bf16* __restrict__ dK,
                                                               const bf16* __restrict__ dV, const bf16* __restrict__ q,
                                                               const bf16* __restrict__ k, int ld, const bf16* __restrict__ wq,
                                                               const bf16* __restrict__ wk, const float* __restrict__ cos_t,
                                                               const float* __restrict__ sin_t, bf16* __restrict__ dq,
                                                               bf16* __restrict__ dk, bf16* __restrict__ dv, int ldd, int B,
                                                               int Lseg, int seq_off, int Ltot, int H, float eps) {
  pdl_grid_sync();
  const long long w = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const long long total = static_cast<long long>(B) * Lseg * H;
  if (w >= total) return;
  const int h = static_cast<int>(w % H);
  const long long tok = w / H;
  const int l = static_cast<int>(tok % Lseg);
  const int b = static_cast<int>(tok / Lseg);
  const int pos = seq_off + l;
  const size_t src = static_cast<size_t>(tok) * ld + h * 128 + lane * 4;
  const size_t dsto = static_cast<size_t>(tok) * ldd + h * 128 + lane * 4;
  const size_t hm = ((static_cast<size_t>(b) * H + h) * Ltot + pos) * 128 + lane * 4;
  float c[4], s[4];
  *reinterpret_cast<float4*>(c) = *reinterpret_cast<const float4*>(cos_t + static_cast<size_t>(pos) * 128 + lane * 4);
  *reinterpret_cast<float4*>(s) = *reinterpret_cast<const float4*>(sin_t + static_cast<size_t>(pos) * 128 + lane * 4);
#pragma unroll
  for (int which = 0; which < 2; ++which) {
    const bf16* in = which == 0 ? q : k;
    const bf16* wgt = which == 0 ? wq : wk;
    const bf16* din = which == 0 ? dQ : dK;
    bf16* out = which == 0 ? dq : dk;
    float x[4], wv[4], g[4];
    ld4(in + src, x);
    ld4(wgt + lane * 4, wv);
    ld4(din + hm, g);
    // RoPE^T
    float dy[4];
    dy[0] = g[0] * c[0] + g[1] * s[1];
    dy[1] = g[1] * c[1] - g[0] * s[0];
    dy[2] = g[2] * c[2] + g[3] * s[3];
    dy[3] = g[3] * c[3] - g[2] * s[2];
    float ss = x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3];
    const float rstd = rsqrtf(warp_sum(ss) * (1.0f / 128.0f) + eps);
    float xn[4], dxn[4];
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      xn[i] = x[i] * rstd;
      dxn[i] = dy[i] * wv[i];
      dot += dxn[i] * xn[i];
    }
    dot = warp_sum(dot) * (1.0f / 128.0f);
    float o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = rstd * (dxn[i] - xn[i] * dot);
    st4(out + dsto, o);
  }
  *reinterpret_cast<uint2*>(dv + dsto) = *reinterpret_cast<const uint2*>(dV + hm);
}

// ------------------------------------------------------------------------------------------------
__global__ void silu_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, long long n) {
  pdl_grid_sync();
  const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 2;
  if (i + 1 < n) {
    float2 v = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(x + i));
    *reinterpret_cast<uint32_t*>(y + i) = pack_bf16x2(silu_f(v.x), silu_f(v.y));
  } else if (i < n) {
    y[i] = __float2bfloat16_rn(silu_f(__bfloat162float(x[i])));
  }
}

// out[b, :] = bf16([cos(t*f) | sin(t*f)]), f_i = exp(-ln(max_period) * i / half), t = bf16(bf16(t_in / div) * mult)
// (the trainer passes timestep / 1000 in fp32, the bf16 model multiplies by 1000 in bf16)
__global__ void timestep_embed_kernel(const float* __restrict__ t_in, bf16* __restrict__ out, int B, int dim, float max_period,
                                      float div, float mult) {
  pdl_grid_sync();
  const int b = blockIdx.x;
  const int half = dim / 2;
  // mult > 0: the bf16 model re-scales in bf16 (FLUX).  mult <= 0: t = t_in / div stays fp32 (Wan2.1: `timesteps_proj` runs on
  // the trainer's fp32 timestep, 0..1000)
  const float t = mult > 0.f ? bf16_round(bf16_round(t_in[b] / div) * mult) : t_in[b] / div;
  for (int i = threadIdx.x; i < half; i += blockDim.x) {
    const float f = expf(-logf(max_period) * static_cast<float>(i) / static_cast<float>(half));
    const float ang = t * f;
    out[static_cast<size_t>(b) * dim + i] = __float2bfloat16_rn(cosf(ang));
    out[static_cast<size_t>(b) * dim + half + i] = __float2bfloat16_rn(sinf(ang));
  }
}

// out = bf16(a + b) (+ c)
__global__ void add_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, const bf16* __restrict__ c,
                           bf16* __restrict__ y, long long n) {
  pdl_grid_sync();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = bf16_round(__bfloat162float(a[i]) + __bfloat162float(b[i]));
  if (c) v = bf16_round(v + __bfloat162float(c[i]));
  y[i] = __float2bfloat16_rn(v);
}

#define B200_LN_DISPATCH(D_, CALL)                                                                  \
  switch ((D_) / 256) {                                                                             \
    case 1: CALL(1); break;                                                                         \
    case 2: CALL(2); break;                                                                         \
    case 3: CALL(3); break;                                                                         \
    case 4: CALL(4); break;                                                                         \
    case 6: CALL(6); break;                                                                         \
    case 8: CALL(8); break;                                                                         \
    case 12: CALL(12); break;                                                                       \
    case 16: CALL(16); break;                                                                       \
    case 20: CALL(20); break;                                                                       \
    default: set_error("LayerNorm width %d unsupported (need D/256 in {1,2,3,4,6,8,12,16,20})", (D_)); \
      return B200_ERR_INVALID;                                                                      \
  }

}  // namespace b200

using namespace b200;

extern "C" int b200_ln_modulate_fwd(b200_ctx* ctx, const void* x, int ldx, const void* shift, const void* scale, int ldmod,
                                    int rows_per_sample, void* out, int ldo, void* mean, void* rstd, int M, int D, float eps,
                                    void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  B200_REQUIRE(x && out && M > 0 && D > 0 && D % 256 == 0, "b200_ln_modulate_fwd: bad args M=%d D=%d", M, D);
  B200_REQUIRE(ldx % 8 == 0 && ldo % 8 == 0 && ldmod % 8 == 0 && rows_per_sample > 0, "b200_ln_modulate_fwd: alignment");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int grid = (M + 7) / 8;
#define CALL(N_)                                                                                                         \
  do {                                                                                                                   \
    auto kern = ln_modulate_fwd_kernel<N_>;                                                                              \
    constexpr int kSm = 8 * N_ * 256 * 2;                                                                                \
    static bool configured = false;                                                                                      \
    if (!configured) {                                                                                                   \
      B200_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSm));                     \
      configured = true;                                                                                                 \
    }                                                                                                                    \
    B200_KLAUNCH(kern, grid, 256, kSm, st, (const bf16*)x, ldx, (const bf16*)shift, (const bf16*)scale, ldmod,           \
                 rows_per_sample, (bf16*)out, ldo, (float*)mean, (float*)rstd, M, eps);                                  \
  } while (0)
  B200_LN_DISPATCH(D, CALL)
#undef CALL
  B200_CUDA_CHECK(cudaGetLastError());
  ctx->launches.fetch_add(1);
  return B200_OK;
}

extern "C" int b200_ln_modulate_bwd(b200_ctx* ctx, const void* dy, int lddy, const void* x, int ldx, const void* mean,
                                    const void* rstd, const void* scale, int ldmod, int rows_per_sample, const void* dres,
                                    int lddres, void* out, int ldo, int M, int D, void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  B200_REQUIRE(dy && x && mean && rstd && out && M > 0 && D % 256 == 0, "b200_ln_modulate_bwd: bad args M=%d D=%d", M, D);
  B200_REQUIRE(lddy % 8 == 0 && ldx % 8 == 0 && ldo % 8 == 0 && rows_per_sample > 0, "b200_ln_modulate_bwd: alignment");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int grid = (M + 7) / 8;
#define CALL(N_)                                                                                                         \
  do {                                                                                                                   \
    auto kern = ln_modulate_bwd_kernel<N_>;                                                                              \
    constexpr int kSm = 8 * 2 * N_ * 256 * 2;                                                                            \
    static bool configured = false;                                                                                      \
    if (!configured) {                                                                                                   \
      B200_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSm));                     \
      configured = true;                                                                                                 \
    }                                                                                                                    \
    B200_KLAUNCH(kern, grid, 256, kSm, st, (const bf16*)dy, lddy, (const bf16*)x, ldx, (const float*)mean,               \
                 (const float*)rstd, (const bf16*)scale, ldmod, rows_per_sample, (const bf16*)dres, lddres, (bf16*)out,  \
                 ldo, M);                                                                                                \
  } while (0)
  B200_LN_DISPATCH(D, CALL)
#undef CALL
  B200_CUDA_CHECK(cudaGetLastError());
  ctx->launches.fetch_add(1);
  return B200_OK;
}

extern "C" int b200_col_reduce(b200_ctx* ctx, const void* a, int lda, const void* b, int ldb, const void* mean,
                               const void* rstd, const void* g, int ldg, void* mul_out, int ldmul, void* sum_a, void* sum_ab,
                               int ldsum, int rows_per_sample, int M, int D, void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  B200_REQUIRE(a && M > 0 && D > 0 && D % 2 == 0 && rows_per_sample > 0, "b200_col_reduce: bad args");
  B200_REQUIRE(lda % 2 == 0 && ldb % 2 == 0 && ldg % 2 == 0 && ldmul % 2 == 0, "b200_col_reduce: leading dims must be even");
  B200_REQUIRE((mean == nullptr) == (rstd == nullptr), "b200_col_reduce: mean and rstd go together");
  if (mul_out) B200_REQUIRE(g != nullptr, "b200_col_reduce: mul_out needs g");
  const int samples = (M + rows_per_sample - 1) / rows_per_sample;
  const int chunks = (rows_per_sample + kColRows - 1) / kColRows;
  dim3 grid((D / 2 + 255) / 256, samples * chunks);
  B200_KLAUNCH(col_reduce_kernel, grid, 256, 0, reinterpret_cast<cudaStream_t>(stream), 
      (const bf16*)a, lda, (const bf16*)b, ldb, (const float*)mean, (const float*)rstd, (const bf16*)g, ldg, (bf16*)mul_out,
      ldmul, (float*)sum_a, (float*)sum_ab, ldsum, rows_per_sample, M, D);
  B200_CUDA_CHECK(cudaGetLastError());
  ctx->launches.fetch_add(1);
  return B200_OK;
}

extern "C" int b200_qk_norm_rope_fwd(b200_ctx* ctx, const void* q, const void* k, const void* v, int ld, const void* wq,
                                     const void* wk, const void* cos_t, const void* sin_t, void* Q, void* K, void* V, int B,
                                     int Lseg, int seq_off, int Ltot, int H, int head_dim, float eps, void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  B200_REQUIRE(head_dim == 128, "b200_qk_norm_rope_fwd: head_dim %d (only 128)", head_dim);
  B200_REQUIRE(q && k && v && wq && wk && cos_t && sin_t && Q && K && V && ld % 4 == 0, "b200_qk_norm_rope_fwd: bad args");
  B200_REQUIRE(seq_off >= 0 && seq_off + Lseg <= Ltot, "b200_qk_norm_rope_fwd: segment out of range");
  const long long warps = static_cast<long long>(B) * Lseg * H;
  const unsigned grid = static_cast<unsigned>((warps + 7) / 8);
  B200_KLAUNCH(qk_norm_rope_fwd_kernel, grid, 256, 0, reinterpret_cast<cudaStream_t>(stream), 
      (const bf16*)q, (const bf16*)k, (const bf16*)v, ld, (const bf16*)wq, (const bf16*)wk, (const float*)cos_t,
      (const float*)sin_t, (bf16*)Q, (bf16*)K, (bf16*)V, B, Lseg, seq_off, Ltot, H, eps);
  B200_CUDA_CHECK(cudaGetLastError());
  ctx->launches.fetch_add(1);
  return B200_OK;
}

extern "C" int b200_qk_norm_rope_bwd(b200_ctx* ctx, const void* dQ, const void* dK, const void* dV, const void* q,
                                     const void* k, int ld, const void* wq, const void* wk, const void* cos_t,
                                     const void* sin_t, void* dq, void* dk, void* dv, int ldd, int B, int Lseg, int seq_off,
                                     int Ltot, int H, int head_dim, float eps, void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  B200_REQUIRE(head_dim == 128, "b200_qk_norm_rope_bwd: head_dim %d (only 128)", head_dim);
  B200_REQUIRE(dQ && dK && dV && q && k && wq && wk && dq && dk && dv && ld % 4 == 0 && ldd % 4 == 0,
               "b200_qk_norm_rope_bwd: bad args");
  const long long warps = static_cast<long long>(B) * Lseg * H;
  const unsigned grid = static_cast<unsigned>((warps + 7) / 8);
  B200_KLAUNCH(qk_norm_rope_bwd_kernel, grid, 256, 0, reinterpret_cast<cudaStream_t>(stream), 
      (const bf16*)dQ, (const bf16*)dK, (const bf16*)dV, (const bf16*)q, (const bf16*)k, ld, (const bf16*)wq,
      (const bf16*)wk, (const float*)cos_t, (const float*)sin_t, (bf16*)dq, (bf16*)dk, (bf16*)dv, ldd, B, Lseg, seq_off,
      Ltot, H, eps);
  B200_CUDA_CHECK(cudaGetLastError());
  ctx->launches.fetch_add(1);
  return B200_OK;
}

extern "C" int b200_silu(b200_ctx* ctx, const void* x, void* y, int64_t n, void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  B200_REQUIRE(x && y && n > 0, "b200_silu: bad args");
  const unsigned grid = static_cast<unsigned>(((n + 1) / 2 + 255) / 256);
  B200_KLAUNCH(silu_kernel, grid, 256, 0, reinterpret_cast<cudaStream_t>(stream), (const bf16*)x, (bf16*)y, n);
  B200_CUDA_CHECK(cudaGetLastError());
  ctx->launches.fetch_add(1);
  return B200_OK;
}

extern "C" int b200_timestep_embed(b200_ctx* ctx, const void* t01, void* out, int B, int dim, float max_period, float div,
                                   float mult, void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  B200_REQUIRE(t01 && out && B > 0 && dim > 0 && dim % 2 == 0 && div != 0.f, "b200_timestep_embed: bad args");
  B200_KLAUNCH(timestep_embed_kernel, B, 128, 0, reinterpret_cast<cudaStream_t>(stream), (const float*)t01, (bf16*)out, B, dim,
                                                                               max_period, div, mult);
  B200_CUDA_CHECK(cudaGetLastError());
  ctx->launches.fetch_add(1);
  return B200_OK;
}

extern "C" int b200_add_bf16(b200_ctx* ctx, const void* a, const void* b, const void* c, void* y, int64_t n, void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  B200_REQUIRE(a && b && y && n > 0, "b200_add_bf16: bad args");
  const unsigned grid = static_cast<unsigned>((n + 255) / 256);
  B200_KLAUNCH(add_kernel, grid, 256, 0, reinterpret_cast<cudaStream_t>(stream), (const bf16*)a, (const bf16*)b, (const bf16*)c,
                                                                       (bf16*)y, n);
  B200_CUDA_CHECK(cudaGetLastError());
  ctx->launches.fetch_add(1);
  return B200_OK;
}
