// Attention for the head dims the tcgen05 kernels do not cover (128 < d <= 256: SD1.5's 1280-channel levels have 8 heads of
// 160), on the CUDA cores.  Such heads only occur at the deepest UNet levels, i.e. at 64 .. 1024 tokens (SD1.5 at 512^2: 256 and
// 64), where the whole attention is < 1 GFLOP: a few CTAs of fp32 FMAs, not a TMEM pipeline.  Token-major operands are read as
// they leave the projections (q [B L, H d], k / v [B Lk, H d]; no head-major re-layout), softmax statistics in fp32, same
// math as the reference's F.scaled_dot_product_attention (toolkit/models/wan21/wan_attn.py:70-76 shows the call; diffusers'
// AttnProcessor2_0 makes the same one for the UNet).
//
// Mapping: 4 consecutive lanes own one row (query row in the forward / dQ kernels, key row in the dK/dV kernel), each lane
// DPT = d / 4 of its head-dim channels in registers; the other side streams through shared memory in tiles of 32 rows; dot
// products are completed with two xor-shuffles.  Forward: scores of a tile first, then ONE running-max update per tile.
#include "common.cuh"
#include "ctx.h"

namespace b200 {

constexpr int kSaRows = 32;    // rows per CTA (x 4 lanes = 128 threads)
constexpr int kSaTile = 32;    // streamed rows per shared-memory tile
constexpr float kSaLog2e = 1.4426950408889634f;

struct SmallAttnArgs {
  const bf16 *q, *k, *v, *o, *dO;
  int ldq, ldk, ldv, ldo, lddo;
  bf16 *out, *dq, *dk, *dv;
  int ldout, lddq, lddk, lddv;
  float *lse, *delta;  // [B, H, L]
  int B, H, L, Lk;
  float scale;
};

template <int DPT>
__device__ __forceinline__ void sa_load_row(const bf16* p, bool live, float (&v)[DPT]) {
#pragma unroll
  for (int c = 0; c < DPT / 8; ++c) {
    uint4 u = make_uint4(0u, 0u, 0u, 0u);
    if (live) u = *reinterpret_cast<const uint4*>(p + c * 8);
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), cc = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    v[c * 8 + 0] = a.x; v[c * 8 + 1] = a.y; v[c * 8 + 2] = b.x; v[c * 8 + 3] = b.y;
    v[c * 8 + 4] = cc.x; v[c * 8 + 5] = cc.y; v[c * 8 + 6] = d.x; v[c * 8 + 7] = d.y;
  }
}
template <int DPT>
__device__ __forceinline__ void sa_store_row(bf16* p, const float (&v)[DPT], float mul) {
#pragma unroll
  for (int c = 0; c < DPT / 8; ++c) {
    uint4 u;
    u.x = pack_bf16x2(v[c * 8 + 0] * mul, v[c * 8 + 1] * mul);
    u.y = pack_bf16x2(v[c * 8 + 2] * mul, v[c * 8 + 3] * mul);
    u.z = pack_bf16x2(v[c * 8 + 4] * mul, v[c * 8 + 5] * mul);
    u.w = pack_bf16x2(v[c * 8 + 6] * mul, v[c * 8 + 7] * mul);
    *reinterpret_cast<uint4*>(p + c * 8) = u;
  }
}
// dot of my DPT channels with the same channels of a bf16 row in shared memory
template <int DPT>
__device__ __forceinline__ float sa_dot(const float (&a)[DPT], const bf16* srow) {
  float s0 = 0.f, s1 = 0.f;  // two chains
#pragma unroll
  for (int c = 0; c < DPT / 8; ++c) {
    const uint4 u = *reinterpret_cast<const uint4*>(srow + c * 8);
    const float2 x0 = unpack_bf16x2(u.x), x1 = unpack_bf16x2(u.y), x2 = unpack_bf16x2(u.z), x3 = unpack_bf16x2(u.w);
    s0 += a[c * 8 + 0] * x0.x + a[c * 8 + 2] * x1.x + a[c * 8 + 4] * x2.x + a[c * 8 + 6] * x3.x;
    s1 += a[c * 8 + 1] * x0.y + a[c * 8 + 3] * x1.y + a[c * 8 + 5] * x2.y + a[c * 8 + 7] * x3.y;
  }
  return s0 + s1;
}
// acc += w * (bf16 row in shared memory)
template <int DPT>
__device__ __forceinline__ void sa_axpy(float (&acc)[DPT], float w, const bf16* srow) {
#pragma unroll
  for (int c = 0; c < DPT / 8; ++c) {
    const uint4 u = *reinterpret_cast<const uint4*>(srow + c * 8);
    const float2 x0 = unpack_bf16x2(u.x), x1 = unpack_bf16x2(u.y), x2 = unpack_bf16x2(u.z), x3 = unpack_bf16x2(u.w);
    acc[c * 8 + 0] += w * x0.x; acc[c * 8 + 1] += w * x0.y; acc[c * 8 + 2] += w * x1.x; acc[c * 8 + 3] += w * x1.y;
    acc[c * 8 + 4] += w * x2.x; acc[c * 8 + 5] += w * x2.y; acc[c * 8 + 6] += w * x3.x; acc[c * 8 + 7] += w * x3.y;
  }
}
__device__ __forceinline__ float sa_quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}
// stage rows [r0, r0 + 32) of a token-major [rows, ld] operand (head columns [hc, hc + d)) into shared memory as [32][d] bf16
__device__ __forceinline__ void sa_stage(bf16* dst, const bf16* src, int ld, long long row_base, int r0, int nrows, int d) {
  const int per_row = d / 8;
  for (int idx = threadIdx.x; idx < kSaTile * per_row; idx += blockDim.x) {
    const int r = idx / per_row, c = idx % per_row;
    uint4 u = make_uint4(0u, 0u, 0u, 0u);
    if (r0 + r < nrows) u = *reinterpret_cast<const uint4*>(src + (row_base + r0 + r) * ld + c * 8);
    *reinterpret_cast<uint4*>(dst + r * d + c * 8) = u;
  }
}

// ---------------------------------------------------------------------------------------------------------------- forward
template <int DPT>
__global__ void __launch_bounds__(128) small_attn_fwd_kernel(const SmallAttnArgs g) {
  constexpr int D = DPT * 4;
  extern __shared__ __align__(16) uint8_t sa_smem[];
  bf16* sK = reinterpret_cast<bf16*>(sa_smem);
  bf16* sV = sK + kSaTile * D;
  pdl_grid_sync();
  const int bh = blockIdx.y, b = bh / g.H, h = bh % g.H;
  const int row = blockIdx.x * kSaRows + (threadIdx.x >> 2);
  const int sub = threadIdx.x & 3;
  const bool live = row < g.L;
  const int hc = h * D + sub * DPT;
  float q[DPT], o[DPT];
  sa_load_row<DPT>(g.q + (static_cast<long long>(b) * g.L + row) * g.ldq + hc, live, q);
#pragma unroll
  for (int i = 0; i < DPT; ++i) o[i] = 0.f;
  const float c2 = g.scale * kSaLog2e;
  float m = -INFINITY, l = 0.f;
  const long long kv_base = static_cast<long long>(b) * g.Lk;
  for (int j0 = 0; j0 < g.Lk; j0 += kSaTile) {
    __syncthreads();  // the previous tile has been consumed
    sa_stage(sK, g.k + h * D, g.ldk, kv_base, j0, g.Lk, D);
    sa_stage(sV, g.v + h * D, g.ldv, kv_base, j0, g.Lk, D);
    __syncthreads();
    float s[kSaTile];
    float mt = m;
#pragma unroll
    for (int jb = 0; jb < kSaTile; jb += 4) {  // 4 independent dot chains in flight (one chain is 40 dependent FMAs)
      float part[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) part[u] = sa_dot<DPT>(q, sK + (jb + u) * D + sub * DPT);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float dot = sa_quad_sum(part[u]);
        s[jb + u] = (j0 + jb + u < g.Lk) ? dot * c2 : -INFINITY;
        mt = fmaxf(mt, s[jb + u]);
      }
    }
    const float corr = (m == -INFINITY) ? 0.f : exp2f(m - mt);  // (mt is finite: every tile has at least one live key)
    l *= corr;
#pragma unroll
    for (int i = 0; i < DPT; ++i) o[i] *= corr;
#pragma unroll
    for (int j = 0; j < kSaTile; ++j) {
      const float p = exp2f(s[j] - mt);  // masked keys: exp2(-inf) = 0
      l += p;
      sa_axpy<DPT>(o, p, sV + j * D + sub * DPT);
    }
    m = mt;
  }
  if (live) {
    sa_store_row<DPT>(g.out + (static_cast<long long>(b) * g.L + row) * g.ldout + hc, o, 1.0f / l);
    if (sub == 0) g.lse[static_cast<long long>(bh) * g.L + row] = (m + log2f(l)) / kSaLog2e;  // natural log, like the tcgen05 path
  }
}

// ------------------------------------------------------------------------------------------------- backward: dQ (+ delta)
template <int DPT>
__global__ void __launch_bounds__(128) small_attn_bwd_q_kernel(const SmallAttnArgs g) {
  constexpr int D = DPT * 4;
  extern __shared__ __align__(16) uint8_t sa_smem[];
  bf16* sK = reinterpret_cast<bf16*>(sa_smem);
  bf16* sV = sK + kSaTile * D;
  pdl_grid_sync();
  const int bh = blockIdx.y, b = bh / g.H, h = bh % g.H;
  const int row = blockIdx.x * kSaRows + (threadIdx.x >> 2);
  const int sub = threadIdx.x & 3;
  const bool live = row < g.L;
  const int hc = h * D + sub * DPT;
  const long long qrow = static_cast<long long>(b) * g.L + row;
  float q[DPT], dO[DPT], dq[DPT];
  sa_load_row<DPT>(g.q + qrow * g.ldq + hc, live, q);
  sa_load_row<DPT>(g.dO + qrow * g.lddo + hc, live, dO);
  float delta;
  {
    float ov[DPT];
    sa_load_row<DPT>(g.o + qrow * g.ldo + hc, live, ov);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < DPT; ++i) s += ov[i] * dO[i];
    delta = sa_quad_sum(s);
  }
#pragma unroll
  for (int i = 0; i < DPT; ++i) dq[i] = 0.f;
  const float lse2 = live ? g.lse[static_cast<long long>(bh) * g.L + row] * kSaLog2e : INFINITY;  // dead rows: p = 0
  if (live && sub == 0) g.delta[static_cast<long long>(bh) * g.L + row] = delta;
  const float c2 = g.scale * kSaLog2e;
  const long long kv_base = static_cast<long long>(b) * g.Lk;
  for (int j0 = 0; j0 < g.Lk; j0 += kSaTile) {
    __syncthreads();
    sa_stage(sK, g.k + h * D, g.ldk, kv_base, j0, g.Lk, D);
    sa_stage(sV, g.v + h * D, g.ldv, kv_base, j0, g.Lk, D);
    __syncthreads();
#pragma unroll 2
    for (int jb = 0; jb < kSaTile; jb += 2) {  // 4 x 2 independent dot chains, then the shuffles, then the updates
      float pd[2], pp[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        pd[u] = sa_dot<DPT>(q, sK + (jb + u) * D + sub * DPT);
        pp[u] = sa_dot<DPT>(dO, sV + (jb + u) * D + sub * DPT);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const float dot = sa_quad_sum(pd[u]), dp = sa_quad_sum(pp[u]);
        const float p = (j0 + jb + u < g.Lk) ? exp2f(dot * c2 - lse2) : 0.f;
        sa_axpy<DPT>(dq, p * (dp - delta) * g.scale, sK + (jb + u) * D + sub * DPT);
      }
    }
  }
  if (live) sa_store_row<DPT>(g.dq + qrow * g.lddq + hc, dq, 1.0f);
}

// ------------------------------------------------------------------------------------------------- backward: dK and dV
template <int DPT>
__global__ void __launch_bounds__(128) small_attn_bwd_kv_kernel(const SmallAttnArgs g) {
  constexpr int D = DPT * 4;
  extern __shared__ __align__(16) uint8_t sa_smem[];
  bf16* sQ = reinterpret_cast<bf16*>(sa_smem);
  bf16* sD = sQ + kSaTile * D;
  float* sL = reinterpret_cast<float*>(sD + kSaTile * D);  // [32] lse * log2e (+inf for dead rows), then [32] delta
  pdl_grid_sync();
  const int bh = blockIdx.y, b = bh / g.H, h = bh % g.H;
  const int row = blockIdx.x * kSaRows + (threadIdx.x >> 2);  // key row
  const int sub = threadIdx.x & 3;
  const bool live = row < g.Lk;
  const int hc = h * D + sub * DPT;
  const long long krow = static_cast<long long>(b) * g.Lk + row;
  float k[DPT], v[DPT], dk[DPT], dv[DPT];
  sa_load_row<DPT>(g.k + krow * g.ldk + hc, live, k);
  sa_load_row<DPT>(g.v + krow * g.ldv + hc, live, v);
#pragma unroll
  for (int i = 0; i < DPT; ++i) dk[i] = dv[i] = 0.f;
  const float c2 = g.scale * kSaLog2e;
  const long long q_base = static_cast<long long>(b) * g.L;
  for (int i0 = 0; i0 < g.L; i0 += kSaTile) {
    __syncthreads();
    sa_stage(sQ, g.q + h * D, g.ldq, q_base, i0, g.L, D);
    sa_stage(sD, g.dO + h * D, g.lddo, q_base, i0, g.L, D);
    if (threadIdx.x < kSaTile) {
      const bool ok = i0 + threadIdx.x < g.L;
      sL[threadIdx.x] = ok ? g.lse[static_cast<long long>(bh) * g.L + i0 + threadIdx.x] * kSaLog2e : INFINITY;
      sL[kSaTile + threadIdx.x] = ok ? g.delta[static_cast<long long>(bh) * g.L + i0 + threadIdx.x] : 0.f;
    }
    __syncthreads();
#pragma unroll 2
    for (int ib = 0; ib < kSaTile; ib += 2) {
      float pd[2], pp[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        pd[u] = sa_dot<DPT>(k, sQ + (ib + u) * D + sub * DPT);
        pp[u] = sa_dot<DPT>(v, sD + (ib + u) * D + sub * DPT);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int i = ib + u;
        const float dot = sa_quad_sum(pd[u]), dp = sa_quad_sum(pp[u]);
        const float p = exp2f(dot * c2 - sL[i]);  // dead query rows: exp2(-inf) = 0
        sa_axpy<DPT>(dv, p, sD + i * D + sub * DPT);
        sa_axpy<DPT>(dk, p * (dp - sL[kSaTile + i]) * g.scale, sQ + i * D + sub * DPT);
      }
    }
  }
  if (live) {
    sa_store_row<DPT>(g.dk + krow * g.lddk + hc, dk, 1.0f);
    sa_store_row<DPT>(g.dv + krow * g.lddv + hc, dv, 1.0f);
  }
}

template <int DPT>
static int sa_launch_fwd(const SmallAttnArgs& a, cudaStream_t st) {
  constexpr int smem = 2 * kSaTile * DPT * 4 * 2;
  auto kern = small_attn_fwd_kernel<DPT>;
  static bool configured = false;
  if (!configured) {
    B200_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  dim3 grid((a.L + kSaRows - 1) / kSaRows, a.B * a.H);
  B200_KLAUNCH(kern, grid, 128, smem, st, a);
  B200_CUDA_CHECK(cudaGetLastError());
  return B200_OK;
}
template <int DPT>
static int sa_launch_bwd(const SmallAttnArgs& a, cudaStream_t st) {
  constexpr int smem = 2 * kSaTile * DPT * 4 * 2 + 2 * kSaTile * 4;
  auto kq = small_attn_bwd_q_kernel<DPT>;
  auto kkv = small_attn_bwd_kv_kernel<DPT>;
  static bool configured = false;
  if (!configured) {
    B200_CUDA_CHECK(cudaFuncSetAttribute(kq, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    B200_CUDA_CHECK(cudaFuncSetAttribute(kkv, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  dim3 gq((a.L + kSaRows - 1) / kSaRows, a.B * a.H), gkv((a.Lk + kSaRows - 1) / kSaRows, a.B * a.H);
  B200_KLAUNCH(kq, gq, 128, smem, st, a);  // also writes delta, which the dK/dV kernel reads
  B200_CUDA_CHECK(cudaGetLastError());
  B200_KLAUNCH(kkv, gkv, 128, smem, st, a);
  B200_CUDA_CHECK(cudaGetLastError());
  return B200_OK;
}

static int sa_check(b200_ctx* ctx, int B, int H, int L, int Lk, int d) {
  B200_REQUIRE(B > 0 && H > 0 && L > 0 && Lk > 0, "b200_attn_small: bad shape B=%d H=%d L=%d Lk=%d", B, H, L, Lk);
  B200_REQUIRE(d == 160 || d == 192 || d == 256,
               "b200_attn_small: head dim %d (this kernel serves 160 / 192 / 256; <= 128 belongs to b200_attn_fwd_xd)", d);
  return B200_OK;
}

}  // namespace b200

using namespace b200;

extern "C" int b200_attn_small_fwd(b200_ctx* ctx, const void* q, int ldq, const void* k, int ldk, const void* v, int ldv,
                                   void* out, int ldo, void* lse, int B, int H, int L, int Lk, int head_dim, float scale,
                                   void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  if ((rc = sa_check(ctx, B, H, L, Lk, head_dim))) return rc;
  B200_REQUIRE(q && k && v && out && lse && ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0,
               "b200_attn_small_fwd: null operand or leading dim not a multiple of 8");
  SmallAttnArgs a = {};
  a.q = (const bf16*)q; a.k = (const bf16*)k; a.v = (const bf16*)v;
  a.ldq = ldq; a.ldk = ldk; a.ldv = ldv;
  a.out = (bf16*)out; a.ldout = ldo;
  a.lse = (float*)lse;
  a.B = B; a.H = H; a.L = L; a.Lk = Lk; a.scale = scale;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  rc = head_dim == 160 ? sa_launch_fwd<40>(a, st) : head_dim == 192 ? sa_launch_fwd<48>(a, st) : sa_launch_fwd<64>(a, st);
  if (rc) return rc;
  ctx->launches.fetch_add(1);
  return B200_OK;
}

extern "C" int b200_attn_small_bwd(b200_ctx* ctx, const void* q, int ldq, const void* k, int ldk, const void* v, int ldv,
                                   const void* o, int ldo, const void* dO, int lddo, const void* lse, void* delta, void* dq,
                                   int lddq, void* dk, int lddk, void* dv, int lddv, int B, int H, int L, int Lk, int head_dim,
                                   float scale, void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  if ((rc = sa_check(ctx, B, H, L, Lk, head_dim))) return rc;
  B200_REQUIRE(q && k && v && o && dO && lse && delta && dq && dk && dv, "b200_attn_small_bwd: null operand");
  B200_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0 && lddo % 8 == 0 && lddq % 8 == 0 && lddk % 8 == 0 &&
                   lddv % 8 == 0, "b200_attn_small_bwd: leading dims must be multiples of 8");
  SmallAttnArgs a = {};
  a.q = (const bf16*)q; a.k = (const bf16*)k; a.v = (const bf16*)v; a.o = (const bf16*)o; a.dO = (const bf16*)dO;
  a.ldq = ldq; a.ldk = ldk; a.ldv = ldv; a.ldo = ldo; a.lddo = lddo;
  a.dq = (bf16*)dq; a.dk = (bf16*)dk; a.dv = (bf16*)dv;
  a.lddq = lddq; a.lddk = lddk; a.lddv = lddv;
  a.lse = (float*)lse; a.delta = (float*)delta;
  a.B = B; a.H = H; a.L = L; a.Lk = Lk; a.scale = scale;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  rc = head_dim == 160 ? sa_launch_bwd<40>(a, st) : head_dim == 192 ? sa_launch_bwd<48>(a, st) : sa_launch_bwd<64>(a, st);
  if (rc) return rc;
  ctx->launches.fetch_add(2);
  return B200_OK;
}
