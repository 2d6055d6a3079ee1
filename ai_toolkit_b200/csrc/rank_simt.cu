// Rank-side products of ONE adapter with a small live rank (r <= 16), on the CUDA cores:
//   out[M, 64] = bf16(alpha * row_alpha[s] * X[M, K] . op(W)),   columns >= r written as zeros
//   TRANS = 0:  W = A_pack [64, K], rows    >= r ignored        Z = X A^T      (forward,  toolkit/network_mixins.py:304-342)
//   TRANS = 1:  W = B_pack [K, 64], columns >= r ignored        T = dY B       (backward of lora_up)
// Why not the tensor-core skinny GEMM (skinny_gemm.cu): these launches are pure streaming of X (28-141 MB for FLUX, 5-40 MB for
// SDXL) with 2 r FLOP per element, and the tcgen05 pipeline's fixed cost (TMEM allocation, tensor-map fetch, cluster barriers
// and the DSMEM reduction of the split-K partials) made them 16-27 us each on FLUX and 12.5 us on SDXL's 2048-token levels
// (profiles/r2_step_launches_warm.md, r2_sdxl_profile.md) -- several times their HBM time.  They also pad the rank to 64
// columns of MMA work.  Here a warp owns R rows and the whole contraction: lane l holds k = 256 c + 8 l .. + 7 of every chunk c,
// 16-byte loads of X straight from global (next chunk prefetched), the r x 256 weight slice of the chunk staged once per block
// in shared memory (double-buffered, lane-major so the reads are conflict-free), packed fp32 FMAs on rank pairs, and one
// shuffle reduce-scatter of the R x r accumulators at the end.  Used for r <= 16 (groups of adapters with more live columns
// and larger ranks stay on the tensor-core kernel).
#include "attn_common.cuh"
#include "common.cuh"
#include "ctx.h"

namespace b200 {

template <int N, int W>
__device__ __forceinline__ void reduce_scatter_step(float (&v)[64], int lane) {  // N live values -> N / 2
  const bool upper = (lane & W) != 0;
#pragma unroll
  for (int i = 0; i < N / 2; ++i) {
    const float send = upper ? v[i] : v[i + N / 2];
    const float keep = upper ? v[i + N / 2] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, W);
  }
}

// RK = live rank padded to 4 / 8 / 16; R rows per warp; NT threads per block
template <int RK, int TRANS>
__global__ void __launch_bounds__(128) rank_simt_kernel(const bf16* __restrict__ X, int ldx, const bf16* __restrict__ Wt, int ldw,
                                                        bf16* __restrict__ out, int ldo, int M, int K, float alpha,
                                                        const float* __restrict__ row_alpha, int rows_per_sample) {
  constexpr int R = RK == 4 ? 8 : 4;
  constexpr int NV = R * RK;       // accumulators per lane (64 or 32)
  constexpr int DEPTH = R == 4 ? 3 : 2;  // chunks of X in flight per lane (v1 had 1: 1.4 TB/s, latency-bound)
  constexpr int P = RK / 8 ? RK / 8 : 1;  // 16-byte parts of the live columns of one B_pack row (TRANS = 1; RK = 4: half a part)
  __shared__ __align__(16) uint4 wsm[2][RK * 32];  // chunk slice, lane-major: TRANS 0: [j][lane]; TRANS 1: [e][p][lane] (RK >= 8)
  pdl_grid_sync();
  const int nwarp = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = (blockIdx.x * nwarp + warp) * R;
  const int nchunk = (K + 255) / 256;

  // ---- stage the weight slice of chunk c into buffer b (all threads of the block)
  auto stage = [&](int c, int b) {
    if (TRANS == 0) {
      for (int idx = threadIdx.x; idx < RK * 32; idx += blockDim.x) {
        const int j = idx >> 5, l = idx & 31;
        const int k = c * 256 + l * 8;
        uint4 u = make_uint4(0u, 0u, 0u, 0u);
        if (k < K) u = *reinterpret_cast<const uint4*>(Wt + static_cast<size_t>(j) * ldw + k);
        wsm[b][idx] = u;
      }
    } else if (RK >= 8) {
      for (int idx = threadIdx.x; idx < 256 * P; idx += blockDim.x) {
        const int n = idx / P, p = idx % P;   // row of B_pack inside the chunk, 16-byte part of its live columns
        const int k = c * 256 + n;
        uint4 u = make_uint4(0u, 0u, 0u, 0u);
        if (k < K) u = *reinterpret_cast<const uint4*>(Wt + static_cast<size_t>(k) * ldw + p * 8);
        wsm[b][((n & 7) * P + p) * 32 + (n >> 3)] = u;
      }
    } else {  // RK = 4: 8 bytes per row; two rows (e, e + 1) share one 16-byte slot: [e / 2][lane]
      uint2* w2 = reinterpret_cast<uint2*>(&wsm[b][0]);
      for (int idx = threadIdx.x; idx < 256; idx += blockDim.x) {
        const int k = c * 256 + idx;
        uint2 u = make_uint2(0u, 0u);
        if (k < K) u = *reinterpret_cast<const uint2*>(Wt + static_cast<size_t>(k) * ldw);
        const int e = idx & 7, l = idx >> 3;
        w2[((e >> 1) * 32 + l) * 2 + (e & 1)] = u;
      }
    }
  };
  auto load_x = [&](int c, uint4 (&xr)[R]) {
    const int k = c * 256 + lane * 8;
#pragma unroll
    for (int i = 0; i < R; ++i) {
      const int row = min(row0 + i, M - 1);
      xr[i] = (k < K && row0 < M) ? *reinterpret_cast<const uint4*>(X + static_cast<size_t>(row) * ldx + k) : make_uint4(0u, 0u, 0u, 0u);
    }
  };

  float2 acc[R][RK / 2];
#pragma unroll
  for (int i = 0; i < R; ++i)
#pragma unroll
    for (int j = 0; j < RK / 2; ++j) acc[i][j] = make_float2(0.f, 0.f);
  uint4 xq[DEPTH][R];
  stage(0, 0);
#pragma unroll
  for (int u = 0; u < DEPTH; ++u) load_x(u, xq[u]);  // (chunks past K load nothing and read as zeros)
  __syncthreads();
  for (int c0 = 0; c0 < nchunk; c0 += DEPTH) {
#pragma unroll
    for (int u = 0; u < DEPTH; ++u) {
      const int c = c0 + u;
      if (c < nchunk) {  // block-uniform
        const int b = c & 1;
        float x[R][8];
#pragma unroll
        for (int i = 0; i < R; ++i) {
          const float2 a0 = unpack_bf16x2(xq[u][i].x), a1 = unpack_bf16x2(xq[u][i].y), a2 = unpack_bf16x2(xq[u][i].z),
                       a3 = unpack_bf16x2(xq[u][i].w);
          x[i][0] = a0.x; x[i][1] = a0.y; x[i][2] = a1.x; x[i][3] = a1.y; x[i][4] = a2.x; x[i][5] = a2.y; x[i][6] = a3.x; x[i][7] = a3.y;
        }
        load_x(c + DEPTH, xq[u]);              // refill this slot: DEPTH chunks in flight while this one is computed
        if (c + 1 < nchunk) stage(c + 1, b ^ 1);  // buffer b ^ 1 was last read in iteration c - 1 (barrier at the end of it)
        if (TRANS == 0) {
#pragma unroll
          for (int jp = 0; jp < RK / 2; ++jp) {
            const uint4 w0 = wsm[b][(2 * jp) * 32 + lane], w1 = wsm[b][(2 * jp + 1) * 32 + lane];
            const uint32_t u0[4] = {w0.x, w0.y, w0.z, w0.w}, u1[4] = {w1.x, w1.y, w1.z, w1.w};
#pragma unroll
            for (int eh = 0; eh < 4; ++eh) {
              const float2 a = unpack_bf16x2(u0[eh]), bb = unpack_bf16x2(u1[eh]);  // (A[j][e], A[j][e+1]), (A[j+1][e], A[j+1][e+1])
              const float2 p0 = make_float2(a.x, bb.x), p1 = make_float2(a.y, bb.y);
#pragma unroll
              for (int i = 0; i < R; ++i) {
                acc[i][jp] = ffma2(make_float2(x[i][2 * eh], x[i][2 * eh]), p0, acc[i][jp]);
                acc[i][jp] = ffma2(make_float2(x[i][2 * eh + 1], x[i][2 * eh + 1]), p1, acc[i][jp]);
              }
            }
          }
        } else if (RK >= 8) {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
#pragma unroll
            for (int p = 0; p < P; ++p) {
              const uint4 w = wsm[b][(e * P + p) * 32 + lane];  // B_pack[k_e][8 p .. 8 p + 7]: rank pairs as packed words
              const uint32_t uw[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float2 wp = unpack_bf16x2(uw[q]);
#pragma unroll
                for (int i = 0; i < R; ++i) acc[i][p * 4 + q] = ffma2(make_float2(x[i][e], x[i][e]), wp, acc[i][p * 4 + q]);
              }
            }
          }
        } else {
#pragma unroll
          for (int e2 = 0; e2 < 4; ++e2) {
            const uint4 w = wsm[b][e2 * 32 + lane];  // rows e = 2 e2 (x, y) and 2 e2 + 1 (z, w)
            const uint32_t uw[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
              for (int q = 0; q < 2; ++q) {
                const float2 wp = unpack_bf16x2(uw[h * 2 + q]);
#pragma unroll
                for (int i = 0; i < R; ++i)
                  acc[i][q] = ffma2(make_float2(x[i][2 * e2 + h], x[i][2 * e2 + h]), wp, acc[i][q]);
              }
          }
        }
        __syncthreads();  // buffer b free for chunk c + 2, buffer b ^ 1 complete for chunk c + 1
      }
    }
  }
  if (row0 >= M) return;
  // ---- reduce over the lanes: the (row, rank) accumulators are scattered, lane l ends with NV / 32 of the totals
  float v[64];
#pragma unroll
  for (int i = 0; i < R; ++i)
#pragma unroll
    for (int j = 0; j < RK / 2; ++j) {
      v[i * RK + 2 * j] = acc[i][j].x;
      v[i * RK + 2 * j + 1] = acc[i][j].y;
    }
  if (NV == 64) reduce_scatter_step<64, 16>(v, lane);
  reduce_scatter_step<NV == 64 ? 32 : 32, NV == 64 ? 8 : 16>(v, lane);
  reduce_scatter_step<NV == 64 ? 16 : 16, NV == 64 ? 4 : 8>(v, lane);
  reduce_scatter_step<NV == 64 ? 8 : 8, NV == 64 ? 2 : 4>(v, lane);
  reduce_scatter_step<NV == 64 ? 4 : 4, NV == 64 ? 1 : 2>(v, lane);
  if (NV == 32) reduce_scatter_step<2, 1>(v, lane);
  // original index of v[i]: the lane bits select the kept halves, most significant step first
  int idx;
  if (NV == 64)
    idx = ((lane >> 4) & 1) * 32 + ((lane >> 3) & 1) * 16 + ((lane >> 2) & 1) * 8 + ((lane >> 1) & 1) * 4 + (lane & 1) * 2;
  else
    idx = ((lane >> 4) & 1) * 16 + ((lane >> 3) & 1) * 8 + ((lane >> 2) & 1) * 4 + ((lane >> 1) & 1) * 2 + (lane & 1);
  {
    const int row = row0 + idx / RK, col = idx % RK;
    if (row < M) {
      float a = alpha;
      if (row_alpha) a *= row_alpha[row / rows_per_sample];
      bf16* o = out + static_cast<size_t>(row) * ldo + col;
      if (NV == 64)
        *reinterpret_cast<uint32_t*>(o) = pack_bf16x2(v[0] * a, v[1] * a);
      else
        *o = __float2bfloat16_rn(v[0] * a);
    }
  }
  // ---- zero the dead columns [RK, 64) of the warp's rows (the fused GEMM multiplies them by the zero columns of B_pack)
  constexpr int Z8 = (64 - RK) / 4;  // 8-byte stores per row
  for (int t = lane; t < R * Z8; t += 32) {
    const int row = row0 + t / Z8;
    if (row < M) *reinterpret_cast<uint2*>(out + static_cast<size_t>(row) * ldo + RK + (t % Z8) * 4) = make_uint2(0u, 0u);
  }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_rank_gemm(b200_ctx* ctx, const void* X, int ldx, const void* W, int ldw, int trans_w, void* out, int ldo,
                              int M, int K, int r_live, float alpha, const void* row_alpha, int rows_per_sample, void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  B200_REQUIRE(X && W && out && M > 0 && K > 0, "b200_rank_gemm: bad args M=%d K=%d", M, K);
  B200_REQUIRE(r_live >= 1 && r_live <= 16, "b200_rank_gemm: live rank %d not in 1..16 (larger ranks: b200_gemm_bf16)", r_live);
  B200_REQUIRE(K % 8 == 0 && ldx % 8 == 0 && ldw % 8 == 0 && ldo % 4 == 0, "b200_rank_gemm: K, ldx, ldw must be multiples of 8");
  B200_REQUIRE((reinterpret_cast<uintptr_t>(X) & 15u) == 0 && (reinterpret_cast<uintptr_t>(W) & 15u) == 0 &&
                   (reinterpret_cast<uintptr_t>(out) & 7u) == 0, "b200_rank_gemm: operand alignment");
  if (row_alpha) B200_REQUIRE(rows_per_sample > 0, "b200_rank_gemm: row_alpha needs rows_per_sample");
  const int RK = r_live <= 4 ? 4 : (r_live <= 8 ? 8 : 16);
  const int R = RK == 4 ? 8 : 4;
  const int warps = (M + R - 1) / R;
  // 4 warps per block when that still gives every SM a block, else 2 (small M: more blocks in flight)
  const int wpb = (warps / 4 >= ctx->sm_count) ? 4 : 2;
  const int grid = (warps + wpb - 1) / wpb;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define RS(RK_, TR_)                                                                                                      \
  do {                                                                                                                   \
    auto kern = rank_simt_kernel<RK_, TR_>;                                                                              \
    B200_KLAUNCH(kern, grid, wpb * 32, 0, st, (const bf16*)X, ldx, (const bf16*)W, ldw, (bf16*)out, ldo, M, K, alpha,     \
                 (const float*)row_alpha, rows_per_sample > 0 ? rows_per_sample : 1);                                    \
  } while (0)
  if (trans_w) {
    if (RK == 4) RS(4, 1); else if (RK == 8) RS(8, 1); else RS(16, 1);
  } else {
    if (RK == 4) RS(4, 0); else if (RK == 8) RS(8, 0); else RS(16, 0);
  }
#undef RS
  B200_CUDA_CHECK(cudaGetLastError());
  ctx->launches.fetch_add(1);
  return B200_OK;
}
