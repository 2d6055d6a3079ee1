// LoRA-wrapped Linear for a handful of rows (M <= 8): the AdaLN modulation projections
// `norm1.linear`, `norm1_context.linear`, `norm.linear` act on the [B, D] conditioning vector, so the
// work is pure weight streaming (3072 x 18432 bf16 = 113 MB per call) and belongs on the HBM roofline,
// not on a 128-row tensor-core tile.  One warp per output feature, 16-byte loads, fp32 accumulate; the
// rank-r side uses the fp32 master weights directly (no bf16 copies).
//
// Replaces toolkit/network_mixins.py:304-342 for these modules:
//   y = bf16( bf16(x W^T + bias) + bf16( c * (x A^T) B^T ) ),   c = multiplier * scale
// and its backward w.r.t. A and B (the conditioning vector itself has no trainable ancestor).
#include "common.cuh"
#include "ctx.h"

namespace b200 {

constexpr int kMaxRows = 8;
constexpr int kMaxRank = 128;

// z[b, j] = c * sum_k x[b,k] A[j,k]          one warp per (b, j)
__global__ void __launch_bounds__(256) lora_gemv_down_kernel(const bf16* __restrict__ x, int ldx, const float* __restrict__ A,
                                                             float c, const float* __restrict__ row_c,
                                                             float* __restrict__ z, int Bm, int r, int K) {
  pdl_grid_sync();
  const int w = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (w >= Bm * r) return;
  const int b = w / r, j = w % r;
  const bf16* xr = x + static_cast<size_t>(b) * ldx;
  const float* ar = A + static_cast<size_t>(j) * K;
  float acc = 0.f;
  for (int k = lane; k < K; k += 32) acc += __bfloat162float(xr[k]) * ar[k];
  acc = warp_sum(acc);
  if (lane == 0) z[b * r + j] = c * (row_c ? row_c[b] : 1.0f) * acc;  // per-sample multiplier (network_mixins.py:311-322)
}

// y[b, n] for all b; one warp per n.  x is staged in shared memory once per block.
template <int BM>
__global__ void __launch_bounds__(256) lora_gemv_fwd_kernel(const bf16* __restrict__ x, int ldx, const bf16* __restrict__ W,
                                                            int ldw, const bf16* __restrict__ bias, const float* __restrict__ z,
                                                            const float* __restrict__ Bw, int r, bf16* __restrict__ y, int ldy,
                                                            int N, int K) {
  pdl_grid_sync();
  extern __shared__ uint8_t smem_raw[];
  bf16* xs = reinterpret_cast<bf16*>(smem_raw);  // [BM][K]
  for (int i = threadIdx.x * 8; i < BM * K; i += blockDim.x * 8) {
    const int b = i / K, k = i % K;
    *reinterpret_cast<uint4*>(xs + i) = *reinterpret_cast<const uint4*>(x + static_cast<size_t>(b) * ldx + k);
  }
  __syncthreads();
  const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  const bf16* wr = W + static_cast<size_t>(n) * ldw;
  float acc[BM];
#pragma unroll
  for (int b = 0; b < BM; ++b) acc[b] = 0.f;
  for (int k = lane * 8; k < K; k += 256) {
    uint4 u = *reinterpret_cast<const uint4*>(wr + k);
    float w[8];
    float2 t0 = unpack_bf16x2(u.x), t1 = unpack_bf16x2(u.y), t2 = unpack_bf16x2(u.z), t3 = unpack_bf16x2(u.w);
    w[0] = t0.x; w[1] = t0.y; w[2] = t1.x; w[3] = t1.y; w[4] = t2.x; w[5] = t2.y; w[6] = t3.x; w[7] = t3.y;
#pragma unroll
    for (int b = 0; b < BM; ++b) {
      uint4 xu = *reinterpret_cast<const uint4*>(xs + b * K + k);
      float2 x0 = unpack_bf16x2(xu.x), x1 = unpack_bf16x2(xu.y), x2 = unpack_bf16x2(xu.z), x3 = unpack_bf16x2(xu.w);
      acc[b] += w[0] * x0.x + w[1] * x0.y + w[2] * x1.x + w[3] * x1.y + w[4] * x2.x + w[5] * x2.y + w[6] * x3.x +
                w[7] * x3.y;
    }
  }
#pragma unroll
  for (int b = 0; b < BM; ++b) acc[b] = warp_sum(acc[b]);
  if (lane == 0) {
    const float bv = bias ? __bfloat162float(bias[n]) : 0.f;
#pragma unroll
    for (int b = 0; b < BM; ++b) {
      float v = bf16_round(acc[b] + bv);
      if (r > 0) {
        float l = 0.f;
        for (int j = 0; j < r; ++j) l += z[b * r + j] * Bw[static_cast<size_t>(n) * r + j];
        v = bf16_round(v + bf16_round(l));
      }
      y[static_cast<size_t>(b) * ldy + n] = __float2bfloat16_rn(v);
    }
  }
}

// blocks [0, nb_db): dBw[n, j] += sum_b dy[b,n] z[b,j]                 (thread per (n, j))
// blocks [nb_db, ..): t[b, j]   += c * sum_{n in 256-row slab} dy[b,n] Bw[n,j]   (thread (sub, j): j fastest, so the
//                     reads of Bw rows are coalesced; fp32 atomics into t, which the caller zeroed)
__global__ void __launch_bounds__(256) lora_gemv_bwd1_kernel(const float* __restrict__ dy, int lddy, const float* __restrict__ z,
                                                             const float* __restrict__ Bw, float c,
                                                             const float* __restrict__ row_c, float* __restrict__ dBw,
                                                             float* __restrict__ t, int Bm, int r, int N, int nb_db) {
  pdl_grid_sync();
  if (static_cast<int>(blockIdx.x) < nb_db) {
    const long long idx = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
    if (idx >= static_cast<long long>(N) * r) return;
    const int n = static_cast<int>(idx / r), j = static_cast<int>(idx % r);
    float acc = 0.f;
    for (int b = 0; b < Bm; ++b) acc += dy[static_cast<size_t>(b) * lddy + n] * z[b * r + j];
    dBw[idx] += acc;
  } else {
    const int groups = 256 / r;
    const int j = threadIdx.x % r, sub = threadIdx.x / r;
    if (sub >= groups) return;
    const int n0 = (blockIdx.x - nb_db) * 256;
    const int n1 = min(n0 + 256, N);
    float acc[kMaxRows];
#pragma unroll
    for (int b = 0; b < kMaxRows; ++b) acc[b] = 0.f;
    for (int n = n0 + sub; n < n1; n += groups) {
      const float w = Bw[static_cast<size_t>(n) * r + j];
#pragma unroll
      for (int b = 0; b < kMaxRows; ++b)
        if (b < Bm) acc[b] += dy[static_cast<size_t>(b) * lddy + n] * w;
    }
#pragma unroll
    for (int b = 0; b < kMaxRows; ++b)
      if (b < Bm) atomicAdd(t + b * r + j, c * (row_c ? row_c[b] : 1.0f) * acc[b]);
  }
}

// dA[j, k] += sum_b t[b,j] x[b,k]
__global__ void __launch_bounds__(256) lora_gemv_bwd2_kernel(const float* __restrict__ t, const bf16* __restrict__ x, int ldx,
                                                             float* __restrict__ dA, int Bm, int r, int K) {
  pdl_grid_sync();
  const long long idx = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (idx >= static_cast<long long>(r) * K) return;
  const int j = static_cast<int>(idx / K), k = static_cast<int>(idx % K);
  float acc = 0.f;
  for (int b = 0; b < Bm; ++b) acc += t[b * r + j] * __bfloat162float(x[static_cast<size_t>(b) * ldx + k]);
  dA[idx] += acc;
}

}  // namespace b200

using namespace b200;

extern "C" int b200_lora_gemv_fwd(b200_ctx* ctx, const void* x, int ldx, const void* W, int ldw, const void* bias,
                                  const void* A, const void* Bw, int r, float c, void* y, int ldy, void* z, int Bm, int N,
                                  int K, void* stream) {
  return b200_lora_gemv_fwd_rows(ctx, x, ldx, W, ldw, bias, A, Bw, r, c, nullptr, y, ldy, z, Bm, N, K, stream);
}

extern "C" int b200_lora_gemv_fwd_rows(b200_ctx* ctx, const void* x, int ldx, const void* W, int ldw, const void* bias,
                                       const void* A, const void* Bw, int r, float c, const void* row_c, void* y, int ldy,
                                       void* z, int Bm, int N, int K, void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  B200_REQUIRE(x && W && y && Bm >= 1 && Bm <= kMaxRows && N > 0 && K > 0, "b200_lora_gemv_fwd: bad args Bm=%d N=%d K=%d", Bm,
               N, K);
  B200_REQUIRE(K % 8 == 0 && ldx % 8 == 0 && ldw % 8 == 0, "b200_lora_gemv_fwd: K, ldx, ldw must be multiples of 8");
  B200_REQUIRE(r >= 0 && r <= kMaxRank, "b200_lora_gemv_fwd: rank %d", r);
  if (r > 0) B200_REQUIRE(A && Bw && z, "b200_lora_gemv_fwd: rank > 0 needs A, B and z");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (r > 0) {
    B200_KLAUNCH(lora_gemv_down_kernel, (Bm * r + 7) / 8, 256, 0, st, (const bf16*)x, ldx, (const float*)A, c, (const float*)row_c, (float*)z, Bm, r, K);
    B200_CUDA_CHECK(cudaGetLastError());
    ctx->launches.fetch_add(1);
  }
  const size_t smem = static_cast<size_t>(Bm) * K * sizeof(bf16);
  B200_REQUIRE(smem <= 200 * 1024, "b200_lora_gemv_fwd: %d x %d input does not fit shared memory", Bm, K);
  const int grid = (N + 7) / 8;
#define GEMV_CASE(BM_)                                                                                                 \
  case BM_: {                                                                                                          \
    auto kern = lora_gemv_fwd_kernel<BM_>;                                                                             \
    if (smem > 48 * 1024) B200_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    B200_KLAUNCH(kern, grid, 256, smem, st, (const bf16*)x, ldx, (const bf16*)W, ldw, (const bf16*)bias, (const float*)z,       \
                                  (const float*)Bw, r, (bf16*)y, ldy, N, K);                                           \
  } break;
  switch (Bm) {
    GEMV_CASE(1) GEMV_CASE(2) GEMV_CASE(3) GEMV_CASE(4) GEMV_CASE(5) GEMV_CASE(6) GEMV_CASE(7) GEMV_CASE(8)
  }
#undef GEMV_CASE
  B200_CUDA_CHECK(cudaGetLastError());
  ctx->launches.fetch_add(1);
  return B200_OK;
}

extern "C" int b200_lora_gemv_bwd(b200_ctx* ctx, const void* dy, int lddy, const void* x, int ldx, const void* z,
                                  const void* A, const void* Bw, int r, float c, void* dA, void* dBw, void* t_ws, int Bm,
                                  int N, int K, void* stream) {
  return b200_lora_gemv_bwd_rows(ctx, dy, lddy, x, ldx, z, A, Bw, r, c, nullptr, dA, dBw, t_ws, Bm, N, K, stream);
}

extern "C" int b200_lora_gemv_bwd_rows(b200_ctx* ctx, const void* dy, int lddy, const void* x, int ldx, const void* z,
                                       const void* A, const void* Bw, int r, float c, const void* row_c, void* dA, void* dBw,
                                       void* t_ws, int Bm, int N, int K, void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  B200_REQUIRE(dy && x && z && A && Bw && dA && dBw && t_ws, "b200_lora_gemv_bwd: null argument");
  B200_REQUIRE(Bm >= 1 && Bm <= kMaxRows && r >= 1 && r <= kMaxRank && N > 0 && K > 0, "b200_lora_gemv_bwd: bad shape");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int nb_db = static_cast<int>((static_cast<long long>(N) * r + 255) / 256);
  const int nb_t = (N + 255) / 256;
  B200_CUDA_CHECK(cudaMemsetAsync(t_ws, 0, sizeof(float) * Bm * r, st));
  B200_KLAUNCH(lora_gemv_bwd1_kernel, nb_db + nb_t, 256, 0, st, (const float*)dy, lddy, (const float*)z, (const float*)Bw, c,
                                                      (const float*)row_c, (float*)dBw, (float*)t_ws, Bm, r, N, nb_db);
  B200_CUDA_CHECK(cudaGetLastError());
  const int nb_da = static_cast<int>((static_cast<long long>(r) * K + 255) / 256);
  B200_KLAUNCH(lora_gemv_bwd2_kernel, nb_da, 256, 0, st, (const float*)t_ws, (const bf16*)x, ldx, (float*)dA, Bm, r, K);
  B200_CUDA_CHECK(cudaGetLastError());
  ctx->launches.fetch_add(2);
  return B200_OK;
}
