// Rank-side GEMMs of the LoRA adapters:  out[M, 64] = bf16(alpha * A[M, K] . op(B)[64, K]^T)
//   Z = X A_pack^T   (forward, B K-major)     T = dY B_pack   (backward, B stored [K, 64] = MN-major)
// M x 64 outputs give only M/128 output tiles (36 for the 4608-token stream), far fewer than 148 SMs, and each
// CTA would have to stream its whole [128, K] operand slab alone.  So the contraction is split over a thread-block
// CLUSTER: S CTAs each run the TMA -> tcgen05 pipeline over K/S, park their fp32 partial tile in shared memory,
// and the partials are reduced through distributed shared memory (ld.shared::cluster) — no fp32 round trip
// through HBM, no second kernel, bf16 result written once.
#include <cstdlib>

#include "common.cuh"
#include "ctx.h"

namespace b200 {

struct SkinnyArgs {
  int M, n_store, kb_total, S;
  bf16* out;
  int ldo;
  float alpha;
  const float* row_alpha;
  int rows_per_sample;
};

constexpr int kSkStages = 6;
constexpr uint32_t kSkA = 128 * 64 * 2, kSkB = 64 * 64 * 2;
constexpr int kSkSmem = 1024 + kSkStages * (kSkA + kSkB) + 128 * 64 * 4 + 16 * 8 + 16;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float ld_dsmem_f32(uint32_t local_addr, uint32_t rank) {
  uint32_t remote;
  asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_addr), "r"(rank));
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(remote));
  return v;
}

__device__ __forceinline__ void st_dsmem_f32(uint32_t local_addr, uint32_t rank, float v) {
  uint32_t remote;
  asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_addr), "r"(rank));
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote), "f"(v) : "memory");
}

// PUSH = 0: validated round-1 reduction (every CTA parks its whole partial tile locally, the owner of a row block PULLS the
//   S partials with ld.shared::cluster — 64 dependent remote loads per thread, latency-bound: the cluster kernel measured
//   18.5 us against 16 us for the persistent kernel at K = 3072 although it uses 4x the SMs).
// PUSH = 1 (round-2 candidate, B200_SKINNY_PUSH=1): every CTA PUSHES each row block of its partial tile straight into the
//   owner's receive buffer with st.shared::cluster (fire and forget), one cluster barrier, then the owner sums S local
//   slices with ordinary LDS.
template <int B_MN, int PUSH = 0>
__global__ void __launch_bounds__(192, 1)
skinny_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const SkinnyArgs g) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_base_1024(smem_raw);
  float* partial = reinterpret_cast<float*>(smem + kSkStages * (kSkA + kSkB));  // [64 cols][128 rows]
  uint64_t* full = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(partial) + 128 * 64 * 4);
  uint64_t* empty = full + kSkStages;
  uint64_t* done = empty + kSkStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  const int warp = warp_id_uniform(), lane = threadIdx.x & 31;
  const uint32_t rank = __shfl_sync(0xffffffffu, cluster_ctarank(), 0);
  const int m_blk = blockIdx.x / g.S;
  const int kb_per = (g.kb_total + g.S - 1) / g.S;
  const int kb0 = static_cast<int>(rank) * kb_per;
  const int kb1 = min(g.kb_total, kb0 + kb_per);
  const int nkb = max(0, kb1 - kb0);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kSkStages; ++s) {
        mbar_init(&full[s], 1);
        mbar_init(&empty[s], 1);
      }
      mbar_init(done, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<64>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();  // everything above touched only shared / tensor memory
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < nkb; ++i) {
        const int st = i % kSkStages;
        const uint32_t ph = (i / kSkStages) & 1;
        mbar_wait(&empty[st], ph ^ 1u, 30);
        uint8_t* sa = smem + st * (kSkA + kSkB);
        mbar_arrive_expect_tx(&full[st], kSkA + kSkB);
        const int kc = (kb0 + i) * 64;
        tma_load_2d(sa, &tmA, &full[st], kc, m_blk * 128);
        if (B_MN)
          tma_load_2d(sa + kSkA, &tmB, &full[st], 0, kc);  // [64 k rows][64 n] as stored
        else
          tma_load_2d(sa + kSkA, &tmB, &full[st], kc, 0);  // [64 n rows][64 k]
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, B_MN);
    const uint64_t dA = umma_desc_sw128(smem_u32(smem), 1024, 16);
    const uint64_t dB = umma_desc_sw128(smem_u32(smem + kSkA), 1024, B_MN ? 8192 : 16);
    for (int i = 0; i < nkb; ++i) {
      const int st = i % kSkStages;
      const uint32_t ph = (i / kSkStages) & 1;
      mbar_wait(&full[st], ph, 31);
      tc_fence_after();
      const uint64_t off = static_cast<uint64_t>(st) * ((kSkA + kSkB) >> 4);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16_ss_w(tmem_base, dA + off + 2u * k, dB + off + (B_MN ? 128u : 2u) * k, idesc, (i > 0 || k > 0) ? 1u : 0u);
      umma_commit_w(&empty[st]);
    }
    umma_commit_w(done);
  } else {
    // ---- epilogue warps: TMEM -> transposed fp32 partial tile in shared memory
    const int q = warp & 3;
    const int row = q * 32 + lane;
    mbar_wait(done, 0, 32);
    tc_fence_after();
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      if (nkb > 0) {
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c * 32, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0u;
      }
      if (PUSH) {
        // receive buffer of the owner: [source rank][64 cols][rows_per]; lanes = consecutive rows -> contiguous remote stores
        const int rows_per = 128 / g.S;
        const uint32_t owner = static_cast<uint32_t>(row / rows_per);
        const uint32_t base = smem_u32(partial) + ((rank * 64u + c * 32u) * rows_per + (row % rows_per)) * 4u;
#pragma unroll
        for (int i = 0; i < 32; ++i) st_dsmem_f32(base + static_cast<uint32_t>(i * rows_per) * 4u, owner, __uint_as_float(v[i]));
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) partial[(c * 32 + i) * 128 + row] = __uint_as_float(v[i]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_barrier();  // every CTA's partial tile is complete and visible cluster-wide
  if (PUSH) {
    if (warp >= 2) {  // sum the S slices that arrived in MY receive buffer: rows [rank * rows_per, +rows_per)
      const int rows_per = 128 / g.S;
      const int t = (warp - 2) * 32 + lane;
      const int rl = t % rows_per;
      const int cgrp = t / rows_per;
      const int ncg = 128 / rows_per;
      const int cols_per = 64 / ncg;
      const int grow = m_blk * 128 + static_cast<int>(rank) * rows_per + rl;
      if (grow < g.M) {
        float a = g.alpha;
        if (g.row_alpha) a *= g.row_alpha[grow / g.rows_per_sample];
        for (int c = 0; c < cols_per; c += 2) {
          const int col = cgrp * cols_per + c;
          float s0 = 0.f, s1 = 0.f;
#pragma unroll
          for (int p = 0; p < 8; ++p) {
            if (p < g.S) {
              s0 += partial[(p * 64 + col) * rows_per + rl];
              s1 += partial[(p * 64 + col + 1) * rows_per + rl];
            }
          }
          if (col < g.n_store)
            *reinterpret_cast<uint32_t*>(g.out + static_cast<size_t>(grow) * g.ldo + col) = pack_bf16x2(s0 * a, s1 * a);
        }
      }
    }
  } else if (warp >= 2) {
    // CTA `rank` reduces rows [rank * 128 / S, ...) over all S partial tiles; 128 threads: lane -> row, warp -> 16 columns
    const int rows_per = 128 / g.S;  // S in {1, 2, 4, 8}
    const int t = (warp - 2) * 32 + lane;
    const int rl = t % rows_per;
    const int cgrp = t / rows_per;             // 0 .. 128/rows_per - 1
    const int ncg = 128 / rows_per;            // column groups
    const int cols_per = 64 / ncg;             // columns per thread (>= 2 for S <= 8... S=8: rows 16, ncg 8, cols 8)
    const int r_in = static_cast<int>(rank) * rows_per + rl;
    const int grow = m_blk * 128 + r_in;
    if (grow < g.M) {
      float a = g.alpha;
      if (g.row_alpha) a *= g.row_alpha[grow / g.rows_per_sample];
      const uint32_t base = smem_u32(partial);
      // issue every remote load of a column pair before the first add (DSMEM latency ~200 cycles: keep 2 S in flight)
      for (int c = 0; c < cols_per; c += 2) {
        const int col = cgrp * cols_per + c;
        float v0[8], v1[8];
#pragma unroll
        for (int p = 0; p < 8; ++p) {
          if (p < g.S) {
            v0[p] = ld_dsmem_f32(base + ((col)*128 + r_in) * 4, p);
            v1[p] = ld_dsmem_f32(base + ((col + 1) * 128 + r_in) * 4, p);
          }
        }
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int p = 0; p < 8; ++p) {
          if (p < g.S) {
            s0 += v0[p];
            s1 += v1[p];
          }
        }
        if (col < g.n_store)
          *reinterpret_cast<uint32_t*>(g.out + static_cast<size_t>(grow) * g.ldo + col) = pack_bf16x2(s0 * a, s1 * a);
      }
    }
  }
  cluster_barrier();  // nobody leaves while a peer may still read its shared memory
  if (warp == 1) tmem_dealloc<64>(tmem_base);
}

template <int B_MN, int PUSH = 0>
static int launch_skinny(b200_ctx* ctx, const CUtensorMap& ta, const CUtensorMap& tb, const SkinnyArgs& a, cudaStream_t stream) {
  auto kern = skinny_gemm_kernel<B_MN, PUSH>;
  static bool configured = false;
  if (!configured) {
    B200_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSkSmem));
    configured = true;
  }
  const int m_tiles = (a.M + 127) / 128;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(m_tiles * a.S, 1, 1);
  cfg.blockDim = dim3(192, 1, 1);
  cfg.dynamicSmemBytes = kSkSmem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = a.S;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
#if B200_PDL
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.numAttrs = 2;
#endif
  B200_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, ta, tb, a));
  ctx->launches.fetch_add(1);
  return B200_OK;
}

// called from b200_gemm_bf16 (gemm_tcgen05.cu) for N <= 64, bf16 output, no fused epilogue
int skinny_gemm_dispatch(b200_ctx* ctx, const b200_gemm_desc* d, cudaStream_t stream) {
  SkinnyArgs a;
  a.M = d->M;
  a.n_store = d->N;
  a.kb_total = (d->K0 + 63) / 64;
  a.out = reinterpret_cast<bf16*>(d->out);
  a.ldo = d->ldo;
  a.alpha = d->alpha;
  a.row_alpha = reinterpret_cast<const float*>(d->row_alpha);
  a.rows_per_sample = d->rows_per_sample > 0 ? d->rows_per_sample : 1;
  const int m_tiles = (d->M + 127) / 128;
  int S = 1;
  while (S < 8 && m_tiles * S * 2 <= ctx->sm_count && a.kb_total >= 4 * S) S *= 2;
  a.S = S;
  CUtensorMap ta, tb;
  int rc = make_tmap_bf16_2d(ctx, &ta, d->A0, d->M, d->K0, d->lda0, 64, 128);
  if (rc) return rc;
  if (d->trans_b)
    rc = make_tmap_bf16_2d(ctx, &tb, d->B0, d->K0, d->N, d->ldb0, 64, 64);
  else
    rc = make_tmap_bf16_2d(ctx, &tb, d->B0, d->N, d->K0, d->ldb0, 64, 64);
  if (rc) return rc;
  static int push = -1;  // B200_SKINNY_PUSH=1: round-2 candidate reduction (opt-in)
  if (push < 0) {
    const char* e = getenv("B200_SKINNY_PUSH");
    push = (e && atoi(e) == 1) ? 1 : 0;
  }
  if (push) return d->trans_b ? launch_skinny<1, 1>(ctx, ta, tb, a, stream) : launch_skinny<0, 1>(ctx, ta, tb, a, stream);
  return d->trans_b ? launch_skinny<1>(ctx, ta, tb, a, stream) : launch_skinny<0>(ctx, ta, tb, a, stream);
}

}  // namespace b200
