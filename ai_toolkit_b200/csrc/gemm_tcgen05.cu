// Persistent, warp-specialised bf16 GEMM for sm_100a: TMA -> shared memory (SWIZZLE_128B) ->
// tcgen05.mma with the fp32 accumulator in TMEM -> tcgen05.ld epilogue with the LoRA / AdaLN fusions.
//
//   acc = A0 . B0^T  (+ A1 . B1^T)      both operands K-major (row-major with K contiguous)
//
// The second contraction segment is the LoRA up-projection: it lands in the SAME TMEM tile as the
// frozen base GEMM, so the adapter costs one extra 64-wide k-block and no extra pass over the output
// (reference: toolkit/network_mixins.py:304-342 runs it as ~8 separate full-size elementwise kernels).
//
// Roles (320 threads):  warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2..9 = epilogue (two per TMEM sub-partition, alternating 32-column chunks).  Two TMEM accumulator stages let the epilogue
// of tile i overlap the main loop of tile i+1.
//
// CG == 2 pairs two CTAs of a cluster on one 256 x BN tile (tcgen05.mma.cta_group::2): each CTA loads
// its own 128 rows of A and HALF of the B tile, which halves the L2->SM traffic per flop.
#include "common.cuh"
#include "ctx.h"

namespace b200 {

struct GemmArgs {
  int M, N;
  int kb0, kb1;  // 64-wide k-blocks in segment 0 / 1
  int splits;    // split-K factor (out_f32 only)
  int group_m;   // rasterisation group (tile rows per group)
  const bf16* bias;
  const bf16* res;
  int ldres;
  const bf16* gate;
  int ldgate;
  int rows_per_sample;
  const bf16* aux_in;
  int ldaux_in;
  bf16* aux_out;
  int ldaux_out;
  void* out;
  int ldo;
  int act;
  int act_ncols;  // activation / aux_out apply to columns n < act_ncols only
  int f32_mode;   // 0 = bf16 epilogue, 1 = fp32 partial store (split-K), 2 = fp32 atomic accumulate
  int f32_trans;  // f32_mode 2: write out[n * ldo + m] instead of out[m * ldo + n]
  int n_store;    // f32 modes: only columns [0, n_store) are written
  float alpha;    // accumulator scale applied before the epilogue
  // dual-problem mode (the two wgrads of one adapter in one launch): tile rows >= dual_mt0 belong to problem 1, which
  // reads its operands through the segment-1 tensor maps and writes out1; both share K (tokens), N = 64, alpha, splits
  int dual_mt0;
  int M0, M1;
  void* out1;
  int ldo1;
  int f32_trans1;
  const float* row_alpha;  // optional per-sample scale [ceil(M / rows_per_sample)]
};

constexpr uint32_t kPeerMask = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address (pair leader)

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_cluster(uint32_t bar_addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(bar_addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_cg2(uint32_t smem_dst, const CUtensorMap* tmap, uint32_t bar_addr, int c0,
                                                int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
      "[%2];" ::"r"(smem_dst),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar_addr), "r"(c0), "r"(c1)
      : "memory");
}
// warp-converged variants (all 32 lanes execute, one elected lane issues; see common.cuh)
__device__ __forceinline__ void umma_bf16_ss_cg2_w(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                   uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_cg2_mc_w(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}" ::
          "r"(smem_u32(bar)),
      "h"(mask)
      : "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc_cg2(uint32_t* smem_result) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "n"(NCOLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc_cg2(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}

// tile index -> (split, m_blk, n_blk) with grouped rasterisation: tiles of `group_m` consecutive tile
// rows are visited column by column so that one wave re-uses both operand panels out of L2.
__device__ __forceinline__ void tile_coords(int t, int m_tiles, int n_tiles, int group_m, int& split, int& m_blk,
                                            int& n_blk) {
  const int per_split = m_tiles * n_tiles;
  split = t / per_split;
  int r = t - split * per_split;
  const int group_size = group_m * n_tiles;
  const int grp = r / group_size;
  const int first_m = grp * group_m;
  const int gm = min(group_m, m_tiles - first_m);
  const int in_grp = r - grp * group_size;
  m_blk = first_m + in_grp % gm;
  n_blk = in_grp / gm;
}

__device__ __forceinline__ void load8_bf16(const bf16* p, float (&v)[8]) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
__device__ __forceinline__ void store8_bf16(bf16* p, const float (&v)[8]) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]);
  u.y = pack_bf16x2(v[2], v[3]);
  u.z = pack_bf16x2(v[4], v[5]);
  u.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}

// Epilogue for 32 consecutive columns of one output row held in registers.
__device__ __forceinline__ void epilogue_row32(const GemmArgs& g, int row, int n0, int split, const uint32_t (&r)[32]) {
  if (g.dual_mt0 > 0) {  // dual wgrad: fp32 atomic accumulate only
    const bool p1 = row >= g.dual_mt0 * 128;
    const int lrow = p1 ? row - g.dual_mt0 * 128 : row;
    if (lrow >= (p1 ? g.M1 : g.M0)) return;
    float* o = reinterpret_cast<float*>(p1 ? g.out1 : g.out);
    const int ldo = p1 ? g.ldo1 : g.ldo;
    const bool tr = p1 ? g.f32_trans1 != 0 : g.f32_trans != 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int n = n0 + j;
      if (n < g.n_store) {
        const size_t idx = tr ? static_cast<size_t>(n) * ldo + lrow : static_cast<size_t>(lrow) * ldo + n;
        atomicAdd(o + idx, g.alpha * __uint_as_float(r[j]));
      }
    }
    return;
  }
  if (row >= g.M) return;
  const int sample = (g.gate != nullptr || g.row_alpha != nullptr) ? row / g.rows_per_sample : 0;
  const float alpha = g.alpha * (g.row_alpha != nullptr ? g.row_alpha[sample] : 1.0f);
  if (g.f32_mode == 1) {
    float* o = reinterpret_cast<float*>(g.out) + (static_cast<size_t>(split) * g.M + row) * g.ldo + n0;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      if (n0 + j < g.n_store) {
        float4 v = make_float4(alpha * __uint_as_float(r[j]), alpha * __uint_as_float(r[j + 1]),
                               alpha * __uint_as_float(r[j + 2]), alpha * __uint_as_float(r[j + 3]));
        *reinterpret_cast<float4*>(o + j) = v;
      }
    }
    return;
  }
  if (g.f32_mode == 2) {
    float* o = reinterpret_cast<float*>(g.out);
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int n = n0 + j;
      if (n < g.n_store) {
        const size_t idx = g.f32_trans ? static_cast<size_t>(n) * g.ldo + row : static_cast<size_t>(row) * g.ldo + n;
        atomicAdd(o + idx, alpha * __uint_as_float(r[j]));
      }
    }
    return;
  }
  // (bf16 epilogue is warp-cooperative: epilogue_warp32_bf16)
}

// ---- warp-cooperative bf16 epilogue ---------------------------------------------------------------------
// tcgen05.ld hands every thread ONE ROW (32 columns).  Reading / writing global memory in that shape makes each
// warp instruction touch 32 different 128-byte lines with 16 bytes each (measured: every extra full-size stream —
// aux_out, aux_in, res — cost ~12 % of the GEMM).  So the [32 rows x 32 cols] bf16 block is transposed through a
// 2 KB per-warp shared-memory scratch (XOR-swizzled 16-byte chunks, conflict-free both ways): global accesses are
// then 4 instructions of 8 rows x 64 contiguous bytes.
__device__ __forceinline__ uint32_t epi_off(int row, int chunk) { return row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4); }

// block [row0 .. row0+31] x [n0 .. n0+31] of a row-major bf16 matrix: issue the 4 coalesced loads (8 rows x 64 B each) ...
__device__ __forceinline__ void epi_prefetch(const bf16* base, int ld, int row0, int n0, int M, int N, int lane, uint4 (&u)[4]) {
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int rr = p * 8 + (lane >> 2), ch = lane & 3;
    const int grow = row0 + rr, gn = n0 + ch * 8;
    u[p] = make_uint4(0u, 0u, 0u, 0u);
    if (grow < M && gn < N) u[p] = *reinterpret_cast<const uint4*>(base + static_cast<size_t>(grow) * ld + gn);
  }
}
// ... and later transpose them through the warp's scratch into this thread's row (32 floats).  Issued a whole chunk
// ahead (the first one before the accumulator is even ready), so the global-load latency is off the critical path.
__device__ __forceinline__ void epi_unpack(uint8_t* scratch, const uint4 (&u)[4], int lane, float (&out)[32]) {
  __syncwarp();
#pragma unroll
  for (int p = 0; p < 4; ++p) *reinterpret_cast<uint4*>(scratch + epi_off(p * 8 + (lane >> 2), lane & 3)) = u[p];
  __syncwarp();
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {
    const uint4 w = *reinterpret_cast<const uint4*>(scratch + epi_off(lane, ch));
    const float2 a = unpack_bf16x2(w.x), b = unpack_bf16x2(w.y), c = unpack_bf16x2(w.z), d = unpack_bf16x2(w.w);
    out[ch * 8 + 0] = a.x; out[ch * 8 + 1] = a.y; out[ch * 8 + 2] = b.x; out[ch * 8 + 3] = b.y;
    out[ch * 8 + 4] = c.x; out[ch * 8 + 5] = c.y; out[ch * 8 + 6] = d.x; out[ch * 8 + 7] = d.y;
  }
}
__device__ __forceinline__ void epi_load_block(uint8_t* scratch, const bf16* base, int ld, int row0, int n0, int M, int N,
                                               int lane, float (&out)[32]) {
  uint4 u[4];
  epi_prefetch(base, ld, row0, n0, M, N, lane, u);
  epi_unpack(scratch, u, lane, out);
}
// this thread's row (32 floats, already bf16-exact) -> block of a row-major bf16 matrix, columns < ncols_limit only
__device__ __forceinline__ void epi_store_block(uint8_t* scratch, bf16* base, int ld, int row0, int n0, int M, int ncols_limit,
                                                int lane, const float (&v)[32]) {
  __syncwarp();
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {
    uint4 u;
    u.x = pack_bf16x2(v[ch * 8 + 0], v[ch * 8 + 1]);
    u.y = pack_bf16x2(v[ch * 8 + 2], v[ch * 8 + 3]);
    u.z = pack_bf16x2(v[ch * 8 + 4], v[ch * 8 + 5]);
    u.w = pack_bf16x2(v[ch * 8 + 6], v[ch * 8 + 7]);
    *reinterpret_cast<uint4*>(scratch + epi_off(lane, ch)) = u;
  }
  __syncwarp();
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int rr = p * 8 + (lane >> 2), ch = lane & 3;
    const int grow = row0 + rr, gn = n0 + ch * 8;
    if (grow < M && gn < ncols_limit)
      *reinterpret_cast<uint4*>(base + static_cast<size_t>(grow) * ld + gn) = *reinterpret_cast<const uint4*>(scratch + epi_off(rr, ch));
  }
}

// 32 rows (row0 + lane) x 32 columns (n0 ..): the whole warp must call this together.
// `pf` holds the prefetched block of the "extra input" stream: res when present, else aux_in.
__device__ __forceinline__ void epilogue_warp32_bf16(const GemmArgs& g, uint8_t* scratch, int row0, int lane, int n0,
                                                     const uint32_t (&r)[32], const uint4 (&pf)[4]) {
  const int row = row0 + lane;
  const int rowc = row < g.M ? row : g.M - 1;  // clamp for per-row parameter loads; stores are guarded
  const int sample = (g.gate != nullptr || g.row_alpha != nullptr) ? rowc / g.rows_per_sample : 0;
  const float alpha = g.alpha * (g.row_alpha != nullptr ? g.row_alpha[sample] : 1.0f);
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = alpha * __uint_as_float(r[i]);
  if (g.bias != nullptr) {
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      if (n0 + j < g.N) {
        float b[8];
        load8_bf16(g.bias + n0 + j, b);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[j + i] += b[i];
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = bf16_round(v[i]);
  const int act_lim = min(g.act_ncols, g.N);
  if (g.aux_out != nullptr && n0 < act_lim) epi_store_block(scratch, g.aux_out, g.ldaux_out, row0, n0, g.M, act_lim, lane, v);
  if (g.act == B200_ACT_GELU_TANH) {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (n0 + i < act_lim) v[i] = bf16_round(gelu_tanh(v[i]));
  }
  if (g.aux_in != nullptr) {
    float a[32];
    if (g.res == nullptr)
      epi_unpack(scratch, pf, lane, a);
    else
      epi_load_block(scratch, g.aux_in, g.ldaux_in, row0, n0, g.M, g.N, lane, a);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = bf16_round(v[i] * gelu_tanh_grad(a[i]));
  }
  if (g.gate != nullptr) {
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      if (n0 + j < g.N) {
        float gt[8];
        load8_bf16(g.gate + static_cast<size_t>(sample) * g.ldgate + n0 + j, gt);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[j + i] = bf16_round(v[j + i] * gt[i]);
      }
    }
  }
  if (g.res != nullptr) {
    float rs[32];
    epi_unpack(scratch, pf, lane, rs);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = bf16_round(v[i] + rs[i]);
  }
  epi_store_block(scratch, reinterpret_cast<bf16*>(g.out), g.ldo, row0, n0, g.M, g.N, lane, v);
}

template <int CG, int BN, int STAGES, int A_MN, int B_MN>
struct GemmCfg {
  static constexpr int BM = 128;
  static constexpr int BK = 64;
  static constexpr int BNL = BN / CG;  // rows of the B tile this CTA loads
  static constexpr uint32_t A_BYTES = BM * BK * 2;
  static constexpr uint32_t B_BYTES = BNL * BK * 2;
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = 2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512));  // two accumulators, BN apart
  static constexpr size_t EPI_SCRATCH = 8 * 2048;  // per epilogue warp: 32 x 64 B transpose buffer
  static constexpr size_t SMEM_BYTES = 1024 + static_cast<size_t>(STAGES) * STAGE_BYTES + EPI_SCRATCH + (2 * STAGES + 4) * 8 + 16;
  static_assert(BN % 16 == 0 && BN >= 32 && BN <= 256, "invalid UMMA N");
  static_assert(TMEM_COLS == 64 || TMEM_COLS == 128 || TMEM_COLS == 256 || TMEM_COLS == 512, "TMEM cols pow2");
  static_assert(SMEM_BYTES <= 232448, "exceeds 227 KB of shared memory");
};

template <int CG, int BN, int STAGES, int A_MN, int B_MN>
__global__ void __launch_bounds__(320, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmB0,
                 const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
                 const GemmArgs g) {
  pdl_launch_dependents();
  using C = GemmCfg<CG, BN, STAGES, A_MN, B_MN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_base_1024(smem_raw);
  uint8_t* epi_scratch = smem + STAGES * C::STAGE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(epi_scratch + C::EPI_SCRATCH);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = warp_id_uniform();
  const int lane = threadIdx.x & 31;
  uint32_t cta_rank = 0;
  if (CG == 2) {
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
    cta_rank = __shfl_sync(0xffffffffu, cta_rank, 0);  // provably warp-uniform for the compiler
  }
  const bool leader = (cta_rank == 0);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmB0);
    if (g.kb1 > 0) {
      tma_prefetch_desc(&tmA1);
      tma_prefetch_desc(&tmB1);
    }
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&full[s], CG);  // one arrive.expect_tx per producing CTA
        mbar_init(&empty[s], 1);  // one tcgen05.commit
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(&tfull[a], 1);
        mbar_init(&tempty[a], 256 * CG);  // every epilogue thread of every CTA of the pair
      }
      fence_barrier_init();
    }
    __syncwarp();
    if (CG == 2)
      tmem_alloc_cg2<C::TMEM_COLS>(tmem_slot);
    else
      tmem_alloc<C::TMEM_COLS>(tmem_slot);
  }
  tc_fence_before();
  if (CG == 2)
    cluster_sync_all();
  else
    __syncthreads();
  tc_fence_after();
  pdl_wait();  // everything above touched only shared / tensor memory
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  const int m_tiles = (g.M + C::BM * CG - 1) / (C::BM * CG);
  const int n_tiles = (g.N + BN - 1) / BN;
  const int splits = g.splits > 1 ? g.splits : 1;
  const int total_tiles = m_tiles * n_tiles * splits;
  const int kb_total = g.kb0 + g.kb1;
  const int kb_per_split = (kb_total + splits - 1) / splits;
  const int worker = blockIdx.x / CG;
  const int nworkers = gridDim.x / CG;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = worker; t < total_tiles; t += nworkers) {
        int split, m_blk, n_blk;
        tile_coords(t, m_tiles, n_tiles, g.group_m, split, m_blk, n_blk);
        int row_a = m_blk * C::BM * CG + static_cast<int>(cta_rank) * C::BM;
        if (g.dual_mt0 > 0 && m_blk >= g.dual_mt0) row_a -= g.dual_mt0 * C::BM;  // problem 1 rows restart at 0
        const int row_b = n_blk * BN + static_cast<int>(cta_rank) * C::BNL;
        const int kb_begin = split * kb_per_split;
        const int kb_end = min(kb_total, kb_begin + kb_per_split);
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1u, 1);
          uint8_t* sa = smem + stage * C::STAGE_BYTES;
          uint8_t* sb = sa + C::A_BYTES;
          const bool seg1 = g.dual_mt0 > 0 ? (m_blk >= g.dual_mt0) : (kb >= g.kb0);
          const CUtensorMap* ma = seg1 ? &tmA1 : &tmA0;
          const CUtensorMap* mb = seg1 ? &tmB1 : &tmB0;
          const int kc = ((seg1 && g.dual_mt0 == 0) ? kb - g.kb0 : kb) * C::BK;
          // K-major operand: one box {64 k, rows}.  MN-major operand (stored [k][mn], mn contiguous):
          // boxes of {64 mn, 64 k}, 8 KB each, one per 64 rows of the tile (descriptor LBO = 8192).
          if (CG == 2) {
            const uint32_t bar = smem_u32(&full[stage]) & kPeerMask;  // the pair leader's barrier
            mbar_arrive_expect_tx_cluster(bar, C::STAGE_BYTES);
            if (A_MN) {
#pragma unroll
              for (int i = 0; i < C::BM / 64; ++i) tma_load_2d_cg2(smem_u32(sa) + i * 8192, ma, bar, row_a + 64 * i, kc);
            } else {
              tma_load_2d_cg2(smem_u32(sa), ma, bar, kc, row_a);
            }
            if (B_MN) {
#pragma unroll
              for (int i = 0; i < C::BNL / 64; ++i) tma_load_2d_cg2(smem_u32(sb) + i * 8192, mb, bar, row_b + 64 * i, kc);
            } else {
              tma_load_2d_cg2(smem_u32(sb), mb, bar, kc, row_b);
            }
          } else {
            mbar_arrive_expect_tx(&full[stage], C::STAGE_BYTES);
            if (A_MN) {
#pragma unroll
              for (int i = 0; i < C::BM / 64; ++i) tma_load_2d(sa + i * 8192, ma, &full[stage], row_a + 64 * i, kc);
            } else {
              tma_load_2d(sa, ma, &full[stage], kc, row_a);
            }
            if (B_MN) {
#pragma unroll
              for (int i = 0; i < C::BNL / 64; ++i) tma_load_2d(sb + i * 8192, mb, &full[stage], row_b + 64 * i, kc);
            } else {
              tma_load_2d(sb, mb, &full[stage], kc, row_b);
            }
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (converged warp, one elected lane)
    if (leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(C::BM * CG, BN, A_MN, B_MN);
      // k-step of 16: +32 B inside the swizzle row (K-major) or +16 rows * 128 B (MN-major), in 16-byte units
      constexpr uint32_t a_kstep = A_MN ? 128u : 2u;
      constexpr uint32_t b_kstep = B_MN ? 128u : 2u;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = worker; t < total_tiles; t += nworkers) {
        int split, m_blk, n_blk;
        tile_coords(t, m_tiles, n_tiles, g.group_m, split, m_blk, n_blk);
        const int kb_begin = split * kb_per_split;
        const int kb_end = min(kb_total, kb_begin + kb_per_split);
        mbar_wait(&tempty[acc], acc_phase ^ 1u, 2);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(&full[stage], phase, 3);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
          const uint32_t sb = sa + C::A_BYTES;
          const uint64_t da = umma_desc_sw128(sa, 1024, A_MN ? 8192 : 16);
          const uint64_t db = umma_desc_sw128(sb, 1024, B_MN ? 8192 : 16);
#pragma unroll
          for (int k = 0; k < C::BK / 16; ++k) {
            const uint32_t accum = (kb > kb_begin || k > 0) ? 1u : 0u;
            if (CG == 2)
              umma_bf16_ss_cg2_w(d_tmem, da + a_kstep * k, db + b_kstep * k, idesc, accum);
            else
              umma_bf16_ss_w(d_tmem, da + a_kstep * k, db + b_kstep * k, idesc, accum);
          }
          if (CG == 2)
            umma_commit_cg2_mc_w(&empty[stage]);
          else
            umma_commit_w(&empty[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (CG == 2)
          umma_commit_cg2_mc_w(&tfull[acc]);
        else
          umma_commit_w(&tfull[acc]);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9)
    // two warps per TMEM sub-partition: warp (q, half) drains the 32-column chunks c with (c & 1) == half
    const int q = warp & 3;  // TMEM sub-partition this warp may read
    const int half = (warp - 2) >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = worker; t < total_tiles; t += nworkers) {
      int split, m_blk, n_blk;
      tile_coords(t, m_tiles, n_tiles, g.group_m, split, m_blk, n_blk);
      const int kb_begin = split * kb_per_split;
      const int kb_end = min(kb_total, kb_begin + kb_per_split);
      const int row = m_blk * C::BM * CG + static_cast<int>(cta_rank) * C::BM + q * 32 + lane;
      const int n_base = n_blk * BN;
      // extra input stream of the bf16 epilogue (residual, or the saved pre-activation for gelu'): fetched one 32-column
      // chunk ahead; the first chunk is requested before the accumulator of this tile is even complete
      const bool bf16_epi = (g.f32_mode == 0 && g.dual_mt0 == 0);
      const bf16* in_ptr = g.res != nullptr ? g.res : g.aux_in;
      const int in_ld = g.res != nullptr ? g.ldres : g.ldaux_in;
      const int n_limit = g.f32_mode ? g.n_store : g.N;
      uint4 pf[4] = {};
      if (bf16_epi && in_ptr != nullptr && n_base + half * 32 < n_limit)
        epi_prefetch(in_ptr, in_ld, row - lane, n_base + half * 32, g.M, g.N, lane, pf);
      mbar_wait(&tfull[acc], acc_phase, 4);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BN);
#pragma unroll 1
      for (int c = half; c < BN / 32; c += 2) {
        if (n_base + c * 32 >= n_limit) break;
        uint32_t r[32];
        if (kb_end > kb_begin) {
          tmem_ld_32x32(t_addr + static_cast<uint32_t>(c * 32), r);
          tmem_ld_wait();
        } else {  // empty split-K slice: its partial result is exactly zero
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = 0u;
        }
        if (!bf16_epi) {
          epilogue_row32(g, row, n_base + c * 32, split, r);
        } else {
          uint4 cur[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) cur[i] = pf[i];
          if (in_ptr != nullptr && c + 2 < BN / 32 && n_base + (c + 2) * 32 < n_limit)
            epi_prefetch(in_ptr, in_ld, row - lane, n_base + (c + 2) * 32, g.M, g.N, lane, pf);
          epilogue_warp32_bf16(g, epi_scratch + (warp - 2) * 2048, row - lane, lane, n_base + c * 32, r, cur);
        }
      }
      tc_fence_before();
      if (CG == 2)
        mbar_arrive_cluster(smem_u32(&tempty[acc]) & kPeerMask);
      else
        mbar_arrive(&tempty[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
  }

  // ------------------------------------------------------------------ teardown
  tc_fence_before();
  if (CG == 2)
    cluster_sync_all();
  else
    __syncthreads();
  tc_fence_after();
  if (warp == 1) {
    if (CG == 2)
      tmem_dealloc_cg2<C::TMEM_COLS>(tmem_base);
    else
      tmem_dealloc<C::TMEM_COLS>(tmem_base);
  }
}

template <int CG, int BN, int STAGES, int A_MN, int B_MN>
static int launch_gemm(b200_ctx* ctx, const CUtensorMap& a0, const CUtensorMap& b0, const CUtensorMap& a1,
                       const CUtensorMap& b1, const GemmArgs& args, cudaStream_t stream) {
  using C = GemmCfg<CG, BN, STAGES, A_MN, B_MN>;
  auto kern = gemm_bf16_kernel<CG, BN, STAGES, A_MN, B_MN>;
  static bool configured = false;
  if (!configured) {
    B200_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES));
    configured = true;
  }
  const int m_tiles = (args.M + C::BM * CG - 1) / (C::BM * CG);
  const int n_tiles = (args.N + BN - 1) / BN;
  const int splits = args.splits > 1 ? args.splits : 1;
  const int total = m_tiles * n_tiles * splits;
  int workers = ctx->sm_count / CG;
  if (total < workers) workers = total;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(workers * CG, 1, 1);
  cfg.blockDim = dim3(320, 1, 1);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int n_attr = 0;
  if (CG == 2) {
    attr[n_attr].id = cudaLaunchAttributeClusterDimension;
    attr[n_attr].val.clusterDim.x = CG;
    attr[n_attr].val.clusterDim.y = 1;
    attr[n_attr].val.clusterDim.z = 1;
    ++n_attr;
  }
#if B200_PDL
  attr[n_attr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[n_attr].val.programmaticStreamSerializationAllowed = 1;
  ++n_attr;
#endif
  cfg.attrs = attr;
  cfg.numAttrs = n_attr;
  B200_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, a0, b0, a1, b1, args));
  ctx->launches.fetch_add(1);
  return B200_OK;
}

}  // namespace b200

namespace b200 {
int skinny_gemm_dispatch(b200_ctx* ctx, const b200_gemm_desc* d, cudaStream_t stream);  // skinny_gemm.cu
}

extern "C" int b200_gemm_bf16(b200_ctx* ctx, const b200_gemm_desc* d, void* stream_v) {
  using namespace b200;
  int rc = check_ctx(ctx);
  if (rc != B200_OK) return rc;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  B200_REQUIRE(d != nullptr, "b200_gemm_bf16: null descriptor");
  B200_REQUIRE(d->M > 0 && d->N > 0 && d->K0 > 0 && d->K1 >= 0, "b200_gemm_bf16: bad shape M=%d N=%d K0=%d K1=%d", d->M,
               d->N, d->K0, d->K1);
  B200_REQUIRE(d->A0 && d->B0 && d->out, "b200_gemm_bf16: null operand");
  const bool f32 = d->f32_mode != 0;
  B200_REQUIRE(d->f32_mode >= 0 && d->f32_mode <= 2, "b200_gemm_bf16: f32_mode %d", d->f32_mode);
  if (!f32) B200_REQUIRE(d->N % 8 == 0 && d->ldo % 8 == 0, "b200_gemm_bf16: N (%d) and ldo (%d) must be multiples of 8",
                         d->N, d->ldo);
  if (!d->trans_a) B200_REQUIRE(d->K0 % 8 == 0 && d->K1 % 8 == 0, "b200_gemm_bf16: K0/K1 must be multiples of 8");
  if (d->trans_a) B200_REQUIRE(d->M % 8 == 0, "b200_gemm_bf16: trans_a needs M %% 8 == 0");
  if (d->trans_b) B200_REQUIRE(d->N % 8 == 0, "b200_gemm_bf16: trans_b needs N %% 8 == 0");
  if (d->K1 > 0) B200_REQUIRE(d->A1 && d->B1, "b200_gemm_bf16: K1 > 0 but A1/B1 null");
  if (d->gate) B200_REQUIRE(d->rows_per_sample > 0 && d->ldgate % 8 == 0, "b200_gemm_bf16: gate needs rows_per_sample");
  if (d->row_alpha) B200_REQUIRE(d->rows_per_sample > 0, "b200_gemm_bf16: row_alpha needs rows_per_sample");
  if (d->res) B200_REQUIRE(d->ldres % 8 == 0, "b200_gemm_bf16: ldres %% 8");
  if (d->aux_in) B200_REQUIRE(d->ldaux_in % 8 == 0, "b200_gemm_bf16: ldaux_in %% 8");
  if (d->aux_out) B200_REQUIRE(d->ldaux_out % 8 == 0, "b200_gemm_bf16: ldaux_out %% 8");
  const int splits = d->splits > 1 ? d->splits : 1;
  if (splits > 1) B200_REQUIRE(f32, "b200_gemm_bf16: split-K needs an fp32 output mode");
  if (f32)
    B200_REQUIRE(!d->bias && !d->res && !d->gate && !d->aux_in && !d->aux_out && d->act == 0,
                 "b200_gemm_bf16: fp32 output excludes the fused epilogue");
  if (d->f32_mode == 1) B200_REQUIRE(d->ldo % 4 == 0 && !d->f32_trans, "b200_gemm_bf16: f32 partial store needs ldo %% 4");

  int config = d->config;
  const bool plain_bf16 = !f32 && !d->trans_a && d->K1 == 0 && !d->bias && !d->res && !d->gate && !d->aux_in && !d->aux_out &&
                          d->act == 0 && d->N <= 64;
  // Rank-side GEMMs (N <= 64).  Measured on B200, L2-hot, persistent 128 x 64 kernel vs cluster split-K + DSMEM reduce:
  //   M 4608, K 3072: 16 vs 18.5 us | K 9216: 35 vs 27 | K 21504: 71 vs 58 | M 4096, K 12288: 43 vs 21 | M 512, K 9216: 33 vs 11
  // -> the cluster variant wins as soon as the contraction is long or the row tiles are few.
  const int sk_mtiles = (d->M + 127) / 128;
  static int sk_always = -1;  // B200_SKINNY_ALWAYS=1 (tuning): cluster kernel for every plain rank-side GEMM
  if (sk_always < 0) {
    const char* e = getenv("B200_SKINNY_ALWAYS");
    sk_always = (e && atoi(e) == 1) ? 1 : 0;
  }
  const bool sk_cluster = plain_bf16 && (d->K0 >= 6144 || sk_mtiles <= 16 || sk_always) && d->K0 >= 512;
  if ((config == B200_GEMM_AUTO && sk_cluster) || config == B200_GEMM_SKINNY_CLUSTER) {
    B200_REQUIRE(plain_bf16, "b200_gemm_bf16: SKINNY_CLUSTER needs N <= 64, bf16 output, one segment and no fused epilogue");
    return skinny_gemm_dispatch(ctx, d, stream);
  }
  if (config == B200_GEMM_AUTO) {
    if (d->N <= 64)
      config = B200_GEMM_1CTA_N64;
    else if (d->N <= 128)
      config = B200_GEMM_1CTA_N128;
    else {
      static int def = -1;
      if (def < 0) {
        const char* e = getenv("B200_GEMM_CG");
        def = (e && e[0] == '1') ? B200_GEMM_1CTA_N256 : B200_GEMM_2CTA_N256;
      }
      config = def;
      // few rows (the 512-token text stream): 256 x 256 pair tiles would leave most SMs idle; 128 x 128 tiles give
      // 4x the CTAs at the cost of operand re-reads that L2 absorbs
      const long long pair_tiles = static_cast<long long>((d->M + 255) / 256) * ((d->N + 255) / 256);
      static int pair_min = -1;  // B200_GEMM_PAIR_MIN (tuning): fewer pair tiles than this -> 128 x 128 tiles; default SMs / 2
      if (pair_min < 0) {
        const char* e = getenv("B200_GEMM_PAIR_MIN");
        pair_min = (e && atoi(e) > 0) ? atoi(e) : ctx->sm_count / 2;
      }
      if (pair_tiles < pair_min) {
        config = B200_GEMM_1CTA_N128;
        // wave quantisation of the 128-row tiles: cost = waves x tile width.  SDXL's [2048, 1280] outputs are 160 tiles of
        // 128 x 128 on 148 SMs (2 waves, the second one 8 % full); 128 x 160 tiles make it one wave.  B200_GEMM_WAVE_TUNE=0: off
        static int wave_tune = -1;
        if (wave_tune < 0) {
          const char* e = getenv("B200_GEMM_WAVE_TUNE");
          wave_tune = (e && e[0] == '0') ? 0 : 1;
        }
        if (wave_tune && !f32 && !d->trans_a) {
          const long long mt = (d->M + 127) / 128;
          auto cost = [&](int bn) {
            const long long tiles = mt * ((d->N + bn - 1) / bn);
            return ((tiles + ctx->sm_count - 1) / ctx->sm_count) * bn;
          };
          long long best = cost(128);
          if (!d->trans_b && cost(160) < best) {
            best = cost(160);
            config = B200_GEMM_1CTA_N160;
          }
          if (cost(192) < best) config = B200_GEMM_1CTA_N192;
        }
      }
    }
  }

  GemmArgs a;
  a.M = d->M;
  a.N = d->N;
  a.kb0 = (d->K0 + 63) / 64;
  a.kb1 = (d->K1 + 63) / 64;
  a.splits = splits;
  a.group_m = 8;
  a.bias = reinterpret_cast<const bf16*>(d->bias);
  a.res = reinterpret_cast<const bf16*>(d->res);
  a.ldres = d->ldres;
  a.gate = reinterpret_cast<const bf16*>(d->gate);
  a.ldgate = d->ldgate;
  a.rows_per_sample = d->rows_per_sample > 0 ? d->rows_per_sample : 1;
  a.aux_in = reinterpret_cast<const bf16*>(d->aux_in);
  a.ldaux_in = d->ldaux_in;
  a.aux_out = reinterpret_cast<bf16*>(d->aux_out);
  a.ldaux_out = d->ldaux_out;
  a.out = d->out;
  a.ldo = d->ldo;
  a.act = d->act;
  a.act_ncols = (d->act_ncols > 0 && d->act_ncols < d->N) ? d->act_ncols : d->N;
  a.f32_mode = d->f32_mode;
  a.f32_trans = d->f32_trans;
  a.n_store = (d->n_store > 0 && d->n_store < d->N) ? d->n_store : d->N;
  a.alpha = d->alpha;
  a.row_alpha = reinterpret_cast<const float*>(d->row_alpha);
  a.dual_mt0 = 0;
  a.M0 = a.M1 = 0;
  a.out1 = nullptr;
  a.ldo1 = 0;
  a.f32_trans1 = 0;

  int cg = 1, bn = 256;
  switch (config) {
    case B200_GEMM_1CTA_N256: cg = 1; bn = 256; break;
    case B200_GEMM_2CTA_N256: cg = 2; bn = 256; break;
    case B200_GEMM_1CTA_N128: cg = 1; bn = 128; break;
    case B200_GEMM_1CTA_N64: cg = 1; bn = 64; break;
    case B200_GEMM_1CTA_N160: cg = 1; bn = 160; break;
    case B200_GEMM_1CTA_N192: cg = 1; bn = 192; break;
    default: set_error("b200_gemm_bf16: unknown config %d", config); return B200_ERR_INVALID;
  }
  {
    // Rasterisation: tiles of `group_m` consecutive tile-rows are visited column by column, so the A rows of a group
    // stay in L2 while B streams past once per group.  DRAM traffic ~ A + B * (m_tiles / group_m): make the group as
    // tall as ~30 MB of A allows (the whole M for the K = 3072 layers, where A is the 28 MB activation).
    const long long row_tile_bytes = 128LL * cg * (static_cast<long long>(d->K0) + d->K1) * 2;
    const int m_tiles_h = (d->M + 128 * cg - 1) / (128 * cg);
    // measured sweep (tools/sweep_group_m.py): if all of A fits (<= 32 MB: the K = 3072 layers) keep every tile-row in one
    // group; otherwise ~16 MB of A per group (2 tile-rows at K = 12288 .. 21504) is the best or within 0.5 % of it
    long long gm = m_tiles_h;
    if (row_tile_bytes * m_tiles_h > (32LL << 20)) gm = (16LL << 20) / (row_tile_bytes > 0 ? row_tile_bytes : 1);
    if (gm < 2) gm = 2;
    if (gm > m_tiles_h) gm = m_tiles_h;
    a.group_m = static_cast<int>(gm);
    const char* e = getenv("B200_GEMM_GROUP_M");
    if (e) a.group_m = atoi(e) > 0 ? atoi(e) : a.group_m;
  }
  const int ta = d->trans_a ? 1 : 0, tb = d->trans_b ? 1 : 0;
  const uint32_t box_b_rows = static_cast<uint32_t>(bn / cg);
  // K-major operand [rows = M|N, cols = K]: box {64 k, tile rows}.  MN-major operand [rows = K, cols = M|N]: box {64, 64}.
  auto map_a = [&](CUtensorMap* t, const void* p, int K, int ld) {
    return ta ? make_tmap_bf16_2d(ctx, t, p, K, d->M, ld, 64, 64) : make_tmap_bf16_2d(ctx, t, p, d->M, K, ld, 64, 128);
  };
  auto map_b = [&](CUtensorMap* t, const void* p, int K, int ld) {
    return tb ? make_tmap_bf16_2d(ctx, t, p, K, d->N, ld, 64, 64)
              : make_tmap_bf16_2d(ctx, t, p, d->N, K, ld, 64, box_b_rows);
  };
  CUtensorMap tA0, tB0, tA1, tB1;
  rc = map_a(&tA0, d->A0, d->K0, d->lda0);
  if (rc) return rc;
  rc = map_b(&tB0, d->B0, d->K0, d->ldb0);
  if (rc) return rc;
  if (d->K1 > 0) {
    rc = map_a(&tA1, d->A1, d->K1, d->lda1);
    if (rc) return rc;
    rc = map_b(&tB1, d->B1, d->K1, d->ldb1);
    if (rc) return rc;
  } else {
    tA1 = tA0;
    tB1 = tB0;
  }
#define B200_LAUNCH(CG_, BN_, ST_, TA_, TB_) \
  if (cg == CG_ && bn == BN_ && ta == TA_ && tb == TB_) return launch_gemm<CG_, BN_, ST_, TA_, TB_>(ctx, tA0, tB0, tA1, tB1, a, stream)
  B200_LAUNCH(2, 256, 6, 0, 0);   // forward:  Y = X W^T
  B200_LAUNCH(2, 256, 6, 0, 1);   // dgrad:    dX = dY W
  B200_LAUNCH(1, 256, 4, 0, 0);
  B200_LAUNCH(1, 256, 4, 0, 1);
  B200_LAUNCH(1, 128, 6, 0, 0);
  B200_LAUNCH(1, 128, 6, 0, 1);
  B200_LAUNCH(1, 160, 5, 0, 0);
  B200_LAUNCH(1, 192, 5, 0, 0);
  B200_LAUNCH(1, 192, 5, 0, 1);
  B200_LAUNCH(1, 64, 8, 0, 0);    // rank-side: Z = X A^T
  B200_LAUNCH(1, 64, 8, 0, 1);    // rank-side: T = dY B
  B200_LAUNCH(1, 64, 8, 1, 1);    // wgrad:     dB = dY^T Z, dA^T = X^T T
#undef B200_LAUNCH
  set_error("b200_gemm_bf16: no kernel for config %d trans_a %d trans_b %d", config, ta, tb);
  return B200_ERR_INVALID;
}

// dB[out, r] += alpha * dY^T Zc   and   dA[r, in] += alpha * T^T X   in ONE launch (both contract over the tokens)
extern "C" int b200_lora_wgrad(b200_ctx* ctx, const void* dY, int lddy, const void* Zc, int ldz, const void* X, int ldx,
                               const void* T, int ldt, void* dB, void* dA, int tokens, int out_dim, int in_dim, int r,
                               int zcols, float alpha, int splits, void* stream_v) {
  using namespace b200;
  int rc = check_ctx(ctx);
  if (rc != B200_OK) return rc;
  B200_REQUIRE(dY && Zc && X && T && dB && dA, "b200_lora_wgrad: null operand");
  B200_REQUIRE(tokens > 0 && out_dim % 8 == 0 && in_dim % 8 == 0 && r > 0 && r <= 64 && zcols % 8 == 0 && zcols >= r && zcols <= 64,
               "b200_lora_wgrad: bad shape tokens=%d out=%d in=%d r=%d zcols=%d", tokens, out_dim, in_dim, r, zcols);
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  GemmArgs a = {};
  const int mt0 = (out_dim + 127) / 128, mt1 = (in_dim + 127) / 128;
  a.M = mt0 * 128 + in_dim;  // virtual row space: problem 0 padded to whole tiles, then problem 1
  a.N = 64;
  a.kb0 = (tokens + 63) / 64;
  a.kb1 = 0;
  a.splits = splits > 1 ? splits : 1;
  a.group_m = mt0 + mt1;
  a.rows_per_sample = 1;
  a.out = dB;
  a.ldo = r;
  a.f32_mode = 2;
  a.f32_trans = 0;
  a.n_store = r;
  a.act_ncols = 64;
  a.alpha = alpha;
  a.dual_mt0 = mt0;
  a.M0 = out_dim;
  a.M1 = in_dim;
  a.out1 = dA;
  a.ldo1 = in_dim;
  a.f32_trans1 = 1;
  CUtensorMap tA0, tB0, tA1, tB1;
  if ((rc = make_tmap_bf16_2d(ctx, &tA0, dY, tokens, out_dim, lddy, 64, 64))) return rc;
  if ((rc = make_tmap_bf16_2d(ctx, &tB0, Zc, tokens, zcols, ldz, 64, 64))) return rc;
  if ((rc = make_tmap_bf16_2d(ctx, &tA1, X, tokens, in_dim, ldx, 64, 64))) return rc;
  if ((rc = make_tmap_bf16_2d(ctx, &tB1, T, tokens, zcols, ldt, 64, 64))) return rc;
  return launch_gemm<1, 64, 8, 1, 1>(ctx, tA0, tB0, tA1, tB1, a, stream);
}
