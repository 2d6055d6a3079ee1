// Attention forward, variant 5: TWO Q tiles per CTA in ping-pong (the FlashAttention-4 schedule on tcgen05 / TMEM).
//
// Why: in the one-tile kernels (attention.cu variant 1, attention_r2.cu variant 3) the chain
//   S_j MMA -> softmax_j -> P_j V_j MMA -> (softmax_{j+1} needs O stable / P buffer free)
// leaves the tensor pipe idle while the 128 x 128 exponentials are computed: 44-47 % tensor-pipe activity, 1022 TF/s vs
// 1486 TF/s for the library SDPA on the same box (gpurun_out/r2_trip.log).  Here one CTA owns 256 query rows as two
// independent 128-row tiles A and B with their own S / P and O accumulators in TMEM (S_A S_B O_A O_B = 4 x 128 columns =
// all 512); while the softmax warp-group of A turns S_A(j) into P_A(j), the tensor pipe runs P_B(j-1) V and S_B(j), and
// vice versa.  Per KV block the tensor pipe has 4 x 512 = 2048 cycles of work and each softmax warp-group one tile
// (128 rows x 128 columns, ONE thread per row: no cross-thread row statistics at all).
//
//   warp 0      TMA: Q_A, Q_B once; K_j, V_j through 2-stage rings
//   warp 1      MMA issue (converged warp, elected lane):  S_A(0) S_B(0) | P V_A(j) S_A(j+1) P V_B(j) S_B(j+1) ...
//               S_t(j+1) overwrites the TMEM columns that hold P_t(j): it is issued AFTER P V_t(j) by the same thread
//               and tcgen05.mma executes in issue order, so no barrier is needed for that hazard.
//   warps 2-3   idle (keep the softmax warp-groups aligned to the TMEM lane quarters: warp % 4 = quarter)
//   warps 4-7   softmax of tile A, warps 8-11 softmax of tile B: thread = one row.
//               pass 1: row max of the raw scores (tcgen05.ld, 3-input max); lazy rescale of O only when the running
//                       maximum grows by more than 2^8 (then, and only then, it waits for P V(j-1));
//               pass 2: P = exp2(S c - m) as packed FFMA2, a share of the exponentials on the FMA pipe (polynomial) because
//                       MUFU alone is exactly as slow as the MMAs (16 ex2 / clk / SM), bf16 pairs stored IN PLACE over S
//                       (P chunk c lands in columns already consumed), then one arrive on p_full[t].
// Softmax arithmetic and rounding points are those of variant 3 (validated against torch SDPA in fp32).
#include "attn_common.cuh"

namespace b200 {

// POLY (template): of every 4 column pairs, how many use the FMA-pipe polynomial (0 = all MUFU); B200_ATTN_PP_POLY picks it
constexpr int kPPThreads = 384;
constexpr int kPP3Threads = 576;  // variant 7: TMA warp, MMA warp, 16 softmax warps (two threads per row)
constexpr int kPP3Smem = 1024 + 6 * 32768 + 32 * 8 + 2 * 1024 * 4;
constexpr int kPP2Threads = 320;  // variant 6: no idle warps (more registers per thread for the whole-row P in registers)
constexpr int kPPSmem = 1024 + 6 * 32768 + 32 * 8;

__device__ __forceinline__ uint8_t* smem_align1024_pp(uint8_t* raw) {
  return raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);  // offset arithmetic: stays in the shared address space
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float max32(const uint32_t (&v)[32], float m) {
  const float* x = reinterpret_cast<const float*>(v);
  float m0 = m, m1 = -INFINITY;
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    m0 = fmax3(m0, x[i], x[i + 1]);
    m1 = fmax3(m1, x[i + 2], x[i + 3]);
  }
  return fmaxf(m0, m1);
}
__device__ __forceinline__ void mask32(uint32_t (&v)[32], int nvalid) {
#pragma unroll
  for (int i = 0; i < 32; ++i)
    if (i >= nvalid) v[i] = 0xff800000u;  // -inf
}

template <int POLY>
__global__ void __launch_bounds__(kPPThreads, 1)
attn_fwd_pp_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const AttnFwdArgs g) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024_pp(smem_raw);
  uint8_t* sQ = smem;               // [2 tiles][2 halves of 64 head-dim columns][128 rows x 128 B]
  uint8_t* sK = sQ + 2 * 32768;     // [2 stages]
  uint8_t* sV = sK + 2 * 32768;     // [2 stages]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 2 * 32768);
  uint64_t* q_full = bars;          // 1
  uint64_t* k_full = bars + 1;      // [2]
  uint64_t* k_empty = bars + 3;     // [2]
  uint64_t* v_full = bars + 5;      // [2]
  uint64_t* v_empty = bars + 7;     // [2]
  uint64_t* s_full = bars + 9;      // [2 tiles]
  uint64_t* p_full = bars + 11;     // [2 tiles]
  uint64_t* pv_done = bars + 13;    // [2 tiles]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

  const int warp = warp_id_uniform(), lane = threadIdx.x & 31;
  const int bh = blockIdx.y;
  const int q0 = blockIdx.x * 256;
  const int n_kv = (g.Lk + 127) / 128;
  const long long row_base = static_cast<long long>(bh) * g.L;   // query rows / lse
  const long long kv_base = static_cast<long long>(bh) * g.Lk;  // key / value rows

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(q_full, 1);
      for (int s = 0; s < 2; ++s) {
        mbar_init(&k_full[s], 1);
        mbar_init(&k_empty[s], 1);
        mbar_init(&v_full[s], 1);
        mbar_init(&v_empty[s], 1);
        mbar_init(&s_full[s], 1);
        mbar_init(&p_full[s], 128);
        mbar_init(&pv_done[s], 1);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();  // everything above touched only shared / tensor memory
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, 65536);
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int qrow = static_cast<int>(row_base + q0 + t * 128);
        tma_load_2d(sQ + t * 32768, &tmQ, q_full, 0, qrow);
        tma_load_2d(sQ + t * 32768 + 16384, &tmQ, q_full, 64, qrow);
      }
      for (int j = 0; j < n_kv; ++j) {
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        const int kvrow = static_cast<int>(kv_base + j * 128);
        mbar_wait(&k_empty[s], ph ^ 1u, 30);
        mbar_arrive_expect_tx(&k_full[s], 32768);
        tma_load_2d(sK + s * 32768, &tmK, &k_full[s], 0, kvrow);
        tma_load_2d(sK + s * 32768 + 16384, &tmK, &k_full[s], 64, kvrow);
        mbar_wait(&v_empty[s], ph ^ 1u, 31);
        mbar_arrive_expect_tx(&v_full[s], 32768);
#pragma unroll
        for (int jc = 0; jc < 2; ++jc)
#pragma unroll
          for (int ih = 0; ih < 2; ++ih)
            tma_load_2d(sV + s * 32768 + (jc * 2 + ih) * 8192, &tmV, &v_full[s], jc * 64, kvrow + ih * 64);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idS = umma_idesc_bf16(128, 128, 0, 0);
    constexpr uint32_t idPV = umma_idesc_bf16(128, 128, 0, 1);
    mbar_wait(q_full, 0, 32);
    const uint64_t dQ0 = umma_desc_sw128(smem_u32(sQ), 1024, 16);
    const uint64_t dK0 = umma_desc_sw128(smem_u32(sK), 1024, 16);
    const uint64_t dV0 = umma_desc_sw128(smem_u32(sV), 1024, 16384);
    auto issue_S = [&](int t, int j) {  // S_t(j) = Q_t K_j^T into columns [128 t, 128 t + 128)
      const uint64_t dq = dQ0 + static_cast<uint64_t>(t * (32768 >> 4));
      const uint64_t dk = dK0 + static_cast<uint64_t>((j & 1) * (32768 >> 4));
      {  // one elect for the 8 k-steps (attn_common.cuh): both operands K-major, halves of 64 head-dim columns 16 KB apart
        constexpr int HH = 16384 >> 4;
        umma_bf16_ss_w_x8<2, 4, 6, HH, HH + 2, HH + 4, HH + 6, 2, 4, 6, HH, HH + 2, HH + 4, HH + 6>(
            tmem_base + static_cast<uint32_t>(t) * 128u, dq, dk, idS, 0u);
      }
      umma_commit_w(&s_full[t]);
    };
    auto issue_PV = [&](int t, int j) {  // O_t += P_t(j) V_j, A = P straight from TMEM (bf16 pairs in S_t's first 64 columns)
      const uint64_t dv = dV0 + static_cast<uint64_t>((j & 1) * (32768 >> 4));
      {  // V MN-major: k-steps 2 KB apart inside a 64-row half, halves 8 KB apart
        constexpr int VS = 2048 >> 4, VH = 8192 >> 4;
        umma_bf16_ts_w_x8<8, 16, 24, 32, 40, 48, 56, VS, 2 * VS, 3 * VS, VH, VH + VS, VH + 2 * VS, VH + 3 * VS>(
            tmem_base + 256u + static_cast<uint32_t>(t) * 128u, tmem_base + static_cast<uint32_t>(t) * 128u, dv, idPV,
            j > 0 ? 1u : 0u);
      }
    };
    mbar_wait(&k_full[0], 0, 33);
    tc_fence_after();
    issue_S(0, 0);
    issue_S(1, 0);
    umma_commit_w(&k_empty[0]);
    for (int j = 0; j < n_kv; ++j) {
      const int s = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      const bool more = j + 1 < n_kv;
      mbar_wait(&v_full[s], ph, 34);
      if (more) mbar_wait(&k_full[s ^ 1], ((j + 1) >> 1) & 1, 35);
      // ---- tile A
      mbar_wait(&p_full[0], j & 1, 36);
      tc_fence_after();
      issue_PV(0, j);
      umma_commit_w(&pv_done[0]);
      if (more) issue_S(0, j + 1);
      // ---- tile B
      mbar_wait(&p_full[1], j & 1, 37);
      tc_fence_after();
      issue_PV(1, j);
      umma_commit_w(&v_empty[s]);
      umma_commit_w(&pv_done[1]);
      if (more) {
        issue_S(1, j + 1);
        umma_commit_w(&k_empty[s ^ 1]);
      }
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- softmax / epilogue: one thread per row
    const int t = (warp - 4) >> 2;
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int qi = q0 + t * 128 + r;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t tS = tmem_base + static_cast<uint32_t>(t) * 128u + lane_off;
    const uint32_t tO = tmem_base + 256u + static_cast<uint32_t>(t) * 128u + lane_off;
    const float c2 = g.scale * kLog2e;
    const float2 c22 = make_float2(c2, c2);
    const int tail = g.Lk & 127;
    float m_used = -INFINITY, l_sum = 0.f;
    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(&s_full[t], j & 1, 38);
      tc_fence_after();
      const bool ragged = (tail != 0) && (j == n_kv - 1);
      uint32_t a[32], b[32];
      // ---- pass 1: row max of the raw scores
      float mx = -INFINITY;
      tmem_ld_32x32(tS, a);
      tmem_ld_32x32(tS + 32, b);
      tmem_ld_wait();
      if (ragged) {
        mask32(a, tail);
        mask32(b, tail - 32);
      }
      mx = max32(a, mx);
      mx = max32(b, mx);
      tmem_ld_32x32(tS + 64, a);
      tmem_ld_32x32(tS + 96, b);
      tmem_ld_wait();
      if (ragged) {
        mask32(a, tail - 64);
        mask32(b, tail - 96);
      }
      mx = max32(a, mx);
      mx = max32(b, mx);
      const float m_new = fmaxf(m_used, mx * c2);  // c2 > 0: scaling commutes with the max
      const bool need = (m_new > m_used + 8.0f);   // also true on the first block (m_used = -inf)
      if (__any_sync(0xffffffffu, need)) {
        const float f = need ? ex2(m_used - m_new) : 1.0f;  // first block: ex2(-inf) = 0
        if (need) {
          m_used = m_new;
          l_sum *= f;
        }
        if (j > 0) {  // O_t must be stable: P V_t(j-1) is the newest MMA accumulating into it
          mbar_wait(&pv_done[t], (j - 1) & 1, 39);
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            uint32_t o[32];
            tmem_ld_32x32(tO + c * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
            tmem_st_32x32(tO + c * 32, o);
          }
          tmem_st_wait();
        }
      }
      // ---- pass 2: P = exp2(S c - m_used) -> bf16 pairs, stored in place over S (chunk c -> columns [16 c, 16 c + 16))
      const float2 nm2 = make_float2(-m_used, -m_used);
      float2 ls[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
      tmem_ld_32x32(tS, a);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t(&cur)[32] = (c & 1) ? b : a;
        uint32_t(&nxt)[32] = (c & 1) ? a : b;
        tmem_ld_wait();
        if (c < 3) tmem_ld_32x32(tS + (c + 1) * 32, nxt);  // in flight while this chunk is computed
        if (ragged) mask32(cur, tail - c * 32);
        const float* x = reinterpret_cast<const float*>(cur);
        uint32_t pk[16];
#pragma unroll
        for (int k2 = 0; k2 < 16; ++k2) {
          const bool poly = (k2 % 4) < POLY;
          const float2 e = ffma2(make_float2(x[2 * k2], x[2 * k2 + 1]), c22, nm2);
          const float2 p2 = poly ? ex2_poly2(e) : make_float2(ex2(e.x), ex2(e.y));
          ls[k2 & 1] = fadd2(ls[k2 & 1], p2);
          pk[k2] = pack_bf16x2(p2.x, p2.y);
        }
        tmem_st_32x16(tS + c * 16, pk);
      }
      l_sum += (ls[0].x + ls[0].y) + (ls[1].x + ls[1].y);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[t]);
    }
    mbar_wait(&pv_done[t], (n_kv - 1) & 1, 40);
    tc_fence_after();
    const float inv = 1.0f / l_sum;
    const bool live = qi < g.L;
    bf16* orow = nullptr;
    if (live) {
      const int bb = bh / g.H, hh = bh % g.H;
      if (qi < g.split)
        orow = g.o0 + (static_cast<size_t>(bb) * g.split + qi) * g.ld0 + hh * 128;
      else
        orow = g.o1 + (static_cast<size_t>(bb) * (g.L - g.split) + (qi - g.split)) * g.ld1 + hh * 128;
      g.lse[row_base + qi] = (m_used + log2f(l_sum)) * kLn2;
    }
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {  // tcgen05.ld is warp-collective: every lane loads, only live rows store
      uint32_t o[32];
      tmem_ld_32x32(tO + c * 32, o);
      tmem_ld_wait();
      if (live) {
#pragma unroll
        for (int k8 = 0; k8 < 4; ++k8) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(o[k8 * 8 + 0]) * inv, __uint_as_float(o[k8 * 8 + 1]) * inv);
          u.y = pack_bf16x2(__uint_as_float(o[k8 * 8 + 2]) * inv, __uint_as_float(o[k8 * 8 + 3]) * inv);
          u.z = pack_bf16x2(__uint_as_float(o[k8 * 8 + 4]) * inv, __uint_as_float(o[k8 * 8 + 5]) * inv);
          u.w = pack_bf16x2(__uint_as_float(o[k8 * 8 + 6]) * inv, __uint_as_float(o[k8 * 8 + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + c * 32 + k8 * 8) = u;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// DBG (timing experiments only, results are WRONG): 1 = no exponentials (P = the scaled score), 2 = no softmax work at all
// (the warps only hand the barriers on): what the MMA / barrier pipeline alone costs.  B200_ATTN_PP_DBG selects it.
template <int POLY, int DBG = 0>
__global__ void __launch_bounds__(kPP2Threads, 1)
attn_fwd_pp2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const AttnFwdArgs g) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024_pp(smem_raw);
  uint8_t* sQ = smem;               // [2 tiles][2 halves of 64 head-dim columns][128 rows x 128 B]
  uint8_t* sK = sQ + 2 * 32768;     // [2 stages]
  uint8_t* sV = sK + 2 * 32768;     // [2 stages]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 2 * 32768);
  uint64_t* q_full = bars;          // 1
  uint64_t* k_full = bars + 1;      // [2]
  uint64_t* k_empty = bars + 3;     // [2]
  uint64_t* v_full = bars + 5;      // [2]
  uint64_t* v_empty = bars + 7;     // [2]
  uint64_t* s_full = bars + 9;      // [2 tiles]
  uint64_t* p_full = bars + 11;     // [2 tiles]
  uint64_t* pv_done = bars + 13;    // [2 tiles]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

  const int warp = warp_id_uniform(), lane = threadIdx.x & 31;
  const int bh = blockIdx.y;
  const int q0 = blockIdx.x * 256;
  const int n_kv = (g.Lk + 127) / 128;
  const long long row_base = static_cast<long long>(bh) * g.L;   // query rows / lse
  const long long kv_base = static_cast<long long>(bh) * g.Lk;  // key / value rows

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(q_full, 1);
      for (int s = 0; s < 2; ++s) {
        mbar_init(&k_full[s], 1);
        mbar_init(&k_empty[s], 1);
        mbar_init(&v_full[s], 1);
        mbar_init(&v_empty[s], 1);
        mbar_init(&s_full[s], 1);
        mbar_init(&p_full[s], 128);
        mbar_init(&pv_done[s], 1);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();  // everything above touched only shared / tensor memory
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  if (warp == 0) {
    if (lane == 0) {
      const bool full_d = g.dlive != 64;  // dlive 64: the second 64 head-dim columns are zero padding, never loaded / multiplied
      mbar_arrive_expect_tx(q_full, full_d ? 65536 : 32768);
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int qrow = static_cast<int>(row_base + q0 + t * 128);
        tma_load_2d(sQ + t * 32768, &tmQ, q_full, 0, qrow);
        if (full_d) tma_load_2d(sQ + t * 32768 + 16384, &tmQ, q_full, 64, qrow);
      }
      for (int j = 0; j < n_kv; ++j) {
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        const int kvrow = static_cast<int>(kv_base + j * 128);
        mbar_wait(&k_empty[s], ph ^ 1u, 30);
        mbar_arrive_expect_tx(&k_full[s], full_d ? 32768 : 16384);
        tma_load_2d(sK + s * 32768, &tmK, &k_full[s], 0, kvrow);
        if (full_d) tma_load_2d(sK + s * 32768 + 16384, &tmK, &k_full[s], 64, kvrow);
        mbar_wait(&v_empty[s], ph ^ 1u, 31);
        mbar_arrive_expect_tx(&v_full[s], full_d ? 32768 : 16384);
#pragma unroll
        for (int jc = 0; jc < 2; ++jc)
#pragma unroll
          for (int ih = 0; ih < 2; ++ih)
            if (jc == 0 || full_d) tma_load_2d(sV + s * 32768 + (jc * 2 + ih) * 8192, &tmV, &v_full[s], jc * 64, kvrow + ih * 64);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idS = umma_idesc_bf16(128, 128, 0, 0);
    constexpr uint32_t idPV128 = umma_idesc_bf16(128, 128, 0, 1);
    constexpr uint32_t idPV64 = umma_idesc_bf16(128, 64, 0, 1);  // N = 64: the first MN-major atom of V = head-dim columns 0..63
    const bool full_d = g.dlive != 64;
    const uint32_t idPV = full_d ? idPV128 : idPV64;
    mbar_wait(q_full, 0, 32);
    const uint64_t dQ0 = umma_desc_sw128(smem_u32(sQ), 1024, 16);
    const uint64_t dK0 = umma_desc_sw128(smem_u32(sK), 1024, 16);
    const uint64_t dV0 = umma_desc_sw128(smem_u32(sV), 1024, 16384);
    auto issue_S = [&](int t, int j) {  // S_t(j) = Q_t K_j^T into columns [128 t, 128 t + 128)
      const uint64_t dq = dQ0 + static_cast<uint64_t>(t * (32768 >> 4));
      const uint64_t dk = dK0 + static_cast<uint64_t>((j & 1) * (32768 >> 4));
      {  // one elect for the 8 k-steps (attn_common.cuh): both operands K-major, halves of 64 head-dim columns 16 KB apart
        constexpr int HH = 16384 >> 4;
        if (full_d)
          umma_bf16_ss_w_x8<2, 4, 6, HH, HH + 2, HH + 4, HH + 6, 2, 4, 6, HH, HH + 2, HH + 4, HH + 6>(
              tmem_base + static_cast<uint32_t>(t) * 128u, dq, dk, idS, 0u);
        else  // the contraction over the zero half contributes nothing
          umma_bf16_ss_w_x4<2, 4, 6, 2, 4, 6>(tmem_base + static_cast<uint32_t>(t) * 128u, dq, dk, idS, 0u);
      }
      umma_commit_w(&s_full[t]);
    };
    auto issue_PV = [&](int t, int j) {  // O_t += P_t(j) V_j, A = P straight from TMEM (bf16 pairs in S_t's first 64 columns)
      const uint64_t dv = dV0 + static_cast<uint64_t>((j & 1) * (32768 >> 4));
      {  // V MN-major: k-steps 2 KB apart inside a 64-row half, halves 8 KB apart
        constexpr int VS = 2048 >> 4, VH = 8192 >> 4;
        umma_bf16_ts_w_x8<8, 16, 24, 32, 40, 48, 56, VS, 2 * VS, 3 * VS, VH, VH + VS, VH + 2 * VS, VH + 3 * VS>(
            tmem_base + 256u + static_cast<uint32_t>(t) * 128u, tmem_base + static_cast<uint32_t>(t) * 128u, dv, idPV,
            j > 0 ? 1u : 0u);
      }
    };
    mbar_wait(&k_full[0], 0, 33);
    tc_fence_after();
    issue_S(0, 0);
    issue_S(1, 0);
    umma_commit_w(&k_empty[0]);
    for (int j = 0; j < n_kv; ++j) {
      const int s = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      const bool more = j + 1 < n_kv;
      mbar_wait(&v_full[s], ph, 34);
      if (more) mbar_wait(&k_full[s ^ 1], ((j + 1) >> 1) & 1, 35);
      // ---- tile A
      mbar_wait(&p_full[0], j & 1, 36);
      tc_fence_after();
      issue_PV(0, j);
      umma_commit_w(&pv_done[0]);
      if (more) issue_S(0, j + 1);
      // ---- tile B
      mbar_wait(&p_full[1], j & 1, 37);
      tc_fence_after();
      issue_PV(1, j);
      umma_commit_w(&v_empty[s]);
      umma_commit_w(&pv_done[1]);
      if (more) {
        issue_S(1, j + 1);
        umma_commit_w(&k_empty[s ^ 1]);
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax / epilogue: one thread per row
    // warps 2-5 = tile A, 6-9 = tile B; warp & 3 is the TMEM lane quarter the warp may access (2, 3, 0, 1 within a tile)
    const int t = (warp - 2) >> 2;
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int qi = q0 + t * 128 + r;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t tS = tmem_base + static_cast<uint32_t>(t) * 128u + lane_off;
    const uint32_t tO = tmem_base + 256u + static_cast<uint32_t>(t) * 128u + lane_off;
    const float c2 = g.scale * kLog2e;
    const float2 c22 = make_float2(c2, c2);
    const int tail = g.Lk & 127;
    float m_used = -INFINITY, l_sum = 0.f;
    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(&s_full[t], j & 1, 38);
      tc_fence_after();
      const bool ragged = (tail != 0) && (j == n_kv - 1);
      uint32_t a[32], b[32], pk[64];
      // P = exp2(S c - m_used) for the whole row into REGISTERS (bf16 pairs) + the row's partial sum; S stays intact in TMEM
      auto pass2 = [&]() -> float {
        const float2 nm2 = make_float2(-m_used, -m_used);
        float2 ls[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
        tmem_ld_32x32(tS, a);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t(&cur)[32] = (c & 1) ? b : a;
          uint32_t(&nxt)[32] = (c & 1) ? a : b;
          tmem_ld_wait();
          if (c < 3) tmem_ld_32x32(tS + (c + 1) * 32, nxt);  // in flight while this chunk is computed
          if (ragged) mask32(cur, tail - c * 32);
          const float* x = reinterpret_cast<const float*>(cur);
#pragma unroll
          for (int k2 = 0; k2 < 16; ++k2) {
            const bool poly = (k2 % 4) < POLY;
            const float2 e = ffma2(make_float2(x[2 * k2], x[2 * k2 + 1]), c22, nm2);
            const float2 p2 = DBG == 1 ? e : (poly ? ex2_poly2(e) : make_float2(ex2(e.x), ex2(e.y)));
            ls[k2 & 1] = fadd2(ls[k2 & 1], p2);
            pk[c * 16 + k2] = pack_bf16x2(p2.x, p2.y);
          }
        }
        return (ls[0].x + ls[0].y) + (ls[1].x + ls[1].y);
      };
      // OPTIMISTIC: keep the running maximum of the earlier blocks.  bf16 and fp32 share the exponent range, so P may exceed 1
      // by many orders of magnitude without losing relative precision; only when the row's partial sum shows that the scores
      // have outgrown m_used by more than ~2^30 (or on the first block) is the exact row maximum computed and O rescaled.
      float bsum = 0.f;
      bool need = (j == 0) && DBG != 2;
      if (DBG == 2) {
        tc_fence_before();
        mbar_arrive(&p_full[t]);
        continue;
      }
      if (!need) {
        bsum = pass2();
        need = DBG == 0 && !(bsum < 1.0e9f);  // also catches inf / nan
      }
      if (__any_sync(0xffffffffu, need)) {
        float mx = -INFINITY;
        tmem_ld_32x32(tS, a);
        tmem_ld_32x32(tS + 32, b);
        tmem_ld_wait();
        if (ragged) {
          mask32(a, tail);
          mask32(b, tail - 32);
        }
        mx = max32(a, mx);
        mx = max32(b, mx);
        tmem_ld_32x32(tS + 64, a);
        tmem_ld_32x32(tS + 96, b);
        tmem_ld_wait();
        if (ragged) {
          mask32(a, tail - 64);
          mask32(b, tail - 96);
        }
        mx = max32(a, mx);
        mx = max32(b, mx);
        const float m_new = fmaxf(m_used, mx * c2);
        const bool grow = need && (m_new > m_used);
        const float f = grow ? ex2(m_used - m_new) : 1.0f;  // first block: ex2(-inf) = 0
        if (grow) {
          m_used = m_new;
          l_sum *= f;
        }
        if (j > 0) {  // O_t must be stable: P V_t(j-1) is the newest MMA accumulating into it
          mbar_wait(&pv_done[t], (j - 1) & 1, 39);
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            uint32_t o[32];
            tmem_ld_32x32(tO + c * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
            tmem_st_32x32(tO + c * 32, o);
          }
          tmem_st_wait();
        }
        bsum = pass2();
      }
      l_sum += bsum;
#pragma unroll
      for (int c = 0; c < 4; ++c)
        tmem_st_32x16(tS + c * 16, *reinterpret_cast<const uint32_t(*)[16]>(&pk[c * 16]));
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[t]);
    }
    mbar_wait(&pv_done[t], (n_kv - 1) & 1, 40);
    tc_fence_after();
    const float inv = 1.0f / l_sum;
    const bool live = qi < g.L;
    bf16* orow = nullptr;
    if (live) {
      const int bb = bh / g.H, hh = bh % g.H;
      if (qi < g.split)
        orow = g.o0 + (static_cast<size_t>(bb) * g.split + qi) * g.ld0 + hh * 128;
      else
        orow = g.o1 + (static_cast<size_t>(bb) * (g.L - g.split) + (qi - g.split)) * g.ld1 + hh * 128;
      g.lse[row_base + qi] = (m_used + log2f(l_sum)) * kLn2;
    }
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {  // tcgen05.ld is warp-collective: every lane loads, only live rows store
      uint32_t o[32];
      if (g.dlive == 64 && c >= 2) {  // padded head-dim columns: the accumulator was never written there -> exact zeros
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = 0u;
      } else {
        tmem_ld_32x32(tO + c * 32, o);
        tmem_ld_wait();
      }
      if (live) {
#pragma unroll
        for (int k8 = 0; k8 < 4; ++k8) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(o[k8 * 8 + 0]) * inv, __uint_as_float(o[k8 * 8 + 1]) * inv);
          u.y = pack_bf16x2(__uint_as_float(o[k8 * 8 + 2]) * inv, __uint_as_float(o[k8 * 8 + 3]) * inv);
          u.z = pack_bf16x2(__uint_as_float(o[k8 * 8 + 4]) * inv, __uint_as_float(o[k8 * 8 + 5]) * inv);
          u.w = pack_bf16x2(__uint_as_float(o[k8 * 8 + 6]) * inv, __uint_as_float(o[k8 * 8 + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + c * 32 + k8 * 8) = u;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// Variant 7: the optimistic ping-pong kernel with TWO threads per row (16 softmax warps).  Diagnostics of variant 6
// (profiles/r2_attn_diag.log): MMA / barrier pipeline alone 170 us, whole kernel 213 us -- the one-thread-per-row softmax
// is a serial chain on ONE active warp per scheduler.  Here thread (row, h) owns columns [64 h, 64 h + 64): half the chain, two
// warps per scheduler and tile.  The optimistic maximum removes the per-block row-max exchange; what remains per block is one
// flag exchange through shared memory + a 64-thread named barrier per (tile, lane quarter), which also orders "both halves have
// read S" before either half stores its P over it.
template <int POLY>
__global__ void __launch_bounds__(kPP3Threads, 1)
attn_fwd_pp3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const AttnFwdArgs g) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024_pp(smem_raw);
  uint8_t* sQ = smem;               // [2 tiles][2 halves of 64 head-dim columns][128 rows x 128 B]
  uint8_t* sK = sQ + 2 * 32768;     // [2 stages]
  uint8_t* sV = sK + 2 * 32768;     // [2 stages]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 2 * 32768);
  uint64_t* q_full = bars;          // 1
  uint64_t* k_full = bars + 1;      // [2]
  uint64_t* k_empty = bars + 3;     // [2]
  uint64_t* v_full = bars + 5;      // [2]
  uint64_t* v_empty = bars + 7;     // [2]
  uint64_t* s_full = bars + 9;      // [2 tiles]
  uint64_t* p_full = bars + 11;     // [2 tiles]
  uint64_t* pv_done = bars + 13;    // [2 tiles]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);
  float* xs = reinterpret_cast<float*>(bars + 32);  // [2 parity][2 tiles][2 halves][128 rows]: "needs the exact maximum" flags
  float* xm = xs + 1024;                            // same shape: half-row maxima (fallback path) / final row sums

  const int warp = warp_id_uniform(), lane = threadIdx.x & 31;
  const int bh = blockIdx.y;
  const int q0 = blockIdx.x * 256;
  const int n_kv = (g.Lk + 127) / 128;
  const long long row_base = static_cast<long long>(bh) * g.L;   // query rows / lse
  const long long kv_base = static_cast<long long>(bh) * g.Lk;  // key / value rows

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(q_full, 1);
      for (int s = 0; s < 2; ++s) {
        mbar_init(&k_full[s], 1);
        mbar_init(&k_empty[s], 1);
        mbar_init(&v_full[s], 1);
        mbar_init(&v_empty[s], 1);
        mbar_init(&s_full[s], 1);
        mbar_init(&p_full[s], 256);
        mbar_init(&pv_done[s], 1);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();  // everything above touched only shared / tensor memory
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, 65536);
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int qrow = static_cast<int>(row_base + q0 + t * 128);
        tma_load_2d(sQ + t * 32768, &tmQ, q_full, 0, qrow);
        tma_load_2d(sQ + t * 32768 + 16384, &tmQ, q_full, 64, qrow);
      }
      for (int j = 0; j < n_kv; ++j) {
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        const int kvrow = static_cast<int>(kv_base + j * 128);
        mbar_wait(&k_empty[s], ph ^ 1u, 30);
        mbar_arrive_expect_tx(&k_full[s], 32768);
        tma_load_2d(sK + s * 32768, &tmK, &k_full[s], 0, kvrow);
        tma_load_2d(sK + s * 32768 + 16384, &tmK, &k_full[s], 64, kvrow);
        mbar_wait(&v_empty[s], ph ^ 1u, 31);
        mbar_arrive_expect_tx(&v_full[s], 32768);
#pragma unroll
        for (int jc = 0; jc < 2; ++jc)
#pragma unroll
          for (int ih = 0; ih < 2; ++ih)
            tma_load_2d(sV + s * 32768 + (jc * 2 + ih) * 8192, &tmV, &v_full[s], jc * 64, kvrow + ih * 64);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idS = umma_idesc_bf16(128, 128, 0, 0);
    constexpr uint32_t idPV = umma_idesc_bf16(128, 128, 0, 1);
    mbar_wait(q_full, 0, 32);
    const uint64_t dQ0 = umma_desc_sw128(smem_u32(sQ), 1024, 16);
    const uint64_t dK0 = umma_desc_sw128(smem_u32(sK), 1024, 16);
    const uint64_t dV0 = umma_desc_sw128(smem_u32(sV), 1024, 16384);
    auto issue_S = [&](int t, int j) {  // S_t(j) = Q_t K_j^T into columns [128 t, 128 t + 128)
      const uint64_t dq = dQ0 + static_cast<uint64_t>(t * (32768 >> 4));
      const uint64_t dk = dK0 + static_cast<uint64_t>((j & 1) * (32768 >> 4));
      {  // one elect for the 8 k-steps (attn_common.cuh): both operands K-major, halves of 64 head-dim columns 16 KB apart
        constexpr int HH = 16384 >> 4;
        umma_bf16_ss_w_x8<2, 4, 6, HH, HH + 2, HH + 4, HH + 6, 2, 4, 6, HH, HH + 2, HH + 4, HH + 6>(
            tmem_base + static_cast<uint32_t>(t) * 128u, dq, dk, idS, 0u);
      }
      umma_commit_w(&s_full[t]);
    };
    auto issue_PV = [&](int t, int j) {  // O_t += P_t(j) V_j, A = P straight from TMEM (bf16 pairs in S_t's first 64 columns)
      const uint64_t dv = dV0 + static_cast<uint64_t>((j & 1) * (32768 >> 4));
      {  // V MN-major: k-steps 2 KB apart inside a 64-row half, halves 8 KB apart
        constexpr int VS = 2048 >> 4, VH = 8192 >> 4;
        umma_bf16_ts_w_x8<8, 16, 24, 32, 40, 48, 56, VS, 2 * VS, 3 * VS, VH, VH + VS, VH + 2 * VS, VH + 3 * VS>(
            tmem_base + 256u + static_cast<uint32_t>(t) * 128u, tmem_base + static_cast<uint32_t>(t) * 128u, dv, idPV,
            j > 0 ? 1u : 0u);
      }
    };
    mbar_wait(&k_full[0], 0, 33);
    tc_fence_after();
    issue_S(0, 0);
    issue_S(1, 0);
    umma_commit_w(&k_empty[0]);
    for (int j = 0; j < n_kv; ++j) {
      const int s = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      const bool more = j + 1 < n_kv;
      mbar_wait(&v_full[s], ph, 34);
      if (more) mbar_wait(&k_full[s ^ 1], ((j + 1) >> 1) & 1, 35);
      // ---- tile A
      mbar_wait(&p_full[0], j & 1, 36);
      tc_fence_after();
      issue_PV(0, j);
      umma_commit_w(&pv_done[0]);
      if (more) issue_S(0, j + 1);
      // ---- tile B
      mbar_wait(&p_full[1], j & 1, 37);
      tc_fence_after();
      issue_PV(1, j);
      umma_commit_w(&v_empty[s]);
      umma_commit_w(&pv_done[1]);
      if (more) {
        issue_S(1, j + 1);
        umma_commit_w(&k_empty[s ^ 1]);
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax / epilogue: two threads per row
    // softmax warps sw = warp - 2 in [0, 16): tile = sw >> 3; half h = (sw & 7) >> 2; TMEM lane quarter = warp & 3
    const int sw = warp - 2;
    const int t = sw >> 3;
    const int h = (sw & 7) >> 2;
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int qi = q0 + t * 128 + r;
    const int bar_id = 1 + t * 4 + q;  // named barrier of the two warps that share these 32 rows
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t tS = tmem_base + static_cast<uint32_t>(t) * 128u + lane_off;
    const uint32_t tO = tmem_base + 256u + static_cast<uint32_t>(t) * 128u + lane_off;
    const float c2 = g.scale * kLog2e;
    const float2 c22 = make_float2(c2, c2);
    const int tail = g.Lk & 127;
    float m_used = -INFINITY, l_sum = 0.f;
    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(&s_full[t], j & 1, 38);
      tc_fence_after();
      const bool ragged = (tail != 0) && (j == n_kv - 1);
      const int par = j & 1;
      float* my_xs = xs + ((par * 2 + t) * 2 + h) * 128 + r;
      const float* his_xs = xs + ((par * 2 + t) * 2 + (h ^ 1)) * 128 + r;
      float* my_xm = xm + ((par * 2 + t) * 2 + h) * 128 + r;
      const float* his_xm = xm + ((par * 2 + t) * 2 + (h ^ 1)) * 128 + r;
      uint32_t a[32], b[32], pk[32];
      // my 64 columns of P into registers + their partial sum; S stays intact in TMEM
      auto pass2 = [&]() -> float {
        const float2 nm2 = make_float2(-m_used, -m_used);
        float2 ls[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
        tmem_ld_32x32(tS + h * 64, a);
        tmem_ld_32x32(tS + h * 64 + 32, b);
        tmem_ld_wait();
        if (ragged) {
          mask32(a, tail - h * 64);
          mask32(b, tail - h * 64 - 32);
        }
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const float* x = reinterpret_cast<const float*>(c ? b : a);
#pragma unroll
          for (int k2 = 0; k2 < 16; ++k2) {
            const bool poly = (k2 % 4) < POLY;
            const float2 e = ffma2(make_float2(x[2 * k2], x[2 * k2 + 1]), c22, nm2);
            const float2 p2 = poly ? ex2_poly2(e) : make_float2(ex2(e.x), ex2(e.y));
            ls[k2 & 1] = fadd2(ls[k2 & 1], p2);
            pk[c * 16 + k2] = pack_bf16x2(p2.x, p2.y);
          }
        }
        return (ls[0].x + ls[0].y) + (ls[1].x + ls[1].y);
      };
      float bsum = 0.f;
      bool need = (j == 0);
      if (!need) {
        bsum = pass2();
        need = !(bsum < 1.0e9f);  // also catches inf / nan
      }
      *my_xs = need ? 1.0f : 0.0f;
      asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");  // both halves have READ their S columns and published their flag
      need = need || (*his_xs != 0.0f);                           // row-level decision: identical in both threads of the row
      if (__any_sync(0xffffffffu, need)) {                        // same 32 rows in both warps -> same branch in both warps
        float mx = -INFINITY;
        tmem_ld_32x32(tS + h * 64, a);
        tmem_ld_32x32(tS + h * 64 + 32, b);
        tmem_ld_wait();
        if (ragged) {
          mask32(a, tail - h * 64);
          mask32(b, tail - h * 64 - 32);
        }
        mx = max32(a, mx);
        mx = max32(b, mx);
        *my_xm = mx * c2;
        asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
        const float m_new = fmaxf(m_used, fmaxf(mx * c2, *his_xm));
        const bool grow = need && (m_new > m_used);
        const float f = grow ? ex2(m_used - m_new) : 1.0f;  // first block: ex2(-inf) = 0
        if (grow) {
          m_used = m_new;
          l_sum *= f;
        }
        if (j > 0) {  // O_t must be stable: P V_t(j-1) is the newest MMA accumulating into it
          mbar_wait(&pv_done[t], (j - 1) & 1, 39);
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            uint32_t o[32];
            tmem_ld_32x32(tO + h * 64 + c * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
            tmem_st_32x32(tO + h * 64 + c * 32, o);
          }
          tmem_st_wait();
        }
        bsum = pass2();
        asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");  // the partner has re-read its S columns: P may overwrite them
      }
      l_sum += bsum;
      // my 64 K-elements of P = 32-bit columns [32 h, 32 h + 32) of the S region
      tmem_st_32x16(tS + h * 32, *reinterpret_cast<const uint32_t(*)[16]>(&pk[0]));
      tmem_st_32x16(tS + h * 32 + 16, *reinterpret_cast<const uint32_t(*)[16]>(&pk[16]));
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[t]);
    }
    // combine the two half-row sums (m_used is identical in both threads of a row)
    {
      const int par = n_kv & 1;
      float* my_xm = xm + ((par * 2 + t) * 2 + h) * 128 + r;
      const float* his_xm = xm + ((par * 2 + t) * 2 + (h ^ 1)) * 128 + r;
      *my_xm = l_sum;
      asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
      l_sum += *his_xm;
    }
    mbar_wait(&pv_done[t], (n_kv - 1) & 1, 40);
    tc_fence_after();
    const float inv = 1.0f / l_sum;
    const bool live = qi < g.L;
    bf16* orow = nullptr;
    if (live) {
      const int bb = bh / g.H, hh = bh % g.H;
      if (qi < g.split)
        orow = g.o0 + (static_cast<size_t>(bb) * g.split + qi) * g.ld0 + hh * 128 + h * 64;
      else
        orow = g.o1 + (static_cast<size_t>(bb) * (g.L - g.split) + (qi - g.split)) * g.ld1 + hh * 128 + h * 64;
      if (h == 0) g.lse[row_base + qi] = (m_used + log2f(l_sum)) * kLn2;
    }
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {  // tcgen05.ld is warp-collective: every lane loads, only live rows store
      uint32_t o[32];
      tmem_ld_32x32(tO + h * 64 + c * 32, o);
      tmem_ld_wait();
      if (live) {
#pragma unroll
        for (int k8 = 0; k8 < 4; ++k8) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(o[k8 * 8 + 0]) * inv, __uint_as_float(o[k8 * 8 + 1]) * inv);
          u.y = pack_bf16x2(__uint_as_float(o[k8 * 8 + 2]) * inv, __uint_as_float(o[k8 * 8 + 3]) * inv);
          u.z = pack_bf16x2(__uint_as_float(o[k8 * 8 + 4]) * inv, __uint_as_float(o[k8 * 8 + 5]) * inv);
          u.w = pack_bf16x2(__uint_as_float(o[k8 * 8 + 6]) * inv, __uint_as_float(o[k8 * 8 + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + c * 32 + k8 * 8) = u;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

template <int POLY>
static int launch_pp(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnFwdArgs& a,
                     cudaStream_t stream) {
  auto kern = attn_fwd_pp_kernel<POLY>;
  static bool configured = false;
  if (!configured) {
    B200_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kPPSmem));
    configured = true;
  }
  dim3 grid((a.L + 255) / 256, a.B * a.H);
  B200_KLAUNCH(kern, grid, kPPThreads, kPPSmem, stream, tq, tk, tv, a);
  B200_CUDA_CHECK(cudaGetLastError());
  return B200_OK;
}

template <int POLY, int DBG = 0>
static int launch_pp2(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnFwdArgs& a,
                      cudaStream_t stream) {
  auto kern = attn_fwd_pp2_kernel<POLY, DBG>;
  static bool configured = false;
  if (!configured) {
    B200_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kPPSmem));
    configured = true;
  }
  dim3 grid((a.L + 255) / 256, a.B * a.H);
  B200_KLAUNCH(kern, grid, kPP2Threads, kPPSmem, stream, tq, tk, tv, a);
  B200_CUDA_CHECK(cudaGetLastError());
  return B200_OK;
}

template <int POLY>
static int launch_pp3(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnFwdArgs& a,
                      cudaStream_t stream) {
  auto kern = attn_fwd_pp3_kernel<POLY>;
  static bool configured = false;
  if (!configured) {
    B200_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kPP3Smem));
    configured = true;
  }
  dim3 grid((a.L + 255) / 256, a.B * a.H);
  B200_KLAUNCH(kern, grid, kPP3Threads, kPP3Smem, stream, tq, tk, tv, a);
  B200_CUDA_CHECK(cudaGetLastError());
  return B200_OK;
}

// variant 7: optimistic ping-pong, two threads per row
int attn_fwd_pp3_launch(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnFwdArgs& a,
                        cudaStream_t stream) {
  static int poly = -1;
  if (poly < 0) {
    const char* e = getenv("B200_ATTN_PP_POLY");
    poly = (e && atoi(e) >= 0 && atoi(e) <= 3) ? atoi(e) : 0;
  }
  switch (poly) {
    case 1: return launch_pp3<1>(tq, tk, tv, a, stream);
    case 2: return launch_pp3<2>(tq, tk, tv, a, stream);
    default: return launch_pp3<0>(tq, tk, tv, a, stream);
  }
}

// variant 6: the ping-pong kernel with the OPTIMISTIC running maximum (no row-max pass on the common path)
int attn_fwd_pp2_launch(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnFwdArgs& a,
                        cudaStream_t stream) {
  static int poly = -1;
  if (poly < 0) {  // measured (gpurun_out/r2_pp2.log): POLY 0 209.9 us, 1 216.5 us, 2 221.5 us -- without the max pass the
    const char* e = getenv("B200_ATTN_PP_POLY");  // softmax is issue-bound, and the polynomial costs 5x the issue slots of MUFU
    poly = (e && atoi(e) >= 0 && atoi(e) <= 3) ? atoi(e) : 0;
  }
  static int dbg = -1;
  if (dbg < 0) {
    const char* e = getenv("B200_ATTN_PP_DBG");
    dbg = e ? atoi(e) : 0;
  }
  if (dbg == 1) return launch_pp2<0, 1>(tq, tk, tv, a, stream);
  if (dbg == 2) return launch_pp2<0, 2>(tq, tk, tv, a, stream);
  switch (poly) {
    case 0: return launch_pp2<0>(tq, tk, tv, a, stream);
    case 2: return launch_pp2<2>(tq, tk, tv, a, stream);
    case 3: return launch_pp2<3>(tq, tk, tv, a, stream);
    default: return launch_pp2<1>(tq, tk, tv, a, stream);
  }
}

int attn_fwd_pp_launch(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnFwdArgs& a,
                       cudaStream_t stream) {
  static int poly = -1;
  if (poly < 0) {
    const char* e = getenv("B200_ATTN_PP_POLY");
    poly = (e && atoi(e) >= 0 && atoi(e) <= 3) ? atoi(e) : 1;
  }
  switch (poly) {
    case 0: return launch_pp<0>(tq, tk, tv, a, stream);
    case 2: return launch_pp<2>(tq, tk, tv, a, stream);
    case 3: return launch_pp<3>(tq, tk, tv, a, stream);
    default: return launch_pp<1>(tq, tk, tv, a, stream);
  }
}

}  // namespace b200
