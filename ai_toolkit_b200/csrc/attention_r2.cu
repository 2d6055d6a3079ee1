// Round-2 CANDIDATE attention kernels (opt-in: B200_ATTN_FWD=3|4, B200_ATTN_BWD=2|3|4|5).  They compile and their SASS
// has been read, but they have NOT run on a B200 yet; the default path stays the validated kernels of attention.cu
// until tools/r2_attn_trip.sh has shown parity and a speed-up on the GPU box.
//
// What the round-1 profiles (profiles/r1_attn_*_source_hotspots.md, warp-state samples per SASS line) showed:
//   forward   the 8 softmax warps are never idle (barrier waits < 10 % of their samples) while the tensor pipe is 45 %
//             active: the kernel is bound by the ISSUE RATE of the softmax (522 instructions per thread and 128 x 128
//             tile = 8.2 per element, two warps per scheduler issuing 48 % of the cycles).
//   backward  each softmax group is a serial chain  S/dP MMA -> softmax -> dV/dK MMA -> next S/dP MMA  (its TMEM buffer is
//             reused), so two tiles take T_softmax + 1280 cycles; T_softmax ~ 2700 cycles, 41 % of it stalled on
//             long-scoreboard loads: generic-address shared-memory loads of the column statistics (the 1024-byte
//             alignment through uintptr_t loses the shared address space), global loads consumed immediately by the
//             prefetch, local-memory loads of indexed TMEM-address arrays.
// Candidates:
//   fwd  variant 3 (NT = 2) / 4 (NT = 4 threads per row, 16 softmax warps):
//        * exponent argument as ONE packed FFMA2 per column pair (s c - m), row sum with FADD2, max over raw scores;
//        * one "P V done" barrier per P buffer: softmax(j) waits for P V(j-2) before storing P(j) and for P V(j-1)
//          only when a row of its warp really rescales O, so the exponentials of tile j+1 overlap P V(j);
//        * shared-memory pointers stay in the shared address space (LDS / STS), no local-memory arrays.
//   bwd  variant 3 (NH = 1, same 8 softmax warps) / 2 (NH = 2 threads per row, 16 softmax warps):
//        * LDS for the column statistics, raw prefetch (scaled when stored), arithmetic TMEM addresses;
//        * S / dP read in 16-column chunks, the next chunk in flight while the current one is computed;
//        * NH = 2: thread (row, hh) owns columns [32 hh, 32 hh + 32) and writes its bf16 P / dS pairs inside its own
//          range (columns 48 hh ...), so no thread overwrites what another still reads.
//   bwd  variant 4 (NH = 2) / 5 (NH = 1): the dK/dV pass of variant 2 / 3 plus a dQ pass whose next S / dP MMA is
//        issued as soon as the softmax has LOADED the current one (dS in its own TMEM region): see attn_bwd_q2_kernel.
#include <type_traits>

#include "attn_common.cuh"

namespace b200 {

#ifndef B200_ATTN_POLY_R2
#define B200_ATTN_POLY_R2 1
#endif
constexpr int kPolyPairsOf4R2 = B200_ATTN_POLY_R2;  // of every 4 column pairs, how many use ex2_poly (0 = all MUFU)

__device__ __forceinline__ uint8_t* smem_align1024(uint8_t* raw) {
  // offset arithmetic (not uintptr_t rounding) keeps the pointer in the shared address space
  return raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
}
template <int N>
__device__ __forceinline__ void named_bar_sync(int id) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(N) : "memory");
}

// =================================================================================================
// forward: NT threads per row
// =================================================================================================
template <int NT>
struct FwdR2 {
  static constexpr int kThreads = 64 + 128 * NT;
  static constexpr int kCW = 128 / NT;                 // S / O columns per thread
  static constexpr int kBarBytes = 24 * 8;             // 18 barriers / slots used
  static constexpr int kXch = 2 * NT * 128;            // floats: [2 parity][NT][128 rows]
  static constexpr int kSmem = 1024 + 5 * 32768 + kBarBytes + kXch * 4;
};

template <int NT>
__global__ void __launch_bounds__(FwdR2<NT>::kThreads, 1)
attn_fwd_nt_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const AttnFwdArgs g) {
  pdl_launch_dependents();
  using C = FwdR2<NT>;
  constexpr int CW = C::kCW;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + 32768;
  uint8_t* sV = sK + 2 * 32768;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 2 * 32768);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = bars + 3;
  uint64_t* v_full = bars + 5;
  uint64_t* v_empty = bars + 7;
  uint64_t* s_full = bars + 9;
  uint64_t* s_empty = bars + 11;
  uint64_t* p_full = bars + 13;   // [2]: one per P buffer (a single barrier could be lapped now that softmax(j+1)
                                  //      no longer waits for P V(j): the MMA warp's parity wait must stay <= 1 phase behind)
  uint64_t* pv_done = bars + 15;  // [2]: one per P buffer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);
  float* xch = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + C::kBarBytes);

  const int warp = warp_id_uniform(), lane = threadIdx.x & 31;
  const int bh = blockIdx.y;
  const int q0 = blockIdx.x * 128;
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
        mbar_init(&s_empty[s], 128 * NT);
        mbar_init(&pv_done[s], 1);
        mbar_init(&p_full[s], 128 * NT);
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
  const uint32_t tO = tmem_base + 256u;
  auto S_at = [&](int b) -> uint32_t { return tmem_base + static_cast<uint32_t>(b) * 128u; };
  auto P_at = [&](int b) -> uint32_t { return tmem_base + 384u + static_cast<uint32_t>(b) * 64u; };  // bf16 pairs

  if (warp == 0) {
    if (lane == 0) {
      const int qrow = static_cast<int>(row_base + q0);
      mbar_arrive_expect_tx(q_full, 32768);
      tma_load_2d(sQ, &tmQ, q_full, 0, qrow);
      tma_load_2d(sQ + 16384, &tmQ, q_full, 64, qrow);
      for (int j = 0; j < n_kv; ++j) {
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        const int kvrow = static_cast<int>(kv_base + j * 128);
        mbar_wait(&k_empty[s], ph ^ 1u, 10);
        mbar_arrive_expect_tx(&k_full[s], 32768);
        tma_load_2d(sK + s * 32768, &tmK, &k_full[s], 0, kvrow);
        tma_load_2d(sK + s * 32768 + 16384, &tmK, &k_full[s], 64, kvrow);
        mbar_wait(&v_empty[s], ph ^ 1u, 11);
        mbar_arrive_expect_tx(&v_full[s], 32768);
#pragma unroll
        for (int jc = 0; jc < 2; ++jc)
#pragma unroll
          for (int ih = 0; ih < 2; ++ih)
            tma_load_2d(sV + s * 32768 + (jc * 2 + ih) * 8192, &tmV, &v_full[s], jc * 64, kvrow + ih * 64);
      }
    }
  } else if (warp == 1) {
    // converged MMA warp: one elected lane issues each tcgen05 instruction (common.cuh)
    constexpr uint32_t idS = umma_idesc_bf16(128, 128, 0, 0);
    constexpr uint32_t idPV = umma_idesc_bf16(128, 128, 0, 1);
    mbar_wait(q_full, 0, 12);
    const uint64_t dQ0 = umma_desc_sw128(smem_u32(sQ), 1024, 16);
    const uint64_t dK0 = umma_desc_sw128(smem_u32(sK), 1024, 16);
    const uint64_t dV0 = umma_desc_sw128(smem_u32(sV), 1024, 16384);
    auto issue_S = [&](int j) {
      const int s = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      mbar_wait(&k_full[s], ph, 13);
      mbar_wait(&s_empty[s], ph ^ 1u, 14);
      tc_fence_after();
      const uint64_t dk = dK0 + static_cast<uint64_t>(s * (32768 >> 4));
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const uint32_t off = (kk >> 2) * (16384 >> 4) + 2u * (kk & 3);
        umma_bf16_ss_w(S_at(s), dQ0 + off, dk + off, idS, kk > 0 ? 1u : 0u);
      }
      umma_commit_w(&k_empty[s]);
      umma_commit_w(&s_full[s]);
    };
    auto issue_PV = [&](int j) {
      const int s = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      mbar_wait(&v_full[s], ph, 15);
      mbar_wait(&p_full[s], ph, 16);
      tc_fence_after();
      const uint64_t dv = dV0 + static_cast<uint64_t>(s * (32768 >> 4));
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const uint32_t off = (kk >> 2) * (8192 >> 4) + (kk & 3) * (2048 >> 4);
        umma_bf16_ts_w(tO, P_at(s) + kk * 8, dv + off, idPV, (j > 0 || kk > 0) ? 1u : 0u);  // A = P straight from TMEM
      }
      umma_commit_w(&v_empty[s]);
      umma_commit_w(&pv_done[s]);
    };
    issue_S(0);
    for (int j = 0; j < n_kv; ++j) {
      if (j + 1 < n_kv) issue_S(j + 1);
      issue_PV(j);
    }
  } else {
    // ---------------------------------------------------------------- softmax / epilogue
    // NT threads per row: thread (r, h) owns columns [CW h, CW h + CW) of the S / P / O row r
    const int q = warp & 3;
    const int h = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const int qi = q0 + r;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const float c2 = g.scale * kLog2e;
    const float2 c22 = make_float2(c2, c2);
    float m_used = -INFINITY, l_sum = 0.f;
    for (int j = 0; j < n_kv; ++j) {
      const int s = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      mbar_wait(&s_full[s], ph, 17);
      tc_fence_after();
      uint32_t v[CW];
#pragma unroll
      for (int c = 0; c < CW / 32; ++c)
        tmem_ld_32x32(S_at(s) + lane_off + h * CW + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&v[c * 32]));
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&s_empty[s]);  // S is in registers: the tensor core may overwrite this buffer
      const int nvalid = g.Lk - (j * 128 + h * CW);  // columns of my part that exist
      if (nvalid < CW) {  // ragged last tile only: missing columns become -inf IN PLACE (no second copy of the row)
#pragma unroll
        for (int i = 0; i < CW; ++i)
          if (i >= nvalid) v[i] = 0xff800000u;
      }
      const float* x = reinterpret_cast<const float*>(v);  // RAW scores; scaled inside the FFMA2 below
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < CW; ++i) mx = fmaxf(mx, x[i]);
      mx *= c2;  // c2 > 0: max and rounding commute
      float* xm = xch + (j & 1) * (NT * 128);
      xm[h * 128 + r] = mx;
      named_bar_sync<32 * NT>(1 + q);  // only the NT warps that share these 32 rows (one named barrier per lane quarter)
      float m_new = m_used;
#pragma unroll
      for (int t = 0; t < NT; ++t) m_new = fmaxf(m_new, xm[t * 128 + r]);
      const bool need = (m_new > m_used + 8.0f);  // also true for the first tile (m_used = -inf)
      const bool any_need = __any_sync(0xffffffffu, need);
      if (any_need) {
        const float f = need ? ex2(m_used - m_new) : 1.0f;  // first tile: ex2(-inf) = 0
        if (need) {
          m_used = m_new;
          l_sum *= f;
        }
        if (j > 0) {
          // O must be stable: P V(j-1) is the newest MMA accumulating into it (P V(j) needs the P written below).
          // At most one phase behind: softmax(j-1) waited for P V(j-3) on this barrier before storing its P.
          mbar_wait(&pv_done[s ^ 1], ((j - 1) >> 1) & 1, 20);
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < CW / 32; ++c) {
            uint32_t o[32];
            tmem_ld_32x32(tO + lane_off + h * CW + c * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
            tmem_st_32x32(tO + lane_off + h * CW + c * 32, o);
          }
          tmem_st_wait();
        }
      }
      // P = exp2(S c - m_used) -> bf16 pairs -> TMEM (the A operand of the P V MMA; no shared-memory round trip)
      uint32_t pk[CW / 2];
      const float2 nm2 = make_float2(-m_used, -m_used);
      float2 ls[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};  // independent packed partial sums
#pragma unroll
      for (int k2 = 0; k2 < CW / 2; ++k2) {
        const bool poly = (k2 % 4) < kPolyPairsOf4R2;
        const float2 a = ffma2(make_float2(x[2 * k2], x[2 * k2 + 1]), c22, nm2);
        const float2 p2 = poly ? ex2_poly2(a) : make_float2(ex2(a.x), ex2(a.y));
        ls[k2 & 1] = fadd2(ls[k2 & 1], p2);
        pk[k2] = pack_bf16x2(p2.x, p2.y);
      }
      l_sum += (ls[0].x + ls[0].y) + (ls[1].x + ls[1].y);
      if (j >= 2) {  // P V(j-2) has finished reading this P buffer (one phase behind at most: see softmax(j-2))
        mbar_wait(&pv_done[s], ((j >> 1) - 1) & 1, 18);
        tc_fence_after();
      }
      if constexpr (CW == 64)
        tmem_st_32x32(P_at(s) + lane_off + h * 32, *reinterpret_cast<const uint32_t(*)[32]>(&pk[0]));
      else
        tmem_st_32x16(P_at(s) + lane_off + h * 16, *reinterpret_cast<const uint32_t(*)[16]>(&pk[0]));
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[s]);
    }
    // combine the partial row sums
    float* xl = xch + (n_kv & 1) * (NT * 128);
    xl[h * 128 + r] = l_sum;
    named_bar_sync<32 * NT>(1 + q);
    l_sum = 0.f;
#pragma unroll
    for (int t = 0; t < NT; ++t) l_sum += xl[t * 128 + r];
    mbar_wait(&pv_done[(n_kv - 1) & 1], ((n_kv - 1) >> 1) & 1, 19);
    tc_fence_after();
    const float inv = 1.0f / l_sum;
    const bool live = qi < g.L;
    bf16* orow = nullptr;
    if (live) {
      const int b = bh / g.H, hh = bh % g.H;
      if (qi < g.split)
        orow = g.o0 + (static_cast<size_t>(b) * g.split + qi) * g.ld0 + hh * 128 + h * CW;
      else
        orow = g.o1 + (static_cast<size_t>(b) * (g.L - g.split) + (qi - g.split)) * g.ld1 + hh * 128 + h * CW;
      if (h == 0) g.lse[row_base + qi] = (m_used + log2f(l_sum)) * kLn2;
    }
#pragma unroll 1
    for (int c = 0; c < CW / 32; ++c) {  // tcgen05.ld is warp-collective: every lane loads, only live rows store
      uint32_t o[32];
      tmem_ld_32x32(tO + lane_off + h * CW + c * 32, o);
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

// =================================================================================================
// backward: NH threads per row
// =================================================================================================
constexpr int kBwdR2Stages = 3;
constexpr int kBwdR2Smem = 1024 + 2 * 32768 + kBwdR2Stages * 32768 + 16 * 8 + 16 + 8 * 128 * 4;

// MODE_KV = 1: stationary (R0, R1) = (K_j, V_j), streamed (T0, T1) = (Q_i, dO_i); outputs dV (acc0), dK (acc1)
// MODE_KV = 0: stationary (R0, R1) = (Q_i, dO_i) held in TMEM, streamed (T0, T1) = (K_j, V_j); output dQ (acc0)
template <int MODE_KV, int NH>
__global__ void __launch_bounds__(64 + 256 * NH, 1)
attn_bwd_r2_kernel(const __grid_constant__ CUtensorMap tmR0, const __grid_constant__ CUtensorMap tmR1,
                   const __grid_constant__ CUtensorMap tmT0, const __grid_constant__ CUtensorMap tmT1, const AttnBwdArgs g) {
  pdl_launch_dependents();
  constexpr int kStages = kBwdR2Stages;
  constexpr int CW = 64 / NH;       // S / dP columns per thread
  constexpr int NCH = CW / 16;      // 16-column chunks per thread
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint8_t* sR0 = smem;
  uint8_t* sR1 = sR0 + 32768;
  uint8_t* sT = sR1 + 32768;  // stage st: T0 at sT + st*32768, T1 at +16384
  uint64_t* bars = reinterpret_cast<uint64_t*>(sT + kStages * 32768);
  uint64_t* r_full = bars;
  uint64_t* t_full = bars + 1;
  uint64_t* t_empty = bars + 4;
  uint64_t* x_full = bars + 7;
  uint64_t* r_tmem = bars + 9;    // MODE_Q: the stationary operands are in TMEM
  uint64_t* pb_full = bars + 11;  // [2]: one per softmax group / TMEM buffer
  uint64_t* done_bar = bars + 13;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
  float* colws = reinterpret_cast<float*>(bars + 16);  // per softmax warp: CW x lse2, CW x delta*scale of its columns

  const int warp = warp_id_uniform(), lane = threadIdx.x & 31;
  const int bh = blockIdx.y;
  const int r0 = blockIdx.x * 128;
  const int n_t = (g.Lt + 63) / 64;                               // streamed side
  const long long row_base = static_cast<long long>(bh) * g.L;   // stationary side (and its outputs)
  const long long t_base = static_cast<long long>(bh) * g.Lt;    // streamed side

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmR0);
    tma_prefetch_desc(&tmR1);
    tma_prefetch_desc(&tmT0);
    tma_prefetch_desc(&tmT1);
  }
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(r_full, 1);
      for (int s = 0; s < kStages; ++s) {
        mbar_init(&t_full[s], 1);
        mbar_init(&t_empty[s], 1);
      }
      mbar_init(&x_full[0], 1);
      mbar_init(&x_full[1], 1);
      mbar_init(r_tmem, 256 * NH);
      mbar_init(&pb_full[0], 128 * NH);
      mbar_init(&pb_full[1], 128 * NH);
      mbar_init(done_bar, 1);
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
  auto X0 = [&](int b) -> uint32_t { return tmem_base + static_cast<uint32_t>(b) * 64u; };          // S  -> P
  auto X1 = [&](int b) -> uint32_t { return tmem_base + 128u + static_cast<uint32_t>(b) * 64u; };   // dP -> dS
  const uint32_t tA0 = tmem_base + 256u, tA1 = tmem_base + 384u;
  const uint32_t tR0 = tmem_base + 384u, tR1 = tmem_base + 448u;  // MODE_Q: Q_i / dO_i as bf16 pairs (one accumulator only)
  // column (inside X) of the K slice kk of the packed bf16 row: NH = 1 packs into columns 0..31, NH = 2 into 0..15 | 48..63
  auto pcol = [](int kk) -> uint32_t { return NH == 1 ? kk * 8 : (kk < 2 ? kk * 8 : 48 + (kk - 2) * 8); };

  if (warp == 0) {
    if (lane == 0) {
      const int rrow = static_cast<int>(row_base + r0);
      if (MODE_KV) {
        mbar_arrive_expect_tx(r_full, 65536);
        tma_load_2d(sR0, &tmR0, r_full, 0, rrow);
        tma_load_2d(sR0 + 16384, &tmR0, r_full, 64, rrow);
        tma_load_2d(sR1, &tmR1, r_full, 0, rrow);
        tma_load_2d(sR1 + 16384, &tmR1, r_full, 64, rrow);
      }
      for (int i = 0; i < n_t; ++i) {
        const int st = i % kStages;
        const uint32_t ph = (i / kStages) & 1;
        const int trow = static_cast<int>(t_base + i * 64);
        mbar_wait(&t_empty[st], ph ^ 1u, 20000 + i);
        mbar_arrive_expect_tx(&t_full[st], 32768);
        uint8_t* d = sT + st * 32768;
        tma_load_2d(d, &tmT0, &t_full[st], 0, trow);
        tma_load_2d(d + 8192, &tmT0, &t_full[st], 64, trow);
        tma_load_2d(d + 16384, &tmT1, &t_full[st], 0, trow);
        tma_load_2d(d + 16384 + 8192, &tmT1, &t_full[st], 64, trow);
      }
    }
  } else if (warp == 1) {
    // converged MMA warp, elected issue (common.cuh)
    constexpr uint32_t idA = umma_idesc_bf16(128, 64, 0, 0);
    constexpr uint32_t idB128 = umma_idesc_bf16(128, 128, 0, 1);
    constexpr uint32_t idB64 = umma_idesc_bf16(128, 64, 0, 1);  // N = 64: head-dim columns 0..63 (first MN-major atom of the streamed tile)
    const bool full_d = g.dlive != 64;  // 64: the second half of the head dim is zero padding -> contractions / outputs over it skipped
    const uint32_t idB = full_d ? idB128 : idB64;
    if (MODE_KV)
      mbar_wait(r_full, 0, 21);
    else
      mbar_wait(r_tmem, 0, 21);  // the softmax warps have stored Q_i / dO_i into TMEM
    tc_fence_after();
    const uint64_t dR0 = umma_desc_sw128(smem_u32(sR0), 1024, 16);
    const uint64_t dR1 = umma_desc_sw128(smem_u32(sR1), 1024, 16);
    const uint64_t dTk = umma_desc_sw128(smem_u32(sT), 1024, 16);    // streamed tiles read K-major (phase A)
    const uint64_t dTm = umma_desc_sw128(smem_u32(sT), 1024, 8192);  // the same bytes read MN-major (phase B)
    auto issue_A = [&](int i) {
      const int st = i % kStages;
      const uint32_t ph = (i / kStages) & 1;
      const int xb = i & 1;
      mbar_wait(&t_full[st], ph, 22000 + i);
      // X[xb] was last read by B(i-2) (P / dS alias it); MMAs of one thread execute in issue order, so no barrier
      tc_fence_after();
      const uint64_t d0 = dTk + static_cast<uint64_t>(st * (32768 >> 4));
      const uint64_t d1 = d0 + (16384 >> 4);
      if (g.dbg != 3) {
        // offsets of the k-steps: A K-major [128 x 128] as two 64-column halves 16 KB apart, B [64 x 128] halves 8 KB apart
        constexpr int AH = 16384 >> 4, BH = 8192 >> 4;
        if (!full_d) {  // contraction over head-dim columns 0..63 only
          if (MODE_KV) {
            umma_bf16_ss_w_x4<2, 4, 6, 2, 4, 6>(X0(xb), dR0, d0, idA, 0u);
            umma_bf16_ss_w_x4<2, 4, 6, 2, 4, 6>(X1(xb), dR1, d1, idA, 0u);
          } else {
            umma_bf16_ts_w_x4<8, 16, 24, 2, 4, 6>(X0(xb), tR0, d0, idA, 0u);
            umma_bf16_ts_w_x4<8, 16, 24, 2, 4, 6>(X1(xb), tR1, d1, idA, 0u);
          }
        } else if (MODE_KV) {
          umma_bf16_ss_w_x8<2, 4, 6, AH, AH + 2, AH + 4, AH + 6, 2, 4, 6, BH, BH + 2, BH + 4, BH + 6>(X0(xb), dR0, d0, idA, 0u);
          umma_bf16_ss_w_x8<2, 4, 6, AH, AH + 2, AH + 4, AH + 6, 2, 4, 6, BH, BH + 2, BH + 4, BH + 6>(X1(xb), dR1, d1, idA, 0u);
        } else {
          umma_bf16_ts_w_x8<8, 16, 24, 32, 40, 48, 56, 2, 4, 6, BH, BH + 2, BH + 4, BH + 6>(X0(xb), tR0, d0, idA, 0u);
          umma_bf16_ts_w_x8<8, 16, 24, 32, 40, 48, 56, 2, 4, 6, BH, BH + 2, BH + 4, BH + 6>(X1(xb), tR1, d1, idA, 0u);
        }
      } else {
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const uint32_t offa = (kk >> 2) * (16384 >> 4) + 2u * (kk & 3);
        const uint32_t offb = (kk >> 2) * (8192 >> 4) + 2u * (kk & 3);
        if (MODE_KV)
          umma_bf16_ss_w(X0(xb), dR0 + offa, d0 + offb, idA, kk > 0 ? 1u : 0u);
        else
          umma_bf16_ts_w(X0(xb), tR0 + kk * 8, d0 + offb, idA, kk > 0 ? 1u : 0u);
      }
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const uint32_t offa = (kk >> 2) * (16384 >> 4) + 2u * (kk & 3);
        const uint32_t offb = (kk >> 2) * (8192 >> 4) + 2u * (kk & 3);
        if (MODE_KV)
          umma_bf16_ss_w(X1(xb), dR1 + offa, d1 + offb, idA, kk > 0 ? 1u : 0u);
        else
          umma_bf16_ts_w(X1(xb), tR1 + kk * 8, d1 + offb, idA, kk > 0 ? 1u : 0u);
      }
      }
      umma_commit_w(&x_full[xb]);
    };
    auto issue_B = [&](int i) {
      const int st = i % kStages;
      mbar_wait(&pb_full[i & 1], (i >> 1) & 1, 24000 + i + 100000 * MODE_KV);
      tc_fence_after();
      const uint64_t m0 = dTm + static_cast<uint64_t>(st * (32768 >> 4));  // T0 tile, MN-major view
      const uint64_t m1 = m0 + (16384 >> 4);                                // T1 tile
      const uint32_t acc = i > 0 ? 1u : 0u;
      const int xb = i & 1;
      if (g.dbg != 3) {
        constexpr int MS = 2048 >> 4;                       // MN-major k-steps of the streamed tile are 2 KB apart
        constexpr int P1 = NH == 1 ? 8 : 8, P2 = NH == 1 ? 16 : 48, P3 = NH == 1 ? 24 : 56;  // pcol(kk)
        if (MODE_KV) {
          umma_bf16_ts_w_x4<P1, P2, P3, MS, 2 * MS, 3 * MS>(tA0, X0(xb), m1, idB, acc);  // dV += P^T dO_i
          umma_bf16_ts_w_x4<P1, P2, P3, MS, 2 * MS, 3 * MS>(tA1, X1(xb), m0, idB, acc);  // dK += dS^T Q_i
        } else {
          umma_bf16_ts_w_x4<P1, P2, P3, MS, 2 * MS, 3 * MS>(tA0, X1(xb), m0, idB, acc);  // dQ += dS K_j
        }
      } else {
        if (MODE_KV) {
  #pragma unroll
          for (int kk = 0; kk < 4; ++kk)  // dV += P^T dO_i
            umma_bf16_ts_w(tA0, X0(xb) + pcol(kk), m1 + kk * (2048 >> 4), idB, (kk > 0) ? 1u : acc);
  #pragma unroll
          for (int kk = 0; kk < 4; ++kk)  // dK += dS^T Q_i
            umma_bf16_ts_w(tA1, X1(xb) + pcol(kk), m0 + kk * (2048 >> 4), idB, (kk > 0) ? 1u : acc);
        } else {
  #pragma unroll
          for (int kk = 0; kk < 4; ++kk)  // dQ += dS K_j
            umma_bf16_ts_w(tA0, X1(xb) + pcol(kk), m0 + kk * (2048 >> 4), idB, (kk > 0) ? 1u : acc);
        }
      }
      umma_commit_w(&t_empty[st]);
    };
    issue_A(0);
    for (int i = 0; i < n_t; ++i) {
      if (i + 1 < n_t) issue_A(i + 1);
      issue_B(i);
    }
    umma_commit_w(done_bar);
  } else {
    // Two softmax groups of 4 NH warps: group gq owns the tiles i == gq (mod 2), i.e. always the TMEM buffer X[gq];
    // thread (row r, part hh) owns columns [CW hh, CW hh + CW) of the 64-column tile.  A group cannot lap: its next
    // x_full needs its own previous pb_full.
    const int q = warp & 3;
    const int gq = ((warp - 2) >> 2) & 1;
    const int hh = (warp - 2) >> 3;  // 0 for NH = 1
    const int r = q * 32 + lane;
    const int ri = r0 + r;  // kv index (MODE_KV) or q index
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const float c2 = g.scale * kLog2e;
    // lse / delta are indexed by QUERY position: the streamed side of the dK/dV pass, the stationary side of the dQ pass
    const float* lse_bh = g.lse + (MODE_KV ? t_base : row_base);
    const float* delta_bh = g.delta + (MODE_KV ? t_base : row_base);
    float* myws = colws + (warp - 2) * (2 * CW);  // per warp: CW x lse2, CW x delta*scale
    float my_lse2 = 0.f, my_dls = 0.f;
    if (!MODE_KV && ri < g.L) {
      my_lse2 = lse_bh[ri] * kLog2e;
      my_dls = delta_bh[ri] * g.scale;
    }
    if (!MODE_KV) {
      // group 0 stores Q_i, group 1 dO_i: thread (r, hh) stores bf16 [128 hh / NH, +128 / NH) of its row as they lie
      // in memory = TMEM columns [64 hh / NH, +64 / NH)
      constexpr int HB = 128 / NH;  // bf16 per thread
      const bf16* src = (gq == 0 ? g.r0 : g.r1) + (row_base + ri) * 128 + hh * HB;
      const uint32_t dst = (gq == 0 ? tR0 : tR1) + lane_off + hh * (HB / 2);
#pragma unroll
      for (int c = 0; c < HB / 64; ++c) {
        uint32_t w[32];
#pragma unroll
        for (int k4 = 0; k4 < 8; ++k4) {
          uint4 u = make_uint4(0u, 0u, 0u, 0u);
          if (ri < g.L) u = *reinterpret_cast<const uint4*>(src + c * 64 + k4 * 8);
          w[k4 * 4 + 0] = u.x;
          w[k4 * 4 + 1] = u.y;
          w[k4 * 4 + 2] = u.z;
          w[k4 * 4 + 3] = u.w;
        }
        tmem_st_32x32(dst + c * 32, w);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(r_tmem);
    }
    // software prefetch of the per-column statistics (MODE_KV): RAW values for the group's next tile; they are scaled
    // when stored to shared memory one tile later, so the global-load latency is really hidden
    constexpr int NPF = CW / 32;
    float nl[NPF], nd[NPF];
#pragma unroll
    for (int k = 0; k < NPF; ++k) nl[k] = nd[k] = 0.f;
    auto fetch_cols = [&](int i) {
#pragma unroll
      for (int k = 0; k < NPF; ++k) {
        const int c = i * 64 + hh * CW + k * 32 + lane;
        nl[k] = c < g.Lt ? lse_bh[c] : 0.f;
        nd[k] = c < g.Lt ? delta_bh[c] : 0.f;
      }
    };
    if (MODE_KV && gq < n_t) fetch_cols(gq);
    for (int i = gq; i < n_t; i += 2) {
      const uint32_t xph = (i >> 1) & 1;
      if (MODE_KV) {
        __syncwarp();
#pragma unroll
        for (int k = 0; k < NPF; ++k) {
          myws[k * 32 + lane] = nl[k] * kLog2e;
          myws[CW + k * 32 + lane] = nd[k] * g.scale;
        }
        __syncwarp();
        if (i + 2 < n_t) fetch_cols(i + 2);
      }
      mbar_wait(&x_full[gq], xph, 25000 + i + 100000 * MODE_KV);
      tc_fence_after();
      if (g.dbg == 2) {  // timing experiment: the MMA / barrier pipeline without any softmax work
        tc_fence_before();
        mbar_arrive(&pb_full[gq]);
        continue;
      }
      const int nvalid = g.Lt - (i * 64 + hh * CW);  // columns of my part that exist (<= 0: none)
      uint32_t pp[CW / 2], dd[CW / 2];             // packed bf16 pairs of my columns: P and dS
      uint32_t sv[2][16], dv[2][16];               // two chunks in flight (ping-pong)
      const uint32_t cs = X0(gq) + lane_off + hh * CW, cd = X1(gq) + lane_off + hh * CW;
      // P = exp2(S c2 - lse2), dS = P (dP scale - delta scale) for 16 columns, as packed fp32 pairs (FFMA2 / FMUL2).
      // MASKED is only instantiated for the ragged last tile: the selects cost 20 % of the loop when always executed.
      const float2 c22 = make_float2(c2, c2), sc2 = make_float2(g.scale, g.scale);
      auto chunk = [&](auto masked_tag, const uint32_t(&s16)[16], const uint32_t(&d16)[16], int ch) {
        constexpr bool MASKED = decltype(masked_tag)::value;
#pragma unroll
        for (int k = 0; k < 16; k += 4) {
          float l4[4], d4[4];
          if (MODE_KV) {
            *reinterpret_cast<float4*>(l4) = *reinterpret_cast<const float4*>(myws + ch * 16 + k);
            *reinterpret_cast<float4*>(d4) = *reinterpret_cast<const float4*>(myws + CW + ch * 16 + k);
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              l4[e] = my_lse2;
              d4[e] = my_dls;
            }
          }
#pragma unroll
          for (int e = 0; e < 4; e += 2) {
            const float2 a = ffma2(make_float2(__uint_as_float(s16[k + e]), __uint_as_float(s16[k + e + 1])), c22,
                                   make_float2(-l4[e], -l4[e + 1]));
            float2 pr = make_float2(ex2(a.x), ex2(a.y));
            if (MASKED) {
              pr.x = (ch * 16 + k + e < nvalid) ? pr.x : 0.f;
              pr.y = (ch * 16 + k + e + 1 < nvalid) ? pr.y : 0.f;
            }
            const float2 t = ffma2(make_float2(__uint_as_float(d16[k + e]), __uint_as_float(d16[k + e + 1])), sc2,
                                   make_float2(-d4[e], -d4[e + 1]));
            const float2 ds = fmul2(pr, t);
            pp[ch * 8 + (k + e) / 2] = pack_bf16x2(pr.x, pr.y);
            dd[ch * 8 + (k + e) / 2] = pack_bf16x2(ds.x, ds.y);
          }
        }
      };
      auto tile = [&](auto masked_tag) {
        tmem_ld_32x16(cs, sv[0]);
        tmem_ld_32x16(cd, dv[0]);
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          tmem_ld_wait();  // chunk ch has arrived
          if (ch + 1 < NCH) {  // next chunk in flight while this one is computed (tcgen05.ld is asynchronous until wait::ld)
            tmem_ld_32x16(cs + (ch + 1) * 16, sv[(ch + 1) & 1]);
            tmem_ld_32x16(cd + (ch + 1) * 16, dv[(ch + 1) & 1]);
          }
          chunk(masked_tag, sv[ch & 1], dv[ch & 1], ch);
        }
      };
      if (nvalid >= CW)  // warp-uniform
        tile(std::false_type{});
      else
        tile(std::true_type{});
      // bf16 pairs back into TMEM INSIDE this thread's own column range (all of its loads have completed):
      //   NH = 1: columns 0..31        NH = 2: columns 48 hh .. 48 hh + 15
      const uint32_t wcol = hh * (64 - CW / 2);
      if constexpr (CW == 64) {
        if (MODE_KV) tmem_st_32x32(X0(gq) + lane_off + wcol, pp);
        tmem_st_32x32(X1(gq) + lane_off + wcol, dd);
      } else {
        if (MODE_KV) tmem_st_32x16(X0(gq) + lane_off + wcol, pp);
        tmem_st_32x16(X1(gq) + lane_off + wcol, dd);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&pb_full[gq]);
    }
    mbar_wait(done_bar, 0, 27);
    tc_fence_after();
    // epilogue: 8 NH warps = 4 lane quarters x 2 NH column blocks
    constexpr int CB = 128 / (2 * NH);  // accumulator columns per warp
    const int cb = (warp - 2) >> 2;
#pragma unroll 1
    for (int which = 0; which < (MODE_KV ? 2 : 1); ++which) {
      const uint32_t ta = which == 0 ? tA0 : tA1;
      bf16* out = which == 0 ? g.out0 : g.out1;
#pragma unroll 1
      for (int c = 0; c < CB / 32; ++c) {
        uint32_t v[32];
        if (g.dlive == 64 && cb * CB + c * 32 >= 64) {  // padded head-dim columns: never accumulated -> exact zeros
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0u;
        } else {
          tmem_ld_32x32(ta + lane_off + cb * CB + c * 32, v);
          tmem_ld_wait();
        }
        if (ri < g.L) {
          bf16* orow = out + (row_base + ri) * 128 + cb * CB + c * 32;
#pragma unroll
          for (int k8 = 0; k8 < 4; ++k8) {
            uint4 u;
            u.x = pack_bf16x2(__uint_as_float(v[k8 * 8 + 0]), __uint_as_float(v[k8 * 8 + 1]));
            u.y = pack_bf16x2(__uint_as_float(v[k8 * 8 + 2]), __uint_as_float(v[k8 * 8 + 3]));
            u.z = pack_bf16x2(__uint_as_float(v[k8 * 8 + 4]), __uint_as_float(v[k8 * 8 + 5]));
            u.w = pack_bf16x2(__uint_as_float(v[k8 * 8 + 6]), __uint_as_float(v[k8 * 8 + 7]));
            *reinterpret_cast<uint4*>(orow + k8 * 8) = u;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// =================================================================================================
// backward, dQ pass with EARLY issue of the next S / dP MMA (variant 4; the dK/dV pass is variant 2's kernel).
// In attn_bwd_r2_kernel<0> the group's chain is  A(i) -> softmax(i) -> B(i) -> A(i+2)  because dS is written over dP inside
// X[g].  The dQ pass has only ONE accumulator, so 128 TMEM columns are free: dS goes to its own region D[g], A(i+2) may
// overwrite X[g] as soon as softmax(i) has LOADED S / dP (x_free), and the tensor pipe works on it while softmax(i) still
// computes.  Q_i / dO_i stay in shared memory (SS MMAs, as in the dK/dV pass; the TMEM-resident form measured neutral).
// Protocol checked by tools/mbar_model.py::run_bwd_q2 (softmax(i) must wait d_free = B(i-2) before it rewrites D[g]).
// =================================================================================================
constexpr int kBwdQ2Stages = 4;
constexpr int kBwdQ2Smem = 1024 + 2 * 32768 + kBwdQ2Stages * 32768 + 24 * 8;

template <int NH>
__global__ void __launch_bounds__(64 + 256 * NH, 1)
attn_bwd_q2_kernel(const __grid_constant__ CUtensorMap tmR0, const __grid_constant__ CUtensorMap tmR1,
                   const __grid_constant__ CUtensorMap tmT0, const __grid_constant__ CUtensorMap tmT1, const AttnBwdArgs g) {
  pdl_launch_dependents();
  constexpr int kStages = kBwdQ2Stages;
  constexpr int CW = 64 / NH;   // S / dP columns per thread
  constexpr int NCH = CW / 16;  // 16-column chunks per thread
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint8_t* sR0 = smem;          // Q_i  [128 x 128] bf16, two 64-column SWIZZLE_128B halves
  uint8_t* sR1 = sR0 + 32768;   // dO_i
  uint8_t* sT = sR1 + 32768;    // stage st: K_j tile at sT + st*32768, V_j tile at +16384 (64 rows each)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sT + kStages * 32768);
  uint64_t* r_full = bars;
  uint64_t* t_full = bars + 1;    // [4]
  uint64_t* t_empty = bars + 5;   // [4]
  uint64_t* x_full = bars + 9;    // [2] S / dP of the group's tile are in TMEM
  uint64_t* x_free = bars + 11;   // [2] the group has loaded them: X[g] may be overwritten
  uint64_t* pb_full = bars + 13;  // [2] dS of the group's tile is in D[g]
  uint64_t* d_free = bars + 15;   // [2] the dQ MMA has read D[g]
  uint64_t* done_bar = bars + 17;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

  const int warp = warp_id_uniform(), lane = threadIdx.x & 31;
  const int bh = blockIdx.y;
  const int r0 = blockIdx.x * 128;
  const int n_t = (g.L + 63) / 64;
  const long long row_base = static_cast<long long>(bh) * g.L;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmR0);
    tma_prefetch_desc(&tmR1);
    tma_prefetch_desc(&tmT0);
    tma_prefetch_desc(&tmT1);
  }
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(r_full, 1);
      for (int s = 0; s < kStages; ++s) {
        mbar_init(&t_full[s], 1);
        mbar_init(&t_empty[s], 1);
      }
      for (int s = 0; s < 2; ++s) {
        mbar_init(&x_full[s], 1);
        mbar_init(&x_free[s], 128 * NH);
        mbar_init(&pb_full[s], 128 * NH);
        mbar_init(&d_free[s], 1);
      }
      mbar_init(done_bar, 1);
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
  auto X0 = [&](int b) -> uint32_t { return tmem_base + static_cast<uint32_t>(b) * 64u; };         // S
  auto X1 = [&](int b) -> uint32_t { return tmem_base + 128u + static_cast<uint32_t>(b) * 64u; };  // dP
  auto Dg = [&](int b) -> uint32_t { return tmem_base + 384u + static_cast<uint32_t>(b) * 32u; };  // dS, bf16 pairs
  const uint32_t tA0 = tmem_base + 256u;                                                            // dQ accumulator

  if (warp == 0) {
    if (lane == 0) {
      const int rrow = static_cast<int>(row_base + r0);
      mbar_arrive_expect_tx(r_full, 65536);
      tma_load_2d(sR0, &tmR0, r_full, 0, rrow);
      tma_load_2d(sR0 + 16384, &tmR0, r_full, 64, rrow);
      tma_load_2d(sR1, &tmR1, r_full, 0, rrow);
      tma_load_2d(sR1 + 16384, &tmR1, r_full, 64, rrow);
      for (int i = 0; i < n_t; ++i) {
        const int st = i % kStages;
        const uint32_t ph = (i / kStages) & 1;
        const int trow = static_cast<int>(row_base + i * 64);
        mbar_wait(&t_empty[st], ph ^ 1u, 20000 + i);
        mbar_arrive_expect_tx(&t_full[st], 32768);
        uint8_t* d = sT + st * 32768;
        tma_load_2d(d, &tmT0, &t_full[st], 0, trow);
        tma_load_2d(d + 8192, &tmT0, &t_full[st], 64, trow);
        tma_load_2d(d + 16384, &tmT1, &t_full[st], 0, trow);
        tma_load_2d(d + 16384 + 8192, &tmT1, &t_full[st], 64, trow);
      }
    }
  } else if (warp == 1) {
    // converged MMA warp, elected issue (common.cuh)
    constexpr uint32_t idA = umma_idesc_bf16(128, 64, 0, 0);
    constexpr uint32_t idB = umma_idesc_bf16(128, 128, 0, 1);
    mbar_wait(r_full, 0, 21);
    tc_fence_after();
    const uint64_t dR0 = umma_desc_sw128(smem_u32(sR0), 1024, 16);
    const uint64_t dR1 = umma_desc_sw128(smem_u32(sR1), 1024, 16);
    const uint64_t dTk = umma_desc_sw128(smem_u32(sT), 1024, 16);    // K_j / V_j read K-major (S, dP)
    const uint64_t dTm = umma_desc_sw128(smem_u32(sT), 1024, 8192);  // K_j read MN-major (dQ += dS K_j)
    auto issue_A = [&](int i) {  // S = Q_i K_j^T -> X0[g], dP = dO_i V_j^T -> X1[g]
      const int st = i % kStages;
      const uint32_t ph = (i / kStages) & 1;
      const int xb = i & 1;
      mbar_wait(&t_full[st], ph, 22000 + i);
      tc_fence_after();
      const uint64_t d0 = dTk + static_cast<uint64_t>(st * (32768 >> 4));
      const uint64_t d1 = d0 + (16384 >> 4);
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const uint32_t offa = (kk >> 2) * (16384 >> 4) + 2u * (kk & 3);
        const uint32_t offb = (kk >> 2) * (8192 >> 4) + 2u * (kk & 3);
        umma_bf16_ss_w(X0(xb), dR0 + offa, d0 + offb, idA, kk > 0 ? 1u : 0u);
      }
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const uint32_t offa = (kk >> 2) * (16384 >> 4) + 2u * (kk & 3);
        const uint32_t offb = (kk >> 2) * (8192 >> 4) + 2u * (kk & 3);
        umma_bf16_ss_w(X1(xb), dR1 + offa, d1 + offb, idA, kk > 0 ? 1u : 0u);
      }
      umma_commit_w(&x_full[xb]);
    };
    issue_A(0);
    if (n_t > 1) issue_A(1);
    for (int i = 0; i < n_t; ++i) {
      const int xb = i & 1;
      const uint32_t gph = (i >> 1) & 1;
      if (i + 2 < n_t) {  // X[xb] is free as soon as the group has its S / dP in registers
        mbar_wait(&x_free[xb], gph, 23000 + i);
        issue_A(i + 2);
      }
      mbar_wait(&pb_full[xb], gph, 24000 + i);
      tc_fence_after();
      const int st = i % kStages;
      const uint64_t m0 = dTm + static_cast<uint64_t>(st * (32768 >> 4));  // K_j tile, MN-major view
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)  // dQ += dS K_j   (A = dS from TMEM: 64 kv = 4 K steps of 8 columns)
        umma_bf16_ts_w(tA0, Dg(xb) + kk * 8, m0 + kk * (2048 >> 4), idB, (kk > 0 || i > 0) ? 1u : 0u);
      umma_commit_w(&t_empty[st]);
      umma_commit_w(&d_free[xb]);
    }
    umma_commit_w(done_bar);
  } else {
    const int q = warp & 3;
    const int gq = ((warp - 2) >> 2) & 1;
    const int hh = (warp - 2) >> 3;  // 0 for NH = 1
    const int r = q * 32 + lane;
    const int ri = r0 + r;  // q index
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const float c2 = g.scale * kLog2e;
    float my_lse2 = 0.f, my_dls = 0.f;
    if (ri < g.L) {
      my_lse2 = (g.lse + row_base)[ri] * kLog2e;
      my_dls = (g.delta + row_base)[ri] * g.scale;
    }
    const float2 c22 = make_float2(c2, c2), sc2 = make_float2(g.scale, g.scale);
    const float2 nl2 = make_float2(-my_lse2, -my_lse2), nd2 = make_float2(-my_dls, -my_dls);
    for (int i = gq; i < n_t; i += 2) {
      const uint32_t gph = (i >> 1) & 1;
      mbar_wait(&x_full[gq], gph, 25000 + i);
      tc_fence_after();
      const int nvalid = g.L - (i * 64 + hh * CW);  // kv columns of my part that exist (<= 0: none)
      uint32_t dd[CW / 2];
      uint32_t sv[2][16], dv[2][16];
      const uint32_t cs = X0(gq) + lane_off + hh * CW, cd = X1(gq) + lane_off + hh * CW;
      auto chunk = [&](auto masked_tag, const uint32_t(&s16)[16], const uint32_t(&d16)[16], int ch) {
        constexpr bool MASKED = decltype(masked_tag)::value;
#pragma unroll
        for (int k = 0; k < 16; k += 2) {
          const float2 a = ffma2(make_float2(__uint_as_float(s16[k]), __uint_as_float(s16[k + 1])), c22, nl2);
          float2 pr = make_float2(ex2(a.x), ex2(a.y));
          if (MASKED) {
            pr.x = (ch * 16 + k < nvalid) ? pr.x : 0.f;
            pr.y = (ch * 16 + k + 1 < nvalid) ? pr.y : 0.f;
          }
          const float2 t = ffma2(make_float2(__uint_as_float(d16[k]), __uint_as_float(d16[k + 1])), sc2, nd2);
          const float2 ds = fmul2(pr, t);
          dd[ch * 8 + k / 2] = pack_bf16x2(ds.x, ds.y);
        }
      };
      auto tile = [&](auto masked_tag) {
        tmem_ld_32x16(cs, sv[0]);
        tmem_ld_32x16(cd, dv[0]);
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          tmem_ld_wait();  // chunk ch has arrived
          if (ch + 1 < NCH) {
            tmem_ld_32x16(cs + (ch + 1) * 16, sv[(ch + 1) & 1]);
            tmem_ld_32x16(cd + (ch + 1) * 16, dv[(ch + 1) & 1]);
          } else {
            tc_fence_before();
            mbar_arrive(&x_free[gq]);  // all of my S / dP are in registers: the next S / dP MMA may overwrite X[gq]
          }
          chunk(masked_tag, sv[ch & 1], dv[ch & 1], ch);
        }
      };
      if (nvalid >= CW)  // warp-uniform
        tile(std::false_type{});
      else
        tile(std::true_type{});
      if (i >= 2) {  // the dQ MMA of tile i-2 has read D[gq]
        mbar_wait(&d_free[gq], ((i >> 1) - 1) & 1, 26000 + i);
        tc_fence_after();
      }
      if constexpr (CW == 64)
        tmem_st_32x32(Dg(gq) + lane_off, dd);
      else
        tmem_st_32x16(Dg(gq) + lane_off + hh * 16, dd);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&pb_full[gq]);
    }
    mbar_wait(done_bar, 0, 27);
    tc_fence_after();
    // epilogue: 8 NH warps = 4 lane quarters x 2 NH column blocks of the dQ accumulator
    constexpr int CB = 128 / (2 * NH);
    const int cb = (warp - 2) >> 2;
#pragma unroll 1
    for (int c = 0; c < CB / 32; ++c) {
      uint32_t v[32];
      tmem_ld_32x32(tA0 + lane_off + cb * CB + c * 32, v);
      tmem_ld_wait();
      if (ri < g.L) {
        bf16* orow = g.out0 + (row_base + ri) * 128 + cb * CB + c * 32;
#pragma unroll
        for (int k8 = 0; k8 < 4; ++k8) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(v[k8 * 8 + 0]), __uint_as_float(v[k8 * 8 + 1]));
          u.y = pack_bf16x2(__uint_as_float(v[k8 * 8 + 2]), __uint_as_float(v[k8 * 8 + 3]));
          u.z = pack_bf16x2(__uint_as_float(v[k8 * 8 + 4]), __uint_as_float(v[k8 * 8 + 5]));
          u.w = pack_bf16x2(__uint_as_float(v[k8 * 8 + 6]), __uint_as_float(v[k8 * 8 + 7]));
          *reinterpret_cast<uint4*>(orow + k8 * 8) = u;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// =================================================================================================
// host-side launchers (called from b200_attn_fwd / b200_attn_bwd when the environment selects a candidate)
// =================================================================================================
template <int NT>
static int launch_fwd_nt(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnFwdArgs& a,
                         cudaStream_t stream) {
  auto kern = attn_fwd_nt_kernel<NT>;
  static bool configured = false;
  if (!configured) {
    B200_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdR2<NT>::kSmem));
    configured = true;
  }
  dim3 grid((a.L + 127) / 128, a.B * a.H);
  B200_KLAUNCH(kern, grid, FwdR2<NT>::kThreads, FwdR2<NT>::kSmem, stream, tq, tk, tv, a);
  B200_CUDA_CHECK(cudaGetLastError());
  return B200_OK;
}

int attn_fwd_r2_launch(int variant, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnFwdArgs& a,
                       cudaStream_t stream) {
  return variant == 4 ? launch_fwd_nt<4>(tq, tk, tv, a, stream) : launch_fwd_nt<2>(tq, tk, tv, a, stream);
}

template <int NH>
static int launch_bwd_nh(const CUtensorMap& k128, const CUtensorMap& v128, const CUtensorMap& q64, const CUtensorMap& d64,
                         const CUtensorMap& q128, const CUtensorMap& d128, const CUtensorMap& k64, const CUtensorMap& v64,
                         const AttnBwdArgs& akv, const AttnBwdArgs& aq, int B, int H, cudaStream_t stream) {
  auto kkv = attn_bwd_r2_kernel<1, NH>;
  auto kq = attn_bwd_r2_kernel<0, NH>;
  static bool configured = false;
  if (!configured) {
    B200_CUDA_CHECK(cudaFuncSetAttribute(kkv, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdR2Smem));
    B200_CUDA_CHECK(cudaFuncSetAttribute(kq, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdR2Smem));
    configured = true;
  }
  dim3 grid_kv((akv.L + 127) / 128, B * H), grid_q((aq.L + 127) / 128, B * H);
  B200_KLAUNCH(kkv, grid_kv, 64 + 256 * NH, kBwdR2Smem, stream, k128, v128, q64, d64, akv);
  B200_CUDA_CHECK(cudaGetLastError());
  B200_KLAUNCH(kq, grid_q, 64 + 256 * NH, kBwdR2Smem, stream, q128, d128, k64, v64, aq);
  B200_CUDA_CHECK(cudaGetLastError());
  return B200_OK;
}

template <int NH>
static int launch_bwd_q2(const CUtensorMap& k128, const CUtensorMap& v128, const CUtensorMap& q64, const CUtensorMap& d64,
                         const CUtensorMap& q128, const CUtensorMap& d128, const CUtensorMap& k64, const CUtensorMap& v64,
                         const AttnBwdArgs& akv, const AttnBwdArgs& aq, int B, int H, cudaStream_t stream) {
  auto kkv = attn_bwd_r2_kernel<1, NH>;
  auto kq = attn_bwd_q2_kernel<NH>;
  static bool configured = false;
  if (!configured) {
    B200_CUDA_CHECK(cudaFuncSetAttribute(kkv, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdR2Smem));
    B200_CUDA_CHECK(cudaFuncSetAttribute(kq, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdQ2Smem));
    configured = true;
  }
  dim3 grid((akv.L + 127) / 128, B * H);
  B200_KLAUNCH(kkv, grid, 64 + 256 * NH, kBwdR2Smem, stream, k128, v128, q64, d64, akv);
  B200_CUDA_CHECK(cudaGetLastError());
  B200_KLAUNCH(kq, grid, 64 + 256 * NH, kBwdQ2Smem, stream, q128, d128, k64, v64, aq);
  B200_CUDA_CHECK(cudaGetLastError());
  return B200_OK;
}

int attn_bwd_r2_launch(int variant, const CUtensorMap& k128, const CUtensorMap& v128, const CUtensorMap& q64,
                       const CUtensorMap& d64, const CUtensorMap& q128, const CUtensorMap& d128, const CUtensorMap& k64,
                       const CUtensorMap& v64, const AttnBwdArgs& akv, const AttnBwdArgs& aq, int B, int H,
                       cudaStream_t stream) {
  if (variant == 4) return launch_bwd_q2<2>(k128, v128, q64, d64, q128, d128, k64, v64, akv, aq, B, H, stream);
  if (variant == 5) return launch_bwd_q2<1>(k128, v128, q64, d64, q128, d128, k64, v64, akv, aq, B, H, stream);
  if (variant == 2) return launch_bwd_nh<2>(k128, v128, q64, d64, q128, d128, k64, v64, akv, aq, B, H, stream);
  return launch_bwd_nh<1>(k128, v128, q64, d64, q128, d128, k64, v64, akv, aq, B, H, stream);
}

}  // namespace b200
