// Flash attention for the DiT blocks on tcgen05 / TMEM / TMA (head_dim 128, non-causal, joint text||image
// sequence), forward and backward.  Operands are head-major [B, H, L, 128] bf16 (written by
// b200_qk_norm_rope_fwd); the forward writes O token-major straight into the operand buffers of the
// following projection GEMMs.
//
// Forward (one CTA per 128-row Q tile of one (b, h); 192 threads = TMA warp, MMA warp, 4 softmax warps):
//   S = Q K_j^T        tcgen05.mma, both operands K-major from SWIZZLE_128B tiles, accumulator in TMEM (2 buffers)
//   P = exp2(S c - m)  one thread per row reads its row with tcgen05.ld (no shuffles), lazy rescaling of O
//                      (only when the running max grows by more than 2^8), P written bf16 to swizzled smem
//   O += P V_j         tcgen05.mma, V consumed MN-major exactly as it lies in HBM ([kv, d], d contiguous)
// Backward = three kernels: delta/dO re-layout, dK/dV (kv-stationary) and dQ (q-stationary); the two
// tensor-core kernels are one template: the stationary 128-row operands stay in smem, 64-row tiles of the
// other side stream through a 3-stage TMA ring, S and dP go to double-buffered TMEM, P / dS are written
// bf16 to smem once and consumed both K-major and MN-major (same bytes, two descriptors).
//
// Reference semantics: torch SDPA as called by diffusers' FLUX attention processor (in-tree equivalent
// extensions_built_in/diffusion_models/chroma/src/math.py:13-30); backward = its autograd.
#include "attn_common.cuh"

namespace b200 {

// =================================================================================================
// forward
// =================================================================================================
constexpr int kFwdSmem = 1024 + 32768 + 2 * 32768 + 2 * 32768 + 16 * 8 + 16 + 2 * 2 * 128 * 4;
constexpr int kAttnThreads = 320;  // TMA warp, MMA warp, 8 softmax warps (two per TMEM sub-partition)

__global__ void __launch_bounds__(kAttnThreads, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnFwdArgs g) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_base_1024(smem_raw);
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
  uint64_t* p_full = bars + 13;
  uint64_t* pv_done = bars + 14;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);
  float* xch = reinterpret_cast<float*>(bars + 16);  // [2 parity][2 halves][128 rows] row-max / row-sum exchange

  const int warp = warp_id_uniform(), lane = threadIdx.x & 31;
  const int bh = blockIdx.y;
  const int q0 = blockIdx.x * 128;
  const int n_kv = (g.L + 127) / 128;
  const long long row_base = static_cast<long long>(bh) * g.L;

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
        mbar_init(&s_empty[s], 256);
      }
      mbar_init(p_full, 256);
      mbar_init(pv_done, 1);
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
  const uint32_t tS[2] = {tmem_base, tmem_base + 128u};
  const uint32_t tO = tmem_base + 256u;
  const uint32_t tP[2] = {tmem_base + 384u, tmem_base + 448u};  // bf16 P, two K elements per 32-bit column

  if (warp == 0) {
    if (lane == 0) {
      const int qrow = static_cast<int>(row_base + q0);
      mbar_arrive_expect_tx(q_full, 32768);
      tma_load_2d(sQ, &tmQ, q_full, 0, qrow);
      tma_load_2d(sQ + 16384, &tmQ, q_full, 64, qrow);
      for (int j = 0; j < n_kv; ++j) {
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        const int kvrow = static_cast<int>(row_base + j * 128);
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
    {  // all 32 lanes run the control flow (converged); one elected lane issues each tcgen05 instruction
      constexpr uint32_t idS = umma_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idPV = umma_idesc_bf16(128, 128, 0, 1);
      mbar_wait(q_full, 0, 12);
      // descriptors: one base per operand buffer, compile-time offsets per k-step (address field is (addr >> 4))
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
          umma_bf16_ss_w(tS[s], dQ0 + off, dk + off, idS, kk > 0 ? 1u : 0u);
        }
        umma_commit_w(&k_empty[s]);
        umma_commit_w(&s_full[s]);
      };
      auto issue_PV = [&](int j) {
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&v_full[s], ph, 15);
        mbar_wait(p_full, j & 1, 16);
        tc_fence_after();
        const uint64_t dv = dV0 + static_cast<uint64_t>(s * (32768 >> 4));
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t off = (kk >> 2) * (8192 >> 4) + (kk & 3) * (2048 >> 4);
          umma_bf16_ts_w(tO, tP[s] + kk * 8, dv + off, idPV, (j > 0 || kk > 0) ? 1u : 0u);  // A = P straight from TMEM
        }
        umma_commit_w(&v_empty[s]);
        umma_commit_w(pv_done);
      };
      issue_S(0);
      for (int j = 0; j < n_kv; ++j) {
        if (j + 1 < n_kv) issue_S(j + 1);
        issue_PV(j);
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax / epilogue
    // two threads per row: thread (r, h) owns columns [64 h, 64 h + 64) of the S / P / O row r
    const int q = warp & 3;
    const int h = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const int qi = q0 + r;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const float c2 = g.scale * kLog2e;
    float m_used = -INFINITY, l_sum = 0.f;
    for (int j = 0; j < n_kv; ++j) {
      const int s = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      mbar_wait(&s_full[s], ph, 17);
      tc_fence_after();
      uint32_t v0[32], v1[32];
      tmem_ld_32x32(tS[s] + lane_off + h * 64, v0);
      tmem_ld_32x32(tS[s] + lane_off + h * 64 + 32, v1);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&s_empty[s]);  // S is in registers: the tensor core may overwrite this buffer
      const int nvalid = g.L - (j * 128 + h * 64);  // columns of my half that exist
      float x[64];
      float mx = -INFINITY;
      if (nvalid >= 64) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          x[i] = __uint_as_float(v0[i]) * c2;
          x[32 + i] = __uint_as_float(v1[i]) * c2;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          x[i] = i < nvalid ? __uint_as_float(v0[i]) * c2 : -INFINITY;
          x[32 + i] = 32 + i < nvalid ? __uint_as_float(v1[i]) * c2 : -INFINITY;
        }
      }
#pragma unroll
      for (int i = 0; i < 64; ++i) mx = fmaxf(mx, x[i]);
      float* xm = xch + (j & 1) * 256;
      xm[h * 128 + r] = mx;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const float m_new = fmaxf(m_used, fmaxf(mx, xm[(h ^ 1) * 128 + r]));
      const bool need = (m_new > m_used + 8.0f);  // also true for the first tile (m_used = -inf)
      const bool any_need = __any_sync(0xffffffffu, need);
      if (j > 0) {
        mbar_wait(pv_done, (j - 1) & 1, 18);  // P buffer free, O stable
        tc_fence_after();
      }
      if (any_need) {
        const float f = need ? ex2(m_used - m_new) : 1.0f;  // first tile: ex2(-inf) = 0
        if (need) {
          m_used = m_new;
          l_sum *= f;
        }
        if (j > 0) {
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            uint32_t v[32];
            tmem_ld_32x32(tO + lane_off + h * 64 + c * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * f);
            tmem_st_32x32(tO + lane_off + h * 64 + c * 32, v);
          }
          tmem_st_wait();
        }
      }
      // P = exp2(S c - m_used) -> bf16 pairs -> TMEM (the A operand of the P V MMA; no shared-memory round trip)
      uint32_t pk[32];
#pragma unroll
      for (int k2 = 0; k2 < 32; ++k2) {
        // every other pair goes to the FMA-pipe polynomial (kPolyShare of the exponentials), the rest to MUFU
        const bool poly = (k2 % 4) < kPolyPairsOf4;
        const float pa = poly ? ex2_poly(x[2 * k2] - m_used) : ex2(x[2 * k2] - m_used);
        const float pb = poly ? ex2_poly(x[2 * k2 + 1] - m_used) : ex2(x[2 * k2 + 1] - m_used);
        l_sum += pa + pb;
        pk[k2] = pack_bf16x2(pa, pb);
      }
      tmem_st_32x32(tP[s] + lane_off + h * 32, pk);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_full);
    }
    // combine the two half-row sums
    float* xl = xch + (n_kv & 1) * 256;
    xl[h * 128 + r] = l_sum;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    l_sum += xl[(h ^ 1) * 128 + r];
    mbar_wait(pv_done, (n_kv - 1) & 1, 19);
    tc_fence_after();
    const float inv = 1.0f / l_sum;
    const bool live = qi < g.L;
    bf16* orow = nullptr;
    if (live) {
      const int b = bh / g.H, hh = bh % g.H;
      if (qi < g.split)
        orow = g.o0 + (static_cast<size_t>(b) * g.split + qi) * g.ld0 + hh * 128 + h * 64;
      else
        orow = g.o1 + (static_cast<size_t>(b) * (g.L - g.split) + (qi - g.split)) * g.ld1 + hh * 128 + h * 64;
      if (h == 0) g.lse[row_base + qi] = (m_used + log2f(l_sum)) * kLn2;
    }
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {  // tcgen05.ld is warp-collective: every lane loads, only live rows store
      uint32_t v[32];
      tmem_ld_32x32(tO + lane_off + h * 64 + c * 32, v);
      tmem_ld_wait();
      if (live) {
#pragma unroll
        for (int k8 = 0; k8 < 4; ++k8) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(v[k8 * 8 + 0]) * inv, __uint_as_float(v[k8 * 8 + 1]) * inv);
          u.y = pack_bf16x2(__uint_as_float(v[k8 * 8 + 2]) * inv, __uint_as_float(v[k8 * 8 + 3]) * inv);
          u.z = pack_bf16x2(__uint_as_float(v[k8 * 8 + 4]) * inv, __uint_as_float(v[k8 * 8 + 5]) * inv);
          u.w = pack_bf16x2(__uint_as_float(v[k8 * 8 + 6]) * inv, __uint_as_float(v[k8 * 8 + 7]) * inv);
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

// -------------------------------------------------------------------------------------------------
// forward, variant 2: TWO independent online-softmax streams per Q tile.  Softmax group g (4 warps, one thread per
// full 128-column row) owns the kv blocks j == g (mod 2), its own S buffer (P is written back in place over S) and its
// own O accumulator; the two partial results are combined in the epilogue (split-KV combine).  While one group works
// through its tcgen05.ld -> exp2 -> tcgen05.st chain the tensor core serves the other, so the per-block latency chain
// is hidden without the row-max exchange / bar.sync of variant 1.
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kAttnThreads, 1)
attn_fwd2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const AttnFwdArgs g) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_base_1024(smem_raw);
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
  uint64_t* p_full = bars + 11;
  uint64_t* pv_done = bars + 13;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);
  float* xch = reinterpret_cast<float*>(bars + 16);  // [2 groups][128 rows][2]: (m, l) of each stream

  const int warp = warp_id_uniform(), lane = threadIdx.x & 31;
  const int bh = blockIdx.y;
  const int q0 = blockIdx.x * 128;
  const int n_kv = (g.L + 127) / 128;
  const long long row_base = static_cast<long long>(bh) * g.L;

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
  const uint32_t tS[2] = {tmem_base, tmem_base + 128u};         // S of stream g; P (bf16 pairs) overwrites its first 64 columns
  const uint32_t tO[2] = {tmem_base + 256u, tmem_base + 384u};  // O accumulator of stream g

  if (warp == 0) {
    if (lane == 0) {
      const int qrow = static_cast<int>(row_base + q0);
      mbar_arrive_expect_tx(q_full, 32768);
      tma_load_2d(sQ, &tmQ, q_full, 0, qrow);
      tma_load_2d(sQ + 16384, &tmQ, q_full, 64, qrow);
      for (int j = 0; j < n_kv; ++j) {
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        const int kvrow = static_cast<int>(row_base + j * 128);
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
    {  // converged MMA warp, elected issue
      constexpr uint32_t idS = umma_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idPV = umma_idesc_bf16(128, 128, 0, 1);
      mbar_wait(q_full, 0, 12);
      const uint64_t dQ0 = umma_desc_sw128(smem_u32(sQ), 1024, 16);
      const uint64_t dK0 = umma_desc_sw128(smem_u32(sK), 1024, 16);
      const uint64_t dV0 = umma_desc_sw128(smem_u32(sV), 1024, 16384);
      auto issue_S = [&](int j) {  // S(j) overwrites the buffer whose P was read by PV(j-2): same thread, issue order
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&k_full[s], ph, 13);
        tc_fence_after();
        const uint64_t dk = dK0 + static_cast<uint64_t>(s * (32768 >> 4));
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t off = (kk >> 2) * (16384 >> 4) + 2u * (kk & 3);
          umma_bf16_ss_w(tS[s], dQ0 + off, dk + off, idS, kk > 0 ? 1u : 0u);
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
          umma_bf16_ts_w(tO[s], tS[s] + kk * 8, dv + off, idPV, (j >= 2 || kk > 0) ? 1u : 0u);  // A = P in TMEM
        }
        umma_commit_w(&v_empty[s]);
        umma_commit_w(&pv_done[s]);
      };
      issue_S(0);
      if (n_kv > 1) issue_S(1);
      for (int j = 0; j < n_kv; ++j) {
        issue_PV(j);
        if (j + 2 < n_kv) issue_S(j + 2);
      }
    }
  } else {
    const int q = warp & 3;
    const int gq = (warp - 2) >> 2;  // stream / group
    const int r = q * 32 + lane;
    const int qi = q0 + r;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const float c2 = g.scale * kLog2e;
    float m_used = -INFINITY, l_sum = 0.f;
    for (int j = gq; j < n_kv; j += 2) {
      const uint32_t ph = (j >> 1) & 1;
      mbar_wait(&s_full[gq], ph, 17);
      tc_fence_after();
      const int nvalid = g.L - j * 128;
      // pass 1: row max (S stays in TMEM; reading it twice is cheaper than 128 live registers)
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(tS[gq] + lane_off + c * 32, v);
        tmem_ld_wait();
        if (nvalid >= 128) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, (c * 32 + i < nvalid) ? __uint_as_float(v[i]) : -INFINITY);
        }
      }
      mx *= c2;  // c2 > 0: max commutes with the scale
      const float m_new = fmaxf(m_used, mx);
      const bool need = (m_new > m_used + 8.0f);
      const bool any_need = __any_sync(0xffffffffu, need);
      if (j >= 2) {
        mbar_wait(&pv_done[gq], ((j >> 1) - 1) & 1, 18);  // PV of this stream's previous block: O_g stable
        tc_fence_after();
      }
      if (any_need) {
        const float f = need ? ex2(m_used - m_new) : 1.0f;
        if (need) {
          m_used = m_new;
          l_sum *= f;
        }
        if (j >= 2) {
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            uint32_t v[32];
            tmem_ld_32x32(tO[gq] + lane_off + c * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * f);
            tmem_st_32x32(tO[gq] + lane_off + c * 32, v);
          }
          tmem_st_wait();
        }
      }
      // pass 2: P = exp2(S c2 - m) in place: chunk c (32 fp32 columns) -> 16 packed bf16 columns at 16 c (already consumed)
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(tS[gq] + lane_off + c * 32, v);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int k2 = 0; k2 < 16; ++k2) {
          const float xa = fmaf(__uint_as_float(v[2 * k2]), c2, -m_used), xb = fmaf(__uint_as_float(v[2 * k2 + 1]), c2, -m_used);
          const bool poly = (k2 % 4) < kPolyPairsOf4;
          float pa = poly ? ex2_poly(xa) : ex2(xa);
          float pb = poly ? ex2_poly(xb) : ex2(xb);
          if (nvalid < 128) {
            pa = (c * 32 + 2 * k2 < nvalid) ? pa : 0.f;
            pb = (c * 32 + 2 * k2 + 1 < nvalid) ? pb : 0.f;
          }
          l_sum += pa + pb;
          pk[k2] = pack_bf16x2(pa, pb);
        }
        tmem_st_32x16(tS[gq] + lane_off + c * 16, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[gq]);
    }
    // ---- combine the two streams
    xch[(gq * 128 + r) * 2] = m_used;
    xch[(gq * 128 + r) * 2 + 1] = l_sum;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const float m0 = xch[r * 2], l0 = xch[r * 2 + 1], m1 = xch[(128 + r) * 2], l1 = xch[(128 + r) * 2 + 1];
    const bool has1 = n_kv > 1;
    const float m = has1 ? fmaxf(m0, m1) : m0;
    const float f0 = ex2(m0 - m), f1 = has1 ? ex2(m1 - m) : 0.f;
    const float l = l0 * f0 + (has1 ? l1 * f1 : 0.f);
    const int nb0 = (n_kv + 1) / 2, nb1 = n_kv / 2;  // blocks of stream 0 / 1
    mbar_wait(&pv_done[0], (nb0 - 1) & 1, 19);
    if (has1) mbar_wait(&pv_done[1], (nb1 - 1) & 1, 19);
    tc_fence_after();
    const float inv = 1.0f / l;
    const float w0 = f0 * inv, w1 = f1 * inv;
    const bool live = qi < g.L;
    bf16* orow = nullptr;
    if (live) {
      const int b = bh / g.H, hh = bh % g.H;
      if (qi < g.split)
        orow = g.o0 + (static_cast<size_t>(b) * g.split + qi) * g.ld0 + hh * 128 + gq * 64;
      else
        orow = g.o1 + (static_cast<size_t>(b) * (g.L - g.split) + (qi - g.split)) * g.ld1 + hh * 128 + gq * 64;
      if (gq == 0) g.lse[row_base + qi] = (m + log2f(l)) * kLn2;
    }
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {  // group gq writes output columns [64 gq, 64 gq + 64)
      uint32_t a0[32], a1[32];
      tmem_ld_32x32(tO[0] + lane_off + gq * 64 + c * 32, a0);
      if (has1) tmem_ld_32x32(tO[1] + lane_off + gq * 64 + c * 32, a1);
      tmem_ld_wait();
      if (live) {
#pragma unroll
        for (int k8 = 0; k8 < 4; ++k8) {
          float o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            o[e] = __uint_as_float(a0[k8 * 8 + e]) * w0;
            if (has1) o[e] = fmaf(__uint_as_float(a1[k8 * 8 + e]), w1, o[e]);
          }
          uint4 u;
          u.x = pack_bf16x2(o[0], o[1]);
          u.y = pack_bf16x2(o[2], o[3]);
          u.z = pack_bf16x2(o[4], o[5]);
          u.w = pack_bf16x2(o[6], o[7]);
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
// backward
// =================================================================================================
// delta[b,h,l] = sum_d dO * O ;  dOh = dO re-laid head-major.  One warp per (b, l, h).
__global__ void __launch_bounds__(256) attn_delta_kernel(const bf16* __restrict__ o0, int ld0, const bf16* __restrict__ o1, int ld1,
                                                         const bf16* __restrict__ do0, int ldd0,
                                                         const bf16* __restrict__ do1, int ldd1, float* __restrict__ delta,
                                                         bf16* __restrict__ dOh, int B, int H, int L, int split) {
  pdl_grid_sync();
  const long long w = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (w >= static_cast<long long>(B) * L * H) return;
  const int h = static_cast<int>(w % H);
  const long long tok = w / H;
  const int l = static_cast<int>(tok % L);
  const int b = static_cast<int>(tok / L);
  const bf16 *po, *pd;
  if (l < split) {
    po = o0 + (static_cast<size_t>(b) * split + l) * ld0 + h * 128 + lane * 4;
    pd = do0 + (static_cast<size_t>(b) * split + l) * ldd0 + h * 128 + lane * 4;
  } else {
    po = o1 + (static_cast<size_t>(b) * (L - split) + (l - split)) * ld1 + h * 128 + lane * 4;
    pd = do1 + (static_cast<size_t>(b) * (L - split) + (l - split)) * ldd1 + h * 128 + lane * 4;
  }
  const uint2 uo = *reinterpret_cast<const uint2*>(po);
  const uint2 ud = *reinterpret_cast<const uint2*>(pd);
  const float2 a0 = unpack_bf16x2(uo.x), a1 = unpack_bf16x2(uo.y), d0 = unpack_bf16x2(ud.x), d1 = unpack_bf16x2(ud.y);
  float s = a0.x * d0.x + a0.y * d0.y + a1.x * d1.x + a1.y * d1.y;
  s = warp_sum(s);
  const size_t hm = ((static_cast<size_t>(b) * H + h) * L + l);
  if (lane == 0) delta[hm] = s;
  *reinterpret_cast<uint2*>(dOh + hm * 128 + lane * 4) = ud;
}

constexpr int kBwdStages = 3;
constexpr int kBwdSmem = 1024 + 2 * 32768 + kBwdStages * 32768 + 16 * 8 + 16 + 8 * 128 * 4;

// MODE_KV = 1: stationary (R0, R1) = (K_j, V_j), streamed (T0, T1) = (Q_i, dO_i); outputs dV (acc0), dK (acc1)
// MODE_KV = 0: stationary (R0, R1) = (Q_i, dO_i), streamed (T0, T1) = (K_j, V_j); output  dQ (acc0)
template <int MODE_KV>
__global__ void __launch_bounds__(kAttnThreads, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmR0, const __grid_constant__ CUtensorMap tmR1,
                const __grid_constant__ CUtensorMap tmT0, const __grid_constant__ CUtensorMap tmT1, const AttnBwdArgs g) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_base_1024(smem_raw);
  uint8_t* sR0 = smem;
  uint8_t* sR1 = sR0 + 32768;
  uint8_t* sT = sR1 + 32768;  // stage st: T0 at sT + st*32768, T1 at +16384
  uint64_t* bars = reinterpret_cast<uint64_t*>(sT + kBwdStages * 32768);
  uint64_t* r_full = bars;
  uint64_t* t_full = bars + 1;
  uint64_t* t_empty = bars + 4;
  uint64_t* x_full = bars + 7;
  uint64_t* x_empty = bars + 9;
  uint64_t* pb_full = bars + 11;  // [2]: one per softmax group / TMEM buffer
  uint64_t* done_bar = bars + 13;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
  float* colws = reinterpret_cast<float*>(bars + 16);  // per softmax warp: 32 x (lse2, delta*scale) of its columns

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
      for (int s = 0; s < kBwdStages; ++s) {
        mbar_init(&t_full[s], 1);
        mbar_init(&t_empty[s], 1);
      }
      for (int s = 0; s < 2; ++s) {
        mbar_init(&x_full[s], 1);
        mbar_init(&x_empty[s], 256);  // [0] doubles as 'stationary operands are in TMEM' (MODE_Q)
      }
      mbar_init(&pb_full[0], 128);
      mbar_init(&pb_full[1], 128);
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
  const uint32_t tX0[2] = {tmem_base, tmem_base + 64u};
  const uint32_t tX1[2] = {tmem_base + 128u, tmem_base + 192u};
  const uint32_t tA0 = tmem_base + 256u, tA1 = tmem_base + 384u;
  // MODE_Q: only one accumulator (dQ), so columns 384..511 hold the stationary A operands Q_i / dO_i as bf16 pairs:
  // phase A then reads A from TMEM and only the 64-row streamed tile from shared memory (the N = 64 SS form needs
  // 6 KB of smem reads per 32-cycle MMA).  Measured: correct, but end-to-end neutral (886 vs ~870 us backward), so
  // shared-memory bandwidth is not what holds this kernel at ~52 % tensor-pipe activity.
  const uint32_t tR0 = tmem_base + 384u, tR1 = tmem_base + 448u;

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
        const int st = i % kBwdStages;
        const uint32_t ph = (i / kBwdStages) & 1;
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
    {  // converged MMA warp, elected issue (see common.cuh)
      constexpr uint32_t idA = umma_idesc_bf16(128, 64, 0, 0);
      constexpr uint32_t idB = umma_idesc_bf16(128, 128, 0, 1);
      const uint32_t tR0u = tR0, tR1u = tR1;
      if (MODE_KV)
        mbar_wait(r_full, 0, 21);
      else
        mbar_wait(&x_empty[0], 0, 21);  // the softmax warps have stored Q_i / dO_i into TMEM
      tc_fence_after();
      const uint64_t dR0 = umma_desc_sw128(smem_u32(sR0), 1024, 16);
      const uint64_t dR1 = umma_desc_sw128(smem_u32(sR1), 1024, 16);
      const uint64_t dTk = umma_desc_sw128(smem_u32(sT), 1024, 16);     // streamed tiles read K-major (phase A)
      const uint64_t dTm = umma_desc_sw128(smem_u32(sT), 1024, 8192);   // the same bytes read MN-major (phase B)
      auto issue_A = [&](int i) {
        const int st = i % kBwdStages;
        const uint32_t ph = (i / kBwdStages) & 1;
        const int xb = i & 1;
        mbar_wait(&t_full[st], ph, 22000 + i);
        // X[xb] was last read by B(i-2) (P / dS alias it); MMAs of one thread execute in issue order, so no barrier
        tc_fence_after();
        const uint64_t d0 = dTk + static_cast<uint64_t>(st * (32768 >> 4));
        const uint64_t d1 = d0 + (16384 >> 4);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t offa = (kk >> 2) * (16384 >> 4) + 2u * (kk & 3);
          const uint32_t offb = (kk >> 2) * (8192 >> 4) + 2u * (kk & 3);
          if (MODE_KV)
            umma_bf16_ss_w(tX0[xb], dR0 + offa, d0 + offb, idA, kk > 0 ? 1u : 0u);
          else
            umma_bf16_ts_w(tX0[xb], tR0u + kk * 8, d0 + offb, idA, kk > 0 ? 1u : 0u);
        }
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t offa = (kk >> 2) * (16384 >> 4) + 2u * (kk & 3);
          const uint32_t offb = (kk >> 2) * (8192 >> 4) + 2u * (kk & 3);
          if (MODE_KV)
            umma_bf16_ss_w(tX1[xb], dR1 + offa, d1 + offb, idA, kk > 0 ? 1u : 0u);
          else
            umma_bf16_ts_w(tX1[xb], tR1u + kk * 8, d1 + offb, idA, kk > 0 ? 1u : 0u);
        }
        umma_commit_w(&x_full[xb]);
      };
      auto issue_B = [&](int i) {
        const int st = i % kBwdStages;
        mbar_wait(&pb_full[i & 1], (i >> 1) & 1, 24000 + i + 100000 * MODE_KV);
        tc_fence_after();
        const uint64_t m0 = dTm + static_cast<uint64_t>(st * (32768 >> 4));  // T0 tile, MN-major view
        const uint64_t m1 = m0 + (16384 >> 4);                                // T1 tile
        const uint32_t acc = i > 0 ? 1u : 0u;
        const int xb = i & 1;
        // A operands come from TMEM: each row's 64 bf16 values sit in the first 32 columns of X (two per column)
        if (MODE_KV) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)  // dV += P^T dO_i
            umma_bf16_ts_w(tA0, tX0[xb] + kk * 8, m1 + kk * (2048 >> 4), idB, (kk > 0) ? 1u : acc);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)  // dK += dS^T Q_i
            umma_bf16_ts_w(tA1, tX1[xb] + kk * 8, m0 + kk * (2048 >> 4), idB, (kk > 0) ? 1u : acc);
        } else {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)  // dQ += dS K_j
            umma_bf16_ts_w(tA0, tX1[xb] + kk * 8, m0 + kk * (2048 >> 4), idB, (kk > 0) ? 1u : acc);
        }
        umma_commit_w(&t_empty[st]);
      };
      issue_A(0);
      for (int i = 0; i < n_t; ++i) {
        if (i + 1 < n_t) issue_A(i + 1);
        issue_B(i);
      }
      umma_commit_w(done_bar);
    }
  } else {
    // Two softmax groups of 4 warps: group gq owns the tiles i == gq (mod 2), i.e. always the TMEM buffer X[gq], one
    // thread per full 64-column row.  Each group has two MMA tile-times to turn S/dP into P/dS, so the tcgen05.ld ->
    // ex2 -> tcgen05.st latency chain of one group hides behind the other's; a group cannot lap (its next x_full needs
    // its own previous pb_full).
    const int q = warp & 3;
    const int gq = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const int ri = r0 + r;  // kv index (MODE_KV) or q index
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const float c2 = g.scale * kLog2e;
    const float* lse_bh = g.lse + row_base;
    const float* delta_bh = g.delta + row_base;
    float* myws = colws + (warp - 2) * 128;  // per warp: 64 x lse2, 64 x delta*scale of the tile's columns
    float my_lse2 = 0.f, my_dls = 0.f;
    if (!MODE_KV && ri < g.L) {
      my_lse2 = lse_bh[ri] * kLog2e;
      my_dls = delta_bh[ri] * g.scale;
    }
    if (!MODE_KV) {
      // group 0 stores Q_i, group 1 stores dO_i: one 256-byte row per thread, bf16 pairs as they lie in memory
      const bf16* src = (gq == 0 ? g.r0 : g.r1) + (row_base + ri) * 128;
      const uint32_t dst = (gq == 0 ? tR0 : tR1) + lane_off;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
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
      mbar_arrive(&x_empty[0]);
    }
    // software prefetch of the per-column statistics (MODE_KV): registers for the group's next tile
    float nl0 = 0.f, nl1 = 0.f, nd0 = 0.f, nd1 = 0.f;
    auto fetch_cols = [&](int i) {
      const int c = i * 64 + lane;
      nl0 = c < g.L ? lse_bh[c] * kLog2e : 0.f;
      nd0 = c < g.L ? delta_bh[c] * g.scale : 0.f;
      nl1 = c + 32 < g.L ? lse_bh[c + 32] * kLog2e : 0.f;
      nd1 = c + 32 < g.L ? delta_bh[c + 32] * g.scale : 0.f;
    };
    if (MODE_KV && gq < n_t) fetch_cols(gq);
    for (int i = gq; i < n_t; i += 2) {
      const uint32_t xph = (i >> 1) & 1;
      const int c0 = i * 64;  // first streamed index (q for MODE_KV, kv otherwise) of this tile
      if (MODE_KV) {
        __syncwarp();
        myws[lane] = nl0;
        myws[32 + lane] = nl1;
        myws[64 + lane] = nd0;
        myws[96 + lane] = nd1;
        __syncwarp();
        if (i + 2 < n_t) fetch_cols(i + 2);
      }
      mbar_wait(&x_full[gq], xph, 25000 + i + 100000 * MODE_KV);
      tc_fence_after();
      const int nvalid = g.L - c0;
      uint32_t pp[32], dd[32];  // packed bf16 pairs of the 64 columns: P and dS
#pragma unroll
      for (int hlf = 0; hlf < 2; ++hlf) {
        uint32_t sv[32], dv[32];
        tmem_ld_32x32(tX0[gq] + lane_off + hlf * 32, sv);
        tmem_ld_32x32(tX1[gq] + lane_off + hlf * 32, dv);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 32; k += 4) {
          float l4[4], d4[4];
          if (MODE_KV) {
            *reinterpret_cast<float4*>(l4) = *reinterpret_cast<const float4*>(myws + hlf * 32 + k);
            *reinterpret_cast<float4*>(d4) = *reinterpret_cast<const float4*>(myws + 64 + hlf * 32 + k);
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              l4[e] = my_lse2;
              d4[e] = my_dls;
            }
          }
          float p4[4], s4[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float p = ex2(__uint_as_float(sv[k + e]) * c2 - l4[e]);
            if (nvalid < 64) p = (hlf * 32 + k + e < nvalid) ? p : 0.f;  // only the last tile has missing columns
            p4[e] = p;
            s4[e] = p * (__uint_as_float(dv[k + e]) * g.scale - d4[e]);
          }
          pp[hlf * 16 + k / 2] = pack_bf16x2(p4[0], p4[1]);
          pp[hlf * 16 + k / 2 + 1] = pack_bf16x2(p4[2], p4[3]);
          dd[hlf * 16 + k / 2] = pack_bf16x2(s4[0], s4[1]);
          dd[hlf * 16 + k / 2 + 1] = pack_bf16x2(s4[2], s4[3]);
        }
      }
      // bf16 pairs back into TMEM over the columns this thread just read (S -> P, dP -> dS): 64 values = 32 columns
      if (MODE_KV) tmem_st_32x32(tX0[gq] + lane_off, pp);
      tmem_st_32x32(tX1[gq] + lane_off, dd);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&pb_full[gq]);
    }
    mbar_wait(done_bar, 0, 27);
    tc_fence_after();
    // epilogue: the 8 warps split the accumulator columns (gq = column half)
#pragma unroll 1
    for (int which = 0; which < (MODE_KV ? 2 : 1); ++which) {
      const uint32_t ta = which == 0 ? tA0 : tA1;
      bf16* out = which == 0 ? g.out0 : g.out1;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(ta + lane_off + gq * 64 + c * 32, v);
        tmem_ld_wait();
        if (ri < g.L) {
          bf16* orow = out + (row_base + ri) * 128 + gq * 64 + c * 32;
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

}  // namespace b200

using namespace b200;

static int attn_maps(b200_ctx* ctx, const void* p, uint64_t rows, uint32_t box_rows, CUtensorMap* out) {
  return make_tmap_bf16_2d(ctx, out, p, rows, 128, 128, 64, box_rows);
}

extern "C" int b200_attn_fwd(b200_ctx* ctx, const void* Q, const void* K, const void* V, void* o0, int ld0, void* o1, int ld1,
                             void* lse, int B, int H, int L, int split, float scale, void* stream) {
  return b200_attn_fwd_xd(ctx, Q, K, V, o0, ld0, o1, ld1, lse, B, H, L, L, split, scale, 128, stream);
}

extern "C" int b200_attn_fwd_x(b200_ctx* ctx, const void* Q, const void* K, const void* V, void* o0, int ld0, void* o1,
                               int ld1, void* lse, int B, int H, int L, int Lk, int split, float scale, void* stream) {
  return b200_attn_fwd_xd(ctx, Q, K, V, o0, ld0, o1, ld1, lse, B, H, L, Lk, split, scale, 128, stream);
}

extern "C" int b200_attn_fwd_xd(b200_ctx* ctx, const void* Q, const void* K, const void* V, void* o0, int ld0, void* o1,
                                int ld1, void* lse, int B, int H, int L, int Lk, int split, float scale, int head_live,
                                void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  B200_REQUIRE(head_live == 128 || head_live == 64, "b200_attn_fwd: head_live %d (128, or 64 = zero-padded half heads)", head_live);
  B200_REQUIRE(Q && K && V && o1 && lse && B > 0 && H > 0 && L > 0 && Lk > 0, "b200_attn_fwd: bad args");
  B200_REQUIRE(split >= 0 && split <= L && (split == 0 || o0 != nullptr), "b200_attn_fwd: bad split %d", split);
  B200_REQUIRE(ld1 % 8 == 0 && (split == 0 || ld0 % 8 == 0), "b200_attn_fwd: output leading dims must be multiples of 8");
  const uint64_t rows = static_cast<uint64_t>(B) * H * L, rows_kv = static_cast<uint64_t>(B) * H * Lk;
  B200_REQUIRE(rows < (1ull << 31) && rows_kv < (1ull << 31), "b200_attn_fwd: too many rows");
  CUtensorMap tq, tk, tv;
  if ((rc = attn_maps(ctx, Q, rows, 128, &tq))) return rc;
  if ((rc = attn_maps(ctx, K, rows_kv, 128, &tk))) return rc;
  if ((rc = attn_maps(ctx, V, rows_kv, 64, &tv))) return rc;
  static bool configured = false;
  // same-box log gpurun_out/r2_trip.log (round 2, 24 x 4608 x 128): variant 3 (attention_r2.cu: packed FFMA2 softmax, scale
  // folded into the exponent FFMA) 255.3 us = 1022 TF/s; variant 4 (16 softmax warps) 256.3 us; variant 1 278.2 us;
  // variant 2 (two streams) 304 us in round 1.  Variant 5 (attention_pp.cu: two Q tiles per CTA in ping-pong, one softmax
  // thread per row) 221.6 us = 1177 TF/s vs 257.8 us for variant 3 in the same log (gpurun_out/r2_pp.log); parity and unit
  // tests green with it.  Variant 6 = the same kernel with the OPTIMISTIC running maximum (no row-max pass on the common path,
  // whole-row P held in registers until the partial sum has been checked): 209.9 us = 1243 TF/s with all exponentials on MUFU
  // vs 221.3 us for variant 5 in the same log (gpurun_out/r2_pp2.log).  B200_ATTN_FWD=1..5 selects the others for A/B runs.
  static int variant = 6;
  if (!configured) {
    B200_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmem));
    B200_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmem));
    const char* e = getenv("B200_ATTN_FWD");
    if (e && atoi(e) >= 1 && atoi(e) <= 7) variant = atoi(e);
    configured = true;
  }
  AttnFwdArgs a{(bf16*)o0, ld0, (bf16*)o1, ld1, (float*)lse, B, H, L, split, scale, Lk, head_live};
  B200_REQUIRE(head_live == 128 || variant == 6, "b200_attn_fwd: head_live 64 needs forward variant 6 (the default)");
  B200_REQUIRE(Lk == L || variant >= 3, "b200_attn_fwd: cross attention (Lk != L) needs forward variant 3, 4 or 5");
  if (variant == 7) {
    if ((rc = attn_fwd_pp3_launch(tq, tk, tv, a, reinterpret_cast<cudaStream_t>(stream)))) return rc;
    ctx->launches.fetch_add(1);
    return B200_OK;
  }
  if (variant == 6) {
    if ((rc = attn_fwd_pp2_launch(tq, tk, tv, a, reinterpret_cast<cudaStream_t>(stream)))) return rc;
    ctx->launches.fetch_add(1);
    return B200_OK;
  }
  if (variant == 5) {
    if ((rc = attn_fwd_pp_launch(tq, tk, tv, a, reinterpret_cast<cudaStream_t>(stream)))) return rc;
    ctx->launches.fetch_add(1);
    return B200_OK;
  }
  if (variant >= 3) {
    if ((rc = attn_fwd_r2_launch(variant, tq, tk, tv, a, reinterpret_cast<cudaStream_t>(stream)))) return rc;
    ctx->launches.fetch_add(1);
    return B200_OK;
  }
  dim3 grid((L + 127) / 128, B * H);
  if (variant == 1)
    B200_KLAUNCH(attn_fwd_kernel, grid, kAttnThreads, kFwdSmem, reinterpret_cast<cudaStream_t>(stream), tq, tk, tv, a);
  else
    B200_KLAUNCH(attn_fwd2_kernel, grid, kAttnThreads, kFwdSmem, reinterpret_cast<cudaStream_t>(stream), tq, tk, tv, a);
  B200_CUDA_CHECK(cudaGetLastError());
  ctx->launches.fetch_add(1);
  return B200_OK;
}

extern "C" int b200_attn_bwd(b200_ctx* ctx, const void* Q, const void* K, const void* V, const void* o0, int ld0,
                             const void* o1, int ld1, const void* do0, int ldd0, const void* do1, int ldd1, const void* lse,
                             void* delta, void* dOh, void* dQ, void* dK, void* dV, int B, int H, int L, int split,
                             float scale, void* stream) {
  return b200_attn_bwd_xd(ctx, Q, K, V, o0, ld0, o1, ld1, do0, ldd0, do1, ldd1, lse, delta, dOh, dQ, dK, dV, B, H, L, L, split,
                          scale, 128, stream);
}

extern "C" int b200_attn_bwd_x(b200_ctx* ctx, const void* Q, const void* K, const void* V, const void* o0, int ld0,
                               const void* o1, int ld1, const void* do0, int ldd0, const void* do1, int ldd1, const void* lse,
                               void* delta, void* dOh, void* dQ, void* dK, void* dV, int B, int H, int L, int Lk, int split,
                               float scale, void* stream) {
  return b200_attn_bwd_xd(ctx, Q, K, V, o0, ld0, o1, ld1, do0, ldd0, do1, ldd1, lse, delta, dOh, dQ, dK, dV, B, H, L, Lk, split,
                          scale, 128, stream);
}

extern "C" int b200_attn_bwd_xd(b200_ctx* ctx, const void* Q, const void* K, const void* V, const void* o0, int ld0,
                                const void* o1, int ld1, const void* do0, int ldd0, const void* do1, int ldd1, const void* lse,
                                void* delta, void* dOh, void* dQ, void* dK, void* dV, int B, int H, int L, int Lk, int split,
                                float scale, int head_live, void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  B200_REQUIRE(head_live == 128 || head_live == 64, "b200_attn_bwd: head_live %d (128, or 64 = zero-padded half heads)", head_live);
  B200_REQUIRE(Q && K && V && o1 && do1 && lse && delta && dOh && dQ && dK && dV, "b200_attn_bwd: null argument");
  B200_REQUIRE(B > 0 && H > 0 && L > 0 && Lk > 0 && split >= 0 && split <= L && (split == 0 || (o0 && do0)), "b200_attn_bwd: bad shape");
  B200_REQUIRE(ld1 % 4 == 0 && ldd1 % 4 == 0 && ld0 % 4 == 0 && ldd0 % 4 == 0, "b200_attn_bwd: leading dims %% 4");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const uint64_t rows = static_cast<uint64_t>(B) * H * L, rows_kv = static_cast<uint64_t>(B) * H * Lk;
  B200_REQUIRE(rows < (1ull << 31) && rows_kv < (1ull << 31), "b200_attn_bwd: too many rows");
  const long long warps = static_cast<long long>(B) * L * H;
  B200_KLAUNCH(attn_delta_kernel, static_cast<unsigned>((warps + 7) / 8), 256, 0, st, (const bf16*)o0, ld0, (const bf16*)o1, ld1,
                                                                            (const bf16*)do0, ldd0, (const bf16*)do1, ldd1,
                                                                            (float*)delta, (bf16*)dOh, B, H, L, split);
  B200_CUDA_CHECK(cudaGetLastError());
  CUtensorMap q128, k128, v128, d128, q64, k64, v64, d64;
  if ((rc = attn_maps(ctx, Q, rows, 128, &q128))) return rc;
  if ((rc = attn_maps(ctx, K, rows_kv, 128, &k128))) return rc;
  if ((rc = attn_maps(ctx, V, rows_kv, 128, &v128))) return rc;
  if ((rc = attn_maps(ctx, dOh, rows, 128, &d128))) return rc;
  if ((rc = attn_maps(ctx, Q, rows, 64, &q64))) return rc;
  if ((rc = attn_maps(ctx, K, rows_kv, 64, &k64))) return rc;
  if ((rc = attn_maps(ctx, V, rows_kv, 64, &v64))) return rc;
  if ((rc = attn_maps(ctx, dOh, rows, 64, &d64))) return rc;
  static bool configured = false;
  // same-box log gpurun_out/r2_trip.log (round 2, 24 x 4608 x 128): variant 2 (attention_r2.cu: LDS, deferred prefetch
  // scaling, 16-column TMEM chunks) 787.7 us = 828 TF/s algorithmic; 3: 797.3; 5: 813.1; 4: 817.3; variant 1: 888.1 us
  static int variant = 2;
  if (!configured) {
    B200_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem));
    B200_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem));
    const char* e = getenv("B200_ATTN_BWD");
    if (e && atoi(e) >= 1 && atoi(e) <= 5) variant = atoi(e);
    configured = true;
  }
  dim3 grid((L + 127) / 128, B * H);
  // dK/dV pass: keys / values stationary (length Lk), queries streamed (L); dQ pass: the other way round
  static int bwd_dbg = -1;
  if (bwd_dbg < 0) {
    const char* e = getenv("B200_ATTN_BWD_DBG");
    bwd_dbg = e ? atoi(e) : 0;
  }
  AttnBwdArgs akv{(const float*)lse, (const float*)delta, nullptr, nullptr, (bf16*)dV, (bf16*)dK, Lk, scale, L, bwd_dbg, head_live};
  AttnBwdArgs aq{(const float*)lse, (const float*)delta, (const bf16*)Q, (const bf16*)dOh, (bf16*)dQ, nullptr, L, scale, Lk, bwd_dbg, head_live};
  B200_REQUIRE(head_live == 128 || ((variant == 2 || variant == 3) && bwd_dbg != 3),
               "b200_attn_bwd: head_live 64 needs backward variant 2 (the default) or 3");
  B200_REQUIRE(Lk == L || variant == 2 || variant == 3, "b200_attn_bwd: cross attention (Lk != L) needs backward variant 2 or 3");
  if (variant >= 2) {
    if ((rc = attn_bwd_r2_launch(variant, k128, v128, q64, d64, q128, d128, k64, v64, akv, aq, B, H, st))) return rc;
    ctx->launches.fetch_add(3);
    return B200_OK;
  }
  B200_KLAUNCH(attn_bwd_kernel<1>, grid, kAttnThreads, kBwdSmem, st, k128, v128, q64, d64, akv);
  B200_CUDA_CHECK(cudaGetLastError());
  B200_KLAUNCH(attn_bwd_kernel<0>, grid, kAttnThreads, kBwdSmem, st, q128, d128, k64, v64, aq);
  B200_CUDA_CHECK(cudaGetLastError());
  ctx->launches.fetch_add(3);
  return B200_OK;
}
