// Batch-preparation, loss and layout kernels added for the UNet / eps-prediction side of the path (all HBM-bound):
//   * DDPM add_noise                      diffusers DDPMScheduler.add_noise as configured by toolkit/sampler.py:31-50
//                                         (scaled-linear betas 0.00085 -> 0.012, 1000 steps), called through
//                                         toolkit/stable_diffusion_model.py:1854-1876
//   * the general training loss           extensions_built_in/sd_trainer/SDTrainer.py:522-1052 default ('mse') path:
//                                         target = flow (noise - latents, :644-646) | eps (noise, :650) | v (:623-625) |
//                                         a given tensor (prior prediction, :619-621); per-sample weights (timestep weights
//                                         :923-943, loss_multiplier :994, SNR-gamma :1001-1011 -- all scalars per sample);
//                                         mask multiplier (:953-959)
//   * NCHW <-> rows, im2col / col2im      Conv2d LoRA (toolkit/lora_special.py:95-104: down = Conv k x k, up = 1 x 1) and
//                                         the frozen conv it wraps as ONE fused GEMM over rows [B Ho Wo, C kh kw]
#include "common.cuh"
#include "ctx.h"

namespace b200 {

__device__ __forceinline__ size_t packed_index2(int b, int c, int h, int w, int C, int H, int W) {
  const int h2 = h >> 1, ph = h & 1, w2 = w >> 1, pw = w & 1;
  return ((static_cast<size_t>(b) * (H >> 1) + h2) * (W >> 1) + w2) * (C * 4) + c * 4 + ph * 2 + pw;
}

// noisy = sqrt(ac[t]) x0 + sqrt(1 - ac[t]) noise, in the arithmetic of a bf16 torch tensor expression:
//   ac = bf16(alphas_cumprod[t]); sa = bf16(sqrt(ac)); sb = bf16(sqrt(bf16(1 - ac)));
//   noisy = bf16( bf16(sa x0) + bf16(sb noise) )
// The timestep is an INTEGER index into the table (bit-exact index op).
__global__ void __launch_bounds__(256) ddpm_add_noise_kernel(const bf16* __restrict__ x0, const bf16* __restrict__ noise,
                                                             const long long* __restrict__ t,
                                                             const float* __restrict__ alphas_cumprod, int n_train,
                                                             bf16* __restrict__ out, long long per) {
  pdl_grid_sync();
  const int b = blockIdx.y;
  long long ti = t[b];
  ti = ti < 0 ? 0 : (ti >= n_train ? n_train - 1 : ti);
  const float ac = bf16_round(alphas_cumprod[ti]);
  const float sa = bf16_round(sqrtf(ac));
  const float sb = bf16_round(sqrtf(bf16_round(1.0f - ac)));
  const long long i2 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i2 * 2 >= per) return;
  const long long i = static_cast<long long>(b) * per + i2 * 2;
  const float2 a = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(x0 + i));
  const float2 n = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(noise + i));
  const float v0 = bf16_round(sa * a.x) + bf16_round(sb * n.x);
  const float v1 = bf16_round(sa * a.y) + bf16_round(sb * n.y);
  *reinterpret_cast<uint32_t*>(out + i) = pack_bf16x2(v0, v1);
}

// target = given | bf16( bf16(cn noise) - bf16(cl x0) )   (cn = cl = 1: flow matching; cn = 1, cl = 0: eps; v: both from
// the scheduler table);  d = pred - target;  loss_b = w_b mean(d^2 mask);  loss = mean_b loss_b;
// dpred = bf16( 2 d mask w_b gscale / (per B) )
__global__ void __launch_bounds__(256) train_loss_kernel(const bf16* __restrict__ pred, const bf16* __restrict__ x0,
                                                         const bf16* __restrict__ noise, const bf16* __restrict__ target,
                                                         const float* __restrict__ coef_noise,
                                                         const float* __restrict__ coef_latent,
                                                         const float* __restrict__ sample_weight,
                                                         const float* __restrict__ mask, int mask_channels,
                                                         bf16* __restrict__ dpred, float* __restrict__ loss_per_sample,
                                                         float* __restrict__ loss_total, int B, int C, int H, int W, int pack,
                                                         float gscale) {
  pdl_grid_sync();
  const long long i2 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long per = static_cast<long long>(C) * H * W;
  const int b = blockIdx.y;
  const float wb = sample_weight ? sample_weight[b] : 1.0f;
  float ss = 0.f;
  if (i2 * 2 < per) {
    const long long rem = i2 * 2;
    const long long i = static_cast<long long>(b) * per + rem;
    const int c = static_cast<int>(rem / (static_cast<long long>(H) * W));
    const int h = static_cast<int>((rem / W) % H);
    const int w = static_cast<int>(rem % W);
    const size_t o = (pack & 1) ? packed_index2(b, c, h, w, C, H, W) : static_cast<size_t>(i);
    float t0, t1;
    if (target) {  // pack & 2: the target lies in the prediction's (packed) layout, e.g. a prior prediction of the same model
      const float2 tg = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(target + ((pack & 2) ? o : static_cast<size_t>(i))));
      t0 = tg.x;
      t1 = tg.y;
    } else {
      const float2 a = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(x0 + i));
      const float2 n = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(noise + i));
      const float cn = coef_noise ? coef_noise[b] : 1.0f;
      const float cl = coef_latent ? coef_latent[b] : 1.0f;
      t0 = bf16_round(bf16_round(cn * n.x) - bf16_round(cl * a.x));
      t1 = bf16_round(bf16_round(cn * n.y) - bf16_round(cl * a.y));
    }
    const float2 p = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(pred + o));
    float m0 = 1.0f, m1 = 1.0f;
    if (mask) {
      const size_t mi = (static_cast<size_t>(b) * mask_channels + (mask_channels == 1 ? 0 : c)) * H * W +
                        static_cast<size_t>(h) * W + w;
      m0 = mask[mi];
      m1 = mask[mi + 1];
    }
    const float d0 = p.x - t0, d1 = p.y - t1;
    ss = d0 * d0 * m0 + d1 * d1 * m1;
    const float k = 2.0f * gscale * wb / (static_cast<float>(per) * static_cast<float>(B));
    if (dpred) *reinterpret_cast<uint32_t*>(dpred + o) = pack_bf16x2(d0 * m0 * k, d1 * m1 * k);
  }
  __shared__ float red[8];
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < 8 ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) {
      atomicAdd(loss_per_sample + b, wb * v / static_cast<float>(per));
      atomicAdd(loss_total, wb * v / (static_cast<float>(per) * static_cast<float>(B)));
    }
  }
}

// NCHW [B, C, HW] <-> rows [B HW, C] through a 32 x 32 shared-memory tile (both sides coalesced).
// to_rows = 1: out rows (ld = ldr) from src NCHW; to_rows = 0: out NCHW from src rows.
__global__ void __launch_bounds__(256) nchw_rows_kernel(const bf16* __restrict__ src, bf16* __restrict__ dst, int C, int HW,
                                                        int ldr, int to_rows) {
  pdl_grid_sync();
  __shared__ bf16 tile[32][33];
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 8 rows of 32 threads
  if (to_rows) {
    for (int r = ty; r < 32; r += 8) {  // tile[c][p]: read along p
      const int c = c0 + r, p = p0 + tx;
      if (c < C && p < HW) tile[r][tx] = src[(static_cast<size_t>(b) * C + c) * HW + p];
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {  // write along c
      const int p = p0 + r, c = c0 + tx;
      if (c < C && p < HW) dst[(static_cast<size_t>(b) * HW + p) * ldr + c] = tile[tx][r];
    }
  } else {
    for (int r = ty; r < 32; r += 8) {  // tile[p][c]: read along c
      const int p = p0 + r, c = c0 + tx;
      if (c < C && p < HW) tile[r][tx] = src[(static_cast<size_t>(b) * HW + p) * ldr + c];
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {  // write along p
      const int c = c0 + r, p = p0 + tx;
      if (c < C && p < HW) dst[(static_cast<size_t>(b) * C + c) * HW + p] = tile[tx][r];
    }
  }
}

struct ConvGeom {
  int B, C, H, W, kh, kw, sh, sw, ph, pw, Ho, Wo, ld;  // ld: row stride of the column matrix (>= C kh kw, elements)
};

// cols[(b, oy, ox), (c, ky, kx)] = x[b, c, oy sh + ky - ph, ox sw + kx - pw]  (zero outside); columns >= C kh kw (row
// padding up to ld) are zero-filled.  One thread per column element, columns fastest (coalesced writes; reads hit L1 / L2).
__global__ void __launch_bounds__(256) im2col_kernel(const bf16* __restrict__ x, bf16* __restrict__ cols, const ConvGeom g) {
  pdl_grid_sync();
  const long long rows = static_cast<long long>(g.B) * g.Ho * g.Wo;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= rows * g.ld) return;
  const long long row = idx / g.ld;
  const int col = static_cast<int>(idx % g.ld);
  const int kk = g.kh * g.kw;
  bf16 v = __float2bfloat16_rn(0.f);
  if (col < g.C * kk) {
    const int c = col / kk, ky = (col % kk) / g.kw, kx = col % g.kw;
    const int b = static_cast<int>(row / (g.Ho * g.Wo));
    const int oy = static_cast<int>((row / g.Wo) % g.Ho), ox = static_cast<int>(row % g.Wo);
    const int y = oy * g.sh + ky - g.ph, xx = ox * g.sw + kx - g.pw;
    if (y >= 0 && y < g.H && xx >= 0 && xx < g.W) v = x[((static_cast<size_t>(b) * g.C + c) * g.H + y) * g.W + xx];
  }
  cols[idx] = v;
}

// dx[b, c, y, x] (+)= sum over (ky, kx) with (y + ph - ky) % sh == 0 ... of dcols[(b, oy, ox), (c, ky, kx)]
// gather form (no atomics, deterministic); fp32 accumulate, one bf16 store.  accumulate != 0: added to the existing dx.
__global__ void __launch_bounds__(256) col2im_kernel(const bf16* __restrict__ dcols, bf16* __restrict__ dx, const ConvGeom g,
                                                     int accumulate) {
  pdl_grid_sync();
  const long long n = static_cast<long long>(g.B) * g.C * g.H * g.W;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const int xx = static_cast<int>(idx % g.W), y = static_cast<int>((idx / g.W) % g.H);
  const int c = static_cast<int>((idx / (static_cast<long long>(g.W) * g.H)) % g.C);
  const int b = static_cast<int>(idx / (static_cast<long long>(g.W) * g.H * g.C));
  float acc = accumulate ? __bfloat162float(dx[idx]) : 0.f;
  for (int ky = 0; ky < g.kh; ++ky) {
    const int ty = y + g.ph - ky;
    if (ty < 0 || ty % g.sh != 0) continue;
    const int oy = ty / g.sh;
    if (oy >= g.Ho) continue;
    for (int kx = 0; kx < g.kw; ++kx) {
      const int tx = xx + g.pw - kx;
      if (tx < 0 || tx % g.sw != 0) continue;
      const int ox = tx / g.sw;
      if (ox >= g.Wo) continue;
      const size_t row = (static_cast<size_t>(b) * g.Ho + oy) * g.Wo + ox;
      acc += __bfloat162float(dcols[row * g.ld + (c * g.kh + ky) * g.kw + kx]);
    }
  }
  dx[idx] = __float2bfloat16_rn(acc);
}

// y[i, :] *= s[i / rows_per_sample]-free elementwise mask: z[row, col] = z[row, col] * mask[row, col] (bf16 * fp32 -> bf16);
// the LoRA dropout / rank-dropout masks of toolkit/network_mixins.py:197-239 applied to the rank-side activations
__global__ void __launch_bounds__(256) mask_rows_kernel(bf16* __restrict__ z, int ldz, const float* __restrict__ row_mask,
                                                        int ldm, const float* __restrict__ col_mask, int rows_per_sample,
                                                        long long rows, int cols) {
  pdl_grid_sync();
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= rows * cols) return;
  const long long r = idx / cols;
  const int c = static_cast<int>(idx % cols);
  float v = __bfloat162float(z[r * ldz + c]);
  if (row_mask) v *= row_mask[r * ldm + c];
  if (col_mask) v *= col_mask[(rows_per_sample > 0 ? r / rows_per_sample : 0) * cols + c];
  z[r * ldz + c] = __float2bfloat16_rn(v);
}

}  // namespace b200

using namespace b200;

extern "C" int b200_ddpm_add_noise(b200_ctx* ctx, const void* latents, const void* noise, const void* timesteps_i64,
                                   const void* alphas_cumprod_f32, int n_train, void* out, int B, int64_t per_sample,
                                   void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  B200_REQUIRE(latents && noise && timesteps_i64 && alphas_cumprod_f32 && out && B > 0 && per_sample > 0 && n_train > 0,
               "b200_ddpm_add_noise: bad args");
  B200_REQUIRE(per_sample % 2 == 0, "b200_ddpm_add_noise: elements per sample must be even");
  dim3 grid(static_cast<unsigned>((per_sample / 2 + 255) / 256), B);
  B200_KLAUNCH(ddpm_add_noise_kernel, grid, 256, 0, reinterpret_cast<cudaStream_t>(stream), (const bf16*)latents,
               (const bf16*)noise, (const long long*)timesteps_i64, (const float*)alphas_cumprod_f32, n_train, (bf16*)out,
               static_cast<long long>(per_sample));
  B200_CUDA_CHECK(cudaGetLastError());
  ctx->launches.fetch_add(1);
  return B200_OK;
}

extern "C" int b200_train_loss(b200_ctx* ctx, const void* pred, const void* latents, const void* noise, const void* target,
                               const void* coef_noise, const void* coef_latent, const void* sample_weight, const void* mask,
                               int mask_channels, void* dpred, void* loss_per_sample, void* loss_total, int B, int C, int H,
                               int W, int pack, float gscale, void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  B200_REQUIRE(pred && loss_per_sample && loss_total && B > 0 && C > 0 && H > 0 && W > 0, "b200_train_loss: bad args");
  B200_REQUIRE(target || (latents && noise), "b200_train_loss: needs a target tensor or latents + noise");
  B200_REQUIRE(W % 2 == 0 && (!(pack & 1) || H % 2 == 0), "b200_train_loss: H/W must be even");
  B200_REQUIRE(!mask || mask_channels == 1 || mask_channels == C, "b200_train_loss: mask channels %d (1 or %d)", mask_channels, C);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  B200_CUDA_CHECK(cudaMemsetAsync(loss_per_sample, 0, sizeof(float) * B, st));
  B200_CUDA_CHECK(cudaMemsetAsync(loss_total, 0, sizeof(float), st));
  const long long pairs = static_cast<long long>(C) * H * W / 2;
  dim3 grid(static_cast<unsigned>((pairs + 255) / 256), B);
  B200_KLAUNCH(train_loss_kernel, grid, 256, 0, st, (const bf16*)pred, (const bf16*)latents, (const bf16*)noise,
               (const bf16*)target, (const float*)coef_noise, (const float*)coef_latent, (const float*)sample_weight,
               (const float*)mask, mask_channels, (bf16*)dpred, (float*)loss_per_sample, (float*)loss_total, B, C, H, W, pack,
               gscale);
  B200_CUDA_CHECK(cudaGetLastError());
  ctx->launches.fetch_add(1);
  return B200_OK;
}

extern "C" int b200_nchw_rows(b200_ctx* ctx, const void* src, void* dst, int B, int C, int HW, int ld_rows, int to_rows,
                              void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  B200_REQUIRE(src && dst && B > 0 && C > 0 && HW > 0 && ld_rows >= C, "b200_nchw_rows: bad args");
  dim3 grid((HW + 31) / 32, (C + 31) / 32, B);
  B200_KLAUNCH(nchw_rows_kernel, grid, 256, 0, reinterpret_cast<cudaStream_t>(stream), (const bf16*)src, (bf16*)dst, C, HW,
               ld_rows, to_rows);
  B200_CUDA_CHECK(cudaGetLastError());
  ctx->launches.fetch_add(1);
  return B200_OK;
}

static int conv_geom(ConvGeom& g, int B, int C, int H, int W, int kh, int kw, int sh, int sw, int ph, int pw, int ld) {
  B200_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && kh > 0 && kw > 0 && sh > 0 && sw > 0 && ph >= 0 && pw >= 0,
               "conv geometry: bad args");
  g = ConvGeom{B, C, H, W, kh, kw, sh, sw, ph, pw, (H + 2 * ph - kh) / sh + 1, (W + 2 * pw - kw) / sw + 1, ld};
  B200_REQUIRE(g.Ho > 0 && g.Wo > 0 && ld >= C * kh * kw, "conv geometry: empty output or ld %d < %d", ld, C * kh * kw);
  return B200_OK;
}

extern "C" int b200_im2col(b200_ctx* ctx, const void* x, void* cols, int B, int C, int H, int W, int kh, int kw, int sh, int sw,
                           int ph, int pw, int ld, void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  B200_REQUIRE(x && cols, "b200_im2col: null argument");
  ConvGeom g;
  if ((rc = conv_geom(g, B, C, H, W, kh, kw, sh, sw, ph, pw, ld))) return rc;
  const long long n = static_cast<long long>(B) * g.Ho * g.Wo * ld;
  B200_KLAUNCH(im2col_kernel, static_cast<unsigned>((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream),
               (const bf16*)x, (bf16*)cols, g);
  B200_CUDA_CHECK(cudaGetLastError());
  ctx->launches.fetch_add(1);
  return B200_OK;
}

extern "C" int b200_col2im(b200_ctx* ctx, const void* dcols, void* dx, int B, int C, int H, int W, int kh, int kw, int sh,
                           int sw, int ph, int pw, int ld, int accumulate, void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  B200_REQUIRE(dcols && dx, "b200_col2im: null argument");
  ConvGeom g;
  if ((rc = conv_geom(g, B, C, H, W, kh, kw, sh, sw, ph, pw, ld))) return rc;
  const long long n = static_cast<long long>(B) * C * H * W;
  B200_KLAUNCH(col2im_kernel, static_cast<unsigned>((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream),
               (const bf16*)dcols, (bf16*)dx, g, accumulate);
  B200_CUDA_CHECK(cudaGetLastError());
  ctx->launches.fetch_add(1);
  return B200_OK;
}

extern "C" int b200_mask_rows(b200_ctx* ctx, void* z, int ldz, const void* row_mask, int ldm, const void* col_mask,
                              int rows_per_sample, int64_t rows, int cols, void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  B200_REQUIRE(z && rows > 0 && cols > 0 && (row_mask || col_mask), "b200_mask_rows: bad args");
  const long long n = rows * cols;
  B200_KLAUNCH(mask_rows_kernel, static_cast<unsigned>((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream),
               (bf16*)z, ldz, (const float*)row_mask, ldm, (const float*)col_mask, rows_per_sample, static_cast<long long>(rows),
               cols);
  B200_CUDA_CHECK(cudaGetLastError());
  ctx->launches.fetch_add(1);
  return B200_OK;
}
