// Batch preparation, loss and optimizer kernels of the LoRA training step (all HBM-bound, fp32 math):
//   * flow-matching add_noise fused with FLUX 2x2 patch packing
//       toolkit/samplers/custom_flowmatch_sampler.py:91-102, toolkit/stable_diffusion_model.py:2157-2172
//   * target = noise - latents, per-sample MSE, and dLoss/dPred in packed layout
//       extensions_built_in/sd_trainer/SDTrainer.py:644-646, :916, :987-990, :1013
//   * global grad-norm, clip, AdamW (eps as given by the caller, reference uses 1e-6), optional EMA
//       SDTrainer.py:2278-2283, toolkit/optimizer.py:78-79, toolkit/ema.py:100-152
//   * fp32 master -> bf16 padded operand copies of lora_down / lora_up for the tensor-core GEMMs
#include "common.cuh"
#include "ctx.h"

namespace b200 {

// index of element (b, c, h, w) of a [B, C, H, W] latent in the packed [B, (H/2)(W/2), C*4] layout
__device__ __forceinline__ size_t packed_index(int b, int c, int h, int w, int C, int H, int W) {
  const int h2 = h >> 1, ph = h & 1, w2 = w >> 1, pw = w & 1;
  return ((static_cast<size_t>(b) * (H >> 1) + h2) * (W >> 1) + w2) * (C * 4) + c * 4 + ph * 2 + pw;
}

// noisy = bf16( (1 - t/1000) * x0 + (t/1000) * noise )   [fp32 math, as the fp32 timestep tensor promotes it]
__global__ void flow_add_noise_kernel(const bf16* __restrict__ x0, const bf16* __restrict__ noise, const float* __restrict__ t,
                                      bf16* __restrict__ out, int B, int C, int H, int W, int pack) {
  pdl_grid_sync();
  const long long i2 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;  // pair index
  const long long per = static_cast<long long>(C) * H * W;
  if (i2 * 2 >= per * B) return;
  const long long i = i2 * 2;
  const int b = static_cast<int>(i / per);
  const long long rem = i % per;
  const int c = static_cast<int>(rem / (H * W));
  const int h = static_cast<int>((rem / W) % H);
  const int w = static_cast<int>(rem % W);
  const float tt = t[b] / 1000.0f;
  const float2 a = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(x0 + i));
  const float2 n = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(noise + i));
  const float v0 = (1.0f - tt) * a.x + tt * n.x;
  const float v1 = (1.0f - tt) * a.y + tt * n.y;
  const size_t o = pack ? packed_index(b, c, h, w, C, H, W) : static_cast<size_t>(i);
  *reinterpret_cast<uint32_t*>(out + o) = pack_bf16x2(v0, v1);
}

// target = bf16(noise - x0);  d = pred - target;  loss += d^2 / (per * B);  dpred = bf16(2 d gscale / (per * B))
__global__ void __launch_bounds__(256) flow_loss_kernel(const bf16* __restrict__ pred, const bf16* __restrict__ x0,
                                                        const bf16* __restrict__ noise, bf16* __restrict__ dpred,
                                                        float* __restrict__ loss_per_sample, float* __restrict__ loss_total,
                                                        int B, int C, int H, int W, int pack, float gscale) {
  pdl_grid_sync();
  const long long i2 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long per = static_cast<long long>(C) * H * W;
  const int b = blockIdx.y;
  float ss = 0.f;
  if (i2 * 2 < per) {
    const long long rem = i2 * 2;
    const long long i = static_cast<long long>(b) * per + rem;
    const int c = static_cast<int>(rem / (H * W));
    const int h = static_cast<int>((rem / W) % H);
    const int w = static_cast<int>(rem % W);
    const float2 a = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(x0 + i));
    const float2 n = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(noise + i));
    const size_t o = pack ? packed_index(b, c, h, w, C, H, W) : static_cast<size_t>(i);
    const float2 p = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(pred + o));
    const float d0 = p.x - bf16_round(n.x - a.x);
    const float d1 = p.y - bf16_round(n.y - a.y);
    ss = d0 * d0 + d1 * d1;
    const float k = 2.0f * gscale / (static_cast<float>(per) * static_cast<float>(B));
    if (dpred) *reinterpret_cast<uint32_t*>(dpred + o) = pack_bf16x2(d0 * k, d1 * k);
  }
  __shared__ float red[8];
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < 8 ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) {
      atomicAdd(loss_per_sample + b, v / static_cast<float>(per));
      atomicAdd(loss_total, v / (static_cast<float>(per) * static_cast<float>(B)));
    }
  }
}

// sum of squares of an fp32 vector -> double accumulator (caller zeroes it)
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, long long n, double* __restrict__ out) {
  pdl_grid_sync();
  float acc = 0.f;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x * 4;
  for (long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 3 < n) {
      const float4 v = *reinterpret_cast<const float4*>(g + i);
      acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    } else {
      for (long long j = i; j < n; ++j) acc += g[j] * g[j];
    }
  }
  __shared__ float red[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < 8 ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) atomicAdd(out, static_cast<double>(v));
  }
}

// hyper: [0] lr, [1] beta1, [2] beta2, [3] eps, [4] weight_decay, [5] max_norm (<=0: no clipping),
//        [6] ema_decay (<=0: no EMA), [7] grad_prescale (e.g. 1/world when the all-reduce summed)
// state: 64 bytes; int64 [0] = optimizer steps taken so far; int64 [1] != 0: EMA warm-up (use_num_updates); floats at
//        byte 16: derived per-step scalars
//        d[0] grad scale (clip * prescale), d[1] step_size = lr / (1 - b1^t), d[2] sqrt(1 - b2^t),
//        d[3] EMA decay of this step, d[4] total grad norm (after prescale, before clipping)
__global__ void adamw_prepare_kernel(const double* __restrict__ sumsq, const float* __restrict__ hyper, long long* state,
                                     float* __restrict__ norm_out) {
  pdl_grid_sync();
  float* d = reinterpret_cast<float*>(reinterpret_cast<char*>(state) + 16);
  const long long step = state[0] + 1;  // 1-based index of this step, as torch.optim.AdamW counts
  state[0] = step;
  const double lr = hyper[0], b1 = hyper[1], b2 = hyper[2];
  const float max_norm = hyper[5], ema_decay = hyper[6], pre = hyper[7];
  const float total_norm = static_cast<float>(sqrt(*sumsq)) * pre;
  float clip = 1.0f;
  if (max_norm > 0.f) clip = fminf(1.0f, max_norm / (total_norm + 1e-6f));  // torch.nn.utils.clip_grad_norm_
  d[0] = clip * pre;
  d[1] = static_cast<float>(lr / (1.0 - pow(b1, static_cast<double>(step))));
  d[2] = static_cast<float>(sqrt(1.0 - pow(b2, static_cast<double>(step))));
  // toolkit/ema.py:121-128: the warm-up min(decay, (1 + n) / (10 + n)) exists only with `use_num_updates=True`; the
  // trainer builds its EMA without it (BaseSDTrainProcess.py:798-803), so the default (state[1] == 0) is a constant decay
  const float nn = static_cast<float>(step);  // num_updates after its increment
  const bool warm_up = state[1] != 0;
  d[3] = ema_decay > 0.f ? (warm_up ? fminf(ema_decay, (1.0f + nn) / (10.0f + nn)) : ema_decay) : 0.f;
  d[4] = total_norm;
  if (norm_out) *norm_out = total_norm;
}

__global__ void __launch_bounds__(256) clip_adamw_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                         float* __restrict__ m, float* __restrict__ v,
                                                         float* __restrict__ ema, const float* __restrict__ hyper,
                                                         const long long* __restrict__ state, long long n) {
  pdl_grid_sync();
  const float* d = reinterpret_cast<const float*>(reinterpret_cast<const char*>(state) + 16);
  const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], wd = hyper[4];
  const float gs = d[0], step_size = d[1], bc2_sqrt = d[2], ema_d = d[3];
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i] * gs;
  float pi = p[i] * (1.0f - lr * wd);
  const float mi = b1 * m[i] + (1.0f - b1) * gi;
  const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  pi -= step_size * (mi / denom);
  p[i] = pi;
  m[i] = mi;
  v[i] = vi;
  if (ema != nullptr && ema_d > 0.f) {
    const float e = ema[i];
    ema[i] = e - (1.0f - ema_d) * (e - pi);
  }
}

struct RepackEntry {
  long long src_off;  // element offset into the flat fp32 parameter buffer
  long long dst_off;  // element offset into the bf16 pack buffer
  int rows, cols;     // source tensor shape [rows, cols]
  int dst_ld;         // destination leading dimension (>= cols)
  int pad;
};

__global__ void __launch_bounds__(256) repack_kernel(const float* __restrict__ flat, bf16* __restrict__ pack,
                                                     const RepackEntry* __restrict__ tab) {
  pdl_grid_sync();
  const RepackEntry e = tab[blockIdx.y];
  const long long n = static_cast<long long>(e.rows) * e.cols;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / e.cols), c = static_cast<int>(i % e.cols);
    pack[e.dst_off + static_cast<long long>(r) * e.dst_ld + c] = __float2bfloat16_rn(flat[e.src_off + i]);
  }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_flow_add_noise(b200_ctx* ctx, const void* latents, const void* noise, const void* t, void* out, int B,
                                   int C, int H, int W, int pack, void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  B200_REQUIRE(latents && noise && t && out && B > 0 && C > 0 && H > 0 && W > 0, "b200_flow_add_noise: bad args");
  B200_REQUIRE(W % 2 == 0 && (!pack || H % 2 == 0), "b200_flow_add_noise: H/W must be even");
  const long long pairs = static_cast<long long>(B) * C * H * W / 2;
  B200_KLAUNCH(flow_add_noise_kernel, static_cast<unsigned>((pairs + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream), 
      (const bf16*)latents, (const bf16*)noise, (const float*)t, (bf16*)out, B, C, H, W, pack);
  B200_CUDA_CHECK(cudaGetLastError());
  ctx->launches.fetch_add(1);
  return B200_OK;
}

extern "C" int b200_flow_loss(b200_ctx* ctx, const void* pred, const void* latents, const void* noise, void* dpred,
                              void* loss_per_sample, void* loss_total, int B, int C, int H, int W, int pack, float gscale,
                              void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  B200_REQUIRE(pred && latents && noise && loss_per_sample && loss_total && B > 0, "b200_flow_loss: bad args");
  B200_REQUIRE(W % 2 == 0 && (!pack || H % 2 == 0), "b200_flow_loss: H/W must be even");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  B200_CUDA_CHECK(cudaMemsetAsync(loss_per_sample, 0, sizeof(float) * B, st));
  B200_CUDA_CHECK(cudaMemsetAsync(loss_total, 0, sizeof(float), st));
  const long long pairs = static_cast<long long>(C) * H * W / 2;
  dim3 grid(static_cast<unsigned>((pairs + 255) / 256), B);
  B200_KLAUNCH(flow_loss_kernel, grid, 256, 0, st, (const bf16*)pred, (const bf16*)latents, (const bf16*)noise, (bf16*)dpred,
                                         (float*)loss_per_sample, (float*)loss_total, B, C, H, W, pack, gscale);
  B200_CUDA_CHECK(cudaGetLastError());
  ctx->launches.fetch_add(1);
  return B200_OK;
}

extern "C" int b200_grad_sumsq(b200_ctx* ctx, const void* g, int64_t n, void* sumsq_f64, void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  B200_REQUIRE(g && sumsq_f64 && n > 0, "b200_grad_sumsq: bad args");
  B200_REQUIRE((reinterpret_cast<uintptr_t>(g) & 15) == 0, "b200_grad_sumsq: g must be 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  B200_CUDA_CHECK(cudaMemsetAsync(sumsq_f64, 0, sizeof(double), st));
  long long blocks = (n / 4 + 255) / 256;
  const long long cap = static_cast<long long>(ctx->sm_count) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  B200_KLAUNCH(sumsq_kernel, static_cast<unsigned>(blocks), 256, 0, st, (const float*)g, n, (double*)sumsq_f64);
  B200_CUDA_CHECK(cudaGetLastError());
  ctx->launches.fetch_add(1);
  return B200_OK;
}

extern "C" int b200_clip_adamw(b200_ctx* ctx, void* p, void* g, void* m, void* v, void* ema, const void* sumsq_f64,
                               const void* hyper, void* state, int64_t n, void* norm_out, void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  B200_REQUIRE(p && g && m && v && sumsq_f64 && hyper && state && n > 0, "b200_clip_adamw: bad args");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  B200_KLAUNCH(adamw_prepare_kernel, 1, 1, 0, st, (const double*)sumsq_f64, (const float*)hyper, (long long*)state, (float*)norm_out);
  B200_CUDA_CHECK(cudaGetLastError());
  B200_KLAUNCH(clip_adamw_kernel, static_cast<unsigned>((n + 255) / 256), 256, 0, st, 
      (float*)p, (const float*)g, (float*)m, (float*)v, (float*)ema, (const float*)hyper, (const long long*)state, n);
  B200_CUDA_CHECK(cudaGetLastError());
  ctx->launches.fetch_add(2);
  return B200_OK;
}

extern "C" int b200_repack_lora(b200_ctx* ctx, const void* flat_f32, void* pack_bf16, const void* table, int n_entries,
                                void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  B200_REQUIRE(flat_f32 && pack_bf16 && table && n_entries > 0, "b200_repack_lora: bad args");
  static_assert(sizeof(RepackEntry) == 32, "RepackEntry layout is part of the C ABI");
  dim3 grid(16, n_entries);
  B200_KLAUNCH(repack_kernel, grid, 256, 0, reinterpret_cast<cudaStream_t>(stream), (const float*)flat_f32, (bf16*)pack_bf16,
                                                                          (const RepackEntry*)table);
  B200_CUDA_CHECK(cudaGetLastError());
  ctx->launches.fetch_add(1);
  return B200_OK;
}
