// Micro-benchmark (round-2 decision input): how fast can CTAs add fp32 tiles into an L2-resident global accumulator?
// The attention backward runs TWO tensor-core kernels because dQ needs a reduction over the kv tiles; folding dQ into the
// dK/dV kernel would remove 2 of 7 MMAs and one softmax recompute, but every (kv tile, q tile) pair would then add a
// [64 q x 128 d] fp32 partial (32 KB) into dQ: 2.0 GB of reductions per backward call at FLUX shapes (24 heads, L = 4608).
// That is only viable if the L2 reduction path sustains >~ 4 TB/s.  Three ways are timed:
//   mode 0  cp.reduce.async.bulk.global.shared::cta.add.f32  (TMA bulk reduce of a contiguous smem tile)
//   mode 1  red.global.add.f32, one float per lane (coalesced 128 B per warp instruction)
//   mode 2  red.global.add.v4.f32, four floats per lane
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/bulk_reduce tools/microbench/bulk_reduce.cu
// Run:    gpurun_out/bulk_reduce            (prints GB/s per mode and tile size; nothing else in the repo depends on it)
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x)                                                                                  \
  do {                                                                                         \
    cudaError_t e_ = (x);                                                                      \
    if (e_ != cudaSuccess) {                                                                   \
      std::fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));          \
      std::exit(1);                                                                            \
    }                                                                                          \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// acc: [n_tiles][tile_floats] fp32 (the dQ accumulator of all heads: L2-resident at 57 MB)
template <int MODE>
__global__ void __launch_bounds__(256) reduce_kernel(float* __restrict__ acc, int n_tiles, int tile_floats, int iters) {
  extern __shared__ __align__(128) float tile[];
  for (int i = threadIdx.x; i < 2 * tile_floats; i += blockDim.x) tile[i] = 1.0f;
  __syncthreads();
  // the access pattern of the fused kernel: CTA (kv tile j of head h) walks over the q tiles of ITS head; the 36 kv-CTAs
  // of a head hit the same q tile at roughly the same time (worst case for the L2 atomic units)
  const int tiles_per_head = 72;
  const int head = (blockIdx.x / 36) % (n_tiles / tiles_per_head);
  for (int it = 0; it < iters; ++it) {
    const int t = head * tiles_per_head + (it + blockIdx.x) % tiles_per_head;
    float* dst = acc + static_cast<size_t>(t) * tile_floats;
    if (MODE == 0) {
      if (threadIdx.x == 0) {  // two tiles in flight per CTA (double-buffered staging, as the real kernel would do)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst),
                     "r"(smem_u32(tile + (it & 1) * tile_floats)), "r"(tile_floats * 4)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");  // the OTHER staging buffer may be rewritten
      }
      __syncthreads();
    } else if (MODE == 1) {
      for (int i = threadIdx.x; i < tile_floats; i += blockDim.x)
        asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dst + i), "f"(tile[i]) : "memory");
    } else {
      for (int i = threadIdx.x * 4; i < tile_floats; i += blockDim.x * 4)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + i), "f"(tile[i]), "f"(tile[i + 1]),
                     "f"(tile[i + 2]), "f"(tile[i + 3])
                     : "memory");
    }
  }
  if (MODE == 0 && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // drain before exit
}

template <int MODE>
static void run(float* acc, int n_tiles, int tile_floats, int ctas, int iters) {
  const int smem = 2 * tile_floats * 4;
  CK(cudaFuncSetAttribute(reduce_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  reduce_kernel<MODE><<<ctas, 256, smem>>>(acc, n_tiles, tile_floats, 4);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  reduce_kernel<MODE><<<ctas, 256, smem>>>(acc, n_tiles, tile_floats, iters);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  const double bytes = static_cast<double>(ctas) * iters * tile_floats * 4.0;
  std::printf("mode %d  tile %6d B  ctas %4d  iters %4d : %8.3f ms  %8.1f GB/s of reductions\n", MODE, smem / 2, ctas, iters, ms,
              bytes / ms / 1e6);
}

int main() {
  const int heads = 24, tiles_per_head = 72;  // L = 4608 = 72 q tiles of 64 rows
  for (int tile_floats : {64 * 128, 32 * 128}) {
    const int n_tiles = heads * tiles_per_head * (64 * 128 / tile_floats);
    float* acc = nullptr;
    CK(cudaMalloc(&acc, static_cast<size_t>(n_tiles) * tile_floats * 4));
    CK(cudaMemset(acc, 0, static_cast<size_t>(n_tiles) * tile_floats * 4));
    for (int ctas : {148, 296, 864}) {
      run<0>(acc, n_tiles, tile_floats, ctas, 72 * 4);
      run<1>(acc, n_tiles, tile_floats, ctas, 72);
      run<2>(acc, n_tiles, tile_floats, ctas, 72);
    }
    CK(cudaFree(acc));
  }
  return 0;
}
