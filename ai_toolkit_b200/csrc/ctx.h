// Host-side context shared by the C-ABI entry points: device properties, launch counter, the
// driver entry point for TMA descriptor encoding (fetched at run time so the library loads on a
// machine without libcuda), and a small cache of encoded tensor maps.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <mutex>
#include <string>
#include <unordered_map>

#include "../../include/b200_lora.h"

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct TmapKey {
  const void* ptr;
  uint64_t d0, d1, ld;
  uint32_t b0, b1;
  uint32_t dtype_swz;
  bool operator==(const TmapKey& o) const {
    return ptr == o.ptr && d0 == o.d0 && d1 == o.d1 && ld == o.ld && b0 == o.b0 && b1 == o.b1 &&
           dtype_swz == o.dtype_swz;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    h ^= k.d0 * 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h ^= k.d1 * 0xC2B2AE3D27D4EB4Full + (h << 6) + (h >> 2);
    h ^= k.ld * 0x165667B19E3779F9ull + (h << 6) + (h >> 2);
    h ^= (static_cast<size_t>(k.b0) << 32 | k.b1) + (h << 6) + (h >> 2);
    h ^= k.dtype_swz + (h << 6) + (h >> 2);
    return h;
  }
};

struct b200_ctx {
  int device = 0;
  int sm_count = 0;
  int cc_major = 0;
  int cc_minor = 0;
  std::atomic<int64_t> launches{0};
  PFN_encodeTiled encode = nullptr;
  std::mutex mu;
  std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> tmaps;
  void* nccl_comm = nullptr;  // ncclComm_t when b200_comm_init has been called
  int world = 1, rank = 0;
};

namespace b200 {
void set_error(const char* fmt, ...);
int check_ctx(b200_ctx* ctx);
// 2-D bf16 row-major [rows, cols] with leading dimension ld (elements); box = {box_cols, box_rows};
// SWIZZLE_128B (box_cols * 2 bytes must be 128).
int make_tmap_bf16_2d(b200_ctx* ctx, CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_cols, uint32_t box_rows);
}  // namespace b200

#define B200_CUDA_CHECK(expr)                                                                    \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess) {                                                                     \
      b200::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return B200_ERR_CUDA;                                                                      \
    }                                                                                            \
  } while (0)

#define B200_REQUIRE(cond, ...)      \
  do {                               \
    if (!(cond)) {                   \
      b200::set_error(__VA_ARGS__);  \
      return B200_ERR_INVALID;       \
    }                                \
  } while (0)
