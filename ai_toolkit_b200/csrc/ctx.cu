// Context, error reporting and TMA descriptor encoding for the C ABI (include/b200_lora.h).
#include "ctx.h"

#include <stdarg.h>
#include <stdio.h>
#include <string.h>

namespace b200 {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_ctx(b200_ctx* ctx) {
  if (ctx == nullptr) {
    set_error("null b200_ctx");
    return B200_ERR_INVALID;
  }
  if (ctx->cc_major != 10) {
    set_error("device %d is sm_%d%d; this library contains sm_100a code only and has no fallback", ctx->device,
              ctx->cc_major, ctx->cc_minor);
    return B200_ERR_ARCH;
  }
  return B200_OK;
}

int make_tmap_bf16_2d(b200_ctx* ctx, CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_cols, uint32_t box_rows) {
  B200_REQUIRE(ctx->encode != nullptr, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  B200_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "TMA operand %p is not 16-byte aligned", ptr);
  B200_REQUIRE(ld % 8 == 0, "TMA operand leading dimension %llu is not a multiple of 8 elements",
               (unsigned long long)ld);
  B200_REQUIRE(box_cols * 2 == 128 && box_rows >= 1 && box_rows <= 256, "bad TMA box %u x %u", box_cols, box_rows);
  TmapKey key{ptr, cols, rows, ld, box_cols, box_rows, 1u};
  {
    std::lock_guard<std::mutex> lk(ctx->mu);
    auto it = ctx->tmaps.find(key);
    if (it != ctx->tmaps.end()) {
      *out = it->second;
      return B200_OK;
    }
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = ctx->encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d) ptr=%p rows=%llu cols=%llu ld=%llu box=%ux%u", (int)r, ptr,
              (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_cols, box_rows);
    return B200_ERR_CUDA;
  }
  {
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (ctx->tmaps.size() > 65536) ctx->tmaps.clear();
    ctx->tmaps.emplace(key, *out);
  }
  return B200_OK;
}

}  // namespace b200

extern "C" {

int b200_version(void) { return 100; }

const char* b200_last_error(void) { return b200::g_err; }

int b200_ctx_create(b200_ctx** out, int device) {
  if (out == nullptr) {
    b200::set_error("b200_ctx_create: null out pointer");
    return B200_ERR_INVALID;
  }
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    b200::set_error("b200_ctx_create: no CUDA device (%s); this library has no CPU path",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    return B200_ERR_CUDA;
  }
  B200_REQUIRE(device >= 0 && device < ndev, "b200_ctx_create: device %d out of range [0,%d)", device, ndev);
  b200_ctx* ctx = new b200_ctx();
  ctx->device = device;
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) {
    delete ctx;
    b200::set_error("cudaGetDeviceProperties failed: %s", cudaGetErrorString(e));
    return B200_ERR_CUDA;
  }
  ctx->sm_count = prop.multiProcessorCount;
  ctx->cc_major = prop.major;
  ctx->cc_minor = prop.minor;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) ctx->encode = reinterpret_cast<PFN_encodeTiled>(fn);
  if (ctx->cc_major != 10) {
    b200::set_error("device %d is sm_%d%d, need sm_100 (B200)", device, prop.major, prop.minor);
    delete ctx;
    return B200_ERR_ARCH;
  }
  *out = ctx;
  return B200_OK;
}

int b200_ctx_destroy(b200_ctx* ctx) {
  delete ctx;
  return B200_OK;
}

int64_t b200_ctx_launch_count(const b200_ctx* ctx) { return ctx ? ctx->launches.load() : 0; }

}  // extern "C"
