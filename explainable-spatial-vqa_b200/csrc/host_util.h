// Host-side plumbing shared by the C-ABI translation units: thread-local error string, CUDA error
// translation and TMA tensor-map construction (driver entry point resolved at run time so the library
// links against cudart only and still loads on a machine without a GPU driver).
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/b200vqa.h"

namespace b200vqa {

void set_error(const char* fmt, ...);
const char* get_error();

#define B200VQA_CUDA_OK(expr)                                                                  \
  do {                                                                                         \
    cudaError_t e__ = (expr);                                                                  \
    if (e__ != cudaSuccess) {                                                                  \
      ::b200vqa::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__,  \
                           __LINE__);                                                          \
      return e__ == cudaErrorMemoryAllocation ? B200VQA_ERR_OUT_OF_MEMORY : B200VQA_ERR_CUDA;  \
    }                                                                                          \
  } while (0)

#define B200VQA_REQUIRE(cond, ...)              \
  do {                                          \
    if (!(cond)) {                              \
      ::b200vqa::set_error(__VA_ARGS__);        \
      return B200VQA_ERR_BAD_ARGUMENT;          \
    }                                           \
  } while (0)

enum class TmapType { kBF16, kF32 };

// 2D row-major tensor [rows, cols] with leading dimension ld (elements); box = {128 bytes of cols, box_rows};
// 128-byte swizzle; out-of-bounds elements read as zero.
int make_tmap_2d(CUtensorMap* out, const void* base, TmapType type, uint64_t rows, uint64_t cols, uint64_t ld,
                 uint32_t box_rows);
// Verifies the device is compute capability 10.x (B200) - there is no other code path.
int require_sm100(int device, int* num_sms);

}  // namespace b200vqa
