#include "host_util.h"
#include "kernels.h"

#include <cstring>
#include <mutex>
#include <set>
#include <utility>

namespace b200vqa {

namespace {
thread_local char g_err[1024] = "";

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn resolve_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}
}  // namespace

namespace {
bool g_pdl = true;
}
bool pdl_enabled() { return g_pdl; }
void set_pdl_enabled(bool on) { g_pdl = on; }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

const char* get_error() { return g_err; }

static int encode_2d(CUtensorMap* out, const void* base, TmapType type, uint64_t rows, uint64_t cols, uint64_t ld,
                     uint32_t box_cols, uint32_t box_rows, CUtensorMapSwizzle swizzle);

int make_tmap_2d(CUtensorMap* out, const void* base, TmapType type, uint64_t rows, uint64_t cols, uint64_t ld,
                 uint32_t box_rows) {
  const uint32_t esz = type == TmapType::kBF16 ? 2 : 4;
  return encode_2d(out, base, type, rows, cols, ld, 128 / esz, box_rows, CU_TENSOR_MAP_SWIZZLE_128B);
}

static int encode_2d(CUtensorMap* out, const void* base, TmapType type, uint64_t rows, uint64_t cols, uint64_t ld,
                     uint32_t box_cols, uint32_t box_rows, CUtensorMapSwizzle swizzle) {
  EncodeTiledFn enc = resolve_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return B200VQA_ERR_CUDA;
  }
  const uint32_t esz = type == TmapType::kBF16 ? 2 : 4;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || ((ld * esz) & 15)) {
    set_error("TMA operand must be 16-byte aligned (base %p, leading dimension %llu elements)", base,
              (unsigned long long)ld);
    return B200VQA_ERR_BAD_ARGUMENT;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * esz};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, type == TmapType::kBF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                   2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows %llu cols %llu ld %llu box_rows %u)", int(r),
              (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows);
    return B200VQA_ERR_CUDA;
  }
  return B200VQA_OK;
}

cudaError_t ensure_dyn_smem(const void* kernel, int bytes) {
  // the attribute belongs to the (kernel, device) pair: a second handle on another GPU of the same process needs its
  // own opt-in (pipeline slots may also call from different host threads)
  static std::mutex mu;
  static std::set<std::pair<const void*, int>> done;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lock(mu);
  if (done.count({kernel, dev})) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) done.insert({kernel, dev});
  return e;
}

cudaError_t current_device_sms(int* sms) {
  static int cache[64] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev >= 0 && dev < 64 && cache[dev] > 0) {
    *sms = cache[dev];
    return cudaSuccess;
  }
  int n = 0;
  e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return e;
  if (dev >= 0 && dev < 64) cache[dev] = n;
  *sms = n;
  return cudaSuccess;
}

int require_sm100(int device, int* num_sms) {
  int major = 0, minor = 0, sms = 0;
  cudaError_t e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  if (e != cudaSuccess) {
    set_error("cudaDeviceGetAttribute(%d) failed: %s", device, cudaGetErrorString(e));
    return B200VQA_ERR_CUDA;
  }
  if (major != 10) {
    set_error("device %d is sm_%d%d; libb200vqa is built for sm_100a (B200) only and has no other path", device, major,
              minor);
    return B200VQA_ERR_UNSUPPORTED_ARCH;
  }
  if (num_sms) *num_sms = sms;
  return B200VQA_OK;
}

}  // namespace b200vqa
