// fp32 -> fp16 (IQAP) / bf16 (FA) on the HOST, on a pool of worker threads: the host-buffer entry points can halve the bytes that cross
// PCIe (803 KB of fp32 ResNet features per question: the upload, not the GPU, bounds end-to-end throughput) by rounding
// the features to fp16 while the previous chunk is on the wire.  Round-to-nearest-even, as torch's .half(); the device
// path consumes them exactly like a caller-provided fp16 feature store (b200vqa_iqap_forward_host_f16).
//
// The vector paths are compiled per function (target attributes) and chosen at run time: AVX-512F, else AVX2 + F16C,
// else the scalar conversion of cuda_fp16.h.  Non-temporal stores: the destination is pinned staging read by the DMA
// engine next, it must not displace the source from the cache hierarchy.
#include <atomic>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include <unistd.h>

#include <cuda_fp16.h>
#include <immintrin.h>

#include "host_util.h"
#include "kernels.h"

namespace b200vqa {
namespace {

__attribute__((target("avx512f"))) void cvt_avx512(const float* src, uint16_t* dst, size_t n) {
  size_t i = 0;
  for (; i + 32 <= n; i += 32) {
    const __m256i a = _mm512_cvtps_ph(_mm512_loadu_ps(src + i), _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC);
    const __m256i b = _mm512_cvtps_ph(_mm512_loadu_ps(src + i + 16), _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i), a);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 16), b);
  }
  for (; i < n; ++i) {
    const __half h = __float2half_rn(src[i]);
    std::memcpy(dst + i, &h, 2);
  }
}

__attribute__((target("avx2,f16c"))) void cvt_f16c(const float* src, uint16_t* dst, size_t n) {
  size_t i = 0;
  for (; i + 16 <= n; i += 16) {
    const __m128i a = _mm256_cvtps_ph(_mm256_loadu_ps(src + i), _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC);
    const __m128i b = _mm256_cvtps_ph(_mm256_loadu_ps(src + i + 8), _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i), _mm256_set_m128i(b, a));
  }
  for (; i < n; ++i) {
    const __half h = __float2half_rn(src[i]);
    std::memcpy(dst + i, &h, 2);
  }
}

void cvt_scalar(const float* src, uint16_t* dst, size_t n) {
  for (size_t i = 0; i < n; ++i) {
    const __half h = __float2half_rn(src[i]);
    std::memcpy(dst + i, &h, 2);
  }
}

// fp32 -> bf16, round to nearest even (finite inputs: bit-identical to the device's __float2bfloat16_rn)
inline uint16_t bf16_rne_scalar(float x) {
  uint32_t u;
  std::memcpy(&u, &x, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return uint16_t((u >> 16) | 0x40u);  // NaN stays a (quiet) NaN
  return uint16_t((u + 0x7fffu + ((u >> 16) & 1u)) >> 16);
}

__attribute__((target("avx512f,avx512bw"))) void cvt_bf16_avx512(const float* src, uint16_t* dst, size_t n) {
  size_t i = 0;
  const __m512i bias = _mm512_set1_epi32(0x7fff), one = _mm512_set1_epi32(1);
  for (; i + 32 <= n; i += 32) {
    __m512i a = _mm512_loadu_si512(src + i), b = _mm512_loadu_si512(src + i + 16);
    a = _mm512_srli_epi32(_mm512_add_epi32(_mm512_add_epi32(a, bias), _mm512_and_si512(_mm512_srli_epi32(a, 16), one)), 16);
    b = _mm512_srli_epi32(_mm512_add_epi32(_mm512_add_epi32(b, bias), _mm512_and_si512(_mm512_srli_epi32(b, 16), one)), 16);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i), _mm512_cvtepi32_epi16(a));
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 16), _mm512_cvtepi32_epi16(b));
  }
  for (; i < n; ++i) dst[i] = bf16_rne_scalar(src[i]);
}

void cvt_bf16_scalar(const float* src, uint16_t* dst, size_t n) {
  for (size_t i = 0; i < n; ++i) dst[i] = bf16_rne_scalar(src[i]);
}

using CvtFn = void (*)(const float*, uint16_t*, size_t);
CvtFn pick_cvt_bf16() {
  __builtin_cpu_init();
  if (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw")) return cvt_bf16_avx512;
  return cvt_bf16_scalar;
}
CvtFn pick_cvt() {
  __builtin_cpu_init();
  if (__builtin_cpu_supports("avx512f")) return cvt_avx512;
  if (__builtin_cpu_supports("avx2") && __builtin_cpu_supports("f16c")) return cvt_f16c;
  return cvt_scalar;
}

// Minimal fork-join pool: the calling thread takes a share of the blocks too.
class Pool {
 public:
  explicit Pool(int workers) {
    for (int i = 0; i < workers; ++i) threads_.emplace_back([this] { loop(); });
  }
  ~Pool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : threads_) t.join();
  }
  int workers() const { return int(threads_.size()); }
  void run(size_t n_blocks, const std::function<void(size_t)>& fn) {
    std::unique_lock<std::mutex> lk(mu_);
    fn_ = &fn;
    n_blocks_ = n_blocks;
    next_.store(0);
    active_ = int(threads_.size());
    ++epoch_;
    lk.unlock();
    cv_.notify_all();
    work();
    lk.lock();
    done_cv_.wait(lk, [this] { return active_ == 0; });
    fn_ = nullptr;
  }

 private:
  void work() {
    for (;;) {
      const size_t b = next_.fetch_add(1);
      if (b >= n_blocks_) break;
      (*fn_)(b);
    }
  }
  void loop() {
    uint64_t seen = 0;
    std::unique_lock<std::mutex> lk(mu_);
    for (;;) {
      cv_.wait(lk, [&] { return stop_ || epoch_ != seen; });
      if (stop_) return;
      seen = epoch_;
      lk.unlock();
      work();
      lk.lock();
      if (--active_ == 0) done_cv_.notify_one();
    }
  }
  std::vector<std::thread> threads_;
  std::mutex mu_;
  std::condition_variable cv_, done_cv_;
  const std::function<void(size_t)>* fn_ = nullptr;
  size_t n_blocks_ = 0;
  std::atomic<size_t> next_{0};
  int active_ = 0;
  uint64_t epoch_ = 0;
  bool stop_ = false;
};

std::mutex g_pool_mu;   // one conversion at a time per process (the pool is shared by every handle)
Pool* g_pool = nullptr;  // lives as long as the process
int g_pool_threads = 0;
pid_t g_pool_pid = 0;    // a forked child inherits the pointer but not the threads: it builds its own pool

}  // namespace

int host_convert_threads() {
  // the CPUs this process may run on (torchrun ranks share the box), capped: the conversion is memory-bound long before
  cpu_set_t set;
  int n = int(std::thread::hardware_concurrency());
  if (sched_getaffinity(0, sizeof(set), &set) == 0) n = CPU_COUNT(&set);
  if (const char* g = getenv("B200VQA_HOST_THREADS")) n = atoi(g);
  return std::max(1, std::min(n, 32));
}

static void host_convert(CvtFn cvt, const float* src, void* dst, size_t n, int threads);

void host_f32_to_f16(const float* src, void* dst, size_t n, int threads) {
  static const CvtFn cvt = pick_cvt();
  host_convert(cvt, src, dst, n, threads);
}

void host_f32_to_bf16(const float* src, void* dst, size_t n, int threads) {
  static const CvtFn cvt = pick_cvt_bf16();
  host_convert(cvt, src, dst, n, threads);
}

static void host_convert(CvtFn cvt, const float* src, void* dst, size_t n, int threads) {
  uint16_t* out = static_cast<uint16_t*>(dst);
  if (threads <= 0) threads = host_convert_threads();
  constexpr size_t kBlock = size_t(1) << 18;  // elements per task (1 MB of source)
  const size_t n_blocks = (n + kBlock - 1) / kBlock;
  if (threads == 1 || n_blocks <= 1) {
    cvt(src, out, n);
    _mm_sfence();
    return;
  }
  std::lock_guard<std::mutex> lk(g_pool_mu);
  if (!g_pool || g_pool_threads != threads || g_pool_pid != getpid()) {
    if (g_pool && g_pool_pid == getpid()) delete g_pool;  // (never joins threads that only existed in the parent)
    g_pool = new Pool(threads - 1);
    g_pool_threads = threads;
    g_pool_pid = getpid();
  }
  g_pool->run(n_blocks, [&](size_t b) {
    const size_t lo = b * kBlock, len = std::min(kBlock, n - lo);
    cvt(src + lo, out + lo, len);
    _mm_sfence();  // the non-temporal stores are globally visible before the block counts as done
  });
}

}  // namespace b200vqa

extern "C" B200VQA_API int b200vqa_host_f32_to_bf16(const float* src, void* dst, long long n, int threads) {
  B200VQA_REQUIRE(n >= 0, "negative element count");
  if (n == 0) return B200VQA_OK;
  B200VQA_REQUIRE(src && dst, "a required buffer is NULL");
  B200VQA_REQUIRE((reinterpret_cast<uintptr_t>(dst) & 31u) == 0, "dst must be 32-byte aligned");
  b200vqa::host_f32_to_bf16(src, dst, size_t(n), threads);
  return B200VQA_OK;
}

extern "C" B200VQA_API int b200vqa_host_f32_to_f16(const float* src, void* dst, long long n, int threads) {
  B200VQA_REQUIRE(n >= 0, "negative element count");
  if (n == 0) return B200VQA_OK;
  B200VQA_REQUIRE(src && dst, "a required buffer is NULL");
  B200VQA_REQUIRE((reinterpret_cast<uintptr_t>(dst) & 31u) == 0, "dst must be 32-byte aligned");
  b200vqa::host_f32_to_f16(src, dst, size_t(n), threads);
  return B200VQA_OK;
}
