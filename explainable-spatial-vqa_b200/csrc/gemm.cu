// Persistent, warp-specialised tcgen05 GEMM for the executor's dense layers.
//
//   out[M,N] = epilogue( A[M,K] . W[N,K]^T + bias )          A, W row-major ("K-major"), bf16 or tf32
//
// This one kernel family carries every Linear on the hot path:
//   image_proj                      (IQAP:152, FA:48/131)      tf32 inputs, PE + row-remap epilogue
//   MHA in_proj / cross K,V / q     (torch functional.py:6244+) bias epilogue
//   MHA out_proj + residual + norm  (transformer.py _sa_block/_mha_block + norm1/2)  LN epilogue
//   linear1 + ReLU                  (transformer.py _ff_block) ReLU epilogue
//   linear2 + residual + norm       LN epilogue
//
// Structure (one CTA per SM, 192 threads):
//   warp 0      : TMA producer   - A tile [128 x 128B] and W tile [BN x 128B] per k-block into a smem ring
//   warp 1      : MMA issuer     - tcgen05.mma (M=128, N=BN, K=32B) accumulating into TMEM; owns TMEM alloc
//   warps 2..5  : epilogue       - tcgen05.ld the 128 x BN fp32 accumulator (one row per thread), fused
//                                  bias / ReLU / residual+LayerNorm / positional-encoding, bf16 stores
// The accumulator is double-buffered in TMEM (2 x BN columns) so the epilogue of tile i overlaps the
// mainloop of tile i+1. Tiles are walked n-fastest so CTAs running concurrently share A rows in L2.
#include "kernels.h"
#include "ptx.cuh"

namespace b200vqa {

namespace {

constexpr int kBM = 128;          // rows per tile = UMMA M
constexpr int kKBytes = 128;      // bytes of K per k-block = one 128B swizzle row
constexpr int kUmmaKBytes = 32;   // bytes of K per tcgen05.mma (16 bf16 / 8 tf32)
constexpr int kGemmThreads = 192;
constexpr int kEpiWarp0 = 2;
constexpr int kAccStages = 2;

template <int BN>
struct GemmSmem {
  static constexpr int kStageA = kBM * kKBytes;  // 16 KB
  static constexpr int kStageB = BN * kKBytes;   // 16 / 32 KB
  static constexpr int kStage = kStageA + kStageB;
  static constexpr int kStages = (BN == 256) ? 4 : 6;
  static constexpr int kBarOff = kStages * kStage;
  static constexpr int kBytes = kBarOff + 256 /*barriers + tmem ptr*/ + 1024 /*alignment slack*/;
};

template <int BN, int EPI, bool TF32>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w,
               const GemmParams p) {
  using L = GemmSmem<BN>;
  constexpr int kStages = L::kStages;
  constexpr uint32_t kTmemCols = kAccStages * BN;
  static_assert(kTmemCols <= 512, "TMEM budget");
  static_assert(EPI != kEpiBiasResLN || BN == 256, "LN epilogue needs the whole row in one tile");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* acc_full = empty_bar + kStages;
  uint64_t* acc_empty = acc_full + kAccStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + kAccStages);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int elem_bytes = TF32 ? 4 : 2;
  const int bk = kKBytes / elem_bytes;               // elements of K per k-block
  const int num_kb = (p.K + bk - 1) / bk;
  const int tiles_m = (p.M + kBM - 1) / kBM;
  const int tiles_n = p.N / BN;
  const int num_tiles = tiles_m * tiles_n;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_w);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < kAccStages; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 4);  // one arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / tiles_n) * kBM;
        const int n0 = (tile % tiles_n) * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::kStage;
          uint8_t* sb = sa + L::kStageA;
          mbar_expect_tx(&full_bar[stage], L::kStage);
          tma_load_2d(&tm_a, &full_bar[stage], sa, kb * bk, m0);
          tma_load_2d(&tm_w, &full_bar[stage], sb, kb * bk, n0);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(TF32 ? kFmtTF32 : kFmtBF16, kBM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&acc_empty[as], aphase ^ 1);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after_sync();
          const uint32_t sa = smem_u32(smem + stage * L::kStage);
          const uint32_t sb = sa + L::kStageA;
#pragma unroll
          for (int k = 0; k < kKBytes / kUmmaKBytes; ++k) {
            const uint64_t da = make_smem_desc_sw128(sa + k * kUmmaKBytes, 16, 1024);
            const uint64_t db = make_smem_desc_sw128(sb + k * kUmmaKBytes, 16, 1024);
            if (TF32) umma_tf32(d_tmem, da, db, idesc, (kb | k) != 0);
            else      umma_bf16(d_tmem, da, db, idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&acc_full[as]);        // accumulator ready for the epilogue
        if (++as == kAccStages) { as = 0; aphase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (4 warps, row per thread)
    const int quarter = warp & 3;  // TMEM lanes [32*quarter, 32*quarter+32) are the ones this warp may read
    const int row_in_tile = quarter * 32 + lane;
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / tiles_n) * kBM;
      const int n0 = (tile % tiles_n) * BN;
      const int row = m0 + row_in_tile;
      const bool valid = row < p.M;
      mbar_wait(&acc_full[as], aphase);
      __syncwarp();
      tc_fence_after_sync();
      const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(as * BN);

      if constexpr (EPI == kEpiBias || EPI == kEpiBiasRelu) {
        __nv_bfloat16* orow = p.out + size_t(valid ? row : 0) * p.ldc + n0;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t r[32];
          tmem_ld32(taddr + c * 32, r);
          tmem_ld_wait();
          uint32_t o[16];
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            float v0 = __uint_as_float(r[j]);
            float v1 = __uint_as_float(r[j + 1]);
            if (p.bias) {
              v0 += __ldg(p.bias + n0 + c * 32 + j);
              v1 += __ldg(p.bias + n0 + c * 32 + j + 1);
            }
            if (EPI == kEpiBiasRelu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
            o[j >> 1] = pack_bf16x2(v0, v1);
          }
          if (valid) {
            uint4* dst = reinterpret_cast<uint4*>(orow + c * 32);
#pragma unroll
            for (int q = 0; q < 4; ++q) dst[q] = make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
          }
        }
      } else if constexpr (EPI == kEpiBiasPeRemap) {
        const int item = row / p.rows_in;
        const int pos = row - item * p.rows_in;
        __nv_bfloat16* orow = p.out + (size_t(valid ? item : 0) * p.rows_out + p.row_off + (valid ? pos : 0)) * p.ldc + n0;
        const float* perow = p.pe + size_t(p.pe_off + (valid ? pos : 0)) * p.N + n0;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t r[32];
          tmem_ld32(taddr + c * 32, r);
          tmem_ld_wait();
          uint32_t o[16];
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 pe4 = __ldg(reinterpret_cast<const float4*>(perow + c * 32 + j));
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c * 32 + j));
            const float v0 = __uint_as_float(r[j]) + b4.x + pe4.x;
            const float v1 = __uint_as_float(r[j + 1]) + b4.y + pe4.y;
            const float v2 = __uint_as_float(r[j + 2]) + b4.z + pe4.z;
            const float v3 = __uint_as_float(r[j + 3]) + b4.w + pe4.w;
            o[j >> 1] = pack_bf16x2(v0, v1);
            o[(j >> 1) + 1] = pack_bf16x2(v2, v3);
          }
          if (valid) {
            uint4* dst = reinterpret_cast<uint4*>(orow + c * 32);
#pragma unroll
            for (int q = 0; q < 4; ++q) dst[q] = make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
          }
        }
      } else {  // kEpiBiasResLN : whole 256-wide row belongs to this thread
        const __nv_bfloat16* rrow = p.residual + size_t(valid ? row : 0) * p.ldr;
        float sum = 0.f, sumsq = 0.f;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t r[32];
          tmem_ld32(taddr + c * 32, r);
          tmem_ld_wait();
          const uint4* rs = reinterpret_cast<const uint4*>(rrow + c * 32);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 rv = rs[q];
            const uint32_t w[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int j = q * 8 + e * 2;
              const float2 res = unpack_bf16x2(w[e]);
              const float v0 = __uint_as_float(r[j]) + __ldg(p.bias + c * 32 + j) + res.x;
              const float v1 = __uint_as_float(r[j + 1]) + __ldg(p.bias + c * 32 + j + 1) + res.y;
              sum += v0 + v1;
              sumsq += v0 * v0 + v1 * v1;
            }
          }
        }
        const float mean = sum * (1.f / BN);
        const float var = fmaxf(sumsq * (1.f / BN) - mean * mean, 0.f);
        const float rstd = rsqrtf(var + p.eps);
        __nv_bfloat16* orow = p.out + size_t(valid ? row : 0) * p.ldc;
        float* frow = p.out_f32 ? p.out_f32 + size_t(valid ? row : 0) * BN : nullptr;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t r[32];
          tmem_ld32(taddr + c * 32, r);
          tmem_ld_wait();
          const uint4* rs = reinterpret_cast<const uint4*>(rrow + c * 32);
          uint32_t o[16];
          float yv[32];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 rv = rs[q];
            const uint32_t w[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int j = q * 8 + e * 2;
              const float2 res = unpack_bf16x2(w[e]);
              const float v0 = __uint_as_float(r[j]) + __ldg(p.bias + c * 32 + j) + res.x;
              const float v1 = __uint_as_float(r[j + 1]) + __ldg(p.bias + c * 32 + j + 1) + res.y;
              const float y0 = (v0 - mean) * rstd * __ldg(p.gamma + c * 32 + j) + __ldg(p.beta + c * 32 + j);
              const float y1 =
                  (v1 - mean) * rstd * __ldg(p.gamma + c * 32 + j + 1) + __ldg(p.beta + c * 32 + j + 1);
              yv[j] = y0;
              yv[j + 1] = y1;
              o[j >> 1] = pack_bf16x2(y0, y1);
            }
          }
          if (valid) {
            uint4* dst = reinterpret_cast<uint4*>(orow + c * 32);
#pragma unroll
            for (int q = 0; q < 4; ++q) dst[q] = make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
            if (frow) {
              float4* fdst = reinterpret_cast<float4*>(frow + c * 32);
#pragma unroll
              for (int q = 0; q < 8; ++q) fdst[q] = make_float4(yv[4 * q], yv[4 * q + 1], yv[4 * q + 2], yv[4 * q + 3]);
            }
          }
        }
      }

      // all TMEM reads of this accumulator stage are complete -> hand it back to the MMA warp
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[as]);
      if (++as == kAccStages) { as = 0; aphase ^= 1; }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after_sync();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

template <int BN, int EPI, bool TF32>
cudaError_t launch_one(const CUtensorMap& tm_a, const CUtensorMap& tm_w, const GemmParams& p, int num_sms,
                       cudaStream_t stream) {
  using L = GemmSmem<BN>;
  auto kfn = gemm_tc_kernel<BN, EPI, TF32>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kBytes);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  const int tiles = ((p.M + kBM - 1) / kBM) * (p.N / BN);
  if (tiles <= 0) return cudaSuccess;
  const int grid = tiles < num_sms ? tiles : num_sms;
  kfn<<<grid, kGemmThreads, L::kBytes, stream>>>(tm_a, tm_w, p);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// Test-only CUDA-core check GEMM (fp32 accumulate, one output element per thread).
// ------------------------------------------------------------------------------------------
template <typename TA>
__global__ void gemm_check_kernel(const TA* __restrict__ A, const TA* __restrict__ W, const float* __restrict__ bias,
                                  float* __restrict__ out, int M, int N, int K) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = blockIdx.y;
  if (n >= N || m >= M) return;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) acc += float(A[size_t(m) * K + k]) * float(W[size_t(n) * K + k]);
  out[size_t(m) * N + n] = acc + (bias ? bias[n] : 0.f);
}

}  // namespace

cudaError_t launch_gemm(int epilogue, bool tf32, int block_n, const CUtensorMap& tm_a, const CUtensorMap& tm_w,
                        const GemmParams& p, int num_sms, cudaStream_t stream) {
  if (p.N % block_n != 0) return cudaErrorInvalidValue;
  if (epilogue == kEpiBiasResLN && (p.N != 256 || block_n != 256)) return cudaErrorInvalidValue;
#define B200VQA_GEMM_CASE(BN_, EPI_, TF_)                                                   \
  if (block_n == BN_ && epilogue == EPI_ && tf32 == TF_)                                    \
    return launch_one<BN_, EPI_, TF_>(tm_a, tm_w, p, num_sms, stream);
  B200VQA_GEMM_CASE(256, kEpiBias, false)
  B200VQA_GEMM_CASE(128, kEpiBias, false)
  B200VQA_GEMM_CASE(256, kEpiBiasRelu, false)
  B200VQA_GEMM_CASE(128, kEpiBiasRelu, false)
  B200VQA_GEMM_CASE(256, kEpiBiasResLN, false)
  B200VQA_GEMM_CASE(256, kEpiBiasPeRemap, false)
  B200VQA_GEMM_CASE(256, kEpiBiasPeRemap, true)
  B200VQA_GEMM_CASE(256, kEpiBias, true)
#undef B200VQA_GEMM_CASE
  return cudaErrorInvalidValue;
}

cudaError_t launch_gemm_check(bool a_is_f32, const void* A, const void* W, const float* bias, float* out, int M,
                              int N, int K, cudaStream_t stream) {
  dim3 block(128), grid((N + 127) / 128, M);
  if (a_is_f32)
    gemm_check_kernel<float><<<grid, block, 0, stream>>>(static_cast<const float*>(A), static_cast<const float*>(W),
                                                         bias, out, M, N, K);
  else
    gemm_check_kernel<__nv_bfloat16><<<grid, block, 0, stream>>>(static_cast<const __nv_bfloat16*>(A),
                                                                 static_cast<const __nv_bfloat16*>(W), bias, out, M,
                                                                 N, K);
  return cudaGetLastError();
}

}  // namespace b200vqa
