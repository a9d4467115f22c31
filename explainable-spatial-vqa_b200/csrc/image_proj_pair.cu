// image_proj (IQAP:152 / FA:48,131): out[item, row_off + pos, :] = feat[item, pos, :] . W^T + bias + PE[pe_off + pos], N = 256,
// straight from the caller's features (fp32 read as tf32, or an fp16 / bf16 feature store), written into the encoder
// input rows - the reference's torch.cat.
//
// The GEMM is HBM-bound on paper (803 KB of fp32 features per question for 103 MFLOP), but one CTA per 128-row tile also
// pulls the whole weight (1 MB fp32) through its SM for every tile: two thirds of what crosses each SM's shared memory
// is W coming back from L2, and the tile rate is set by that fill, not by HBM (0.24 ms per 1024 questions where the
// features alone take 0.13 ms).  Here a CTA PAIR (cta_group::2) owns 256 rows and each CTA loads HALF of every W
// k-block (the B operand is split by N across the pair).
//
// Warp roles (320 threads): warp 0 TMA producer (ring of six 32-KB stages: 16 KB of A + 16 KB of W per CTA; bytes are
// counted on the leader's mbarriers), warp 1 MMA issuer (leader only; whole warp walks the loop, one elected lane
// issues from uniform registers), warps 2-9 epilogue (two per TMEM lane quarter; two accumulator stages, so the
// epilogue of a tile runs under the MMAs of the next).
#include <cstdio>

#include "host_util.h"
#include "kernels.h"
#include "ptx.cuh"

namespace b200vqa {
namespace {

constexpr int kIpThreads = 320;
constexpr int kIpStages = 6;
constexpr int kIpStage = 32768;  // per CTA: A k-block [128 rows x 128 B] | W k-block half [128 of the 256 outputs x 128 B]

struct IpSmem {
  static constexpr int kOffRing = 0;
  static constexpr int kOffBar = kIpStages * kIpStage;
  static constexpr int kOffBias = kOffBar + 256;
  static constexpr int kBytes = kOffBias + 1024;
};

// Bounded by the clock (about one second): a protocol bug traps instead of hanging the GPU.
__device__ __forceinline__ void ip_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 2000000000ll) {
      printf("b200vqa: image_proj_pair wait timed out (block %d thread %d barrier +%d parity %u)\n", blockIdx.x, threadIdx.x,
             int(smem_u32(bar) & 255u), parity);
      __trap();
    }
  }
}

// FMT: kFmtTF32 (fp32 elements, 32 per k-block), kFmtF16 / kFmtBF16 (64 per k-block)
template <uint32_t FMT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kIpThreads, 1)
image_proj_pair_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w,
                       const GemmParams p, const uint64_t a_policy) {
  using L = IpSmem;
  constexpr int kElems = FMT == kFmtTF32 ? 32 : 64;  // elements of K per 128-byte k-block row
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* ring = smem + L::kOffRing;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kOffBar);
  uint64_t* ring_full = bars;                 // [6] leader: both CTAs' halves of the stage have landed
  uint64_t* acc_empty = bars + kIpStages;     // [2] leader: accumulator read out by the epilogues of both CTAs (16 warps)
  uint64_t* ring_empty = bars + kIpStages + 2;      // [6] both (multicast commit): the MMAs that read the stage retired
  uint64_t* acc_full = bars + 2 * kIpStages + 2;    // [2] both (multicast commit)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kIpStages + 4);
  float* s_bias = reinterpret_cast<float*>(smem + L::kOffBias);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int n_sup = (p.M + 255) / 256;
  const int num_kb = p.K / kElems;
  const int sup0 = blockIdx.x >> 1, sup_step = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_w);
    for (int i = 0; i < kIpStages; ++i) {
      mbar_init(&ring_full[i], 1);
      mbar_init(&ring_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 16);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_pair<512>(tmem_slot);
  if (threadIdx.x >= 64) s_bias[threadIdx.x - 64] = p.bias ? __ldg(p.bias + threadIdx.x - 64) : 0.f;  // 256 epilogue threads
  tc_fence_before_sync();
  __syncthreads();
  cluster_arrive_release();  // both CTAs' barriers exist before anything is signalled across the pair
  cluster_wait_acquire();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    const bool el = elect_one();
    uint32_t u = 0;
    for (int sup = sup0; sup < n_sup; sup += sup_step) {
      for (int kb = 0; kb < num_kb; ++kb, ++u) {
        const uint32_t st = u % kIpStages;
        ip_wait(&ring_empty[st], ((u / kIpStages) & 1) ^ 1);
        if (el) {
          if (leader) mbar_expect_tx(&ring_full[st], 2 * kIpStage);
          // the features are read exactly once: evict-first keeps 803 KB per question from displacing what the
          // concurrent decode chains re-use in L2
          tma_load_2d_pair_hint(&tm_a, &ring_full[st], ring + st * kIpStage, kb * kElems, sup * 256 + int(rank) * 128, a_policy);
          tma_load_2d_pair(&tm_w, &ring_full[st], ring + st * kIpStage + 16384, kb * kElems, int(rank) * 128);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    if (leader) {
      const bool el = elect_one();
      constexpr uint32_t idesc = make_idesc(FMT, 256, 256, 0, 0);
      const uint64_t d_base = make_smem_desc_sw128(smem_u32(ring), 16, 1024);
      uint32_t u = 0, t = 0;
      for (int sup = sup0; sup < n_sup; sup += sup_step, ++t) {
        const uint32_t as = t & 1;
        ip_wait(&acc_empty[as], ((t >> 1) & 1) ^ 1);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + as * 256;
        for (int kb = 0; kb < num_kb; ++kb, ++u) {
          const uint32_t st = u % kIpStages;
          ip_wait(&ring_full[st], (u / kIpStages) & 1);
          tc_fence_after_sync();
          const uint64_t da = d_base + uint64_t((st * kIpStage) >> 4);
          const uint64_t db = da + uint64_t(16384 >> 4);
          if (el) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {  // 32 bytes of K per MMA (8 tf32 / 16 half-precision elements)
              if constexpr (FMT == kFmtTF32)
                umma_tf32_pair(d_tmem, da + uint64_t((k * 32) >> 4), db + uint64_t((k * 32) >> 4), idesc, (kb | k) != 0);
              else
                umma_bf16_pair(d_tmem, da + uint64_t((k * 32) >> 4), db + uint64_t((k * 32) >> 4), idesc, (kb | k) != 0);
            }
            umma_commit_pair(&ring_empty[st]);
          }
          __syncwarp();
        }
        if (el) umma_commit_pair(&acc_full[as]);
        __syncwarp();
      }
    }
  } else {
    // ---- epilogue: row of this CTA's tile == TMEM lane; the two warps of a lane quarter take 128 columns each
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const uint32_t lane_off = uint32_t(quarter * 32) << 16;
    uint32_t t = 0;
    for (int sup = sup0; sup < n_sup; sup += sup_step, ++t) {
      const uint32_t as = t & 1;
      const int row = sup * 256 + int(rank) * 128 + quarter * 32 + lane;
      const bool valid = row < p.M;
      const int item = (valid ? row : 0) / p.rows_in;
      const int pos = (valid ? row : 0) - item * p.rows_in;
      __nv_bfloat16* orow = p.out + (size_t(item) * p.rows_out + p.row_off + pos) * p.ldc + half * 128;
      const float* perow = p.pe + size_t(p.pe_off + pos) * kD + half * 128;
      ip_wait(&acc_full[as], (t >> 1) & 1);
      __syncwarp();
      tc_fence_after_sync();
      const uint32_t taddr = tmem_base + as * 256 + half * 128 + lane_off;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld32(taddr + c * 32, r);
        float4 pe4[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) pe4[j] = __ldg(reinterpret_cast<const float4*>(perow + c * 32 + j * 4));
        tmem_ld_wait();
        if (c == 3) {  // the accumulator stage is in registers: the MMAs of the tile after next may overwrite it
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(cluster_map_shared(smem_u32(&acc_empty[as]), 0));
        }
        uint32_t o[16];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b4 = *reinterpret_cast<const float4*>(s_bias + half * 128 + c * 32 + j * 4);
          o[2 * j] = pack_bf16x2(__uint_as_float(r[4 * j]) + b4.x + pe4[j].x, __uint_as_float(r[4 * j + 1]) + b4.y + pe4[j].y);
          o[2 * j + 1] =
              pack_bf16x2(__uint_as_float(r[4 * j + 2]) + b4.z + pe4[j].z, __uint_as_float(r[4 * j + 3]) + b4.w + pe4[j].w);
        }
        if (valid) {
          uint4* dst = reinterpret_cast<uint4*>(orow + c * 32);
#pragma unroll
          for (int q = 0; q < 4; ++q) dst[q] = make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
        }
      }
    }
  }

  // neither CTA may leave (or free its TMEM) while the other can still read its shared memory or signal its barriers
  tc_fence_before_sync();
  __syncthreads();
  cluster_arrive_release();
  cluster_wait_acquire();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after_sync();
    tmem_dealloc_pair<512>(tmem_base);
  }
}

}  // namespace

// fmt: 0 = fp32 features (tf32 MMA), 1 = fp16, 2 = bf16.  tm_a: [M, K] box 128 B x 128 rows; tm_w: [256, K] box 128 B x 128 rows.
cudaError_t launch_image_proj_pair(int fmt, const CUtensorMap& tm_a, const CUtensorMap& tm_w, const GemmParams& p,
                                   cudaStream_t stream, bool a_evict_first) {
  if (p.M <= 0) return cudaSuccess;
  const int elems = fmt == 0 ? 32 : 64;
  if (p.N != kD || p.K <= 0 || p.K % elems != 0 || !p.out || !p.pe || p.rows_in <= 0) return cudaErrorInvalidValue;
  int num_sms = 0;
  cudaError_t e = current_device_sms(&num_sms);
  if (e != cudaSuccess) return e;
  const int sups = (p.M + 255) / 256;
  const int pairs = sups < num_sms / 2 ? sups : num_sms / 2;
  const dim3 grid(2 * pairs), block(kIpThreads);
  const uint64_t a_policy = a_evict_first ? kL2EvictFirst : kL2EvictNormal;
#define B200VQA_IP_CASE(F)                                                                              \
  {                                                                                                     \
    auto kfn = image_proj_pair_kernel<F>;                                                               \
    e = ensure_dyn_smem(reinterpret_cast<const void*>(kfn), IpSmem::kBytes);                            \
    if (e != cudaSuccess) return e;                                                                     \
    return launch_kernel(kfn, grid, block, IpSmem::kBytes, stream, false, tm_a, tm_w, p, a_policy);     \
  }
  if (fmt == 0) B200VQA_IP_CASE(kFmtTF32)
  if (fmt == 1) B200VQA_IP_CASE(kFmtF16)
  if (fmt == 2) B200VQA_IP_CASE(kFmtBF16)
#undef B200VQA_IP_CASE
  return cudaErrorInvalidValue;
}

}  // namespace b200vqa
