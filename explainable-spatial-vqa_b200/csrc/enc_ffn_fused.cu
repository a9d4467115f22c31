// Encoder feed-forward block in ONE kernel (torch TransformerEncoderLayer._ff_block + norm2, IQAP:173 / FA:42):
//     y = LayerNorm(x + W2 relu(W1 x + b1) + b2)        [+ nn.Transformer's final encoder norm on top, FA:42]
// for M = questions x 256 rows.  As a GEMM pair the hidden activations [M, ff] go to HBM and come back (IQAP, 1024
// questions: 1.07 GB written + 1.07 GB read for 0.13 GB of input).  Here a persistent CTA owns a 128-row tile and walks
// the hidden dimension in slices of 128 units, entirely on chip:
//     H_j = relu(X . W1[j]^T + b1[j])     tcgen05, fp32 accumulator in TMEM (two buffers) -> bf16 in shared memory (two)
//     Y  += H_j . W2[:, j]^T              tcgen05, accumulator [128 x 256] fp32 resident in TMEM for the whole tile
// and the epilogue adds b2 + residual and normalises the row.  HBM sees x once and y once.
//
// Warp roles (448 threads): warp 0 TMA producer (x tile; weights through a ring of three 32-KB units, from L2 after the
// first tile), warp 1 MMA issuer, warps 2-5 hidden epilogue (TMEM -> bias, ReLU -> swizzled K-major bf16 tile),
// warps 6-13 LayerNorm epilogue (two warps per TMEM lane quarter, 128 columns each).  GEMM1 of slice j+1 is issued
// before GEMM2 of slice j, so the tensor pipe never waits for the hidden epilogue; the LayerNorm epilogue of tile i
// runs under the first slices of tile i+1.
//
// Bound: every 128-row tile pulls all of W1 and W2 through its SM (2 * 256 * ff bytes: 2 MB at ff = 2048) - the
// shared-memory fill rate, not the tensor pipe, sets the pace.
#include "host_util.h"
#include "kernels.h"
#include "ptx.cuh"

namespace b200vqa {
namespace {

constexpr int kEfThreads = 448;
constexpr int kEfRing = 3;
constexpr int kEfUnit = 32768;

struct EfSmem {
  static constexpr int kOffX = 0;                       // x tile: 4 k-blocks of [128 rows x 64]
  static constexpr int kOffH = 65536;                   // 2 x H slice: 2 panels of [128 rows x 64 hidden]
  static constexpr int kOffRing = kOffH + 2 * 32768;    // weight units
  static constexpr int kOffBar = kOffRing + kEfRing * kEfUnit;
  static constexpr int kOffStats = kOffBar + 256;       // float2 [2 column halves][128 rows]
  static constexpr int kBytes = kOffStats + 2048;
};
static_assert(EfSmem::kBytes <= 232448, "shared memory budget");

__device__ __forceinline__ float4 ldg4f(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

template <bool LN2>
__global__ void __launch_bounds__(kEfThreads, 1)
enc_ffn_fused_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w1,
                     const __grid_constant__ CUtensorMap tm_w2, const EncFfnParams p) {
  using L = EfSmem;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sX = smem + L::kOffX;
  uint8_t* sH = smem + L::kOffH;
  uint8_t* sRing = smem + L::kOffRing;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kOffBar);
  uint64_t* ring_full = bars;        // [3] weight unit landed
  uint64_t* ring_empty = bars + 3;   // [3] the MMAs that read it have retired
  uint64_t* x_full = bars + 6;
  uint64_t* x_free = bars + 7;       // last GEMM1 of the tile retired
  uint64_t* ht_full = bars + 8;      // [2] H accumulator complete
  uint64_t* ht_free = bars + 10;     // [2] ... and read out by the hidden epilogue
  uint64_t* hs_full = bars + 12;     // [2] H (bf16) written to shared memory
  uint64_t* hs_free = bars + 14;     // [2] ... and consumed by GEMM2
  uint64_t* y_full = bars + 16;
  uint64_t* y_free = bars + 17;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);
  float2* s_stats = reinterpret_cast<float2*>(smem + L::kOffStats);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (p.M + 127) / 128;
  const int n_sl = p.n_slices;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_w1);
    tma_prefetch_desc(&tm_w2);
    for (int i = 0; i < kEfRing; ++i) {
      mbar_init(&ring_full[i], 1);
      mbar_init(&ring_empty[i], 1);
    }
    mbar_init(x_full, 1);
    mbar_init(x_free, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ht_full[i], 1);
      mbar_init(&ht_free[i], 128);
      mbar_init(&hs_full[i], 128);
      mbar_init(&hs_free[i], 1);
    }
    mbar_init(y_full, 1);
    mbar_init(y_free, 256);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t dY = tmem_base;  // 256 columns; H accumulators: 128 columns each at +256 and +384

  if (warp == 0) {
    if (lane == 0) {
      uint32_t u = 0, t = 0;
      auto slot_wait = [&]() -> uint32_t {
        const uint32_t slot = u % kEfRing;
        mbar_wait(&ring_empty[slot], ((u / kEfRing) & 1) ^ 1);
        mbar_expect_tx(&ring_full[slot], kEfUnit);
        ++u;
        return slot;
      };
      auto put_w1 = [&](int j, int half) {  // [128 hidden x 128 of K]: two k-blocks
        const uint32_t slot = slot_wait();
        tma_load_2d(&tm_w1, &ring_full[slot], sRing + slot * kEfUnit, (2 * half) * 64, j * 128);
        tma_load_2d(&tm_w1, &ring_full[slot], sRing + slot * kEfUnit + 16384, (2 * half + 1) * 64, j * 128);
      };
      auto put_w2 = [&](int j, int half) {  // [256 out x 64 hidden]
        const uint32_t slot = slot_wait();
        tma_load_2d(&tm_w2, &ring_full[slot], sRing + slot * kEfUnit, j * 128 + half * 64, 0);
      };
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
        mbar_wait(x_free, (t & 1) ^ 1);
        mbar_expect_tx(x_full, 65536);
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) tma_load_2d(&tm_x, x_full, sX + kb * 16384, kb * 64, tile * 128);
        // the order the MMA warp consumes them in
        put_w1(0, 0);
        put_w1(0, 1);
        for (int j = 0; j < n_sl; ++j) {
          if (j + 1 < n_sl) {
            put_w1(j + 1, 0);
            put_w1(j + 1, 1);
          }
          put_w2(j, 0);
          put_w2(j, 1);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc1 = make_idesc(kFmtBF16, 128, 128, 0, 0);
      constexpr uint32_t idesc2 = make_idesc(kFmtBF16, 128, 256, 0, 0);
      uint32_t u = 0, s = 0, t = 0;
      auto unit_wait = [&]() -> uint32_t {
        const uint32_t slot = u % kEfRing;
        mbar_wait(&ring_full[slot], (u / kEfRing) & 1);
        tc_fence_after_sync();
        ++u;
        return slot;
      };
      auto gemm1 = [&](uint32_t sidx) {  // H[sidx & 1] = X . W1[j]^T : M=128, N=128, K=256
        const uint32_t buf = sidx & 1;
        mbar_wait(&ht_free[buf], ((sidx >> 1) & 1) ^ 1);
        tc_fence_after_sync();
        const uint32_t dH = tmem_base + 256 + buf * 128;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const uint32_t slot = unit_wait();
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int kk = half * 8 + k;
            const uint32_t a = smem_u32(sX) + (kk / 4) * 16384 + (kk % 4) * 32;
            const uint32_t b = smem_u32(sRing) + slot * kEfUnit + (k / 4) * 16384 + (k % 4) * 32;
            umma_bf16(dH, make_smem_desc_sw128(a, 16, 1024), make_smem_desc_sw128(b, 16, 1024), idesc1, kk != 0);
          }
          umma_commit(&ring_empty[slot]);
        }
        umma_commit(&ht_full[buf]);
      };
      auto gemm2 = [&](int j, uint32_t sidx) {  // Y += H[sidx & 1] . W2[:, j]^T : M=128, N=256, K=128
        const uint32_t buf = sidx & 1;
        mbar_wait(&hs_full[buf], (sidx >> 1) & 1);
        if (j == 0) mbar_wait(y_free, (t & 1) ^ 1);
        tc_fence_after_sync();
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const uint32_t slot = unit_wait();
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t a = smem_u32(sH) + buf * 32768 + half * 16384 + k * 32;
            const uint32_t b = smem_u32(sRing) + slot * kEfUnit + k * 32;
            umma_bf16(dY, make_smem_desc_sw128(a, 16, 1024), make_smem_desc_sw128(b, 16, 1024), idesc2,
                      !(j == 0 && half == 0 && k == 0));
          }
          umma_commit(&ring_empty[slot]);
        }
        umma_commit(&hs_free[buf]);
      };
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
        mbar_wait(x_full, t & 1);
        tc_fence_after_sync();
        gemm1(s);
        if (n_sl == 1) umma_commit(x_free);
        for (int j = 0; j < n_sl; ++j) {
          if (j + 1 < n_sl) {
            gemm1(s + j + 1);
            if (j + 2 == n_sl) umma_commit(x_free);  // the last GEMM1 of the tile: x may be replaced when it retires
          }
          gemm2(j, s + j);
        }
        umma_commit(y_full);
        s += n_sl;
      }
    }
  } else if (warp < 6) {
    // ---- hidden epilogue: row r of the tile == TMEM lane r
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_off = uint32_t(quarter * 32) << 16;
    uint32_t s = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      for (int j = 0; j < n_sl; ++j, ++s) {
        const uint32_t buf = s & 1, ph = (s >> 1) & 1;
        const float* b1 = p.b1 + j * 128;
        uint8_t* hbuf = sH + buf * 32768;
        mbar_wait(&hs_free[buf], ph ^ 1);  // GEMM2 of slice s - 2 no longer reads this buffer
        mbar_wait(&ht_full[buf], ph);
        __syncwarp();
        tc_fence_after_sync();
        const uint32_t dH = tmem_base + 256 + buf * 128 + lane_off;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          tmem_ld32(dH + c * 32, v);
          tmem_ld_wait();
          if (c == 3) {  // the accumulator is in registers: GEMM1 of slice s + 2 may overwrite it
            tc_fence_before_sync();
            mbar_arrive(&ht_free[buf]);
          }
          uint32_t o[16];
#pragma unroll
          for (int jj = 0; jj < 32; jj += 4) {
            const float4 b4 = ldg4f(b1 + c * 32 + jj);
            o[jj >> 1] =
                pack_bf16x2(fmaxf(__uint_as_float(v[jj]) + b4.x, 0.f), fmaxf(__uint_as_float(v[jj + 1]) + b4.y, 0.f));
            o[(jj >> 1) + 1] =
                pack_bf16x2(fmaxf(__uint_as_float(v[jj + 2]) + b4.z, 0.f), fmaxf(__uint_as_float(v[jj + 3]) + b4.w, 0.f));
          }
          // K-major, 128-byte swizzled A operand: panel = 64 hidden units, 16-byte chunk index XOR (row & 7)
          uint8_t* prow = hbuf + (c >> 1) * 16384 + r * 128;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int chunk = (c & 1) * 4 + q;
            *reinterpret_cast<uint4*>(prow + ((chunk ^ (r & 7)) << 4)) =
                make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
          }
        }
        fence_proxy_async_smem();
        mbar_arrive(&hs_full[buf]);
      }
    }
  } else {
    // ---- LayerNorm epilogue: warps w and w + 4 share a TMEM lane quarter and split the 256 columns
    const int quarter = warp & 3;
    const int half = (warp - 6) >> 2;
    const int rt = quarter * 32 + lane;  // row inside the tile
    const uint32_t taddr = dY + (uint32_t(quarter * 32) << 16) + half * 128;
    const float* b2 = p.b2 + half * 128;
    uint32_t t = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
      const int row = tile * 128 + rt;
      const bool valid = row < p.M;
      const __nv_bfloat16* res = p.residual + size_t(valid ? row : 0) * kD + half * 128;
      mbar_wait(y_full, t & 1);
      __syncwarp();
      tc_fence_after_sync();
      // v = acc + b2 + residual for 32 columns of this thread's row
      auto load_chunk = [&](int c, float (&v)[32]) {
        uint32_t a[32];
        tmem_ld32(taddr + c * 32, a);
        uint4 rr[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) rr[q] = __ldg(reinterpret_cast<const uint4*>(res + c * 32 + q * 8));
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t w4[4] = {rr[q].x, rr[q].y, rr[q].z, rr[q].w};
          const float4 c0 = ldg4f(b2 + c * 32 + q * 8), c1 = ldg4f(b2 + c * 32 + q * 8 + 4);
          const float bb[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 x2 = unpack_bf16x2(w4[e]);
            v[q * 8 + 2 * e] = __uint_as_float(a[q * 8 + 2 * e]) + bb[2 * e] + x2.x;
            v[q * 8 + 2 * e + 1] = __uint_as_float(a[q * 8 + 2 * e + 1]) + bb[2 * e + 1] + x2.y;
          }
        }
      };
      // row statistics from the two column halves (sum and sum of squares in fp32; the rows are O(1) activations)
      auto combine = [&](float s1, float s2, float& mean, float& rstd) {
        s_stats[half * 128 + rt] = make_float2(s1, s2);
        named_bar_sync(1 + quarter, 64);
        const float2 o = s_stats[(half ^ 1) * 128 + rt];
        mean = (s1 + o.x) * (1.f / kD);
        const float var = fmaxf((s2 + o.y) * (1.f / kD) - mean * mean, 0.f);
        rstd = rsqrtf(var + p.eps);
        named_bar_sync(1 + quarter, 64);  // both have read before the next statistics are written
      };
      float mean, rstd, mean2 = 0.f, rstd2 = 1.f;
      {
        float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          float v[32];
          load_chunk(c, v);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            s1 += v[i];
            s2 = fmaf(v[i], v[i], s2);
          }
        }
        combine(s1, s2, mean, rstd);
      }
      const float* g1 = p.gamma + half * 128;
      const float* t1 = p.beta + half * 128;
      if constexpr (LN2) {
        // second LayerNorm on top (nn.Transformer's final encoder norm): statistics of the normalised row
        float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          float v[32];
          load_chunk(c, v);
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 g4 = ldg4f(g1 + c * 32 + i), t4 = ldg4f(t1 + c * 32 + i);
            const float y0 = (v[i] - mean) * rstd * g4.x + t4.x, y1 = (v[i + 1] - mean) * rstd * g4.y + t4.y;
            const float y2 = (v[i + 2] - mean) * rstd * g4.z + t4.z, y3 = (v[i + 3] - mean) * rstd * g4.w + t4.w;
            s1 += (y0 + y1) + (y2 + y3);
            s2 = fmaf(y0, y0, fmaf(y1, y1, fmaf(y2, y2, fmaf(y3, y3, s2))));
          }
        }
        combine(s1, s2, mean2, rstd2);
      }
      __nv_bfloat16* orow = p.out + size_t(valid ? row : 0) * kD + half * 128;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        float v[32];
        load_chunk(c, v);
        if (c == 3) {  // last read of the accumulator: GEMM2 of the next tile may start
          tc_fence_before_sync();
          mbar_arrive(y_free);
        }
        uint32_t o[16];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 g4 = ldg4f(g1 + c * 32 + i), t4 = ldg4f(t1 + c * 32 + i);
          float y0 = (v[i] - mean) * rstd * g4.x + t4.x, y1 = (v[i + 1] - mean) * rstd * g4.y + t4.y;
          float y2 = (v[i + 2] - mean) * rstd * g4.z + t4.z, y3 = (v[i + 3] - mean) * rstd * g4.w + t4.w;
          if constexpr (LN2) {
            const float4 h4 = ldg4f(p.gamma2 + half * 128 + c * 32 + i), u4 = ldg4f(p.beta2 + half * 128 + c * 32 + i);
            y0 = (y0 - mean2) * rstd2 * h4.x + u4.x;
            y1 = (y1 - mean2) * rstd2 * h4.y + u4.y;
            y2 = (y2 - mean2) * rstd2 * h4.z + u4.z;
            y3 = (y3 - mean2) * rstd2 * h4.w + u4.w;
          }
          o[i >> 1] = pack_bf16x2(y0, y1);
          o[(i >> 1) + 1] = pack_bf16x2(y2, y3);
        }
        if (valid) {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<uint4*>(orow + c * 32 + q * 8) = make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
        }
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after_sync();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace

cudaError_t launch_enc_ffn_fused(const CUtensorMap& tm_x, const CUtensorMap& tm_w1, const CUtensorMap& tm_w2,
                                 const EncFfnParams& p, cudaStream_t stream) {
  if (p.M <= 0) return cudaSuccess;
  if (p.n_slices < 1 || !p.b1 || !p.b2 || !p.residual || !p.gamma || !p.beta || !p.out) return cudaErrorInvalidValue;
  int num_sms = 0;
  cudaError_t e = current_device_sms(&num_sms);
  if (e != cudaSuccess) return e;
  const int tiles = (p.M + 127) / 128;
  const int grid = tiles < num_sms ? tiles : num_sms;
  if (p.gamma2) {
    auto kfn = enc_ffn_fused_kernel<true>;
    e = ensure_dyn_smem(reinterpret_cast<const void*>(kfn), EfSmem::kBytes);
    if (e != cudaSuccess) return e;
    return launch_kernel(kfn, dim3(grid), dim3(kEfThreads), EfSmem::kBytes, stream, false, tm_x, tm_w1, tm_w2, p);
  }
  auto kfn = enc_ffn_fused_kernel<false>;
  e = ensure_dyn_smem(reinterpret_cast<const void*>(kfn), EfSmem::kBytes);
  if (e != cudaSuccess) return e;
  return launch_kernel(kfn, dim3(grid), dim3(kEfThreads), EfSmem::kBytes, stream, false, tm_x, tm_w1, tm_w2, p);
}

}  // namespace b200vqa
