// Encoder feed-forward block in ONE kernel (torch TransformerEncoderLayer._ff_block + norm2, IQAP:173 / FA:42):
//     y = LayerNorm(x + W2 relu(W1 x + b1) + b2)        [+ nn.Transformer's final encoder norm on top, FA:42]
// for M = questions x 256 rows.  As a GEMM pair the hidden activations [M, ff] go to HBM and come back (IQAP, 1024
// questions: 1.07 GB written + 1.07 GB read for 0.13 GB of input).  Here a persistent CTA PAIR (cta_group::2, the two
// SMs of a TPC) owns 256 rows - 128 per CTA - and walks the hidden dimension in slices of 128 units, entirely on chip:
//     H_j = relu(X . W1[j]^T + b1[j])     tcgen05 M = 256, N = 128; fp32 accumulator in TMEM (two buffers) -> bf16 in
//                                         shared memory (two buffers), each CTA its own 128 rows
//     Y  += H_j . W2[:, j]^T              tcgen05 M = 256, N = 256; accumulator resident in TMEM for the whole tile
// and the epilogue adds b2 + residual and normalises the row.  HBM sees x once and y once.
//
// Why a pair: every 128-row tile needs ALL of W1 and W2 (2 * 256 * ff bytes: 2 MB at ff = 2048).  One CTA alone pulls
// them through its own shared memory and reads them again for the MMA - 384 KB of shared-memory traffic per slice against
// 2048 cycles of tensor work at 128 B/clk: the shared-memory pipe, not the tensor pipe, set the pace (measured: no
// faster than the GEMM pair through HBM).  With cta_group::2 each CTA holds HALF of every weight tile (the B operand is
// split by N across the pair) and the tensor cores of both SMs read both halves.
//
// Warp roles (512 threads per CTA, aligned to warpgroups for setmaxnreg): warp 0 TMA producer (its CTA's x rows and weight halves; ring of six 16-KB units;
// completion is counted on the LEADER's mbarriers), warp 1 MMA issuer (leader CTA only; tcgen05.commit multicasts to
// the mbarriers of both CTAs), warps 4-7 hidden epilogue (TMEM -> bias, ReLU -> swizzled K-major bf16 tile), warps 8-15
// LayerNorm epilogue (two warps per TMEM lane quarter, 128 columns each).  GEMM1 of slice j+1 is issued before GEMM2 of
// slice j, so the tensor pipe never waits for the hidden epilogue; the LayerNorm epilogue of tile i runs under the
// first slices of tile i+1.
#include <cstdio>
#include <cstdlib>

#include "host_util.h"
#include "kernels.h"
#include "ptx.cuh"

namespace b200vqa {
namespace {

constexpr int kEfE1PerQuarter = 1;  // hidden-epilogue warps per TMEM lane quarter (2 measured no faster: the chain is
                                    // bound by the shared-memory pipe it shares with the MMA operand reads, not by issue)
constexpr int kEfE1Warps = 4 * kEfE1PerQuarter;
constexpr int kEfThreads = (4 + kEfE1Warps + 8) * 32;
constexpr int kEfE1Cols = 128 / kEfE1PerQuarter;  // columns of a slice per hidden-epilogue thread
constexpr int kEfRegE1 = kEfE1PerQuarter == 1 ? 104 : 64;
constexpr int kEfRegLN = kEfE1PerQuarter == 1 ? 184 : 152;
constexpr int kEfRing = 6;
constexpr int kEfUnit = 16384;

struct EfSmem {
  static constexpr int kOffX = 0;                       // this CTA's x rows: 4 k-blocks of [128 rows x 64]
  static constexpr int kOffH = 65536;                   // 2 x H slice: 2 panels of [128 rows x 64 hidden]
  static constexpr int kOffRing = kOffH + 2 * 32768;    // weight units (this CTA's half of each)
  static constexpr int kOffBar = kOffRing + kEfRing * kEfUnit;
  static constexpr int kOffStats = kOffBar + 256;       // float [2 column halves][128 rows]
  static constexpr int kOffB1 = kOffStats + 1024;       // 2 x this slice's 128 linear1 biases
  static constexpr int kBytes = kOffB1 + 1024;
};
static_assert(EfSmem::kBytes <= 232448, "shared memory budget");

// bf16x2 {lo = max(lo, 0), hi = max(hi, 0)}: conversion and ReLU in one instruction
__device__ __forceinline__ uint32_t pack_relu_bf16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ float4 ldg4f(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
// Streaming accesses bypass L1: with 227 KB of the SM carved out as shared memory ~24 KB of L1 remain, and they are kept
// for the per-column constants (b2, gamma, beta) every LayerNorm thread re-reads for every tile
__device__ __forceinline__ uint4 ld_stream16(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ float ld_stream_f32(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_stream16(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.global.L1::no_allocate.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// Bounded by the clock (about one second): a protocol bug traps instead of hanging the GPU.
__device__ __forceinline__ void ef_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 2000000000ll) {
      printf("b200vqa: enc_ffn_fused wait timed out (block %d thread %d barrier +%d parity %u)\n", blockIdx.x, threadIdx.x,
             int(smem_u32(bar) & 255u), parity);
      __trap();
    }
  }
}

// wait + cycles spent in it (debug counters, B200VQA_ENC_FFN_DBG=1)
__device__ __forceinline__ void ef_wait_t(uint64_t* bar, uint32_t parity, long long& acc, bool on) {
  if (!on) {
    ef_wait(bar, parity);
    return;
  }
  const long long t0 = clock64();
  ef_wait(bar, parity);
  acc += clock64() - t0;
}

// Register reallocation between warpgroups (all four warps of a warpgroup execute the same one)
template <int N>
__device__ __forceinline__ void setmaxnreg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}

template <bool LN2>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kEfThreads, 1)
enc_ffn_fused_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w1,
                     const __grid_constant__ CUtensorMap tm_w2, const EncFfnParams p) {
  using L = EfSmem;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sX = smem + L::kOffX;
  uint8_t* sH = smem + L::kOffH;
  uint8_t* sRing = smem + L::kOffRing;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kOffBar);
  // waited on by the leader's MMA thread only (the peer's copies stay unused):
  uint64_t* ring_full = bars;        // [6] both halves of a weight unit landed (bytes of both CTAs)
  uint64_t* x_full = bars + 6;       // both x tiles landed
  uint64_t* ht_free = bars + 7;      // [2] H accumulator read out by the hidden epilogues of both CTAs (4 warps each)
  uint64_t* hs_full = bars + 9;      // [2] first K panel (64 hidden units) of H (bf16) written to shared memory in both CTAs
  uint64_t* hs_full2 = bars + 25;    // [2] ... second panel: GEMM2 starts on the first while the epilogue converts the second
  uint64_t* y_free = bars + 11;      // Y read out by the LayerNorm epilogues of both CTAs (4 warps each)
  // signalled in both CTAs by the leader's multicast commits:
  uint64_t* ring_empty = bars + 12;  // [6] the MMAs that read the unit have retired
  uint64_t* x_free = bars + 18;      // last GEMM1 of the tile retired
  uint64_t* ht_full = bars + 19;     // [2] H accumulator complete
  uint64_t* hs_free = bars + 21;     // [2] H (bf16) consumed by GEMM2
  uint64_t* y_full = bars + 23;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);
  float* s_stats = reinterpret_cast<float*>(smem + L::kOffStats);
  float* s_b1 = reinterpret_cast<float*>(smem + L::kOffB1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int n_sup = (p.M + 255) / 256;  // 256-row tiles of the pair
  const int n_sl = p.n_slices;
  const int sup0 = blockIdx.x >> 1, sup_step = gridDim.x >> 1;

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
      mbar_init(&ht_free[i], 2 * kEfE1Warps);
      mbar_init(&hs_full[i], 8);   // four warps per panel and CTA
      mbar_init(&hs_full2[i], 8);
      mbar_init(&hs_free[i], 1);
    }
    mbar_init(y_full, 1);
    mbar_init(y_free, 16);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_pair<512>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  // both CTAs' barriers exist before anything is signalled across the pair
  cluster_arrive_release();
  cluster_wait_acquire();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t dY = tmem_base;  // 256 columns; H accumulators: 128 columns each at +256 and +384

  // 512 threads start with 128 registers each; the LayerNorm warps need their 128 accumulator values in registers (see
  // below) and take what the producer / MMA warpgroup and the hidden epilogue release: 40 + 104 + 2 x 184 = 4 x 128
  // (only what the CTA itself releases can be re-acquired: asking for more blocks setmaxnreg.inc for ever).
  // (setmaxnreg sits at the top of each role's branch: that is where ptxas takes the branch's register budget from)
  if (warp == 0) {
    setmaxnreg_dec<40>();
    {
      // (whole warp walks the loop, one elected lane issues: the TMA operands stay in uniform registers)
      const bool el = elect_one();
      uint32_t u = 0, t = 0;
      const bool dbg = p.dbg != nullptr && blockIdx.x == 0;
      long long w_empty = 0;
      auto slot_wait = [&]() -> uint32_t {
        const uint32_t slot = u % kEfRing;
        ef_wait_t(&ring_empty[slot], ((u / kEfRing) & 1) ^ 1, w_empty, dbg);
        if (leader && el) mbar_expect_tx(&ring_full[slot], 2 * kEfUnit);
        ++u;
        return slot;
      };
      // GEMM1 unit: this CTA's 64 of the slice's 128 hidden units x 128 of K (two k-blocks of [64 x 64])
      auto put_w1 = [&](int j, int half) {
        const uint32_t slot = slot_wait();
        const int row = j * 128 + int(rank) * 64;
        if (el) {
          tma_load_2d_pair(&tm_w1, &ring_full[slot], sRing + slot * kEfUnit, (2 * half) * 64, row);
          tma_load_2d_pair(&tm_w1, &ring_full[slot], sRing + slot * kEfUnit + 8192, (2 * half + 1) * 64, row);
        }
        __syncwarp();
      };
      // GEMM2 unit: this CTA's 128 of the 256 output columns x 64 hidden units
      auto put_w2 = [&](int j, int half) {
        const uint32_t slot = slot_wait();
        if (el) tma_load_2d_pair(&tm_w2, &ring_full[slot], sRing + slot * kEfUnit, j * 128 + half * 64, int(rank) * 128);
        __syncwarp();
      };
      for (int sup = sup0; sup < n_sup; sup += sup_step, ++t) {
        ef_wait(x_free, (t & 1) ^ 1);
        if (el) {
          if (leader) mbar_expect_tx(x_full, 2 * 65536);
#pragma unroll
          for (int kb = 0; kb < 4; ++kb)
            tma_load_2d_pair(&tm_x, x_full, sX + kb * 16384, kb * 64, sup * 256 + int(rank) * 128);
        }
        __syncwarp();
        // the order the MMA thread consumes them in
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
      if (dbg && lane == 0) p.dbg[15] = w_empty;
    }
  } else if (warp == 1) {
    setmaxnreg_dec<40>();
    if (leader) {
      // The WHOLE warp runs this loop (waits included) and one elected lane issues: with uniform control flow the
      // descriptors live in uniform registers.  Under `if (lane == 0)` the compiler wraps every tcgen05.mma in an
      // elect / R2UR.BROADCAST loop - ~16 instructions and ~100 cycles per MMA, more than the 64 cycles an
      // M = 256, N = 128 MMA takes.
      const bool el = elect_one();
      constexpr uint32_t idesc1 = make_idesc(kFmtBF16, 256, 128, 0, 0);
      constexpr uint32_t idesc2 = make_idesc(kFmtBF16, 256, 256, 0, 0);
      // descriptors = base + (byte offset >> 4): the tiles are 1024-byte aligned and below 256 KB, so the 14-bit address
      // field never carries
      const uint64_t dx_base = make_smem_desc_sw128(smem_u32(sX), 16, 1024);
      const uint64_t dh_base = make_smem_desc_sw128(smem_u32(sH), 16, 1024);
      const uint64_t dr_base = make_smem_desc_sw128(smem_u32(sRing), 16, 1024);
      uint32_t u = 0, s = 0, t = 0;
      const bool dbg = p.dbg != nullptr && blockIdx.x == 0;
      long long w_ring = 0, w_htfree = 0, w_hsfull = 0, w_yfree = 0, w_xfull = 0;
      const long long t_begin = clock64();
      auto unit_wait = [&]() -> uint32_t {
        const uint32_t slot = u % kEfRing;
        ef_wait_t(&ring_full[slot], (u / kEfRing) & 1, w_ring, dbg);
        tc_fence_after_sync();
        ++u;
        return slot;
      };
      auto gemm1 = [&](uint32_t sidx) {  // H[sidx & 1] = X . W1[j]^T : M=256, N=128, K=256
        const uint32_t buf = sidx & 1;
        ef_wait_t(&ht_free[buf], ((sidx >> 1) & 1) ^ 1, w_htfree, dbg);
        tc_fence_after_sync();
        const uint32_t dH = tmem_base + 256 + buf * 128;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const uint32_t slot = unit_wait();
          const uint64_t db = dr_base + uint64_t((slot * kEfUnit) >> 4);
          if (el) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int kk = half * 8 + k;
              umma_bf16_pair(dH, dx_base + uint64_t(((kk / 4) * 16384 + (kk % 4) * 32) >> 4),
                             db + uint64_t(((k / 4) * 8192 + (k % 4) * 32) >> 4), idesc1, kk != 0);
            }
            umma_commit_pair(&ring_empty[slot]);
          }
          __syncwarp();
        }
        if (el) umma_commit_pair(&ht_full[buf]);
        __syncwarp();
      };
      auto gemm2 = [&](int j, uint32_t sidx) {  // Y += H[sidx & 1] . W2[:, j]^T : M=256, N=256, K=128
        const uint32_t buf = sidx & 1;
        if (j == 0) ef_wait_t(y_free, (t & 1) ^ 1, w_yfree, dbg);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          ef_wait_t(half == 0 ? &hs_full[buf] : &hs_full2[buf], (sidx >> 1) & 1, w_hsfull, dbg);
          tc_fence_after_sync();
          const uint32_t slot = unit_wait();
          const uint64_t da = dh_base + uint64_t((buf * 32768 + half * 16384) >> 4);
          const uint64_t db = dr_base + uint64_t((slot * kEfUnit) >> 4);
          if (el) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_pair(dY, da + uint64_t((k * 32) >> 4), db + uint64_t((k * 32) >> 4), idesc2,
                             !(j == 0 && half == 0 && k == 0));
            umma_commit_pair(&ring_empty[slot]);
          }
          __syncwarp();
        }
        if (el) umma_commit_pair(&hs_free[buf]);
        __syncwarp();
      };
      for (int sup = sup0; sup < n_sup; sup += sup_step, ++t) {
        ef_wait_t(x_full, t & 1, w_xfull, dbg);
        tc_fence_after_sync();
        gemm1(s);
        if (n_sl == 1 && el) umma_commit_pair(x_free);
        for (int j = 0; j < n_sl; ++j) {
          if (j + 1 < n_sl) {
            gemm1(s + j + 1);
            // the last GEMM1 of the tile: x may be replaced when it retires
            if (j + 2 == n_sl && el) umma_commit_pair(x_free);
          }
          gemm2(j, s + j);
        }
        if (el) umma_commit_pair(y_full);
        __syncwarp();
        s += n_sl;
      }
      if (dbg && lane == 0) {
        p.dbg[0] = clock64() - t_begin;
        p.dbg[1] = w_ring;
        p.dbg[2] = w_htfree;
        p.dbg[3] = w_hsfull;
        p.dbg[4] = w_yfree;
        p.dbg[5] = w_xfull;
        p.dbg[6] = t;
      }
    }
  } else if (warp < 4) {
    setmaxnreg_dec<40>();  // (two spare warps: the roles are aligned to warpgroups for setmaxnreg)
  } else if (warp < 4 + kEfE1Warps) {
    setmaxnreg_dec<kEfRegE1>();
    // ---- hidden epilogue: row r of this CTA's tile == TMEM lane r
    const int quarter = warp & 3;
    const int ch = (warp - 4) >> 2;  // column group of the slice (with two groups: == 64-unit K panel of the H tile)
    const int r = quarter * 32 + lane;
    const int e_tid = threadIdx.x - 128;
    const uint32_t lane_off = uint32_t(quarter * 32) << 16;
    uint32_t s = 0;
    const bool dbg = p.dbg != nullptr && blockIdx.x == 0 && warp == 4 && lane == 0;
    long long w_hsfree = 0, w_htfull = 0, t_top = 0, t_body = 0;
    const long long t_begin = clock64();
    // The slice's biases go through shared memory (two buffers, one value per thread of the first four warps, fetched
    // one slice ahead): with 227 KB of the SM carved out as shared memory there is next to no L1, and a global load in
    // the loop below is an L2 round trip per 32 columns
    float b_next = e_tid < 128 ? ld_stream_f32(p.b1 + e_tid) : 0.f;
    for (int sup = sup0; sup < n_sup; sup += sup_step) {
      for (int j = 0; j < n_sl; ++j, ++s) {
        const uint32_t buf = s & 1, ph = (s >> 1) & 1;
        float* sb = s_b1 + buf * 128;
        uint8_t* hbuf = sH + buf * 32768;
        long long q0 = 0, q1 = 0, q2 = 0;
        if (dbg) q0 = clock64();
        if (e_tid < 128) {
          sb[e_tid] = b_next;  // its readers of slice s - 2 passed the barrier of slice s - 1
          b_next = ld_stream_f32(p.b1 + (j + 1 < n_sl ? j + 1 : 0) * 128 + e_tid);
        }
        named_bar_sync(5, kEfE1Warps * 32);
        if (dbg) q1 = clock64();
        ef_wait_t(&hs_free[buf], ph ^ 1, w_hsfree, dbg);  // GEMM2 of slice s - 2 no longer reads this buffer
        ef_wait_t(&ht_full[buf], ph, w_htfull, dbg);
        __syncwarp();
        tc_fence_after_sync();
        const uint32_t dH = tmem_base + 256 + buf * 128 + ch * kEfE1Cols + lane_off;
        if (dbg) q2 = clock64();
#pragma unroll
        for (int c = 0; c < kEfE1Cols / 32; ++c) {
          uint32_t v[32];
          tmem_ld32(dH + c * 32, v);
          tmem_ld_wait();
          // K-major, 128-byte swizzled A operand: panel = 64 hidden units, 16-byte chunk index XOR (row & 7)
          const int cc = ch * (kEfE1Cols / 32) + c;  // 32-column chunk of the slice
          uint8_t* prow = hbuf + (cc >> 1) * 16384 + r * 128;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 ba = *reinterpret_cast<const float4*>(sb + cc * 32 + q * 8);
            const float4 bb = *reinterpret_cast<const float4*>(sb + cc * 32 + q * 8 + 4);
            uint4 o;
            o.x = pack_relu_bf16x2(__uint_as_float(v[q * 8 + 0]) + ba.x, __uint_as_float(v[q * 8 + 1]) + ba.y);
            o.y = pack_relu_bf16x2(__uint_as_float(v[q * 8 + 2]) + ba.z, __uint_as_float(v[q * 8 + 3]) + ba.w);
            o.z = pack_relu_bf16x2(__uint_as_float(v[q * 8 + 4]) + bb.x, __uint_as_float(v[q * 8 + 5]) + bb.y);
            o.w = pack_relu_bf16x2(__uint_as_float(v[q * 8 + 6]) + bb.z, __uint_as_float(v[q * 8 + 7]) + bb.w);
            const int chunk = (cc & 1) * 4 + q;
            *reinterpret_cast<uint4*>(prow + ((chunk ^ (r & 7)) << 4)) = o;
          }
          if ((cc & 1) && c + 1 < kEfE1Cols / 32) {
            // the first 64-unit K panel is complete: GEMM2 of this slice starts on it under the conversion of the second
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(cluster_map_shared(smem_u32(&hs_full[buf]), 0));
          }
        }
        // (the last tcgen05.ld completed above) GEMM1 of slice s + 2 may overwrite H; GEMM2 of this slice may read the
        // panel just finished
        tc_fence_before_sync();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive_cluster(cluster_map_shared(smem_u32(&ht_free[buf]), 0));
          mbar_arrive_cluster(cluster_map_shared(smem_u32(kEfE1PerQuarter == 2 && ch == 0 ? &hs_full[buf] : &hs_full2[buf]), 0));
        }
        if (dbg) {
          t_top += q1 - q0;
          t_body += clock64() - q2;
        }
      }
    }
    if (dbg) {
      p.dbg[8] = clock64() - t_begin;
      p.dbg[9] = w_hsfree;
      p.dbg[10] = w_htfull;
      p.dbg[11] = t_top;
      p.dbg[14] = t_body;
    }
  } else {
    setmaxnreg_inc<kEfRegLN>();
    // ---- LayerNorm epilogue: warps w and w + 4 share a TMEM lane quarter and split the 256 columns.  The thread's 128
    // accumulator values are read into registers in one go and Y is handed back at once (GEMM2 of the next tile starts
    // two slices after this one ends); bias, residual, statistics and the store then run under the next tile's MMAs.
    const int quarter = warp & 3;
    const int half = (warp - 4 - kEfE1Warps) >> 2;
    const int rt = quarter * 32 + lane;  // row inside this CTA's tile
    const uint32_t taddr = dY + (uint32_t(quarter * 32) << 16) + half * 128;
    const float* b2 = p.b2 + half * 128;
    const float* g1 = p.gamma + half * 128;
    const float* t1 = p.beta + half * 128;
    uint32_t t = 0;
    const bool dbg = p.dbg != nullptr && blockIdx.x == 0 && warp == 4 + kEfE1Warps && lane == 0;
    long long w_yfull = 0;
    const long long t_begin = clock64();
    // row statistics from the two column halves (sum and sum of squares in fp32; the rows are O(1) activations)
    auto combine = [&](float s1, float s2, float& mean, float& rstd) {
      s_stats[half * 128 + rt] = s1;
      named_bar_sync(1 + quarter, 64);
      const float o1 = s_stats[(half ^ 1) * 128 + rt];
      named_bar_sync(1 + quarter, 64);  // both have read before the slot is reused
      s_stats[half * 128 + rt] = s2;
      named_bar_sync(1 + quarter, 64);
      const float o2 = s_stats[(half ^ 1) * 128 + rt];
      named_bar_sync(1 + quarter, 64);
      mean = (s1 + o1) * (1.f / kD);
      const float var = fmaxf((s2 + o2) * (1.f / kD) - mean * mean, 0.f);
      rstd = rsqrtf(var + p.eps);
    };
    for (int sup = sup0; sup < n_sup; sup += sup_step, ++t) {
      const int row = sup * 256 + int(rank) * 128 + rt;
      const bool valid = row < p.M;
      const __nv_bfloat16* res = p.residual + size_t(valid ? row : 0) * kD + half * 128;
      ef_wait_t(y_full, t & 1, w_yfull, dbg);
      __syncwarp();
      tc_fence_after_sync();
      uint32_t acc[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld32(taddr + c * 32, acc[c]);
      tmem_ld_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(cluster_map_shared(smem_u32(y_free), 0));
      // v = acc + b2 + residual, in place
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint4 rr[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) rr[q] = ld_stream16(res + c * 32 + q * 8);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t w4[4] = {rr[q].x, rr[q].y, rr[q].z, rr[q].w};
          const float4 c0 = ldg4f(b2 + c * 32 + q * 8), c1 = ldg4f(b2 + c * 32 + q * 8 + 4);
          const float bb[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 x2 = unpack_bf16x2(w4[e]);
            const float v0 = __uint_as_float(acc[c][q * 8 + 2 * e]) + bb[2 * e] + x2.x;
            const float v1 = __uint_as_float(acc[c][q * 8 + 2 * e + 1]) + bb[2 * e + 1] + x2.y;
            acc[c][q * 8 + 2 * e] = __float_as_uint(v0);
            acc[c][q * 8 + 2 * e + 1] = __float_as_uint(v1);
            s1 += v0 + v1;
            s2 = fmaf(v0, v0, fmaf(v1, v1, s2));
          }
        }
      }
      float mean, rstd;
      combine(s1, s2, mean, rstd);
      if constexpr (LN2) {
        // second LayerNorm on top (nn.Transformer's final encoder norm): normalise in place, statistics of the result
        s1 = 0.f;
        s2 = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 g4 = ldg4f(g1 + c * 32 + i), t4 = ldg4f(t1 + c * 32 + i);
            const float y0 = (__uint_as_float(acc[c][i]) - mean) * rstd * g4.x + t4.x;
            const float y1 = (__uint_as_float(acc[c][i + 1]) - mean) * rstd * g4.y + t4.y;
            const float y2 = (__uint_as_float(acc[c][i + 2]) - mean) * rstd * g4.z + t4.z;
            const float y3 = (__uint_as_float(acc[c][i + 3]) - mean) * rstd * g4.w + t4.w;
            acc[c][i] = __float_as_uint(y0);
            acc[c][i + 1] = __float_as_uint(y1);
            acc[c][i + 2] = __float_as_uint(y2);
            acc[c][i + 3] = __float_as_uint(y3);
            s1 += (y0 + y1) + (y2 + y3);
            s2 = fmaf(y0, y0, fmaf(y1, y1, fmaf(y2, y2, fmaf(y3, y3, s2))));
          }
        }
        combine(s1, s2, mean, rstd);
      }
      const float* gg = LN2 ? p.gamma2 + half * 128 : g1;
      const float* tt = LN2 ? p.beta2 + half * 128 : t1;
      __nv_bfloat16* orow = p.out + size_t(valid ? row : 0) * kD + half * 128;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t o[16];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 g4 = ldg4f(gg + c * 32 + i), t4 = ldg4f(tt + c * 32 + i);
          const float y0 = (__uint_as_float(acc[c][i]) - mean) * rstd * g4.x + t4.x;
          const float y1 = (__uint_as_float(acc[c][i + 1]) - mean) * rstd * g4.y + t4.y;
          const float y2 = (__uint_as_float(acc[c][i + 2]) - mean) * rstd * g4.z + t4.z;
          const float y3 = (__uint_as_float(acc[c][i + 3]) - mean) * rstd * g4.w + t4.w;
          o[i >> 1] = pack_bf16x2(y0, y1);
          o[(i >> 1) + 1] = pack_bf16x2(y2, y3);
        }
        if (valid) {
#pragma unroll
          for (int q = 0; q < 4; ++q) st_stream16(orow + c * 32 + q * 8, o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
        }
      }
    }
    if (dbg) {
      p.dbg[12] = clock64() - t_begin;
      p.dbg[13] = w_yfull;
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

cudaError_t launch_enc_ffn_fused(const CUtensorMap& tm_x, const CUtensorMap& tm_w1, const CUtensorMap& tm_w2,
                                 const EncFfnParams& p, cudaStream_t stream) {
  if (p.M <= 0) return cudaSuccess;
  if (p.n_slices < 1 || !p.b1 || !p.b2 || !p.residual || !p.gamma || !p.beta || !p.out) return cudaErrorInvalidValue;
  int num_sms = 0;
  cudaError_t e = current_device_sms(&num_sms);
  if (e != cudaSuccess) return e;
  const int sups = (p.M + 255) / 256;
  const int pairs = sups < num_sms / 2 ? sups : num_sms / 2;
  const int grid = 2 * pairs;  // the cluster shape (2,1,1) is part of the kernel
  if (getenv("B200VQA_ENC_FFN_DBG") && !p.dbg) {
    // cycles the roles of CTA 0 spend waiting on each other (synchronous: debugging only)
    long long* d = nullptr;
    long long hbuf[20] = {};
    if (cudaMalloc(&d, sizeof(hbuf)) != cudaSuccess) return cudaErrorMemoryAllocation;
    cudaMemset(d, 0, sizeof(hbuf));
    EncFfnParams q = p;
    q.dbg = d;
    e = launch_enc_ffn_fused(tm_x, tm_w1, tm_w2, q, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e == cudaSuccess) e = cudaMemcpy(hbuf, d, sizeof(hbuf), cudaMemcpyDeviceToHost);
    cudaFree(d);
    fprintf(stderr,
            "b200vqa enc_ffn_fused (M %d, %d slices): MMA thread %lld cycles over %lld tiles: waits ring %lld, H-tmem free "
            "%lld, H-smem full %lld, Y free %lld, x %lld | hidden epilogue %lld: waits H-smem free %lld, H-tmem full %lld | "
            "[bias staging + barrier %lld, tmem -> smem + signal %lld] | LN epilogue %lld: waits Y full %lld | producer waits empty "
            "%lld\n",
            p.M, p.n_slices, hbuf[0], hbuf[6], hbuf[1], hbuf[2], hbuf[3], hbuf[4], hbuf[5], hbuf[8], hbuf[9], hbuf[10],
            hbuf[11], hbuf[14], hbuf[12], hbuf[13], hbuf[15]);
    return e;
  }
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
