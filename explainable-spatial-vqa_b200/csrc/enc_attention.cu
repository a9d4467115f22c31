// Fused encoder self-attention on tcgen05 for the ~200-token image+program sequence.
//
// Replaces the scaled-dot-product attention inside nn.TransformerEncoderLayer.self_attn
// (IQAP:118,173; FA:42,135 -> torch functional.py:6244+): softmax(q k^T / sqrt(dh)) v per head, with a
// key-length mask for the batched FA path (the reference is batch-1 and never pads).
//
// One CTA per (question, head) handles BOTH 128-query tiles of the sequence (<= 256 rows), so K and V are
// fetched once.  The whole K and V of a head sit in shared memory - no online-softmax rescaling:
//   warp 0      TMA: Q tile 0, Q tile 1, K, V (128B-swizzled boxes of the packed q|k|v activations)
//   warp 1      tcgen05.mma  S0 = Q0 K^T, S1 = Q1 K^T (two 256-column TMEM accumulators), later O0 = P0 V, O1 = P1 V
//   warps 2-9   softmax group of tile 0, warps 10-17 softmax group of tile 1.  A query row is shared by TWO threads
//               (same TMEM lane, alternating 32-column chunks) so the per-row dependent chain is halved; the pair
//               exchanges its row maximum and row sum through shared memory.  tcgen05.ld S, masked softmax in fp32
//               (exp2), P -> smem as the bf16 K-major A operand of the second MMA, then O / rowsum -> bf16 global
// The groups run concurrently on all four scheduler partitions, and tile 0's P.V overlaps tile 1's softmax.
// V is consumed directly as an MN-major B operand (no transpose).  P tiles reuse the shared memory of K and Q
// once both score MMAs have retired.
#include "kernels.h"
#include "ptx.cuh"

namespace b200vqa {
namespace {

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

constexpr int kAttnThreads = 64 + 512;  // TMA warp, MMA warp, 2 tiles x 8 softmax warps
constexpr int kKeysMax = 256;

template <int DH>
struct AttnSmem {
  static constexpr int kPanels = DH / 64;
  static constexpr int kQ = 128 * DH * 2;       // one query tile: 16 / 32 KB
  static constexpr int kK = kKeysMax * DH * 2;  // 32 / 64 KB
  static constexpr int kV = kK;
  static constexpr int kP = 128 * kKeysMax * 2;  // 64 KB per probability tile
  // dh = 64 : region A = [K | Q0 | Q1] (64 KB, later P0), region B = P1 (64 KB), V (32 KB)          -> 160 KB
  // dh = 128: region A = K (64 KB, later P0), region B = [Q0 | Q1] (64 KB, later P1), V (64 KB)     -> 192 KB
  static constexpr int kOffK = 0;
  static constexpr int kOffQ0 = (DH == 64) ? kK : kP;
  static constexpr int kOffQ1 = kOffQ0 + kQ;
  static constexpr int kOffP0 = 0;
  static constexpr int kOffP1 = kP;
  static constexpr int kOffV = 2 * kP;
  static constexpr int kOffXch = kOffV + kV;          // [2 tiles][2 halves][128 rows] float2 (max, sum)
  static constexpr int kOffBar = kOffXch + 2 * 2 * 128 * 8;
  static constexpr int kBytes = kOffBar + 128;
};

template <int DH>
__global__ void __launch_bounds__(kAttnThreads, 1)
enc_attention_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                     const EncAttnParams p) {
  using L = AttnSmem<DH>;
  constexpr int kPanels = L::kPanels;

  const int h = blockIdx.x % p.nhead;
  const int b = blockIdx.x / p.nhead;
  int len = p.lens ? p.lens[b] : p.const_len;
  len = len < 1 ? 1 : (len > kKeysMax ? kKeysMax : len);
  const bool two = len > 128;  // the second query tile holds valid rows
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const size_t row0 = size_t(b) * kLP;

  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sK = smem + L::kOffK;
  uint8_t* sQ[2] = {smem + L::kOffQ0, smem + L::kOffQ1};
  uint8_t* sP[2] = {smem + L::kOffP0, smem + L::kOffP1};
  uint8_t* sV = smem + L::kOffV;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kOffBar);
  uint64_t* bar_qk = bars + 0;
  uint64_t* bar_v = bars + 1;
  uint64_t* bar_s = bars + 2;  // [2] score accumulator of tile t complete
  uint64_t* bar_p = bars + 4;  // [2] P of tile t written (256 arrivals: two threads per row)
  uint64_t* bar_o = bars + 6;  // [2] output accumulator of tile t complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int keys16 = (len + 15) & ~15;  // keys processed by the tensor core (multiple of the UMMA K / N step)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_kv);
    mbar_init(bar_qk, 1);
    mbar_init(bar_v, 1);
    for (int t = 0; t < 2; ++t) {
      mbar_init(&bar_s[t], 1);
      mbar_init(&bar_p[t], 256);
      mbar_init(&bar_o[t], 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {  // uniform control flow up to here: the TMA operands come from uniform registers
      mbar_expect_tx(bar_qk, 2 * L::kQ + L::kK);
#pragma unroll
      for (int pn = 0; pn < kPanels; ++pn) {
        tma_load_2d(&tm_kv, bar_qk, sK + pn * 32768, kD + h * DH + pn * 64, int(row0));
        tma_load_2d(&tm_q, bar_qk, sQ[0] + pn * 16384, h * DH + pn * 64, int(row0));
        tma_load_2d(&tm_q, bar_qk, sQ[1] + pn * 16384, h * DH + pn * 64, int(row0) + 128);
      }
      mbar_expect_tx(bar_v, L::kV);
#pragma unroll
      for (int pn = 0; pn < kPanels; ++pn)
        tma_load_2d(&tm_kv, bar_v, sV + pn * 32768, 2 * kD + h * DH + pn * 64, int(row0));
    }
    __syncwarp();
  } else if (warp == 1) {
    // The whole warp waits and one elected lane issues: the descriptors stay in uniform registers (see gemm_tc_kernel;
    // the N = DH MMAs of P V take 32-64 cycles each, issued from one thread's vector registers ~100)
    const bool el = elect_one();
    const uint64_t dk0 = make_smem_desc_sw128(smem_u32(sK), 16, 1024);
    const uint64_t dv0 = make_smem_desc_sw128(smem_u32(sV), 32768, 1024);
    // ---- S_t = Q_t K^T : M=128, N=keys16, K=DH
    mbar_wait(bar_qk, 0);
    tc_fence_after_sync();
    const uint32_t idesc_s = make_idesc(kFmtBF16, 128, uint32_t(keys16), 0, 0);
    for (int t = 0; t < 2; ++t) {
      const uint64_t dq0 = make_smem_desc_sw128(smem_u32(sQ[t]), 16, 1024);
      if (el) {
        if (t == 0 || two) {
#pragma unroll
          for (int k = 0; k < DH / 16; ++k)
            umma_bf16(tmem_base + t * 256, dq0 + uint64_t(((k / 4) * 16384 + (k % 4) * 32) >> 4),
                      dk0 + uint64_t(((k / 4) * 32768 + (k % 4) * 32) >> 4), idesc_s, k != 0);
        }
        umma_commit(&bar_s[t]);  // bar_s[1] also tells the softmax groups that K and Q are no longer read
      }
      __syncwarp();
    }
    // ---- O_t = P_t V : M=128, N=DH, K=keys16   (V is the MN-major B operand)
    const uint32_t idesc_o = make_idesc(kFmtBF16, 128, DH, 0, 1);
    const int nkk = keys16 / 16;
    mbar_wait(bar_v, 0);
    for (int t = 0; t < 2; ++t) {
      if (t == 1 && !two) break;
      mbar_wait(&bar_p[t], 0);
      tc_fence_after_sync();
      const uint64_t dp0 = make_smem_desc_sw128(smem_u32(sP[t]), 16, 1024);
      if (el) {
        for (int kk = 0; kk < nkk; ++kk)  // 16 key rows of 128 B per step of V
          umma_bf16(tmem_base + t * 256, dp0 + uint64_t(((kk / 4) * 16384 + (kk % 4) * 32) >> 4),
                    dv0 + uint64_t((kk * 2048) >> 4), idesc_o, kk != 0);
        umma_commit(&bar_o[t]);
      }
      __syncwarp();
    }
  } else {
    const int t = (warp - 2) >> 3;          // query tile of this softmax group
    const int hc = ((warp - 2) >> 2) & 1;   // which half of the row's 32-column chunks (alternating) this thread takes
    const int quarter = warp & 3;           // TMEM lane quarter this warp may access
    const int r = quarter * 32 + lane;      // query row inside the tile == TMEM lane
    const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(t * 256);
    const float sl2 = p.scale * 1.4426950408889634f;
    __nv_bfloat16* orow = p.out + (row0 + t * 128 + r) * kD + h * DH;
    float2* xch = reinterpret_cast<float2*>(smem + L::kOffXch);  // [t][hc][r]
    float2* mine = xch + (t * 2 + hc) * 128 + r;
    const float2* theirs = xch + (t * 2 + (hc ^ 1)) * 128 + r;
    const uint32_t pair_bar = 1 + t * 4 + quarter;  // the two warps sharing these 32 rows
    constexpr int kOutCols = DH / 2;                 // output columns written by this thread

    if (t == 1 && !two) {
      // a tile that holds only padding rows: keep them finite (they are masked as keys downstream)
#pragma unroll
      for (int c = 0; c < kOutCols / 8; ++c) reinterpret_cast<uint4*>(orow + hc * kOutCols)[c] = make_uint4(0, 0, 0, 0);
    } else {
      mbar_wait(&bar_s[t], 0);
      __syncwarp();
      tc_fence_after_sync();
      const int nchunks = (len + 31) / 32;
      const int nfull = len / 32;  // chunks without masked keys: the per-element length test is paid by the last one only

      float mx = -INFINITY;
      for (int c = hc; c < nchunks; c += 2) {
        uint32_t v[32];
        tmem_ld32(taddr + c * 32, v);
        tmem_ld_wait();
        if (c < nfull) {
#pragma unroll
          for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (c * 32 + j < len) mx = fmaxf(mx, __uint_as_float(v[j]));
        }
      }
      mine->x = mx;
      named_bar_sync(pair_bar, 64);
      mx = fmaxf(mx, theirs->x);
      // P_t overwrites K / Q: both score MMAs must have retired
      if (t == 0) mbar_wait(&bar_s[1], 0);
      const float mxs = mx * sl2;
      float sum = 0.f;
      for (int c = hc; c < nchunks; c += 2) {
        uint32_t v[32];
        tmem_ld32(taddr + c * 32, v);
        tmem_ld_wait();
        uint32_t o[16];
        const bool full = c < nfull;
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          // scores are <= the row maximum: the argument is <= 0, ex2.approx needs no range handling
          float p0 = ex2_approx(fmaf(__uint_as_float(v[j]), sl2, -mxs));
          float p1 = ex2_approx(fmaf(__uint_as_float(v[j + 1]), sl2, -mxs));
          if (!full) {
            p0 = (c * 32 + j < len) ? p0 : 0.f;
            p1 = (c * 32 + j + 1 < len) ? p1 : 0.f;
          }
          const uint32_t pb = pack_bf16x2(p0, p1);
          // normalise by what the tensor core will actually sum: the bf16-rounded weights
          sum += __uint_as_float(pb << 16) + __uint_as_float(pb & 0xffff0000u);
          o[j >> 1] = pb;
        }
        uint8_t* prow = sP[t] + (c >> 1) * 16384 + r * 128;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int chunk = (c & 1) * 4 + q;
          *reinterpret_cast<uint4*>(prow + ((chunk ^ (r & 7)) << 4)) =
              make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
        }
      }
      // (with an odd chunk count the last 32-column half of the final 64-key panel is never written; the PV MMA
      //  reads keys16 <= 32*nchunks columns, so it is never read either)
      mine->y = sum;
      fence_proxy_async_smem();
      tc_fence_before_sync();
      mbar_arrive(&bar_p[t]);

      mbar_wait(&bar_o[t], 0);
      __syncwarp();
      tc_fence_after_sync();
      named_bar_sync(pair_bar, 64);  // partner's partial row sum is visible
      const float inv = 1.f / (sum + theirs->y);
#pragma unroll
      for (int c = 0; c < kOutCols / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(taddr + hc * kOutCols + c * 32, v);
        tmem_ld_wait();
        uint32_t o[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2)
          o[j >> 1] = pack_bf16x2(__uint_as_float(v[j]) * inv, __uint_as_float(v[j + 1]) * inv);
        uint4* dst = reinterpret_cast<uint4*>(orow + hc * kOutCols + c * 32);
#pragma unroll
        for (int q = 0; q < 4; ++q) dst[q] = make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
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


// ------------------------------------------------------------------------------------------
// dh = 64 (IQAP, four heads): ONE query tile per CTA, TWO CTAs per SM.
//
// The kernel above owns its SM alone (160 KB, all 512 TMEM columns) and runs its phases one after the other - load, score
// MMA, two softmax passes over TMEM, P.V MMA, store: 12.4 us per (question, head), tensor pipe 9.5 % busy.  Here a CTA
// handles one 128-query tile: K | Q | spare (64 KB, overwritten by P once the score MMA has retired) + V (32 KB) = 96 KB,
// and the output accumulator reuses the score accumulator's TMEM columns (dead once P is in shared memory), so a CTA
// allocates 256 columns and two CTAs share an SM: one CTA's loads and MMAs run under the other's softmax.  K and V are
// fetched once per tile (twice per head) - from L2, the packed q|k|v rows of a question were just written.
// ------------------------------------------------------------------------------------------
constexpr int kTileThreads = 64 + 256;  // TMA warp, MMA warp, 8 softmax warps (two threads per query row)

struct TileSmem {
  static constexpr int kOffK = 0;                 // 256 keys x 64 x 2 B = 32 KB
  static constexpr int kOffQ = 32768;             // 128 queries = 16 KB
  static constexpr int kOffP = 0;                 // 128 x 256 x 2 B = 64 KB over K | Q | spare
  static constexpr int kOffV = 65536;             // 32 KB
  static constexpr int kOffXch = kOffV + 32768;   // [2 halves][128 rows] float2 (max, sum)
  static constexpr int kOffBar = kOffXch + 2 * 128 * 8;
  static constexpr int kBytes = kOffBar + 64;
};

__global__ void __launch_bounds__(kTileThreads, 2)
enc_attention_tile_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                          const EncAttnParams p) {
  using L = TileSmem;
  constexpr int DH = 64;
  const int t = blockIdx.x & 1;  // query tile
  const int h = (blockIdx.x >> 1) % p.nhead;
  const int b = (blockIdx.x >> 1) / p.nhead;
  int len = p.lens ? p.lens[b] : p.const_len;
  len = len < 1 ? 1 : (len > kKeysMax ? kKeysMax : len);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const size_t row0 = size_t(b) * kLP;

  if (t == 1 && len <= 128) {
    // a tile that holds only padding rows: keep them finite (they are masked as keys downstream)
    if (threadIdx.x < 256) {
      const int r = threadIdx.x >> 1, half = threadIdx.x & 1;
      uint4* dst = reinterpret_cast<uint4*>(p.out + (row0 + 128 + r) * kD + h * DH + half * 32);
#pragma unroll
      for (int c = 0; c < 4; ++c) dst[c] = make_uint4(0, 0, 0, 0);
    }
    return;
  }

  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sK = smem + L::kOffK;
  uint8_t* sQ = smem + L::kOffQ;
  uint8_t* sP = smem + L::kOffP;
  uint8_t* sV = smem + L::kOffV;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kOffBar);
  uint64_t* bar_qk = bars + 0;
  uint64_t* bar_v = bars + 1;
  uint64_t* bar_s = bars + 2;  // score accumulator complete (K and Q are no longer read)
  uint64_t* bar_p = bars + 3;  // P written (256 arrivals)
  uint64_t* bar_o = bars + 4;  // output accumulator complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
  const int keys16 = (len + 15) & ~15;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_kv);
    mbar_init(bar_qk, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_p, 256);
    mbar_init(bar_o, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<256>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {  // uniform control flow up to here: the TMA operands come from uniform registers
      mbar_expect_tx(bar_qk, 32768 + 16384);
      tma_load_2d(&tm_kv, bar_qk, sK, kD + h * DH, int(row0));
      tma_load_2d(&tm_q, bar_qk, sQ, h * DH, int(row0) + t * 128);
      mbar_expect_tx(bar_v, 32768);
      tma_load_2d(&tm_kv, bar_v, sV, 2 * kD + h * DH, int(row0));
    }
    __syncwarp();
  } else if (warp == 1) {
    // The whole warp waits and one elected lane issues: the descriptors stay in uniform registers (see gemm_tc_kernel;
    // the 16 N = 64 MMAs of P V take 32 cycles each, issued from vector registers ~100)
    const bool el = elect_one();
    const uint64_t dq0 = make_smem_desc_sw128(smem_u32(sQ), 16, 1024);
    const uint64_t dk0 = make_smem_desc_sw128(smem_u32(sK), 16, 1024);
    const uint64_t dp0 = make_smem_desc_sw128(smem_u32(sP), 16, 1024);
    const uint64_t dv0 = make_smem_desc_sw128(smem_u32(sV), 32768, 1024);
    mbar_wait(bar_qk, 0);
    tc_fence_after_sync();
    const uint32_t idesc_s = make_idesc(kFmtBF16, 128, uint32_t(keys16), 0, 0);
    if (el) {
#pragma unroll
      for (int k = 0; k < DH / 16; ++k) umma_bf16(tmem_base, dq0 + uint64_t((k * 32) >> 4), dk0 + uint64_t((k * 32) >> 4), idesc_s, k != 0);
      umma_commit(bar_s);
    }
    __syncwarp();
    // O = P V overwrites the score accumulator's first 64 columns: every thread has read its scores (twice) and
    // published P before bar_p completes
    const uint32_t idesc_o = make_idesc(kFmtBF16, 128, DH, 0, 1);
    const int nkk = keys16 / 16;
    mbar_wait(bar_v, 0);
    mbar_wait(bar_p, 0);
    tc_fence_after_sync();
    if (el) {
      for (int kk = 0; kk < nkk; ++kk)
        umma_bf16(tmem_base, dp0 + uint64_t(((kk / 4) * 16384 + (kk % 4) * 32) >> 4), dv0 + uint64_t((kk * 2048) >> 4), idesc_o,
                  kk != 0);
      umma_commit(bar_o);
    }
    __syncwarp();
  } else {
    const int hc = ((warp - 2) >> 2) & 1;   // which half of the row's 32-column chunks (alternating) this thread takes
    const int quarter = warp & 3;           // TMEM lane quarter this warp may access
    const int r = quarter * 32 + lane;      // query row inside the tile == TMEM lane
    const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16);
    const float sl2 = p.scale * 1.4426950408889634f;
    __nv_bfloat16* orow = p.out + (row0 + t * 128 + r) * kD + h * DH;
    float2* xch = reinterpret_cast<float2*>(smem + L::kOffXch);  // [hc][r]
    float2* mine = xch + hc * 128 + r;
    const float2* theirs = xch + (hc ^ 1) * 128 + r;
    const uint32_t pair_bar = 1 + quarter;  // the two warps sharing these 32 rows

    mbar_wait(bar_s, 0);
    __syncwarp();
    tc_fence_after_sync();
    const int nchunks = (len + 31) / 32;
    const int nfull = len / 32;  // chunks without masked keys: the per-element length test is paid by the last one only
    float mx = -INFINITY;
    for (int c = hc; c < nchunks; c += 2) {
      uint32_t v[32];
      tmem_ld32(taddr + c * 32, v);
      tmem_ld_wait();
      if (c < nfull) {
#pragma unroll
        for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (c * 32 + j < len) mx = fmaxf(mx, __uint_as_float(v[j]));
      }
    }
    mine->x = mx;
    named_bar_sync(pair_bar, 64);
    mx = fmaxf(mx, theirs->x);
    const float mxs = mx * sl2;
    float sum = 0.f;
    for (int c = hc; c < nchunks; c += 2) {
      uint32_t v[32];
      tmem_ld32(taddr + c * 32, v);
      tmem_ld_wait();
      uint32_t o[16];
      const bool full = c < nfull;
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        // scores are <= the row maximum: the argument is <= 0, ex2.approx needs no range handling
        float p0 = ex2_approx(fmaf(__uint_as_float(v[j]), sl2, -mxs));
        float p1 = ex2_approx(fmaf(__uint_as_float(v[j + 1]), sl2, -mxs));
        if (!full) {
          p0 = (c * 32 + j < len) ? p0 : 0.f;
          p1 = (c * 32 + j + 1 < len) ? p1 : 0.f;
        }
        const uint32_t pb = pack_bf16x2(p0, p1);
        // normalise by what the tensor core will actually sum: the bf16-rounded weights (a bf16 is the upper half of
        // the fp32 with the same value)
        sum += __uint_as_float(pb << 16) + __uint_as_float(pb & 0xffff0000u);
        o[j >> 1] = pb;
      }
      uint8_t* prow = sP + (c >> 1) * 16384 + r * 128;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int chunk = (c & 1) * 4 + q;
        *reinterpret_cast<uint4*>(prow + ((chunk ^ (r & 7)) << 4)) =
            make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
      }
    }
    mine->y = sum;
    fence_proxy_async_smem();
    tc_fence_before_sync();  // this thread's TMEM reads precede the output MMA that overwrites the columns
    mbar_arrive(bar_p);

    mbar_wait(bar_o, 0);
    __syncwarp();
    tc_fence_after_sync();
    named_bar_sync(pair_bar, 64);  // partner's partial row sum is visible
    const float inv = 1.f / (sum + theirs->y);
    {
      uint32_t v[32];
      tmem_ld32(taddr + hc * 32, v);
      tmem_ld_wait();
      uint32_t o[16];
#pragma unroll
      for (int j = 0; j < 32; j += 2)
        o[j >> 1] = pack_bf16x2(__uint_as_float(v[j]) * inv, __uint_as_float(v[j + 1]) * inv);
      uint4* dst = reinterpret_cast<uint4*>(orow + hc * 32);
#pragma unroll
      for (int q = 0; q < 4; ++q) dst[q] = make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after_sync();
    tmem_dealloc<256>(tmem_base);
  }
}

template <int DH>
cudaError_t launch_dh(const CUtensorMap& tm_q, const CUtensorMap& tm_kv, const EncAttnParams& p,
                      cudaStream_t stream) {
  using L = AttnSmem<DH>;
  auto kfn = enc_attention_kernel<DH>;
  {
    cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void*>(kfn), L::kBytes);
    if (e != cudaSuccess) return e;
  }
  kfn<<<p.B * p.nhead, kAttnThreads, L::kBytes, stream>>>(tm_q, tm_kv, p);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_enc_attention(const CUtensorMap& tm_q, const CUtensorMap& tm_kv, const __nv_bfloat16* /*qkv*/,
                                 const EncAttnParams& p, cudaStream_t stream) {
  if (p.B <= 0) return cudaSuccess;
  const int dh = kD / p.nhead;
  if (dh == 64 && !p.one_cta_per_head) {
    cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void*>(enc_attention_tile_kernel), TileSmem::kBytes);
    if (e != cudaSuccess) return e;
    enc_attention_tile_kernel<<<p.B * p.nhead * 2, kTileThreads, TileSmem::kBytes, stream>>>(tm_q, tm_kv, p);
    return cudaGetLastError();
  }
  if (dh == 64) return launch_dh<64>(tm_q, tm_kv, p, stream);
  if (dh == 128) return launch_dh<128>(tm_q, tm_kv, p, stream);
  return cudaErrorInvalidValue;
}

}  // namespace b200vqa
