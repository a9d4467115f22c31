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
// Structure (one CTA per SM, 320 threads):
//   warp 0      : TMA producer   - operand tiles into shared memory (128-byte swizzle)
//   warp 1      : MMA issuer     - tcgen05.mma (M=128, N=BN, K=32 B) accumulating into TMEM; owns the TMEM allocation
//   warps 2..9  : epilogue       - two warps per TMEM lane quarter (four for LayerNorm), each taking every other
//                                  32-column chunk: tcgen05.ld -> fused bias / ReLU / residual+LayerNorm -> bf16 chunk
//                                  transposed through a swizzled shared-memory tile -> coalesced 16-byte stores
// The accumulator is double-buffered in TMEM (2 x BN columns) so the epilogue of tile i overlaps the mainloop
// of tile i+1.
//
// Operand staging has two modes, chosen per launch:
//   W-stationary (K <= 4 k-blocks, i.e. every K=256 layer): the CTA walks a contiguous range of tiles in
//     n-major order, keeps its W[n-tile] (BN x K) resident in shared memory and streams only A through a
//     4-stage ring.  With K=256 a 128x256 tile is just 2048 tensor-core cycles; re-fetching W per tile would
//     need ~180 GB/s per SM from L2, three times what the L2 can give all 148 SMs at once.
//   streaming (K > 256: linear2 with K = ff, image_proj with K = 1024): A and W k-blocks share a 4-stage ring.
#include "kernels.h"
#include "ptx.cuh"

namespace b200vqa {

namespace {

constexpr int kBM = 128;          // rows per tile = UMMA M
constexpr int kKBytes = 128;      // bytes of K per k-block = one 128B swizzle row
constexpr int kUmmaKBytes = 32;   // bytes of K per tcgen05.mma (16 bf16 / 8 tf32)
// epilogue warps: 8 (two per TMEM lane quarter); the LayerNorm epilogue is instruction-latency-bound with two warps per
// scheduler, so it runs 16 (four per quarter, 64 of the 256 columns each)
__host__ __device__ constexpr bool is_ln_epi(int epi) { return epi == kEpiBiasResLN || epi == kEpiBiasResLN2; }
__host__ __device__ constexpr int epi_warps(int epi) { return is_ln_epi(epi) ? 16 : 8; }
__host__ __device__ constexpr int gemm_threads(int epi) { return 64 + epi_warps(epi) * 32; }
constexpr int kMaxEpiWarps = 16;
constexpr int kAccStages = 2;
constexpr int kStages = 4;
constexpr int kWstatMaxKb = 4;    // W-stationary when the whole K fits in 4 k-blocks
constexpr int kWstatKb = 4;

template <int BN, bool LN = false>
struct GemmSmem {
  static constexpr int kStageA = kBM * kKBytes;             // 16 KB
  static constexpr int kStageB = BN * kKBytes;              // 8 / 16 / 32 KB
  // BN = 64 serves the decode chain (M = one branch of questions, K = 256, a single tile per CTA): two A stages and a
  // single-buffered epilogue staging bring the CTA from 130 KB to 82 KB, so that two of them - or one beside the
  // memory-attention CTAs of another branch - share an SM instead of each holding one
  // The LayerNorm epilogue (BN = 256) runs three stages: the 16 KB it gives up hold bias | gamma | beta (| gamma2 |
  // beta2) of the row - with ~200 KB of the SM carved out as shared memory there is next to no L1, and every table
  // look-up in the epilogue's chunk loops was an L2 round trip
  static constexpr int kNSt = BN == 64 ? 2 : (LN ? 3 : kStages);
  static constexpr int kStg = BN == 64 ? 16384 : 32768;
  static constexpr int kRingW = kWstatKb * kStageB + kNSt * kStageA;      // W-stationary: W (4 blocks) + A ring
  static constexpr int kRingS = kNSt * (kStageA + kStageB);               // streaming: stage = [A | W]
  static constexpr int kRing = kRingW > kRingS ? kRingW : kRingS;
  // staging: 8 warps x two 32-row x 64-byte tiles (double-buffered), 16 warps x one tile, or 8 x one tile (BN = 64)
  static constexpr int kOffStg = kRing;
  static constexpr int kOffXch = kOffStg + kStg;            // LayerNorm / argmax exchange between the warps of a quarter
  static constexpr int kOffBar = kOffXch + 2048;
  static constexpr int kOffTab = kOffBar + 128 /*barriers + tmem ptr*/;
  static constexpr int kBytes = kOffTab + (LN ? 5 * BN * 4 : 0);
};

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// Writes one 32-row x 32-column bf16 chunk (row per lane in registers) to global memory with coalesced stores:
// transpose through a 2 KB shared tile, then every store instruction covers 8 rows x 64 contiguous bytes.  Plain
// (generic-proxy) stores: a shared->global TMA store would need a fence.proxy.async per chunk (~450 cycles measured).
__device__ __forceinline__ void store_chunk_bf16(uint8_t* buf, const uint32_t (&o)[16], __nv_bfloat16* out, int ldc,
                                                 int row0, int col0, int M, int lane) {
  // 16-byte units of a 64-byte row are XOR-swizzled with (row >> 1) & 3: both the row-per-lane writes and the
  // 8-rows-per-instruction reads are then bank-conflict free (un-swizzled, the writes are 4-way conflicted and the
  // 8 epilogue warps saturate the shared-memory pipe: ~380 cycles per chunk measured)
#pragma unroll
  for (int q = 0; q < 4; ++q)
    *reinterpret_cast<uint4*>(buf + lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4)) =
        make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
  __syncwarp();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int rr = (lane >> 2) + 8 * k;
    const uint4 val = *reinterpret_cast<const uint4*>(buf + rr * 64 + (((lane & 3) ^ ((rr >> 1) & 3)) << 4));
    if (row0 + rr < M) *reinterpret_cast<uint4*>(out + size_t(row0 + rr) * ldc + col0 + (lane & 3) * 8) = val;
  }
  __syncwarp();
}

template <int BN, int EPI, bool TF32>
__global__ void __launch_bounds__(gemm_threads(EPI), BN == 64 ? 2 : 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w, const GemmParams p) {
  using L = GemmSmem<BN, is_ln_epi(EPI)>;
  constexpr uint32_t kTmemCols = kAccStages * BN;
  static_assert(kTmemCols <= 512 && kTmemCols >= 32, "TMEM budget");
  static_assert(!is_ln_epi(EPI) || BN == 256, "LN epilogue needs the whole row in one tile");
  constexpr int kChunks = BN / 32;            // 32-column chunks per tile
  constexpr int kEpiWarps = epi_warps(EPI);
  constexpr int kSplit = kEpiWarps / 4;           // warps sharing a TMEM lane quarter (they split the columns)
  constexpr int kMyChunks = kChunks / kSplit;     // per epilogue warp
  constexpr int kNSt = L::kNSt;
  constexpr int kStgPerWarp = L::kStg / kEpiWarps;  // 4 KB (two buffers) or 2 KB (one buffer)
  constexpr uint32_t kStgMask = kStgPerWarp == 4096 ? 1u : 0u;
  static_assert(kChunks % kSplit == 0, "column chunks must divide among the warps of a quarter");

  // 128-byte swizzled operand tiles need 1024-byte alignment; the whole 227 KB budget is in use, so there is no
  // slack to round up - the alignment attribute is relied upon and checked
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kOffBar);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* acc_full = empty_bar + kStages;
  uint64_t* acc_empty = acc_full + kAccStages;
  uint64_t* w_full = acc_empty + kAccStages;
  uint64_t* w_empty = w_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_empty + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // optional clock64() stamps of CTA 0's pipeline stages (tools/microbench_gemm.py); a null pointer costs one branch
  const bool dbg = p.dbg_clk != nullptr && blockIdx.x == 0;
#define B200VQA_STAMP(i) \
  if (dbg) p.dbg_clk[i] = clock64()
  if (threadIdx.x == 0) B200VQA_STAMP(0);

  constexpr int elem_bytes = TF32 ? 4 : 2;
  constexpr int bk = kKBytes / elem_bytes;           // elements of K per k-block
  const int num_kb = (p.K + bk - 1) / bk;
  const int tiles_m = (p.M + kBM - 1) / kBM;
  const int tiles_n = p.N / BN;
  const int num_tiles = tiles_m * tiles_n;
  const bool wstat = num_kb <= kWstatMaxKb;
  // contiguous, balanced tile range of this CTA in n-major order (tile L -> n = L / tiles_m, m = L % tiles_m)
  const int t_begin = int((long long)num_tiles * blockIdx.x / gridDim.x);
  const int t_end = int((long long)num_tiles * (blockIdx.x + 1) / gridDim.x);

  // W-stationary layout: [W block 0..3 | A ring]; streaming layout: stage s = [A | W]
  uint8_t* sW = smem;
  uint8_t* sAring = smem + kWstatMaxKb * L::kStageB;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_w);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < kAccStages; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], kEpiWarps);  // one arrive per epilogue warp
    }
    mbar_init(w_full, 1);
    mbar_init(w_empty, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
  pdl_launch_dependents();  // the next kernel of the stream may start its own prologue now
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) B200VQA_STAMP(1);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    // (whole warp walks the loop, one elected lane issues: UTMALDG takes its operands from uniform registers too, and
    // under `if (lane == 0)` every load costs an elect / R2UR.BROADCAST loop of ~100 cycles)
    {
      const bool el = elect_one();
      int stage = 0;
      uint32_t phase = 0;
      int cur_n = -1;
      uint32_t wphase = 0;
      if (wstat && t_begin < t_end) {
        // weights do not depend on the previous kernel: fetch the first W tile before waiting for it
        const int nt = t_begin / tiles_m;
        if (el) {
          mbar_expect_tx(w_full, uint32_t(num_kb) * L::kStageB);
          for (int kb = 0; kb < num_kb; ++kb) tma_load_2d(&tm_w, w_full, sW + kb * L::kStageB, kb * bk, nt * BN);
        }
        __syncwarp();
        cur_n = nt;
        wphase = 1;
      }
      pdl_wait();  // A (and everything the epilogue reads / overwrites) belongs to the previous kernel until here
      if (el) B200VQA_STAMP(2);
      for (int tile = t_begin; tile < t_end; ++tile) {
        const int nt = tile / tiles_m;
        const int m0 = (tile - nt * tiles_m) * kBM;
        const int n0 = nt * BN;
        const int a_koff = p.a_group_cols > 0 ? (n0 / p.a_group_cols) * p.K : 0;  // grouped GEMM: A columns of this head
        if (wstat && nt != cur_n) {
          mbar_wait(w_empty, wphase ^ 1);  // every MMA that read the previous W has retired
          if (el) {
            mbar_expect_tx(w_full, uint32_t(num_kb) * L::kStageB);
            for (int kb = 0; kb < num_kb; ++kb) tma_load_2d(&tm_w, w_full, sW + kb * L::kStageB, kb * bk, n0);
          }
          __syncwarp();
          cur_n = nt;
          wphase ^= 1;
        }
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (el) {
            if (wstat) {
              mbar_expect_tx(&full_bar[stage], L::kStageA);
              tma_load_2d(&tm_a, &full_bar[stage], sAring + stage * L::kStageA, kb * bk + a_koff, m0);
            } else {
              uint8_t* sa = smem + stage * (L::kStageA + L::kStageB);
              mbar_expect_tx(&full_bar[stage], L::kStageA + L::kStageB);
              tma_load_2d(&tm_a, &full_bar[stage], sa, kb * bk + a_koff, m0);
              tma_load_2d(&tm_w, &full_bar[stage], sa + L::kStageA, kb * bk, n0);
            }
          }
          __syncwarp();
          if (++stage == kNSt) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // The whole warp walks the loop (waits included) and one elected lane issues: with warp-uniform control flow the
    // descriptors are computed in uniform registers.  Under `if (lane == 0)` every tcgen05.mma is wrapped in an
    // elect / R2UR.BROADCAST loop - ~16 instructions and ~100 cycles per MMA, as long as a BN = 256 MMA itself and three
    // times a BN = 64 one (the decode chain's GEMMs).
    {
      const bool el = elect_one();
      const uint32_t idesc = make_idesc(TF32 ? kFmtTF32 : (p.f16 ? kFmtF16 : kFmtBF16), kBM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      int cur_n = -1;
      uint32_t wphase = 0;
      for (int tile = t_begin; tile < t_end; ++tile) {
        const int nt = tile / tiles_m;
        if (wstat && nt != cur_n) {
          mbar_wait(w_full, wphase);
          wphase ^= 1;
          cur_n = nt;
        }
        mbar_wait(&acc_empty[as], aphase ^ 1);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after_sync();
          if (kb == 0 && el) B200VQA_STAMP(4);
          if (kb == num_kb - 1 && el) B200VQA_STAMP(5);
          uint32_t sa, sb;
          if (wstat) {
            sa = smem_u32(sAring + stage * L::kStageA);
            sb = smem_u32(sW + kb * L::kStageB);
          } else {
            sa = smem_u32(smem + stage * (L::kStageA + L::kStageB));
            sb = sa + L::kStageA;
          }
          // descriptor = base + (byte offset >> 4): tiles are 1024-byte aligned, the 14-bit address field never carries
          const uint64_t da0 = make_smem_desc_sw128(sa, 16, 1024), db0 = make_smem_desc_sw128(sb, 16, 1024);
          if (el) {
#pragma unroll
            for (int k = 0; k < kKBytes / kUmmaKBytes; ++k) {
              const uint64_t da = da0 + uint64_t((k * kUmmaKBytes) >> 4);
              const uint64_t db = db0 + uint64_t((k * kUmmaKBytes) >> 4);
              if (TF32) umma_tf32(d_tmem, da, db, idesc, (kb | k) != 0);
              else      umma_bf16(d_tmem, da, db, idesc, (kb | k) != 0);
            }
            umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
          }
          __syncwarp();
          if (++stage == kNSt) { stage = 0; phase ^= 1; }
        }
        if (el) umma_commit(&acc_full[as]);        // accumulator ready for the epilogue
        if (wstat) {
          const bool last_with_w = (tile + 1 == t_end) || ((tile + 1) / tiles_m != nt);
          if (last_with_w && el) umma_commit(w_empty);
        }
        __syncwarp();
        if (++as == kAccStages) { as = 0; aphase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (8 warps)
    const int ew = warp - 2;
    const int quarter = warp & 3;  // TMEM lanes [32*quarter, 32*quarter+32) are the ones this warp may read
    const int half = ew >> 2;      // which of the kSplit warps of this quarter: takes chunks half, half+kSplit, ...
    const int row_in_tile = quarter * 32 + lane;
    uint8_t* stg = smem + L::kOffStg + ew * kStgPerWarp;
    float2* xch = reinterpret_cast<float2*>(smem + L::kOffXch);  // [8][32] float2 (8-warp epilogues)
    float* xchf = reinterpret_cast<float*>(smem + L::kOffXch);   // [16][32] float  (LayerNorm epilogue)
    int as = 0;
    uint32_t aphase = 0;
    uint32_t nstore = 0;  // chunks written by this warp (staging buffer = nstore & kStgMask)
    [[maybe_unused]] float* btab = reinterpret_cast<float*>(smem + L::kOffXch);  // bias of the current n-tile
    [[maybe_unused]] int btab_nt = -1;
    [[maybe_unused]] const float* ltab = reinterpret_cast<const float*>(smem + L::kOffTab);
    [[maybe_unused]] uint4 rnext[4];  // LayerNorm epilogue: the residual chunk requested ahead
    if constexpr (is_ln_epi(EPI)) {
      // bias | gamma | beta (| gamma2 | beta2) of the 256 columns (N = BN: one n-tile) into shared memory, once per CTA
      // (weights: safe before the dependency wait)
      float* wtab = reinterpret_cast<float*>(smem + L::kOffTab);
      for (int i = ew * 32 + lane; i < BN; i += kEpiWarps * 32) {
        wtab[i] = p.bias ? __ldg(p.bias + i) : 0.f;
        wtab[BN + i] = __ldg(p.gamma + i);
        wtab[2 * BN + i] = __ldg(p.beta + i);
        if constexpr (EPI == kEpiBiasResLN2) {
          wtab[3 * BN + i] = __ldg(p.gamma2 + i);
          wtab[4 * BN + i] = __ldg(p.beta2 + i);
        }
      }
      named_bar_sync(6, kEpiWarps * 32);
    }
    pdl_wait();

    for (int tile = t_begin; tile < t_end; ++tile) {
      const int nt = tile / tiles_m;
      const int m0 = (tile - nt * tiles_m) * kBM;
      const int n0 = nt * BN;
      const int row = m0 + row_in_tile;
      const bool valid = row < p.M;
      if constexpr (EPI == kEpiBias || EPI == kEpiBiasRelu) {
        // bias of this n-tile into shared memory (once per n-tile, while the operands are still in flight): the
        // chunk loops then read it with broadcast LDS instead of paying an L2 round trip per chunk
        if (nt != btab_nt) {
          named_bar_sync(6, kEpiWarps * 32);  // everyone is done with the previous n-tile's table
          for (int i = ew * 32 + lane; i < BN; i += kEpiWarps * 32) btab[i] = p.bias ? __ldg(p.bias + n0 + i) : 0.f;
          named_bar_sync(6, kEpiWarps * 32);
          btab_nt = nt;
        }
      }
      // LayerNorm epilogue: the first residual chunk of a tile is requested while the previous tile is still being
      // normalised and stored (below), so no HBM round trip sits in front of the dependent chunk loops
      if constexpr (is_ln_epi(EPI)) {
        if (tile == t_begin) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int grow = m0 + quarter * 32 + (lane >> 2) + 8 * k;
            rnext[k] = grow < p.M ? *reinterpret_cast<const uint4*>(p.residual + size_t(grow) * p.ldr + n0 + half * 32 +
                                                                      (lane & 3) * 8)
                                  : make_uint4(0, 0, 0, 0);
          }
        }
      }
      mbar_wait(&acc_full[as], aphase);
      __syncwarp();
      tc_fence_after_sync();
      if (ew == 0 && lane == 0) B200VQA_STAMP(6);
      const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(as * BN);

      if constexpr (EPI == kEpiBias || EPI == kEpiBiasRelu) {
        // software-pipelined: the TMEM load of chunk i+1 is in flight while chunk i is converted and stored
        uint32_t rbuf[2][32];
        tmem_ld32(taddr + half * 32, rbuf[0]);
#pragma unroll
        for (int i = 0; i < kMyChunks; ++i) {
          const int c = half + kSplit * i;
          uint32_t(&r)[32] = rbuf[i & 1];
          tmem_ld_wait();
          if (i + 1 < kMyChunks) {
            tmem_ld32(taddr + (c + kSplit) * 32, rbuf[(i + 1) & 1]);
          } else {  // all TMEM reads of this stage are done -> hand it back to the MMA warp
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[as]);
          }
          uint32_t o[16];
          const float* bptr = btab + c * 32;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b4 = *reinterpret_cast<const float4*>(bptr + j);
            float v0 = __uint_as_float(r[j]) + b4.x, v1 = __uint_as_float(r[j + 1]) + b4.y;
            float v2 = __uint_as_float(r[j + 2]) + b4.z, v3 = __uint_as_float(r[j + 3]) + b4.w;
            if (EPI == kEpiBiasRelu) {
              v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f);
            }
            o[j >> 1] = pack_bf16x2(v0, v1);
            o[(j >> 1) + 1] = pack_bf16x2(v2, v3);
          }
          store_chunk_bf16(stg + (nstore & kStgMask) * 2048, o, p.out, p.ldc, m0 + quarter * 32, n0 + c * 32, p.M, lane);
          ++nstore;
        }
      } else if constexpr (EPI == kEpiBiasPeRemap) {
        const int item = row / p.rows_in;
        const int pos = row - item * p.rows_in;
        __nv_bfloat16* orow =
            p.out + (size_t(valid ? item : 0) * p.rows_out + p.row_off + (valid ? pos : 0)) * p.ldc + n0;
        const float* perow = p.pe + size_t(p.pe_off + (valid ? pos : 0)) * p.N + n0;
#pragma unroll 1
        for (int i = 0; i < kMyChunks; ++i) {
          const int c = half + kSplit * i;
          uint32_t r[32];
          tmem_ld32(taddr + c * 32, r);
          tmem_ld_wait();
          if (i == kMyChunks - 1) {
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[as]);
          }
          uint32_t o[16];
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 pe4 = ldg4(perow + c * 32 + j);
            const float4 b4 = ldg4(p.bias + n0 + c * 32 + j);
            o[j >> 1] = pack_bf16x2(__uint_as_float(r[j]) + b4.x + pe4.x, __uint_as_float(r[j + 1]) + b4.y + pe4.y);
            o[(j >> 1) + 1] =
                pack_bf16x2(__uint_as_float(r[j + 2]) + b4.z + pe4.z, __uint_as_float(r[j + 3]) + b4.w + pe4.w);
          }
          if (valid) {
            uint4* dst = reinterpret_cast<uint4*>(orow + c * 32);
#pragma unroll
            for (int q = 0; q < 4; ++q) dst[q] = make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
          }
        }
      } else if constexpr (EPI == kEpiLstm) {
        // chunks half, half+2, half+4, half+6 of this tile = gates i, f, g, o of 32 hidden units of this row
        const int hidden = p.N >> 2;
        float v[kMyChunks][32];
        long long tokv = p.lstm_tokens ? p.lstm_tokens[size_t(valid ? row : 0) * p.lstm_tok_ld + p.lstm_tok_col]
                                       : (long long)p.lstm_token_const;
        tokv = tokv < 0 ? 0 : (tokv >= p.lstm_vocab ? p.lstm_vocab - 1 : tokv);
        const float* trow = p.lstm_table + size_t(tokv) * p.N + n0;
#pragma unroll
        for (int i = 0; i < kMyChunks; ++i) {
          const int c = half + kSplit * i;
          uint32_t r[32];
          tmem_ld32(taddr + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 t4 = ldg4(trow + c * 32 + j);
            v[i][j] = __uint_as_float(r[j]) + t4.x;
            v[i][j + 1] = __uint_as_float(r[j + 1]) + t4.y;
            v[i][j + 2] = __uint_as_float(r[j + 2]) + t4.z;
            v[i][j + 3] = __uint_as_float(r[j + 3]) + t4.w;
          }
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[as]);
        if (valid) {
          const int u0 = (n0 >> 2) + half * 32;  // first hidden unit of this thread
          float* crow = p.lstm_c + size_t(row) * hidden + u0;
          __nv_bfloat16* hrow = p.lstm_h + size_t(row) * hidden + u0;
          float* hfrow = p.lstm_h_f32 ? p.lstm_h_f32 + size_t(row) * hidden + u0 : nullptr;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            float cprev[8], hn[8];
            *reinterpret_cast<float4*>(cprev) = *reinterpret_cast<const float4*>(crow + j);
            *reinterpret_cast<float4*>(cprev + 4) = *reinterpret_cast<const float4*>(crow + j + 4);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float ig = 1.f / (1.f + expf(-v[0][j + e]));
              const float fg = 1.f / (1.f + expf(-v[1][j + e]));
              const float gg = tanhf(v[2][j + e]);
              const float og = 1.f / (1.f + expf(-v[3][j + e]));
              const float cn = fg * cprev[e] + ig * gg;
              cprev[e] = cn;
              hn[e] = og * tanhf(cn);
            }
            *reinterpret_cast<float4*>(crow + j) = *reinterpret_cast<const float4*>(cprev);
            *reinterpret_cast<float4*>(crow + j + 4) = *reinterpret_cast<const float4*>(cprev + 4);
            *reinterpret_cast<uint4*>(hrow + j) = make_uint4(pack_bf16x2(hn[0], hn[1]), pack_bf16x2(hn[2], hn[3]),
                                                             pack_bf16x2(hn[4], hn[5]), pack_bf16x2(hn[6], hn[7]));
            if (hfrow) {
              *reinterpret_cast<float4*>(hfrow + j) = *reinterpret_cast<const float4*>(hn);
              *reinterpret_cast<float4*>(hfrow + j + 4) = *reinterpret_cast<const float4*>(hn + 4);
            }
          }
        }
      } else if constexpr (EPI == kEpiHead) {
        // logits of this warp's columns, running argmax in increasing column order (first maximum wins)
        float best = -INFINITY;
        int besti = 0x7fffffff;
        float* lrow = (p.logits && valid) ? p.logits + (size_t(row) * p.logits_T + p.head_t) * p.head_V : nullptr;
#pragma unroll 1
        for (int i = 0; i < kMyChunks; ++i) {
          const int c = half + kSplit * i;
          if (n0 + c * 32 >= p.head_V) break;
          uint32_t r[32];
          tmem_ld32(taddr + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int col = n0 + c * 32 + j;
            if (col < p.head_V) {
              const float lg = __uint_as_float(r[j]) + __ldg(p.bias + col);
              if (lrow) lrow[col] = lg;
              if (lg > best) { best = lg; besti = col; }
            }
          }
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[as]);
        // combine with the partner warp (the other half of the columns of the same rows)
        const int partner = ew ^ 4;
        xch[ew * 32 + lane] = make_float2(best, __int_as_float(besti));
        named_bar_sync(1 + quarter, 64);
        const float2 other = xch[partner * 32 + lane];
        const int oi = __float_as_int(other.y);
        if (other.x > best || (other.x == best && oi < besti)) { best = other.x; besti = oi; }
        named_bar_sync(1 + quarter, 64);  // xch is reused by the next tile
        if (besti >= p.head_V) besti = 0;  // all-NaN row: keep the index in range
        if (valid && half == 0) p.tok[size_t(row) * p.tok_ld + p.head_t + 1] = besti;
        if (p.pe_next) {
          // next decoder input x_next[row] = emb[next token] + pe[t+1].  The warp walks its 32 rows together: one
          // row per iteration, 4 columns per lane of this warp's half [128*half, 128*half + 128) - coalesced 512-byte
          // embedding reads and 256-byte stores (a row per lane would touch 32 different lines per instruction)
          long long nxt = (valid && p.forced) ? p.forced[size_t(row) * p.forced_ld + p.head_t] : (long long)besti;
          nxt = nxt < 0 ? 0 : (nxt >= p.vocab ? p.vocab - 1 : nxt);
          const int nx = int(nxt);
          const int colx = half * 128 + lane * 4;
          const float4 pe4 = ldg4(p.pe_next + colx);
          const int row0 = m0 + quarter * 32;
          const int rows_here = p.M - row0;  // rows of this warp inside the matrix (may be <= 0 or > 32)
#pragma unroll 8
          for (int r = 0; r < 32; ++r) {
            const int tr = __shfl_sync(0xffffffffu, nx, r);
            if (r < rows_here) {
              const float4 e = ldg4(p.emb + size_t(tr) * kD + colx);
              *reinterpret_cast<uint2*>(p.x_next + size_t(row0 + r) * kD + colx) =
                  make_uint2(pack_bf16x2(e.x + pe4.x, e.y + pe4.y), pack_bf16x2(e.z + pe4.z, e.w + pe4.w));
            }
          }
        }
      } else {  // kEpiBiasResLN : this warp holds 32 rows x (BN/2) columns of v = acc + bias + residual in registers
        float v[kMyChunks][32];
#pragma unroll
        for (int i = 0; i < kMyChunks; ++i) {
          const int c = half + kSplit * i;
          uint32_t r[32];
          tmem_ld32(taddr + c * 32, r);
          // residual chunk (32 rows x 64 B): coalesced 16-byte loads (8 rows per instruction), transposed through smem
          uint8_t* buf = stg + (i & kStgMask) * 2048;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            *reinterpret_cast<uint4*>(buf + ((lane >> 2) + 8 * k) * 64 +
                                      (((lane & 3) ^ ((((lane >> 2) + 8 * k) >> 1) & 3)) << 4)) = rnext[k];
          if (i + 1 < kMyChunks) {  // next chunk's residual is in flight while this one is combined
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int grow = m0 + quarter * 32 + (lane >> 2) + 8 * k;
              rnext[k] = grow < p.M ? *reinterpret_cast<const uint4*>(p.residual + size_t(grow) * p.ldr + n0 +
                                                                        (c + kSplit) * 32 + (lane & 3) * 8)
                                    : make_uint4(0, 0, 0, 0);
            }
          } else if (tile + 1 < t_end) {  // ... and the first chunk of the NEXT tile under this tile's statistics + stores
            const int m0n = (tile + 1 - ((tile + 1) / tiles_m) * tiles_m) * kBM;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int grow = m0n + quarter * 32 + (lane >> 2) + 8 * k;
              rnext[k] = grow < p.M ? *reinterpret_cast<const uint4*>(p.residual + size_t(grow) * p.ldr + n0 + half * 32 +
                                                                        (lane & 3) * 8)
                                    : make_uint4(0, 0, 0, 0);
            }
          }
          __syncwarp();
          tmem_ld_wait();
          const float* bptr = ltab + c * 32;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 rv = *reinterpret_cast<const uint4*>(buf + lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4));
            const uint32_t w4[4] = {rv.x, rv.y, rv.z, rv.w};
            const float4 b0 = *reinterpret_cast<const float4*>(bptr + q * 8),
                         b1 = *reinterpret_cast<const float4*>(bptr + q * 8 + 4);
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 res = unpack_bf16x2(w4[e]);
              const int j = q * 8 + e * 2;
              v[i][j] = __uint_as_float(r[j]) + bb[e * 2] + res.x;
              v[i][j + 1] = __uint_as_float(r[j + 1]) + bb[e * 2 + 1] + res.y;
            }
          }
          __syncwarp();  // buffer (i & 1) is rewritten two chunks later
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[as]);

        // row statistics over all BN columns: the kSplit warps of this quarter exchange partial sums (two-pass)
        const int xbase = (ew & 3) * 32 + lane;  // slot of warp (ew & 3) + 4*j is xchf[xbase + 128*j]
        float mean, rstd;
        auto row_stats = [&]() {
          float s = 0.f;
#pragma unroll
          for (int i = 0; i < kMyChunks; ++i)
#pragma unroll
            for (int j = 0; j < 32; ++j) s += v[i][j];
          xchf[ew * 32 + lane] = s;
          named_bar_sync(1 + quarter, kSplit * 32);
          float tot = 0.f;
#pragma unroll
          for (int j = 0; j < kSplit; ++j) tot += xchf[xbase + 128 * j];
          mean = tot * (1.f / BN);
          float sq = 0.f;
#pragma unroll
          for (int i = 0; i < kMyChunks; ++i)
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float d = v[i][j] - mean;
              sq = fmaf(d, d, sq);
            }
          named_bar_sync(1 + quarter, kSplit * 32);  // every partner has read the sums before they are overwritten
          xchf[ew * 32 + lane] = sq;
          named_bar_sync(1 + quarter, kSplit * 32);
          float tsq = 0.f;
#pragma unroll
          for (int j = 0; j < kSplit; ++j) tsq += xchf[xbase + 128 * j];
          rstd = rsqrtf(tsq * (1.f / BN) + p.eps);
          named_bar_sync(1 + quarter, kSplit * 32);  // ... and the squares before the next sums
        };
        row_stats();
        const float* gamma = ltab + BN;
        const float* beta = ltab + 2 * BN;
        if constexpr (EPI == kEpiBiasResLN2) {
          // a second LayerNorm on top (nn.Transformer's final encoder norm, FA:42): normalise in registers, then take the
          // statistics of the result - one kernel and one 512-byte-per-row round trip through HBM fewer
#pragma unroll
          for (int i = 0; i < kMyChunks; ++i) {
            const int c = half + kSplit * i;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 g4 = *reinterpret_cast<const float4*>(gamma + c * 32 + j),
                           t4 = *reinterpret_cast<const float4*>(beta + c * 32 + j);
              v[i][j] = (v[i][j] - mean) * rstd * g4.x + t4.x;
              v[i][j + 1] = (v[i][j + 1] - mean) * rstd * g4.y + t4.y;
              v[i][j + 2] = (v[i][j + 2] - mean) * rstd * g4.z + t4.z;
              v[i][j + 3] = (v[i][j + 3] - mean) * rstd * g4.w + t4.w;
            }
          }
          row_stats();
          gamma = ltab + 3 * BN;
          beta = ltab + 4 * BN;
        }

        float* frow = (p.out_f32 && valid) ? p.out_f32 + size_t(row) * BN : nullptr;
#pragma unroll
        for (int i = 0; i < kMyChunks; ++i) {
          const int c = half + kSplit * i;
          const float* gptr = gamma + c * 32;
          const float* btptr = beta + c * 32;
          uint32_t o[16];
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 g4 = *reinterpret_cast<const float4*>(gptr + j), t4 = *reinterpret_cast<const float4*>(btptr + j);
            const float y0 = (v[i][j] - mean) * rstd * g4.x + t4.x;
            const float y1 = (v[i][j + 1] - mean) * rstd * g4.y + t4.y;
            const float y2 = (v[i][j + 2] - mean) * rstd * g4.z + t4.z;
            const float y3 = (v[i][j + 3] - mean) * rstd * g4.w + t4.w;
            o[j >> 1] = pack_bf16x2(y0, y1);
            o[(j >> 1) + 1] = pack_bf16x2(y2, y3);
            if (frow) *reinterpret_cast<float4*>(frow + c * 32 + j) = make_float4(y0, y1, y2, y3);
          }
          store_chunk_bf16(stg + (nstore & kStgMask) * 2048, o, p.out, p.ldc, m0 + quarter * 32, n0 + c * 32, p.M, lane);
          ++nstore;
        }
      }
      if (++as == kAccStages) { as = 0; aphase ^= 1; }
    }
    if (ew == 0 && lane == 0) B200VQA_STAMP(7);
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after_sync();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
  if (threadIdx.x == 0) B200VQA_STAMP(9);
#undef B200VQA_STAMP
}

template <int BN, int EPI, bool TF32>
cudaError_t launch_one(const CUtensorMap& tm_a, const CUtensorMap& tm_w, const GemmParams& p, int num_sms,
                       cudaStream_t stream) {
  using L = GemmSmem<BN, is_ln_epi(EPI)>;
  auto kfn = gemm_tc_kernel<BN, EPI, TF32>;
  {
    cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void*>(kfn), L::kBytes);
    if (e != cudaSuccess) return e;
  }
  const int tiles = ((p.M + kBM - 1) / kBM) * (p.N / BN);
  if (tiles <= 0) return cudaSuccess;
  const int grid = tiles < num_sms ? tiles : num_sms;
  return launch_kernel(kfn, dim3(grid), dim3(gemm_threads(EPI)), L::kBytes, stream, p.pdl, tm_a, tm_w, p);
}

// ------------------------------------------------------------------------------------------
// Latency-optimised out_proj + residual + LayerNorm for the decode chain (M = questions of one branch, N = K = 256).
//
// The persistent kernel above gives a 128-row m-tile to ONE SM, whose epilogue then owns 128 x 256 values: at decode
// sizes (1..8 m-tiles) that epilogue IS the kernel (12 k of 20 k cycles measured).  Here a cluster of 4 CTAs shares the
// m-tile, each computing a 128 x 64 column slice (A tile 64 KB + W slice 32 KB per SM instead of 64 + 128 KB).  The
// LayerNorm statistics are per-thread (mean, M2) pairs over 32 columns, sent to every peer's shared memory with
// st.async (the arriving bytes complete the receiver's mbarrier: no cluster-wide barrier on the critical path) and
// combined with the parallel-variance formula (exact two-pass statistics per part, no E[x^2] - mean^2 cancellation).
//   warp 0     : TMA loads + tcgen05.mma issue (one thread), owns the 64 TMEM columns
//   warps 1..8 : epilogue; warp w reads TMEM lane quarter (w & 3), columns 32 * ((w - 1) >> 2) of the slice
// ------------------------------------------------------------------------------------------
constexpr int kLnCl = 4;                      // CTAs per cluster = column slices of the 256-wide row
constexpr int kLnBN = 256 / kLnCl;            // 64 columns per CTA
constexpr int kLnMaxStages = 8;               // k-block stages of shared memory (A 16 KB + W 8 KB each)
constexpr int kLnThreads = 32 + 8 * 32;
// NKB = K / 64 k-blocks of 128 B: 4 (K = 256, every k-block has its own stage), 8 (K = 512) or 16 (K = 1024: the fused
// value + output projection of the absorbed cross-attention; k-blocks 8..15 refill the stages as their MMAs retire)
template <int NKB>
struct LnClSmem {
  static constexpr int kStages = NKB < kLnMaxStages ? NKB : kLnMaxStages;
  static constexpr int kOffA = 0;                                   // kStages x [128 rows x 128 B]
  static constexpr int kOffW = kOffA + kStages * kBM * kKBytes;     // kStages x [64 rows x 128 B]
  static constexpr int kOffStg = kOffW + kStages * kLnBN * kKBytes; // 8 warps x 2 KB
  static constexpr int kOffPart = kOffStg + 8 * 2048;               // float2 [2 * kLnCl parts][128 rows]
  static constexpr int kOffTab = kOffPart + 2 * kLnCl * kBM * 8;    // bias | gamma | beta of this slice (3 x 64 fp32)
  static constexpr int kOffBar = kOffTab + 3 * kLnBN * 4;
  static constexpr int kBytes = kOffBar + 256;
};

template <int NKB>
__global__ void __cluster_dims__(kLnCl, 1, 1) __launch_bounds__(kLnThreads, 1)
gemm_ln_cluster_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w,
                       const GemmParams p) {
  using L = LnClSmem<NKB>;
  constexpr int NS = L::kStages;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kOffBar);  // [NS] A k-blocks (refills: A + W)
  uint64_t* w_full = full_bar + NS;                                     // [NS] W k-blocks of the first pass
  uint64_t* empty_bar = w_full + NS;                                    // [NS] the stage's MMAs have retired
  uint64_t* acc_full = empty_bar + NS;
  uint64_t* stat_full = acc_full + 1;  // completes when all 2 * kLnCl statistic parts of the 128 rows have arrived
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(stat_full + 1);
  float* tab = reinterpret_cast<float*>(smem + L::kOffTab);
  float2* part = reinterpret_cast<float2*>(smem + L::kOffPart);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();   // == blockIdx.x : column slice
  const int m0 = blockIdx.y * kBM;
  const int n0 = int(rank) * kLnBN;
  const bool dbg = p.dbg_clk != nullptr && blockIdx.x == 0 && blockIdx.y == 0;
#define B200VQA_STAMP(i) \
  if (dbg) p.dbg_clk[i] = clock64()
  if (threadIdx.x == 0) B200VQA_STAMP(0);

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tm_a);
      tma_prefetch_desc(&tm_w);
      for (int kb = 0; kb < NS; ++kb) {
        mbar_init(&full_bar[kb], 1);
        mbar_init(&w_full[kb], 1);
        mbar_init(&empty_bar[kb], 1);
      }
      mbar_init(acc_full, 1);
      mbar_init(stat_full, 1);
      fence_mbar_init();
      mbar_expect_tx(stat_full, 2 * kLnCl * kBM * 8);
    }
    __syncwarp();
    // weights do not depend on the previous kernel (issued from uniform registers: see gemm_tc_kernel's producer)
    if (elect_one()) {
      for (int kb = 0; kb < NS; ++kb) {
        mbar_expect_tx(&w_full[kb], kLnBN * kKBytes);
        tma_load_2d(&tm_w, &w_full[kb], smem + L::kOffW + kb * kLnBN * kKBytes, kb * 64, n0);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    tmem_alloc<kLnBN>(tmem_slot);
  }
  // this slice's bias | gamma | beta (weights: safe before the dependency wait).  Only requested here - the value is
  // parked in shared memory after the residual loads are in flight, so no barrier waits for this L2 round trip
  float tabv = 0.f;
  const int tabi = (warp - 1) * 32 + lane;
  if (warp >= 1 && tabi < 3 * kLnBN) {
    const float* src = tabi < kLnBN ? p.bias : (tabi < 2 * kLnBN ? p.gamma : p.beta);
    tabv = src ? __ldg(src + n0 + (tabi & (kLnBN - 1))) : (tabi >= kLnBN && tabi < 2 * kLnBN ? 1.f : 0.f);
  }
  pdl_launch_dependents();
  tc_fence_before_sync();
  __syncthreads();  // this CTA's barriers and TMEM pointer are set
  // cluster-wide "my barriers are initialised, my shared memory may be written": only arrive here, the wait sits right
  // before the first store into a peer, where it has long completed
  cluster_arrive_release();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) B200VQA_STAMP(1);

  if (warp == 0) {
    // producer + MMA issuer.  The whole warp walks the loop and one elected lane issues, so that the descriptors stay in
    // uniform registers (see gemm_tc_kernel): 16 to 64 MMAs of 32 cycles each are issue-bound otherwise.
    const bool el = elect_one();
    pdl_wait();  // A belongs to the previous kernel until here
    if (el) {
      B200VQA_STAMP(2);
      for (int kb = 0; kb < NS; ++kb) {
        mbar_expect_tx(&full_bar[kb], kBM * kKBytes);
        tma_load_2d(&tm_a, &full_bar[kb], smem + L::kOffA + kb * kBM * kKBytes, kb * 64, m0);
      }
    }
    __syncwarp();
    constexpr uint32_t idesc = make_idesc(kFmtBF16, kBM, kLnBN, 0, 0);
    const uint64_t da_base = make_smem_desc_sw128(smem_u32(smem + L::kOffA), 16, 1024);
    const uint64_t db_base = make_smem_desc_sw128(smem_u32(smem + L::kOffW), 16, 1024);
    for (int kb = 0; kb < NKB; ++kb) {
      const int st = kb % NS;
      if (kb < NS) mbar_wait(&w_full[st], 0);
      mbar_wait(&full_bar[st], uint32_t(kb / NS) & 1u);
      tc_fence_after_sync();
      if (el) {
        if (kb == 0) B200VQA_STAMP(4);
        if (kb == NKB - 1) B200VQA_STAMP(5);
        const uint64_t da0 = da_base + uint64_t((st * kBM * kKBytes) >> 4);
        const uint64_t db0 = db_base + uint64_t((st * kLnBN * kKBytes) >> 4);
#pragma unroll
        for (int k = 0; k < kKBytes / kUmmaKBytes; ++k)
          umma_bf16(tmem_base, da0 + uint64_t((k * kUmmaKBytes) >> 4), db0 + uint64_t((k * kUmmaKBytes) >> 4), idesc,
                    (kb | k) != 0);
      }
      __syncwarp();
      if constexpr (NKB > NS) {
        if (kb + NS < NKB && el) umma_commit(&empty_bar[st]);
        // refill the previous k-block's stage (its MMAs retire while this k-block's run): k-block kb - 1 + NS
        if (kb >= 1 && kb - 1 + NS < NKB) {
          const int sp = (kb - 1) % NS, nk = kb - 1 + NS;
          mbar_wait(&empty_bar[sp], uint32_t((kb - 1) / NS) & 1u);
          if (el) {
            mbar_expect_tx(&full_bar[sp], (kBM + kLnBN) * kKBytes);
            tma_load_2d(&tm_a, &full_bar[sp], smem + L::kOffA + sp * kBM * kKBytes, nk * 64, m0);
            tma_load_2d(&tm_w, &full_bar[sp], smem + L::kOffW + sp * kLnBN * kKBytes, nk * 64, n0);
          }
          __syncwarp();
        }
      }
    }
    if (el) umma_commit(acc_full);
    __syncwarp();
    cluster_wait_acquire();
  } else {
    const int ew = warp - 1;
    const int quarter = warp & 3;
    const int half = ew >> 2;
    const int row_in_tile = quarter * 32 + lane;
    const int row = m0 + row_in_tile;
    const bool valid = row < p.M;
    const int col0 = n0 + half * 32;
    uint8_t* buf = smem + L::kOffStg + ew * 2048;
    pdl_wait();  // the residual is the previous kernel's output
    // residual chunk (32 rows x 64 B): coalesced 16-byte loads, transposed to a row per lane through shared memory
    uint4 rv4[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int grow = m0 + quarter * 32 + (lane >> 2) + 8 * k;
      rv4[k] = grow < p.M
                   ? *reinterpret_cast<const uint4*>(p.residual + size_t(grow) * p.ldr + col0 + (lane & 3) * 8)
                   : make_uint4(0, 0, 0, 0);
    }
    if (tabi < 3 * kLnBN) tab[tabi] = tabv;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int rr = (lane >> 2) + 8 * k;
      *reinterpret_cast<uint4*>(buf + rr * 64 + (((lane & 3) ^ ((rr >> 1) & 3)) << 4)) = rv4[k];
    }
    named_bar_sync(1, 8 * 32);  // table complete (all epilogue warps); also orders this warp's staging writes
    float v[32];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint4 rv = *reinterpret_cast<const uint4*>(buf + lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4));
      const uint32_t w4[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 res = unpack_bf16x2(w4[e]);
        v[q * 8 + e * 2] = res.x + tab[half * 32 + q * 8 + e * 2];
        v[q * 8 + e * 2 + 1] = res.y + tab[half * 32 + q * 8 + e * 2 + 1];
      }
    }
    __syncwarp();
    mbar_wait(acc_full, 0);
    __syncwarp();
    tc_fence_after_sync();
    if (ew == 0 && lane == 0) B200VQA_STAMP(6);
    {
      uint32_t r[32];
      tmem_ld32(tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(half * 32), r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] += __uint_as_float(r[j]);
    }
    // exact statistics of this thread's 32 values, shared with every CTA of the cluster
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) s += v[j];
    const float mloc = s * (1.f / 32.f);
    float m2 = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float d = v[j] - mloc;
      m2 = fmaf(d, d, m2);
    }
    const uint32_t slot_addr = smem_u32(part + (int(rank) * 2 + half) * kBM + row_in_tile);
    const uint32_t bar_addr = smem_u32(stat_full);
    __syncwarp();
    cluster_wait_acquire();  // every peer has initialised its barriers (arrived at kernel start)
#pragma unroll
    for (uint32_t dst = 0; dst < kLnCl; ++dst)
      st_async_f32x2(cluster_map_shared(slot_addr, dst), mloc, m2, cluster_map_shared(bar_addr, dst));
    if (ew == 0 && lane == 0) B200VQA_STAMP(10);
    mbar_wait(stat_full, 0);
    if (ew == 0 && lane == 0) B200VQA_STAMP(11);
    float pm[2 * kLnCl], msum = 0.f, m2sum = 0.f;
#pragma unroll
    for (int i = 0; i < 2 * kLnCl; ++i) {
      const float2 pr = part[i * kBM + row_in_tile];
      pm[i] = pr.x;
      msum += pr.x;
      m2sum += pr.y;
    }
    const float mean = msum * (1.f / (2 * kLnCl));
    float dev = 0.f;
#pragma unroll
    for (int i = 0; i < 2 * kLnCl; ++i) dev = fmaf(pm[i] - mean, pm[i] - mean, dev);
    const float rstd = rsqrtf((m2sum + 32.f * dev) * (1.f / 256.f) + p.eps);

    float* frow = (p.out_f32 && valid) ? p.out_f32 + size_t(row) * 256 + col0 : nullptr;
    uint32_t o[16];
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 g4 = *reinterpret_cast<const float4*>(tab + kLnBN + half * 32 + j);
      const float4 t4 = *reinterpret_cast<const float4*>(tab + 2 * kLnBN + half * 32 + j);
      const float y0 = (v[j] - mean) * rstd * g4.x + t4.x;
      const float y1 = (v[j + 1] - mean) * rstd * g4.y + t4.y;
      const float y2 = (v[j + 2] - mean) * rstd * g4.z + t4.z;
      const float y3 = (v[j + 3] - mean) * rstd * g4.w + t4.w;
      o[j >> 1] = pack_bf16x2(y0, y1);
      o[(j >> 1) + 1] = pack_bf16x2(y2, y3);
      if (frow) *reinterpret_cast<float4*>(frow + j) = make_float4(y0, y1, y2, y3);
    }
    store_chunk_bf16(buf, o, p.out, p.ldc, m0 + quarter * 32, col0, p.M, lane);
    if (ew == 0 && lane == 0) B200VQA_STAMP(7);
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc<kLnBN>(tmem_base);
  }
  if (threadIdx.x == 0) B200VQA_STAMP(9);
#undef B200VQA_STAMP
}

template <int NKB>
cudaError_t launch_ln_cluster_k(const CUtensorMap& tm_a, const CUtensorMap& tm_w, const GemmParams& p,
                                cudaStream_t stream) {
  using L = LnClSmem<NKB>;
  {
    cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void*>(gemm_ln_cluster_kernel<NKB>), L::kBytes);
    if (e != cudaSuccess) return e;
  }
  const int tiles_m = (p.M + kBM - 1) / kBM;
  if (tiles_m <= 0) return cudaSuccess;
  return launch_kernel(gemm_ln_cluster_kernel<NKB>, dim3(kLnCl, tiles_m), dim3(kLnThreads), L::kBytes, stream, p.pdl,
                       tm_a, tm_w, p);
}

cudaError_t launch_ln_cluster(const CUtensorMap& tm_a, const CUtensorMap& tm_w, const GemmParams& p,
                              cudaStream_t stream) {
  if (p.K == 256) return launch_ln_cluster_k<4>(tm_a, tm_w, p, stream);
  if (p.K == 512) return launch_ln_cluster_k<8>(tm_a, tm_w, p, stream);
  if (p.K == 1024) return launch_ln_cluster_k<16>(tm_a, tm_w, p, stream);
  return cudaErrorInvalidValue;
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
  if (block_n <= 0 || p.N % block_n != 0) return cudaErrorInvalidValue;
  if (epilogue == kEpiHead && (p.N != block_n || p.head_V > p.N)) return cudaErrorInvalidValue;
  if (epilogue == kEpiLstm && (block_n != 256 || p.N % 256 != 0)) return cudaErrorInvalidValue;
  if (epilogue == kEpiBiasResLN && block_n == kLnBN) {  // cluster-of-4 variant for the decode-sized launches
    if (p.N != 256 || (p.K != 256 && p.K != 512 && p.K != 1024) || tf32) return cudaErrorInvalidValue;
    return launch_ln_cluster(tm_a, tm_w, p, stream);
  }
  if (is_ln_epi(epilogue) && (p.N != 256 || block_n != 256)) return cudaErrorInvalidValue;
  if (epilogue == kEpiBiasResLN2 && (!p.gamma2 || !p.beta2)) return cudaErrorInvalidValue;
#define B200VQA_GEMM_CASE(BN_, EPI_, TF_)                                                   \
  if (block_n == BN_ && epilogue == EPI_ && tf32 == TF_)                                    \
    return launch_one<BN_, EPI_, TF_>(tm_a, tm_w, p, num_sms, stream);
  B200VQA_GEMM_CASE(256, kEpiBias, false)
  B200VQA_GEMM_CASE(128, kEpiBias, false)
  B200VQA_GEMM_CASE(64, kEpiBias, false)
  B200VQA_GEMM_CASE(256, kEpiBiasRelu, false)
  B200VQA_GEMM_CASE(128, kEpiBiasRelu, false)
  B200VQA_GEMM_CASE(64, kEpiBiasRelu, false)
  B200VQA_GEMM_CASE(256, kEpiBiasResLN, false)
  B200VQA_GEMM_CASE(256, kEpiBiasResLN2, false)
  B200VQA_GEMM_CASE(256, kEpiBiasPeRemap, false)
  B200VQA_GEMM_CASE(256, kEpiBiasPeRemap, true)
  B200VQA_GEMM_CASE(256, kEpiLstm, false)
  B200VQA_GEMM_CASE(64, kEpiHead, true)
  B200VQA_GEMM_CASE(256, kEpiHead, true)
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
