// Persistent decode kernel: every position x layer of a greedy decode in ONE launch (kernels.h: DecPersistParams).
//
// Reference semantics: VQAModel.autoregressive_program_generation (IQAP:190-241) and greedy_decode's loop (FA:137-145)
// in the KV-cached, absorbed-cross-attention form the per-kernel chain in api.cu implements (enqueue_decode_rows): per
// position t and decoder layer l
//     qkv = x W_in^T + b                       G1  (tensor cores, columns split over the cluster)
//     a   = self-attention(q, K-cache + k, V-cache + v)          R2  (one warp per question)
//     x1  = LN1(x + a W_o^T + b)               G3 + R4
//     q'  = x1 W_qk^T + b_qk                   G5  (absorbed queries, nhead x 256)
//     u_h = softmax(q'_h . M^T / sqrt(dh)) M   R6  (the HBM-bound stream over the encoder memory, warp-level MMAs)
//     c   = concat_h(u_h W_v,h^T) + b_v        G7
//     x2  = LN2(x1 + c W_co^T + b)             G8 + R9
//     x3  = LN3(x2 + relu(x2 W1^T + b1) W2^T + b2)               G10 + G11 (hidden slice stays in shared memory) + R12
//     last layer: (final norm), logits = x3 W_head^T + b, argmax (first maximum), x_next = emb[tok] + pe[t+1]   R12
//
// Work split.  A cluster of 8 CTAs owns 64 questions from the first position to the last; clusters are independent.
//   G phases: CTA r computes output columns [r N/8, (r+1) N/8) of all 64 questions: tcgen05.mma with M = 64 (accumulator
//             row m lives in TMEM lane (m % 16) + 32 (m / 16)), A = the activation tile copied from the L2-resident
//             scratch into 128-byte-swizzled shared memory by the worker warps, B = weight chunks of up to 64 rows that a
//             producer thread streams from L2 through a 4-slot TMA ring (weights are static: the producer runs ahead
//             across phases and stages), epilogue straight from TMEM to the scratch buffers.
//   R phases: CTA r owns questions 8r..8r+7, worker warp w owns question 8r+w: no intra-CTA synchronisation at all.
//   Between phases the 8 CTAs meet at an mbarrier rendezvous (one remote arrive per CTA pair, release/acquire at
//   cluster scope); the producer and MMA-issue warps never take part, so weight prefetch is not stopped by it.
//   R6 takes the ring's shared memory for the per-warp memory tiles (16 rows x 512 B, two slots per warp, each warp
//   its own TMA producer); the weight producer resumes when the 8 warps have finished (attn_done).
#include <cstdio>
#include <cstdlib>

#include "kernels.h"
#include "ptx.cuh"

namespace b200vqa {
namespace {

constexpr int kDpCluster = 8;
constexpr int kDpTile = 64;                 // questions per cluster = UMMA M
constexpr int kDpWorkers = 8;               // worker warps = questions per CTA
constexpr int kDpThreads = (kDpWorkers + 2) * 32;
constexpr int kDpSlots = 4;
constexpr int kDpSlotBytes = 32768;         // one weight chunk: <= 64 rows x 256 K, or 256 rows x 64 K
constexpr int kDpABytes = 4 * 64 * 128;     // A operand: 4 k-blocks of [64 rows x 128 B]
constexpr int kDpOffA = 0;
constexpr int kDpOffR = kDpOffA + kDpABytes;
constexpr int kDpRBytes = kDpSlots * kDpSlotBytes;
constexpr int kDpOffBar = kDpOffR + kDpRBytes;
constexpr int kDpOffSp = kDpOffBar + 512;   // per worker warp: 256 B softmax-weight scratch
constexpr int kDpSmem = kDpOffSp + kDpWorkers * 256;
constexpr int kDpWarpRing = kDpRBytes / kDpWorkers;  // 16 KB: two 16-row memory tiles per worker warp
static_assert(kDpWarpRing == 2 * 8192, "per-warp memory ring");

__device__ __forceinline__ uint4 ldcg16(const void* p) { return __ldcg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ float4 ldcg_f4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void unpack8(const uint4& u, float (&v)[8]) {
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  return make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}
__device__ __forceinline__ void ldg8(const float* src, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(src));
  const float4 b = __ldg(reinterpret_cast<const float4*>(src) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// TMEM -> registers: this warp's 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// cluster-scope mbarrier pieces of the rendezvous
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded by the clock (about two seconds): a protocol bug traps instead of hanging the GPU box.
constexpr long long kDpTimeoutCycles = 4000000000ll;
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (clock64() - t0 > kDpTimeoutCycles) {
      printf("b200vqa: cluster rendezvous timed out (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void dp_wait(uint64_t* bar, uint32_t parity, int what) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > kDpTimeoutCycles) {
      printf("b200vqa: decode_persist wait %d timed out (block %d,%d thread %d)\n", what, blockIdx.x, blockIdx.y,
             threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void fence_acq_rel_cluster() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }

// byte offset of the 16-byte chunk holding channels [c, c + 8) of row r inside a swizzled 16-row memory tile
// (4 column blocks of [16 rows x 128 B])
__device__ __forceinline__ uint32_t mtile_off(int r, int c) {
  return uint32_t((c >> 6) * 2048 + r * 128 + ((((c & 63) >> 3) ^ (r & 7)) << 4));
}

// LayerNorm of a 256-wide row held 8 values per lane (two-pass statistics, like ffn_reduce_ln_kernel)
__device__ __forceinline__ void warp_layernorm(float (&v)[8], const float* gamma, const float* beta, float eps, int lane) {
  float s1 = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) s1 += v[j];
  const float mean = warp_sum(s1) * (1.f / kD);
  float s2 = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) s2 += (v[j] - mean) * (v[j] - mean);
  const float rstd = rsqrtf(warp_sum(s2) * (1.f / kD) + eps);
  float g[8], t[8];
  ldg8(gamma + lane * 8, g);
  ldg8(beta + lane * 8, t);
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = (v[j] - mean) * rstd * g[j] + t[j];
}

template <int NH>
__global__ void __cluster_dims__(kDpCluster, 1, 1) __launch_bounds__(kDpThreads, 1)
decode_persist_kernel(const __grid_constant__ CUtensorMap tm_wa, const __grid_constant__ CUtensorMap tm_wb,
                      const __grid_constant__ CUtensorMap tm_mem, const DecPersistParams p) {
  static_assert(NH == 2 || NH == 4, "heads");
  constexpr int DH = kD / NH;
  constexpr int NQ = NH * kD / kDpCluster;  // absorbed-query columns per CTA (128 | 64)
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sA = smem + kDpOffA;
  const uint32_t sA_u32 = smem_u32(sA);
  const uint32_t sR_u32 = smem_u32(smem + kDpOffR);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kDpOffBar);
  uint64_t* full = bars;                    // [kDpSlots] weight chunk landed
  uint64_t* empty = bars + kDpSlots;        // [kDpSlots] the chunk's MMAs have retired
  uint64_t* a_full = bars + 2 * kDpSlots;   // A operand written by the 256 worker threads
  uint64_t* acc_full = a_full + 1;          // accumulator complete
  uint64_t* attn_done = a_full + 2;         // the 8 worker warps are done with the ring's shared memory (R6)
  uint64_t* cb_bar = a_full + 3;            // cluster rendezvous: one arrival per CTA
  uint64_t* memfull = a_full + 4;           // [kDpWorkers][2] memory tile landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(memfull + 2 * kDpWorkers);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = int(cluster_ctarank());  // == blockIdx.x: column slice of the G phases, row group of the R phases
  const int tile0 = blockIdx.y * kDpTile;
  const int ff = p.ff, ffs = p.ff / kDpCluster;  // hidden units per CTA (a multiple of 64)
  const int nkb2 = ffs / 64;
  // slab A row offsets inside a layer
  const int off_in = 0, off_out = 3 * kD, off_qk = 4 * kD, off_v = off_qk + NH * kD, off_co = off_v + kD,
            off_w1 = off_co + kD;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_wa);
    tma_prefetch_desc(&tm_wb);
    tma_prefetch_desc(&tm_mem);
    for (int i = 0; i < kDpSlots; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(a_full, kDpWorkers * 32);
    mbar_init(acc_full, 1);
    mbar_init(attn_done, kDpWorkers);
    mbar_init(cb_bar, kDpCluster);
    for (int i = 0; i < 2 * kDpWorkers; ++i) mbar_init(&memfull[i], 1);
    fence_mbar_init();
  }
  if (warp == kDpWorkers + 1) tmem_alloc<256>(tmem_slot);
  // the ring starts with finite contents: a memory tile's stale rows only ever get softmax weight exactly 0
  for (int i = threadIdx.x; i < kDpRBytes / 16; i += kDpThreads)
    reinterpret_cast<uint4*>(smem + kDpOffR)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  // every CTA's barriers exist before any peer arrives on them
  cluster_arrive_release();
  cluster_wait_acquire();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const int n_stage = p.steps * p.n_layers;
  // test hook: with dbg_stop = k >= 1 the workers return after their k-th rendezvous of the first stage (after the 6th
  // for k = 5: the memory tiles in flight are drained first); producer and MMA issue only what the workers consume
  auto dbg_ok = [&](int need) { return p.dbg_stop < 0 || p.dbg_stop >= need; };

  if (warp == kDpWorkers) {
    // ------------------------------------------------------------------ weight producer
    if (lane == 0) {
      uint32_t cc = 0;
      auto slot_wait = [&]() {
        const uint32_t slot = cc % kDpSlots;
        if (cc >= kDpSlots) dp_wait(&empty[slot], ((cc / kDpSlots) - 1) & 1u, 1);
        return slot;
      };
      // rows [row0, row0 + nrows) x 256 K of slab A: per k-block [nrows x 128 B], 32-row boxes
      auto put_a = [&](int row0, int nrows) {
        const uint32_t slot = slot_wait();
        const uint32_t dst = sR_u32 + slot * kDpSlotBytes;
        mbar_expect_tx(&full[slot], uint32_t(nrows) * 512u);
        for (int kb = 0; kb < 4; ++kb)
          for (int rb = 0; rb < nrows; rb += 32)
            tma_load_2d_u32(&tm_wa, &full[slot], dst + kb * nrows * 128 + rb * 128, kb * 64, row0 + rb);
        ++cc;
      };
      // one k-block of linear2: 256 rows x 64 K of slab B
      auto put_b = [&](int row0) {
        const uint32_t slot = slot_wait();
        const uint32_t dst = sR_u32 + slot * kDpSlotBytes;
        mbar_expect_tx(&full[slot], 256u * 128u);
        for (int rb = 0; rb < 256; rb += 32) tma_load_2d_u32(&tm_wb, &full[slot], dst + rb * 128, 0, row0 + rb);
        ++cc;
      };
      for (int st = 0; st < n_stage; ++st) {
        const int l = st % p.n_layers;
        const int base = l * p.rows_per_layer;
        put_a(base + off_in + 96 * r, 64);
        put_a(base + off_in + 96 * r + 64, 32);
        if (!dbg_ok(3)) break;
        put_a(base + off_out + 32 * r, 32);
        if (!dbg_ok(5)) break;
        for (int c = 0; c < NQ; c += 64) put_a(base + off_qk + NQ * r + c, 64);
        dp_wait(attn_done, uint32_t(st) & 1u, 2);  // R6 owns the ring's shared memory until here
        if (!dbg_ok(7)) break;
        put_a(base + off_v + 32 * r, 32);
        if (!dbg_ok(8)) break;
        put_a(base + off_co + 32 * r, 32);
        if (!dbg_ok(10)) break;
        for (int c = 0; c < ffs; c += 64) put_a(base + off_w1 + ffs * r + c, 64);
        for (int j = 0; j < nkb2; ++j) put_b((l * (ff / 64) + (ffs * r) / 64 + j) * 256);
        if (p.dbg_stop >= 0) break;
      }
    }
  } else if (warp == kDpWorkers + 1) {
    // ------------------------------------------------------------------ MMA issue
    if (lane == 0) {
      uint32_t cc = 0, ga = 0;
      // D[64 x N] (+)= A[64 x 256] . chunk^T, chunks of slab A rows; columns of D advance with the chunks
      auto gemm_a = [&](int n_total) {
        dp_wait(a_full, ga & 1u, 3);
        ++ga;
        tc_fence_after_sync();
        for (int n0 = 0; n0 < n_total;) {
          const int nrows = (n_total - n0) >= 64 ? 64 : 32;
          const uint32_t slot = cc % kDpSlots;
          dp_wait(&full[slot], (cc / kDpSlots) & 1u, 4);
          tc_fence_after_sync();
          const uint32_t sb = sR_u32 + slot * kDpSlotBytes;
          const uint32_t idesc = make_idesc(kFmtBF16, 64, uint32_t(nrows), 0, 0);
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const uint32_t a = sA_u32 + (k >> 2) * 8192 + (k & 3) * 32;
            const uint32_t b = sb + (k >> 2) * nrows * 128 + (k & 3) * 32;
            umma_bf16(tmem_base + uint32_t(n0), make_smem_desc_sw128(a, 16, 1024), make_smem_desc_sw128(b, 16, 1024),
                      idesc, k != 0);
          }
          umma_commit(&empty[slot]);
          ++cc;
          n0 += nrows;
        }
        umma_commit(acc_full);
      };
      for (int st = 0; st < n_stage; ++st) {
        gemm_a(96);       // G1
        if (!dbg_ok(3)) break;
        gemm_a(32);       // G3
        if (!dbg_ok(5)) break;
        gemm_a(NQ);       // G5
        if (!dbg_ok(7)) break;
        gemm_a(32);       // G7
        if (!dbg_ok(8)) break;
        gemm_a(32);       // G8
        if (!dbg_ok(10)) break;
        gemm_a(ffs);      // G10
        {                 // G11: D[64 x 256] = H[64 x ffs] . W2[:, slice]^T, one chunk per 64 hidden units
          dp_wait(a_full, ga & 1u, 5);
          ++ga;
          tc_fence_after_sync();
          const uint32_t idesc = make_idesc(kFmtBF16, 64, 256, 0, 0);
          for (int j = 0; j < nkb2; ++j) {
            const uint32_t slot = cc % kDpSlots;
            dp_wait(&full[slot], (cc / kDpSlots) & 1u, 6);
            tc_fence_after_sync();
            const uint32_t sb = sR_u32 + slot * kDpSlotBytes;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem_base, make_smem_desc_sw128(sA_u32 + j * 8192 + k * 32, 16, 1024),
                        make_smem_desc_sw128(sb + k * 32, 16, 1024), idesc, (j | k) != 0);
            umma_commit(&empty[slot]);
            ++cc;
          }
          umma_commit(acc_full);
        }
        if (p.dbg_stop >= 0) break;
      }
    }
  } else {
    // ------------------------------------------------------------------ workers
    const int w = warp;                       // question 8 r + w in the R phases
    const int wt = threadIdx.x;               // 0..255
    const int qrow = tile0 + r * kDpWorkers + w;
    const bool qvalid = qrow < p.B;
    // G-phase epilogue: TMEM lane quarter w & 3 (lanes 0..15 hold rows 16 (w & 3) + lane), column half w >> 2
    const int erow = tile0 + 16 * (w & 3) + lane;
    const bool evalid = lane < 16 && erow < p.B;
    const int ehalf = w >> 2;
    const uint32_t tlane = uint32_t(32 * (w & 3)) << 16;
    // A-operand copy: thread -> (row wt / 4, k-block wt % 4)
    const int arow = wt >> 2, akb = wt & 3;
    const bool avalid = tile0 + arow < p.B;
    const uint32_t a_dst = sA_u32 + akb * 8192 + arow * 128;
    uint32_t cb_phase = 0, acc_phase = 0, mem_phase = 0;  // mem_phase: bit s = parity of memory-tile slot s
    int n_cb = 0;
    bool stop = false;
    __nv_bfloat16* sp = reinterpret_cast<__nv_bfloat16*>(smem + kDpOffSp + w * 256);  // [NH][16] softmax weights
    const uint32_t ring = sR_u32 + w * kDpWarpRing;

    auto rendezvous = [&]() {
      fence_acq_rel_cluster();
      named_bar_sync(1, kDpWorkers * 32);
      if (wt == 0) {
        const uint32_t local = smem_u32(cb_bar);
#pragma unroll
        for (uint32_t dst = 0; dst < kDpCluster; ++dst) mbar_arrive_remote(cluster_map_shared(local, dst));
      }
      mbar_wait_cluster(cb_bar, cb_phase);
      cb_phase ^= 1u;
      ++n_cb;
      if (p.dbg_stop >= 0 && n_cb >= p.dbg_stop) stop = true;
    };
    // activation tile [64 rows x 256] at src (leading dimension ld elements) -> swizzled K-major A operand
    auto load_a = [&](const __nv_bfloat16* src, int ld) {
      uint4 v[8];
      const __nv_bfloat16* s = src + size_t(tile0 + arow) * ld + akb * 64;
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = avalid ? ldcg16(s + j * 8) : make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_dst + uint32_t((j ^ (arow & 7)) << 4)),
                     "r"(v[j].x), "r"(v[j].y), "r"(v[j].z), "r"(v[j].w)
                     : "memory");
      fence_proxy_async_smem();
      tc_fence_before_sync();
      mbar_arrive(a_full);
    };
    auto wait_acc = [&]() {
      dp_wait(acc_full, acc_phase, 7);
      acc_phase ^= 1u;
      __syncwarp();
      tc_fence_after_sync();
    };
    // f(col, v[16]) for this thread's row and every 16-column group of its half of the n_local accumulator columns
    auto epilogue = [&](int n_local, auto&& f) {
      const int nh = n_local >> 1;
      for (int c = 0; c < nh; c += 16) {
        uint32_t v[16];
        tmem_ld16(tmem_base + tlane + uint32_t(ehalf * nh + c), v);
        tmem_ld_wait();
        if (evalid) f(ehalf * nh + c, v);
      }
      tc_fence_before_sync();
    };
    auto issue_tile = [&](int i) {  // lane 0: memory rows [16 i, 16 i + 16) of this warp's question
      const int s = i & 1;
      const uint32_t dst = ring + s * 8192;
      mbar_expect_tx(&memfull[w * 2 + s], 8192);
#pragma unroll
      for (int cb = 0; cb < 4; ++cb)
        tma_load_2d_u32(&tm_mem, &memfull[w * 2 + s], dst + cb * 2048, cb * 64, qrow * kLP + i * 16);
    };
    int mlen = 0;
    if (qvalid) {
      mlen = p.lens ? p.lens[qrow] : p.const_len;
      mlen = mlen > kLP ? kLP : (mlen < 1 ? 1 : mlen);
    }
    const int n_tiles = (mlen + 15) >> 4;

    if (p.stagger_cycles > 0 && (blockIdx.y & 1)) {
      const long long t0 = clock64();
      while (clock64() - t0 < p.stagger_cycles) {}
    }

    // optional timeline of CTA (0, 0): clock64() of worker thread 0 at 16 points of each of the first 8 stages,
    // + [16] cycles its warp waited for memory tiles in R6, [17] R6 start, [18] queries loaded (tools/)
    const bool stamping = p.dbg_clk != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && wt == 0;
#define DP_STAMP(i) \
  if (stamping && st < 8) p.dbg_clk[st * 24 + (i)] = clock64()
    for (int st = 0; st < n_stage && !stop; ++st) {
      const int t = st / p.n_layers, l = st % p.n_layers;
      const bool last = l == p.n_layers - 1;
      const DecLayerDev L = p.layers[l];
      const __nv_bfloat16* xin = l == 0 ? p.dx : p.dxo[(l - 1) & 1];
      __nv_bfloat16* xout = p.dxo[l & 1];

      // ---------------------------------------------------------------- G1: qkv
      DP_STAMP(0);
      load_a(xin, kD);
      DP_STAMP(1);
      wait_acc();
      DP_STAMP(2);
      epilogue(96, [&](int col, const uint32_t (&v)[16]) {
        const int n = 96 * r + col;
        float b[16];
        ldg8(L.b_in + n, *reinterpret_cast<float(*)[8]>(b));
        ldg8(L.b_in + n + 8, *reinterpret_cast<float(*)[8]>(b + 8));
        uint32_t o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
          o[j] = pack_bf16x2(__uint_as_float(v[2 * j]) + b[2 * j], __uint_as_float(v[2 * j + 1]) + b[2 * j + 1]);
        uint4* dst = reinterpret_cast<uint4*>(p.dqkv + size_t(erow) * 3 * kD + n);
        dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
        dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
      });
      DP_STAMP(3);
      rendezvous();
      DP_STAMP(4);
      if (stop) break;

      // ---------------------------------------------------------------- R2: self-attention over <= 32 keys
      if (qvalid) {
        constexpr int LPH = DH / 8;  // lanes per head
        const int n_old = t;         // cached keys; this position's own key comes from the projection
        const uint4 zero = make_uint4(0, 0, 0, 0);
        __nv_bfloat16* kcache = p.kc[l] + size_t(qrow) * p.t_max * kD + lane * 8;
        __nv_bfloat16* vcache = p.vc[l] + size_t(qrow) * p.t_max * kD + lane * 8;
        const __nv_bfloat16* qkv = p.dqkv + size_t(qrow) * 3 * kD + lane * 8;
        const uint4 fq = ldcg16(qkv), fk = ldcg16(qkv + kD), fv = ldcg16(qkv + 2 * kD);
        // K rows first, V rows after the scores: both sets at once (128 registers) would not fit beside the scores
        uint4 kr[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) kr[u] = u < n_old ? ldcg16(kcache + size_t(u) * kD) : zero;
        float q[8];
        unpack8(fq, q);
        const float sl2 = rsqrtf(float(DH)) * 1.4426950408889634f;
#pragma unroll
        for (int e = 0; e < 8; ++e) q[e] *= sl2;
        *reinterpret_cast<uint4*>(kcache + size_t(t) * kD) = fk;
        *reinterpret_cast<uint4*>(vcache + size_t(t) * kD) = fv;
        auto score = [&](const uint4& raw) {
          float kx[8];
          unpack8(raw, kx);
          float sc = 0.f;
#pragma unroll
          for (int e = 0; e < 8; ++e) sc = fmaf(q[e], kx[e], sc);
#pragma unroll
          for (int o = LPH / 2; o > 0; o >>= 1) sc += __shfl_xor_sync(0xffffffffu, sc, o);
          return sc;
        };
        float sc[32];
#pragma unroll
        for (int u = 0; u < 16; ++u) sc[u] = u < n_old ? score(kr[u]) : -INFINITY;
        if (n_old > 16) {
#pragma unroll
          for (int u = 0; u < 16; ++u) kr[u] = 16 + u < n_old ? ldcg16(kcache + size_t(16 + u) * kD) : zero;
#pragma unroll
          for (int u = 0; u < 16; ++u) sc[16 + u] = 16 + u < n_old ? score(kr[u]) : -INFINITY;
        } else {
#pragma unroll
          for (int u = 0; u < 16; ++u) sc[16 + u] = -INFINITY;
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) kr[u] = u < n_old ? ldcg16(vcache + size_t(u) * kD) : zero;  // now the V rows
        const float s_new = score(fk);
        float mx = s_new;
#pragma unroll
        for (int j = 0; j < 32; ++j) mx = fmaxf(mx, sc[j]);
        float acc[8], sum;
        {
          const float pn = exp2f(s_new - mx);
          float vx[8];
          unpack8(fv, vx);
          sum = pn;
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[e] = pn * vx[e];
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const float pj = exp2f(sc[u] - mx);  // exactly 0 for the padding entries
          float vx[8];
          unpack8(kr[u], vx);
          sum += pj;
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[e] = fmaf(pj, vx[e], acc[e]);
        }
        if (n_old > 16) {
#pragma unroll
          for (int u = 0; u < 16; ++u) kr[u] = 16 + u < n_old ? ldcg16(vcache + size_t(16 + u) * kD) : zero;
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            const float pj = exp2f(sc[16 + u] - mx);
            float vx[8];
            unpack8(kr[u], vx);
            sum += pj;
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] = fmaf(pj, vx[e], acc[e]);
          }
        }
        const float inv = 1.f / sum;
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] *= inv;
        *reinterpret_cast<uint4*>(p.dattn + size_t(qrow) * kD + lane * 8) = pack8(acc);
      }
      rendezvous();
      DP_STAMP(5);
      if (stop) break;

      // ---------------------------------------------------------------- G3: out_proj + residual -> pre-LN sums
      auto out_proj_epilogue = [&](const float* bias, const __nv_bfloat16* residual) {
        epilogue(32, [&](int col, const uint32_t (&v)[16]) {
          const int n = 32 * r + col;
          float b[16], res[16];
          ldg8(bias + n, *reinterpret_cast<float(*)[8]>(b));
          ldg8(bias + n + 8, *reinterpret_cast<float(*)[8]>(b + 8));
          const __nv_bfloat16* rs = residual + size_t(erow) * kD + n;
          unpack8(ldcg16(rs), *reinterpret_cast<float(*)[8]>(res));
          unpack8(ldcg16(rs + 8), *reinterpret_cast<float(*)[8]>(res + 8));
          float4* dst = reinterpret_cast<float4*>(p.dpre + size_t(erow) * kD + n);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            dst[j] = make_float4(__uint_as_float(v[4 * j]) + b[4 * j] + res[4 * j],
                                 __uint_as_float(v[4 * j + 1]) + b[4 * j + 1] + res[4 * j + 1],
                                 __uint_as_float(v[4 * j + 2]) + b[4 * j + 2] + res[4 * j + 2],
                                 __uint_as_float(v[4 * j + 3]) + b[4 * j + 3] + res[4 * j + 3]);
        });
      };
      auto ln_rows = [&](const float* gamma, const float* beta, __nv_bfloat16* dst) {
        if (qvalid) {
          float v[8];
          const float4 a = ldcg_f4(p.dpre + size_t(qrow) * kD + lane * 8);
          const float4 b = ldcg_f4(p.dpre + size_t(qrow) * kD + lane * 8 + 4);
          v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
          warp_layernorm(v, gamma, beta, p.eps, lane);
          *reinterpret_cast<uint4*>(dst + size_t(qrow) * kD + lane * 8) = pack8(v);
        }
      };
      load_a(p.dattn, kD);
      wait_acc();
      out_proj_epilogue(L.b_out, xin);
      rendezvous();
      DP_STAMP(6);
      if (stop) break;
      // ---------------------------------------------------------------- R4: LN1
      ln_rows(L.n1w, L.n1b, p.dx1);
      rendezvous();
      DP_STAMP(7);
      if (stop) break;

      // ---------------------------------------------------------------- G5: absorbed queries
      load_a(p.dx1, kD);
      wait_acc();
      DP_STAMP(8);
      // every weight chunk issued so far has been consumed: the ring's shared memory is this warp's until attn_done
      if (lane == 0) {
        if (n_tiles > 0) issue_tile(0);
        if (n_tiles > 1) issue_tile(1);
      }
      epilogue(NQ, [&](int col, const uint32_t (&v)[16]) {
        const int n = NQ * r + col;
        float b[16];
        ldg8(L.b_qk + n, *reinterpret_cast<float(*)[8]>(b));
        ldg8(L.b_qk + n + 8, *reinterpret_cast<float(*)[8]>(b + 8));
        uint32_t o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
          o[j] = pack_bf16x2(__uint_as_float(v[2 * j]) + b[2 * j], __uint_as_float(v[2 * j + 1]) + b[2 * j + 1]);
        uint4* dst = reinterpret_cast<uint4*>(p.dq + size_t(erow) * NH * kD + n);
        dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
        dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
      });
      rendezvous();
      DP_STAMP(9);
      // (a debug stop must still drain the two tiles in flight: fall through R6 and stop at its rendezvous)

      // ---------------------------------------------------------------- R6: absorbed cross-attention on the memory
      {
        DP_STAMP(17);
        const int g = lane >> 2, q4 = lane & 3;
        const float sl2 = rsqrtf(float(DH)) * 1.4426950408889634f;
        uint32_t bq[16][2];
        {
          const __nv_bfloat16* qp = p.dq + (size_t(qrow) * NH + (g < NH ? g : 0)) * kD + 2 * q4;
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            bq[k][0] = (qvalid && g < NH) ? __ldcg(reinterpret_cast<const uint32_t*>(qp + 16 * k)) : 0u;
            bq[k][1] = (qvalid && g < NH) ? __ldcg(reinterpret_cast<const uint32_t*>(qp + 16 * k + 8)) : 0u;
          }
        }
        DP_STAMP(18);
        long long waited = 0;
        float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
        float acc[16][4];
#pragma unroll
        for (int mt = 0; mt < 16; ++mt)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[mt][e] = 0.f;
        for (int i = 0; i < n_tiles; ++i) {
          const int s = i & 1;
          const long long tw0 = stamping ? clock64() : 0;
          dp_wait(&memfull[w * 2 + s], (mem_phase >> s) & 1u, 8);
          if (stamping) waited += clock64() - tw0;
          mem_phase ^= 1u << s;
          const uint32_t tile = ring + s * 8192;
          float c[4][4];
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int e = 0; e < 4; ++e) c[a][e] = 0.f;
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            uint32_t af[4];
            ldmatrix_x4(tile + mtile_off(lane & 15, k * 16 + (lane >> 4) * 8), af);
            mma_bf16_16816(c[k & 3], af, bq[k][0], bq[k][1]);
          }
          // sv[0,1] = S[row g][heads 2 q4, 2 q4 + 1], sv[2,3] = the same heads of row g + 8
          float sv[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) sv[e] = (c[0][e] + c[1][e] + c[2][e] + c[3][e]) * sl2;
          const int r0 = i * 16 + g;
          if (r0 >= mlen) sv[0] = sv[1] = -INFINITY;
          if (r0 + 8 >= mlen) sv[2] = sv[3] = -INFINITY;
          float tm0 = fmaxf(sv[0], sv[2]), tm1 = fmaxf(sv[1], sv[3]);
#pragma unroll
          for (int o = 4; o <= 16; o <<= 1) {
            tm0 = fmaxf(tm0, __shfl_xor_sync(0xffffffffu, tm0, o));
            tm1 = fmaxf(tm1, __shfl_xor_sync(0xffffffffu, tm1, o));
          }
          const float mn0 = fmaxf(m_run[0], tm0), mn1 = fmaxf(m_run[1], tm1);  // finite: row 16 i is valid
          const float al0 = exp2f(m_run[0] - mn0), al1 = exp2f(m_run[1] - mn1);
          m_run[0] = mn0;
          m_run[1] = mn1;
          const float p00 = exp2f(sv[0] - mn0), p01 = exp2f(sv[1] - mn1), p10 = exp2f(sv[2] - mn0),
                      p11 = exp2f(sv[3] - mn1);
          l_run[0] = l_run[0] * al0 + p00 + p10;
          l_run[1] = l_run[1] * al1 + p01 + p11;
          if (2 * q4 < NH) {
            sp[(2 * q4) * 16 + g] = __float2bfloat16(p00);
            sp[(2 * q4 + 1) * 16 + g] = __float2bfloat16(p01);
            sp[(2 * q4) * 16 + g + 8] = __float2bfloat16(p10);
            sp[(2 * q4 + 1) * 16 + g + 8] = __float2bfloat16(p11);
          }
          __syncwarp();
          // B fragments of P^T: rows (k) = memory rows of the tile, columns (n) = heads
          uint32_t pb0 = 0u, pb1 = 0u;
          if (g < NH) {
            pb0 = *reinterpret_cast<const uint32_t*>(sp + g * 16 + 2 * q4);
            pb1 = *reinterpret_cast<const uint32_t*>(sp + g * 16 + 2 * q4 + 8);
          }
          __syncwarp();
#pragma unroll
          for (int mt = 0; mt < 16; ++mt) {
            acc[mt][0] *= al0;
            acc[mt][1] *= al1;
            acc[mt][2] *= al0;
            acc[mt][3] *= al1;
          }
          // A = M^T through ldmatrix.trans: channels on the 16 MMA rows, the tile's 16 memory rows on K
          const int tr = ((lane >> 4) & 1) * 8 + (lane & 7);
          const int tc = ((lane >> 3) & 1) * 8;
#pragma unroll
          for (int mt = 0; mt < 16; ++mt) {
            uint32_t am[4];
            ldmatrix_x4_trans(tile + mtile_off(tr, mt * 16 + tc), am);
            mma_bf16_16816(acc[mt], am, pb0, pb1);
          }
          __syncwarp();
          if (lane == 0 && i + 2 < n_tiles) issue_tile(i + 2);
        }
        if (qvalid) {
          float l0 = l_run[0], l1 = l_run[1];
#pragma unroll
          for (int o = 4; o <= 16; o <<= 1) {
            l0 += __shfl_xor_sync(0xffffffffu, l0, o);
            l1 += __shfl_xor_sync(0xffffffffu, l1, o);
          }
          const float inv0 = 1.f / l0, inv1 = 1.f / l1;
          // this lane: heads 2 q4, 2 q4 + 1 of channels 16 mt + g and + 8 -> [NH][256] bf16 in the warp's ring, then
          // 16-byte coalesced stores
          __nv_bfloat16* su = reinterpret_cast<__nv_bfloat16*>(smem + kDpOffR + w * kDpWarpRing);
          if (2 * q4 < NH) {
#pragma unroll
            for (int mt = 0; mt < 16; ++mt) {
              su[(2 * q4) * kD + mt * 16 + g] = __float2bfloat16(acc[mt][0] * inv0);
              su[(2 * q4 + 1) * kD + mt * 16 + g] = __float2bfloat16(acc[mt][1] * inv1);
              su[(2 * q4) * kD + mt * 16 + g + 8] = __float2bfloat16(acc[mt][2] * inv0);
              su[(2 * q4 + 1) * kD + mt * 16 + g + 8] = __float2bfloat16(acc[mt][3] * inv1);
            }
          }
          __syncwarp();
          uint4* dst = reinterpret_cast<uint4*>(p.du + size_t(qrow) * NH * kD);
#pragma unroll
          for (int ch = lane; ch < NH * 32; ch += 32) dst[ch] = reinterpret_cast<const uint4*>(su)[ch];
          __syncwarp();
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(attn_done);
        if (stamping && st < 8) p.dbg_clk[st * 24 + 16] = waited;
      }
      DP_STAMP(10);
      rendezvous();
      DP_STAMP(11);
      if (stop) break;

      // ---------------------------------------------------------------- G7: per-head value projection
      load_a(p.du + ((32 * r) / DH) * kD, NH * kD);
      wait_acc();
      epilogue(32, [&](int col, const uint32_t (&v)[16]) {
        const int n = 32 * r + col;
        float b[16];
        ldg8(L.b_v + n, *reinterpret_cast<float(*)[8]>(b));
        ldg8(L.b_v + n + 8, *reinterpret_cast<float(*)[8]>(b + 8));
        uint32_t o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
          o[j] = pack_bf16x2(__uint_as_float(v[2 * j]) + b[2 * j], __uint_as_float(v[2 * j + 1]) + b[2 * j + 1]);
        uint4* dst = reinterpret_cast<uint4*>(p.dattn + size_t(erow) * kD + n);
        dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
        dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
      });
      rendezvous();
      DP_STAMP(12);
      if (stop) break;

      // ---------------------------------------------------------------- G8 + R9: cross out_proj + residual, LN2
      load_a(p.dattn, kD);
      wait_acc();
      out_proj_epilogue(L.b_co, p.dx1);
      rendezvous();
      if (stop) break;
      ln_rows(L.n2w, L.n2b, p.dx2);
      rendezvous();
      DP_STAMP(13);
      if (stop) break;

      // ---------------------------------------------------------------- G10: hidden slice -> shared memory (A of G11)
      load_a(p.dx2, kD);
      wait_acc();
      epilogue(ffs, [&](int col, const uint32_t (&v)[16]) {
        float b[16];
        ldg8(L.b1 + ffs * r + col, *reinterpret_cast<float(*)[8]>(b));
        ldg8(L.b1 + ffs * r + col + 8, *reinterpret_cast<float(*)[8]>(b + 8));
        uint32_t o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
          o[j] = pack_bf16x2(fmaxf(__uint_as_float(v[2 * j]) + b[2 * j], 0.f),
                             fmaxf(__uint_as_float(v[2 * j + 1]) + b[2 * j + 1], 0.f));
        const int hr = 16 * (w & 3) + lane;  // row inside the tile
        const uint32_t base = sA_u32 + (col >> 6) * 8192 + hr * 128;
        const int ch = (col & 63) >> 3;
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(base + uint32_t((ch ^ (hr & 7)) << 4)), "r"(o[0]),
                     "r"(o[1]), "r"(o[2]), "r"(o[3])
                     : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(base + uint32_t(((ch + 1) ^ (hr & 7)) << 4)),
                     "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7])
                     : "memory");
      });
      // (rows of the tile past the batch keep the zeros load_a wrote: finite, and their accumulator rows are never stored)
      fence_proxy_async_smem();
      tc_fence_before_sync();
      mbar_arrive(a_full);
      // ---------------------------------------------------------------- G11: partial sums of this hidden slice
      wait_acc();
      epilogue(256, [&](int col, const uint32_t (&v)[16]) {
        float4* dst = reinterpret_cast<float4*>(p.partial + (size_t(r) * p.part_rows + erow) * kD + col);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          dst[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                               __uint_as_float(v[4 * j + 3]));
      });
      rendezvous();
      DP_STAMP(14);
      if (stop) break;

      // ---------------------------------------------------------------- R12: reduce + LN3 (+ final norm, head, next input)
      if (qvalid) {
        float v[8];
        {
          float b2[8], res[8];
          ldg8(L.b2 + lane * 8, b2);
          unpack8(ldcg16(p.dx2 + size_t(qrow) * kD + lane * 8), res);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = b2[j] + res[j];
        }
        {
          float4 a[kDpCluster], b[kDpCluster];
          const float* src = p.partial + size_t(qrow) * kD + lane * 8;
#pragma unroll
          for (int u = 0; u < kDpCluster; ++u) {
            a[u] = ldcg_f4(src + size_t(u) * p.part_rows * kD);
            b[u] = ldcg_f4(src + size_t(u) * p.part_rows * kD + 4);
          }
#pragma unroll
          for (int u = 0; u < kDpCluster; ++u) {
            v[0] += a[u].x; v[1] += a[u].y; v[2] += a[u].z; v[3] += a[u].w;
            v[4] += b[u].x; v[5] += b[u].y; v[6] += b[u].z; v[7] += b[u].w;
          }
        }
        warp_layernorm(v, L.n3w, L.n3b, p.eps, lane);
        *reinterpret_cast<uint4*>(xout + size_t(qrow) * kD + lane * 8) = pack8(v);
        if (last) {
          if (p.fn_gamma) warp_layernorm(v, p.fn_gamma, p.fn_beta, p.eps, lane);
          // vocabulary head: 32 tokens per pass, lane j ends with the logit of token 32 grp + j
          float best = -INFINITY;
          int besti = 0x7fffffff;
          const int ngrp = (p.head_V + 31) >> 5;
#pragma unroll 1
          for (int grp = 0; grp < ngrp; ++grp) {
            float x[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int tokv = grp * 32 + i;
              float a = 0.f;
              if (tokv < p.head_V) {  // warp-uniform
                float wv[8];
                ldg8(p.head_w + size_t(tokv) * kD + lane * 8, wv);
                a = v[0] * wv[0];
#pragma unroll
                for (int j = 1; j < 8; ++j) a = fmaf(v[j], wv[j], a);
              }
              x[i] = a;
            }
            // transposed warp reduction (31 shuffles for 32 sums)
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
              const bool upper = (lane & off) != 0;
#pragma unroll
              for (int i = 0; i < off; ++i) {
                const float send = upper ? x[i] : x[i + off];
                const float keep = upper ? x[i + off] : x[i];
                x[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
              }
            }
            const int tokv = grp * 32 + lane;
            const float lg = tokv < p.head_V ? x[0] + __ldg(p.head_b + tokv) : -INFINITY;
            if (p.logits && tokv < p.head_V) p.logits[(size_t(qrow) * p.logits_T + t) * p.head_V + tokv] = lg;
            if (lg > best) {  // strict: the lowest index wins ties inside a lane (groups ascend)
              best = lg;
              besti = tokv;
            }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
            if (ob > best || (ob == best && oi < besti)) {
              best = ob;
              besti = oi;
            }
          }
          if (besti >= p.head_V) besti = 0;
          if (lane == 0) p.tok[size_t(qrow) * p.tok_ld + t + 1] = besti;
          if (t + 1 < p.steps) {  // next decoder input = emb[next token] + pe[t + 1]
            long long nxt = p.forced ? p.forced[size_t(qrow) * p.forced_ld + t] : (long long)besti;
            nxt = nxt < 0 ? 0 : (nxt >= p.vocab ? p.vocab - 1 : nxt);
            float e[8], pe[8];
            ldg8(p.emb + size_t(nxt) * kD + lane * 8, e);
            ldg8(p.pe + size_t(t + 1) * kD + lane * 8, pe);
#pragma unroll
            for (int j = 0; j < 8; ++j) e[j] += pe[j];
            *reinterpret_cast<uint4*>(p.dx + size_t(qrow) * kD + lane * 8) = pack8(e);
          }
        }
      }
      rendezvous();
      DP_STAMP(15);
    }
#undef DP_STAMP
  }

  // nobody leaves while a peer may still arrive on its barriers
  __syncwarp();
  tc_fence_before_sync();
  cluster_arrive_release();
  cluster_wait_acquire();
  if (warp == kDpWorkers + 1) {
    tc_fence_after_sync();
    tmem_dealloc<256>(tmem_base);
  }
}

__global__ void pack_w2_kblocks_kernel(const __nv_bfloat16* __restrict__ w2, __nv_bfloat16* __restrict__ out, int ff) {
  // out row kb * 256 + n = W2[n][64 kb .. 64 kb + 63]; one 16-byte chunk per thread
  const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t total = size_t(kD) * ff / 8;
  if (i >= total) return;
  const int ch = int(i & 7);
  const size_t row = i >> 3;
  const int n = int(row % kD), kb = int(row / kD);
  reinterpret_cast<uint4*>(out)[i] = *reinterpret_cast<const uint4*>(w2 + size_t(n) * ff + kb * 64 + ch * 8);
}

}  // namespace

cudaError_t launch_pack_w2_kblocks(const __nv_bfloat16* w2, __nv_bfloat16* out, int ff, cudaStream_t stream) {
  const size_t total = size_t(kD) * ff / 8;
  pack_w2_kblocks_kernel<<<unsigned((total + 255) / 256), 256, 0, stream>>>(w2, out, ff);
  return cudaGetLastError();
}

cudaError_t launch_decode_persist(const CUtensorMap& tm_wa, const CUtensorMap& tm_wb, const CUtensorMap& tm_mem,
                                  const DecPersistParams& p, cudaStream_t stream) {
  if (p.B <= 0 || p.steps <= 0) return cudaSuccess;
  if (p.steps > kDecPersistMaxSteps || p.ff % (64 * kDpCluster) != 0 || p.ff / kDpCluster > 256 || p.head_V < 1 ||
      p.n_layers < 1)
    return cudaErrorInvalidValue;
  const dim3 grid(kDpCluster, (p.B + kDpTile - 1) / kDpTile);
  if (getenv("B200VQA_PERSIST_VERBOSE")) {
    static bool once = false;
    if (!once) {
      once = true;
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = grid;
      cfg.blockDim = dim3(kDpThreads);
      cfg.dynamicSmemBytes = kDpSmem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = kDpCluster;
      at[0].val.clusterDim.y = 1;
      at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      int n = -1;
      ensure_dyn_smem(reinterpret_cast<const void*>(decode_persist_kernel<4>), kDpSmem);
      cudaError_t e = cudaOccupancyMaxActiveClusters(&n, decode_persist_kernel<4>, &cfg);
      fprintf(stderr, "b200vqa: decode_persist grid %d x %d, smem %d B, max active clusters %d (%s)\n", grid.x, grid.y,
              kDpSmem, n, cudaGetErrorString(e));
    }
  }
  if (p.nhead == 4) {
    cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void*>(decode_persist_kernel<4>), kDpSmem);
    if (e != cudaSuccess) return e;
    return launch_kernel(decode_persist_kernel<4>, grid, dim3(kDpThreads), kDpSmem, stream, false, tm_wa, tm_wb, tm_mem, p);
  }
  if (p.nhead == 2) {
    cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void*>(decode_persist_kernel<2>), kDpSmem);
    if (e != cudaSuccess) return e;
    return launch_kernel(decode_persist_kernel<2>, grid, dim3(kDpThreads), kDpSmem, stream, false, tm_wa, tm_wb, tm_mem, p);
  }
  return cudaErrorInvalidValue;
}

}  // namespace b200vqa
