// Decode-phase kernels: one new position per question and step.  These are bandwidth- or latency-bound
// (M = batch rows, one query row per question), so they run on CUDA cores with 16-byte coalesced accesses;
// the dense layers between them go through the tcgen05 GEMM (gemm.cu).
//
//   dec_embed_start   x0 = emb[start] + pe[0]                                   (IQAP:205-216, FA:136-139)
//   row_attn          one query row against len key/value rows, all heads       (IQAP:223-227, FA:141)
//   (the vocabulary head + argmax + next embedding is the kEpiHead epilogue of the tensor-core GEMM, gemm.cu)
//   publish_tokens    library token buffer -> programs / ys / HBM step cache    (IQAP:239, FA:120-121)
#include <algorithm>
#include <cstdlib>

#include "kernels.h"
#include "ptx.cuh"

namespace b200vqa {
namespace {

__device__ __forceinline__ void store_bf16x8(__nv_bfloat16* dst, const float (&v)[8]) {
  *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]),
                                              pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&v)[8]) {
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
__device__ __forceinline__ void load_f32x8(const float* src, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(src));
  const float4 b = __ldg(reinterpret_cast<const float4*>(src) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ uint4 ld_stream16(const void* p) {
  uint4 r;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p)
               : "memory");
  return r;
}

__global__ void dec_embed_start_kernel(const DecEmbedParams p) {
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (b >= p.B) return;
  float e[8], pe8[8], v[8];
  long long st = p.start_tokens ? p.start_tokens[size_t(b) * p.start_ld] : (long long)p.start_token;
  st = st < 0 ? 0 : (st >= p.vocab ? p.vocab - 1 : st);
  load_f32x8(p.emb + size_t(st) * kD + lane * 8, e);
  load_f32x8(p.pe + lane * 8, pe8);
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = e[j] + pe8[j];
  store_bf16x8(p.x + size_t(b) * kD + lane * 8, v);
  if (lane == 0) p.tok[size_t(b) * p.tok_ld] = st;
}

// ------------------------------------------------------------------------------------------------
// One CTA (8 warps) per question.  A key row is 256 bf16 = 512 B = one 16-byte chunk per lane, so every
// warp-wide load is one fully coalesced row; each warp takes every 8th key, four keys in flight per pass.
// Scores of all heads live in shared memory between the two passes: K and V are each read exactly once.
// ------------------------------------------------------------------------------------------------
constexpr int kAttnWarps = 8;
constexpr int kKeysInFlight = 4;

template <int DH>
__global__ void __launch_bounds__(kAttnWarps * 32) row_attn_kernel(const RowAttnParams p) {
  constexpr int LPH = DH / 8;  // lanes per head
  __shared__ float s_sc[4][kLP];
  __shared__ float s_red[kAttnWarps][kD];
  __shared__ float s_inv[4];
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = lane / LPH;
  pdl_launch_dependents();
  pdl_wait();
  int len = p.lens ? p.lens[b] : p.const_len;
  len = len > kLP ? kLP : len;

  // self-attention: this position's key / value are appended to the caches for the later positions; the attention
  // below takes row append_pos straight from new_k / new_v, so nothing waits for the appended rows to land
  const __nv_bfloat16* fresh_k = p.new_k ? p.new_k + size_t(b) * p.ld_new + lane * 8 : nullptr;
  const __nv_bfloat16* fresh_v = p.new_k ? p.new_v + size_t(b) * p.ld_new + lane * 8 : nullptr;
  if (p.new_k && threadIdx.x >= 64 && threadIdx.x < 128) {  // warps 2 and 3: their first row batch is the shortest
    const bool is_v = threadIdx.x >= 96;
    __nv_bfloat16* dst = (is_v ? p.v_app : p.k_app) + (size_t(b) * p.rows_per_q + p.append_pos) * p.ld + lane * 8;
    *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(is_v ? fresh_v : fresh_k);
  }

  float q[8];
  unpack8(*reinterpret_cast<const uint4*>(p.q + size_t(b) * p.ldq + lane * 8), q);
  const float sl2 = rsqrtf(float(DH)) * 1.4426950408889634f;  // 1/sqrt(dh) * log2(e): softmax via exp2
#pragma unroll
  for (int e = 0; e < 8; ++e) q[e] *= sl2;

  const __nv_bfloat16* kbase = p.k + size_t(b) * p.rows_per_q * p.ld + lane * 8;
  const __nv_bfloat16* vbase = p.v + size_t(b) * p.rows_per_q * p.ld + lane * 8;

  // software-pipelined: the next batch of rows is in flight while the current one is consumed, and the first
  // batch of V rows is requested before the softmax barrier so the HBM stream never drains
  constexpr int kStep = kAttnWarps * kKeysInFlight;
  auto load_rows = [&](const __nv_bfloat16* base, const __nv_bfloat16* fresh, int j0, uint4 (&raw)[kKeysInFlight]) {
#pragma unroll
    for (int u = 0; u < kKeysInFlight; ++u) {
      const int j = j0 + u * kAttnWarps;
      const __nv_bfloat16* src = (fresh && j == p.append_pos) ? fresh : base + size_t(j) * p.ld;
      raw[u] = (j < len) ? ld_stream16(src) : make_uint4(0, 0, 0, 0);
    }
  };
  uint4 cur[kKeysInFlight], nxt[kKeysInFlight];
  load_rows(kbase, fresh_k, warp, cur);
  for (int j0 = warp; j0 < len; j0 += kStep) {
    if (j0 + kStep < len) load_rows(kbase, fresh_k, j0 + kStep, nxt);
#pragma unroll
    for (int u = 0; u < kKeysInFlight; ++u) {
      const int j = j0 + u * kAttnWarps;
      float kx[8];
      unpack8(cur[u], kx);
      float s = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) s = fmaf(q[e], kx[e], s);
#pragma unroll
      for (int o = LPH / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if ((lane % LPH) == 0 && j < len) s_sc[head][j] = s;
    }
#pragma unroll
    for (int u = 0; u < kKeysInFlight; ++u) cur[u] = nxt[u];
  }
  load_rows(vbase, fresh_v, warp, cur);  // first V batch: in flight across the softmax
  __syncthreads();

  if (warp < p.nhead) {  // softmax statistics of head `warp`
    float mx = -INFINITY;
    for (int j = lane; j < len; j += 32) mx = fmaxf(mx, s_sc[warp][j]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < len; j += 32) {
      const float e = exp2f(s_sc[warp][j] - mx);
      s_sc[warp][j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    if (lane == 0) s_inv[warp] = 1.f / sum;
  }
  __syncthreads();

  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  for (int j0 = warp; j0 < len; j0 += kStep) {
    if (j0 + kStep < len) load_rows(vbase, fresh_v, j0 + kStep, nxt);
#pragma unroll
    for (int u = 0; u < kKeysInFlight; ++u) {
      const int j = j0 + u * kAttnWarps;
      if (j < len) {
        float vx[8];
        unpack8(cur[u], vx);
        const float pj = s_sc[head][j];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = fmaf(pj, vx[e], acc[e]);
      }
    }
#pragma unroll
    for (int u = 0; u < kKeysInFlight; ++u) cur[u] = nxt[u];
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) s_red[warp][lane * 8 + e] = acc[e];
  __syncthreads();
  {
    const int t = threadIdx.x;  // one output channel per thread
    float o = 0.f;
#pragma unroll
    for (int w = 0; w < kAttnWarps; ++w) o += s_red[w][t];
    p.out[size_t(b) * kD + t] = __float2bfloat16(o * s_inv[t / DH]);
  }
}

// ------------------------------------------------------------------------------------------------
// Decoder self-attention of the first 32 positions: ONE WARP per question, no shared memory, no CTA barrier.  A lane
// owns 8 channels (one 16-byte chunk of every 512-byte row, so each warp-wide load is one coalesced row); the cached key
// and value rows (at most 31) are requested up front, 16 at a time, the new row comes straight from the projection and
// is appended to the caches on the side.  Scores are reduced over the DH / 8 lanes of a head with shuffles and stay in
// registers; softmax via exp2.  (row_attn_kernel above spends three CTA barriers and a shared-memory reduction on the
// same 14 KB of data: 9.8 us per 1024 questions under ncu.)
// ------------------------------------------------------------------------------------------------
constexpr int kSelfWarpMaxLen = 32;

template <int DH>
__global__ void __launch_bounds__(128) self_attn_warp_kernel(const RowAttnParams p) {
  constexpr int LPH = DH / 8;  // lanes per head
  const int lane = threadIdx.x & 31;
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  pdl_launch_dependents();
  pdl_wait();
  if (b >= p.B) return;
  const int len = p.const_len;  // 1 .. 32 keys, the last one is this position's own
  const int n_old = len - 1;
  const uint4 zero = make_uint4(0, 0, 0, 0);
  const __nv_bfloat16* kbase = p.k + size_t(b) * p.rows_per_q * p.ld + lane * 8;
  const __nv_bfloat16* vbase = p.v + size_t(b) * p.rows_per_q * p.ld + lane * 8;
  const uint4 fk = ld_stream16(p.new_k + size_t(b) * p.ld_new + lane * 8);
  const uint4 fv = ld_stream16(p.new_v + size_t(b) * p.ld_new + lane * 8);
  uint4 kr[16], vr[16];
#pragma unroll
  for (int u = 0; u < 16; ++u) kr[u] = u < n_old ? ld_stream16(kbase + size_t(u) * p.ld) : zero;
#pragma unroll
  for (int u = 0; u < 16; ++u) vr[u] = u < n_old ? ld_stream16(vbase + size_t(u) * p.ld) : zero;
  float q[8];
  unpack8(*reinterpret_cast<const uint4*>(p.q + size_t(b) * p.ldq + lane * 8), q);
  const float sl2 = rsqrtf(float(DH)) * 1.4426950408889634f;  // 1/sqrt(dh) * log2(e): softmax via exp2
#pragma unroll
  for (int e = 0; e < 8; ++e) q[e] *= sl2;
  *reinterpret_cast<uint4*>(p.k_app + (size_t(b) * p.rows_per_q + p.append_pos) * p.ld + lane * 8) = fk;
  *reinterpret_cast<uint4*>(p.v_app + (size_t(b) * p.rows_per_q + p.append_pos) * p.ld + lane * 8) = fv;

  auto score = [&](const uint4& raw) {
    float kx[8];
    unpack8(raw, kx);
    float sc = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) sc = fmaf(q[e], kx[e], sc);
#pragma unroll
    for (int o = LPH / 2; o > 0; o >>= 1) sc += __shfl_xor_sync(0xffffffffu, sc, o);
    return sc;  // the same value in every lane of the head
  };
  float sc[kSelfWarpMaxLen];
#pragma unroll
  for (int u = 0; u < 16; ++u) sc[u] = u < n_old ? score(kr[u]) : -INFINITY;
  if (len > 16) {
#pragma unroll
    for (int u = 0; u < 16; ++u) kr[u] = 16 + u < n_old ? ld_stream16(kbase + size_t(16 + u) * p.ld) : zero;
#pragma unroll
    for (int u = 0; u < 16; ++u) sc[16 + u] = 16 + u < n_old ? score(kr[u]) : -INFINITY;
  } else {
#pragma unroll
    for (int u = 0; u < 16; ++u) sc[16 + u] = -INFINITY;
  }
  const float s_new = score(fk);
  float mx = s_new;
#pragma unroll
  for (int j = 0; j < kSelfWarpMaxLen; ++j) mx = fmaxf(mx, sc[j]);
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
    unpack8(vr[u], vx);
    sum += pj;
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = fmaf(pj, vx[e], acc[e]);
  }
  if (len > 16) {
#pragma unroll
    for (int u = 0; u < 16; ++u) vr[u] = 16 + u < n_old ? ld_stream16(vbase + size_t(16 + u) * p.ld) : zero;
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const float pj = exp2f(sc[16 + u] - mx);
      float vx[8];
      unpack8(vr[u], vx);
      sum += pj;
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = fmaf(pj, vx[e], acc[e]);
    }
  }
  const float inv = 1.f / sum;
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] *= inv;
  store_bf16x8(p.out + size_t(b) * kD + lane * 8, acc);
}

// ------------------------------------------------------------------------------------------------
// Absorbed cross-attention (MemAttnParams, kernels.h): persistent CTAs (4 warps, six per SM), ONE pass over the memory.
//
// The memory rows m_j (256 bf16 = 512 B) stream through a double-buffered shared-memory ring of 32-row tiles, filled
// by TMA (one elected thread, 128-byte swizzle, mbarrier completion) and running ahead ACROSS question boundaries.  Every tile is consumed once with an online softmax: HBM traffic per question is len * 512 B -
// half of reading a projected K row and a V row - and nothing is re-read.  With NH query vectors per row a CUDA-core
// form needs 8*NH FMAs per 16 loaded bytes and is issue-bound (measured 66 us per 1024 questions vs 47 us for the K|V
// kernel), so both contractions run on warp-level tensor-core MMAs (m16n8k16 bf16, fp32 accumulate):
//   scores  S[j, h] = sum_d M[j, d] q'[h, d]   A = memory tile (ldmatrix), B = absorbed queries (registers, heads padded
//                                              to the 8 MMA columns); warp w takes 16 rows and half of the channels
//   softmax warp h (< NH) owns head h: sums the partial scores of the tile's rows (two per lane), keeps the running
//           max / sum, publishes exp2(S - max) as bf16 and the rescale factor of the accumulators
//   values  U^T[d, h] += sum_j M[j, d] P[h, j]  A = the same tile through ldmatrix.trans (channels on the MMA rows),
//                                              B = published weights (heads on the 8 columns); warp w owns channels
//                                              [32 w, 32 w + 32)
// (A 128-row tcgen05 atom would be 97% padding here: 2-4 query rows per question.)  Rows past the sequence length
// are read as they lie in the memory buffer (finite: the encoder writes all 256 rows) and get weight exactly 0.
// ------------------------------------------------------------------------------------------------
constexpr int kMemTileRows = kMemAttnTileRows;
constexpr int kMemRowGroups = kMemTileRows / 16;            // 16-row MMA groups per tile
constexpr int kMemWarps = 4;                                // small CTAs: six fit an SM (registers and 33 KB of ring each)
constexpr int kMemKSplit = kMemWarps / kMemRowGroups;       // warps sharing a row group split the 256 channels
constexpr int kMemMTiles = kD / kMemWarps / 16;             // 16-channel output tiles per warp in the value product
constexpr int kMemKSteps = 16 / kMemKSplit;                 // 16-channel MMA steps per warp in the score phase
constexpr int kMemCtasPerSm = 6;                            // occupancy the kernel is compiled for (registers, 33 KB of ring each)
constexpr int kMemLaunchCtasPerSm = 4;                      // ... and the persistent CTAs actually launched per SM (see launch_mem_attn)
constexpr int kMemStages = 2;
constexpr int kMemBlockBytes = kMemTileRows * 128;          // one 64-channel column block of a tile (TMA box)
constexpr int kMemStageBytes = 4 * kMemBlockBytes;
constexpr int kMemAttnSmem = kMemStages * kMemStageBytes + 1024;  // + slack to align the ring to the swizzle atom
constexpr int kScPad = kMemTileRows + 8;  // row pitch of the score exchange: heads land in different banks

// byte offset of the 16-byte chunk holding channels [c, c + 8) of row r inside a swizzled tile
__device__ __forceinline__ uint32_t mem_tile_off(int r, int c) {
  return uint32_t((c >> 6) * kMemBlockBytes + r * 128 + ((((c & 63) >> 3) ^ (r & 7)) << 4));
}

template <int NH>
__global__ void __launch_bounds__(kMemWarps * 32, kMemCtasPerSm)
mem_attn_kernel(const __grid_constant__ CUtensorMap tm_mem, const MemAttnParams p) {
  static_assert(NH == 2 || NH == 4, "heads");
  extern __shared__ uint8_t ring_raw[];
  __shared__ float s_part[kMemKSplit][NH][kScPad];   // partial scores of the current tile (channel slices)
  __shared__ __align__(16) __nv_bfloat16 s_p[NH][kMemTileRows];  // softmax weights of the current tile
  __shared__ float s_alpha[NH], s_inv[NH];
  // absorbed queries of the current / next question (row pitch 132 words: heads land in different banks); they arrive
  // with the question's first tile, so no global-load latency sits between two questions
  __shared__ __align__(16) uint32_t s_q[2][NH][132];
  __shared__ __align__(8) uint64_t full_bar[kMemStages];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, q4 = lane & 3;
  const int rg = warp % kMemRowGroups;  // 16-row group of the tile this warp scores ...
  const int kh = warp / kMemRowGroups;  // ... over channels [kh * 16 * kMemKSteps, +16 * kMemKSteps)
  const uint32_t ring_u32 = (smem_u32(ring_raw) + 1023u) & ~1023u;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_mem);
    for (int s = 0; s < kMemStages; ++s) mbar_init(&full_bar[s], 1);
    fence_mbar_init();
  }
  pdl_launch_dependents();
  __syncthreads();
  pdl_wait();
  auto len_of = [&](int q) {
    const int l = p.lens ? p.lens[q] : p.const_len;
    return l > kLP ? kLP : (l < 1 ? 1 : l);  // producer and consumer must agree on >= 1 tile per question
  };

  // Producer cursor: the CTA walks questions blockIdx.x, + gridDim.x, ... and the tiles inside each; loads run
  // kMemStages - 1 tiles ahead of the consumers across question boundaries.  EVERY thread keeps the cursor (it only
  // ever changes in CTA-uniform control flow, so the compiler keeps it in uniform registers) and one elected lane of
  // warp 0 issues: from a single thread's vector registers every TMA / bulk-copy costs an elect / R2UR.BROADCAST loop
  // of ~100 cycles, four to eight of them per tile on the warp the whole CTA waits for at the next barrier.
  int pq = blockIdx.x, pt = 0, p_len = 0, issued = 0, p_questions = 0;
  const bool p_el = warp == 0 ? elect_one() : false;
  auto issue_next = [&]() {
    if (pq < p.B) {
      if (pt == 0) p_len = __shfl_sync(0xffffffffu, len_of(pq), 0);  // (tells the compiler the value is warp-uniform)
      const int st = issued % kMemStages;
      const uint32_t dst = ring_u32 + st * kMemStageBytes;
      const int row = pq * int(p.rows_per_q) + pt * kMemTileRows;
      const uint64_t pol = p.l2_evict_first ? kL2EvictFirst : kL2EvictNormal;  // MemAttnParams: the stream must not thrash L2
      if (warp == 0 && p_el) {
        mbar_expect_tx(&full_bar[st], kMemStageBytes + (pt == 0 ? NH * kD * 2 : 0));
        if (pt == 0) {
#pragma unroll
          for (int h = 0; h < NH; ++h)
            bulk_load_u32(smem_u32(&s_q[p_questions & 1][h][0]), p.qp + (size_t(pq) * NH + h) * kD, kD * 2,
                          &full_bar[st]);
        }
#pragma unroll
        for (int cb = 0; cb < 4; ++cb)
          tma_load_2d_u32_hint(&tm_mem, &full_bar[st], dst + cb * kMemBlockBytes, cb * 64, row, pol);
      }
      __syncwarp();
      if (pt == 0) ++p_questions;
      if ((++pt) * kMemTileRows >= p_len) {
        pt = 0;
        pq += gridDim.x;
      }
    }
    ++issued;
  };
  for (int i = 0; i < kMemStages - 1; ++i) issue_next();
  const float sl2 = rsqrtf(float(kD / NH)) * 1.4426950408889634f;  // 1/sqrt(dh) * log2(e): softmax via exp2
  int consumed = 0, questions = 0;

  for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
    const int len = len_of(b);
    const int n_tiles = (len + kMemTileRows - 1) / kMemTileRows;
    // absorbed queries as B fragments of this warp's channel slice: bq[k] = q'[head g][16 ks + 2 q4 + {0,1}],
    // [.. + 8 + {0,1}] with ks = kMemKSteps kh + k; heads >= NH are zero columns.  Filled from s_q at the first tile.
    uint32_t bq[kMemKSteps][2];
    float m_run = -INFINITY, l_run = 0.f;  // softmax warps: running max (warp-uniform) and this lane's share of the sum
    // acc[mt] = U^T[channels 32 w + 16 mt + g (+8)][heads 2 q4, 2 q4 + 1]: the value product is computed transposed
    // (channels on the 16 MMA rows, heads on the 8 columns) - half the MMAs of heads-on-rows
    float acc[kMemMTiles][4];
#pragma unroll
    for (int mt = 0; mt < kMemMTiles; ++mt)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[mt][e] = 0.f;

    for (int i = 0; i < n_tiles; ++i) {
      __syncthreads();  // everyone is done with the previous tile: its stage, s_part, s_p and s_alpha are free
      issue_next();
      const int st = consumed % kMemStages;
      mbar_wait(&full_bar[st], uint32_t(consumed / kMemStages) & 1u);
      const uint32_t tile_u32 = ring_u32 + st * kMemStageBytes;
      ++consumed;
      if (i == 0) {
        const uint32_t* qrow = &s_q[questions & 1][g < NH ? g : 0][kh * kMemKSteps * 8];
#pragma unroll
        for (int k = 0; k < kMemKSteps; ++k) {
          bq[k][0] = g < NH ? qrow[k * 8 + q4] : 0u;
          bq[k][1] = g < NH ? qrow[k * 8 + q4 + 4] : 0u;
        }
        ++questions;
      }
      {
        // partial scores of this warp's 16 rows over its channel slice
        const int ar = rg * 16 + (lane & 15);
        const int ac = kh * kMemKSteps * 16 + (lane >> 4) * 8;
        float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < kMemKSteps; k += 2) {
          uint32_t a0[4], a1[4];
          ldmatrix_x4(tile_u32 + mem_tile_off(ar, ac + k * 16), a0);
          ldmatrix_x4(tile_u32 + mem_tile_off(ar, ac + k * 16 + 16), a1);
          mma_bf16_16816(c0, a0, bq[k][0], bq[k][1]);
          mma_bf16_16816(c1, a1, bq[k + 1][0], bq[k + 1][1]);
        }
        // c[0,1] = S[row g][heads 2 q4, 2 q4 + 1], c[2,3] = the same heads of row g + 8
        if (2 * q4 < NH) {
          const int r = rg * 16 + g;
          s_part[kh][2 * q4][r] = c0[0] + c1[0];
          s_part[kh][2 * q4 + 1][r] = c0[1] + c1[1];
          s_part[kh][2 * q4][r + 8] = c0[2] + c1[2];
          s_part[kh][2 * q4 + 1][r + 8] = c0[3] + c1[3];
        }
      }
      __syncthreads();
      if (warp < NH) {  // online softmax of head `warp`: lane = rows lane, lane + 32, ... of the tile
        float sv[kMemTileRows / 32];
        float mt = -INFINITY;
#pragma unroll
        for (int rr = 0; rr < kMemTileRows / 32; ++rr) {
          const int r = rr * 32 + lane;
          float v = s_part[0][warp][r];
#pragma unroll
          for (int kp = 1; kp < kMemKSplit; ++kp) v += s_part[kp][warp][r];
          sv[rr] = (i * kMemTileRows + r < len) ? v * sl2 : -INFINITY;
          mt = fmaxf(mt, sv[rr]);
        }
        const float m_new = fmaxf(m_run, warp_max(mt));  // finite: every tile holds at least one valid row
        const float alpha = exp2f(m_run - m_new);
        m_run = m_new;
        l_run *= alpha;
#pragma unroll
        for (int rr = 0; rr < kMemTileRows / 32; ++rr) {
          const float pe = exp2f(sv[rr] - m_new);
          l_run += pe;
          s_p[warp][rr * 32 + lane] = __float2bfloat16(pe);
        }
        if (lane == 0) s_alpha[warp] = alpha;
      }
      __syncthreads();
      // B fragments of P^T: pb[ks][0] = P[head g][16 ks + 2 q4 + {0,1}], pb[ks][1] = the same + 8; heads >= NH are zero
      uint32_t pb[kMemRowGroups][2];
#pragma unroll
      for (int ks = 0; ks < kMemRowGroups; ++ks) {
        const uint32_t* pr = reinterpret_cast<const uint32_t*>(&s_p[g < NH ? g : 0][ks * 16 + 2 * q4]);
        pb[ks][0] = g < NH ? pr[0] : 0u;
        pb[ks][1] = g < NH ? pr[4] : 0u;
      }
      {
        const float al0 = s_alpha[2 * q4 < NH ? 2 * q4 : 0], al1 = s_alpha[2 * q4 + 1 < NH ? 2 * q4 + 1 : 0];
#pragma unroll
        for (int mt = 0; mt < kMemMTiles; ++mt) {
          acc[mt][0] *= al0;
          acc[mt][1] *= al1;
          acc[mt][2] *= al0;
          acc[mt][3] *= al1;
        }
      }
      // A = M^T through ldmatrix.trans: matrix m = lane / 8 -> memory rows (m >> 1) * 8 + lane % 8,
      // channels 32 w + 16 mt + (m & 1) * 8
      const int tr = ((lane >> 4) & 1) * 8 + (lane & 7);
      const int tc = warp * (kD / kMemWarps) + ((lane >> 3) & 1) * 8;
#pragma unroll
      for (int ks = 0; ks < kMemRowGroups; ++ks) {
#pragma unroll
        for (int mt = 0; mt < kMemMTiles; ++mt) {
          uint32_t am[4];
          ldmatrix_x4_trans(tile_u32 + mem_tile_off(ks * 16 + tr, tc + mt * 16), am);
          mma_bf16_16816(acc[mt], am, pb[ks][0], pb[ks][1]);
        }
      }
    }
    // normalise and store: this lane holds heads 2 q4, 2 q4 + 1 of channels 32 w + 16 mt + g and + 8
    if (warp < NH) {
      const float l = warp_sum(l_run);
      if (lane == 0) s_inv[warp] = 1.f / l;
    }
    __syncthreads();
    if (2 * q4 < NH) {
      const float inv0 = s_inv[2 * q4], inv1 = s_inv[2 * q4 + 1];
      __nv_bfloat16* o0 = p.out + (size_t(b) * NH + 2 * q4) * kD + warp * (kD / kMemWarps) + g;
      __nv_bfloat16* o1 = o0 + kD;
#pragma unroll
      for (int mt = 0; mt < kMemMTiles; ++mt) {
        o0[mt * 16] = __float2bfloat16(acc[mt][0] * inv0);
        o1[mt * 16] = __float2bfloat16(acc[mt][1] * inv1);
        o0[mt * 16 + 8] = __float2bfloat16(acc[mt][2] * inv0);
        o1[mt * 16 + 8] = __float2bfloat16(acc[mt][3] * inv1);
      }
    }
  }
}

__global__ void publish_tokens_kernel(const PublishParams p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.B * p.n_cols) return;
  const int b = i / p.n_cols, j = i % p.n_cols;
  const long long t = p.tok[size_t(b) * p.tok_ld + p.src_col0 + j];
  if (p.out_i64) p.out_i64[size_t(b) * p.out_ld + j] = t;
  if (p.out_i32) {
    if (p.n_steps && p.step >= p.n_steps[b]) return;
    const long long val = (j == 0 || !p.forced) ? t : p.forced[size_t(b) * p.forced_ld + j - 1];
    p.out_i32[size_t(b) * p.out_ld + j] = int(val);
  }
}

inline int ceil_div(long long a, long long b) { return int((a + b - 1) / b); }

}  // namespace

cudaError_t launch_dec_embed_start(const DecEmbedParams& p, cudaStream_t stream) {
  dec_embed_start_kernel<<<ceil_div((long long)p.B * 32, 256), 256, 0, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_row_attn(const RowAttnParams& p, cudaStream_t stream) {
  const int dh = kD / p.nhead;
  // decoder self-attention of the first 32 positions: one warp per question
  if (p.warp_form && p.new_k && p.new_v && !p.lens && p.const_len >= 1 && p.const_len <= kSelfWarpMaxLen &&
      p.append_pos == p.const_len - 1 && p.k_app == p.k && p.v_app == p.v && p.B > 0) {
    const dim3 grid((p.B * 32 + 127) / 128);
    if (dh == 64) return launch_kernel(self_attn_warp_kernel<64>, grid, dim3(128), 0, stream, p.pdl, p);
    if (dh == 128) return launch_kernel(self_attn_warp_kernel<128>, grid, dim3(128), 0, stream, p.pdl, p);
  }
  if (dh == 64) return launch_kernel(row_attn_kernel<64>, dim3(p.B), dim3(kAttnWarps * 32), 0, stream, p.pdl, p);
  if (dh == 128) return launch_kernel(row_attn_kernel<128>, dim3(p.B), dim3(kAttnWarps * 32), 0, stream, p.pdl, p);
  return cudaErrorInvalidValue;
}

cudaError_t launch_mem_attn(const CUtensorMap& tm_mem, const MemAttnParams& p, cudaStream_t stream) {
  if (p.B <= 0) return cudaSuccess;
  int num_sms = 0;
  cudaError_t e = current_device_sms(&num_sms);
  if (e != cudaSuccess) return e;
  // persistent CTAs, each streaming its questions back to back.  FOUR per SM although six would fit: 592 concurrent
  // streams reach 0.90 of the measured copy peak at 4096 questions where 888 reach 0.78 (0.79 / 0.73 at 2048, 0.69 / 0.65
  // at 1024; 3 and 5 per SM, one CTA per question and a three-stage ring all measure in between or below:
  // tools/microbench_mem_attn.py, profiles/r2_microbench_mem_attn_grid.txt) - fewer, longer streams
  int per_sm = kMemLaunchCtasPerSm;
  if (const char* g = getenv("B200VQA_MEM_ATTN_CTAS_PER_SM")) per_sm = std::min(kMemCtasPerSm, std::max(1, atoi(g)));  // A/B runs
  // ... and every CTA the same number of questions: 1024 questions on 592 CTAs would leave 432 CTAs with two questions
  // and 160 with one (the launch lasts as long as two), 512 CTAs get exactly two each
  const int max_ctas = per_sm * num_sms;
  const int per_cta = (p.B + max_ctas - 1) / max_ctas;
  int grid = (p.B + per_cta - 1) / per_cta;
  if (getenv("B200VQA_MEM_ATTN_UNBALANCED")) grid = p.B < max_ctas ? p.B : max_ctas;  // A/B runs
  e = ensure_dyn_smem(reinterpret_cast<const void*>(mem_attn_kernel<4>), kMemAttnSmem);
  if (e == cudaSuccess) e = ensure_dyn_smem(reinterpret_cast<const void*>(mem_attn_kernel<2>), kMemAttnSmem);
  if (e != cudaSuccess) return e;
  if (p.nhead == 4)
    return launch_kernel(mem_attn_kernel<4>, dim3(grid), dim3(kMemWarps * 32), kMemAttnSmem, stream, p.pdl, tm_mem, p);
  if (p.nhead == 2)
    return launch_kernel(mem_attn_kernel<2>, dim3(grid), dim3(kMemWarps * 32), kMemAttnSmem, stream, p.pdl, tm_mem, p);
  return cudaErrorInvalidValue;
}

cudaError_t launch_publish_tokens(const PublishParams& p, cudaStream_t stream) {
  const long long n = (long long)p.B * p.n_cols;
  if (n <= 0) return cudaSuccess;
  publish_tokens_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace b200vqa
