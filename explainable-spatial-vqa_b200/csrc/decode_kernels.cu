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
constexpr int kMemCtasPerSm = 6;
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

  // Producer cursor (thread 0): the CTA walks questions blockIdx.x, + gridDim.x, ... and the tiles inside each; loads
  // run kMemStages - 1 tiles ahead of the consumers across question boundaries.
  int pq = blockIdx.x, pt = 0, p_len = 0, issued = 0, p_questions = 0;
  auto issue_next = [&]() {
    if (pq < p.B) {
      if (pt == 0) p_len = len_of(pq);
      const int st = issued % kMemStages;
      const uint32_t dst = ring_u32 + st * kMemStageBytes;
      const int row = pq * int(p.rows_per_q) + pt * kMemTileRows;
      mbar_expect_tx(&full_bar[st], kMemStageBytes + (pt == 0 ? NH * kD * 2 : 0));
      if (pt == 0) {
#pragma unroll
        for (int h = 0; h < NH; ++h)
          bulk_load_u32(smem_u32(&s_q[p_questions & 1][h][0]), p.qp + (size_t(pq) * NH + h) * kD, kD * 2,
                        &full_bar[st]);
        ++p_questions;
      }
#pragma unroll
      for (int cb = 0; cb < 4; ++cb) tma_load_2d_u32(&tm_mem, &full_bar[st], dst + cb * kMemBlockBytes, cb * 64, row);
      if ((++pt) * kMemTileRows >= p_len) {
        pt = 0;
        pq += gridDim.x;
      }
    }
    ++issued;
  };
  if (threadIdx.x == 0)
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
      if (threadIdx.x == 0) issue_next();
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

// ------------------------------------------------------------------------------------------------
// Absorbed cross-attention on tcgen05 (same algorithm and MemAttnParams as above; the warp-MMA form is bound by the
// legacy HMMA pipe at ~1070 cycles per 32 rows per SM).  A cluster of two CTAs serves one question, one 128-row half of
// the memory each (TMA, 4 swizzled 64-channel blocks = 64 KB; three CTAs per SM keep ~190 KB of loads in flight):
//   scores  S[j, h]  = sum_d M[j, d] q'[h, d]   UMMA M = 128 memory rows (K-major A = the tile), N = 16 (heads padded;
//                                               rows >= NH of the B operand are never written: garbage there only
//                                               reaches accumulator columns nobody reads), K = 256: 16 instructions
//   softmax thread j owns row j (= TMEM lane): per-head max / sum over the 128 rows through warp shuffles + one
//           4-warp exchange; P[h, j] = exp2(S - max_tile) as bf16 into a K-major swizzled B operand
//   values  U^T[d, h] = sum_j M[j, d] P[h, j]   the same tile as the MN-major A operand (channels on the 128 accumulator
//                                               lanes, two halves), N = 16, K = 128 rows: 2 x 8 instructions
//   combine the two CTAs hold (max_i, sum_i, U_i) of their halves relative to their own maxima; each sends the other
//           CTA's 128 channels and its statistics through distributed shared memory, one cluster barrier, and writes
//           u[b, h, 128 rank + t] = (w_0 U_0 + w_1 U_1) / (w_0 l_0 + w_1 l_1), w_i = exp2(max_i - max).
// Rows past the sequence length are read as they lie in the buffer (finite) and get weight exactly 0; a half with no
// valid row (len <= 128) loads nothing and contributes (max -inf, sum 0, U 0).
// ------------------------------------------------------------------------------------------------
constexpr int kMtRows = 128;
constexpr int kMtThreads = 160;                  // warps 0-3: softmax / combine (TMEM lane quarter = warp), warp 4: TMA + MMA issue
constexpr int kMtBlockBytes = kMtRows * 128;     // one 64-channel column block of the tile
constexpr int kMtTileBytes = 4 * kMtBlockBytes;
constexpr int kMtOffQ = kMtTileBytes;            // 4 x 1 KB: heads 0..7 of the query operand per 64-channel block ...
constexpr int kMtQSbo = 4096;                    // ... "heads 8..15" alias whatever lies 4 KB further (never read back)
constexpr int kMtOffP = kMtOffQ + 4096;          // 2 x 1 KB: heads 0..7 of P per 64-row block
constexpr int kMtPSbo = 2048;
constexpr int kMtOffRecv = kMtOffP + 2048;       // [128] float4: the peer's partial product for this CTA's channels, odd questions;
                                                 // even questions: the unused head rows 4..7 of the query blocks
constexpr int kMtOffMisc = kMtOffRecv + 2048;    // statistics, mbarriers, TMEM slot
constexpr int kMtSmem = kMtOffMisc + 512;        // (the aliased operand rows reach kMtOffQ + 3 KB + 4 KB + 1 KB)
constexpr int kMtCtasPerSm = 3;
static_assert(kMtOffQ + 3 * 1024 + kMtQSbo + 1024 <= kMtSmem && kMtOffP + 1024 + kMtPSbo + 1024 <= kMtSmem, "aliases");

__device__ __forceinline__ long long mt_clock() { return clock64(); }
__device__ __forceinline__ long long mt_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return (long long)t;
}

template <int NH>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kMtThreads, kMtCtasPerSm)
mem_attn_tc_kernel(const __grid_constant__ CUtensorMap tm_mem, const __grid_constant__ CUtensorMap tm_q,
                   const MemAttnParams p) {
  static_assert(NH == 2 || NH == 4, "heads");
  extern __shared__ __align__(1024) uint8_t mt_smem[];
  const uint32_t sbase = smem_u32(mt_smem);
  if ((sbase & 1023u) != 0) __trap();
  float* s_wmax = reinterpret_cast<float*>(mt_smem + kMtOffMisc);  // [warp][4]
  float* s_wsum = s_wmax + 16;                                     // [warp][4]
  float* s_peer = s_wsum + 16;                                     // [parity][0..3 max, 4..7 sum] of the peer's half
  uint64_t* bars = reinterpret_cast<uint64_t*>(mt_smem + kMtOffMisc + 256);
  uint64_t* bar_full = bars + 0;  // [4] 64-channel block of the tile (+ of the queries) landed
  uint64_t* bar_s = bars + 4;     // score accumulator complete
  uint64_t* bar_p = bars + 5;     // P written (128 arrivals)
  uint64_t* bar_u = bars + 6;     // value accumulators complete: tile, queries and P are free again
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();  // which 128-row half of the memory
  const int n_clusters = gridDim.x >> 1;

  if (threadIdx.x == 128) {
    tma_prefetch_desc(&tm_mem);
    tma_prefetch_desc(&tm_q);
    for (int i = 0; i < 4; ++i) mbar_init(&bar_full[i], 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_p, 128);
    mbar_init(bar_u, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<64>(tmem_slot);
  pdl_launch_dependents();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  cluster_arrive_release();  // generation 0: this CTA's shared memory is live (the peer waits before its remote stores)
  const uint32_t tmem = *tmem_slot;
  pdl_wait();

  // rows of this CTA's half of question q that take part (<= 0: none; the half then contributes max -inf, sum 0, U 0)
  auto valid_of = [&](int q) {
    int len = p.lens ? p.lens[q] : p.const_len;
    len = len > kLP ? kLP : (len < 1 ? 1 : len);
    return min(kMtRows, len - int(rank) * kMtRows);
  };
  long long* dbg = p.dbg;
  auto stamp = [&](int q, int k, long long v) {
    if (dbg) dbg[(size_t(q) * 2 + rank) * 16 + k] = v;
  };

  if (warp == 4) {
    // ---- control warp: one thread issues the loads and the MMAs; the single 64 KB stage is refilled for the next
    // question as soon as the value MMAs of the current one have retired, so the combine step overlaps the load
    auto issue_loads = [&](int q) {
      const int row = q * int(p.rows_per_q) + int(rank) * kMtRows;
#pragma unroll
      for (int cb = 0; cb < 4; ++cb) {
        mbar_expect_tx(&bar_full[cb], kMtBlockBytes + NH * 128);
        tma_load_2d_u32(&tm_q, &bar_full[cb], sbase + kMtOffQ + cb * 1024, cb * 64, q * NH);
        tma_load_2d_u32(&tm_mem, &bar_full[cb], sbase + cb * kMtBlockBytes, cb * 64, row);
      }
      stamp(q, 3, mt_clock());
    };
    int b = blockIdx.x >> 1;
    int valid = valid_of(b);
    if (lane == 0 && valid > 0) issue_loads(b);
    uint32_t ph = 0;  // phase of the mbarriers: they complete once per question with valid rows
    while (b < p.B) {
      const int nb = b + n_clusters;
      const int nvalid = nb < p.B ? valid_of(nb) : 0;
      if (lane == 0) {
        if (valid > 0) {
          constexpr uint32_t idesc_s = make_idesc(kFmtBF16, 128, 16, 0, 0);
#pragma unroll
          for (int k = 0; k < kD / 16; ++k) {
            if (k % 4 == 0) {
              mbar_wait(&bar_full[k / 4], ph);
              tc_fence_after_sync();
            }
            umma_bf16(tmem, make_smem_desc_sw128(sbase + (k / 4) * kMtBlockBytes + (k % 4) * 32, 16, 1024),
                      make_smem_desc_sw128(sbase + kMtOffQ + (k / 4) * 1024 + (k % 4) * 32, 16, kMtQSbo), idesc_s,
                      k != 0);
          }
          umma_commit(bar_s);
          stamp(b, 4, mt_clock());
          mbar_wait(bar_p, ph);
          tc_fence_after_sync();
          constexpr uint32_t idesc_v = make_idesc(kFmtBF16, 128, 16, 1, 0);
#pragma unroll
          for (int half = 0; half < 2; ++half)
#pragma unroll
            for (int k = 0; k < kMtRows / 16; ++k)
              umma_bf16(tmem + 16 + half * 16,
                        make_smem_desc_sw128(sbase + half * 2 * kMtBlockBytes + k * 2048, kMtBlockBytes, 1024),
                        make_smem_desc_sw128(sbase + kMtOffP + (k / 4) * 1024 + (k % 4) * 32, 16, kMtPSbo), idesc_v,
                        k != 0);
          umma_commit(bar_u);
          mbar_wait(bar_u, ph);  // the stage, the queries and P are free again
        }
        if (nvalid > 0) issue_loads(nb);
      }
      if (valid > 0) ph ^= 1u;
      __syncwarp();
      // every thread of the cluster arrives once per question (generation 0 = start-up); this warp waits only just
      // before its next arrival, so it never stalls on the softmax warps' exchange
      cluster_wait_acquire();
      cluster_arrive_release();
      b = nb;
      valid = nvalid;
    }
    cluster_wait_acquire();
  } else {
    const int r = threadIdx.x;  // memory row of the half == TMEM lane; later: output channel 128 rank + r
    const uint32_t tlane = tmem + (uint32_t(warp * 32) << 16);
    const float sl2 = rsqrtf(float(kD / NH)) * 1.4426950408889634f;  // 1/sqrt(dh) * log2(e): softmax via exp2
    uint32_t ph = 0, it = 0;
    cluster_wait_acquire();  // generation 0: the peer CTA has started, its shared memory may be written
    for (int b = blockIdx.x >> 1; b < p.B; b += n_clusters, ++it) {
      const int valid = valid_of(b);
      if (r == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        stamp(b, 0, smid);
        stamp(b, 1, mt_globaltimer());
        stamp(b, 2, mt_clock());
        stamp(b, 11, it);
      }
      float m_own[NH], l_own[NH], u_own[4] = {0.f, 0.f, 0.f, 0.f}, u_snd[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        m_own[h] = -INFINITY;
        l_own[h] = 0.f;
      }
      if (valid > 0) {
        mbar_wait(bar_s, ph);
        __syncwarp();
        tc_fence_after_sync();
        if (r == 0) stamp(b, 5, mt_clock());
        uint32_t sv[4];
        tmem_ld4(tlane, sv);
        tmem_ld_wait();
        float v[NH];
#pragma unroll
        for (int h = 0; h < NH; ++h) {
          v[h] = r < valid ? __uint_as_float(sv[h]) * sl2 : -INFINITY;
          const float wm = warp_max(v[h]);
          if (lane == 0) s_wmax[warp * 4 + h] = wm;
        }
        named_bar_sync(1, 128);
        uint8_t* prow = mt_smem + kMtOffP + (r >> 6) * 1024 + (r & 7) * 2;
#pragma unroll
        for (int h = 0; h < NH; ++h) {
          // finite: row 0 of a half with valid rows is valid
          m_own[h] = fmaxf(fmaxf(s_wmax[h], s_wmax[4 + h]), fmaxf(s_wmax[8 + h], s_wmax[12 + h]));
          const __nv_bfloat16 pb = __float2bfloat16(exp2f(v[h] - m_own[h]));
          *reinterpret_cast<__nv_bfloat16*>(prow + h * 128 + ((((r & 63) >> 3) ^ h) << 4)) = pb;
          const float ws = warp_sum(__bfloat162float(pb));  // normalise by what the tensor core will actually sum
          if (lane == 0) s_wsum[warp * 4 + h] = ws;
        }
        fence_proxy_async_smem();
        tc_fence_before_sync();
        mbar_arrive(bar_p);
        if (r == 0) stamp(b, 6, mt_clock());
        named_bar_sync(1, 128);
#pragma unroll
        for (int h = 0; h < NH; ++h) l_own[h] = (s_wsum[h] + s_wsum[4 + h]) + (s_wsum[8 + h] + s_wsum[12 + h]);
        mbar_wait(bar_u, ph);
        __syncwarp();
        tc_fence_after_sync();
        if (r == 0) stamp(b, 7, mt_clock());
        uint32_t a[4], c[4];
        tmem_ld4(tlane + 16 + rank * 16, a);
        tmem_ld4(tlane + 16 + (rank ^ 1u) * 16, c);
        tmem_ld_wait();
        tc_fence_before_sync();  // the next question's MMAs overwrite these columns only after this thread's bar_p arrival
#pragma unroll
        for (int h = 0; h < NH; ++h) {
          u_own[h] = __uint_as_float(a[h]);
          u_snd[h] = __uint_as_float(c[h]);
        }
        ph ^= 1u;
      }
      // exchange buffers alternate with the question's parity: the peer's stores of question it + 2 follow its wait
      // on generation it + 2, which needs this thread's arrival of question it + 1 - after the reads below
      const uint32_t par = it & 1u;
      const uint32_t recv_off = par ? kMtOffRecv + r * 16 : kMtOffQ + (r >> 5) * 1024 + 512 + (r & 31) * 16;
      st_cluster_f32x4(cluster_map_shared(sbase + recv_off, rank ^ 1u), u_snd[0], u_snd[1], u_snd[2], u_snd[3]);
      if (r < NH) {
        const uint32_t peer_stats = cluster_map_shared(smem_u32(s_peer + par * 8), rank ^ 1u);
        st_cluster_f32(peer_stats + r * 4, m_own[r]);
        st_cluster_f32(peer_stats + 16 + r * 4, l_own[r]);
      }
      cluster_arrive_release();
      cluster_wait_acquire();
      if (r == 0) stamp(b, 8, mt_clock());
      const float4 rc = *reinterpret_cast<const float4*>(mt_smem + recv_off);
      const float u_rcv[4] = {rc.x, rc.y, rc.z, rc.w};
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        const float mp = s_peer[par * 8 + h], lp = s_peer[par * 8 + 4 + h];
        const float m = fmaxf(m_own[h], mp);
        const float w0 = exp2f(m_own[h] - m), w1 = exp2f(mp - m);
        const float o = (w0 * u_own[h] + w1 * u_rcv[h]) / (w0 * l_own[h] + w1 * lp);
        p.out[(size_t(b) * NH + h) * kD + rank * kMtRows + r] = __float2bfloat16(o);
      }
      if (r == 0) {
        stamp(b, 9, mt_clock());
        stamp(b, 10, mt_globaltimer());
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc<64>(tmem);
  }
}

// ------------------------------------------------------------------------------------------------
// The same tcgen05 arithmetic as mem_attn_tc_kernel with ONE persistent CTA per SM that owns whole questions: the
// 128-row halves stream through a ring of three 64 KB stages filled by a producer thread that runs ahead across
// question boundaries (two stages are always loading while the third is being consumed), the MMA thread issues the
// score MMAs of tile t + 1 before the value MMAs of tile t (score accumulators and P double-buffered by tile parity),
// and the halves of a question are combined inside the CTA (value accumulators per (question parity, half of the
// memory, half of the channels) in TMEM) - no cluster, no distributed shared memory.
// ------------------------------------------------------------------------------------------------
constexpr int kMrStages = 3;
constexpr int kMrThreads = 192;  // warps 0-3 softmax / combine, warp 4 TMA producer, warp 5 MMA issue
constexpr int kMrOffQ = kMrStages * kMtTileBytes;  // [2 question parities] 4 x 1 KB query blocks (heads 8..15 alias + 4 KB)
constexpr int kMrOffP = kMrOffQ + 2 * 4096;        // [2 tile parities] 2 x 1 KB P blocks (heads 8..15 alias + 2 KB)
constexpr int kMrOffMisc = kMrOffP + 2 * 2048;     // statistics, mbarriers, TMEM slot
constexpr int kMrSmem = kMrOffMisc + 3072;         // the aliased rows of the last P buffer end at kMrOffP + 2048 + 4096
static_assert(kMrOffQ + 4096 + 3 * 1024 + kMtQSbo + 1024 <= kMrSmem && kMrOffP + 2048 + 1024 + kMtPSbo + 1024 <= kMrSmem,
              "aliases");

struct MrTileIter {  // the tiles of this CTA in processing order; every role walks the same sequence
  int q, i, n, g, qi, stride, B;
  const int32_t* lens;
  int const_len;
  __device__ __forceinline__ int tiles_of(int qq) const {
    int len = lens ? lens[qq] : const_len;
    len = len > kLP ? kLP : (len < 1 ? 1 : len);
    return (len + kMtRows - 1) / kMtRows;
  }
  __device__ __forceinline__ int len_of(int qq) const {
    const int len = lens ? lens[qq] : const_len;
    return len > kLP ? kLP : (len < 1 ? 1 : len);
  }
  __device__ __forceinline__ void init(const MemAttnParams& p) {
    q = blockIdx.x; i = 0; g = 0; qi = 0; stride = gridDim.x; B = p.B; lens = p.lens; const_len = p.const_len;
    n = q < B ? tiles_of(q) : 0;
  }
  __device__ __forceinline__ bool done() const { return q >= B; }
  __device__ __forceinline__ void next() {
    ++g;
    if (++i == n) {
      i = 0;
      q += stride;
      ++qi;
      n = q < B ? tiles_of(q) : 0;
    }
  }
};

template <int NH>
__global__ void __launch_bounds__(kMrThreads, 1)
mem_attn_ring_tc_kernel(const __grid_constant__ CUtensorMap tm_mem, const __grid_constant__ CUtensorMap tm_q,
                        const MemAttnParams p) {
  static_assert(NH == 2 || NH == 4, "heads");
  extern __shared__ __align__(1024) uint8_t mr_smem[];
  const uint32_t sbase = smem_u32(mr_smem);
  if ((sbase & 1023u) != 0) __trap();
  float* s_wmax = reinterpret_cast<float*>(mr_smem + kMrOffMisc);  // [warp][4]
  float* s_wsum = s_wmax + 16;                                     // [warp][4]
  uint64_t* bars = reinterpret_cast<uint64_t*>(mr_smem + kMrOffMisc + 256);
  uint64_t* bar_full = bars + 0;    // [3] stage landed (+ the queries with a question's first tile)
  uint64_t* bar_empty = bars + 3;   // [3] the value MMAs that read the stage have retired
  uint64_t* bar_s = bars + 6;       // [2] score accumulator of tile parity complete
  uint64_t* bar_p = bars + 8;       // [2] P of tile parity written (128 arrivals)
  uint64_t* bar_u = bars + 10;      // [2] value accumulators of question parity complete
  uint64_t* bar_ufree = bars + 12;  // [2] ... and read by the combine step (128 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 128) {
    tma_prefetch_desc(&tm_mem);
    tma_prefetch_desc(&tm_q);
    for (int i = 0; i < kMrStages; ++i) {
      mbar_init(&bar_full[i], 1);
      mbar_init(&bar_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_s[i], 1);
      mbar_init(&bar_p[i], 128);
      mbar_init(&bar_u[i], 1);
      mbar_init(&bar_ufree[i], 128);
    }
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<256>(tmem_slot);
  pdl_launch_dependents();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();
  MrTileIter t;
  t.init(p);
  // TMEM columns: scores [tile parity] at 16 * parity; values [question parity][half of the memory][half of the channels]
  auto u_col = [](int qpar, int i, int half) { return uint32_t(32 + ((qpar * 2 + i) * 2 + half) * 16); };

  if (warp == 4) {
    if (lane == 0) {
      for (; !t.done(); t.next()) {
        const int st = t.g % kMrStages;
        mbar_wait(&bar_empty[st], ((t.g / kMrStages) & 1) ^ 1);  // passes at once on a fresh barrier
        mbar_expect_tx(&bar_full[st], kMtTileBytes + (t.i == 0 ? 4 * NH * 128 : 0));
        if (t.i == 0) {
#pragma unroll
          for (int cb = 0; cb < 4; ++cb)
            tma_load_2d_u32(&tm_q, &bar_full[st], sbase + kMrOffQ + (t.qi & 1) * 4096 + cb * 1024, cb * 64, t.q * NH);
        }
        const int row = t.q * int(p.rows_per_q) + t.i * kMtRows;
#pragma unroll
        for (int cb = 0; cb < 4; ++cb)
          tma_load_2d_u32(&tm_mem, &bar_full[st], sbase + st * kMtTileBytes + cb * kMtBlockBytes, cb * 64, row);
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc(kFmtBF16, 128, 16, 0, 0);
      constexpr uint32_t idesc_v = make_idesc(kFmtBF16, 128, 16, 1, 0);
      auto issue_scores = [&](const MrTileIter& x) {
        const int st = x.g % kMrStages;
        mbar_wait(&bar_full[st], (x.g / kMrStages) & 1);
        tc_fence_after_sync();
        const uint32_t tile = sbase + st * kMtTileBytes, qb = sbase + kMrOffQ + (x.qi & 1) * 4096;
#pragma unroll
        for (int k = 0; k < kD / 16; ++k)
          umma_bf16(tmem + (x.g & 1) * 16, make_smem_desc_sw128(tile + (k / 4) * kMtBlockBytes + (k % 4) * 32, 16, 1024),
                    make_smem_desc_sw128(qb + (k / 4) * 1024 + (k % 4) * 32, 16, kMtQSbo), idesc_s, k != 0);
        umma_commit(&bar_s[x.g & 1]);
      };
      if (!t.done()) issue_scores(t);
      while (!t.done()) {
        MrTileIter nx = t;
        nx.next();
        if (!nx.done()) issue_scores(nx);  // S(t + 1) runs under the softmax of tile t
        const int st = t.g % kMrStages, qpar = t.qi & 1;
        mbar_wait(&bar_p[t.g & 1], (t.g >> 1) & 1);
        if (t.i == 0 && t.qi >= 2) mbar_wait(&bar_ufree[qpar], ((t.qi >> 1) - 1) & 1);  // accumulators of question qi - 2 read
        tc_fence_after_sync();
        const uint32_t tile = sbase + st * kMtTileBytes, pb = sbase + kMrOffP + (t.g & 1) * 2048;
#pragma unroll
        for (int half = 0; half < 2; ++half)
#pragma unroll
          for (int k = 0; k < kMtRows / 16; ++k)
            umma_bf16(tmem + u_col(qpar, t.i, half),
                      make_smem_desc_sw128(tile + half * 2 * kMtBlockBytes + k * 2048, kMtBlockBytes, 1024),
                      make_smem_desc_sw128(pb + (k / 4) * 1024 + (k % 4) * 32, 16, kMtPSbo), idesc_v, k != 0);
        umma_commit(&bar_empty[st]);
        if (t.i == t.n - 1) umma_commit(&bar_u[qpar]);
        t = nx;
      }
    }
  } else {
    const int r = threadIdx.x;  // memory row inside a tile == TMEM lane; in the combine step: channels r and 128 + r
    const uint32_t tlane = tmem + (uint32_t(warp * 32) << 16);
    const float sl2 = rsqrtf(float(kD / NH)) * 1.4426950408889634f;  // 1/sqrt(dh) * log2(e): softmax via exp2
    float m_t[2][NH], l_t[2][NH];  // per half of the memory: max and sum of exp2(S - max) of the current question
    for (; !t.done(); t.next()) {
      const int valid = min(kMtRows, t.len_of(t.q) - t.i * kMtRows);
      const int sp = t.g & 1;
      mbar_wait(&bar_s[sp], (t.g >> 1) & 1);  // also: the value MMAs of tile g - 2 (readers of this P buffer) have retired
      __syncwarp();
      tc_fence_after_sync();
      uint32_t sv[4];
      tmem_ld4(tlane + sp * 16, sv);
      tmem_ld_wait();
      float v[NH];
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        v[h] = r < valid ? __uint_as_float(sv[h]) * sl2 : -INFINITY;
        const float wm = warp_max(v[h]);
        if (lane == 0) s_wmax[warp * 4 + h] = wm;
      }
      named_bar_sync(1, 128);
      uint8_t* prow = mr_smem + kMrOffP + sp * 2048 + (r >> 6) * 1024 + (r & 7) * 2;
      float ps[NH];
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        const float m = fmaxf(fmaxf(s_wmax[h], s_wmax[4 + h]), fmaxf(s_wmax[8 + h], s_wmax[12 + h]));  // finite: row 0 is valid
        const __nv_bfloat16 pb = __float2bfloat16(exp2f(v[h] - m));
        *reinterpret_cast<__nv_bfloat16*>(prow + h * 128 + ((((r & 63) >> 3) ^ h) << 4)) = pb;
        ps[h] = __bfloat162float(pb);  // normalise by what the tensor core will actually sum
        if (t.i == 0) m_t[0][h] = m; else m_t[1][h] = m;
      }
      fence_proxy_async_smem();
      tc_fence_before_sync();
      mbar_arrive(&bar_p[sp]);
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        const float ws = warp_sum(ps[h]);
        if (lane == 0) s_wsum[warp * 4 + h] = ws;
      }
      named_bar_sync(1, 128);
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        const float l = (s_wsum[h] + s_wsum[4 + h]) + (s_wsum[8 + h] + s_wsum[12 + h]);
        if (t.i == 0) l_t[0][h] = l; else l_t[1][h] = l;
      }
      if (t.i == t.n - 1) {
        // combine the halves of the question: u = (w_0 U_0 + w_1 U_1) / (w_0 l_0 + w_1 l_1), w_i = exp2(max_i - max)
        const int qpar = t.qi & 1;
        float w0[NH], w1[NH];
#pragma unroll
        for (int h = 0; h < NH; ++h) {
          const float m1 = t.n == 2 ? m_t[1][h] : -INFINITY, l1 = t.n == 2 ? l_t[1][h] : 0.f;
          const float m = fmaxf(m_t[0][h], m1);
          const float a = exp2f(m_t[0][h] - m), b = exp2f(m1 - m);
          const float inv = 1.f / (a * l_t[0][h] + b * l1);
          w0[h] = a * inv;
          w1[h] = b * inv;
        }
        mbar_wait(&bar_u[qpar], (t.qi >> 1) & 1);
        __syncwarp();
        tc_fence_after_sync();
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t a[4], b[4] = {0u, 0u, 0u, 0u};
          tmem_ld4(tlane + u_col(qpar, 0, half), a);
          if (t.n == 2) tmem_ld4(tlane + u_col(qpar, 1, half), b);
          tmem_ld_wait();
#pragma unroll
          for (int h = 0; h < NH; ++h)
            p.out[(size_t(t.q) * NH + h) * kD + half * kMtRows + r] =
                __float2bfloat16(w0[h] * __uint_as_float(a[h]) + (t.n == 2 ? w1[h] * __uint_as_float(b[h]) : 0.f));
        }
        tc_fence_before_sync();
        mbar_arrive(&bar_ufree[qpar]);
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc<256>(tmem);
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
  // persistent: kMemCtasPerSm CTAs per SM, each streaming its questions back to back
  const int grid = p.B < kMemCtasPerSm * num_sms ? p.B : kMemCtasPerSm * num_sms;
  e = ensure_dyn_smem(reinterpret_cast<const void*>(mem_attn_kernel<4>), kMemAttnSmem);
  if (e == cudaSuccess) e = ensure_dyn_smem(reinterpret_cast<const void*>(mem_attn_kernel<2>), kMemAttnSmem);
  if (e != cudaSuccess) return e;
  if (p.nhead == 4)
    return launch_kernel(mem_attn_kernel<4>, dim3(grid), dim3(kMemWarps * 32), kMemAttnSmem, stream, p.pdl, tm_mem, p);
  if (p.nhead == 2)
    return launch_kernel(mem_attn_kernel<2>, dim3(grid), dim3(kMemWarps * 32), kMemAttnSmem, stream, p.pdl, tm_mem, p);
  return cudaErrorInvalidValue;
}

cudaError_t launch_mem_attn_tc(const CUtensorMap& tm_mem, const CUtensorMap& tm_q, const MemAttnParams& p,
                               cudaStream_t stream) {
  if (p.B <= 0) return cudaSuccess;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
      const void* fn = i == 0 ? reinterpret_cast<const void*>(mem_attn_tc_kernel<4>)
                              : reinterpret_cast<const void*>(mem_attn_tc_kernel<2>);
      e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kMtSmem);
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    }
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
  }
  // one cluster of two CTAs per question; persistent (three CTAs per SM) when there are more questions than that
  int clusters = p.B;
  if (p.tc_persistent) clusters = std::min(p.B, kMtCtasPerSm * num_sms / 2);
  if (p.nhead == 4)
    return launch_kernel(mem_attn_tc_kernel<4>, dim3(2 * clusters), dim3(kMtThreads), kMtSmem, stream, p.pdl, tm_mem, tm_q, p);
  if (p.nhead == 2)
    return launch_kernel(mem_attn_tc_kernel<2>, dim3(2 * clusters), dim3(kMtThreads), kMtSmem, stream, p.pdl, tm_mem, tm_q, p);
  return cudaErrorInvalidValue;
}

cudaError_t launch_mem_attn_ring_tc(const CUtensorMap& tm_mem, const CUtensorMap& tm_q, const MemAttnParams& p,
                                    cudaStream_t stream) {
  if (p.B <= 0) return cudaSuccess;
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
  }
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(mem_attn_ring_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMrSmem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(mem_attn_ring_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMrSmem);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  const int grid = std::min(p.B, num_sms);  // persistent: one CTA per SM, questions blockIdx.x, + gridDim.x, ...
  if (p.nhead == 4)
    return launch_kernel(mem_attn_ring_tc_kernel<4>, dim3(grid), dim3(kMrThreads), kMrSmem, stream, p.pdl, tm_mem, tm_q, p);
  if (p.nhead == 2)
    return launch_kernel(mem_attn_ring_tc_kernel<2>, dim3(grid), dim3(kMrThreads), kMrSmem, stream, p.pdl, tm_mem, tm_q, p);
  return cudaErrorInvalidValue;
}

cudaError_t launch_publish_tokens(const PublishParams& p, cudaStream_t stream) {
  const long long n = (long long)p.B * p.n_cols;
  if (n <= 0) return cudaSuccess;
  publish_tokens_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace b200vqa
