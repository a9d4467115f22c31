// Decode-phase kernels: one new position per question and step.  These are bandwidth- or latency-bound
// (M = batch rows, one query row per question), so they run on CUDA cores with 16-byte coalesced accesses;
// the dense layers between them go through the tcgen05 GEMM (gemm.cu).
//
//   dec_embed_start   x0 = emb[start] + pe[0]                                   (IQAP:205-216, FA:136-139)
//   row_attn          one query row against len key/value rows, all heads       (IQAP:223-227, FA:141)
//   (the vocabulary head + argmax + next embedding is the kEpiHead epilogue of the tensor-core GEMM, gemm.cu)
//   publish_tokens    library token buffer -> programs / ys / HBM step cache    (IQAP:239, FA:120-121)
#include <algorithm>

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

  if (p.new_k) {  // append this position's key / value to the caches, then attend over them (self-attention)
    if (threadIdx.x < 64) {
      const bool is_v = threadIdx.x >= 32;
      const __nv_bfloat16* src = (is_v ? p.new_v : p.new_k) + size_t(b) * p.ld_new + lane * 8;
      __nv_bfloat16* dst = (is_v ? p.v_app : p.k_app) + (size_t(b) * p.rows_per_q + p.append_pos) * p.ld + lane * 8;
      *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(src);
    }
    __syncthreads();
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
  auto load_rows = [&](const __nv_bfloat16* base, int j0, uint4 (&raw)[kKeysInFlight]) {
#pragma unroll
    for (int u = 0; u < kKeysInFlight; ++u) {
      const int j = j0 + u * kAttnWarps;
      raw[u] = (j < len) ? ld_stream16(base + size_t(j) * p.ld) : make_uint4(0, 0, 0, 0);
    }
  };
  uint4 cur[kKeysInFlight], nxt[kKeysInFlight];
  load_rows(kbase, warp, cur);
  for (int j0 = warp; j0 < len; j0 += kStep) {
    if (j0 + kStep < len) load_rows(kbase, j0 + kStep, nxt);
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
  load_rows(vbase, warp, cur);  // first V batch: in flight across the softmax
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
    if (j0 + kStep < len) load_rows(vbase, j0 + kStep, nxt);
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
// Absorbed cross-attention (MemAttnParams, kernels.h): one CTA (8 warps) per question, ONE pass over the memory.
//
// The memory rows m_j (256 bf16 = 512 B) stream through a 3-stage shared-memory ring of 64-row tiles (cp.async, rows
// past the length zero-filled); every tile is consumed once with an online softmax, so HBM traffic per question is
// len * 512 B - half of reading a projected K row and a V row - and nothing is re-read.  With NH query vectors per
// row a CUDA-core form needs 8*NH FMAs per 16 loaded bytes and is issue-bound (measured 66 us per 1024 questions vs
// 47 us for the K|V kernel), so both contractions run on warp-level tensor-core MMAs (m16n8k16 bf16, fp32 accumulate):
//   scores  S[j, h] = sum_d M[j, d] q'[h, d]   A = memory tile (ldmatrix), B = absorbed queries (registers, heads padded
//                                              to the 8 MMA columns); warp w takes rows 16 (w & 3), half (w >> 2) of d
//   values  U[h, d] += sum_j P[h, j] M[j, d]   A = exp2(S - running max) (heads padded to the 16 MMA rows), B = the same
//                                              tile through ldmatrix.trans; warp w owns output columns [32 w, 32 w + 32)
// Every warp tracks the running max / sum of "its" head redundantly, so no cross-warp merge is needed at the end.
// Shared rows are padded to 528 B so the eight row addresses of an ldmatrix hit different banks.  (A 128-row tcgen05
// atom would be 97% padding here: 2-4 query rows per question.)
// ------------------------------------------------------------------------------------------------
constexpr int kMemTileRows = 64;
constexpr int kMemRowBytes = kD * 2 + 16;
constexpr int kMemStages = 3;
constexpr int kMemStageBytes = kMemTileRows * kMemRowBytes;
constexpr int kMemAttnSmem = kMemStages * kMemStageBytes;
constexpr int kScPad = kMemTileRows + 8;  // row pitch of the score exchange: heads land in different banks

template <int NH>
__global__ void __launch_bounds__(kAttnWarps * 32, 2) mem_attn_kernel(const MemAttnParams p) {
  static_assert(NH == 2 || NH == 4, "heads");
  extern __shared__ __align__(16) uint8_t ring[];
  __shared__ float s_part[2][NH][kScPad];  // partial scores of the current tile (two halves of d)
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, q4 = lane & 3;
  const int kh = warp >> 2;  // half of the 256 channels this warp contracts in the score MMA
  constexpr uint32_t kHeadLanes = (1u << (4 * NH)) - 1u;  // lanes whose fragment row g is a real head
  pdl_launch_dependents();
  pdl_wait();
  int len = p.lens ? p.lens[b] : p.const_len;
  len = len > kLP ? kLP : len;
  const int n_tiles = (len + kMemTileRows - 1) / kMemTileRows;
  const uint8_t* mbase = reinterpret_cast<const uint8_t*>(p.mem + size_t(b) * p.rows_per_q * kD);
  const uint32_t ring_u32 = smem_u32(ring);

  auto issue = [&](int i) {  // tile i -> stage i % kMemStages
    if (i < n_tiles) {
      const uint32_t dst = ring_u32 + (i % kMemStages) * kMemStageBytes;
#pragma unroll
      for (int c = 0; c < kMemTileRows * 32 / (kAttnWarps * 32); ++c) {
        const int chunk = c * (kAttnWarps * 32) + threadIdx.x;  // 32 16-byte chunks per row
        const int r = chunk >> 5, col = chunk & 31;
        const int j = i * kMemTileRows + r;
        cp_async_16(dst + r * kMemRowBytes + col * 16, mbase + size_t(j < len ? j : 0) * (kD * 2) + col * 16,
                    j < len ? 16u : 0u);
      }
    }
    cp_async_commit();  // (possibly empty) group: keeps the group count in step with the tile number
  };
  for (int i = 0; i < kMemStages - 1; ++i) issue(i);

  // absorbed queries as B fragments of this warp's channel half: bq[k] = q'[head g][16 ks + 2 q4 + {0,1}],
  // [.. + 8 + {0,1}] with ks = 8 kh + k; heads >= NH are zero columns
  uint32_t bq[8][2];
  {
    const uint32_t* qrow = reinterpret_cast<const uint32_t*>(p.qp + (size_t(b) * NH + (g < NH ? g : 0)) * kD) + kh * 64;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      bq[k][0] = g < NH ? __ldg(qrow + k * 8 + q4) : 0u;
      bq[k][1] = g < NH ? __ldg(qrow + k * 8 + q4 + 4) : 0u;
    }
  }
  const float sl2 = rsqrtf(float(kD / NH)) * 1.4426950408889634f;  // 1/sqrt(dh) * log2(e): softmax via exp2

  float m_run = -INFINITY, l_run = 0.f;  // running max / (per-lane partial) sum of head g
  float acc[4][4];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[nt][e] = 0.f;

  for (int i = 0; i < n_tiles; ++i) {
    cp_async_wait<kMemStages - 2>();
    __syncthreads();  // tile i has landed for every thread; stage (i-1) % S and s_part are free again
    issue(i + kMemStages - 1);
    const uint32_t tile_u32 = ring_u32 + (i % kMemStages) * kMemStageBytes;
    {
      // partial scores of rows [16 (w & 3), +16) over channels [128 kh, +128)
      const uint32_t a_addr =
          tile_u32 + ((warp & 3) * 16 + (lane & 15)) * kMemRowBytes + (lane >> 4) * 16 + kh * 256;
      float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int k = 0; k < 8; k += 2) {
        uint32_t a0[4], a1[4];
        ldmatrix_x4(a_addr + k * 32, a0);
        ldmatrix_x4(a_addr + k * 32 + 32, a1);
        mma_bf16_16816(c0, a0, bq[k][0], bq[k][1]);
        mma_bf16_16816(c1, a1, bq[k + 1][0], bq[k + 1][1]);
      }
      // c[0,1] = S[row g][heads 2 q4, 2 q4 + 1], c[2,3] = the same heads of row g + 8
      if (2 * q4 < NH) {
        const int r = (warp & 3) * 16 + g;
        s_part[kh][2 * q4][r] = (c0[0] + c1[0]) * sl2;
        s_part[kh][2 * q4 + 1][r] = (c0[1] + c1[1]) * sl2;
        s_part[kh][2 * q4][r + 8] = (c0[2] + c1[2]) * sl2;
        s_part[kh][2 * q4 + 1][r + 8] = (c0[3] + c1[3]) * sl2;
      }
    }
    __syncthreads();
    // softmax weights of head g for the 64 rows, in A-fragment order: pa[ks][0] = rows 16 ks + 2 q4 + {0,1},
    // pa[ks][2] = rows 16 ks + 8 + 2 q4 + {0,1}; fragment rows >= NH (registers 1 and 3) stay zero
    uint32_t pa[kMemTileRows / 16][4];
    if (g < NH) {
      float sv[kMemTileRows / 16][4];
      float mt = -INFINITY;
      const int jbase = i * kMemTileRows;
#pragma unroll
      for (int ks = 0; ks < kMemTileRows / 16; ++ks) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = ks * 16 + 2 * q4 + (e & 1) + (e >> 1) * 8;
          const float s = s_part[0][g][j] + s_part[1][g][j];
          sv[ks][e] = (jbase + j < len) ? s : -INFINITY;
          mt = fmaxf(mt, sv[ks][e]);
        }
      }
      mt = fmaxf(mt, __shfl_xor_sync(kHeadLanes, mt, 1));
      mt = fmaxf(mt, __shfl_xor_sync(kHeadLanes, mt, 2));
      const float m_new = fmaxf(m_run, mt);  // finite: every tile holds at least one valid row
      const float alpha = exp2f(m_run - m_new);
      m_run = m_new;
      float ls = 0.f;
#pragma unroll
      for (int ks = 0; ks < kMemTileRows / 16; ++ks) {
        float pe[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          pe[e] = exp2f(sv[ks][e] - m_new);
          ls += pe[e];
        }
        pa[ks][0] = pack_bf16x2(pe[0], pe[1]);
        pa[ks][1] = 0u;
        pa[ks][2] = pack_bf16x2(pe[2], pe[3]);
        pa[ks][3] = 0u;
      }
      l_run = l_run * alpha + ls;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        acc[nt][0] *= alpha;
        acc[nt][1] *= alpha;
      }
    } else {
#pragma unroll
      for (int ks = 0; ks < kMemTileRows / 16; ++ks) pa[ks][0] = pa[ks][1] = pa[ks][2] = pa[ks][3] = 0u;
    }
    __syncwarp();
    // ldmatrix.trans addresses: matrix m = lane / 8 -> rows (m & 1) * 8 + lane % 8, columns 32 w + pair * 16 + (m >> 1) * 8
    const uint32_t b_addr =
        tile_u32 + (((lane >> 3) & 1) * 8 + (lane & 7)) * kMemRowBytes + (warp * 32 + (lane >> 4) * 8) * 2;
#pragma unroll
    for (int ks = 0; ks < kMemTileRows / 16; ++ks) {
#pragma unroll
      for (int pair = 0; pair < 2; ++pair) {
        uint32_t bm[4];
        ldmatrix_x4_trans(b_addr + ks * 16 * kMemRowBytes + pair * 32, bm);
        mma_bf16_16816(acc[2 * pair], pa[ks], bm[0], bm[1]);
        mma_bf16_16816(acc[2 * pair + 1], pa[ks], bm[2], bm[3]);
      }
    }
  }
  cp_async_wait<0>();
  // acc[nt][0,1] = U[head g][32 w + 8 nt + 2 q4 + {0,1}] (un-normalised); the sum is spread over the quad
  if (g < NH) {
    l_run += __shfl_xor_sync(kHeadLanes, l_run, 1);
    l_run += __shfl_xor_sync(kHeadLanes, l_run, 2);
    const float inv = 1.f / l_run;
    __nv_bfloat16* orow = p.out + (size_t(b) * NH + g) * kD + warp * 32 + 2 * q4;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
      *reinterpret_cast<uint32_t*>(orow + nt * 8) = pack_bf16x2(acc[nt][0] * inv, acc[nt][1] * inv);
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
  if (dh == 64) return launch_kernel(row_attn_kernel<64>, dim3(p.B), dim3(kAttnWarps * 32), 0, stream, p.pdl, p);
  if (dh == 128) return launch_kernel(row_attn_kernel<128>, dim3(p.B), dim3(kAttnWarps * 32), 0, stream, p.pdl, p);
  return cudaErrorInvalidValue;
}

cudaError_t launch_mem_attn(const MemAttnParams& p, cudaStream_t stream) {
  if (p.B <= 0) return cudaSuccess;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(mem_attn_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMemAttnSmem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(mem_attn_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMemAttnSmem);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  if (p.nhead == 4)
    return launch_kernel(mem_attn_kernel<4>, dim3(p.B), dim3(kAttnWarps * 32), kMemAttnSmem, stream, p.pdl, p);
  if (p.nhead == 2)
    return launch_kernel(mem_attn_kernel<2>, dim3(p.B), dim3(kAttnWarps * 32), kMemAttnSmem, stream, p.pdl, p);
  return cudaErrorInvalidValue;
}

cudaError_t launch_publish_tokens(const PublishParams& p, cudaStream_t stream) {
  const long long n = (long long)p.B * p.n_cols;
  if (n <= 0) return cudaSuccess;
  publish_tokens_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace b200vqa
