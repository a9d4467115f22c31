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

cudaError_t launch_publish_tokens(const PublishParams& p, cudaStream_t stream) {
  const long long n = (long long)p.B * p.n_cols;
  if (n <= 0) return cudaSuccess;
  publish_tokens_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace b200vqa
