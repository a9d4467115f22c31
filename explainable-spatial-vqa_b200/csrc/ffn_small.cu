// Feed-forward block for the decode phase (M = batch rows, one position per question):
//     y = LayerNorm(x + W2 relu(W1 x + b1) + b2)                  (torch TransformerDecoderLayer._ff_block + norm3)
//
// With M <= a few thousand rows a plain GEMM pair leaves the chip idle: linear2 has N = 256 (one n-tile) and
// K = ff, so only M/128 CTAs would each stream the whole 1 MB W2 through one SM.  Here the hidden dimension is
// split instead: CTA (m-tile, slice) computes, entirely on chip,
//     H_slice = relu(X[128x256] . W1[slice]^T + b1[slice])        tcgen05, fp32 accum in TMEM -> bf16 in smem
//     P_slice = H_slice[128x128] . W2[:, slice]^T                 tcgen05, A operand = the smem tile just written
// and stores the fp32 partial P_slice [128 x 256]; ffn_reduce_ln_kernel then sums the ff/128 partials of a row
// in a fixed order (deterministic), adds b2 + residual and applies LayerNorm.  The hidden activations never touch
// HBM, every SM pulls 1/16 of the weights, and the two launches replace linear1 / linear2.
#include "kernels.h"
#include "ptx.cuh"

namespace b200vqa {
namespace {

constexpr int kFfnThreads = 192;
constexpr int kSlice = 128;  // hidden units per CTA

struct FfnSmem {
  static constexpr int kX = 4 * 16384;    // X tile: 4 k-blocks of [128 rows x 64]
  static constexpr int kW1 = 4 * 16384;   // W1 slice: 4 k-blocks of [128 hidden x 64]
  static constexpr int kW2 = 2 * 32768;   // W2 slice: 2 k-blocks of [256 out x 64 hidden]
  static constexpr int kH = 2 * 16384;    // H slice: 2 panels of [128 rows x 64 hidden]
  static constexpr int kOffX = 0;
  static constexpr int kOffW1 = kOffX + kX;
  static constexpr int kOffW2 = kOffW1 + kW1;
  static constexpr int kOffH = kOffX;  // H is written after the first GEMM has retired: it reuses X's shared memory
  static constexpr int kOffBar = kOffW2 + kW2;
  static constexpr int kOffB1 = kOffBar + 64;   // this slice's 128 linear1 biases (fp32)
  static constexpr int kBytes = kOffB1 + 512;   // 192.6 KB: leaves room for row-attention blocks of other branches
};

__global__ void __launch_bounds__(kFfnThreads, 1)
ffn_partial_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w1,
                   const __grid_constant__ CUtensorMap tm_w2, const FfnSmallParams p) {
  using L = FfnSmem;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sX = smem + L::kOffX;
  uint8_t* sW1 = smem + L::kOffW1;
  uint8_t* sW2 = smem + L::kOffW2;
  uint8_t* sH = smem + L::kOffH;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kOffBar);
  uint64_t* bar_in1 = bars + 0;  // X + W1 slice landed
  uint64_t* bar_in2 = bars + 1;  // W2 slice landed
  uint64_t* bar_d1 = bars + 2;   // H accumulator complete
  uint64_t* bar_h = bars + 3;    // H (bf16) written to smem by the 128 epilogue threads
  uint64_t* bar_d2 = bars + 4;   // partial accumulator complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 128;
  const int slice = blockIdx.y;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_w1);
    tma_prefetch_desc(&tm_w2);
    mbar_init(bar_in1, 1);
    mbar_init(bar_in2, 1);
    mbar_init(bar_d1, 1);
    mbar_init(bar_h, 128);
    mbar_init(bar_d2, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  pdl_launch_dependents();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t d1 = tmem_base;        // 128 columns
  const uint32_t d2 = tmem_base + 128;  // 256 columns

  if (warp == 0) {
    // (whole warp, one elected lane issues: the TMA operands stay in uniform registers - see gemm_tc_kernel's producer)
    const bool el = elect_one();
    if (el) {
      // weights first (independent of the previous kernel), activations after the dependency wait
      mbar_expect_tx(bar_in1, L::kX + L::kW1);
#pragma unroll
      for (int kb = 0; kb < 4; ++kb) tma_load_2d(&tm_w1, bar_in1, sW1 + kb * 16384, kb * 64, slice * kSlice);
      mbar_expect_tx(bar_in2, L::kW2);
#pragma unroll
      for (int kb = 0; kb < 2; ++kb) tma_load_2d(&tm_w2, bar_in2, sW2 + kb * 32768, slice * kSlice + kb * 64, 0);
    }
    __syncwarp();
    pdl_wait();
    if (el) {
#pragma unroll
      for (int kb = 0; kb < 4; ++kb) tma_load_2d(&tm_x, bar_in1, sX + kb * 16384, kb * 64, m0);
    }
    __syncwarp();
  } else if (warp == 1) {
    // The whole warp waits and one elected lane issues: the descriptors stay in uniform registers (see gemm_tc_kernel)
    const bool el = elect_one();
    const uint64_t dx0 = make_smem_desc_sw128(smem_u32(sX), 16, 1024);
    const uint64_t dw1 = make_smem_desc_sw128(smem_u32(sW1), 16, 1024);
    const uint64_t dh0 = make_smem_desc_sw128(smem_u32(sH), 16, 1024);
    const uint64_t dw2 = make_smem_desc_sw128(smem_u32(sW2), 16, 1024);
    // ---- H = X . W1_slice^T : M=128, N=128, K=256
    mbar_wait(bar_in1, 0);
    tc_fence_after_sync();
    if (el) {
      constexpr uint32_t idesc1 = make_idesc(kFmtBF16, 128, 128, 0, 0);
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const uint64_t off = uint64_t(((k / 4) * 16384 + (k % 4) * 32) >> 4);
        umma_bf16(d1, dx0 + off, dw1 + off, idesc1, k != 0);
      }
      umma_commit(bar_d1);
    }
    __syncwarp();
    // ---- P = H . W2_slice^T : M=128, N=256, K=128
    mbar_wait(bar_h, 0);
    mbar_wait(bar_in2, 0);
    tc_fence_after_sync();
    if (el) {
      constexpr uint32_t idesc2 = make_idesc(kFmtBF16, 128, 256, 0, 0);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        umma_bf16(d2, dh0 + uint64_t(((k / 4) * 16384 + (k % 4) * 32) >> 4),
                  dw2 + uint64_t(((k / 4) * 32768 + (k % 4) * 32) >> 4), idesc2, k != 0);
      umma_commit(bar_d2);
    }
    __syncwarp();
  } else {
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;  // row inside the tile == TMEM lane
    const uint32_t lane_off = uint32_t(quarter * 32) << 16;
    // the slice's biases are parked in shared memory while the first GEMM runs (weights: safe before the dependency
    // wait), so no L2 round trip sits between the accumulator and H
    float* s_b1 = reinterpret_cast<float*>(smem + L::kOffB1);
    s_b1[r] = __ldg(p.b1 + slice * kSlice + r);
    named_bar_sync(1, 128);
    pdl_wait();  // the partial-sum buffer is still being read by the previous reduce kernel until here
    mbar_wait(bar_d1, 0);
    __syncwarp();
    tc_fence_after_sync();
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t v[32];
      tmem_ld32(d1 + lane_off + c * 32, v);
      tmem_ld_wait();
      uint32_t o[16];
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b4 = *reinterpret_cast<const float4*>(s_b1 + c * 32 + j);
        o[j >> 1] = pack_bf16x2(fmaxf(__uint_as_float(v[j]) + b4.x, 0.f), fmaxf(__uint_as_float(v[j + 1]) + b4.y, 0.f));
        o[(j >> 1) + 1] =
            pack_bf16x2(fmaxf(__uint_as_float(v[j + 2]) + b4.z, 0.f), fmaxf(__uint_as_float(v[j + 3]) + b4.w, 0.f));
      }
      // K-major, 128-byte swizzled A operand: panel = 64 hidden units, 16-byte chunk index XOR (row & 7)
      uint8_t* prow = sH + (c >> 1) * 16384 + r * 128;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int chunk = (c & 1) * 4 + q;
        *reinterpret_cast<uint4*>(prow + ((chunk ^ (r & 7)) << 4)) =
            make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
      }
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    mbar_arrive(bar_h);

    mbar_wait(bar_d2, 0);
    __syncwarp();
    tc_fence_after_sync();
    // fp32 partial tile, 32 rows x 32 columns per warp and chunk: transposed through a 4 KB shared tile (W1's shared
    // memory, free since the first GEMM retired; 16-byte units XOR-swizzled with row & 7: conflict-free both ways) so
    // that every store instruction covers 4 rows x 128 contiguous bytes instead of 32 rows x 16 bytes
    uint8_t* stg = sW1 + quarter * 4096;
    float* pbase = p.partial + (size_t(slice) * p.M + m0 + quarter * 32) * kD;
#pragma unroll 1
    for (int c = 0; c < 8; ++c) {
      uint32_t v[32];
      tmem_ld32(d2 + lane_off + c * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 8; ++q)
        *reinterpret_cast<uint4*>(stg + lane * 128 + ((q ^ (lane & 7)) << 4)) =
            make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int rr = (lane >> 3) + 4 * k, pc = lane & 7;
        const uint4 val = *reinterpret_cast<const uint4*>(stg + rr * 128 + ((pc ^ (rr & 7)) << 4));
        if (m0 + quarter * 32 + rr < p.M) *reinterpret_cast<uint4*>(pbase + size_t(rr) * kD + c * 32 + pc * 4) = val;
      }
      __syncwarp();
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

// One warp per row: y = LayerNorm(sum_s partial[s] + b2 + residual); lane owns 8 consecutive columns.
__global__ void __launch_bounds__(256) ffn_reduce_ln_kernel(const FfnSmallParams p) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();
  pdl_wait();
  if (row >= p.M) return;
  float v[8];
  {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p.b2 + lane * 8));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p.b2 + lane * 8) + 1);
    const uint4 rv = *reinterpret_cast<const uint4*>(p.residual + size_t(row) * kD + lane * 8);
    const float2 r0 = unpack_bf16x2(rv.x), r1 = unpack_bf16x2(rv.y), r2 = unpack_bf16x2(rv.z), r3 = unpack_bf16x2(rv.w);
    v[0] = a.x + r0.x; v[1] = a.y + r0.y; v[2] = a.z + r1.x; v[3] = a.w + r1.y;
    v[4] = b.x + r2.x; v[5] = b.y + r2.y; v[6] = b.z + r3.x; v[7] = b.w + r3.y;
  }
  const float* src = p.partial + size_t(row) * kD + lane * 8;
  const size_t stride = size_t(p.M) * kD;
  // all partials of up to 16 slices are requested before the first add (one L2 round trip, not four); the sum order
  // stays slice 0, 1, 2, ... (deterministic)
  for (int s0 = 0; s0 < p.n_slices; s0 += 16) {
    float4 a[16], b[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      if (s0 + u < p.n_slices) {
        a[u] = *reinterpret_cast<const float4*>(src + (s0 + u) * stride);
        b[u] = *reinterpret_cast<const float4*>(src + (s0 + u) * stride + 4);
      }
    }
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      if (s0 + u < p.n_slices) {
        v[0] += a[u].x; v[1] += a[u].y; v[2] += a[u].z; v[3] += a[u].w;
        v[4] += b[u].x; v[5] += b[u].y; v[6] += b[u].z; v[7] += b[u].w;
      }
    }
  }
  float s1 = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) s1 += v[j];
  const float mean = warp_sum(s1) * (1.f / kD);
  float s2 = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) s2 += (v[j] - mean) * (v[j] - mean);
  const float rstd = rsqrtf(warp_sum(s2) * (1.f / kD) + p.eps);
  const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.gamma + lane * 8));
  const float4 g1 = __ldg(reinterpret_cast<const float4*>(p.gamma + lane * 8) + 1);
  const float4 t0 = __ldg(reinterpret_cast<const float4*>(p.beta + lane * 8));
  const float4 t1 = __ldg(reinterpret_cast<const float4*>(p.beta + lane * 8) + 1);
  const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
  const float t[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
  float y[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) y[j] = (v[j] - mean) * rstd * g[j] + t[j];
  *reinterpret_cast<uint4*>(p.out + size_t(row) * kD + lane * 8) =
      make_uint4(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]), pack_bf16x2(y[4], y[5]), pack_bf16x2(y[6], y[7]));
  if (p.out_f32) {
    if (p.fn_gamma) {  // final decoder norm on top of norm3 (fp32, feeds the vocabulary head)
      float a1 = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) a1 += y[j];
      const float m2 = warp_sum(a1) * (1.f / kD);
      float a2 = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) a2 += (y[j] - m2) * (y[j] - m2);
      const float r2 = rsqrtf(warp_sum(a2) * (1.f / kD) + p.eps);
      const float4 fg0 = __ldg(reinterpret_cast<const float4*>(p.fn_gamma + lane * 8));
      const float4 fg1 = __ldg(reinterpret_cast<const float4*>(p.fn_gamma + lane * 8) + 1);
      const float4 fb0 = __ldg(reinterpret_cast<const float4*>(p.fn_beta + lane * 8));
      const float4 fb1 = __ldg(reinterpret_cast<const float4*>(p.fn_beta + lane * 8) + 1);
      const float fg[8] = {fg0.x, fg0.y, fg0.z, fg0.w, fg1.x, fg1.y, fg1.z, fg1.w};
      const float fb[8] = {fb0.x, fb0.y, fb0.z, fb0.w, fb1.x, fb1.y, fb1.z, fb1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) y[j] = (y[j] - m2) * r2 * fg[j] + fb[j];
    }
    float4* dst = reinterpret_cast<float4*>(p.out_f32 + size_t(row) * kD + lane * 8);
    dst[0] = make_float4(y[0], y[1], y[2], y[3]);
    dst[1] = make_float4(y[4], y[5], y[6], y[7]);
  }
  if (p.head_w) {
    // vocabulary head of this row: lane l ends up with the logits of tokens l and 32 + l
    float lg[2];
#pragma unroll
    for (int grp = 0; grp < 2; ++grp) {
      float x[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int tokv = grp * 32 + i;
        float acc = 0.f;
        if (tokv < p.head_V) {  // warp-uniform
          const float4 w0 = __ldg(reinterpret_cast<const float4*>(p.head_w + size_t(tokv) * kD + lane * 8));
          const float4 w1 = __ldg(reinterpret_cast<const float4*>(p.head_w + size_t(tokv) * kD + lane * 8) + 1);
          acc = y[0] * w0.x;
          acc = fmaf(y[1], w0.y, acc); acc = fmaf(y[2], w0.z, acc); acc = fmaf(y[3], w0.w, acc);
          acc = fmaf(y[4], w1.x, acc); acc = fmaf(y[5], w1.y, acc); acc = fmaf(y[6], w1.z, acc);
          acc = fmaf(y[7], w1.w, acc);
        }
        x[i] = acc;
      }
      // transposed warp reduction (31 shuffles for 32 sums): at distance `off` a lane keeps the half of the remaining
      // values selected by that bit of its lane index and sends the other half; lane l ends with the total of x[l]
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
      lg[grp] = tokv < p.head_V ? x[0] + __ldg(p.head_b + tokv) : -INFINITY;
      if (p.logits && tokv < p.head_V)
        p.logits[(size_t(row) * p.logits_T + p.head_t) * p.head_V + tokv] = lg[grp];
    }
    // argmax, first maximum wins (a NaN never wins; an all-NaN row gives token 0)
    float best = -INFINITY;
    int besti = 0x7fffffff;
    if (lg[0] > best) { best = lg[0]; besti = lane; }
    if (lg[1] > best) { best = lg[1]; besti = 32 + lane; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
      if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
    }
    if (besti >= p.head_V) besti = 0;
    if (lane == 0) p.tok[size_t(row) * p.tok_ld + p.head_t + 1] = besti;
    if (p.pe_next) {  // next decoder input x_next[row] = emb[next token] + pe[t+1]
      long long nxt = p.forced ? p.forced[size_t(row) * p.forced_ld + p.head_t] : (long long)besti;
      nxt = nxt < 0 ? 0 : (nxt >= p.vocab ? p.vocab - 1 : nxt);
      const float4 e0 = __ldg(reinterpret_cast<const float4*>(p.emb + size_t(nxt) * kD + lane * 8));
      const float4 e1 = __ldg(reinterpret_cast<const float4*>(p.emb + size_t(nxt) * kD + lane * 8) + 1);
      const float4 q0 = __ldg(reinterpret_cast<const float4*>(p.pe_next + lane * 8));
      const float4 q1 = __ldg(reinterpret_cast<const float4*>(p.pe_next + lane * 8) + 1);
      *reinterpret_cast<uint4*>(p.x_next + size_t(row) * kD + lane * 8) =
          make_uint4(pack_bf16x2(e0.x + q0.x, e0.y + q0.y), pack_bf16x2(e0.z + q0.z, e0.w + q0.w),
                     pack_bf16x2(e1.x + q1.x, e1.y + q1.y), pack_bf16x2(e1.z + q1.z, e1.w + q1.w));
    }
  }
}

}  // namespace

cudaError_t launch_ffn_small(const CUtensorMap& tm_x, const CUtensorMap& tm_w1, const CUtensorMap& tm_w2,
                             const FfnSmallParams& p, cudaStream_t stream) {
  if (p.M <= 0) return cudaSuccess;
  if (p.ff % kSlice != 0 || p.n_slices != p.ff / kSlice) return cudaErrorInvalidValue;
  if (p.head_w && (p.head_V < 1 || p.head_V > 64 || !p.tok || !p.head_b)) return cudaErrorInvalidValue;
  {
    cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void*>(ffn_partial_kernel), FfnSmem::kBytes);
    if (e != cudaSuccess) return e;
  }
  dim3 grid((p.M + 127) / 128, p.n_slices);
  cudaError_t e = launch_kernel(ffn_partial_kernel, grid, dim3(kFfnThreads), FfnSmem::kBytes, stream, p.pdl, tm_x, tm_w1,
                                tm_w2, p);
  if (e != cudaSuccess) return e;
  return launch_kernel(ffn_reduce_ln_kernel, dim3((p.M * 32 + 255) / 256), dim3(256), 0, stream, p.pdl, p);
}

}  // namespace b200vqa
