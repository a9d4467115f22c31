// Fused encoder self-attention on tcgen05 for the ~200-token image+program sequence.
//
// Replaces the scaled-dot-product attention inside nn.TransformerEncoderLayer.self_attn
// (IQAP:118,173; FA:42,135 -> torch functional.py:6244+): softmax(q k^T / sqrt(dh)) v per head, with a
// key-length mask for the batched FA path (the reference is batch-1 and never pads).
//
// One CTA per (question, head, 128-query tile). The whole K and V of a head (<= 256 rows) sit in shared
// memory, so there is no online-softmax rescaling:
//   warp 0     TMA: Q tile, K and V (128B-swizzled boxes of the packed q|k|v activations)
//   warp 1     tcgen05.mma  S[128 x keys] = Q K^T  into TMEM, later  O[128 x dh] = P V
//   warps 2-5  one query row per thread: tcgen05.ld S, masked softmax in fp32 (exp2), P -> smem as the
//              bf16 K-major A operand of the second MMA, then O / rowsum -> bf16 global
// V is consumed directly as an MN-major B operand (v_mode 0); v_mode 1 transposes it in shared memory
// first (kept as a cross-check of the MN-major descriptor path).
#include "kernels.h"
#include "ptx.cuh"

namespace b200vqa {
namespace {

constexpr int kAttnThreads = 192;
constexpr int kKeysMax = 256;

template <int DH>
struct AttnSmem {
  static constexpr int kPanels = DH / 64;
  static constexpr int kQ = 128 * DH * 2;                 // 16 / 32 KB
  static constexpr int kKP = 65536;                        // K (256 x DH) aliased later by P (128 x 256)
  static constexpr int kV = kKeysMax * DH * 2;             // 32 / 64 KB
  static constexpr int kOffQ = 0;
  static constexpr int kOffKP = kQ;
  static constexpr int kOffV = kOffKP + kKP;
  static constexpr int kOffVt = kOffV + kV;                // only used by v_mode 1
  static constexpr int kBarOffNoVt = kOffVt;
  static constexpr int kBarOffVt = kOffVt + kV;
  static constexpr int bytes(bool vt) { return (vt ? kBarOffVt : kBarOffNoVt) + 128; }
};

// dh = 64 needs 112 KB of shared memory and 256 TMEM columns per CTA: two CTAs share an SM, so one CTA's
// softmax (CUDA cores) overlaps the other's TMA loads and tensor-core work.
template <int DH>
__global__ void __launch_bounds__(kAttnThreads, DH == 64 ? 2 : 1)
enc_attention_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                     const EncAttnParams p) {
  using L = AttnSmem<DH>;
  constexpr int kPanels = L::kPanels;

  const int mt = blockIdx.x & 1;
  const int h = (blockIdx.x >> 1) % p.nhead;
  const int b = (blockIdx.x >> 1) / p.nhead;
  const int len = p.lens ? p.lens[b] : p.const_len;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const size_t row0 = size_t(b) * kLP + mt * 128;

  if (mt * 128 >= len) {
    // a tile that holds only padding rows: keep them finite (they are masked as keys downstream)
    for (int i = threadIdx.x; i < 128 * DH / 8; i += kAttnThreads) {
      const int r = i / (DH / 8), c = i % (DH / 8);
      reinterpret_cast<uint4*>(p.out + (row0 + r) * kD + h * DH)[c] = make_uint4(0, 0, 0, 0);
    }
    return;
  }

  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sQ = smem + L::kOffQ;
  uint8_t* sK = smem + L::kOffKP;
  uint8_t* sP = smem + L::kOffKP;
  uint8_t* sV = smem + L::kOffV;
  uint8_t* sVt = smem + L::kOffVt;
  const bool use_vt = p.v_mode == 1;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (use_vt ? L::kBarOffVt : L::kBarOffNoVt));
  uint64_t* bar_qk = bars + 0;
  uint64_t* bar_v = bars + 1;
  uint64_t* bar_s = bars + 2;
  uint64_t* bar_p = bars + 3;
  uint64_t* bar_o = bars + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

  const int keys16 = (len + 15) & ~15;  // keys processed by the tensor core (multiple of the UMMA K / N step)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_kv);
    mbar_init(bar_qk, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_p, 128);
    mbar_init(bar_o, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<256>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(bar_qk, L::kQ + kKeysMax * DH * 2);
#pragma unroll
      for (int pn = 0; pn < kPanels; ++pn) {
        tma_load_2d(&tm_q, bar_qk, sQ + pn * 16384, h * DH + pn * 64, int(row0));
        tma_load_2d(&tm_kv, bar_qk, sK + pn * 32768, kD + h * DH + pn * 64, b * kLP);
      }
      mbar_expect_tx(bar_v, L::kV);
#pragma unroll
      for (int pn = 0; pn < kPanels; ++pn)
        tma_load_2d(&tm_kv, bar_v, sV + pn * 32768, 2 * kD + h * DH + pn * 64, b * kLP);
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---- S = Q K^T : M=128, N=keys16, K=DH
      mbar_wait(bar_qk, 0);
      tc_fence_after_sync();
      const uint32_t idesc_s = make_idesc(kFmtBF16, 128, uint32_t(keys16), 0, 0);
#pragma unroll
      for (int k = 0; k < DH / 16; ++k) {
        const uint32_t qa = smem_u32(sQ) + (k / 4) * 16384 + (k % 4) * 32;
        const uint32_t ka = smem_u32(sK) + (k / 4) * 32768 + (k % 4) * 32;
        umma_bf16(tmem_base, make_smem_desc_sw128(qa, 16, 1024), make_smem_desc_sw128(ka, 16, 1024), idesc_s,
                  k != 0);
      }
      umma_commit(bar_s);

      // ---- O = P V : M=128, N=DH, K=keys16
      mbar_wait(bar_p, 0);
      mbar_wait(bar_v, 0);
      tc_fence_after_sync();
      const int nkk = keys16 / 16;
      if (!use_vt) {
        const uint32_t idesc_o = make_idesc(kFmtBF16, 128, DH, 0, 1);  // B (= V) is MN-major
        for (int kk = 0; kk < nkk; ++kk) {
          const uint32_t pa = smem_u32(sP) + (kk / 4) * 16384 + (kk % 4) * 32;
          const uint32_t va = smem_u32(sV) + kk * 2048;  // 16 key rows of 128 B
          umma_bf16(tmem_base, make_smem_desc_sw128(pa, 16, 1024), make_smem_desc_sw128(va, 32768, 1024), idesc_o,
                    kk != 0);
        }
      } else {
        const uint32_t idesc_o = make_idesc(kFmtBF16, 128, DH, 0, 0);
        for (int kk = 0; kk < nkk; ++kk) {
          const uint32_t pa = smem_u32(sP) + (kk / 4) * 16384 + (kk % 4) * 32;
          const uint32_t va = smem_u32(sVt) + (kk / 4) * (DH * 128) + (kk % 4) * 32;
          umma_bf16(tmem_base, make_smem_desc_sw128(pa, 16, 1024), make_smem_desc_sw128(va, 16, 1024), idesc_o,
                    kk != 0);
        }
      }
      umma_commit(bar_o);
    }
  } else {
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;  // query row inside the tile == TMEM lane
    const int tid = (warp - 2) * 32 + lane;
    const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16);
    const float sl2 = p.scale * 1.4426950408889634f;

    if (use_vt) {
      // V [key][d] (TMA swizzled) -> V^T [d][key] K-major swizzled, element-wise (cross-check path only)
      mbar_wait(bar_v, 0);
      for (int i = tid; i < kKeysMax * DH; i += 128) {
        const int k = i / DH, d = i % DH;
        const uint32_t src = (d / 64) * 32768 + k * 128 + ((((d % 64) / 8) ^ (k & 7)) << 4) + (d & 7) * 2;
        const uint32_t dst = (k / 64) * (DH * 128) + d * 128 + ((((k % 64) / 8) ^ (d & 7)) << 4) + (k & 7) * 2;
        *reinterpret_cast<uint16_t*>(sVt + dst) = *reinterpret_cast<const uint16_t*>(sV + src);
      }
    }

    mbar_wait(bar_s, 0);
    __syncwarp();
    tc_fence_after_sync();
    const int nchunks = (len + 31) / 32;

    float mx = -INFINITY;
    for (int c = 0; c < nchunks; ++c) {
      uint32_t v[32];
      tmem_ld32(taddr + c * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (c * 32 + j < len) mx = fmaxf(mx, __uint_as_float(v[j]));
    }
    const float mxs = mx * sl2;
    float sum = 0.f;
    for (int c = 0; c < nchunks; ++c) {
      uint32_t v[32];
      tmem_ld32(taddr + c * 32, v);
      tmem_ld_wait();
      uint32_t o[16];
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        float p0 = (c * 32 + j < len) ? exp2f(__uint_as_float(v[j]) * sl2 - mxs) : 0.f;
        float p1 = (c * 32 + j + 1 < len) ? exp2f(__uint_as_float(v[j + 1]) * sl2 - mxs) : 0.f;
        const __nv_bfloat162 pb = __floats2bfloat162_rn(p0, p1);
        const float2 pr = __bfloat1622float2(pb);
        sum += pr.x + pr.y;
        o[j >> 1] = *reinterpret_cast<const uint32_t*>(&pb);
      }
      uint8_t* prow = sP + (c >> 1) * 16384 + r * 128;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int chunk = (c & 1) * 4 + q;
        *reinterpret_cast<uint4*>(prow + ((chunk ^ (r & 7)) << 4)) =
            make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
      }
    }
    // keys in [32*nchunks, keys16) cannot exist (keys16 <= 32*nchunks); P is complete for the MMA
    fence_proxy_async_smem();
    tc_fence_before_sync();
    mbar_arrive(bar_p);

    mbar_wait(bar_o, 0);
    __syncwarp();
    tc_fence_after_sync();
    const float inv = 1.f / sum;
    __nv_bfloat16* orow = p.out + (row0 + r) * kD + h * DH;
#pragma unroll
    for (int c = 0; c < DH / 32; ++c) {
      uint32_t v[32];
      tmem_ld32(taddr + c * 32, v);
      tmem_ld_wait();
      uint32_t o[16];
#pragma unroll
      for (int j = 0; j < 32; j += 2)
        o[j >> 1] = pack_bf16x2(__uint_as_float(v[j]) * inv, __uint_as_float(v[j + 1]) * inv);
      uint4* dst = reinterpret_cast<uint4*>(orow + c * 32);
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
  static int smem_set = 0;
  const int bytes = L::bytes(p.v_mode == 1);
  if (smem_set < bytes) {
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return e;
    smem_set = bytes;
  }
  kfn<<<p.B * p.nhead * 2, kAttnThreads, bytes, stream>>>(tm_q, tm_kv, p);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_enc_attention(const CUtensorMap& tm_q, const CUtensorMap& tm_kv, const __nv_bfloat16* /*qkv*/,
                                 const EncAttnParams& p, cudaStream_t stream) {
  if (p.B <= 0) return cudaSuccess;
  const int dh = kD / p.nhead;
  if (dh == 64) return launch_dh<64>(tm_q, tm_kv, p, stream);
  if (dh == 128) return launch_dh<128>(tm_q, tm_kv, p, stream);
  return cudaErrorInvalidValue;
}

}  // namespace b200vqa
