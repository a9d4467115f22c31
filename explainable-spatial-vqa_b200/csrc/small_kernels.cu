// CUDA-core kernels of the executor path: everything that is a gather, a per-row reduction or a
// one-query attention. They are HBM- or latency-bound (no GEMM shape worth a tensor core), so the rules
// that matter are coalesced 16-byte accesses, enough loads in flight, and no host round trips.
#include <algorithm>

#include "kernels.h"
#include "ptx.cuh"

namespace b200vqa {
namespace {

__device__ __forceinline__ void store_bf16x8(__nv_bfloat16* dst, const float (&v)[8]) {
  *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]),
                                              pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}
__device__ __forceinline__ void load_bf16x8(const __nv_bfloat16* src, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(src);
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
__device__ __forceinline__ void load_f32x8(const float* src, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(src));
  const float4 b = __ldg(reinterpret_cast<const float4*>(src) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// ------------------------------------------------------------------------------------------------
// IQAP: [CLS] row, question rows and padding of the encoder input (IQAP:156-170). One warp per row,
// 8 columns per lane. Image rows 1..196 are written by the image_proj GEMM epilogue.
// ------------------------------------------------------------------------------------------------
__global__ void iqap_embed_kernel(const int64_t* __restrict__ questions, int B, int q_len,
                                  const float* __restrict__ cls, const float* __restrict__ emb, int vocab,
                                  const float* __restrict__ pe, int n_img, __nv_bfloat16* __restrict__ x) {
  const int rows_per_q = kLP - n_img;  // CLS + question + pad
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (gw >= B * rows_per_q) return;
  const int b = gw / rows_per_q;
  const int i = gw % rows_per_q;
  const int row = (i == 0) ? 0 : n_img + i;  // 0, 197, 198, ...
  float v[8];
  if (row == 0) {
    float c[8], p8[8];
    load_f32x8(cls + lane * 8, c);
    load_f32x8(pe + lane * 8, p8);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = c[j] + p8[j];
  } else if (row < 1 + n_img + q_len) {
    long long tok = questions[size_t(b) * q_len + (row - 1 - n_img)];
    tok = tok < 0 ? 0 : (tok >= vocab ? vocab - 1 : tok);
    float e[8], p8[8];
    load_f32x8(emb + size_t(tok) * kD + lane * 8, e);
    load_f32x8(pe + size_t(row) * kD + lane * 8, p8);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = e[j] + p8[j];
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
  }
  store_bf16x8(x + (size_t(b) * kLP + row) * kD + lane * 8, v);
}

// ------------------------------------------------------------------------------------------------
// FA: encoder input of one chain step, gathered through the HBM-resident inference cache
// (run_inference_chain, FA:96-118): src = [func] + cache[dep0] + cache[dep1]; x = [img | emb(src)] + PE.
// One block per question.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fa_build_src_kernel(const FaBuildSrcParams p) {
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  __shared__ int s_tok[64];
  __shared__ int s_len;

  if (threadIdx.x == 0) {
    int n = 0;
    if (p.src_direct) {
      n = p.src_len_in ? p.src_len_in[b] : p.src_ld;
      n = n > 60 ? 60 : n;
      for (int j = 0; j < n; ++j) s_tok[j] = int(p.src_direct[size_t(b) * p.src_ld + j]);
    } else {
      const size_t bs = size_t(b) * p.S + p.step;
      s_tok[n++] = p.func[bs];
      for (int d = 0; d < 2; ++d) {
        const int dep = p.deps[bs * 2 + d];
        // a pointer to a step without cached output contributes nothing (FA:110-115 `cache.get(idx, "")`)
        if (dep >= 0 && dep < p.step && dep < p.n_steps[b]) {
          const int32_t* c = p.cache + (size_t(b) * p.S + dep) * p.T;
          for (int j = 0; j < p.T && n < 60; ++j) s_tok[n++] = c[j];
        }
      }
    }
    const int room = p.pe_len - p.n_img;  // positional table bounds the sequence (FA:40)
    if (n > room) n = room;
    if (n > kLP - p.n_img) n = kLP - p.n_img;
    s_len = n;
    p.lens[b] = p.n_img + n;
  }
  __syncthreads();
  const int n = s_len;

  // image tokens: straight 16-byte copy (PE rows 0..195 were folded in by fa_project_images)
  int img = b;
  if (p.image_idx) {
    img = p.image_idx[b];
    img = img < 0 ? 0 : (img >= p.n_images ? p.n_images - 1 : img);
  }
  const uint4* src = reinterpret_cast<const uint4*>(p.img_tokens + size_t(img) * p.n_img * kD);
  uint4* dst = reinterpret_cast<uint4*>(p.x + size_t(b) * kLP * kD);
  const int n16 = p.n_img * kD / 8;
  for (int i = threadIdx.x; i < n16; i += 256) dst[i] = src[i];

  // text rows + zero padding, one warp per row
  for (int row = p.n_img + warp; row < kLP; row += 8) {
    float v[8];
    const int j = row - p.n_img;
    if (j < n) {
      int tok = s_tok[j];
      tok = tok < 0 ? 0 : (tok >= p.vocab ? p.vocab - 1 : tok);
      float e[8], pe8[8];
      load_f32x8(p.emb + size_t(tok) * kD + lane * 8, e);
      load_f32x8(p.pe + size_t(row) * kD + lane * 8, pe8);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = e[k] + pe8[k];
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = 0.f;
    }
    store_bf16x8(p.x + (size_t(b) * kLP + row) * kD + lane * 8, v);
  }
}

// ------------------------------------------------------------------------------------------------
// Row LayerNorm (nn.Transformer's final encoder norm, FA:42). One warp per row.
// ------------------------------------------------------------------------------------------------
__global__ void layernorm_rows_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                      const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                      int rows) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float v[8];
  load_bf16x8(in + size_t(row) * kD + lane * 8, v);
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += v[j];
  const float mean = warp_sum(s) * (1.f / kD);
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) q += (v[j] - mean) * (v[j] - mean);
  const float rstd = rsqrtf(warp_sum(q) * (1.f / kD) + eps);
  float g[8], bt[8];
  load_f32x8(gamma + lane * 8, g);
  load_f32x8(beta + lane * 8, bt);
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = (v[j] - mean) * rstd * g[j] + bt[j];
  store_bf16x8(out + size_t(row) * kD + lane * 8, v);
}

// fp32 (or bf16: already rounded on the host) [B, C, P] -> bf16 [B*P, C] through a padded 32x32 shared tile.
template <typename TIn>
__global__ void transpose_cast_kernel(const TIn* __restrict__ in, __nv_bfloat16* __restrict__ out, int C, int P) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, pp = p0 + tx;
    tile[i][tx] = (c < C && pp < P) ? float(in[(size_t(b) * C + c) * P + pp]) : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int pp = p0 + i, c = c0 + tx;
    if (pp < P && c < C) out[(size_t(b) * P + pp) * C + c] = __float2bfloat16(tile[tx][i]);
  }
}

__global__ void memory_import_kernel(const float* __restrict__ mem, int S, int B, int ldb,
                                     __nv_bfloat16* __restrict__ x) {
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (gw >= B * kLP) return;
  const int b = gw / kLP, s = gw % kLP;
  float v[8];
  if (s < S) load_f32x8(mem + (size_t(s) * ldb + b) * kD + lane * 8, v);
  else {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
  }
  store_bf16x8(x + size_t(gw) * kD + lane * 8, v);
}

__global__ void memory_export_kernel(const __nv_bfloat16* __restrict__ x, int S, int B, int ldb,
                                     float* __restrict__ mem) {
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (gw >= B * S) return;
  const int s = gw / B, b = gw % B;
  float v[8];
  load_bf16x8(x + (size_t(b) * kLP + s) * kD + lane * 8, v);
  float4* dst = reinterpret_cast<float4*>(mem + (size_t(s) * ldb + b) * kD + lane * 8);
  dst[0] = make_float4(v[0], v[1], v[2], v[3]);
  dst[1] = make_float4(v[4], v[5], v[6], v[7]);
}

// ------------------------------------------------------------------------------------------------
// IQAP answer head (IQAP:122-127,176-179) on the CLS row; fp32 weights, one block per question.
// w0_t is answer_classifier.0.weight transposed to [d, hidden] so thread j reads coalesced.
// pool_rows > 0: the input is the mean of memory rows 1..pool_rows instead (the image tokens) - the bounding-box
// regressor of train_transformer_iqap_bb.py:304-310 (same Linear -> ReLU -> Linear shape, 40 outputs).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) answer_head_kernel(const __nv_bfloat16* __restrict__ memory,
                                                          const float* __restrict__ w0_t,
                                                          const float* __restrict__ b0, int hidden,
                                                          const float* __restrict__ w1,
                                                          const float* __restrict__ b1, int classes, int pool_rows,
                                                          float* __restrict__ out) {
  __shared__ float xs[kD];
  __shared__ float hs[1024];
  const int b = blockIdx.x;
  const __nv_bfloat16* mrow = memory + size_t(b) * kLP * kD + threadIdx.x;  // thread = channel: coalesced 512-byte rows
  if (pool_rows <= 0) {
    xs[threadIdx.x] = __bfloat162float(mrow[0]);
  } else {
    float acc = 0.f;
    int r = 1;
    for (; r + 7 <= pool_rows; r += 8) {  // eight rows in flight
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __bfloat162float(mrow[size_t(r + u) * kD]);
#pragma unroll
      for (int u = 0; u < 8; ++u) acc += v[u];
    }
    for (; r <= pool_rows; ++r) acc += __bfloat162float(mrow[size_t(r) * kD]);
    xs[threadIdx.x] = acc / float(pool_rows);
  }
  __syncthreads();
  for (int j = threadIdx.x; j < hidden; j += 256) {
    float acc = b0[j];
    for (int k = 0; k < kD; ++k) acc = fmaf(xs[k], __ldg(w0_t + size_t(k) * hidden + j), acc);
    hs[j] = fmaxf(acc, 0.f);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c = warp; c < classes; c += 8) {
    float acc = 0.f;
    for (int j = lane; j < hidden; j += 32) acc = fmaf(hs[j], __ldg(w1 + size_t(c) * hidden + j), acc);
    acc = warp_sum(acc);
    if (lane == 0) out[size_t(b) * classes + c] = acc + b1[c];
  }
}

__global__ void cast_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, size_t n) {
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x)
    out[i] = __float2bfloat16(in[i]);
}

__global__ void cast_f16_kernel(const float* __restrict__ in, __half* __restrict__ out, size_t n) {
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x)
    out[i] = __float2half_rn(in[i]);
}

__global__ void transpose_f32_kernel(const float* __restrict__ in, float* __restrict__ out, int R, int C) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < R && c < C) ? in[size_t(r) * C + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < C && r < R) out[size_t(c) * R + r] = tile[threadIdx.x][i];
  }
}

// Program -> chain glue on the device (reference preprocess_questions/utils_programs.py:100-156 `prefix_to_list`):
// a program in PREFIX order becomes chain elements in execution order (= post-order: inputs before consumers, root
// last, exactly tree_to_list's numbering) with dependency pointers.  One thread per question; malformed programs
// (a terminator before the tree closes, more than S nodes) are truncated, never out of bounds.
// (nvcc's "used before its value is set" on tokv / child / nchild is a false positive: emit(node) only ever sees nodes
//  whose three entries were written when the node was pushed; zero-filling 320 ints per thread would cost more than
//  the kernel)
#pragma nv_diag_suppress 549
__global__ void programs_to_chain_kernel(const ProgToChainParams p) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= p.B) return;
  constexpr int kMax = 64;
  int st_node[kMax], st_rem[kMax], tokv[kMax], child[kMax][2], nchild[kMax];
  int sp = 0, cnt = 0;
  const int T = p.T < kMax ? p.T : kMax;
  int32_t* func = p.func + size_t(b) * p.S;
  int32_t* deps = p.deps + size_t(b) * p.S * 2;
  auto emit = [&](int node) {
    const int idx = cnt < p.S ? cnt : p.S - 1;  // truncated programs overwrite the last slot, never past it
    func[idx] = p.func_map[tokv[node]];
    deps[idx * 2] = nchild[node] > 0 ? child[node][0] : -1;
    deps[idx * 2 + 1] = nchild[node] > 1 ? child[node][1] : -1;
    if (cnt < p.S) ++cnt;
    return idx;
  };
  for (int pos = 0; pos < T; ++pos) {
    long long tok = p.programs[size_t(b) * p.T + pos];
    if (tok < 0 || tok >= p.prog_vocab) break;
    const int a = p.arity[tok];
    if (a < 0) break;  // <NULL> / <START> / <END>: end of program
    st_node[sp] = pos;
    st_rem[sp] = a > 2 ? 2 : a;
    tokv[pos] = int(tok);
    nchild[pos] = 0;
    ++sp;
    while (sp > 0 && st_rem[sp - 1] == 0) {
      const int node = st_node[--sp];
      const int idx = emit(node);
      if (sp > 0) {
        const int parent = st_node[sp - 1];
        if (nchild[parent] < 2) child[parent][nchild[parent]++] = idx;
        --st_rem[sp - 1];
      }
    }
    if (sp == 0) break;  // the tree is complete
  }
  while (sp > 0) {  // truncated program: close the open nodes with the inputs they have
    const int node = st_node[--sp];
    const int idx = emit(node);
    if (sp > 0) {
      const int parent = st_node[sp - 1];
      if (nchild[parent] < 2) child[parent][nchild[parent]++] = idx;
    }
  }
  for (int i = cnt; i < p.S; ++i) {
    func[i] = 0;
    deps[i * 2] = deps[i * 2 + 1] = -1;
  }
  p.n_steps[b] = cnt;
}
#pragma nv_diag_default 549

// Holds the stream busy for `cycles` SM clocks: lets the host enqueue a whole step behind it so that the
// profiler's event timestamps see back-to-back kernels instead of host launch latency.
__global__ void delay_kernel(long long cycles) {
  const long long t0 = clock64();
  while (clock64() - t0 < cycles) {
  }
}

inline int ceil_div(long long a, long long b) { return int((a + b - 1) / b); }

}  // namespace

cudaError_t launch_iqap_embed(const int64_t* questions, int B, int q_len, const float* cls, const float* emb,
                              int vocab, const float* pe, int n_img, __nv_bfloat16* x, cudaStream_t stream) {
  const long long warps = (long long)B * (kLP - n_img);
  iqap_embed_kernel<<<ceil_div(warps * 32, 256), 256, 0, stream>>>(questions, B, q_len, cls, emb, vocab, pe, n_img, x);
  return cudaGetLastError();
}

cudaError_t launch_fa_build_src(const FaBuildSrcParams& p, cudaStream_t stream) {
  fa_build_src_kernel<<<p.B, 256, 0, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_layernorm_rows(const __nv_bfloat16* in, __nv_bfloat16* out, const float* gamma,
                                  const float* beta, float eps, int rows, cudaStream_t stream) {
  layernorm_rows_kernel<<<ceil_div((long long)rows * 32, 256), 256, 0, stream>>>(in, out, gamma, beta, eps, rows);
  return cudaGetLastError();
}

cudaError_t launch_transpose_cast(const float* in, __nv_bfloat16* out, int B, int C, int P, cudaStream_t stream) {
  dim3 grid(ceil_div(P, 32), ceil_div(C, 32), B), block(32, 8);
  transpose_cast_kernel<float><<<grid, block, 0, stream>>>(in, out, C, P);
  return cudaGetLastError();
}

cudaError_t launch_transpose_bf16(const __nv_bfloat16* in, __nv_bfloat16* out, int B, int C, int P, cudaStream_t stream) {
  dim3 grid(ceil_div(P, 32), ceil_div(C, 32), B), block(32, 8);
  transpose_cast_kernel<__nv_bfloat16><<<grid, block, 0, stream>>>(in, out, C, P);
  return cudaGetLastError();
}

cudaError_t launch_memory_import(const float* mem, int S, int B, int ldb, __nv_bfloat16* x, cudaStream_t stream) {
  memory_import_kernel<<<ceil_div((long long)B * kLP * 32, 256), 256, 0, stream>>>(mem, S, B, ldb, x);
  return cudaGetLastError();
}

cudaError_t launch_memory_export(const __nv_bfloat16* x, int S, int B, int ldb, float* mem, cudaStream_t stream) {
  memory_export_kernel<<<ceil_div((long long)B * S * 32, 256), 256, 0, stream>>>(x, S, B, ldb, mem);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Evaluation tally on the device (reference inference_transformer_iqap_tally.py:317-344): per question
// predicted answer = first maximum of the answer logits (torch.max), program correct = all T tokens equal;
// counts[0..3] += {both correct, answer only, program only, neither}.  One warp per question.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tally_kernel(const TallyParams p) {
  const int b = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= p.B) return;
  float best = -INFINITY;
  int arg = 0x7fffffff;
  for (int c = lane; c < p.classes; c += 32) {
    const float v = p.answer_logits[size_t(b) * p.classes + c];
    if (v > best || (v == best && c < arg) || arg == 0x7fffffff) { best = v; arg = c; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
    // NaN never wins a comparison in torch.max unless it is present - logits are finite here; lowest index wins ties
    if (oa != 0x7fffffff && (arg == 0x7fffffff || ob > best || (ob == best && oa < arg))) { best = ob; arg = oa; }
  }
  bool same = true;
  for (int t = lane; t < p.T; t += 32)
    same &= p.programs[size_t(b) * p.T + t] == p.gt_programs[size_t(b) * p.T + t];
  const bool prog_ok = __all_sync(0xffffffffu, same);
  if (lane == 0) {
    const bool ans_ok = int64_t(arg) == p.gt_answers[b];
    if (p.pred_answers) p.pred_answers[b] = arg;
    atomicAdd(p.counts + (ans_ok ? (prog_ok ? 0 : 1) : (prog_ok ? 2 : 3)), 1ull);
  }
}

// x[b*kLP + 1 + pos] = img_tok[image_idx[b]*n_img_tokens + pos]  (16 bytes per thread; the image rows of a question's
// block when several questions share one image: image_proj + PE were applied once per image)
__global__ void __launch_bounds__(256) gather_image_rows_kernel(const __nv_bfloat16* __restrict__ img_tok,
                                                                const int32_t* __restrict__ image_idx, int n_img,
                                                                int n_img_tokens, __nv_bfloat16* __restrict__ x) {
  const int b = blockIdx.y;
  int img = image_idx[b];
  img = img < 0 ? 0 : (img >= n_img ? n_img - 1 : img);
  const uint4* src = reinterpret_cast<const uint4*>(img_tok + size_t(img) * n_img_tokens * kD);
  uint4* dst = reinterpret_cast<uint4*>(x + (size_t(b) * kLP + 1) * kD);
  const int n = n_img_tokens * kD / 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = __ldg(src + i);
}

// Absorbed query-key weights of one cross-attention block (kernels.h: launch_absorb_qk).  Block = one output row
// (head h, memory channel i), thread = input channel d; fp32 accumulation over the head dimension.
__global__ void __launch_bounds__(kD) absorb_qk_kernel(const float* __restrict__ w_in, const float* __restrict__ b_in,
                                                       int nhead, __nv_bfloat16* __restrict__ w_qk,
                                                       float* __restrict__ b_qk) {
  const int dh = kD / nhead;
  const int h = blockIdx.x / kD, i = blockIdx.x % kD, d = threadIdx.x;
  const float* wq = w_in;                    // rows 0 .. d-1
  const float* wk = w_in + size_t(kD) * kD;  // rows d .. 2d-1
  float acc = 0.f, accb = 0.f;
  for (int e = 0; e < dh; ++e) {
    const float k = wk[size_t(h * dh + e) * kD + i];
    acc = fmaf(k, wq[size_t(h * dh + e) * kD + d], acc);
    accb = fmaf(k, b_in[h * dh + e], accb);
  }
  w_qk[size_t(blockIdx.x) * kD + d] = __float2bfloat16(acc);
  if (d == 0) b_qk[blockIdx.x] = accb;
}

// Absorbed value-output weights of one cross-attention block (kernels.h: launch_absorb_ov).  Block = one output channel
// n, thread = memory channel i of head blockIdx.y; fp32 accumulation over the head dimension.
__global__ void __launch_bounds__(kD) absorb_ov_kernel(const float* __restrict__ w_in, const float* __restrict__ b_in,
                                                       const float* __restrict__ w_out, const float* __restrict__ b_out,
                                                       int nhead, __nv_bfloat16* __restrict__ w_ov,
                                                       float* __restrict__ b_ov) {
  const int dh = kD / nhead;
  const int n = blockIdx.x, h = blockIdx.y, i = threadIdx.x;
  const float* wv = w_in + size_t(2) * kD * kD;  // rows 2d .. 3d-1
  float acc = 0.f;
  for (int e = 0; e < dh; ++e) acc = fmaf(w_out[size_t(n) * kD + h * dh + e], wv[size_t(h * dh + e) * kD + i], acc);
  w_ov[size_t(n) * nhead * kD + h * kD + i] = __float2bfloat16(acc);
  if (h == 0 && i == 0) {
    float accb = b_out[n];
    for (int c = 0; c < kD; ++c) accb = fmaf(w_out[size_t(n) * kD + c], b_in[2 * kD + c], accb);
    b_ov[n] = accb;
  }
}

cudaError_t launch_absorb_ov(const float* in_proj_weight, const float* in_proj_bias, const float* out_proj_weight,
                             const float* out_proj_bias, int nhead, __nv_bfloat16* w_ov, float* b_ov,
                             cudaStream_t stream) {
  absorb_ov_kernel<<<dim3(kD, nhead), kD, 0, stream>>>(in_proj_weight, in_proj_bias, out_proj_weight, out_proj_bias,
                                                      nhead, w_ov, b_ov);
  return cudaGetLastError();
}

cudaError_t launch_absorb_qk(const float* in_proj_weight, const float* in_proj_bias, int nhead, __nv_bfloat16* w_qk,
                             float* b_qk, cudaStream_t stream) {
  absorb_qk_kernel<<<nhead * kD, kD, 0, stream>>>(in_proj_weight, in_proj_bias, nhead, w_qk, b_qk);
  return cudaGetLastError();
}

cudaError_t launch_tally(const TallyParams& p, cudaStream_t stream) {
  if (p.B <= 0) return cudaSuccess;
  tally_kernel<<<(p.B + 7) / 8, 256, 0, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_gather_image_rows(const __nv_bfloat16* img_tok, const int32_t* image_idx, int n_img, int n_img_tokens,
                                     int B, __nv_bfloat16* x, cudaStream_t stream) {
  if (B <= 0) return cudaSuccess;
  gather_image_rows_kernel<<<dim3(4, B), 256, 0, stream>>>(img_tok, image_idx, n_img, n_img_tokens, x);
  return cudaGetLastError();
}

cudaError_t launch_answer_head(const __nv_bfloat16* memory, int B, const float* w0_t, const float* b0, int hidden,
                               const float* w1, const float* b1, int classes, int pool_rows, float* out,
                               cudaStream_t stream) {
  if (hidden > 1024 || pool_rows >= kLP) return cudaErrorInvalidValue;
  answer_head_kernel<<<B, 256, 0, stream>>>(memory, w0_t, b0, hidden, w1, b1, classes, pool_rows, out);
  return cudaGetLastError();
}

__global__ void scatter_rows_i32_kernel(const int32_t* __restrict__ src, const int32_t* __restrict__ order, int n,
                                        int width, int32_t* __restrict__ dst) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)n * width) return;
  const int r = int(i / width), c = int(i % width);
  dst[(long long)order[r] * width + c] = src[i];
}

cudaError_t launch_scatter_rows_i32(const int32_t* src, const int32_t* order, int n, int width, int32_t* dst,
                                    cudaStream_t stream) {
  const long long total = (long long)n * width;
  if (total <= 0) return cudaSuccess;
  scatter_rows_i32_kernel<<<unsigned((total + 255) / 256), 256, 0, stream>>>(src, order, n, width, dst);
  return cudaGetLastError();
}

cudaError_t launch_programs_to_chain(const ProgToChainParams& p, cudaStream_t stream) {
  if (p.B <= 0) return cudaSuccess;
  programs_to_chain_kernel<<<(p.B + 127) / 128, 128, 0, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_delay(long long cycles, cudaStream_t stream) {
  delay_kernel<<<1, 1, 0, stream>>>(cycles);
  return cudaGetLastError();
}

cudaError_t launch_cast_bf16(const float* in, __nv_bfloat16* out, size_t n, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  const int grid = int(std::min<size_t>((n + 255) / 256, 4096));
  cast_bf16_kernel<<<grid, 256, 0, stream>>>(in, out, n);
  return cudaGetLastError();
}

cudaError_t launch_cast_f16(const float* in, __half* out, size_t n, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  const int grid = int(std::min<size_t>((n + 255) / 256, 4096));
  cast_f16_kernel<<<grid, 256, 0, stream>>>(in, out, n);
  return cudaGetLastError();
}

cudaError_t launch_transpose_f32(const float* in, float* out, int R, int C, cudaStream_t stream) {
  dim3 grid(ceil_div(C, 32), ceil_div(R, 32)), block(32, 8);
  transpose_f32_kernel<<<grid, block, 0, stream>>>(in, out, R, C);
  return cudaGetLastError();
}

}  // namespace b200vqa
