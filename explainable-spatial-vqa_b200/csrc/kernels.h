// Internal (C++) launch interface between the C-ABI layer (api.cu) and the kernel translation units.
// Nothing here crosses the shared-library boundary; the public surface is include/b200vqa.h.
#pragma once
#include <cstdint>
#include <utility>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace b200vqa {

// Launch helper: plain <<<>>> semantics, optionally with programmatic dependent launch (the kernel must call
// pdl_wait() before it touches anything its predecessor produced or still reads).
bool pdl_enabled();
void set_pdl_enabled(bool on);

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                 bool pdl, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl && pdl_enabled()) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// Padded per-question row count of the encoder activations (sequence rows live at [b*kLP, b*kLP+len)).
constexpr int kLP = 256;
constexpr int kD = 256;  // d_model the kernels are specialised for (IQAP:15, FA:157)

// ------------------------------------------------------------------------------------------
// tcgen05 GEMM:  out[M,N] = epilogue( A[M,K] . W[N,K]^T + bias )
// ------------------------------------------------------------------------------------------
enum GemmEpilogue : int {
  kEpiBias = 0,       // bf16 out = acc + bias
  kEpiBiasRelu = 1,   // bf16 out = relu(acc + bias)
  kEpiBiasResLN = 2,  // bf16 out = LayerNorm(acc + bias + residual) * gamma + beta   (N == 256)
  kEpiBiasPeRemap = 3,// bf16 out[row'] = acc + bias + pe[p]; row' = item*rows_out + off + p
  kEpiHead = 4,       // vocabulary head of one decode position: logits = acc + bias (N = padded vocabulary), argmax ->
                      // tok[row, t+1], next input x_next[row] = emb[next] + pe[t+1]   (IQAP:230-236, FA:142-145)
  kEpiLstm = 5,       // one LSTM time step of the program generator: gates = acc (h_prev . W_hh^T) + table[token]
                      // (embedding . W_ih^T + biases, precomputed per vocabulary entry); c, h updated in the epilogue
  kEpiBiasResLN2 = 6  // kEpiBiasResLN followed by a second LayerNorm (gamma2 / beta2) of its output: nn.Transformer's final
                      // encoder norm fused into the last layer's norm2 (its own instantiation: the common epilogue keeps
                      // its register budget)
};

struct GemmParams {
  int M = 0, N = 0, K = 0;
  bool pdl = false;                  // launch with programmatic dependent launch (decode chain)
  int a_group_cols = 0;              // > 0: grouped GEMM - the n-tile at column n0 reads A columns
                                     // [(n0 / a_group_cols) * K, +K) (per-head value projection of absorbed attention)
  bool f16 = false;                  // 2-byte operands are IEEE fp16 instead of bf16 (fp16 feature store -> image_proj)
  bool ln_cluster = false;           // host-side: kEpiBiasResLN call site of the decode chain -> cluster-of-4 kernel
  long long* dbg_clk = nullptr;      // optional: CTA 0 writes clock64() stamps of its pipeline stages (tools/ only)
  const float* bias = nullptr;       // [N] fp32 (may be null)
  __nv_bfloat16* out = nullptr;      // bf16 output
  int ldc = 0;                       // output leading dimension (elements)
  // kEpiBiasResLN
  const __nv_bfloat16* residual = nullptr;
  int ldr = 0;
  const float* gamma = nullptr;
  const float* beta = nullptr;
  float eps = 1e-5f;
  float* out_f32 = nullptr;          // optional fp32 copy of the LN output, leading dim N
  const float* gamma2 = nullptr;     // kEpiBiasResLN2: the second LayerNorm's weight / bias
  const float* beta2 = nullptr;
  // kEpiBiasPeRemap
  int rows_in = 1;                   // GEMM rows per item (196 image tokens)
  int rows_out = 1;                  // output rows per item (kLP, or 196 for the FA image-token store)
  int row_off = 0;                   // first output row inside the item
  const float* pe = nullptr;         // [*, N] fp32 positional-encoding table
  int pe_off = 0;                    // pe row = pe_off + p
  // kEpiHead (A = fp32 decoder output [M, 256] read as tf32, W = fp32 head weight [V, 256])
  int head_V = 0;                    // real vocabulary size (columns >= V are padding)
  int head_t = 0;                    // decode position
  int64_t* tok = nullptr;            // [M, tok_ld]; argmax (lowest index wins ties) written to column t+1
  int tok_ld = 0;
  float* logits = nullptr;           // optional [M, logits_T, V]; row t
  int logits_T = 0;
  const int64_t* forced = nullptr;   // optional teacher forcing: position t+1 is fed forced[row*forced_ld + t]
  int forced_ld = 0;
  const float* emb = nullptr;        // [vocab, 256] decoder embedding
  int vocab = 0;
  const float* pe_next = nullptr;    // pe row t+1, null on the last position
  __nv_bfloat16* x_next = nullptr;   // [M, 256]
  // kEpiLstm (N = 4 * hidden; columns permuted so that an n-tile holds i|f|g|o of 64 hidden units, 64 columns each)
  const float* lstm_table = nullptr;    // [vocab, N] fp32, same column permutation
  const int64_t* lstm_tokens = nullptr; // token of row r = lstm_tokens[r * lstm_tok_ld + lstm_tok_col]; null -> lstm_token_const
  int lstm_tok_ld = 0, lstm_tok_col = 0;
  int lstm_token_const = 0;
  int lstm_vocab = 0;
  float* lstm_c = nullptr;              // [M, hidden] fp32 cell state, updated in place
  __nv_bfloat16* lstm_h = nullptr;      // [M, hidden] bf16 new hidden state (the next step's A operand: ping-pong)
  float* lstm_h_f32 = nullptr;          // optional fp32 copy (feeds the tf32 vocabulary head)
};

// A/W tensor maps: 2D, 128-byte swizzle, box = {128 B of K, 128 rows (A) | BN rows (W)}.
// The output leaves through plain coalesced stores (p.out, leading dimension p.ldc).
cudaError_t launch_gemm(int epilogue, bool tf32, int block_n, const CUtensorMap& tm_a, const CUtensorMap& tm_w,
                        const GemmParams& p, int num_sms, cudaStream_t stream);

// Plain CUDA-core GEMM used ONLY by the test suite to cross-check the tensor-core kernel on the GPU
// at sizes where a host check is too slow. fp32 accumulate over the same bf16 (or fp32) inputs.
cudaError_t launch_gemm_check(bool a_is_f32, const void* A, const void* W, const float* bias, float* out, int M,
                              int N, int K, cudaStream_t stream);

// ------------------------------------------------------------------------------------------
// Encoder self-attention (tcgen05): one CTA per (question, head, 128-query tile).
//   qkv  [B*kLP, 3*kD] bf16 (q | k | v), lens[b] = valid rows (keys >= len are masked), out [B*kLP, kD].
// ------------------------------------------------------------------------------------------
struct EncAttnParams {
  int B = 0;
  int nhead = 4;
  const int32_t* lens = nullptr;  // [B] or null -> const_len
  int const_len = 0;
  __nv_bfloat16* out = nullptr;   // [B*kLP, kD]
  float scale = 0.125f;           // 1/sqrt(dh)
  int v_mode = 0;                 // 0: V as MN-major operand straight from TMA; 1: transposed in smem first
  bool one_cta_per_head = false;  // dh = 64 only: the 160 KB both-tiles-per-CTA kernel instead of one tile per CTA (A/B runs)
};
cudaError_t launch_enc_attention(const CUtensorMap& tm_qkv, const CUtensorMap& tm_v, const __nv_bfloat16* qkv,
                                 const EncAttnParams& p, cudaStream_t stream);

// ------------------------------------------------------------------------------------------
// Small CUDA-core kernels (memory-bound or tiny)
// ------------------------------------------------------------------------------------------
// IQAP encoder-input rows that are not image tokens: CLS row 0 and question rows 197..242 (+PE), pad rows zero.
cudaError_t launch_iqap_embed(const int64_t* questions, int B, int q_len, const float* cls, const float* emb,
                              int vocab, const float* pe, int n_img, __nv_bfloat16* x, cudaStream_t stream);

// FA: x[b] = [img_tokens[b] (PE already added) | emb(src)+PE | zero pad]; src gathered through the step cache.
struct FaBuildSrcParams {
  int B = 0;
  int step = 0;                      // chain step index i
  int S = 0;                         // steps dimension of func/deps/cache
  int T = 20;                        // tokens per cached step output
  const int32_t* func = nullptr;     // [B,S]
  const int32_t* deps = nullptr;     // [B,S,2]  (-1 = none; >= step or unwritten -> contributes nothing)
  const int32_t* n_steps = nullptr;  // [B]
  const int32_t* cache = nullptr;    // [B,S,T]
  const int64_t* src_direct = nullptr;  // alternative: explicit src tokens [B, src_ld] with src_len[b]
  const int32_t* src_len_in = nullptr;
  int src_ld = 0;
  const __nv_bfloat16* img_tokens = nullptr;  // [B,196,kD], or [n_images,196,kD] with image_idx
  const int32_t* image_idx = nullptr;         // optional [B]: image of question b (several questions per image)
  int n_images = 0;
  const float* emb = nullptr;
  int vocab = 0;
  const float* pe = nullptr;
  int pe_len = 0;
  int n_img = 196;
  __nv_bfloat16* x = nullptr;        // [B*kLP, kD]
  int32_t* lens = nullptr;           // [B] out: 196 + src_len
};
cudaError_t launch_fa_build_src(const FaBuildSrcParams& p, cudaStream_t stream);

// Row LayerNorm over kD columns, bf16 in/out (nn.Transformer's final encoder norm, FA:42).
cudaError_t launch_layernorm_rows(const __nv_bfloat16* in, __nv_bfloat16* out, const float* gamma,
                                  const float* beta, float eps, int rows, cudaStream_t stream);

// fp32 [B, C, P] (channel-major, FA:47) -> bf16 [B*P, C]
cudaError_t launch_transpose_cast(const float* in, __nv_bfloat16* out, int B, int C, int P, cudaStream_t stream);
// the same from values already rounded to bf16 (the host-buffer entry's bf16 upload mode)
cudaError_t launch_transpose_bf16(const __nv_bfloat16* in, __nv_bfloat16* out, int B, int C, int P, cudaStream_t stream);

// seq-first fp32 memory [S,B,kD] -> padded bf16 rows [B*kLP, kD] (IQAP:190 entry with caller memory)
// (ldb = batch size of the full seq-first tensor; `mem` already points at this chunk's first question)
cudaError_t launch_memory_import(const float* mem, int S, int B, int ldb, __nv_bfloat16* x, cudaStream_t stream);
// padded bf16 rows -> seq-first fp32 [S,B,kD]
cudaError_t launch_memory_export(const __nv_bfloat16* x, int S, int B, int ldb, float* mem, cudaStream_t stream);

// IQAP answer head on the CLS row: Linear(256,hidden)+ReLU+Linear(hidden,C), fp32 weights (IQAP:122-127,179).
// pool_rows > 0: on the mean of memory rows 1..pool_rows instead (bbox regressor, train_transformer_iqap_bb.py:304-310).
cudaError_t launch_answer_head(const __nv_bfloat16* memory, int B, const float* w0, const float* b0, int hidden,
                               const float* w1, const float* b1, int classes, int pool_rows, float* out,
                               cudaStream_t stream);

// ------------------------------------------------------------------------------------------
// Decoder (decode_kernels.cu).  All decode positions of a question are produced on the device; the greedy
// tokens accumulate in a library-owned int64 buffer tok[B, tok_ld] (column 0 = start token) and are copied
// to the caller's layout once at the end (launch_publish_tokens), so the decode loop itself has a fixed
// launch sequence with fixed arguments and can be replayed as a CUDA graph.
// ------------------------------------------------------------------------------------------
// x0[b] = emb[start] + pe[0]; tok[b,0] = start.
struct DecEmbedParams {
  int B = 0;
  const float* emb = nullptr;
  int vocab = 0;
  const float* pe = nullptr;
  int start_token = 0;
  const int64_t* start_tokens = nullptr;  // optional per-question start token: start_tokens[b*start_ld]
  int start_ld = 0;
  __nv_bfloat16* x = nullptr;             // [B,kD]
  int64_t* tok = nullptr;                 // [B, tok_ld]
  int tok_ld = 0;
};
cudaError_t launch_dec_embed_start(const DecEmbedParams& p, cudaStream_t stream);

// One query row per question against `len` key/value rows (all heads; one CTA per question).  Serves both
// the decoder self-attention over its KV cache (exact form of the reference's causal-mask recompute,
// IQAP:208-227 / FA:137-141; the new position's k/v are appended first) and the cross-attention over the
// projected encoder memory.  Key row j of question b is at k + (b*rows_per_q + j)*ld: 256 contiguous bf16.
struct RowAttnParams {
  int B = 0, nhead = 4;
  bool pdl = false;
  const __nv_bfloat16* q = nullptr;  // [B, ldq], columns 0..255
  int ldq = 0;
  const __nv_bfloat16* k = nullptr;
  const __nv_bfloat16* v = nullptr;
  long long rows_per_q = 0;
  int ld = 0;
  const int32_t* lens = nullptr;     // [B] or null -> const_len
  int const_len = 0;
  // self-attention only: this position's key/value (row b of new_k / new_v, leading dim ld_new) is written to
  // row append_pos of the caches (k_app / v_app alias k / v) before attending
  const __nv_bfloat16* new_k = nullptr;
  const __nv_bfloat16* new_v = nullptr;
  int ld_new = 0;
  int append_pos = 0;
  bool warp_form = true;  // self-attention over <= 32 keys: one warp per question (self_attn_warp_kernel); false = CTA per question
  __nv_bfloat16* k_app = nullptr;
  __nv_bfloat16* v_app = nullptr;
  __nv_bfloat16* out = nullptr;      // [B, kD]
};
cudaError_t launch_row_attn(const RowAttnParams& p, cudaStream_t stream);

// Cross-attention of one decode position straight from the encoder memory ("absorbed" projections):
//   scores_h[j] = (W_k,h^T q_h) . m_j / sqrt(dh)      (the key bias adds a per-head constant: softmax-invariant)
//   u_h        = sum_j softmax(scores_h)[j] m_j       (the value projection W_v,h u_h + b_v,h follows as a GEMM)
// so a decode position reads the memory rows (512 B each) once from HBM instead of a K row and a V row, and the
// per-question K|V projection of the memory disappears.  qp [B, nhead*256] bf16 = absorbed queries W_k,h^T q_h,
// memory [B*rows_per_q, 256] bf16 (through the tensor map), out u [B, nhead*256] bf16.
struct MemAttnParams {
  int B = 0, nhead = 4;
  bool pdl = false;
  const __nv_bfloat16* qp = nullptr;
  long long rows_per_q = 0;
  const int32_t* lens = nullptr;     // [B] or null -> const_len
  int const_len = 0;
  __nv_bfloat16* out = nullptr;
  // the memory rows are fetched evict-first: every decode position of every layer streams the same rows again, but the
  // rows of all the chains in flight (254 MB at 2 x 1024 questions) never fit L2, so keeping them only displaces what
  // does get re-used - weights, KV caches, the FFN partials
  bool l2_evict_first = true;
};
constexpr int kMemAttnTileRows = 32;  // memory rows per shared-memory tile of the absorbed cross-attention
// tm_mem: the memory as a 2D tensor [B * rows_per_q, 256] bf16, 128-byte swizzle, box {64 channels, kMemAttnTileRows}
cudaError_t launch_mem_attn(const CUtensorMap& tm_mem, const MemAttnParams& p, cudaStream_t stream);

// Weight packing for the absorbed cross-attention: in_proj_weight [3d, d] / in_proj_bias [3d] fp32 ->
//   w_qk [nhead*d, d] bf16, row h*d + i = sum_e W_k[h*dh+e][i] * W_q[h*dh+e][:]     b_qk [nhead*d] likewise with b_q
cudaError_t launch_absorb_qk(const float* in_proj_weight, const float* in_proj_bias, int nhead, __nv_bfloat16* w_qk,
                             float* b_qk, cudaStream_t stream);

// ... and for the other side of the attention: the per-head value projection followed by out_proj is one linear map of
// the attention-weighted memory u [nhead*d]:  out_proj(concat_h(W_v,h u_h + b_v,h)) = W_ov u + b_ov with
//   w_ov [d, nhead*d] bf16, w_ov[n][h*d + i] = sum_e W_o[n][h*dh+e] * W_v[h*dh+e][i]     b_ov = b_o + W_o b_v
cudaError_t launch_absorb_ov(const float* in_proj_weight, const float* in_proj_bias, const float* out_proj_weight,
                             const float* out_proj_bias, int nhead, __nv_bfloat16* w_ov, float* b_ov,
                             cudaStream_t stream);

// tok[B, tok_ld] -> caller layout.
//   out_i64 : out[b*out_ld + j] = tok[b, src_col0 + j], j < n_cols                       (IQAP programs, FA `ys`)
//   out_i32 : FA step cache row: out[b*out_ld + j] = (j == 0 || !forced) ? tok[b, j] : forced[b*forced_ld + j-1],
//             skipped for questions with n_steps[b] <= step (the reference never executes those steps)
struct PublishParams {
  int B = 0, n_cols = 0, src_col0 = 0;
  const int64_t* tok = nullptr;
  int tok_ld = 0;
  int64_t* out_i64 = nullptr;
  int32_t* out_i32 = nullptr;
  long long out_ld = 0;
  const int64_t* forced = nullptr;
  int forced_ld = 0;
  const int32_t* n_steps = nullptr;
  int step = 0;
};
cudaError_t launch_publish_tokens(const PublishParams& p, cudaStream_t stream);

// Decode-phase feed-forward block, hidden dimension split over CTAs (ffn_small.cu):
//   out = LayerNorm(residual + W2 relu(W1 x + b1) + b2);  x, residual, out bf16 [M, kD]; partial fp32 [ff/128, M, kD]
// tm_x: x [M, kD] box 128 rows; tm_w1: W1 [ff, kD] box 128 rows; tm_w2: W2 [kD, ff] box 256 rows (all 128B-swizzled)
struct FfnSmallParams {
  int M = 0, ff = 0, n_slices = 0;
  bool pdl = false;
  const float* b1 = nullptr;
  const float* b2 = nullptr;
  const __nv_bfloat16* residual = nullptr;
  const float* gamma = nullptr;
  const float* beta = nullptr;
  float eps = 1e-5f;
  float* partial = nullptr;
  __nv_bfloat16* out = nullptr;
  float* out_f32 = nullptr;  // optional fp32 copy of the output ...
  const float* fn_gamma = nullptr;  // ... after a second LayerNorm when given (nn.Transformer's decoder.norm, FA:42)
  const float* fn_beta = nullptr;
  // Optional vocabulary head of this decode position fused into the reduce kernel (last decoder layer, vocabularies of
  // up to 64 entries): logits = y W_head^T + b in fp32 on the CUDA cores (the row is already in the warp's registers),
  // argmax (lowest index wins ties) -> tok[row, t+1], next input x_next[row] = emb[next] + pe[t+1]; same meaning as
  // the kEpiHead GEMM epilogue's fields (IQAP:230-236)
  const float* head_w = nullptr;     // [V, 256] fp32
  const float* head_b = nullptr;
  int head_V = 0, head_t = 0;
  int64_t* tok = nullptr;
  int tok_ld = 0;
  float* logits = nullptr;
  int logits_T = 0;
  const int64_t* forced = nullptr;
  int forced_ld = 0;
  const float* emb = nullptr;
  int vocab = 0;
  const float* pe_next = nullptr;
  __nv_bfloat16* x_next = nullptr;
};
cudaError_t launch_ffn_small(const CUtensorMap& tm_x, const CUtensorMap& tm_w1, const CUtensorMap& tm_w2,
                             const FfnSmallParams& p, cudaStream_t stream);

// image_proj on CTA pairs (image_proj_pair.cu): the kEpiBiasPeRemap GEMM with N = 256 where each CTA of a pair loads half
// of every weight k-block.  fmt: 0 = fp32 features (tf32 MMA), 1 = fp16, 2 = bf16; p as for launch_gemm.
// a_evict_first: the feature rows (read exactly once) are fetched with the L2 evict-first policy.
cudaError_t launch_image_proj_pair(int fmt, const CUtensorMap& tm_a, const CUtensorMap& tm_w, const GemmParams& p,
                                   cudaStream_t stream, bool a_evict_first = true);

// Host side (host_convert.cu): fp32 -> fp16 (round to nearest even) on a pool of worker threads; dst 32-byte aligned.
void host_f32_to_f16(const float* src, void* dst, size_t n, int threads);
void host_f32_to_bf16(const float* src, void* dst, size_t n, int threads);
int host_convert_threads();

// Encoder feed-forward block + norm2 (+ the final encoder norm) in one persistent kernel (enc_ffn_fused.cu): the
// hidden activations stay in TMEM / shared memory.  tm_x: [M, 256] box 64 x 128 (also the residual), tm_w1: [ff, 256]
// box 64 x 64, tm_w2: [256, ff] box 64 x 128 (a CTA pair splits every weight tile).
struct EncFfnParams {
  int M = 0, n_slices = 0;  // ff = 128 * n_slices
  const float* b1 = nullptr;
  const float* b2 = nullptr;
  const __nv_bfloat16* residual = nullptr;  // = the input rows, row pitch 256
  const float* gamma = nullptr;
  const float* beta = nullptr;
  const float* gamma2 = nullptr;  // optional second LayerNorm
  const float* beta2 = nullptr;
  float eps = 1e-5f;
  __nv_bfloat16* out = nullptr;
  long long* dbg = nullptr;  // optional: wait-cycle counters of CTA 0 (B200VQA_ENC_FFN_DBG)
};
cudaError_t launch_enc_ffn_fused(const CUtensorMap& tm_x, const CUtensorMap& tm_w1, const CUtensorMap& tm_w2,
                                 const EncFfnParams& p, cudaStream_t stream);

// ------------------------------------------------------------------------------------------
// Persistent decode (decode_persist.cu): ALL positions x layers of a greedy decode in one launch.
//
// A cluster of 8 CTAs owns a tile of 64 questions for the whole decode (questions never interact, so clusters never
// synchronise with each other).  Inside the cluster every dense layer is split by output columns over the 8 CTAs
// (tcgen05, M = 64 questions, weights streamed from L2 through a TMA ring, accumulator in TMEM), every per-question
// operation (self-attention over the KV cache, LayerNorm, the absorbed cross-attention over the encoder memory, the
// vocabulary head + argmax + next embedding) is split by rows: CTA r owns questions 8r..8r+7, one warp each.  The
// hand-offs between the two splits go through the (L2-resident) activation scratch below and a cluster-wide mbarrier
// rendezvous - what used to be a kernel boundary (IQAP:208-236, FA:137-145).
// ------------------------------------------------------------------------------------------
constexpr int kDecPersistMaxSteps = 32;  // the warp-per-question self-attention keeps <= 32 keys in registers

// per-layer fp32 vectors (device pointers), one entry per decoder layer, array in device memory
struct DecLayerDev {
  const float *b_in, *b_out, *n1w, *n1b;   // self-attention in_proj / out_proj biases, norm1
  const float *b_qk, *b_v, *b_co, *n2w, *n2b;  // absorbed query bias, value bias, cross out_proj bias, norm2
  const float *b1, *b2, *n3w, *n3b;        // feed-forward biases, norm3
};

struct DecPersistParams {
  int B = 0, steps = 0, n_layers = 0, nhead = 4, ff = 0;
  // weight slab A: rows of 256 bf16 (K-major); per layer [in_proj 768 | out_proj 256 | w_qk nhead*256 | w_v 256 |
  // cross out_proj 256 | linear1 ff]; slab B: linear2 as k-blocks, row (layer * ff/64 + kb) * 256 + n holds
  // W2[n][64 kb .. 64 kb + 63]
  int rows_per_layer = 0;
  const DecLayerDev* layers = nullptr;
  const int32_t* lens = nullptr;  // [B] memory rows per question, or null -> const_len
  int const_len = 0;
  // activation scratch, rows indexed by question (library workspace)
  __nv_bfloat16 *dx = nullptr, *dqkv = nullptr, *dattn = nullptr, *dx1 = nullptr, *dq = nullptr, *du = nullptr,
                *dx2 = nullptr, *dxo[2] = {nullptr, nullptr};
  float* dpre = nullptr;      // [rows, 256] pre-LayerNorm sums
  float* partial = nullptr;   // [8][part_rows][256] feed-forward partial sums
  long long part_rows = 0;
  __nv_bfloat16* const* kc = nullptr;  // [n_layers] device array of self-attention key caches [rows, t_max, 256]
  __nv_bfloat16* const* vc = nullptr;
  int t_max = 0;
  float eps = 1e-5f;
  const float *fn_gamma = nullptr, *fn_beta = nullptr;  // nn.Transformer's final decoder norm (FA:42), or null
  // vocabulary head (fp32, CUDA cores: the row is already in the owning warp's registers)
  const float *head_w = nullptr, *head_b = nullptr;
  int head_V = 0;
  int64_t* tok = nullptr;
  int tok_ld = 0;
  float* logits = nullptr;
  int logits_T = 0;
  const int64_t* forced = nullptr;
  int forced_ld = 0;
  const float *emb = nullptr, *pe = nullptr;  // decoder embedding [vocab, 256], positional table [*, 256]
  int vocab = 0;
  int stagger_cycles = 0;   // start delay of odd tiles: de-phases the clusters' HBM-bound and latency-bound phases
  long long* dbg_clk = nullptr;  // optional [8][24] int64: timeline of CTA (0, 0) over the first 8 stages (tools/)
  int dbg_stop = -1;        // k >= 1: return after the k-th cluster rendezvous of the first stage (tests: inspect the scratch)
};
// tm_wa: slab A [rows, 256] bf16, box {64, 32}; tm_wb: slab B [rows, 64] bf16, box {64, 32};
// tm_mem: encoder memory [B * kLP, 256] bf16, box {64, 16}   (all 128-byte swizzled)
cudaError_t launch_decode_persist(const CUtensorMap& tm_wa, const CUtensorMap& tm_wb, const CUtensorMap& tm_mem,
                                  const DecPersistParams& p, cudaStream_t stream);
// linear2 weight [256, ff] bf16 -> k-block-major slab B rows (see above)
cudaError_t launch_pack_w2_kblocks(const __nv_bfloat16* w2, __nv_bfloat16* out, int ff, cudaStream_t stream);

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device): the attribute is per device
cudaError_t ensure_dyn_smem(const void* kernel, int bytes);
// SM count of the CURRENT device (cached per device)
cudaError_t current_device_sms(int* sms);

// programs [B,T] i64 (prefix order) -> func [B,S], deps [B,S,2], n_steps [B] (all int32) in execution order
struct ProgToChainParams {
  int B = 0, T = 0, S = 0, prog_vocab = 0;
  const int64_t* programs = nullptr;
  const int32_t* arity = nullptr;     // [prog_vocab]: inputs of each token (0 / 1 / 2), negative = end of program
  const int32_t* func_map = nullptr;  // [prog_vocab]: chain function token of each program token
  int32_t* func = nullptr;
  int32_t* deps = nullptr;
  int32_t* n_steps = nullptr;
};
cudaError_t launch_programs_to_chain(const ProgToChainParams& p, cudaStream_t stream);

// Evaluation tally (inference_transformer_iqap_tally.py:317-344); counts[4] u64 are accumulated, not reset.
struct TallyParams {
  int B = 0, classes = 0, T = 0;
  const float* answer_logits = nullptr;   // [B, classes]
  const int64_t* programs = nullptr;      // [B, T] generated
  const int64_t* gt_answers = nullptr;    // [B]
  const int64_t* gt_programs = nullptr;   // [B, T]
  unsigned long long* counts = nullptr;   // {both, answer only, program only, neither}
  int32_t* pred_answers = nullptr;        // optional [B]: argmax of the answer logits
};
cudaError_t launch_tally(const TallyParams& p, cudaStream_t stream);

// Image rows of each question's block copied from per-image tokens (several questions per image).
cudaError_t launch_gather_image_rows(const __nv_bfloat16* img_tok, const int32_t* image_idx, int n_img, int n_img_tokens,
                                     int B, __nv_bfloat16* x, cudaStream_t stream);

// dst[order[r]] = src[r] for rows of `width` int32 (results of a sorted batch back into the caller's order)
cudaError_t launch_scatter_rows_i32(const int32_t* src, const int32_t* order, int n, int width, int32_t* dst,
                                    cudaStream_t stream);

cudaError_t launch_delay(long long cycles, cudaStream_t stream);

// Weight packing (run once per b200vqa_create / refresh): fp32 -> bf16 cast, fp32 [R,C] -> [C,R] transpose.
cudaError_t launch_cast_bf16(const float* in, __nv_bfloat16* out, size_t n, cudaStream_t stream);
cudaError_t launch_cast_f16(const float* in, __half* out, size_t n, cudaStream_t stream);
cudaError_t launch_transpose_f32(const float* in, float* out, int R, int C, cudaStream_t stream);

}  // namespace b200vqa
