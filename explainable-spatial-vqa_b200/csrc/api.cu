// C-ABI layer of libb200vqa.so (include/b200vqa.h): handle lifetime, weight packing, the activation
// workspace in HBM, and the kernel sequences that replace
//   VQAModel.forward + autoregressive_program_generation          (IQAP:136-241)
//   MultiModalTransformer.forward / greedy_decode                 (FA:45-58, 126-146)
//   run_inference_chain with the inference cache resident in HBM  (FA:83-124)
// Everything is enqueued on the caller's stream; only the *_host entry points synchronise.
//
// HBM layout (per handle, sized for `cap` questions, all row-major):
//   x, attn, x1, mem   bf16 [cap*256, 256]   encoder rows of question b live at [b*256, b*256+len)
//   qkv                bf16 [cap*256, 768]   packed q|k|v like torch's in_proj
//   hid                bf16 [cap*256, ff]
//   ckv[l]             bf16 [cap*256, 512]   decoder layer l cross-attention K|V of the memory (once per question)
//   kc[l], vc[l]       bf16 [cap, T, 256]    decoder self-attention KV cache
//   d*                 bf16 [cap, *]         one decode position per question
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <tuple>
#include <vector>

#include <nvtx3/nvToolsExt.h>  // header-only; a no-op unless a profiler injects the NVTX library

#include "host_util.h"
#include "kernels.h"

using namespace b200vqa;

namespace {

constexpr int kMaxDecodeLen = 64;
constexpr int kTokLd = kMaxDecodeLen + 1;

struct MhaPacked {
  __nv_bfloat16* w_in = nullptr;   // [3d, d]
  float* b_in = nullptr;           // [3d]
  __nv_bfloat16* w_out = nullptr;  // [d, d]
  float* b_out = nullptr;
};
struct LayerPacked {
  MhaPacked self_attn, cross_attn;
  __nv_bfloat16* w1 = nullptr;  // [ff, d]
  float* b1 = nullptr;
  __nv_bfloat16* w2 = nullptr;  // [d, ff]
  float* b2 = nullptr;
  float *n1w = nullptr, *n1b = nullptr, *n2w = nullptr, *n2b = nullptr, *n3w = nullptr, *n3b = nullptr;
  // decoder layers: absorbed cross-attention query weights, row h*d + i = W_k,h^T W_q,h (kernels.h: MemAttnParams)
  __nv_bfloat16* w_qk = nullptr;  // [nhead*d, d]
  float* b_qk = nullptr;          // [nhead*d]
  // ... and the value + output projections as one map of the attention-weighted memory (kernels.h: launch_absorb_ov)
  __nv_bfloat16* w_ov = nullptr;  // [d, nhead*d]
  float* b_ov = nullptr;          // [d]
};

// Bump allocator over one cudaMalloc'ed arena (first pass measures, second pass assigns).
struct Arena {
  uint8_t* base = nullptr;
  size_t off = 0;
  template <typename T>
  T* take(size_t n) {
    off = (off + 255) & ~size_t(255);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

struct Workspace {
  int cap = 0;
  int t_max = 0;
  uint8_t* base = nullptr;
  size_t bytes = 0;
  __nv_bfloat16 *x = nullptr, *qkv = nullptr, *attn = nullptr, *x1 = nullptr, *hid = nullptr, *mem = nullptr;
  int32_t* lens = nullptr;
  std::vector<__nv_bfloat16*> ckv, kc, vc;
  __nv_bfloat16* du = nullptr;     // absorbed cross-attention output [drows, nhead*256]
  __nv_bfloat16 *dx = nullptr, *dqkv = nullptr, *dattn = nullptr, *dx1 = nullptr, *dq = nullptr, *dx2 = nullptr,
                *dhid = nullptr, *dxo[2] = {nullptr, nullptr};
  float* dout = nullptr;
  float* ffn_partial = nullptr;    // [ff/128, drows, 256] fp32 partial sums of the decode feed-forward block
  int64_t* tok = nullptr;          // [cap, kTokLd] greedy tokens of the running decode (column 0 = start token)
  __nv_bfloat16* img_t = nullptr;  // FA: transposed + cast image features [cap*196, 1024]
  __nv_bfloat16** kc_dev = nullptr;  // device copies of the kc / vc pointer tables (persistent decode kernel)
  __nv_bfloat16** vc_dev = nullptr;
  size_t drows = 0;                // decode rows (cap rounded up to 128)
};

// NVTX ranges around the phases of a call (encoder, decode, image projection, ...) when B200VQA_NVTX=1: they name the
// spans of a profiler timeline; costs nothing when no tool is attached and is compiled to one branch otherwise.
struct NvtxRange {
  bool on;
  NvtxRange(bool enabled, const char* name) : on(enabled) {
    if (on) nvtxRangePushA(name);
  }
  ~NvtxRange() {
    if (on) nvtxRangePop();
  }
};

using TmapKey = std::tuple<const void*, int, uint64_t, uint64_t, uint64_t, uint32_t, uint32_t>;

// kernel classes for the built-in profiler (b200vqa_profile_*)
enum Tag : int {
  kTagEmbed = 0, kTagImgProj, kTagEncQkv, kTagEncAttn, kTagEncOutLn, kTagEncFfn1, kTagEncFfn2Ln, kTagEncFfnFused, kTagEncFinalLn,
  kTagAnswer, kTagDecCrossKv, kTagDecSelfQkv, kTagDecCrossQ, kTagDecCrossV, kTagDecGemmLn, kTagDecFfn, kTagDecSelfAttn, kTagDecCrossAttn, kTagDecHead, kTagDecPersist, kTagMisc, kNumTags
};
const char* const kTagNames[kNumTags] = {
    "embed_gather", "image_proj_gemm", "enc_qkv_gemm", "enc_attention", "enc_outproj_ln_gemm", "enc_ffn1_gemm",
    "enc_ffn2_ln_gemm", "enc_ffn_fused", "enc_final_ln", "answer_head", "dec_cross_kv_gemm", "dec_self_qkv_gemm", "dec_cross_q_gemm", "dec_cross_v_gemm", "dec_outproj_ln_gemm", "dec_ffn_split", "dec_self_attention",
    "dec_cross_attention", "dec_head_argmax", "dec_persistent", "misc"};

struct ProfRec {
  int tag;
  cudaEvent_t a, b;
};

using GraphKey = std::tuple<int, int, int, int, const void*, const void*>;
struct GraphEntry {
  cudaGraphExec_t exec = nullptr;
  uint64_t launches = 0;  // kernel nodes in the graph (for b200vqa_launch_count)
};

}  // namespace

struct b200vqa_handle {
  int device = 0;
  int num_sms = 148;
  b200vqa_model_desc d{};
  std::vector<b200vqa_encoder_layer_weights> enc_src;
  std::vector<b200vqa_decoder_layer_weights> dec_src;

  uint8_t* wbase = nullptr;
  size_t wbytes = 0;
  float* img_w_f32 = nullptr;           // IQAP: tf32 operand [d, 1024]
  __half* img_w_f16 = nullptr;          // IQAP: fp16 operand for the fp16 feature store (same 10-bit mantissa as tf32)
  __nv_bfloat16* img_tok = nullptr;     // IQAP indexed forward: image_proj + PE of the unique images [n_img, 196, d]
  size_t img_tok_cap = 0;               // images the buffer holds
  __nv_bfloat16* img_w_bf16 = nullptr;  // FA: bf16 operand
  float *img_b = nullptr, *cls = nullptr, *enc_emb = nullptr, *dec_emb = nullptr, *pe_enc = nullptr, *pe_dec = nullptr;
  std::vector<LayerPacked> enc, dec;
  float *enc_fn_w = nullptr, *enc_fn_b = nullptr, *dec_fn_w = nullptr, *dec_fn_b = nullptr;
  float *head_w = nullptr, *head_b = nullptr;   // [V, d] fp32 (tf32 tensor-core operand)
  float *ans_w0t = nullptr, *ans_b0 = nullptr, *ans_w1 = nullptr, *ans_b1 = nullptr;

  Workspace ws;
  std::map<TmapKey, CUtensorMap> tmaps;
  uint64_t launches = 0;
  int cur_tag = kTagMisc;
  bool use_graphs = true;
  bool pdl_chain = false;  // set while the decode loop is being enqueued: its kernels form a PDL chain
  bool absorb = true;          // cross-attention reads the encoder memory directly (B200VQA_NO_ABSORB=1: K|V rows)
  int stagger_us = 0;          // start delay of every other decode branch (B200VQA_BRANCH_STAGGER_US)
  int stagger_mod = 2;
  bool absorb_ov = false;      // B200VQA_ABSORB_OV=1: value + output projection of the absorbed cross-attention folded into
                               // ONE K = nhead*256 LayerNorm GEMM instead of grouped value GEMM + out_proj LayerNorm GEMM
                               // (one launch fewer per layer and position, but four CTAs stream 4x the weight bytes:
                               // measured 4.57 vs 4.47 ms per step, so off by default)
  bool enc_attn_whole_head = false;  // B200VQA_ENC_ATTN_WHOLE_HEAD=1: dh = 64 encoder attention with one CTA per (question, head)
  bool no_fused_final_ln = false;  // B200VQA_NO_FUSED_FINAL_LN=1: nn.Transformer's final encoder norm as its own kernel (A/B runs)
  bool no_warp_self_attn = false;  // B200VQA_NO_WARP_SELF_ATTN=1: decoder self-attention with one CTA per question (A/B runs)
  bool no_fused_head = false;  // B200VQA_NO_FUSED_HEAD=1: vocabulary head as its own tf32 tensor-core GEMM even for vocabularies
                               // of up to 64 entries (A/B runs)
  bool no_ln_cluster = false;  // B200VQA_NO_LN_CLUSTER=1: decode LayerNorm GEMMs on the persistent kernel (A/B runs)
  bool nvtx = false;                  // B200VQA_NVTX=1
  int iqap_start_token = 1;           // Config.SPECIAL_TOKEN_ID (IQAP:24,205); b200vqa_set_start_token
  // b200vqa_set_host_upload(1): fp32 host features are rounded to fp16 on host threads before they cross PCIe
  bool host_upload_f16 = false;
  void* pin16[2] = {nullptr, nullptr};  // pinned fp16 staging of one chunk each
  size_t pin16_bytes = 0;
  cudaEvent_t ev_pin[2] = {nullptr, nullptr};  // the upload out of pin16[i] has completed
  bool pin_pending[2] = {false, false};
  // persistent decode kernel (decode_persist.cu): all positions x layers in one launch
  bool decode_persist = false;        // B200VQA_DECODE=persist|chain
  int persist_stagger_us = 0;         // B200VQA_PERSIST_STAGGER_US: start delay of odd question tiles
  int persist_dbg_stop = -1;          // tests: stop after this many rendezvous of the first stage
  long long* persist_clk = nullptr;   // B200VQA_PERSIST_STAMPS=1: [8][16] stage timeline of CTA (0, 0) (tools/)
  __nv_bfloat16* slab_a = nullptr;    // decoder weights as rows of 256 bf16 (kernels.h: DecPersistParams)
  __nv_bfloat16* slab_b = nullptr;    // linear2 as k-blocks of [256 x 64]
  int slab_rows_per_layer = 0;
  DecLayerDev* dec_layers_dev = nullptr;
  std::map<GraphKey, GraphEntry> graphs;
  cudaStream_t cap_stream = nullptr;  // capture happens here: the caller's stream may be the legacy default stream
  int decode_branches = 8;            // concurrent question ranges inside the decode graph
  // L2 eviction hints on the two pure streams (decode memory rows, image features): B200VQA_NO_L2_HINTS=1 turns them off
  bool l2_hints = true;
  cudaStream_t br_stream[7] = {};
  cudaEvent_t br_done[7] = {};
  cudaEvent_t br_fork = nullptr;
  bool profiling = false;
  std::vector<ProfRec> prof;

  // host-buffer entry point
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
  uint8_t* stage = nullptr;
  size_t stage_bytes = 0;
  int dbg_skip = 0;  // B200VQA_DBG_SKIP: timing decomposition only (results are garbage): 1 cross-attention, 2 self-
                     // attention, 4 feed-forward, 8 out_proj+LN GEMMs, 16 plain GEMMs of the decode chain, 32 encoder,
                     // 64 the vocabulary-head GEMM
  bool no_img_proj_pair = false;  // B200VQA_NO_IMG_PROJ_PAIR=1: image_proj on the one-CTA-per-tile GEMM
  bool no_fused_enc_ffn = false;  // B200VQA_NO_FUSED_ENC_FFN=1: linear1 / linear2 as two GEMMs through HBM
  int small_bn = 64;  // narrowest n-tile of the plain GEMMs (B200VQA_SMALL_BN=128: A/B of fewer, wider decode CTAs)
  int32_t* h_tables = nullptr;  // pinned: the sorted program tables of one fa_run_chain_host call
  size_t h_tables_ints = 0;
  cudaEvent_t ev_tables = nullptr;
  bool tables_pending = false;
};

namespace {

// questions processed per pass of the kernel sequence (bounds the activation workspace: ~2.4 MB per IQAP
// question with ff = 2048, ~1.5 MB per FA question with ff = 512)
int default_cap(const b200vqa_handle* h) { return h->d.kind == B200VQA_MODEL_IQAP ? 2048 : 4096; }

// ---------------------------------------------------------------------------------------------
// weights
// ---------------------------------------------------------------------------------------------
void layout_mha(Arena& a, MhaPacked& m, int d) {
  m.w_in = a.take<__nv_bfloat16>(size_t(3) * d * d);
  m.b_in = a.take<float>(size_t(3) * d);
  m.w_out = a.take<__nv_bfloat16>(size_t(d) * d);
  m.b_out = a.take<float>(d);
}

void layout_weights(b200vqa_handle* h, Arena& a) {
  const auto& d = h->d;
  const int D = d.d_model;
  if (d.kind == B200VQA_MODEL_IQAP) {
    h->img_w_f32 = a.take<float>(size_t(D) * d.img_feat_dim);
    h->img_w_f16 = a.take<__half>(size_t(D) * d.img_feat_dim);
  } else h->img_w_bf16 = a.take<__nv_bfloat16>(size_t(D) * d.img_feat_dim);
  h->img_b = a.take<float>(D);
  h->cls = d.cls_token ? a.take<float>(D) : nullptr;
  h->enc_emb = a.take<float>(size_t(d.enc_vocab) * D);
  h->dec_emb = a.take<float>(size_t(d.dec_vocab) * D);
  h->pe_enc = a.take<float>(size_t(d.pe_enc_len) * D);
  h->pe_dec = a.take<float>(size_t(d.pe_dec_len) * D);
  h->enc.resize(d.n_enc_layers);
  h->dec.resize(d.n_dec_layers);
  for (auto& L : h->enc) {
    layout_mha(a, L.self_attn, D);
    L.w1 = a.take<__nv_bfloat16>(size_t(d.dim_ff) * D);
    L.b1 = a.take<float>(d.dim_ff);
    L.w2 = a.take<__nv_bfloat16>(size_t(d.dim_ff) * D);
    L.b2 = a.take<float>(D);
    L.n1w = a.take<float>(D); L.n1b = a.take<float>(D); L.n2w = a.take<float>(D); L.n2b = a.take<float>(D);
  }
  for (auto& L : h->dec) {
    layout_mha(a, L.self_attn, D);
    layout_mha(a, L.cross_attn, D);
    L.w_qk = a.take<__nv_bfloat16>(size_t(d.nhead) * D * D);
    L.b_qk = a.take<float>(size_t(d.nhead) * D);
    L.w_ov = a.take<__nv_bfloat16>(size_t(d.nhead) * D * D);
    L.b_ov = a.take<float>(D);
    L.w1 = a.take<__nv_bfloat16>(size_t(d.dim_ff) * D);
    L.b1 = a.take<float>(d.dim_ff);
    L.w2 = a.take<__nv_bfloat16>(size_t(d.dim_ff) * D);
    L.b2 = a.take<float>(D);
    L.n1w = a.take<float>(D); L.n1b = a.take<float>(D); L.n2w = a.take<float>(D); L.n2b = a.take<float>(D);
    L.n3w = a.take<float>(D); L.n3b = a.take<float>(D);
  }
  h->slab_rows_per_layer = 3 * D + D + d.nhead * D + D + D + d.dim_ff;
  h->slab_a = a.take<__nv_bfloat16>(size_t(d.n_dec_layers) * h->slab_rows_per_layer * D);
  h->slab_b = a.take<__nv_bfloat16>(size_t(d.n_dec_layers) * d.dim_ff * D);
  h->dec_layers_dev = a.take<DecLayerDev>(d.n_dec_layers);
  if (d.enc_final_norm_weight) { h->enc_fn_w = a.take<float>(D); h->enc_fn_b = a.take<float>(D); }
  if (d.dec_final_norm_weight) { h->dec_fn_w = a.take<float>(D); h->dec_fn_b = a.take<float>(D); }
  h->head_w = a.take<float>(size_t(D) * d.dec_vocab);
  h->head_b = a.take<float>(d.dec_vocab);
  if (d.kind == B200VQA_MODEL_IQAP) {
    h->ans_w0t = a.take<float>(size_t(D) * d.answer_hidden);
    h->ans_b0 = a.take<float>(d.answer_hidden);
    h->ans_w1 = a.take<float>(size_t(d.num_classes) * d.answer_hidden);
    h->ans_b1 = a.take<float>(d.num_classes);
  }
}

#define PACK_OK(expr)                       \
  do {                                      \
    cudaError_t e_ = (expr);                \
    if (e_ != cudaSuccess) return e_;       \
  } while (0)

cudaError_t copy_f32(float* dst, const float* src, size_t n, cudaStream_t s) {
  return cudaMemcpyAsync(dst, src, n * sizeof(float), cudaMemcpyDeviceToDevice, s);
}

cudaError_t pack_mha(const b200vqa_mha_weights& w, MhaPacked& m, int D, cudaStream_t s) {
  PACK_OK(launch_cast_bf16(w.in_proj_weight, m.w_in, size_t(3) * D * D, s));
  PACK_OK(copy_f32(m.b_in, w.in_proj_bias, size_t(3) * D, s));
  PACK_OK(launch_cast_bf16(w.out_proj_weight, m.w_out, size_t(D) * D, s));
  PACK_OK(copy_f32(m.b_out, w.out_proj_bias, D, s));
  return cudaSuccess;
}

cudaError_t pack_weights(b200vqa_handle* h, cudaStream_t s) {
  const auto& d = h->d;
  const int D = d.d_model;
  if (h->img_w_f32) PACK_OK(copy_f32(h->img_w_f32, d.image_proj_weight, size_t(D) * d.img_feat_dim, s));
  if (h->img_w_f16) PACK_OK(launch_cast_f16(d.image_proj_weight, h->img_w_f16, size_t(D) * d.img_feat_dim, s));
  if (h->img_w_bf16) PACK_OK(launch_cast_bf16(d.image_proj_weight, h->img_w_bf16, size_t(D) * d.img_feat_dim, s));
  PACK_OK(copy_f32(h->img_b, d.image_proj_bias, D, s));
  if (h->cls) PACK_OK(copy_f32(h->cls, d.cls_token, D, s));
  PACK_OK(copy_f32(h->enc_emb, d.enc_embedding, size_t(d.enc_vocab) * D, s));
  PACK_OK(copy_f32(h->dec_emb, d.dec_embedding, size_t(d.dec_vocab) * D, s));
  PACK_OK(copy_f32(h->pe_enc, d.pe_enc, size_t(d.pe_enc_len) * D, s));
  PACK_OK(copy_f32(h->pe_dec, d.pe_dec, size_t(d.pe_dec_len) * D, s));
  for (int l = 0; l < d.n_enc_layers; ++l) {
    const auto& w = h->enc_src[l];
    auto& L = h->enc[l];
    PACK_OK(pack_mha(w.self_attn, L.self_attn, D, s));
    PACK_OK(launch_cast_bf16(w.linear1_weight, L.w1, size_t(d.dim_ff) * D, s));
    PACK_OK(copy_f32(L.b1, w.linear1_bias, d.dim_ff, s));
    PACK_OK(launch_cast_bf16(w.linear2_weight, L.w2, size_t(d.dim_ff) * D, s));
    PACK_OK(copy_f32(L.b2, w.linear2_bias, D, s));
    PACK_OK(copy_f32(L.n1w, w.norm1_weight, D, s)); PACK_OK(copy_f32(L.n1b, w.norm1_bias, D, s));
    PACK_OK(copy_f32(L.n2w, w.norm2_weight, D, s)); PACK_OK(copy_f32(L.n2b, w.norm2_bias, D, s));
  }
  for (int l = 0; l < d.n_dec_layers; ++l) {
    const auto& w = h->dec_src[l];
    auto& L = h->dec[l];
    PACK_OK(pack_mha(w.self_attn, L.self_attn, D, s));
    PACK_OK(pack_mha(w.multihead_attn, L.cross_attn, D, s));
    PACK_OK(launch_absorb_qk(w.multihead_attn.in_proj_weight, w.multihead_attn.in_proj_bias, d.nhead, L.w_qk, L.b_qk,
                             s));
    PACK_OK(launch_absorb_ov(w.multihead_attn.in_proj_weight, w.multihead_attn.in_proj_bias,
                             w.multihead_attn.out_proj_weight, w.multihead_attn.out_proj_bias, d.nhead, L.w_ov, L.b_ov,
                             s));
    PACK_OK(launch_cast_bf16(w.linear1_weight, L.w1, size_t(d.dim_ff) * D, s));
    PACK_OK(copy_f32(L.b1, w.linear1_bias, d.dim_ff, s));
    PACK_OK(launch_cast_bf16(w.linear2_weight, L.w2, size_t(d.dim_ff) * D, s));
    PACK_OK(copy_f32(L.b2, w.linear2_bias, D, s));
    PACK_OK(copy_f32(L.n1w, w.norm1_weight, D, s)); PACK_OK(copy_f32(L.n1b, w.norm1_bias, D, s));
    PACK_OK(copy_f32(L.n2w, w.norm2_weight, D, s)); PACK_OK(copy_f32(L.n2b, w.norm2_bias, D, s));
    PACK_OK(copy_f32(L.n3w, w.norm3_weight, D, s)); PACK_OK(copy_f32(L.n3b, w.norm3_bias, D, s));
  }
  // the persistent decode kernel's view of the decoder weights: one K-major slab of 256-wide rows + linear2 k-blocks
  {
    std::vector<DecLayerDev> lay(d.n_dec_layers);
    auto copy_rows = [&](__nv_bfloat16* dst, const __nv_bfloat16* src, size_t rows) {
      return cudaMemcpyAsync(dst, src, rows * D * sizeof(__nv_bfloat16), cudaMemcpyDeviceToDevice, s);
    };
    for (int l = 0; l < d.n_dec_layers; ++l) {
      const auto& L = h->dec[l];
      __nv_bfloat16* base = h->slab_a + size_t(l) * h->slab_rows_per_layer * D;
      PACK_OK(copy_rows(base, L.self_attn.w_in, 3 * D));
      PACK_OK(copy_rows(base + size_t(3) * D * D, L.self_attn.w_out, D));
      PACK_OK(copy_rows(base + size_t(4) * D * D, L.w_qk, size_t(d.nhead) * D));
      PACK_OK(copy_rows(base + size_t(4 + d.nhead) * D * D, L.cross_attn.w_in + size_t(2) * D * D, D));
      PACK_OK(copy_rows(base + size_t(5 + d.nhead) * D * D, L.cross_attn.w_out, D));
      PACK_OK(copy_rows(base + size_t(6 + d.nhead) * D * D, L.w1, d.dim_ff));
      PACK_OK(launch_pack_w2_kblocks(L.w2, h->slab_b + size_t(l) * d.dim_ff * D, d.dim_ff, s));
      DecLayerDev& v = lay[l];
      v.b_in = L.self_attn.b_in; v.b_out = L.self_attn.b_out; v.n1w = L.n1w; v.n1b = L.n1b;
      v.b_qk = L.b_qk; v.b_v = L.cross_attn.b_in + 2 * D; v.b_co = L.cross_attn.b_out; v.n2w = L.n2w; v.n2b = L.n2b;
      v.b1 = L.b1; v.b2 = L.b2; v.n3w = L.n3w; v.n3b = L.n3b;
    }
    PACK_OK(cudaMemcpyAsync(h->dec_layers_dev, lay.data(), lay.size() * sizeof(DecLayerDev), cudaMemcpyHostToDevice, s));
    PACK_OK(cudaStreamSynchronize(s));  // `lay` is pageable host memory
  }
  if (h->enc_fn_w) {
    PACK_OK(copy_f32(h->enc_fn_w, d.enc_final_norm_weight, D, s));
    PACK_OK(copy_f32(h->enc_fn_b, d.enc_final_norm_bias, D, s));
  }
  if (h->dec_fn_w) {
    PACK_OK(copy_f32(h->dec_fn_w, d.dec_final_norm_weight, D, s));
    PACK_OK(copy_f32(h->dec_fn_b, d.dec_final_norm_bias, D, s));
  }
  PACK_OK(copy_f32(h->head_w, d.head_weight, size_t(d.dec_vocab) * D, s));
  PACK_OK(copy_f32(h->head_b, d.head_bias, d.dec_vocab, s));
  if (d.kind == B200VQA_MODEL_IQAP) {
    PACK_OK(launch_transpose_f32(d.answer_w0, h->ans_w0t, d.answer_hidden, D, s));
    PACK_OK(copy_f32(h->ans_b0, d.answer_b0, d.answer_hidden, s));
    PACK_OK(copy_f32(h->ans_w1, d.answer_w1, size_t(d.num_classes) * d.answer_hidden, s));
    PACK_OK(copy_f32(h->ans_b1, d.answer_b1, d.num_classes, s));
  }
  return cudaSuccess;
}

int validate_desc(const b200vqa_model_desc* d) {
  B200VQA_REQUIRE(d != nullptr, "model descriptor is NULL");
  B200VQA_REQUIRE(d->kind == B200VQA_MODEL_IQAP || d->kind == B200VQA_MODEL_FA, "unknown model kind %d", d->kind);
  if (d->d_model != kD) {
    set_error("d_model %d: the sm_100a kernels are specialised for d_model = %d", d->d_model, kD);
    return B200VQA_ERR_UNSUPPORTED_SHAPE;
  }
  if (d->nhead != 2 && d->nhead != 4) {
    set_error("nhead %d: supported head counts are 2 and 4 (head dim 128 / 64)", d->nhead);
    return B200VQA_ERR_UNSUPPORTED_SHAPE;
  }
  if (d->dim_ff <= 0 || d->dim_ff % 256 != 0 || d->img_feat_dim % 32 != 0 || d->n_img_tokens <= 0 ||
      d->n_img_tokens > 200) {
    set_error("unsupported shape: dim_ff %d (multiple of 256), img_feat_dim %d (multiple of 32), n_img_tokens %d",
              d->dim_ff, d->img_feat_dim, d->n_img_tokens);
    return B200VQA_ERR_UNSUPPORTED_SHAPE;
  }
  B200VQA_REQUIRE(d->n_enc_layers >= 1 && d->n_enc_layers <= 16 && d->n_dec_layers >= 1 && d->n_dec_layers <= 16,
                  "layer counts out of range (%d encoder, %d decoder)", d->n_enc_layers, d->n_dec_layers);
  if (d->dec_vocab > 256) {
    set_error("decoder vocabulary %d: the fused head kernel handles up to 256 entries (reference: 44 / 170)",
              d->dec_vocab);
    return B200VQA_ERR_UNSUPPORTED_SHAPE;
  }
  B200VQA_REQUIRE(d->enc_vocab > 0 && d->dec_vocab > 0 && d->pe_enc_len > 0 && d->pe_dec_len > 0,
                  "vocabulary / positional table sizes must be positive");
  B200VQA_REQUIRE(d->image_proj_weight && d->image_proj_bias && d->enc_embedding && d->dec_embedding && d->pe_enc &&
                      d->pe_dec && d->enc_layers && d->dec_layers && d->head_weight && d->head_bias,
                  "a required weight pointer is NULL");
  if (d->kind == B200VQA_MODEL_IQAP) {
    B200VQA_REQUIRE(d->cls_token && d->answer_w0 && d->answer_b0 && d->answer_w1 && d->answer_b1,
                    "IQAP needs cls_token and the answer classifier weights");
    B200VQA_REQUIRE(d->answer_hidden > 0 && d->answer_hidden <= 1024 && d->num_classes > 0,
                    "answer head sizes out of range (hidden %d, classes %d)", d->answer_hidden, d->num_classes);
    B200VQA_REQUIRE(d->answer_pool_rows >= 0 && d->answer_pool_rows <= d->n_img_tokens,
                    "answer_pool_rows %d out of range (0..%d image tokens)", d->answer_pool_rows, d->n_img_tokens);
    if (1 + d->n_img_tokens + d->max_q_len > kLP || d->pe_enc_len < 1 + d->n_img_tokens + d->max_q_len) {
      set_error("IQAP sequence 1+%d+%d exceeds %d rows or the positional table (%d)", d->n_img_tokens, d->max_q_len,
                kLP, d->pe_enc_len);
      return B200VQA_ERR_UNSUPPORTED_SHAPE;
    }
  } else {
    B200VQA_REQUIRE(d->pe_enc_len > d->n_img_tokens, "FA positional table shorter than the image tokens");
  }
  return B200VQA_OK;
}

// ---------------------------------------------------------------------------------------------
// workspace
// ---------------------------------------------------------------------------------------------
void layout_workspace(const b200vqa_handle* h, Workspace& w, Arena& a, int cap, int t_max) {
  const auto& d = h->d;
  const size_t rows = size_t(cap) * kLP;
  w.cap = cap;
  w.t_max = t_max;
  w.x = a.take<__nv_bfloat16>(rows * kD);
  w.qkv = a.take<__nv_bfloat16>(rows * 3 * kD);
  w.attn = a.take<__nv_bfloat16>(rows * kD);
  w.x1 = a.take<__nv_bfloat16>(rows * kD);
  w.hid = a.take<__nv_bfloat16>(rows * d.dim_ff);
  w.mem = a.take<__nv_bfloat16>(rows * kD);
  w.lens = a.take<int32_t>(cap);
  w.ckv.resize(d.n_dec_layers);
  w.kc.resize(d.n_dec_layers);
  w.vc.resize(d.n_dec_layers);
  for (int l = 0; l < d.n_dec_layers; ++l) {
    w.ckv[l] = h->absorb ? nullptr : a.take<__nv_bfloat16>(rows * 2 * kD);
    w.kc[l] = a.take<__nv_bfloat16>(size_t(cap) * t_max * kD);
    w.vc[l] = a.take<__nv_bfloat16>(size_t(cap) * t_max * kD);
  }
  // decode rows are padded to a whole 128-row tile so TMA boxes stay inside the allocation
  const size_t drows = (size_t(cap) + 127) / 128 * 128;
  w.dx = a.take<__nv_bfloat16>(drows * kD);
  w.dqkv = a.take<__nv_bfloat16>(drows * 3 * kD);
  w.dattn = a.take<__nv_bfloat16>(drows * kD);
  w.dx1 = a.take<__nv_bfloat16>(drows * kD);
  w.dq = a.take<__nv_bfloat16>(drows * d.nhead * kD);  // absorbed queries: nhead x 256 per row
  w.du = a.take<__nv_bfloat16>(drows * d.nhead * kD);
  w.dx2 = a.take<__nv_bfloat16>(drows * kD);
  w.dhid = a.take<__nv_bfloat16>(drows * d.dim_ff);
  w.dxo[0] = a.take<__nv_bfloat16>(drows * kD);
  w.dxo[1] = a.take<__nv_bfloat16>(drows * kD);
  w.dout = a.take<float>(drows * kD);
  w.drows = drows;
  w.ffn_partial = a.take<float>(size_t(std::max(d.dim_ff / 128, 8)) * drows * kD);
  w.kc_dev = a.take<__nv_bfloat16*>(d.n_dec_layers);
  w.vc_dev = a.take<__nv_bfloat16*>(d.n_dec_layers);
  w.tok = a.take<int64_t>(size_t(cap) * kTokLd);
  if (d.kind == B200VQA_MODEL_FA) w.img_t = a.take<__nv_bfloat16>(size_t(cap) * d.n_img_tokens * d.img_feat_dim);
}

int ensure_workspace(b200vqa_handle* h, int B, int t_max) {
  const int cap_want = std::min(B, default_cap(h));
  if (h->ws.base && h->ws.cap >= cap_want && h->ws.t_max >= t_max) return B200VQA_OK;
  const int cap = std::max(cap_want, h->ws.cap);
  const int tm = std::max(t_max, h->ws.t_max);
  if (h->ws.base) {
    B200VQA_CUDA_OK(cudaDeviceSynchronize());
    B200VQA_CUDA_OK(cudaFree(h->ws.base));
    h->ws = Workspace{};
    h->tmaps.clear();
    for (auto& kv : h->graphs) cudaGraphExecDestroy(kv.second.exec);
    h->graphs.clear();
  }
  Arena measure;
  Workspace tmp;
  layout_workspace(h, tmp, measure, cap, tm);
  uint8_t* base = nullptr;
  B200VQA_CUDA_OK(cudaMalloc(&base, measure.off + 256));
  // finite contents everywhere: padding rows feed tensor-core tiles whose results are discarded
  B200VQA_CUDA_OK(cudaMemset(base, 0, measure.off + 256));
  Arena a;
  a.base = base;
  layout_workspace(h, h->ws, a, cap, tm);
  h->ws.base = base;
  h->ws.bytes = measure.off + 256;
  B200VQA_CUDA_OK(cudaMemcpy(h->ws.kc_dev, h->ws.kc.data(), h->ws.kc.size() * sizeof(void*), cudaMemcpyHostToDevice));
  B200VQA_CUDA_OK(cudaMemcpy(h->ws.vc_dev, h->ws.vc.data(), h->ws.vc.size() * sizeof(void*), cudaMemcpyHostToDevice));
  return B200VQA_OK;
}

// ---------------------------------------------------------------------------------------------
// GEMM helpers
// ---------------------------------------------------------------------------------------------
// cached 2D operand tensor map: 128-byte swizzle, box = {128 bytes of the inner dimension, box_rows}
int get_tmap(b200vqa_handle* h, const void* base, TmapType type, uint64_t rows, uint64_t cols, uint64_t ld,
             uint32_t box_rows, CUtensorMap* out) {
  // the descriptor is returned BY VALUE (128 B): the cache may be emptied by a later lookup of the same call sequence
  TmapKey key{base, int(type), rows, cols, ld, box_rows, 0u};
  auto it = h->tmaps.find(key);
  if (it == h->tmaps.end()) {
    if (h->tmaps.size() > 4096) h->tmaps.clear();
    CUtensorMap m;
    int rc = make_tmap_2d(&m, base, type, rows, cols, ld, box_rows);
    if (rc != B200VQA_OK) return rc;
    it = h->tmaps.emplace(key, m).first;
  }
  *out = it->second;
  return B200VQA_OK;
}

#define RC_OK(expr)                   \
  do {                                \
    int rc_ = (expr);                 \
    if (rc_ != B200VQA_OK) return rc_; \
  } while (0)

#define LAUNCH_OK(h, expr)                                                                        \
  do {                                                                                            \
    ProfRec pr__{(h)->cur_tag, nullptr, nullptr};                                                 \
    if ((h)->profiling) {                                                                         \
      cudaEventCreate(&pr__.a);                                                                   \
      cudaEventCreate(&pr__.b);                                                                   \
      cudaEventRecord(pr__.a, s);                                                                 \
    }                                                                                             \
    cudaError_t e__ = (expr);                                                                     \
    if (e__ != cudaSuccess) {                                                                     \
      set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__);     \
      return B200VQA_ERR_CUDA;                                                                    \
    }                                                                                             \
    if ((h)->profiling) {                                                                         \
      cudaEventRecord(pr__.b, s);                                                                 \
      (h)->prof.push_back(pr__);                                                                  \
    }                                                                                             \
    ++(h)->launches;                                                                              \
  } while (0)

#define LAUNCH_OK_S(h, expr, stream_) \
  do {                               \
    cudaStream_t s = (stream_);      \
    LAUNCH_OK(h, expr);              \
  } while (0)

// out = epilogue(A[M,K] . W[N,K]^T + bias), A/W bf16 (or fp32 for tf32) row-major
int gemm(b200vqa_handle* h, int epi, bool tf32, const void* A, int M, int K, int lda, const void* W, int N,
         GemmParams p, cudaStream_t s, int a_cols = 0) {
  if (M <= 0) return B200VQA_OK;
  p.pdl = h->pdl_chain;
  const TmapType ty = tf32 ? TmapType::kF32 : TmapType::kBF16;
  // narrower tiles when there are too few 128-row tiles to occupy the SMs (the decode GEMMs, M = batch)
  int bn = 256;
  if (epi == kEpiBias || epi == kEpiBiasRelu) {
    const int tiles_m = (M + 127) / 128;
    if (tiles_m * (N / 256) < h->num_sms / 2 && N % 128 == 0) bn = 128;
    if (tiles_m * (N / 128) < h->num_sms / 2 && N % 64 == 0 && h->small_bn <= 64) bn = 64;
  }
  if (p.a_group_cols > 0) bn = 64;  // grouped GEMM: an n-tile must not straddle two heads
  if (epi == kEpiHead) bn = N;  // N = padded vocabulary (one n-tile); rows >= V of W are zero-filled by TMA
  // the decode chain's out_proj + LayerNorm (M = questions): a cluster of four CTAs per 128-row tile (64-column slices,
  // statistics exchanged through distributed shared memory) instead of one SM owning the whole 128 x 256 epilogue.
  // Chosen per call site, never by M: the two kernels round the statistics differently, and a question's result must
  // not depend on how many other questions share its batch.
  if (epi == kEpiBiasResLN && p.ln_cluster && (K == 256 || K == 512 || K == 1024) && N == 256 && !h->no_ln_cluster)
    bn = 64;
  CUtensorMap ta, tw;
  // rows of A are rounded up to whole tiles only virtually: TMA zero-fills rows >= M
  RC_OK(get_tmap(h, A, ty, uint64_t(M), uint64_t(a_cols > 0 ? a_cols : K), uint64_t(lda), 128, &ta));
  RC_OK(get_tmap(h, W, ty, uint64_t(epi == kEpiHead ? p.head_V : N), uint64_t(K), uint64_t(K), bn, &tw));
  p.M = M;
  p.N = N;
  p.K = K;
  LAUNCH_OK(h, launch_gemm(epi, tf32, bn, ta, tw, p, h->num_sms, s));
  return B200VQA_OK;
}

int gemm_bias(b200vqa_handle* h, bool relu, const __nv_bfloat16* A, int M, int K, const __nv_bfloat16* W, int N,
              const float* bias, __nv_bfloat16* out, cudaStream_t s) {
  GemmParams p;
  p.bias = bias;
  p.out = out;
  p.ldc = N;
  return gemm(h, relu ? kEpiBiasRelu : kEpiBias, false, A, M, K, K, W, N, p, s);
}

int gemm_res_ln(b200vqa_handle* h, const __nv_bfloat16* A, int M, int K, const __nv_bfloat16* W, const float* bias,
                const __nv_bfloat16* residual, const float* gamma, const float* beta, __nv_bfloat16* out,
                float* out_f32, cudaStream_t s, bool decode = false, const float* gamma2 = nullptr,
                const float* beta2 = nullptr) {
  GemmParams p;
  p.ln_cluster = decode;
  p.gamma2 = gamma2;
  p.beta2 = beta2;
  p.bias = bias;
  p.out = out;
  p.ldc = kD;
  p.residual = residual;
  p.ldr = kD;
  p.gamma = gamma;
  p.beta = beta;
  p.eps = h->d.layer_norm_eps;
  p.out_f32 = out_f32;
  return gemm(h, gamma2 ? kEpiBiasResLN2 : kEpiBiasResLN, false, A, M, K, K, W, kD, p, s);
}

// ---------------------------------------------------------------------------------------------
// encoder: x (bf16 rows, [B*256,256]) -> memory; returns the buffer that holds the memory
// ---------------------------------------------------------------------------------------------
int run_encoder(b200vqa_handle* h, int B, const int32_t* lens, int const_len, __nv_bfloat16** memory,
                cudaStream_t s) {
  NvtxRange nvtx_range(h->nvtx, "b200vqa encoder");
  Workspace& w = h->ws;
  const auto& d = h->d;
  const int M = B * kLP;
  __nv_bfloat16* in = w.x;
  __nv_bfloat16* out = w.mem;
  for (int l = 0; l < d.n_enc_layers; ++l) {
    if (h->dbg_skip & 32) break;
    const LayerPacked& L = h->enc[l];
    h->cur_tag = kTagEncQkv;
    RC_OK(gemm_bias(h, false, in, M, kD, L.self_attn.w_in, 3 * kD, L.self_attn.b_in, w.qkv, s));
    {
      h->cur_tag = kTagEncAttn;
      CUtensorMap tq, tkv;
      RC_OK(get_tmap(h, w.qkv, TmapType::kBF16, uint64_t(M), 3 * kD, 3 * kD, 128, &tq));
      RC_OK(get_tmap(h, w.qkv, TmapType::kBF16, uint64_t(M), 3 * kD, 3 * kD, 256, &tkv));
      EncAttnParams ap;
      ap.B = B;
      ap.nhead = d.nhead;
      ap.lens = lens;
      ap.const_len = const_len;
      ap.out = w.attn;
      ap.scale = 1.f / sqrtf(float(kD / d.nhead));
      ap.one_cta_per_head = h->enc_attn_whole_head;
      LAUNCH_OK(h, launch_enc_attention(tq, tkv, w.qkv, ap, s));
    }
    h->cur_tag = kTagEncOutLn;
    RC_OK(gemm_res_ln(h, w.attn, M, kD, L.self_attn.w_out, L.self_attn.b_out, in, L.n1w, L.n1b, w.x1, nullptr, s));
    // the last layer's norm2 also applies nn.Transformer's final encoder norm (FA:42) when there is one
    const bool fuse_final = l == d.n_enc_layers - 1 && h->enc_fn_w && !h->no_fused_final_ln;
    if (!h->no_fused_enc_ffn && d.dim_ff % 128 == 0) {
      // linear1 -> ReLU -> linear2 -> + residual -> LayerNorm in one kernel: the hidden rows never leave the SM
      CUtensorMap tx, tw1, tw2;
      RC_OK(get_tmap(h, w.x1, TmapType::kBF16, uint64_t(M), kD, kD, 128, &tx));
      // each CTA of the pair loads half of a weight tile: 64 of a slice's 128 hidden units / 128 of the 256 outputs
      RC_OK(get_tmap(h, L.w1, TmapType::kBF16, uint64_t(d.dim_ff), kD, kD, 64, &tw1));
      RC_OK(get_tmap(h, L.w2, TmapType::kBF16, kD, uint64_t(d.dim_ff), uint64_t(d.dim_ff), 128, &tw2));
      EncFfnParams fp;
      fp.M = M;
      fp.n_slices = d.dim_ff / 128;
      fp.b1 = L.b1;
      fp.b2 = L.b2;
      fp.residual = w.x1;
      fp.gamma = L.n2w;
      fp.beta = L.n2b;
      fp.gamma2 = fuse_final ? h->enc_fn_w : nullptr;
      fp.beta2 = fuse_final ? h->enc_fn_b : nullptr;
      fp.eps = d.layer_norm_eps;
      fp.out = out;
      h->cur_tag = kTagEncFfnFused;
      LAUNCH_OK(h, launch_enc_ffn_fused(tx, tw1, tw2, fp, s));
    } else {
      h->cur_tag = kTagEncFfn1;
      RC_OK(gemm_bias(h, true, w.x1, M, kD, L.w1, d.dim_ff, L.b1, w.hid, s));
      h->cur_tag = kTagEncFfn2Ln;
      RC_OK(gemm_res_ln(h, w.hid, M, d.dim_ff, L.w2, L.b2, w.x1, L.n2w, L.n2b, out, nullptr, s, false,
                        fuse_final ? h->enc_fn_w : nullptr, fuse_final ? h->enc_fn_b : nullptr));
    }
    std::swap(in, out);
  }
  // `in` now holds the last layer's output
  h->cur_tag = kTagEncFinalLn;
  if (h->enc_fn_w && h->no_fused_final_ln)
    LAUNCH_OK(h, launch_layernorm_rows(in, in, h->enc_fn_w, h->enc_fn_b, d.layer_norm_eps, M, s));
  *memory = in;
  return B200VQA_OK;
}

// ---------------------------------------------------------------------------------------------
// decoder: greedy decode of `steps` positions against `memory`
// ---------------------------------------------------------------------------------------------
struct DecodeIO {
  int start_token = 0;
  const int64_t* start_tokens = nullptr;  // per-question start (teacher-forced forward)
  int start_ld = 0;
  int steps = 0;
  float* logits = nullptr;    // [B, logits_T, V]
  int logits_T = 0;
  const int64_t* forced = nullptr;
  int forced_ld = 0;
};

// The launch sequence of one greedy decode: cross K|V projection of the memory, start embedding, then `steps`
// positions x layers.  Every argument is a library-owned buffer or a value in the graph key, so the same
// sequence can be captured once and replayed (run_decoder).
// Decode positions 0..steps-1 for the questions [b_lo, b_lo + B) of the current chunk on stream `s`: a chain of
// small kernels per position (PDL-linked).  `br` selects this branch's slice of the partial-sum buffer.
int enqueue_decode_rows(b200vqa_handle* h, int b_lo, int B, const __nv_bfloat16* memory, const int32_t* lens,
                        int const_len, const DecodeIO& io, cudaStream_t s) {
  Workspace& w = h->ws;
  const auto& d = h->d;
  const size_t r0 = size_t(b_lo);
  __nv_bfloat16* dx = w.dx + r0 * kD;
  __nv_bfloat16* dqkv = w.dqkv + r0 * 3 * kD;
  __nv_bfloat16* dattn = w.dattn + r0 * kD;
  __nv_bfloat16* dx1 = w.dx1 + r0 * kD;
  __nv_bfloat16* dq = w.dq + r0 * d.nhead * kD;
  __nv_bfloat16* du = w.du + r0 * d.nhead * kD;
  const __nv_bfloat16* mem_b = memory + r0 * kLP * kD;
  __nv_bfloat16* dx2 = w.dx2 + r0 * kD;
  __nv_bfloat16* dxo[2] = {w.dxo[0] + r0 * kD, w.dxo[1] + r0 * kD};
  float* dout = w.dout + r0 * kD;
  float* partial = w.ffn_partial + r0 * size_t(d.dim_ff / 128) * kD;
  int64_t* tok = w.tok + r0 * kTokLd;
  const int32_t* lens_b = lens ? lens + b_lo : nullptr;
  float* logits = io.logits ? io.logits + r0 * io.logits_T * d.dec_vocab : nullptr;
  const int64_t* forced = io.forced ? io.forced + r0 * io.forced_ld : nullptr;
  h->pdl_chain = true;
  struct PdlOff {
    b200vqa_handle* h;
    ~PdlOff() { h->pdl_chain = false; }
  } pdl_off{h};
  const bool fused_head = d.dec_vocab <= 64 && !h->no_fused_head;
  for (int t = 0; t < io.steps; ++t) {
    const __nv_bfloat16* in = dx;
    for (int l = 0; l < d.n_dec_layers; ++l) {
      const LayerPacked& L = h->dec[l];
      const bool last = l == d.n_dec_layers - 1;
      __nv_bfloat16* out = dxo[l & 1];
      __nv_bfloat16* kc = w.kc[l] + r0 * w.t_max * kD;
      __nv_bfloat16* vc = w.vc[l] + r0 * w.t_max * kD;
      const __nv_bfloat16* ckv = h->absorb ? nullptr : w.ckv[l] + r0 * kLP * 2 * kD;
      h->cur_tag = kTagDecSelfQkv;
      if (!(h->dbg_skip & 16)) RC_OK(gemm_bias(h, false, in, B, kD, L.self_attn.w_in, 3 * kD, L.self_attn.b_in, dqkv, s));
      RowAttnParams sp;
      sp.B = B;
      sp.nhead = d.nhead;
      sp.q = dqkv;
      sp.ldq = 3 * kD;
      sp.k = kc;
      sp.v = vc;
      sp.rows_per_q = w.t_max;
      sp.ld = kD;
      sp.const_len = t + 1;
      sp.new_k = dqkv + kD;
      sp.new_v = dqkv + 2 * kD;
      sp.ld_new = 3 * kD;
      sp.append_pos = t;
      sp.warp_form = !h->no_warp_self_attn;
      sp.k_app = kc;
      sp.v_app = vc;
      sp.out = dattn;
      sp.pdl = true;
      h->cur_tag = kTagDecSelfAttn;
      if (!(h->dbg_skip & 2)) LAUNCH_OK(h, launch_row_attn(sp, s));
      h->cur_tag = kTagDecGemmLn;
      if (!(h->dbg_skip & 8))
        RC_OK(gemm_res_ln(h, dattn, B, kD, L.self_attn.w_out, L.self_attn.b_out, in, L.n1w, L.n1b, dx1, nullptr, s, true));
      if (h->absorb) {
        // cross-attention on the encoder memory itself: absorbed queries (N = nhead*256), one pass over the memory
        // rows for all heads, then the per-head value projection as a grouped GEMM
        const int NHD = d.nhead * kD;
        h->cur_tag = kTagDecCrossQ;
        if (!(h->dbg_skip & 16)) RC_OK(gemm_bias(h, false, dx1, B, kD, L.w_qk, NHD, L.b_qk, dq, s));
        MemAttnParams mp;
        mp.B = B;
        mp.nhead = d.nhead;
        mp.qp = dq;
        mp.rows_per_q = kLP;
        mp.lens = lens_b;
        mp.const_len = const_len;
        mp.out = du;
        mp.pdl = true;
        mp.l2_evict_first = h->l2_hints;
        h->cur_tag = kTagDecCrossAttn;
        CUtensorMap tmem_map;
        RC_OK(get_tmap(h, mem_b, TmapType::kBF16, uint64_t(B) * kLP, kD, kD, kMemAttnTileRows, &tmem_map));
        if (!(h->dbg_skip & 1)) LAUNCH_OK(h, launch_mem_attn(tmem_map, mp, s));
        if (!h->absorb_ov && !(h->dbg_skip & 16)) {
          GemmParams vp;
          vp.bias = L.cross_attn.b_in + 2 * kD;
          vp.out = dattn;
          vp.ldc = kD;
          vp.a_group_cols = kD / d.nhead;
          h->cur_tag = kTagDecCrossV;
          RC_OK(gemm(h, kEpiBias, false, du, B, kD, NHD, L.cross_attn.w_in + size_t(2) * kD * kD, kD, vp, s, NHD));
        }
      } else {
        h->cur_tag = kTagDecCrossQ;
        RC_OK(gemm_bias(h, false, dx1, B, kD, L.cross_attn.w_in, kD, L.cross_attn.b_in, dq, s));
        RowAttnParams cp;
        cp.B = B;
        cp.nhead = d.nhead;
        cp.q = dq;
        cp.ldq = kD;
        cp.k = ckv;
        cp.v = ckv + kD;
        cp.rows_per_q = kLP;
        cp.ld = 2 * kD;
        cp.lens = lens_b;
        cp.const_len = const_len;
        cp.out = dattn;
        cp.pdl = true;
        h->cur_tag = kTagDecCrossAttn;
        LAUNCH_OK(h, launch_row_attn(cp, s));
      }
      h->cur_tag = kTagDecGemmLn;
      if (h->absorb && h->absorb_ov)  // x2 = LN2(x1 + W_ov u + b_ov): K = nhead * 256 streamed through the cluster kernel
        RC_OK(gemm_res_ln(h, du, B, d.nhead * kD, L.w_ov, L.b_ov, dx1, L.n2w, L.n2b, dx2, nullptr, s, true));
      else if (!(h->dbg_skip & 8))
        RC_OK(gemm_res_ln(h, dattn, B, kD, L.cross_attn.w_out, L.cross_attn.b_out, dx1, L.n2w, L.n2b, dx2, nullptr, s,
                          true));
      {
        // feed-forward block with the hidden dimension split over CTAs (ffn_small.cu): 2 launches
        CUtensorMap tx, tw1, tw2;
        RC_OK(get_tmap(h, dx2, TmapType::kBF16, uint64_t(B), kD, kD, 128, &tx));
        RC_OK(get_tmap(h, L.w1, TmapType::kBF16, uint64_t(d.dim_ff), kD, kD, 128, &tw1));
        RC_OK(get_tmap(h, L.w2, TmapType::kBF16, kD, uint64_t(d.dim_ff), uint64_t(d.dim_ff), 256, &tw2));
        FfnSmallParams fp;
        fp.M = B;
        fp.ff = d.dim_ff;
        fp.n_slices = d.dim_ff / 128;
        fp.b1 = L.b1;
        fp.b2 = L.b2;
        fp.residual = dx2;
        fp.gamma = L.n3w;
        fp.beta = L.n3b;
        fp.eps = d.layer_norm_eps;
        fp.partial = partial;
        fp.out = out;
        fp.out_f32 = last ? dout : nullptr;
        fp.fn_gamma = last ? h->dec_fn_w : nullptr;
        fp.fn_beta = last ? h->dec_fn_b : nullptr;
        fp.pdl = true;
        if (last && fused_head) {
          // vocabularies of up to 64 entries: the head runs inside the reduce kernel (fp32, the row is in registers)
          fp.head_w = h->head_w;
          fp.head_b = h->head_b;
          fp.head_V = d.dec_vocab;
          fp.head_t = t;
          fp.tok = tok;
          fp.tok_ld = kTokLd;
          fp.logits = logits;
          fp.logits_T = io.logits_T;
          fp.forced = forced;
          fp.forced_ld = io.forced_ld;
          fp.emb = h->dec_emb;
          fp.vocab = d.dec_vocab;
          fp.pe_next = (t + 1 < io.steps) ? h->pe_dec + size_t(t + 1) * kD : nullptr;
          fp.x_next = dx;
        }
        h->cur_tag = kTagDecFfn;
        if (!(h->dbg_skip & 4)) {
          LAUNCH_OK(h, launch_ffn_small(tx, tw1, tw2, fp, s));
          ++h->launches;  // two kernels
        }
      }
      in = out;
    }
    if (!fused_head) {
      // vocabulary head on the tensor cores (tf32 inputs, fp32 accumulate) with argmax + next embedding fused
      GemmParams hp;
      hp.bias = h->head_b;
      hp.head_V = d.dec_vocab;
      hp.head_t = t;
      hp.tok = tok;
      hp.tok_ld = kTokLd;
      hp.logits = logits;
      hp.logits_T = io.logits_T;
      hp.forced = forced;
      hp.forced_ld = io.forced_ld;
      hp.emb = h->dec_emb;
      hp.vocab = d.dec_vocab;
      hp.pe_next = (t + 1 < io.steps) ? h->pe_dec + size_t(t + 1) * kD : nullptr;
      hp.x_next = dx;
      h->cur_tag = kTagDecHead;
      if (!(h->dbg_skip & 64)) RC_OK(gemm(h, kEpiHead, true, dout, B, kD, kD, h->head_w, d.dec_vocab <= 64 ? 64 : 256, hp, s));
    }
  }
  return B200VQA_OK;
}

// Work that precedes the per-position chains: cross-attention K|V of the memory - once per question (the reference
// recomputes it every step) - and the start embedding.
int enqueue_decode_prologue(b200vqa_handle* h, int B, const __nv_bfloat16* memory, const DecodeIO& io,
                            cudaStream_t s) {
  Workspace& w = h->ws;
  const auto& d = h->d;
  const int M = B * kLP;
  h->cur_tag = kTagDecCrossKv;
  for (int l = 0; l < d.n_dec_layers && !h->absorb; ++l) {
    const MhaPacked& ca = h->dec[l].cross_attn;
    RC_OK(gemm_bias(h, false, memory, M, kD, ca.w_in + size_t(kD) * kD, 2 * kD, ca.b_in + kD, w.ckv[l], s));
  }
  DecEmbedParams ep;
  ep.B = B;
  ep.emb = h->dec_emb;
  ep.vocab = d.dec_vocab;
  ep.pe = h->pe_dec;
  ep.start_token = std::min(std::max(io.start_token, 0), d.dec_vocab - 1);
  ep.start_tokens = io.start_tokens;
  ep.start_ld = io.start_ld;
  ep.x = w.dx;
  ep.tok = w.tok;
  ep.tok_ld = kTokLd;
  h->cur_tag = kTagEmbed;
  LAUNCH_OK(h, launch_dec_embed_start(ep, s));
  return B200VQA_OK;
}

bool persist_eligible(const b200vqa_handle* h, const DecodeIO& io) {
  const auto& d = h->d;
  return h->decode_persist && h->absorb && io.steps <= kDecPersistMaxSteps && d.dim_ff % 512 == 0 && d.dim_ff <= 2048;
}

// All positions x layers in ONE launch: a cluster of 8 CTAs per 64 questions (decode_persist.cu).
int enqueue_decode_persist(b200vqa_handle* h, int B, const __nv_bfloat16* memory, const int32_t* lens, int const_len,
                           const DecodeIO& io, cudaStream_t s) {
  Workspace& w = h->ws;
  const auto& d = h->d;
  CUtensorMap ta, tb, tm;
  RC_OK(get_tmap(h, h->slab_a, TmapType::kBF16, uint64_t(d.n_dec_layers) * h->slab_rows_per_layer, kD, kD, 32, &ta));
  RC_OK(get_tmap(h, h->slab_b, TmapType::kBF16, uint64_t(d.n_dec_layers) * (d.dim_ff / 64) * kD, 64, 64, 32, &tb));
  RC_OK(get_tmap(h, memory, TmapType::kBF16, uint64_t(B) * kLP, kD, kD, 16, &tm));
  DecPersistParams p;
  p.B = B;
  p.steps = io.steps;
  p.n_layers = d.n_dec_layers;
  p.nhead = d.nhead;
  p.ff = d.dim_ff;
  p.rows_per_layer = h->slab_rows_per_layer;
  p.layers = h->dec_layers_dev;
  p.lens = lens;
  p.const_len = const_len;
  p.dx = w.dx; p.dqkv = w.dqkv; p.dattn = w.dattn; p.dx1 = w.dx1; p.dq = w.dq; p.du = w.du; p.dx2 = w.dx2;
  p.dxo[0] = w.dxo[0]; p.dxo[1] = w.dxo[1];
  p.dpre = w.dout;
  p.partial = w.ffn_partial;
  p.part_rows = (long long)w.drows;
  p.kc = w.kc_dev;
  p.vc = w.vc_dev;
  p.t_max = w.t_max;
  p.eps = d.layer_norm_eps;
  p.fn_gamma = h->dec_fn_w;
  p.fn_beta = h->dec_fn_b;
  p.head_w = h->head_w;
  p.head_b = h->head_b;
  p.head_V = d.dec_vocab;
  p.tok = w.tok;
  p.tok_ld = kTokLd;
  p.logits = io.logits;
  p.logits_T = io.logits_T;
  p.forced = io.forced;
  p.forced_ld = io.forced_ld;
  p.emb = h->dec_emb;
  p.pe = h->pe_dec;
  p.vocab = d.dec_vocab;
  if (h->persist_stagger_us > 0) {
    int khz = 0;
    B200VQA_CUDA_OK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, h->device));
    p.stagger_cycles = int((long long)h->persist_stagger_us * khz / 1000);
  }
  p.dbg_stop = h->persist_dbg_stop;
  p.dbg_clk = h->persist_clk;
  h->cur_tag = kTagDecPersist;
  LAUNCH_OK(h, launch_decode_persist(ta, tb, tm, p, s));
  return B200VQA_OK;
}

int enqueue_decoder(b200vqa_handle* h, int B, const __nv_bfloat16* memory, const int32_t* lens, int const_len,
                    const DecodeIO& io, cudaStream_t s) {
  RC_OK(enqueue_decode_prologue(h, B, memory, io, s));
  if (persist_eligible(h, io)) return enqueue_decode_persist(h, B, memory, lens, const_len, io, s);
  return enqueue_decode_rows(h, 0, B, memory, lens, const_len, io, s);
}

// Graph-capture form: the batch is split into independent branches (questions never interact) that the GPU runs
// concurrently, so one branch's latency-bound chain of small GEMMs overlaps another branch's HBM-bound
// cross-attention.
int enqueue_decoder_branched(b200vqa_handle* h, int B, const __nv_bfloat16* memory, const int32_t* lens, int const_len,
                             const DecodeIO& io, cudaStream_t s) {
  RC_OK(enqueue_decode_prologue(h, B, memory, io, s));
  int nbr = h->decode_branches;
  while (nbr > 1 && B / nbr < 128) --nbr;
  if (nbr <= 1) return enqueue_decode_rows(h, 0, B, memory, lens, const_len, io, s);
  for (int i = 0; i < nbr - 1; ++i) {
    if (!h->br_stream[i]) {
      B200VQA_CUDA_OK(cudaStreamCreateWithFlags(&h->br_stream[i], cudaStreamNonBlocking));
      B200VQA_CUDA_OK(cudaEventCreateWithFlags(&h->br_done[i], cudaEventDisableTiming));
    }
  }
  if (!h->br_fork) B200VQA_CUDA_OK(cudaEventCreateWithFlags(&h->br_fork, cudaEventDisableTiming));
  B200VQA_CUDA_OK(cudaEventRecord(h->br_fork, s));
  // branch boundaries on multiples of 128 rows so every branch starts on its own GEMM tile
  const int per = ((B + nbr - 1) / nbr + 127) / 128 * 128;
  int rc = B200VQA_OK;
  for (int i = 0; i < nbr && rc == B200VQA_OK; ++i) {
    const int lo = i * per, hi = std::min(B, lo + per);
    if (lo >= hi) break;
    cudaStream_t bs = i == 0 ? s : h->br_stream[i - 1];
    if (i > 0) B200VQA_CUDA_OK(cudaStreamWaitEvent(bs, h->br_fork, 0));
    if (h->stagger_us > 0 && (i % h->stagger_mod) != 0) {
      // identical branches run in lock step: all of them hit the HBM-bound cross-attention together and the
      // latency-bound small kernels together.  A start offset de-phases them so the two kinds of work overlap.
      int khz = 0;
      B200VQA_CUDA_OK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, h->device));
      h->cur_tag = kTagMisc;
      LAUNCH_OK_S(h, launch_delay((long long)(h->stagger_us) * (i % h->stagger_mod) * khz / 1000, bs), bs);
    }
    rc = enqueue_decode_rows(h, lo, hi - lo, memory, lens, const_len, io, bs);
    if (i > 0) {
      B200VQA_CUDA_OK(cudaEventRecord(h->br_done[i - 1], bs));
      B200VQA_CUDA_OK(cudaStreamWaitEvent(s, h->br_done[i - 1], 0));
    }
  }
  return rc;
}

// Greedy decode into ws.tok.  The plain (no logits / no teacher forcing) sequence depends only on
// (B, steps, start token, memory buffer, length source) and library-owned buffers, so it is captured into a
// CUDA graph on first use and replayed afterwards: ~20 launches per position collapse into one graph launch.
int run_decoder(b200vqa_handle* h, int B, const __nv_bfloat16* memory, const int32_t* lens, int const_len,
                const DecodeIO& io, cudaStream_t s) {
  NvtxRange nvtx_range(h->nvtx, "b200vqa greedy decode");
  const bool plain = !io.logits && !io.forced && !io.start_tokens;
  // the persistent kernel is two launches (start embedding + decode): nothing for a graph to collapse
  if (!plain || h->profiling || !h->use_graphs || persist_eligible(h, io))
    return enqueue_decoder(h, B, memory, lens, const_len, io, s);
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  B200VQA_CUDA_OK(cudaStreamIsCapturing(s, &cs));
  if (cs != cudaStreamCaptureStatusNone) return enqueue_decoder(h, B, memory, lens, const_len, io, s);
  // The step-wise executor calls this with the per-step count of still-active questions, which differs from step to
  // step and from batch to batch on a ragged data set: rounded up to a multiple of 128 rows (within the workspace) so
  // that a handful of graphs serve every count.  The surplus rows are earlier steps' (finite) leftovers: they are
  // decoded and never published.
  if (h->d.kind == B200VQA_MODEL_FA && lens == h->ws.lens) B = std::min(h->ws.cap, (B + 127) / 128 * 128);
  GraphKey key{B, io.steps, io.start_token, const_len, static_cast<const void*>(memory), static_cast<const void*>(lens)};
  auto it = h->graphs.find(key);
  if (it == h->graphs.end()) {
    // warm the tensor-map cache and the kernels' one-time attribute calls outside capture
    const uint64_t before = h->launches;
    cudaGraph_t graph = nullptr;
    if (!h->cap_stream) B200VQA_CUDA_OK(cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking));
    B200VQA_CUDA_OK(cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeThreadLocal));
    int rc = enqueue_decoder_branched(h, B, memory, lens, const_len, io, h->cap_stream);
    cudaError_t e = cudaStreamEndCapture(h->cap_stream, &graph);
    if (rc != B200VQA_OK) {
      if (graph) cudaGraphDestroy(graph);
      return rc;
    }
    if (e != cudaSuccess) {
      set_error("cudaStreamEndCapture failed: %s", cudaGetErrorString(e));
      return B200VQA_ERR_CUDA;
    }
    GraphEntry ge;
    ge.launches = h->launches - before;
    h->launches = before;
    e = cudaGraphInstantiate(&ge.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) {
      set_error("cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
      return B200VQA_ERR_CUDA;
    }
    if (h->graphs.size() >= 64) {
      for (auto& kv : h->graphs) cudaGraphExecDestroy(kv.second.exec);
      h->graphs.clear();
    }
    it = h->graphs.emplace(key, ge).first;
  }
  B200VQA_CUDA_OK(cudaGraphLaunch(it->second.exec, s));
  h->launches += it->second.launches;
  return B200VQA_OK;
}

int publish(b200vqa_handle* h, int B, int n_cols, int src_col0, int64_t* out_i64, int32_t* out_i32, long long out_ld,
            const int64_t* forced, int forced_ld, const int32_t* n_steps, int step, cudaStream_t s) {
  PublishParams pp;
  pp.B = B;
  pp.n_cols = n_cols;
  pp.src_col0 = src_col0;
  pp.tok = h->ws.tok;
  pp.tok_ld = kTokLd;
  pp.out_i64 = out_i64;
  pp.out_i32 = out_i32;
  pp.out_ld = out_ld;
  pp.forced = forced;
  pp.forced_ld = forced_ld;
  pp.n_steps = n_steps;
  pp.step = step;
  h->cur_tag = kTagMisc;
  LAUNCH_OK(h, launch_publish_tokens(pp, s));
  return B200VQA_OK;
}

int check_decode_len(const b200vqa_handle* h, int steps) {
  if (steps < 1 || steps > kMaxDecodeLen || steps > h->d.pe_dec_len) {
    set_error("decode length %d out of range (1..%d, positional table %d)", steps, kMaxDecodeLen, h->d.pe_dec_len);
    return B200VQA_ERR_BAD_ARGUMENT;
  }
  return B200VQA_OK;
}

int set_device(const b200vqa_handle* h) {
  B200VQA_CUDA_OK(cudaSetDevice(h->device));
  return B200VQA_OK;
}

// image_proj (+bias +PE of rows 1..196) of n_img unique images into h->img_tok [first .. first + n_img)
int project_unique_images(b200vqa_handle* h, const float* img, size_t first, int n_img, cudaStream_t s) {
  const auto& d = h->d;
  h->cur_tag = kTagImgProj;
  GemmParams p;
  p.bias = h->img_b;
  p.out = h->img_tok + first * d.n_img_tokens * kD;
  p.ldc = kD;
  p.rows_in = d.n_img_tokens;
  p.rows_out = d.n_img_tokens;
  p.row_off = 0;
  p.pe = h->pe_enc;
  p.pe_off = 1;
  return gemm(h, kEpiBiasPeRemap, true, img, n_img * d.n_img_tokens, d.img_feat_dim, d.img_feat_dim, h->img_w_f32, kD, p,
              s);
}

int ensure_img_tok(b200vqa_handle* h, size_t n_img) {
  if (h->img_tok_cap >= n_img) return B200VQA_OK;
  if (h->img_tok) {
    B200VQA_CUDA_OK(cudaDeviceSynchronize());
    B200VQA_CUDA_OK(cudaFree(h->img_tok));
    h->img_tok = nullptr;
    h->img_tok_cap = 0;
    h->tmaps.clear();
  }
  B200VQA_CUDA_OK(cudaMalloc(&h->img_tok, n_img * h->d.n_img_tokens * kD * sizeof(__nv_bfloat16)));
  h->img_tok_cap = n_img;
  return B200VQA_OK;
}

// img != null: one feature block per question; otherwise image rows are gathered from h->img_tok by image_idx
int iqap_chunk(b200vqa_handle* h, const float* img, const int64_t* q, int B, int T, float* answer, int64_t* programs,
               float* step_logits, const int64_t* forced, float* opt_memory, int B_total, int b0, cudaStream_t s,
               const int32_t* image_idx = nullptr, int n_img = 0, bool f16 = false) {
  Workspace& w = h->ws;
  const auto& d = h->d;
  const int S = 1 + d.n_img_tokens + d.max_q_len;
  h->cur_tag = kTagEmbed;
  LAUNCH_OK(h, launch_iqap_embed(q, B, d.max_q_len, h->cls, h->enc_emb, d.enc_vocab, h->pe_enc, d.n_img_tokens, w.x, s));
  if (!img) {
    LAUNCH_OK(h, launch_gather_image_rows(h->img_tok, image_idx, n_img, d.n_img_tokens, B, w.x, s));
  } else {
    // image_proj straight from the caller's fp32 features (tf32 tensor-core math), +bias +PE, written into
    // rows 1..196 of each question's block (this is the reference's torch.cat, IQAP:164)
    h->cur_tag = kTagImgProj;
    GemmParams p;
    p.bias = h->img_b;
    p.out = w.x;
    p.ldc = kD;
    p.rows_in = d.n_img_tokens;
    p.rows_out = kLP;
    p.row_off = 1;
    p.pe = h->pe_enc;
    p.pe_off = 1;
    const int M_img = B * d.n_img_tokens;
    if (!h->no_img_proj_pair && d.img_feat_dim % 64 == 0) {
      // CTA pairs: each CTA loads half of every weight k-block (image_proj_pair.cu)
      CUtensorMap ta, tw;
      const TmapType ty = f16 ? TmapType::kBF16 : TmapType::kF32;  // (2-byte elements: the map only moves bytes)
      RC_OK(get_tmap(h, img, ty, uint64_t(M_img), uint64_t(d.img_feat_dim), uint64_t(d.img_feat_dim), 128, &ta));
      RC_OK(get_tmap(h, f16 ? static_cast<const void*>(h->img_w_f16) : static_cast<const void*>(h->img_w_f32), ty, kD,
                     uint64_t(d.img_feat_dim), uint64_t(d.img_feat_dim), 128, &tw));
      p.M = M_img;
      p.N = kD;
      p.K = d.img_feat_dim;
      LAUNCH_OK(h, launch_image_proj_pair(f16 ? 1 : 0, ta, tw, p, s, h->l2_hints));
    } else if (f16) {  // `img` holds fp16 features: kind::f16 MMA against the fp16 copy of the weight
      p.f16 = true;
      RC_OK(gemm(h, kEpiBiasPeRemap, false, img, M_img, d.img_feat_dim, d.img_feat_dim, h->img_w_f16, kD, p, s));
    } else {
      RC_OK(gemm(h, kEpiBiasPeRemap, true, img, M_img, d.img_feat_dim, d.img_feat_dim, h->img_w_f32, kD, p, s));
    }
  }
  __nv_bfloat16* memory = nullptr;
  RC_OK(run_encoder(h, B, nullptr, S, &memory, s));
  if (opt_memory) {
    // seq-first [S, B_total, d]: this chunk fills columns b0..b0+B
    h->cur_tag = kTagMisc;
    LAUNCH_OK(h, launch_memory_export(memory, S, B, B_total, opt_memory + size_t(b0) * kD, s));
  }
  h->cur_tag = kTagAnswer;
  LAUNCH_OK(h, launch_answer_head(memory, B, h->ans_w0t, h->ans_b0, d.answer_hidden, h->ans_w1, h->ans_b1,
                                  d.num_classes, d.answer_pool_rows, answer, s));
  DecodeIO io;
  io.start_token = h->iqap_start_token;  // Config.SPECIAL_TOKEN_ID (IQAP:24,205)
  io.steps = T;
  io.logits = step_logits;
  io.logits_T = T;
  io.forced = forced;
  io.forced_ld = T;
  RC_OK(run_decoder(h, B, memory, nullptr, S, io, s));
  return publish(h, B, T, 1, programs, nullptr, T, nullptr, 0, nullptr, 0, s);  // drop the <START> column (IQAP:239)
}

}  // namespace

// =================================================================================================
// exported symbols
// =================================================================================================
extern "C" {

B200VQA_API const char* b200vqa_last_error(void) { return get_error(); }
B200VQA_API const char* b200vqa_version(void) { return "b200vqa 0.1 (sm_100a)"; }

B200VQA_API int b200vqa_create(const b200vqa_model_desc* desc, int device, b200vqa_handle** out) {
  B200VQA_REQUIRE(out != nullptr, "out handle pointer is NULL");
  *out = nullptr;
  RC_OK(validate_desc(desc));
  int num_sms = 0;
  RC_OK(require_sm100(device, &num_sms));
  B200VQA_CUDA_OK(cudaSetDevice(device));
  b200vqa_handle* h = new (std::nothrow) b200vqa_handle();
  if (!h) {
    set_error("out of host memory");
    return B200VQA_ERR_OUT_OF_MEMORY;
  }
  h->device = device;
  h->num_sms = num_sms;
  if (const char* g = getenv("B200VQA_NO_GRAPH")) h->use_graphs = !(g[0] && g[0] != '0');
  if (const char* g = getenv("B200VQA_NO_ABSORB")) h->absorb = !(g[0] && g[0] != '0');
  if (const char* g = getenv("B200VQA_BRANCH_STAGGER_US")) h->stagger_us = std::max(0, atoi(g));
  if (const char* g = getenv("B200VQA_BRANCH_STAGGER_MOD")) h->stagger_mod = std::max(2, atoi(g));
  if (const char* g = getenv("B200VQA_ABSORB_OV")) h->absorb_ov = g[0] && g[0] != '0';
  if (const char* g = getenv("B200VQA_ENC_ATTN_WHOLE_HEAD")) h->enc_attn_whole_head = g[0] && g[0] != '0';
  if (const char* g = getenv("B200VQA_NO_FUSED_FINAL_LN")) h->no_fused_final_ln = g[0] && g[0] != '0';
  if (const char* g = getenv("B200VQA_NO_WARP_SELF_ATTN")) h->no_warp_self_attn = g[0] && g[0] != '0';
  if (const char* g = getenv("B200VQA_NO_FUSED_HEAD")) h->no_fused_head = g[0] && g[0] != '0';
  if (const char* g = getenv("B200VQA_NO_LN_CLUSTER")) h->no_ln_cluster = g[0] && g[0] != '0';
  {  // process-wide (the launch helper is shared by every handle): re-evaluated on every create
    const char* g = getenv("B200VQA_NO_PDL");
    set_pdl_enabled(!(g && g[0] && g[0] != '0'));
  }
  if (const char* g = getenv("B200VQA_NVTX")) h->nvtx = g[0] && g[0] != '0';
  if (const char* g = getenv("B200VQA_DECODE")) h->decode_persist = g[0] == 'p';
  if (const char* g = getenv("B200VQA_PERSIST_STAGGER_US")) h->persist_stagger_us = std::max(0, atoi(g));
  if (const char* g = getenv("B200VQA_PERSIST_DBG_STOP")) h->persist_dbg_stop = atoi(g);
  if (const char* g = getenv("B200VQA_PERSIST_STAMPS")) {
    if (g[0] && g[0] != '0' && cudaMalloc(&h->persist_clk, 8 * 24 * sizeof(long long)) == cudaSuccess)
      cudaMemset(h->persist_clk, 0, 8 * 24 * sizeof(long long));
  }
  if (const char* g = getenv("B200VQA_NO_IMG_PROJ_PAIR")) h->no_img_proj_pair = g[0] && g[0] != '0';
  if (const char* g = getenv("B200VQA_NO_FUSED_ENC_FFN")) h->no_fused_enc_ffn = g[0] && g[0] != '0';
  if (const char* g = getenv("B200VQA_DBG_SKIP")) h->dbg_skip = atoi(g);
  if (const char* g = getenv("B200VQA_SMALL_BN")) h->small_bn = atoi(g);
  if (const char* g = getenv("B200VQA_NO_L2_HINTS")) h->l2_hints = !(g[0] && g[0] != '0');
  if (const char* g = getenv("B200VQA_DECODE_BRANCHES")) h->decode_branches = std::min(8, std::max(1, atoi(g)));
  h->d = *desc;
  h->enc_src.assign(desc->enc_layers, desc->enc_layers + desc->n_enc_layers);
  h->dec_src.assign(desc->dec_layers, desc->dec_layers + desc->n_dec_layers);
  h->d.enc_layers = nullptr;
  h->d.dec_layers = nullptr;
  Arena measure;
  layout_weights(h, measure);
  h->wbytes = measure.off + 256;
  cudaError_t e = cudaMalloc(&h->wbase, h->wbytes);
  if (e != cudaSuccess) {
    set_error("cudaMalloc(%zu) for packed weights failed: %s", h->wbytes, cudaGetErrorString(e));
    delete h;
    return B200VQA_ERR_OUT_OF_MEMORY;
  }
  Arena a;
  a.base = h->wbase;
  layout_weights(h, a);
  e = pack_weights(h, nullptr);
  if (e == cudaSuccess) e = cudaStreamSynchronize(nullptr);
  if (e != cudaSuccess) {
    set_error("weight packing failed: %s", cudaGetErrorString(e));
    cudaFree(h->wbase);
    delete h;
    return B200VQA_ERR_CUDA;
  }
  *out = h;
  return B200VQA_OK;
}

B200VQA_API int b200vqa_refresh_weights(b200vqa_handle* h, const b200vqa_model_desc* desc) {
  B200VQA_REQUIRE(h != nullptr, "handle is NULL");
  RC_OK(validate_desc(desc));
  const auto& o = h->d;
  B200VQA_REQUIRE(desc->kind == o.kind && desc->nhead == o.nhead && desc->n_enc_layers == o.n_enc_layers &&
                      desc->n_dec_layers == o.n_dec_layers && desc->dim_ff == o.dim_ff &&
                      desc->enc_vocab == o.enc_vocab && desc->dec_vocab == o.dec_vocab &&
                      desc->pe_enc_len == o.pe_enc_len && desc->pe_dec_len == o.pe_dec_len &&
                      desc->answer_hidden == o.answer_hidden && desc->num_classes == o.num_classes &&
                      desc->answer_pool_rows == o.answer_pool_rows &&
                      desc->img_feat_dim == o.img_feat_dim,
                  "refresh_weights: the new descriptor has different dimensions; create a new handle");
  RC_OK(set_device(h));
  h->d = *desc;
  h->enc_src.assign(desc->enc_layers, desc->enc_layers + desc->n_enc_layers);
  h->dec_src.assign(desc->dec_layers, desc->dec_layers + desc->n_dec_layers);
  h->d.enc_layers = nullptr;
  h->d.dec_layers = nullptr;
  B200VQA_CUDA_OK(cudaDeviceSynchronize());
  B200VQA_CUDA_OK(pack_weights(h, nullptr));
  B200VQA_CUDA_OK(cudaStreamSynchronize(nullptr));
  return B200VQA_OK;
}

B200VQA_API void b200vqa_destroy(b200vqa_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  for (auto& kv : h->graphs) cudaGraphExecDestroy(kv.second.exec);
  if (h->ws.base) cudaFree(h->ws.base);
  if (h->wbase) cudaFree(h->wbase);
  if (h->stage) cudaFree(h->stage);
  if (h->img_tok) cudaFree(h->img_tok);
  if (h->persist_clk) cudaFree(h->persist_clk);
  for (int i = 0; i < 2; ++i) {
    if (h->ev_in[i]) cudaEventDestroy(h->ev_in[i]);
    if (h->ev_free[i]) cudaEventDestroy(h->ev_free[i]);
  }
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  for (int i = 0; i < 2; ++i) {
    if (h->ev_pin[i]) cudaEventDestroy(h->ev_pin[i]);
    if (h->pin16[i]) cudaFreeHost(h->pin16[i]);
  }
  if (h->ev_tables) cudaEventDestroy(h->ev_tables);
  if (h->h_tables) cudaFreeHost(h->h_tables);
  if (h->cap_stream) cudaStreamDestroy(h->cap_stream);
  for (int i = 0; i < 7; ++i) {
    if (h->br_stream[i]) cudaStreamDestroy(h->br_stream[i]);
    if (h->br_done[i]) cudaEventDestroy(h->br_done[i]);
  }
  if (h->br_fork) cudaEventDestroy(h->br_fork);
  delete h;
}

B200VQA_API size_t b200vqa_workspace_bytes(const b200vqa_handle* h, int B) {
  if (!h || B <= 0) return 0;
  Arena measure;
  Workspace tmp;
  layout_workspace(h, tmp, measure, std::min(B, default_cap(h)), 32);
  return measure.off + 256;
}

B200VQA_API uint64_t b200vqa_launch_count(const b200vqa_handle* h) { return h ? h->launches : 0; }

B200VQA_API int b200vqa_set_host_upload(b200vqa_handle* h, int mode) {
  B200VQA_REQUIRE(h != nullptr, "handle is NULL");
  B200VQA_REQUIRE(mode == 0 || mode == 1, "upload mode %d: 0 = bytes as given, 1 = fp32 rounded to fp16 on the host", mode);
  h->host_upload_f16 = mode == 1;
  return B200VQA_OK;
}

B200VQA_API int b200vqa_set_start_token(b200vqa_handle* h, int start_token) {
  B200VQA_REQUIRE(h != nullptr, "handle is NULL");
  B200VQA_REQUIRE(start_token >= 0 && start_token < h->d.dec_vocab, "start token %d outside the decoder vocabulary (%d)",
                  start_token, h->d.dec_vocab);
  h->iqap_start_token = start_token;
  return B200VQA_OK;
}

// ------------------------------------------------------------------------------------------------ profiler
B200VQA_API int b200vqa_profile_num_tags(void) { return kNumTags; }
B200VQA_API const char* b200vqa_profile_tag_name(int tag) { return (tag >= 0 && tag < kNumTags) ? kTagNames[tag] : ""; }

B200VQA_API int b200vqa_profile_begin(b200vqa_handle* h) {
  B200VQA_REQUIRE(h != nullptr, "handle is NULL");
  for (auto& r : h->prof) {
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  h->prof.clear();
  h->profiling = true;
  return B200VQA_OK;
}

B200VQA_API int b200vqa_profile_delay(b200vqa_handle* h, double ms, void* stream) {
  B200VQA_REQUIRE(h != nullptr && ms >= 0 && ms <= 200, "bad arguments");
  RC_OK(set_device(h));
  int khz = 0;
  B200VQA_CUDA_OK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, h->device));
  B200VQA_CUDA_OK(launch_delay((long long)(ms * khz), static_cast<cudaStream_t>(stream)));
  return B200VQA_OK;
}

B200VQA_API int b200vqa_profile_end(b200vqa_handle* h, float* ms_per_tag, int32_t* launches_per_tag) {
  B200VQA_REQUIRE(h != nullptr && ms_per_tag && launches_per_tag, "NULL argument");
  RC_OK(set_device(h));
  h->profiling = false;
  B200VQA_CUDA_OK(cudaDeviceSynchronize());
  for (int t = 0; t < kNumTags; ++t) {
    ms_per_tag[t] = 0.f;
    launches_per_tag[t] = 0;
  }
  for (auto& r : h->prof) {
    float ms = 0.f;
    B200VQA_CUDA_OK(cudaEventElapsedTime(&ms, r.a, r.b));
    ms_per_tag[r.tag] += ms;
    launches_per_tag[r.tag] += 1;
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  h->prof.clear();
  return B200VQA_OK;
}

// ------------------------------------------------------------------------------------------------ IQAP
static int iqap_forward_impl(b200vqa_handle* h, const void* image_features, bool f16, const int64_t* questions, int B,
                             int program_len, float* answer, int64_t* programs, float* opt_step_logits,
                             const int64_t* opt_forced_tokens, float* opt_memory, void* stream) {
  B200VQA_REQUIRE(h != nullptr, "handle is NULL");
  B200VQA_REQUIRE(h->d.kind == B200VQA_MODEL_IQAP, "handle was not created for the IQAP model");
  B200VQA_REQUIRE(B >= 0, "negative batch");
  if (B == 0) return B200VQA_OK;
  B200VQA_REQUIRE(image_features && questions && answer && programs, "a required buffer is NULL");
  RC_OK(check_decode_len(h, program_len));
  RC_OK(set_device(h));
  RC_OK(ensure_workspace(h, B, program_len));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const auto& d = h->d;
  const int cap = h->ws.cap;
  for (int b0 = 0; b0 < B; b0 += cap) {
    const int nb = std::min(cap, B - b0);
    const size_t off = size_t(b0) * d.n_img_tokens * d.img_feat_dim * (f16 ? 2 : 4);
    RC_OK(iqap_chunk(h, reinterpret_cast<const float*>(static_cast<const uint8_t*>(image_features) + off),
                     questions + size_t(b0) * d.max_q_len, nb, program_len, answer + size_t(b0) * d.num_classes,
                     programs + size_t(b0) * program_len,
                     opt_step_logits ? opt_step_logits + size_t(b0) * program_len * d.dec_vocab : nullptr,
                     opt_forced_tokens ? opt_forced_tokens + size_t(b0) * program_len : nullptr, opt_memory, B, b0,
                     s, nullptr, 0, f16));
  }
  return B200VQA_OK;
}

B200VQA_API int b200vqa_iqap_forward(b200vqa_handle* h, const float* image_features, const int64_t* questions, int B,
                         int program_len, float* answer, int64_t* programs, float* opt_step_logits,
                         const int64_t* opt_forced_tokens, float* opt_memory, void* stream) {
  return iqap_forward_impl(h, image_features, false, questions, B, program_len, answer, programs, opt_step_logits,
                           opt_forced_tokens, opt_memory, stream);
}

B200VQA_API int b200vqa_iqap_forward_f16(b200vqa_handle* h, const void* image_features_f16, const int64_t* questions,
                                         int B, int program_len, float* answer, int64_t* programs,
                                         float* opt_step_logits, const int64_t* opt_forced_tokens, float* opt_memory,
                                         void* stream) {
  return iqap_forward_impl(h, image_features_f16, true, questions, B, program_len, answer, programs, opt_step_logits,
                           opt_forced_tokens, opt_memory, stream);
}

B200VQA_API int b200vqa_iqap_forward_indexed(b200vqa_handle* h, const float* image_features, int n_img,
                                             const int32_t* image_idx, const int64_t* questions, int B, int program_len,
                                             float* answer, int64_t* programs, void* stream) {
  B200VQA_REQUIRE(h != nullptr, "handle is NULL");
  B200VQA_REQUIRE(h->d.kind == B200VQA_MODEL_IQAP, "handle was not created for the IQAP model");
  B200VQA_REQUIRE(B >= 0 && n_img >= 0, "negative batch");
  if (B == 0) return B200VQA_OK;
  B200VQA_REQUIRE(n_img >= 1, "questions without images");
  B200VQA_REQUIRE(image_features && image_idx && questions && answer && programs, "a required buffer is NULL");
  RC_OK(check_decode_len(h, program_len));
  RC_OK(set_device(h));
  RC_OK(ensure_workspace(h, B, program_len));
  RC_OK(ensure_img_tok(h, size_t(n_img)));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const auto& d = h->d;
  RC_OK(project_unique_images(h, image_features, 0, n_img, s));
  const int cap = h->ws.cap;
  for (int b0 = 0; b0 < B; b0 += cap) {
    const int nb = std::min(cap, B - b0);
    RC_OK(iqap_chunk(h, nullptr, questions + size_t(b0) * d.max_q_len, nb, program_len,
                     answer + size_t(b0) * d.num_classes, programs + size_t(b0) * program_len, nullptr, nullptr, nullptr,
                     B, b0, s, image_idx + b0, n_img));
  }
  return B200VQA_OK;
}

B200VQA_API int b200vqa_iqap_tally(b200vqa_handle* h, const float* answer_logits, const int64_t* programs,
                                   const int64_t* gt_answers, const int64_t* gt_programs, int B, int program_len,
                                   uint64_t* counts, int32_t* opt_pred_answers, void* stream) {
  B200VQA_REQUIRE(h != nullptr, "handle is NULL");
  B200VQA_REQUIRE(h->d.kind == B200VQA_MODEL_IQAP, "handle was not created for the IQAP model");
  B200VQA_REQUIRE(B >= 0 && program_len >= 1, "shape out of range (B %d, program length %d)", B, program_len);
  if (B == 0) return B200VQA_OK;
  B200VQA_REQUIRE(answer_logits && programs && gt_answers && gt_programs && counts, "a required buffer is NULL");
  RC_OK(set_device(h));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  TallyParams p;
  p.B = B;
  p.classes = h->d.num_classes;
  p.T = program_len;
  p.answer_logits = answer_logits;
  p.programs = programs;
  p.gt_answers = gt_answers;
  p.gt_programs = gt_programs;
  p.counts = reinterpret_cast<unsigned long long*>(counts);
  p.pred_answers = opt_pred_answers;
  h->cur_tag = kTagMisc;
  LAUNCH_OK(h, launch_tally(p, s));
  return B200VQA_OK;
}

B200VQA_API int b200vqa_iqap_decode(b200vqa_handle* h, const float* memory, int S, int B, int program_len, int64_t* programs,
                        float* opt_step_logits, const int64_t* opt_forced_tokens, void* stream) {
  B200VQA_REQUIRE(h != nullptr, "handle is NULL");
  B200VQA_REQUIRE(h->d.kind == B200VQA_MODEL_IQAP, "handle was not created for the IQAP model");
  B200VQA_REQUIRE(B >= 0 && S >= 1 && S <= kLP, "memory shape out of range (S %d, B %d)", S, B);
  if (B == 0) return B200VQA_OK;
  B200VQA_REQUIRE(memory && programs, "a required buffer is NULL");
  RC_OK(check_decode_len(h, program_len));
  RC_OK(set_device(h));
  RC_OK(ensure_workspace(h, B, program_len));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int cap = h->ws.cap;
  const int V = h->d.dec_vocab;
  for (int b0 = 0; b0 < B; b0 += cap) {
    const int nb = std::min(cap, B - b0);
    // seq-first memory interleaves questions: a chunk is the column range [b0, b0+nb) of every row
    h->cur_tag = kTagMisc;
    LAUNCH_OK(h, launch_memory_import(memory + size_t(b0) * kD, S, nb, B, h->ws.mem, s));
    DecodeIO io;
    io.start_token = h->iqap_start_token;
    io.steps = program_len;
    io.logits = opt_step_logits ? opt_step_logits + size_t(b0) * program_len * V : nullptr;
    io.logits_T = program_len;
    io.forced = opt_forced_tokens ? opt_forced_tokens + size_t(b0) * program_len : nullptr;
    io.forced_ld = program_len;
    RC_OK(run_decoder(h, nb, h->ws.mem, nullptr, S, io, s));
    RC_OK(publish(h, nb, program_len, 1, programs + size_t(b0) * program_len, nullptr, program_len, nullptr, 0, nullptr,
                  0, s));
  }
  return B200VQA_OK;
}

static int iqap_forward_host_impl(b200vqa_handle* h, const void* h_img, bool f16, const int64_t* h_q, int B,
                                  int program_len, float* h_answer, int64_t* h_programs, int chunk, void* stream,
                                  bool sync) {
  B200VQA_REQUIRE(h != nullptr, "handle is NULL");
  B200VQA_REQUIRE(h->d.kind == B200VQA_MODEL_IQAP, "handle was not created for the IQAP model");
  B200VQA_REQUIRE(B >= 0, "negative batch");
  if (B == 0) return B200VQA_OK;
  B200VQA_REQUIRE(h_img && h_q && h_answer && h_programs, "a required buffer is NULL");
  // fp32 features rounded to fp16 on the host (b200vqa_set_host_upload): half the bytes on the wire, and the device path
  // of a caller-provided fp16 feature store from there on
  const bool convert = !f16 && h->host_upload_f16;
  if (convert) f16 = true;
  const size_t esz = f16 ? 2 : 4;  // bytes per feature element on the device
  RC_OK(check_decode_len(h, program_len));
  RC_OK(set_device(h));
  const auto& d = h->d;
  if (chunk <= 0) chunk = 512;  // decode is latency-bound: few large chunks beat many small ones
  chunk = std::min({chunk, B, default_cap(h)});
  RC_OK(ensure_workspace(h, chunk, program_len));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!h->copy_stream) {
    B200VQA_CUDA_OK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      B200VQA_CUDA_OK(cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming));
      B200VQA_CUDA_OK(cudaEventCreateWithFlags(&h->ev_free[i], cudaEventDisableTiming));
    }
  }
  // staging: 2 x (features + questions) for the double-buffered upload, + results for the whole batch
  const size_t img_b = (size_t(chunk) * d.n_img_tokens * d.img_feat_dim * esz + 255) & ~size_t(255);
  const size_t q_b = (size_t(chunk) * d.max_q_len * sizeof(int64_t) + 255) & ~size_t(255);
  const size_t ans_b = (size_t(B) * d.num_classes * sizeof(float) + 255) & ~size_t(255);
  const size_t prog_b = (size_t(B) * program_len * sizeof(int64_t) + 255) & ~size_t(255);
  const size_t need = 2 * (img_b + q_b) + ans_b + prog_b;
  if (h->stage_bytes < need) {
    if (h->stage) {
      B200VQA_CUDA_OK(cudaDeviceSynchronize());
      B200VQA_CUDA_OK(cudaFree(h->stage));
      h->stage = nullptr;
      h->stage_bytes = 0;
      h->tmaps.clear();
    }
    B200VQA_CUDA_OK(cudaMalloc(&h->stage, need));
    h->stage_bytes = need;
  }
  uint8_t* p = h->stage;
  float* d_img[2];
  int64_t* d_q[2];
  for (int i = 0; i < 2; ++i) {
    d_img[i] = reinterpret_cast<float*>(p); p += img_b;
    d_q[i] = reinterpret_cast<int64_t*>(p); p += q_b;
  }
  float* d_ans = reinterpret_cast<float*>(p); p += ans_b;
  int64_t* d_prog = reinterpret_cast<int64_t*>(p);

  const size_t per_q = size_t(d.n_img_tokens) * d.img_feat_dim;  // feature elements per question
  if (convert) {
    const size_t pin_b = size_t(chunk) * per_q * 2;
    if (h->pin16_bytes < pin_b) {
      for (int i = 0; i < 2; ++i) {
        if (h->pin_pending[i]) B200VQA_CUDA_OK(cudaEventSynchronize(h->ev_pin[i]));
        h->pin_pending[i] = false;
        if (h->pin16[i]) B200VQA_CUDA_OK(cudaFreeHost(h->pin16[i]));
        h->pin16[i] = nullptr;
      }
      h->pin16_bytes = 0;
      for (int i = 0; i < 2; ++i) B200VQA_CUDA_OK(cudaMallocHost(&h->pin16[i], pin_b));
      h->pin16_bytes = pin_b;
    }
    for (int i = 0; i < 2; ++i)
      if (!h->ev_pin[i]) B200VQA_CUDA_OK(cudaEventCreateWithFlags(&h->ev_pin[i], cudaEventDisableTiming));
  }

  // the copy stream must not overwrite staging that earlier, still unsynchronised work of `s` may read (a previous
  // *_host_async call on this handle): everything enqueued on `s` so far precedes the first upload
  B200VQA_CUDA_OK(cudaEventRecord(h->ev_free[0], s));
  B200VQA_CUDA_OK(cudaStreamWaitEvent(h->copy_stream, h->ev_free[0], 0));
  int it = 0;
  for (int b0 = 0; b0 < B; b0 += chunk, ++it) {
    const int nb = std::min(chunk, B - b0);
    const int slot = it & 1;
    const void* src = static_cast<const uint8_t*>(h_img) + size_t(b0) * per_q * (convert ? 4 : esz);
    if (convert) {
      // the host rounds chunk `it` while chunk it - 1 is on the wire; the staging is free once its own upload (two
      // chunks ago, or a previous call's) has completed
      const bool trace = getenv("B200VQA_HOST_TRACE") != nullptr;
      const auto t0 = std::chrono::steady_clock::now();
      if (h->pin_pending[slot]) {
        B200VQA_CUDA_OK(cudaEventSynchronize(h->ev_pin[slot]));
        h->pin_pending[slot] = false;
      }
      const auto t1 = std::chrono::steady_clock::now();
      host_f32_to_f16(static_cast<const float*>(src), h->pin16[slot], size_t(nb) * per_q, 0);
      if (trace) {
        const auto t2 = std::chrono::steady_clock::now();
        static const auto t_first = t0;
        fprintf(stderr,
                "b200vqa host upload: t = %.2f ms, chunk %d (%d questions): waited %.2f ms for the staging, converted in %.2f ms\n",
                std::chrono::duration<double, std::milli>(t0 - t_first).count(), it, nb,
                std::chrono::duration<double, std::milli>(t1 - t0).count(),
                std::chrono::duration<double, std::milli>(t2 - t1).count());
      }
      src = h->pin16[slot];
    }
    if (it >= 2) B200VQA_CUDA_OK(cudaStreamWaitEvent(h->copy_stream, h->ev_free[slot], 0));
    B200VQA_CUDA_OK(cudaMemcpyAsync(d_img[slot], src, size_t(nb) * per_q * esz, cudaMemcpyHostToDevice, h->copy_stream));
    if (convert) {
      B200VQA_CUDA_OK(cudaEventRecord(h->ev_pin[slot], h->copy_stream));
      h->pin_pending[slot] = true;
    }
    B200VQA_CUDA_OK(cudaMemcpyAsync(d_q[slot], h_q + size_t(b0) * d.max_q_len,
                                    size_t(nb) * d.max_q_len * sizeof(int64_t), cudaMemcpyHostToDevice,
                                    h->copy_stream));
    B200VQA_CUDA_OK(cudaEventRecord(h->ev_in[slot], h->copy_stream));
    B200VQA_CUDA_OK(cudaStreamWaitEvent(s, h->ev_in[slot], 0));
    RC_OK(iqap_chunk(h, d_img[slot], d_q[slot], nb, program_len, d_ans + size_t(b0) * d.num_classes,
                     d_prog + size_t(b0) * program_len, nullptr, nullptr, nullptr, B, b0, s, nullptr, 0, f16));
    B200VQA_CUDA_OK(cudaEventRecord(h->ev_free[slot], s));
  }
  B200VQA_CUDA_OK(cudaMemcpyAsync(h_answer, d_ans, size_t(B) * d.num_classes * sizeof(float), cudaMemcpyDeviceToHost, s));
  B200VQA_CUDA_OK(cudaMemcpyAsync(h_programs, d_prog, size_t(B) * program_len * sizeof(int64_t),
                                  cudaMemcpyDeviceToHost, s));
  if (sync) B200VQA_CUDA_OK(cudaStreamSynchronize(s));
  return B200VQA_OK;
}

// Host buffers, several questions per image: the unique images are uploaded (double-buffered) and projected once, then
// the questions run in chunks whose image rows are gathered on the device.
B200VQA_API int b200vqa_iqap_forward_host_indexed(b200vqa_handle* h, const float* h_img, int n_img,
                                                  const int32_t* h_image_idx, const int64_t* h_q, int B,
                                                  int program_len, float* h_answer, int64_t* h_programs, int chunk,
                                                  void* stream) {
  B200VQA_REQUIRE(h != nullptr, "handle is NULL");
  B200VQA_REQUIRE(h->d.kind == B200VQA_MODEL_IQAP, "handle was not created for the IQAP model");
  B200VQA_REQUIRE(B >= 0 && n_img >= 0, "negative batch");
  if (B == 0) return B200VQA_OK;
  B200VQA_REQUIRE(n_img >= 1, "questions without images");
  B200VQA_REQUIRE(h_img && h_image_idx && h_q && h_answer && h_programs, "a required buffer is NULL");
  RC_OK(check_decode_len(h, program_len));
  RC_OK(set_device(h));
  const auto& d = h->d;
  if (chunk <= 0) chunk = 512;
  chunk = std::min({chunk, B, default_cap(h)});
  const int ichunk = std::min(n_img, 256);  // images per upload: 256 x 803 KB = 205 MB per staging buffer
  RC_OK(ensure_workspace(h, chunk, program_len));
  RC_OK(ensure_img_tok(h, size_t(n_img)));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!h->copy_stream) {
    B200VQA_CUDA_OK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      B200VQA_CUDA_OK(cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming));
      B200VQA_CUDA_OK(cudaEventCreateWithFlags(&h->ev_free[i], cudaEventDisableTiming));
    }
  }
  auto up = [](size_t n) { return (n + 255) & ~size_t(255); };
  const size_t img_b = up(size_t(ichunk) * d.n_img_tokens * d.img_feat_dim * sizeof(float));
  const size_t q_b = up(size_t(B) * d.max_q_len * sizeof(int64_t));
  const size_t idx_b = up(size_t(B) * sizeof(int32_t));
  const size_t ans_b = up(size_t(B) * d.num_classes * sizeof(float));
  const size_t prog_b = up(size_t(B) * program_len * sizeof(int64_t));
  const size_t need = 2 * img_b + q_b + idx_b + ans_b + prog_b;
  if (h->stage_bytes < need) {
    if (h->stage) {
      B200VQA_CUDA_OK(cudaDeviceSynchronize());
      B200VQA_CUDA_OK(cudaFree(h->stage));
      h->stage = nullptr;
      h->stage_bytes = 0;
      h->tmaps.clear();
    }
    B200VQA_CUDA_OK(cudaMalloc(&h->stage, need));
    h->stage_bytes = need;
  }
  uint8_t* p = h->stage;
  float* d_img[2];
  for (int i = 0; i < 2; ++i) { d_img[i] = reinterpret_cast<float*>(p); p += img_b; }
  int64_t* d_q = reinterpret_cast<int64_t*>(p); p += q_b;
  int32_t* d_idx = reinterpret_cast<int32_t*>(p); p += idx_b;
  float* d_ans = reinterpret_cast<float*>(p); p += ans_b;
  int64_t* d_prog = reinterpret_cast<int64_t*>(p);

  // the copy stream must not overwrite staging that earlier work of `s` may still read
  B200VQA_CUDA_OK(cudaEventRecord(h->ev_free[0], s));
  B200VQA_CUDA_OK(cudaStreamWaitEvent(h->copy_stream, h->ev_free[0], 0));
  B200VQA_CUDA_OK(cudaMemcpyAsync(d_q, h_q, size_t(B) * d.max_q_len * sizeof(int64_t), cudaMemcpyHostToDevice,
                                  h->copy_stream));
  B200VQA_CUDA_OK(cudaMemcpyAsync(d_idx, h_image_idx, size_t(B) * sizeof(int32_t), cudaMemcpyHostToDevice,
                                  h->copy_stream));
  int it = 0;
  const size_t per_img = size_t(d.n_img_tokens) * d.img_feat_dim;
  for (int i0 = 0; i0 < n_img; i0 += ichunk, ++it) {
    const int ni = std::min(ichunk, n_img - i0);
    const int slot = it & 1;
    if (it >= 2) B200VQA_CUDA_OK(cudaStreamWaitEvent(h->copy_stream, h->ev_free[slot], 0));
    B200VQA_CUDA_OK(cudaMemcpyAsync(d_img[slot], h_img + size_t(i0) * per_img, size_t(ni) * per_img * sizeof(float),
                                    cudaMemcpyHostToDevice, h->copy_stream));
    B200VQA_CUDA_OK(cudaEventRecord(h->ev_in[slot], h->copy_stream));
    B200VQA_CUDA_OK(cudaStreamWaitEvent(s, h->ev_in[slot], 0));
    RC_OK(project_unique_images(h, d_img[slot], size_t(i0), ni, s));
    B200VQA_CUDA_OK(cudaEventRecord(h->ev_free[slot], s));
  }
  for (int b0 = 0; b0 < B; b0 += chunk) {
    const int nb = std::min(chunk, B - b0);
    RC_OK(iqap_chunk(h, nullptr, d_q + size_t(b0) * d.max_q_len, nb, program_len, d_ans + size_t(b0) * d.num_classes,
                     d_prog + size_t(b0) * program_len, nullptr, nullptr, nullptr, B, b0, s, d_idx + b0, n_img));
  }
  B200VQA_CUDA_OK(cudaMemcpyAsync(h_answer, d_ans, size_t(B) * d.num_classes * sizeof(float), cudaMemcpyDeviceToHost, s));
  B200VQA_CUDA_OK(cudaMemcpyAsync(h_programs, d_prog, size_t(B) * program_len * sizeof(int64_t),
                                  cudaMemcpyDeviceToHost, s));
  B200VQA_CUDA_OK(cudaStreamSynchronize(s));
  return B200VQA_OK;
}

B200VQA_API int b200vqa_iqap_forward_host(b200vqa_handle* h, const float* h_img, const int64_t* h_q, int B, int program_len,
                              float* h_answer, int64_t* h_programs, int chunk, void* stream) {
  return iqap_forward_host_impl(h, h_img, false, h_q, B, program_len, h_answer, h_programs, chunk, stream, true);
}

B200VQA_API int b200vqa_iqap_forward_host_f16(b200vqa_handle* h, const void* h_img_f16, const int64_t* h_q, int B,
                                              int program_len, float* h_answer, int64_t* h_programs, int chunk,
                                              void* stream) {
  return iqap_forward_host_impl(h, h_img_f16, true, h_q, B, program_len, h_answer, h_programs, chunk, stream, true);
}

B200VQA_API int b200vqa_iqap_forward_host_async(b200vqa_handle* h, const float* h_img, const int64_t* h_q, int B,
                                                int program_len, float* h_answer, int64_t* h_programs, int chunk,
                                                void* stream) {
  return iqap_forward_host_impl(h, h_img, false, h_q, B, program_len, h_answer, h_programs, chunk, stream, false);
}

// ------------------------------------------------------------------------------------------------ FA
// src_bf16: the features were rounded to bf16 on the host (the host-buffer entry's bf16 upload mode) - the rounding the
// transpose does anyway, so both forms give identical image tokens
static int fa_project_images_impl(b200vqa_handle* h, const void* image_features, bool src_bf16, int B,
                                  void* img_tokens_bf16, void* stream) {
  B200VQA_REQUIRE(h != nullptr, "handle is NULL");
  B200VQA_REQUIRE(h->d.kind == B200VQA_MODEL_FA, "handle was not created for the FA model");
  B200VQA_REQUIRE(B >= 0, "negative batch");
  if (B == 0) return B200VQA_OK;
  B200VQA_REQUIRE(image_features && img_tokens_bf16, "a required buffer is NULL");
  RC_OK(set_device(h));
  RC_OK(ensure_workspace(h, B, 20));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const auto& d = h->d;
  const int cap = h->ws.cap;
  __nv_bfloat16* out = static_cast<__nv_bfloat16*>(img_tokens_bf16);
  for (int b0 = 0; b0 < B; b0 += cap) {
    const int nb = std::min(cap, B - b0);
    // channel-major fp32 [nb,1024,196] -> token-major bf16 [nb*196,1024] (the reference's view+permute, FA:47,130)
    h->cur_tag = kTagMisc;
    const size_t off = size_t(b0) * d.img_feat_dim * d.n_img_tokens;
    if (src_bf16)
      LAUNCH_OK(h, launch_transpose_bf16(static_cast<const __nv_bfloat16*>(image_features) + off, h->ws.img_t, nb,
                                         d.img_feat_dim, d.n_img_tokens, s));
    else
      LAUNCH_OK(h, launch_transpose_cast(static_cast<const float*>(image_features) + off, h->ws.img_t, nb, d.img_feat_dim,
                                         d.n_img_tokens, s));
    h->cur_tag = kTagImgProj;
    GemmParams p;
    p.bias = h->img_b;
    p.out = out + size_t(b0) * d.n_img_tokens * kD;
    p.ldc = kD;
    p.rows_in = d.n_img_tokens;
    p.rows_out = d.n_img_tokens;
    p.row_off = 0;
    p.pe = h->pe_enc;
    p.pe_off = 0;
    RC_OK(gemm(h, kEpiBiasPeRemap, false, h->ws.img_t, nb * d.n_img_tokens, d.img_feat_dim, d.img_feat_dim,
               h->img_w_bf16, kD, p, s));
  }
  return B200VQA_OK;
}

B200VQA_API int b200vqa_fa_project_images(b200vqa_handle* h, const float* image_features, int B, void* img_tokens_bf16,
                              void* stream) {
  return fa_project_images_impl(h, image_features, false, B, img_tokens_bf16, stream);
}

static int fa_one_step(b200vqa_handle* h, const __nv_bfloat16* img_tokens, FaBuildSrcParams bp, int B, DecodeIO io,
                       cudaStream_t s) {
  Workspace& w = h->ws;
  const auto& d = h->d;
  bp.B = B;
  bp.img_tokens = img_tokens;
  bp.emb = h->enc_emb;
  bp.vocab = d.enc_vocab;
  bp.pe = h->pe_enc;
  bp.pe_len = d.pe_enc_len;
  bp.n_img = d.n_img_tokens;
  bp.x = w.x;
  bp.lens = w.lens;
  h->cur_tag = kTagEmbed;
  LAUNCH_OK(h, launch_fa_build_src(bp, s));
  __nv_bfloat16* memory = nullptr;
  RC_OK(run_encoder(h, B, w.lens, 0, &memory, s));
  return run_decoder(h, B, memory, w.lens, 0, io, s);
}

B200VQA_API int b200vqa_fa_step(b200vqa_handle* h, const void* img_tokens_bf16, const int64_t* src, const int32_t* src_len,
                    int src_ld, int B, int start_token, int max_len, int64_t* out_tokens, float* opt_logits,
                    const int64_t* opt_forced, void* stream) {
  B200VQA_REQUIRE(h != nullptr, "handle is NULL");
  B200VQA_REQUIRE(h->d.kind == B200VQA_MODEL_FA, "handle was not created for the FA model");
  B200VQA_REQUIRE(B >= 0, "negative batch");
  if (B == 0) return B200VQA_OK;
  B200VQA_REQUIRE(img_tokens_bf16 && src && out_tokens, "a required buffer is NULL");
  B200VQA_REQUIRE(src_ld >= 0 && src_ld <= 60 && h->d.n_img_tokens + src_ld <= h->d.pe_enc_len,
                  "src length %d exceeds the encoder positional table (%d rows, %d image tokens) or 60 tokens",
                  src_ld, h->d.pe_enc_len, h->d.n_img_tokens);
  B200VQA_REQUIRE(max_len >= 2, "max_len must be at least 2");
  RC_OK(check_decode_len(h, max_len - 1));
  RC_OK(set_device(h));
  RC_OK(ensure_workspace(h, B, max_len - 1));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const auto& d = h->d;
  const int cap = h->ws.cap;
  const int T = max_len - 1;
  for (int b0 = 0; b0 < B; b0 += cap) {
    const int nb = std::min(cap, B - b0);
    FaBuildSrcParams bp;
    bp.src_direct = src + size_t(b0) * src_ld;
    bp.src_len_in = src_len ? src_len + b0 : nullptr;
    bp.src_ld = src_ld;
    DecodeIO io;
    io.start_token = start_token;
    io.steps = T;
    io.logits = opt_logits ? opt_logits + size_t(b0) * T * d.dec_vocab : nullptr;
    io.logits_T = T;
    io.forced = opt_forced ? opt_forced + size_t(b0) * T : nullptr;
    io.forced_ld = T;
    RC_OK(fa_one_step(h, static_cast<const __nv_bfloat16*>(img_tokens_bf16) + size_t(b0) * d.n_img_tokens * kD, bp, nb,
                      io, s));
    RC_OK(publish(h, nb, max_len, 0, out_tokens + size_t(b0) * max_len, nullptr, max_len, nullptr, 0, nullptr, 0, s));
  }
  return B200VQA_OK;
}

B200VQA_API int b200vqa_fa_forward(b200vqa_handle* h, const void* img_tokens_bf16, const int64_t* src,
                                   const int32_t* src_len, int src_ld, int B, const int64_t* tgt, int T, float* logits,
                                   void* stream) {
  B200VQA_REQUIRE(h != nullptr, "handle is NULL");
  B200VQA_REQUIRE(h->d.kind == B200VQA_MODEL_FA, "handle was not created for the FA model");
  B200VQA_REQUIRE(B >= 0, "negative batch");
  if (B == 0) return B200VQA_OK;
  B200VQA_REQUIRE(img_tokens_bf16 && src && tgt && logits, "a required buffer is NULL");
  B200VQA_REQUIRE(src_ld >= 0 && src_ld <= 60 && h->d.n_img_tokens + src_ld <= h->d.pe_enc_len,
                  "src length %d exceeds the encoder positional table (%d rows, %d image tokens) or 60 tokens",
                  src_ld, h->d.pe_enc_len, h->d.n_img_tokens);
  RC_OK(check_decode_len(h, T));
  RC_OK(set_device(h));
  RC_OK(ensure_workspace(h, B, T));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const auto& d = h->d;
  const int cap = h->ws.cap;
  for (int b0 = 0; b0 < B; b0 += cap) {
    const int nb = std::min(cap, B - b0);
    FaBuildSrcParams bp;
    bp.src_direct = src + size_t(b0) * src_ld;
    bp.src_len_in = src_len ? src_len + b0 : nullptr;
    bp.src_ld = src_ld;
    DecodeIO io;
    io.start_tokens = tgt + size_t(b0) * T;
    io.start_ld = T;
    io.steps = T;
    io.logits = logits + size_t(b0) * T * d.dec_vocab;
    io.logits_T = T;
    io.forced = tgt + size_t(b0) * T + 1;  // position t+1 is fed tgt[b, t+1]; never read at t = T-1
    io.forced_ld = T;
    RC_OK(fa_one_step(h, static_cast<const __nv_bfloat16*>(img_tokens_bf16) + size_t(b0) * d.n_img_tokens * kD, bp, nb,
                      io, s));
  }
  return B200VQA_OK;
}

static int fa_run_chain_impl(b200vqa_handle* h, const void* img_tokens_bf16, const int32_t* image_idx, int n_images,
                             const int32_t* func, const int32_t* deps, const int32_t* n_steps, int B, int S,
                             int start_token, int max_len, int32_t* cache, const int32_t* h_active, float* opt_logits,
                             const int64_t* opt_forced, void* stream) {
  B200VQA_REQUIRE(h != nullptr, "handle is NULL");
  B200VQA_REQUIRE(h->d.kind == B200VQA_MODEL_FA, "handle was not created for the FA model");
  B200VQA_REQUIRE(B >= 0 && S >= 0, "negative batch or step count");
  if (B == 0 || S == 0) return B200VQA_OK;
  B200VQA_REQUIRE(img_tokens_bf16 && func && deps && n_steps && cache, "a required buffer is NULL");
  B200VQA_REQUIRE(max_len >= 2 && 1 + 2 * max_len <= 60, "max_len %d out of range (2..29)", max_len);
  B200VQA_REQUIRE(h->d.n_img_tokens + 1 + 2 * max_len <= std::min(h->d.pe_enc_len, kLP),
                  "1 + 2*max_len source tokens exceed the encoder positional table (%d rows)", h->d.pe_enc_len);
  RC_OK(check_decode_len(h, max_len - 1));
  RC_OK(set_device(h));
  RC_OK(ensure_workspace(h, B, max_len - 1));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const auto& d = h->d;
  const int cap = h->ws.cap;
  const int T = max_len - 1;
  // questions are independent, steps of one question are sequential: chunk over questions, loop steps inside
  for (int b0 = 0; b0 < B; b0 += cap) {
    const int nb_full = std::min(cap, B - b0);
    for (int i = 0; i < S; ++i) {
      int nb = nb_full;
      if (h_active) nb = std::max(0, std::min(nb_full, h_active[i] - b0));
      if (nb == 0) continue;
      FaBuildSrcParams bp;
      bp.step = i;
      bp.S = S;
      bp.T = max_len;
      bp.func = func + size_t(b0) * S;
      bp.deps = deps + size_t(b0) * S * 2;
      bp.n_steps = n_steps + b0;
      bp.cache = cache + size_t(b0) * S * max_len;
      bp.image_idx = image_idx ? image_idx + b0 : nullptr;
      bp.n_images = n_images;
      DecodeIO io;
      io.start_token = start_token;
      io.steps = T;
      io.logits = opt_logits ? opt_logits + (size_t(b0) * S + i) * T * d.dec_vocab : nullptr;
      io.logits_T = S * T;  // row (b, i, t) = b*S*T + i*T + t
      io.forced = opt_forced ? opt_forced + (size_t(b0) * S + i) * T : nullptr;
      io.forced_ld = S * T;
      RC_OK(fa_one_step(h, static_cast<const __nv_bfloat16*>(img_tokens_bf16) +
                               (image_idx ? size_t(0) : size_t(b0) * d.n_img_tokens * kD), bp, nb, io, s));
      // step i's 20 tokens (start token included, FA:120-121) into the HBM cache; with teacher forcing the
      // cache receives the forced tokens so that later steps consume exactly what the caller dictated
      RC_OK(publish(h, nb, max_len, 0, nullptr, cache + (size_t(b0) * S + i) * max_len, (long long)S * max_len, io.forced,
                    io.forced_ld, n_steps + b0, i, s));
    }
  }
  return B200VQA_OK;
}

B200VQA_API int b200vqa_fa_run_chain(b200vqa_handle* h, const void* img_tokens_bf16, const int32_t* func, const int32_t* deps,
                         const int32_t* n_steps, int B, int S, int start_token, int max_len, int32_t* cache,
                         const int32_t* h_active, float* opt_logits, const int64_t* opt_forced, void* stream) {
  return fa_run_chain_impl(h, img_tokens_bf16, nullptr, 0, func, deps, n_steps, B, S, start_token, max_len, cache,
                           h_active, opt_logits, opt_forced, stream);
}

B200VQA_API int b200vqa_fa_run_chain_indexed(b200vqa_handle* h, const void* img_tokens_bf16, int n_images,
                                             const int32_t* image_idx, const int32_t* func, const int32_t* deps,
                                             const int32_t* n_steps, int B, int S, int start_token, int max_len,
                                             int32_t* cache, const int32_t* h_active, float* opt_logits,
                                             const int64_t* opt_forced, void* stream) {
  B200VQA_REQUIRE(image_idx != nullptr && n_images >= 1, "image_idx is NULL or there are no images");
  return fa_run_chain_impl(h, img_tokens_bf16, image_idx, n_images, func, deps, n_steps, B, S, start_token, max_len,
                           cache, h_active, opt_logits, opt_forced, stream);
}

static int shared_ingest_stream(int device, cudaStream_t* out) {
  static std::mutex mu;
  static std::map<int, cudaStream_t> streams;  // lives as long as the process (like the context it belongs to)
  std::lock_guard<std::mutex> lk(mu);
  auto it = streams.find(device);
  if (it == streams.end()) {
    cudaStream_t st = nullptr;
    B200VQA_CUDA_OK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    it = streams.emplace(device, st).first;
  }
  *out = it->second;
  return B200VQA_OK;
}

// Host buffers in, host cache out: the reference's driver loop does one H2D and one D2H PER PROGRAM STEP
// (FA:193-206 -> 109-121); here a sub-batch of questions is uploaded once (double-buffered, projected as it lands),
// executed longest-program-first with the cache resident in HBM, and its cache rows are downloaded once - while the
// next sub-batch uploads.
static int fa_run_chain_host_impl(b200vqa_handle* h, const float* h_img, const int32_t* h_func, const int32_t* h_deps,
                                  const int32_t* h_n_steps, int B, int S, int start_token, int max_len,
                                  int32_t* h_cache, int chunk, void* stream, bool sync) {
  B200VQA_REQUIRE(h != nullptr, "handle is NULL");
  B200VQA_REQUIRE(h->d.kind == B200VQA_MODEL_FA, "handle was not created for the FA model");
  B200VQA_REQUIRE(B >= 0 && S >= 0, "negative batch or step count");
  if (B == 0 || S == 0) return B200VQA_OK;
  B200VQA_REQUIRE(h_img && h_func && h_deps && h_n_steps && h_cache, "a required buffer is NULL");
  B200VQA_REQUIRE(max_len >= 2 && 1 + 2 * max_len <= 60, "max_len %d out of range (2..29)", max_len);
  RC_OK(set_device(h));
  const auto& d = h->d;
  if (chunk <= 0) chunk = 2048;
  chunk = std::min({chunk, B, default_cap(h)});
  const int ichunk = std::min(chunk, 256);  // images per upload: 256 x 803 KB = 205 MB of staging
  RC_OK(ensure_workspace(h, chunk, std::max(max_len - 1, 20)));  // 20: what b200vqa_fa_project_images asks for
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!h->copy_stream) {
    B200VQA_CUDA_OK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      B200VQA_CUDA_OK(cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming));
      B200VQA_CUDA_OK(cudaEventCreateWithFlags(&h->ev_free[i], cudaEventDisableTiming));
    }
  }
  // upload + image projection of sub-batch k+1 run on the ingest stream while k executes on `s`.  ONE ingest stream per
  // device, shared by every handle: with several calls in flight on different (handle, stream) slots their uploads
  // queue up in submission order instead of splitting the PCIe link (three concurrent 60-ms uploads would all land
  // after 180 ms and none of the chains could start earlier)
  cudaStream_t ingest = nullptr;
  RC_OK(shared_ingest_stream(h->device, &ingest));
  auto up = [](size_t n) { return (n + 255) & ~size_t(255); };
  const size_t per_img = size_t(d.n_img_tokens) * d.img_feat_dim;
  const size_t img_b = up(size_t(ichunk) * per_img * sizeof(float));
  const size_t tok_b = up(size_t(chunk) * d.n_img_tokens * kD * sizeof(__nv_bfloat16));
  const size_t func_b = up(size_t(chunk) * S * sizeof(int32_t));
  const size_t deps_b = up(size_t(chunk) * S * 2 * sizeof(int32_t));
  const size_t ns_b = up(size_t(chunk) * sizeof(int32_t));
  const size_t cache_b = up(size_t(chunk) * S * max_len * sizeof(int32_t));
  // image tokens, program tables and caches are double-buffered: sub-batch k+1 is prepared while k still runs
  const size_t need = img_b + 2 * (tok_b + func_b + deps_b + 2 * ns_b + 2 * cache_b);
  if (h->stage_bytes < need) {
    if (h->stage) {
      B200VQA_CUDA_OK(cudaDeviceSynchronize());
      B200VQA_CUDA_OK(cudaFree(h->stage));
      h->stage = nullptr;
      h->stage_bytes = 0;
      h->tmaps.clear();
    }
    B200VQA_CUDA_OK(cudaMalloc(&h->stage, need));
    h->stage_bytes = need;
  }
  uint8_t* p = h->stage;
  float* d_img = reinterpret_cast<float*>(p); p += img_b;
  __nv_bfloat16* d_tok[2];
  int32_t *d_func[2], *d_deps[2], *d_ns[2], *d_order[2], *d_cache_sorted[2], *d_cache[2];
  for (int i = 0; i < 2; ++i) {
    d_tok[i] = reinterpret_cast<__nv_bfloat16*>(p); p += tok_b;
    d_func[i] = reinterpret_cast<int32_t*>(p); p += func_b;
    d_deps[i] = reinterpret_cast<int32_t*>(p); p += deps_b;
    d_ns[i] = reinterpret_cast<int32_t*>(p); p += ns_b;
    d_order[i] = reinterpret_cast<int32_t*>(p); p += ns_b;
    d_cache_sorted[i] = reinterpret_cast<int32_t*>(p); p += cache_b;
    d_cache[i] = reinterpret_cast<int32_t*>(p); p += cache_b;
  }
  // staging may still be read by earlier unsynchronised work of `s`
  B200VQA_CUDA_OK(cudaEventRecord(h->ev_free[0], s));
  B200VQA_CUDA_OK(cudaStreamWaitEvent(ingest, h->ev_free[0], 0));

  // B200VQA_FA_HOST_TRACE=1: GPU timeline of the call (ms since its start) printed to stderr after the final sync
  const bool trace = getenv("B200VQA_FA_HOST_TRACE") != nullptr;
  std::vector<cudaEvent_t> tr_ev;
  std::vector<const char*> tr_name;
  auto mark = [&](const char* name, cudaStream_t st) {
    if (!trace) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    tr_ev.push_back(e);
    tr_name.push_back(name);
  };
  mark("start (compute stream)", s);
  const bool convert = h->host_upload_f16;  // b200vqa_set_host_upload(1): for this model the half-width format is bf16
  int pin_turn = 0;
  if (convert) {
    const size_t pin_b = size_t(ichunk) * per_img * 2;
    if (h->pin16_bytes < pin_b) {
      for (int i = 0; i < 2; ++i) {
        if (h->pin_pending[i]) B200VQA_CUDA_OK(cudaEventSynchronize(h->ev_pin[i]));
        h->pin_pending[i] = false;
        if (h->pin16[i]) B200VQA_CUDA_OK(cudaFreeHost(h->pin16[i]));
        h->pin16[i] = nullptr;
      }
      h->pin16_bytes = 0;
      for (int i = 0; i < 2; ++i) B200VQA_CUDA_OK(cudaMallocHost(&h->pin16[i], pin_b));
      h->pin16_bytes = pin_b;
    }
    for (int i = 0; i < 2; ++i)
      if (!h->ev_pin[i]) B200VQA_CUDA_OK(cudaEventCreateWithFlags(&h->ev_pin[i], cudaEventDisableTiming));
  }
  // ---- host: every sub-batch longest program first (stable), so that finished questions drop out of the later
  // steps.  The sorted tables of the WHOLE call are written to pinned memory up front: a copy from pageable memory would
  // make the host wait for everything queued before it on the stream, i.e. for the previous sub-batch's chains
  const int n_sub = (B + chunk - 1) / chunk;
  const size_t tab_ints = size_t(B) * (size_t(S) * 3 + 2);
  if (h->tables_pending) {  // the previous call on this handle may still be uploading from the table buffer
    B200VQA_CUDA_OK(cudaEventSynchronize(h->ev_tables));
    h->tables_pending = false;
  }
  if (h->h_tables_ints < tab_ints) {
    if (h->h_tables) B200VQA_CUDA_OK(cudaFreeHost(h->h_tables));
    h->h_tables = nullptr;
    h->h_tables_ints = 0;
    B200VQA_CUDA_OK(cudaMallocHost(&h->h_tables, tab_ints * sizeof(int32_t)));
    h->h_tables_ints = tab_ints;
  }
  if (!h->ev_tables) B200VQA_CUDA_OK(cudaEventCreateWithFlags(&h->ev_tables, cudaEventDisableTiming));
  std::vector<int32_t> active(size_t(n_sub) * S);
  auto tab_func = [&](int k) { return h->h_tables + size_t(k) * chunk * (size_t(S) * 3 + 2); };
  for (int k = 0; k < n_sub; ++k) {
    const int b0 = k * chunk;
    const int nb = std::min(chunk, B - b0);
    int32_t* t_func = tab_func(k);
    int32_t* t_deps = t_func + size_t(nb) * S;
    int32_t* t_ns = t_deps + size_t(nb) * S * 2;
    int32_t* t_order = t_ns + nb;
    for (int i = 0; i < nb; ++i) t_order[i] = i;
    std::stable_sort(t_order, t_order + nb, [&](int a, int b) { return h_n_steps[b0 + a] > h_n_steps[b0 + b]; });
    for (int i = 0; i < nb; ++i) {
      const size_t src = size_t(b0) + t_order[i];
      std::memcpy(t_func + size_t(i) * S, h_func + src * S, size_t(S) * sizeof(int32_t));
      std::memcpy(t_deps + size_t(i) * S * 2, h_deps + src * S * 2, size_t(S) * 2 * sizeof(int32_t));
      t_ns[i] = h_n_steps[src];
    }
    for (int i = 0; i < S; ++i) {
      int a = 0;
      while (a < nb && t_ns[a] > i) ++a;
      active[size_t(k) * S + i] = a;
    }
  }
  // ingest stream: features of sub-batch k in caller order, projected as they land (image_proj + PE once per question),
  // then its program tables.  Enqueued one sub-batch AHEAD of the chains on the host as well, so that the next upload
  // is already queued when the host is held up on the launch queue by a chain (hundreds of launches)
  auto enqueue_ingest = [&](int k) -> int {
    const int b0 = k * chunk;
    const int nb = std::min(chunk, B - b0);
    const int par = k & 1;
    // chain k-2 (and the scatter + download behind it) has released d_tok / the tables / the caches of this parity
    if (k >= 2) B200VQA_CUDA_OK(cudaStreamWaitEvent(ingest, h->ev_free[par], 0));
    for (int i0 = 0; i0 < nb; i0 += ichunk) {
      const int ni = std::min(ichunk, nb - i0);
      const float* src = h_img + (size_t(b0) + i0) * per_img;
      if (convert) {
        // the host rounds these images to bf16 (what the transpose on the device does anyway: identical tokens) while
        // the previous group is on the wire; the pinned staging is free once its own upload has completed
        const int ps = pin_turn++ & 1;
        if (h->pin_pending[ps]) {
          B200VQA_CUDA_OK(cudaEventSynchronize(h->ev_pin[ps]));
          h->pin_pending[ps] = false;
        }
        host_f32_to_bf16(src, h->pin16[ps], size_t(ni) * per_img, 0);
        B200VQA_CUDA_OK(cudaMemcpyAsync(d_img, h->pin16[ps], size_t(ni) * per_img * 2, cudaMemcpyHostToDevice, ingest));
        B200VQA_CUDA_OK(cudaEventRecord(h->ev_pin[ps], ingest));
        h->pin_pending[ps] = true;
      } else {
        B200VQA_CUDA_OK(cudaMemcpyAsync(d_img, src, size_t(ni) * per_img * sizeof(float), cudaMemcpyHostToDevice, ingest));
      }
      RC_OK(fa_project_images_impl(h, d_img, convert, ni, d_tok[par] + size_t(i0) * d.n_img_tokens * kD, ingest));
    }
    const int32_t* t_func = tab_func(k);
    const int32_t* t_deps = t_func + size_t(nb) * S;
    const int32_t* t_ns = t_deps + size_t(nb) * S * 2;
    const int32_t* t_order = t_ns + nb;
    B200VQA_CUDA_OK(cudaMemcpyAsync(d_func[par], t_func, size_t(nb) * S * 4, cudaMemcpyHostToDevice, ingest));
    B200VQA_CUDA_OK(cudaMemcpyAsync(d_deps[par], t_deps, size_t(nb) * S * 8, cudaMemcpyHostToDevice, ingest));
    B200VQA_CUDA_OK(cudaMemcpyAsync(d_ns[par], t_ns, size_t(nb) * 4, cudaMemcpyHostToDevice, ingest));
    B200VQA_CUDA_OK(cudaMemcpyAsync(d_order[par], t_order, size_t(nb) * 4, cudaMemcpyHostToDevice, ingest));
    B200VQA_CUDA_OK(cudaMemsetAsync(d_cache_sorted[par], 0xff, size_t(nb) * S * max_len * sizeof(int32_t), ingest));
    B200VQA_CUDA_OK(cudaEventRecord(h->ev_in[par], ingest));
    mark("ingest done", ingest);
    return B200VQA_OK;
  };
  RC_OK(enqueue_ingest(0));
  int sub = 0;
  for (int b0 = 0; b0 < B; b0 += chunk, ++sub) {
    const int nb = std::min(chunk, B - b0);
    const int par = sub & 1;
    // sub-batch sub+1 may be uploaded as soon as chain sub-1 (which uses the same buffers) has been ENQUEUED: its
    // release event exists by then
    if (sub + 1 < n_sub) RC_OK(enqueue_ingest(sub + 1));
    B200VQA_CUDA_OK(cudaStreamWaitEvent(s, h->ev_in[par], 0));
    mark("chain start", s);
    // question i of the sorted batch uses the image tokens of question order[i]
    RC_OK(fa_run_chain_impl(h, d_tok[par], d_order[par], nb, d_func[par], d_deps[par], d_ns[par], nb, S, start_token,
                            max_len, d_cache_sorted[par], &active[size_t(sub) * S], nullptr, nullptr, s));
    h->cur_tag = kTagMisc;
    LAUNCH_OK(h, launch_scatter_rows_i32(d_cache_sorted[par], d_order[par], nb, S * max_len, d_cache[par], s));
    B200VQA_CUDA_OK(cudaMemcpyAsync(h_cache + size_t(b0) * S * max_len, d_cache[par],
                                    size_t(nb) * S * max_len * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    B200VQA_CUDA_OK(cudaEventRecord(h->ev_free[par], s));
    mark("chain + download done", s);
  }
  B200VQA_CUDA_OK(cudaEventRecord(h->ev_tables, ingest));
  h->tables_pending = true;
  if (sync || trace) B200VQA_CUDA_OK(cudaStreamSynchronize(s));
  if (trace) {
    B200VQA_CUDA_OK(cudaStreamSynchronize(ingest));
    for (size_t i = 0; i < tr_ev.size(); ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, tr_ev[0], tr_ev[i]);
      fprintf(stderr, "b200vqa fa_run_chain_host: %-28s %8.2f ms\n", tr_name[i], ms);
      if (i) cudaEventDestroy(tr_ev[i]);
    }
    cudaEventDestroy(tr_ev[0]);
  }
  return B200VQA_OK;
}

B200VQA_API int b200vqa_fa_run_chain_host(b200vqa_handle* h, const float* h_img, const int32_t* h_func,
                                          const int32_t* h_deps, const int32_t* h_n_steps, int B, int S, int start_token,
                                          int max_len, int32_t* h_cache, int chunk, void* stream) {
  return fa_run_chain_host_impl(h, h_img, h_func, h_deps, h_n_steps, B, S, start_token, max_len, h_cache, chunk, stream,
                                true);
}

B200VQA_API int b200vqa_fa_run_chain_host_async(b200vqa_handle* h, const float* h_img, const int32_t* h_func,
                                                const int32_t* h_deps, const int32_t* h_n_steps, int B, int S,
                                                int start_token, int max_len, int32_t* h_cache, int chunk, void* stream) {
  return fa_run_chain_host_impl(h, h_img, h_func, h_deps, h_n_steps, B, S, start_token, max_len, h_cache, chunk, stream,
                                false);
}

// ------------------------------------------------------------------------------------------------ test hooks
B200VQA_API int b200vqa_dbg_gemm(const b200vqa_dbg_gemm_args* a, void* stream) {
  B200VQA_REQUIRE(a != nullptr, "args is NULL");
  int dev = 0, num_sms = 0;
  B200VQA_CUDA_OK(cudaGetDevice(&dev));
  RC_OK(require_sm100(dev, &num_sms));
  const TmapType ty = a->tf32 ? TmapType::kF32 : TmapType::kBF16;
  CUtensorMap ta, tw;
  RC_OK(make_tmap_2d(&ta, a->A, ty, uint64_t(a->M), uint64_t(a->K), uint64_t(a->K), 128));
  RC_OK(make_tmap_2d(&tw, a->W, ty, uint64_t(a->N), uint64_t(a->K), uint64_t(a->K), uint32_t(a->block_n)));
  GemmParams p;
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.bias = a->bias;
  p.out = static_cast<__nv_bfloat16*>(a->out);
  p.ldc = a->ldc;
  p.residual = static_cast<const __nv_bfloat16*>(a->residual);
  p.ldr = a->N;
  p.gamma = a->gamma;
  p.beta = a->beta;
  p.out_f32 = a->out_f32;
  p.rows_in = a->rows_in > 0 ? a->rows_in : 1;
  p.rows_out = a->rows_out > 0 ? a->rows_out : 1;
  p.row_off = a->row_off;
  p.pe = a->pe;
  p.pe_off = a->pe_off;
  p.dbg_clk = a->clk;
  B200VQA_CUDA_OK(launch_gemm(a->epilogue, a->tf32 != 0, a->block_n, ta, tw, p, num_sms,
                              static_cast<cudaStream_t>(stream)));
  return B200VQA_OK;
}

B200VQA_API int b200vqa_dbg_gemm_check(int a_is_f32, const void* A, const void* W, const float* bias, float* out, int M, int N,
                           int K, void* stream) {
  B200VQA_CUDA_OK(launch_gemm_check(a_is_f32 != 0, A, W, bias, out, M, N, K, static_cast<cudaStream_t>(stream)));
  return B200VQA_OK;
}

B200VQA_API int b200vqa_dbg_enc_attention(const void* qkv, const int32_t* lens, int const_len, int B, int nhead, int v_mode,
                              void* out, void* stream) {
  B200VQA_REQUIRE(qkv && out && B > 0 && (nhead == 2 || nhead == 4), "bad arguments");
  int dev = 0;
  B200VQA_CUDA_OK(cudaGetDevice(&dev));
  RC_OK(require_sm100(dev, nullptr));
  CUtensorMap tq, tkv;
  RC_OK(make_tmap_2d(&tq, qkv, TmapType::kBF16, uint64_t(B) * kLP, 3 * kD, 3 * kD, 128));
  RC_OK(make_tmap_2d(&tkv, qkv, TmapType::kBF16, uint64_t(B) * kLP, 3 * kD, 3 * kD, 256));
  EncAttnParams ap;
  ap.B = B;
  ap.nhead = nhead;
  ap.lens = lens;
  ap.const_len = const_len;
  ap.out = static_cast<__nv_bfloat16*>(out);
  ap.scale = 1.f / sqrtf(float(kD / nhead));
  ap.v_mode = v_mode;
  B200VQA_CUDA_OK(launch_enc_attention(tq, tkv, static_cast<const __nv_bfloat16*>(qkv), ap,
                                       static_cast<cudaStream_t>(stream)));
  return B200VQA_OK;
}

B200VQA_API int b200vqa_dbg_workspace(b200vqa_handle* h, int which, void* dst, size_t dst_bytes, size_t* bytes) {
  B200VQA_REQUIRE(h != nullptr && bytes, "NULL argument");
  void* src = nullptr;
  void** ptr = &src;
  B200VQA_REQUIRE(h->ws.base != nullptr, "the workspace has not been allocated yet");
  const Workspace& w = h->ws;
  const size_t r = w.drows, nh = size_t(h->d.nhead);
  switch (which) {
    case 0: *ptr = w.dx; *bytes = r * kD * 2; break;
    case 1: *ptr = w.dqkv; *bytes = r * 3 * kD * 2; break;
    case 2: *ptr = w.dattn; *bytes = r * kD * 2; break;
    case 3: *ptr = w.dx1; *bytes = r * kD * 2; break;
    case 4: *ptr = w.dq; *bytes = r * nh * kD * 2; break;
    case 5: *ptr = w.du; *bytes = r * nh * kD * 2; break;
    case 6: *ptr = w.dx2; *bytes = r * kD * 2; break;
    case 7: *ptr = w.dxo[0]; *bytes = r * kD * 2; break;
    case 8: *ptr = w.dxo[1]; *bytes = r * kD * 2; break;
    case 9: *ptr = w.dout; *bytes = r * kD * 4; break;
    case 10: *ptr = w.tok; *bytes = size_t(w.cap) * kTokLd * 8; break;
    case 11:
      B200VQA_REQUIRE(h->persist_clk != nullptr, "no stage timeline: create the handle with B200VQA_PERSIST_STAMPS=1");
      *ptr = h->persist_clk; *bytes = 8 * 24 * sizeof(long long); break;
    default: set_error("unknown workspace buffer %d", which); return B200VQA_ERR_BAD_ARGUMENT;
  }
  if (dst && dst_bytes > 0) {
    RC_OK(set_device(h));
    B200VQA_CUDA_OK(cudaDeviceSynchronize());
    B200VQA_CUDA_OK(cudaMemcpy(dst, src, std::min(dst_bytes, *bytes), cudaMemcpyDeviceToDevice));
  }
  return B200VQA_OK;
}

B200VQA_API int b200vqa_dbg_mem_attn(const void* qp, const void* memory, const int32_t* lens, int const_len, int B, int nhead,
                                     int impl, void* out, long long* stamps, void* stream) {
  B200VQA_REQUIRE(qp && memory && out && B > 0 && (nhead == 2 || nhead == 4), "bad arguments");
  int dev = 0;
  B200VQA_CUDA_OK(cudaGetDevice(&dev));
  RC_OK(require_sm100(dev, nullptr));
  MemAttnParams mp;
  mp.B = B;
  mp.nhead = nhead;
  mp.qp = static_cast<const __nv_bfloat16*>(qp);
  mp.rows_per_q = kLP;
  mp.lens = lens;
  mp.const_len = const_len;
  mp.out = static_cast<__nv_bfloat16*>(out);
  if (const char* g = getenv("B200VQA_NO_L2_HINTS")) mp.l2_evict_first = !(g[0] && g[0] != '0');
  B200VQA_REQUIRE(impl == 0 && stamps == nullptr, "only the warp-MMA kernel (impl 0) exists");
  CUtensorMap tm;
  RC_OK(make_tmap_2d(&tm, memory, TmapType::kBF16, uint64_t(B) * kLP, kD, kD, kMemAttnTileRows));
  B200VQA_CUDA_OK(launch_mem_attn(tm, mp, static_cast<cudaStream_t>(stream)));
  return B200VQA_OK;
}

}  // extern "C"
