// C-ABI of the LSTM program generator: question tokens -> 27 program tokens.
// Replaces Seq2SeqModel.forward of the reference's code/run_model_lstm_qp.py:277-319:
//   embedding(85,256) -> LSTM(256,512) over the 46 question tokens -> 27 x [embedding(prev token) -> LSTM cell ->
//   Linear(512,44) -> argmax].
//
// Both LSTMs only ever see embedding rows, so the input half of the gates is a table look-up:
//   T[v] = embedding[v] . W_ih^T + b_ih + b_hh        ([vocab, 2048] fp32, built once per weight set)
// and a time step is ONE tensor-core GEMM, gates = T[token] + h_prev . W_hh^T (M = batch, K = 512, N = 2048), whose
// epilogue applies the cell update (gemm.cu, kEpiLstm).  W_hh rows / table columns are permuted so that every
// 256-column n-tile holds i|f|g|o of the same 64 hidden units.  The decoder adds a tf32 head GEMM with fused argmax
// per step.  The 46 + 2*27 launches are PDL-linked and replayed as a CUDA graph per batch size.
#include <algorithm>
#include <cstdlib>
#include <map>
#include <new>

#include "host_util.h"
#include "kernels.h"

using namespace b200vqa;

namespace {
constexpr int kEmb = 256;
constexpr int kHid = 512;
constexpr int kGates = 4 * kHid;
constexpr int kLstmTokLd = 65;
constexpr int kLstmCap = 4096;

// permuted gate column n (tile nt = n / 256, gate g = (n % 256) / 64, unit u = n % 64) <- torch row g*512 + 64*nt + u
__device__ __forceinline__ int torch_gate_row(int n) { return ((n & 255) >> 6) * kHid + (n >> 8) * 64 + (n & 63); }

__global__ void lstm_pack_whh_kernel(const float* __restrict__ w_hh, __nv_bfloat16* __restrict__ out) {
  const int n = blockIdx.x;  // permuted row
  const float* src = w_hh + size_t(torch_gate_row(n)) * kHid;
  for (int k = threadIdx.x; k < kHid; k += blockDim.x) out[size_t(n) * kHid + k] = __float2bfloat16(src[k]);
}

// T[v][n] = emb[v] . w_ih[row(n)] + b_ih[row(n)] + b_hh[row(n)]   (fp32)
__global__ void lstm_table_kernel(const float* __restrict__ emb, const float* __restrict__ w_ih,
                                  const float* __restrict__ b_ih, const float* __restrict__ b_hh,
                                  float* __restrict__ table) {
  const int v = blockIdx.y;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= kGates) return;
  const int r = torch_gate_row(n);
  const float* e = emb + size_t(v) * kEmb;
  const float* w = w_ih + size_t(r) * kEmb;
  float acc = b_ih[r] + b_hh[r];
  for (int k = 0; k < kEmb; ++k) acc = fmaf(e[k], w[k], acc);
  table[size_t(v) * kGates + n] = acc;
}
}  // namespace

struct b200vqa_lstm {
  int device = 0, num_sms = 148;
  b200vqa_lstm_desc d{};
  uint8_t* wbase = nullptr;
  __nv_bfloat16 *whh_enc = nullptr, *whh_dec = nullptr;
  float *t_enc = nullptr, *t_dec = nullptr, *fc_w = nullptr, *fc_b = nullptr;
  // workspace (kLstmCap questions)
  uint8_t* ws = nullptr;
  __nv_bfloat16* h[2] = {nullptr, nullptr};
  float *hf = nullptr, *c = nullptr;
  int64_t *tok = nullptr, *q = nullptr;
  std::map<std::tuple<int, int, int, int>, cudaGraphExec_t> graphs;
  cudaStream_t cap_stream = nullptr;
  uint64_t launches = 0;
};

namespace {

#define RC_OK(expr)                    \
  do {                                 \
    int rc_ = (expr);                  \
    if (rc_ != B200VQA_OK) return rc_; \
  } while (0)

int lstm_step(b200vqa_lstm* h, int B, int src, const __nv_bfloat16* whh, const float* table, const int64_t* tokens,
              int tok_ld, int tok_col, int tok_const, bool want_f32, cudaStream_t s) {
  CUtensorMap ta, tw;
  RC_OK(make_tmap_2d(&ta, h->h[src], TmapType::kBF16, uint64_t(B), kHid, kHid, 128));
  RC_OK(make_tmap_2d(&tw, whh, TmapType::kBF16, kGates, kHid, kHid, 256));
  GemmParams p;
  p.M = B;
  p.N = kGates;
  p.K = kHid;
  p.pdl = true;
  p.lstm_table = table;
  p.lstm_tokens = tokens;
  p.lstm_tok_ld = tok_ld;
  p.lstm_tok_col = tok_col;
  p.lstm_token_const = tok_const;
  p.lstm_vocab = h->d.vocab;
  p.lstm_c = h->c;
  p.lstm_h = h->h[src ^ 1];
  p.lstm_h_f32 = want_f32 ? h->hf : nullptr;
  B200VQA_CUDA_OK(launch_gemm(kEpiLstm, false, 256, ta, tw, p, h->num_sms, s));
  ++h->launches;
  return B200VQA_OK;
}

// the whole generate sequence for B questions already copied into h->q (fixed arguments -> graph-capturable)
int enqueue_generate(b200vqa_lstm* h, int B, int q_len, int T, int start_token, float* logits, const int64_t* forced,
                     cudaStream_t s) {
  B200VQA_CUDA_OK(cudaMemsetAsync(h->h[0], 0, size_t(B) * kHid * sizeof(__nv_bfloat16), s));
  B200VQA_CUDA_OK(cudaMemsetAsync(h->c, 0, size_t(B) * kHid * sizeof(float), s));
  int cur = 0;
  for (int t = 0; t < q_len; ++t) {  // encoder (run_model_lstm_qp.py:294): every position, padding included
    RC_OK(lstm_step(h, B, cur, h->whh_enc, h->t_enc, h->q, q_len, t, 0, false, s));
    cur ^= 1;
  }
  const int bn = h->d.prog_vocab <= 64 ? 64 : 256;
  for (int t = 0; t < T; ++t) {  // decoder (run_model_lstm_qp.py:311-317)
    const int64_t* toks = nullptr;
    int ld = 0, col = 0;
    if (t > 0) {
      toks = forced ? forced : h->tok;
      ld = forced ? T : kLstmTokLd;
      col = forced ? t - 1 : t;
    }
    RC_OK(lstm_step(h, B, cur, h->whh_dec, h->t_dec, toks, ld, col, start_token, true, s));
    cur ^= 1;
    CUtensorMap ta, tw;
    RC_OK(make_tmap_2d(&ta, h->hf, TmapType::kF32, uint64_t(B), kHid, kHid, 128));
    RC_OK(make_tmap_2d(&tw, h->fc_w, TmapType::kF32, uint64_t(h->d.prog_vocab), kHid, kHid, uint32_t(bn)));
    GemmParams p;
    p.M = B;
    p.N = bn;
    p.K = kHid;
    p.pdl = true;
    p.bias = h->fc_b;
    p.head_V = h->d.prog_vocab;
    p.head_t = t;
    p.tok = h->tok;
    p.tok_ld = kLstmTokLd;
    p.logits = logits;
    p.logits_T = T;
    B200VQA_CUDA_OK(launch_gemm(kEpiHead, true, bn, ta, tw, p, h->num_sms, s));
    ++h->launches;
  }
  return B200VQA_OK;
}

}  // namespace

extern "C" {

B200VQA_API int b200vqa_lstm_create(const b200vqa_lstm_desc* d, int device, b200vqa_lstm** out) {
  B200VQA_REQUIRE(out != nullptr, "out handle pointer is NULL");
  *out = nullptr;
  B200VQA_REQUIRE(d != nullptr, "descriptor is NULL");
  if (d->embedding_dim != kEmb || d->hidden_dim != kHid) {
    set_error("LSTM generator kernels are specialised for embedding 256 / hidden 512 (got %d / %d)", d->embedding_dim,
              d->hidden_dim);
    return B200VQA_ERR_UNSUPPORTED_SHAPE;
  }
  if (d->vocab <= 0 || d->prog_vocab <= 0 || d->prog_vocab > 256 || d->prog_vocab > d->vocab) {
    set_error("vocabulary sizes out of range (question %d, program %d <= min(256, question vocab))", d->vocab,
              d->prog_vocab);
    return B200VQA_ERR_UNSUPPORTED_SHAPE;
  }
  B200VQA_REQUIRE(d->embedding && d->enc_w_ih && d->enc_w_hh && d->enc_b_ih && d->enc_b_hh && d->dec_w_ih && d->dec_w_hh &&
                      d->dec_b_ih && d->dec_b_hh && d->fc_w && d->fc_b,
                  "a required weight pointer is NULL");
  int sms = 0;
  RC_OK(require_sm100(device, &sms));
  B200VQA_CUDA_OK(cudaSetDevice(device));
  b200vqa_lstm* h = new (std::nothrow) b200vqa_lstm();
  if (!h) {
    set_error("out of host memory");
    return B200VQA_ERR_OUT_OF_MEMORY;
  }
  h->device = device;
  h->num_sms = sms;
  h->d = *d;
  const size_t whh_b = size_t(kGates) * kHid * sizeof(__nv_bfloat16);
  const size_t tab_b = size_t(d->vocab) * kGates * sizeof(float);
  const size_t fc_b = (size_t(d->prog_vocab) * kHid * sizeof(float) + 255) & ~size_t(255);
  const size_t total = 2 * whh_b + 2 * tab_b + fc_b + 1024;
  if (cudaMalloc(&h->wbase, total) != cudaSuccess) {
    set_error("cudaMalloc(%zu) for the LSTM weights failed", total);
    delete h;
    return B200VQA_ERR_OUT_OF_MEMORY;
  }
  uint8_t* p = h->wbase;
  h->whh_enc = reinterpret_cast<__nv_bfloat16*>(p); p += whh_b;
  h->whh_dec = reinterpret_cast<__nv_bfloat16*>(p); p += whh_b;
  h->t_enc = reinterpret_cast<float*>(p); p += tab_b;
  h->t_dec = reinterpret_cast<float*>(p); p += tab_b;
  h->fc_w = reinterpret_cast<float*>(p); p += fc_b;
  h->fc_b = reinterpret_cast<float*>(p);
  lstm_pack_whh_kernel<<<kGates, 128>>>(d->enc_w_hh, h->whh_enc);
  lstm_pack_whh_kernel<<<kGates, 128>>>(d->dec_w_hh, h->whh_dec);
  dim3 tg(kGates / 128, d->vocab);
  lstm_table_kernel<<<tg, 128>>>(d->embedding, d->enc_w_ih, d->enc_b_ih, d->enc_b_hh, h->t_enc);
  lstm_table_kernel<<<tg, 128>>>(d->embedding, d->dec_w_ih, d->dec_b_ih, d->dec_b_hh, h->t_dec);
  cudaMemcpyAsync(h->fc_w, d->fc_w, size_t(d->prog_vocab) * kHid * sizeof(float), cudaMemcpyDeviceToDevice, nullptr);
  cudaMemcpyAsync(h->fc_b, d->fc_b, size_t(d->prog_vocab) * sizeof(float), cudaMemcpyDeviceToDevice, nullptr);
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaGetLastError();
  // workspace
  const size_t rows = kLstmCap;
  const size_t hb = rows * kHid * sizeof(__nv_bfloat16), fb = rows * kHid * sizeof(float);
  const size_t tb = rows * kLstmTokLd * sizeof(int64_t), qb = rows * 64 * sizeof(int64_t);
  if (e == cudaSuccess) e = cudaMalloc(&h->ws, 2 * hb + 2 * fb + tb + qb);
  if (e != cudaSuccess) {
    set_error("LSTM generator setup failed: %s", cudaGetErrorString(e));
    cudaFree(h->wbase);
    delete h;
    return B200VQA_ERR_CUDA;
  }
  p = h->ws;
  h->h[0] = reinterpret_cast<__nv_bfloat16*>(p); p += hb;
  h->h[1] = reinterpret_cast<__nv_bfloat16*>(p); p += hb;
  h->hf = reinterpret_cast<float*>(p); p += fb;
  h->c = reinterpret_cast<float*>(p); p += fb;
  h->tok = reinterpret_cast<int64_t*>(p); p += tb;
  h->q = reinterpret_cast<int64_t*>(p);
  *out = h;
  return B200VQA_OK;
}

B200VQA_API void b200vqa_lstm_destroy(b200vqa_lstm* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  for (auto& kv : h->graphs) cudaGraphExecDestroy(kv.second);
  if (h->cap_stream) cudaStreamDestroy(h->cap_stream);
  if (h->ws) cudaFree(h->ws);
  if (h->wbase) cudaFree(h->wbase);
  delete h;
}

B200VQA_API uint64_t b200vqa_lstm_launch_count(const b200vqa_lstm* h) { return h ? h->launches : 0; }

B200VQA_API int b200vqa_programs_to_chain(const int64_t* programs, int B, int T, const int32_t* arity,
                                          const int32_t* func_map, int prog_vocab, int S, int32_t* func, int32_t* deps,
                                          int32_t* n_steps, void* stream) {
  B200VQA_REQUIRE(B >= 0 && T >= 1 && T <= 64 && S >= 1 && prog_vocab >= 1, "shape out of range (B %d, T %d, S %d)", B, T, S);
  if (B == 0) return B200VQA_OK;
  B200VQA_REQUIRE(programs && arity && func_map && func && deps && n_steps, "a required buffer is NULL");
  ProgToChainParams p;
  p.B = B;
  p.T = T;
  p.S = S;
  p.prog_vocab = prog_vocab;
  p.programs = programs;
  p.arity = arity;
  p.func_map = func_map;
  p.func = func;
  p.deps = deps;
  p.n_steps = n_steps;
  B200VQA_CUDA_OK(launch_programs_to_chain(p, static_cast<cudaStream_t>(stream)));
  return B200VQA_OK;
}

B200VQA_API int b200vqa_lstm_generate(b200vqa_lstm* h, const int64_t* questions, int q_len, int B, int program_len,
                                      int start_token, int64_t* programs, float* opt_logits, const int64_t* opt_forced,
                                      void* stream) {
  B200VQA_REQUIRE(h != nullptr, "handle is NULL");
  B200VQA_REQUIRE(B >= 0 && q_len >= 1 && q_len <= 64 && program_len >= 1 && program_len < kLstmTokLd,
                  "shape out of range (B %d, question length %d (1..64), program length %d (1..64))", B, q_len,
                  program_len);
  if (B == 0) return B200VQA_OK;
  B200VQA_REQUIRE(questions && programs, "a required buffer is NULL");
  B200VQA_CUDA_OK(cudaSetDevice(h->device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int T = program_len, V = h->d.prog_vocab;
  const int st = std::min(std::max(start_token, 0), h->d.vocab - 1);
  for (int b0 = 0; b0 < B; b0 += kLstmCap) {
    const int nb = std::min(kLstmCap, B - b0);
    B200VQA_CUDA_OK(cudaMemcpyAsync(h->q, questions + size_t(b0) * q_len, size_t(nb) * q_len * sizeof(int64_t),
                                    cudaMemcpyDeviceToDevice, s));
    float* lg = opt_logits ? opt_logits + size_t(b0) * T * V : nullptr;
    const int64_t* fz = opt_forced ? opt_forced + size_t(b0) * T : nullptr;
    const bool plain = !lg && !fz && !getenv("B200VQA_NO_GRAPH");
    if (!plain) {
      RC_OK(enqueue_generate(h, nb, q_len, T, st, lg, fz, s));
    } else {
      auto key = std::make_tuple(nb, q_len, T, st);
      auto it = h->graphs.find(key);
      if (it == h->graphs.end()) {
        if (!h->cap_stream) B200VQA_CUDA_OK(cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking));
        const uint64_t before = h->launches;
        cudaGraph_t graph = nullptr;
        B200VQA_CUDA_OK(cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeThreadLocal));
        int rc = enqueue_generate(h, nb, q_len, T, st, nullptr, nullptr, h->cap_stream);
        cudaError_t e = cudaStreamEndCapture(h->cap_stream, &graph);
        if (rc != B200VQA_OK) {
          if (graph) cudaGraphDestroy(graph);
          return rc;
        }
        cudaGraphExec_t exec = nullptr;
        if (e == cudaSuccess) e = cudaGraphInstantiate(&exec, graph, 0);
        if (graph) cudaGraphDestroy(graph);
        if (e != cudaSuccess) {
          set_error("CUDA graph capture of the LSTM generator failed: %s", cudaGetErrorString(e));
          return B200VQA_ERR_CUDA;
        }
        h->launches = before;
        it = h->graphs.emplace(key, exec).first;
      }
      B200VQA_CUDA_OK(cudaGraphLaunch(it->second, s));
      h->launches += uint64_t(q_len + 2 * T);
    }
    PublishParams pp;
    pp.B = nb;
    pp.n_cols = T;
    pp.src_col0 = 1;
    pp.tok = h->tok;
    pp.tok_ld = kLstmTokLd;
    pp.out_i64 = programs + size_t(b0) * T;
    pp.out_ld = T;
    B200VQA_CUDA_OK(launch_publish_tokens(pp, s));
  }
  return B200VQA_OK;
}

}  // extern "C"
