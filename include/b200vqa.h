/*
 * b200vqa.h - C ABI of libb200vqa.so, the sm_100a (B200) implementation of the batched Program Executor
 * inference path of guoyu-zhang/explainable-spatial-vqa.
 *
 * The reference has no FFI: its "operator interface" for this path is two Python files,
 *   code/inference_transformer_iqap.py                    ("IQAP")
 *   code/inference_transformer_full_annotation_new.py     ("FA")
 * Each entry point below names the reference function (file:line) whose arithmetic it replaces.
 * The Python modules of the same names in explainable-spatial-vqa_b200/ bind these symbols with ctypes
 *  and keep the reference's class / function signatures.
 *
 * Conventions
 *   - plain C types only; every pointer is a raw DEVICE pointer unless the name starts with h_ (host).
 *   - the caller owns all input/output buffers; the library owns packed weights and its workspace.
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*) and performs no host sync,
 *     except the *_host entry points, which return after their results are in the host buffers.
 *   - return value 0 = ok, negative = b200vqa_status; the message is in b200vqa_last_error() (thread local).
 *   - there is no CPU fallback: a device that is not compute capability 10.x is an error.
 */
#ifndef B200VQA_H_
#define B200VQA_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define B200VQA_API __attribute__((visibility("default")))
#else
#define B200VQA_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200vqa_handle b200vqa_handle;

typedef enum b200vqa_status {
  B200VQA_OK = 0,
  B200VQA_ERR_BAD_ARGUMENT = -1,
  B200VQA_ERR_UNSUPPORTED_ARCH = -2,
  B200VQA_ERR_OUT_OF_MEMORY = -3,
  B200VQA_ERR_CUDA = -4,
  B200VQA_ERR_UNSUPPORTED_SHAPE = -5
} b200vqa_status;

typedef enum b200vqa_model_kind {
  B200VQA_MODEL_IQAP = 0, /* VQAModel, IQAP:95-241 */
  B200VQA_MODEL_FA = 1    /* MultiModalTransformer, FA:32-58 */
} b200vqa_model_kind;

/* torch.nn.MultiheadAttention parameters, fp32, state-dict layout (in_proj packed q|k|v on dim 0). */
typedef struct b200vqa_mha_weights {
  const float* in_proj_weight;  /* [3d, d] */
  const float* in_proj_bias;    /* [3d]    */
  const float* out_proj_weight; /* [d, d]  */
  const float* out_proj_bias;   /* [d]     */
} b200vqa_mha_weights;

/* torch.nn.TransformerEncoderLayer (post-norm, ReLU), IQAP:118 / FA:42 */
typedef struct b200vqa_encoder_layer_weights {
  b200vqa_mha_weights self_attn;
  const float* linear1_weight; /* [ff, d] */
  const float* linear1_bias;   /* [ff]    */
  const float* linear2_weight; /* [d, ff] */
  const float* linear2_bias;   /* [d]     */
  const float* norm1_weight;
  const float* norm1_bias;
  const float* norm2_weight;
  const float* norm2_bias;
} b200vqa_encoder_layer_weights;

/* torch.nn.TransformerDecoderLayer (post-norm, ReLU), IQAP:132 / FA:42 */
typedef struct b200vqa_decoder_layer_weights {
  b200vqa_mha_weights self_attn;
  b200vqa_mha_weights multihead_attn;
  const float* linear1_weight;
  const float* linear1_bias;
  const float* linear2_weight;
  const float* linear2_bias;
  const float* norm1_weight;
  const float* norm1_bias;
  const float* norm2_weight;
  const float* norm2_bias;
  const float* norm3_weight;
  const float* norm3_bias;
} b200vqa_decoder_layer_weights;

/* Everything b200vqa_create needs: dimensions + fp32 device pointers exactly as in the module's state_dict. */
typedef struct b200vqa_model_desc {
  int32_t kind;          /* b200vqa_model_kind */
  int32_t d_model;       /* 256 (IQAP:15, FA:157); the kernels are specialised for 256 */
  int32_t img_feat_dim;  /* 1024 (IQAP:17, FA:38) */
  int32_t n_img_tokens;  /* 196 */
  int32_t nhead;         /* IQAP 4 (IQAP:118,132); FA ctor argument (FA:158 uses 2) */
  int32_t n_enc_layers;
  int32_t n_dec_layers;
  int32_t dim_ff;        /* IQAP 2048 (torch default); FA ctor argument (512) */
  int32_t enc_vocab;     /* IQAP question vocab / FA vocab */
  int32_t dec_vocab;     /* IQAP program vocab / FA vocab (= head output size) */
  int32_t max_q_len;     /* IQAP 46 (Config.MAX_QUESTION_LEN); FA: max_text_len */
  int32_t pe_enc_len;    /* rows of pos_encoder.pe */
  int32_t pe_dec_len;    /* rows of pos_decoder.pe */
  int32_t answer_hidden; /* IQAP hidden_dim (256) */
  int32_t num_classes;   /* IQAP answer classes */
  float layer_norm_eps;  /* 1e-5 */
  int32_t answer_pool_rows; /* 0: the answer head reads the [CLS] row of the memory (IQAP:176-179); n > 0: it reads the mean
                             * of memory rows 1..n - the bounding-box regressor of train_transformer_iqap_bb.py:304-310
                             * (answer_w0/b0/w1/b1 = bbox_regressor.0 / .2, num_classes = 40 = 10 boxes x 4) */

  const float* image_proj_weight; /* [d, 1024] */
  const float* image_proj_bias;   /* [d] */
  const float* cls_token;         /* IQAP [d]; FA NULL */
  const float* enc_embedding;     /* IQAP embedding.weight / FA text_embedding.weight  [enc_vocab, d] */
  const float* dec_embedding;     /* IQAP program_decoder_embedding.weight / FA text_embedding.weight */
  const float* pe_enc;            /* [pe_enc_len, d] */
  const float* pe_dec;            /* [pe_dec_len, d] */
  const b200vqa_encoder_layer_weights* enc_layers; /* HOST array of n_enc_layers structs */
  const b200vqa_decoder_layer_weights* dec_layers; /* HOST array of n_dec_layers structs */
  const float* enc_final_norm_weight; /* FA transformer.encoder.norm, else NULL */
  const float* enc_final_norm_bias;
  const float* dec_final_norm_weight; /* FA transformer.decoder.norm, else NULL */
  const float* dec_final_norm_bias;
  const float* head_weight;       /* IQAP program_output.weight / FA output_linear.weight [dec_vocab, d] */
  const float* head_bias;
  const float* answer_w0;         /* IQAP answer_classifier.0 [hidden, d] */
  const float* answer_b0;
  const float* answer_w1;         /* IQAP answer_classifier.3 [classes, hidden] */
  const float* answer_b1;
} b200vqa_model_desc;

/* ---------------------------------------------------------------------------------------------------- */
/* Lifetime                                                                                               */
/* ---------------------------------------------------------------------------------------------------- */

/* Replaces: VQAModel.__init__ + load_state_dict (IQAP:96-134, 255-276) / MultiModalTransformer.__init__
 * (FA:33-44, 175-180): packs the fp32 state-dict tensors into the bf16 / tf32 layouts the kernels read. */
B200VQA_API int b200vqa_create(const b200vqa_model_desc* desc, int device, b200vqa_handle** out);
B200VQA_API void b200vqa_destroy(b200vqa_handle* h);
/* Re-pack after the caller changed parameter values in place (load_state_dict on a live module). */
B200VQA_API int b200vqa_refresh_weights(b200vqa_handle* h, const b200vqa_model_desc* desc);
/* <START> token the IQAP decode begins with: the reference reads Config.SPECIAL_TOKEN_ID at decode time (IQAP:24,205).
 * Default 1.  Applies to every later b200vqa_iqap_* call on this handle. */
B200VQA_API int b200vqa_set_start_token(b200vqa_handle* h, int start_token);
/* Device bytes of activations the library will hold for a batch of B questions. */
B200VQA_API size_t b200vqa_workspace_bytes(const b200vqa_handle* h, int B);
B200VQA_API const char* b200vqa_last_error(void);
B200VQA_API const char* b200vqa_version(void);
/* Number of kernel launches issued by this handle since creation (bench.py's gpu_launches). */
B200VQA_API uint64_t b200vqa_launch_count(const b200vqa_handle* h);

/* Built-in per-kernel-class timer (CUDA events on the caller's stream around every launch). bench.py uses it
 * for the live roofline numbers; it serialises nothing but adds two event records per launch, so it is
 * only enabled for a few untimed steps.  ms_per_tag / launches_per_tag have b200vqa_profile_num_tags() slots. */
B200VQA_API int b200vqa_profile_num_tags(void);
B200VQA_API const char* b200vqa_profile_tag_name(int tag);
B200VQA_API int b200vqa_profile_begin(b200vqa_handle* h);
B200VQA_API int b200vqa_profile_end(b200vqa_handle* h, float* ms_per_tag, int32_t* launches_per_tag);
/* Enqueues a kernel that keeps `stream` busy for about `ms` milliseconds: called before a profiled step so the host
 * can enqueue the whole step behind it and the events see back-to-back kernels, not host launch latency. */
B200VQA_API int b200vqa_profile_delay(b200vqa_handle* h, double ms, void* stream);

/* ---------------------------------------------------------------------------------------------------- */
/* IQAP                                                                                                   */
/* ---------------------------------------------------------------------------------------------------- */

/* Replaces VQAModel.forward (IQAP:136-188) incl. autoregressive_program_generation (IQAP:190-241).
 *   image_features [B,196,1024] f32, questions [B,46] i64  ->  answer [B,classes] f32, programs [B,T] i64
 *   opt_step_logits  : NULL or [B,T,dec_vocab] f32, the program_output logits of every decode position
 *   opt_forced_tokens: NULL or [B,T] i64; when given, position t+1 is fed forced[b,t] instead of the argmax
 *                      (teacher forcing, used by the parity tests; `programs` still receives the argmax)
 *   opt_memory       : NULL or [243,B,256] f32, the encoder output in the reference's seq-first layout */
B200VQA_API int b200vqa_iqap_forward(b200vqa_handle* h, const float* image_features, const int64_t* questions, int B,
                         int program_len, float* answer, int64_t* programs, float* opt_step_logits,
                         const int64_t* opt_forced_tokens, float* opt_memory, void* stream);

/* Feature ingest with image-level de-duplication.  CLEVR has ~10 questions per image and the
 * question file carries `image_idxs` (reference preprocess_questions/preprocess_questions.py:122, read by
 * VQADatasetSingleSample, IQAP:63-72): image_features [n_img,196,1024] f32 holds each image ONCE, image_idx [B] i32
 * names the image of every question.  image_proj (+PE) runs once per image; results are identical to
 * b200vqa_iqap_forward on the expanded [B,196,1024] features.  Device pointers; asynchronous on `stream`. */
B200VQA_API int b200vqa_iqap_forward_indexed(b200vqa_handle* h, const float* image_features, int n_img,
                                             const int32_t* image_idx, const int64_t* questions, int B, int program_len,
                                             float* answer, int64_t* programs, void* stream);
/* Same with HOST buffers: unique images are uploaded once (double-buffered against image_proj), then the questions
 * run in chunks.  Returns once h_answer / h_programs are valid. */
B200VQA_API int b200vqa_iqap_forward_host_indexed(b200vqa_handle* h, const float* h_image_features, int n_img,
                                                  const int32_t* h_image_idx, const int64_t* h_questions, int B,
                                                  int program_len, float* h_answer, int64_t* h_programs, int chunk,
                                                  void* stream);

/* fp16 feature store: the same calls with image features held as IEEE fp16 - half the HBM / PCIe
 * bytes of the reference's float32 HDF5 layout (preprocess_images/extract_features.py:124).  image_proj then runs as
 * an fp16 x fp16 tensor-core GEMM with fp32 accumulation: fp16 carries the same 10-bit mantissa as the tf32 operands
 * of the fp32 path, so the results agree to rounding (conv4 features are post-ReLU, far inside fp16's range). */
B200VQA_API int b200vqa_iqap_forward_f16(b200vqa_handle* h, const void* image_features_f16, const int64_t* questions,
                                         int B, int program_len, float* answer, int64_t* programs,
                                         float* opt_step_logits, const int64_t* opt_forced_tokens, float* opt_memory,
                                         void* stream);
B200VQA_API int b200vqa_iqap_forward_host_f16(b200vqa_handle* h, const void* h_image_features_f16,
                                              const int64_t* h_questions, int B, int program_len, float* h_answer,
                                              int64_t* h_programs, int chunk, void* stream);

/* Evaluation tally on the device (reference inference_transformer_iqap_tally.py:317-344):
 * predicted answer = first maximum of answer_logits[b] (torch.max), program correct iff all program_len tokens
 * match.  counts[4] u64 (device) are ACCUMULATED: {both correct, answer correct / program wrong, answer wrong /
 * program correct, both wrong}.  opt_pred_answers NULL or [B] i32.  Device pointers; asynchronous on `stream`. */
B200VQA_API int b200vqa_iqap_tally(b200vqa_handle* h, const float* answer_logits, const int64_t* programs,
                                   const int64_t* gt_answers, const int64_t* gt_programs, int B, int program_len,
                                   uint64_t* counts, int32_t* opt_pred_answers, void* stream);

/* Replaces VQAModel.autoregressive_program_generation (IQAP:190-241) for a caller-supplied memory
 * [S,B,256] f32 (seq-first, S <= 256). */
B200VQA_API int b200vqa_iqap_decode(b200vqa_handle* h, const float* memory, int S, int B, int program_len, int64_t* programs,
                        float* opt_step_logits, const int64_t* opt_forced_tokens, void* stream);

/* Same as b200vqa_iqap_forward with HOST buffers (pinned or pageable): the host->device copy of the
 * inputs and the device->host copy of the results happen inside the call, chunked and overlapped with
 * compute. Returns once h_answer / h_programs are valid. This is the call bench.py times for "e2e". */
B200VQA_API int b200vqa_iqap_forward_host(b200vqa_handle* h, const float* h_image_features, const int64_t* h_questions, int B,
                              int program_len, float* h_answer, int64_t* h_programs, int chunk, void* stream);
/* Same without the final synchronisation: the (pinned) result buffers are valid once `stream` has drained.  With two
 * handles on two streams the upload of one batch overlaps the latency-bound decode tail of the previous one. */
B200VQA_API int b200vqa_iqap_forward_host_async(b200vqa_handle* h, const float* h_image_features,
                                                const int64_t* h_questions, int B, int program_len, float* h_answer,
                                                int64_t* h_programs, int chunk, void* stream);

/* ---------------------------------------------------------------------------------------------------- */
/* FA (step-wise executor with the inference cache)                                                       */
/* ---------------------------------------------------------------------------------------------------- */

/* Replaces the image half of greedy_decode (FA:129-131): view(B,1024,196).permute(0,2,1) -> image_proj.
 * Output: img_tokens [B,196,256] bf16 with the positional encoding of rows 0..195 already added; reused by
 * every program step of the question (the reference recomputes it per step). */
B200VQA_API int b200vqa_fa_project_images(b200vqa_handle* h, const float* image_features /*[B,1024,196]*/, int B,
                              void* img_tokens_bf16, void* stream);

/* Replaces greedy_decode (FA:126-146), batched: src [B,src_ld] i64 with src_len[b] valid tokens each
 * (the reference is batch 1 and never pads; here keys beyond 196+src_len[b] are masked).
 *   out_tokens [B,max_len] i64 (column 0 = start_token, like `ys`)
 *   opt_logits NULL or [B,max_len-1,vocab] f32; opt_forced NULL or [B,max_len-1] i64 (teacher forcing:
 *   this is also MultiModalTransformer.forward(image, src, tgt), FA:45-58, with tgt = [start | forced]) */
B200VQA_API int b200vqa_fa_step(b200vqa_handle* h, const void* img_tokens_bf16, const int64_t* src, const int32_t* src_len,
                    int src_ld, int B, int start_token, int max_len, int64_t* out_tokens, float* opt_logits,
                    const int64_t* opt_forced, void* stream);

/* Replaces MultiModalTransformer.forward (FA:45-58), the teacher-forced path: logits [B,T,vocab] f32 for
 * tgt [B,T] i64 (position t attends to tgt[:, :t+1]); src as in b200vqa_fa_step. */
B200VQA_API int b200vqa_fa_forward(b200vqa_handle* h, const void* img_tokens_bf16, const int64_t* src,
                                   const int32_t* src_len, int src_ld, int B, const int64_t* tgt, int T, float* logits,
                                   void* stream);

/* Replaces run_inference_chain (FA:83-124) for B questions at once with the cache resident in HBM.
 *   func [B,S] i32      function token of step i
 *   deps [B,S,2] i32    dependency pointers (step indices < i), -1 = none; a pointer to a step that has not
 *                       produced output contributes no tokens (the reference's `cache.get(idx, "")`, FA:110-115)
 *   n_steps [B] i32     steps of question b (<= S)
 *   cache [B,S,max_len] i32  OUT: all max_len tokens of every executed step incl. the start token (FA:120-121)
 *   h_active NULL or HOST int32[S]: h_active[i] = number of leading questions with n_steps > i (callers that
 *                       sort questions by n_steps descending let the library skip finished questions)
 *   opt_logits NULL or [B,S,max_len-1,vocab] f32; opt_forced NULL or [B,S,max_len-1] i64. With opt_forced the
 *   cache receives the FORCED tokens (later steps then consume exactly the tokens the caller dictated) and the
 *   model's own decisions are read from opt_logits - the parity tests feed the oracle's tokens this way. */
B200VQA_API int b200vqa_fa_run_chain(b200vqa_handle* h, const void* img_tokens_bf16, const int32_t* func, const int32_t* deps,
                         const int32_t* n_steps, int B, int S, int start_token, int max_len, int32_t* cache,
                         const int32_t* h_active, float* opt_logits, const int64_t* opt_forced, void* stream);

/* ---------------------------------------------------------------------------------------------------- */
/* LSTM program generator: question tokens -> program tokens                           */
/* ---------------------------------------------------------------------------------------------------- */
typedef struct b200vqa_lstm b200vqa_lstm;

/* fp32 device pointers in torch's nn.LSTM / nn.Embedding / nn.Linear state-dict layout (gate order i|f|g|o). */
typedef struct b200vqa_lstm_desc {
  int32_t vocab;         /* embedding rows (question vocabulary; program tokens index the same table) */
  int32_t embedding_dim; /* 256 */
  int32_t hidden_dim;    /* 512 */
  int32_t prog_vocab;    /* fc rows */
  const float* embedding; /* [vocab, 256] */
  const float* enc_w_ih;  /* encoder.weight_ih_l0 [2048, 256] */
  const float* enc_w_hh;  /* encoder.weight_hh_l0 [2048, 512] */
  const float* enc_b_ih;
  const float* enc_b_hh;
  const float* dec_w_ih;
  const float* dec_w_hh;
  const float* dec_b_ih;
  const float* dec_b_hh;
  const float* fc_w;      /* [prog_vocab, 512] */
  const float* fc_b;
} b200vqa_lstm_desc;

/* Replaces Seq2SeqModel.__init__ + load_state_dict (code/run_model_lstm_qp.py:277-289). */
B200VQA_API int b200vqa_lstm_create(const b200vqa_lstm_desc* desc, int device, b200vqa_lstm** out);
B200VQA_API void b200vqa_lstm_destroy(b200vqa_lstm* h);
B200VQA_API uint64_t b200vqa_lstm_launch_count(const b200vqa_lstm* h);
/* Replaces Seq2SeqModel.forward (code/run_model_lstm_qp.py:291-319): questions [B, q_len] i64 -> programs
 * [B, program_len] i64 (greedy).  opt_logits NULL or [B, program_len, prog_vocab] f32; opt_forced NULL or
 * [B, program_len] i64: decoder step t+1 is fed opt_forced[b, t] (teacher forcing = the `program_targets` branch
 * shifted by the start token).  Asynchronous on `stream`. */
B200VQA_API int b200vqa_lstm_generate(b200vqa_lstm* h, const int64_t* questions, int q_len, int B, int program_len,
                                      int start_token, int64_t* programs, float* opt_logits,
                                      const int64_t* opt_forced, void* stream);

/* Program -> chain glue (reference preprocess_questions/utils_programs.py:100-156, prefix_to_list): programs [B,T]
 * i64 in PREFIX order -> chain arrays in execution order (inputs before consumers, root last): func [B,S] i32 =
 * func_map[token], deps [B,S,2] i32 (-1 = none), n_steps [B] i32.  arity [prog_vocab] i32 gives each token's input
 * count (0 / 1 / 2; negative ends the program).  All device pointers; feeds b200vqa_fa_run_chain directly. */
B200VQA_API int b200vqa_programs_to_chain(const int64_t* programs, int B, int T, const int32_t* arity,
                                          const int32_t* func_map, int prog_vocab, int S, int32_t* func,
                                          int32_t* deps, int32_t* n_steps, void* stream);

/* Replaces the reference's driver loop around run_inference_chain (FA:193-206) and its per-step host round trips
 * (FA:109-121): HOST buffers in, HOST cache out.  h_img [B, 1024, 196] fp32 (the H5 layout, FA:201; pinned memory for
 * full PCIe speed), h_func [B,S], h_deps [B,S,2], h_n_steps [B], h_cache [B,S,max_len] (rows of steps that are not
 * executed are set to -1).  Questions are processed in sub-batches of `chunk` (<= 0: 2048): a sub-batch's features are
 * uploaded and projected on an internal stream while the previous sub-batch executes, each sub-batch runs
 * longest-program-first with the cache in HBM, and its cache is downloaded once.  Synchronises `stream` before
 * returning. */
B200VQA_API int b200vqa_fa_run_chain_host(b200vqa_handle* h, const float* h_img, const int32_t* h_func,
                                          const int32_t* h_deps, const int32_t* h_n_steps, int B, int S, int start_token,
                                          int max_len, int32_t* h_cache, int chunk, void* stream);
/* Same without the final synchronisation: h_cache is valid (and the host inputs may be released) once `stream` has been
 * synchronised by the caller.  Lets independent parts of a batch run on several handles / streams at once: a chain is a
 * dependent sequence of small kernels per program step, so two of them overlap well. */
B200VQA_API int b200vqa_fa_run_chain_host_async(b200vqa_handle* h, const float* h_img, const int32_t* h_func,
                                                const int32_t* h_deps, const int32_t* h_n_steps, int B, int S,
                                                int start_token, int max_len, int32_t* h_cache, int chunk, void* stream);

/* b200vqa_fa_run_chain with image-level de-duplication ("dedup of img_tokens (image_idxs)"):
 * img_tokens_bf16 [n_images,196,256] holds every image once, image_idx [B] i32 (device) names the image of each
 * question (reference preprocess_questions/preprocess_questions.py:122).  Everything else as above. */
B200VQA_API int b200vqa_fa_run_chain_indexed(b200vqa_handle* h, const void* img_tokens_bf16, int n_images,
                                             const int32_t* image_idx, const int32_t* func, const int32_t* deps,
                                             const int32_t* n_steps, int B, int S, int start_token, int max_len,
                                             int32_t* cache, const int32_t* h_active, float* opt_logits,
                                             const int64_t* opt_forced, void* stream);

/* ---------------------------------------------------------------------------------------------------- */
/* Kernel-level entry points used by the test-suite only (not part of the drop-in surface)               */
/* ---------------------------------------------------------------------------------------------------- */
typedef struct b200vqa_dbg_gemm_args {
  int32_t epilogue; /* 0 bias, 1 bias+relu, 2 bias+residual+layernorm, 3 bias+pe+row remap */
  int32_t tf32;     /* A/W are fp32 (tf32 math) instead of bf16 */
  int32_t block_n;  /* 128 or 256 */
  int32_t M, N, K;
  const void* A;    /* [M,K] */
  const void* W;    /* [N,K] */
  const float* bias;
  void* out;        /* bf16 */
  int32_t ldc;
  const void* residual; /* bf16 [M,N] */
  const float* gamma;
  const float* beta;
  float* out_f32;
  int32_t rows_in, rows_out, row_off, pe_off;
  const float* pe;
  long long* clk; /* optional device int64[16]: clock64() stamps of CTA 0's pipeline stages (tools/microbench_gemm.py) */
} b200vqa_dbg_gemm_args;
B200VQA_API int b200vqa_dbg_gemm(const b200vqa_dbg_gemm_args* args, void* stream);
B200VQA_API int b200vqa_dbg_gemm_check(int a_is_f32, const void* A, const void* W, const float* bias, float* out, int M, int N,
                           int K, void* stream);
/* qkv [B*256, 768] bf16 -> out [B*256, 256] bf16; lens NULL -> const_len */
B200VQA_API int b200vqa_dbg_enc_attention(const void* qkv, const int32_t* lens, int const_len, int B, int nhead, int v_mode,
                              void* out, void* stream);
/* absorbed decode cross-attention of one position: qp [B, nhead*256] bf16 (absorbed queries), memory [B*256, 256] bf16
 * -> out [B, nhead*256] bf16 (per-head attention-weighted memory); lens NULL -> const_len.
 * impl must be 0 (the warp-MMA ring kernel); stamps must be NULL (kept for ABI stability). */
B200VQA_API int b200vqa_dbg_mem_attn(const void* qp, const void* memory, const int32_t* lens, int const_len, int B, int nhead,
                         int impl, void* out, long long* stamps, void* stream);
/* Test hook: copies one decode scratch buffer of the handle's workspace (sized by an earlier call) into the device
 * buffer dst (at most dst_bytes) after a device synchronisation; *bytes = the buffer's size.  which: 0 dx, 1 dqkv, 2 dattn, 3 dx1, 4 dq, 5 du, 6 dx2, 7 dxo[0], 8 dxo[1] (bf16), 9 dout / pre-LN (fp32),
 * 10 tok (int64 [cap, 65]).  Used by tests / tools to compare the persistent decode kernel with the per-kernel chain. */
B200VQA_API int b200vqa_dbg_workspace(b200vqa_handle* h, int which, void* dst, size_t dst_bytes, size_t* bytes);

/* Upload mode of b200vqa_iqap_forward_host[_async] (and b200vqa_fa_run_chain_host[_async]: bf16, see
 * b200vqa_host_f32_to_bf16) for fp32 host features: 0 (default) = the bytes as given, 1 = rounded
 * to fp16 (nearest even) on host threads, chunk by chunk, while the previous chunk is on the wire - half the PCIe bytes,
 * then the device path of a caller-provided fp16 feature store (b200vqa_iqap_forward_host_f16): results equal those of
 * that call on the rounded features bit for bit.  Pays when the process has host cores to spare (>= ~8). */
B200VQA_API int b200vqa_set_host_upload(b200vqa_handle* h, int mode);

/* Host helper (no GPU involved): n fp32 values -> fp16, round to nearest even (what torch's .half() gives), on `threads`
 * host threads (<= 0: every CPU this process may run on, at most 32).  dst must be 32-byte aligned.  The *_host entry
 * points use it for the "fp16" upload mode (b200vqa_set_host_upload): the feature rows that the reference uploads as
 * fp32 per sample (IQAP:288-296) cross PCIe at half the size. */
B200VQA_API int b200vqa_host_f32_to_f16(const float* src, void* dst, long long n, int threads);
/* The same to bf16 (round to nearest even = the device's conversion): upload mode 1 of b200vqa_fa_run_chain_host[_async],
 * whose device path rounds the features to bf16 anyway (FA:47 transpose + cast) - the results are bit-identical to mode 0. */
B200VQA_API int b200vqa_host_f32_to_bf16(const float* src, void* dst, long long n, int threads);

#ifdef __cplusplus
}
#endif
#endif /* B200VQA_H_ */
