"""Drop-in for the reference's `code/inference_transformer_iqap.py` ("IQAP") on B200.

Same public surface - `Config`, `PositionalEncoding`, `generate_square_subsequent_mask`, `VQAModel`,
`get_data_info`, `load_model`, `run_inference` - and the same state-dict (parameter / buffer names and
shapes), so `model.load_state_dict(torch.load("best_transformer_iqap.pth"))` works
unchanged.  What differs is underneath: `VQAModel.forward` and `autoregressive_program_generation` do not
run PyTorch modules; they hand raw device pointers to libb200vqa.so (include/b200vqa.h), whose sm_100a
kernels compute the encoder, the answer head and the whole 27-position greedy decode on the GPU with a
KV cache, the cross-attention K/V projected once, and the argmax / token append on device.

The nn.Module children (`nn.TransformerEncoder`, ...) are kept as *parameter containers* only: they give the
state-dict its names and PyTorch's default initialisation.  There is no CPU / eager fallback - calling the
model with CPU tensors raises.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np
import torch
import torch.nn as nn

from . import _native as nat

__all__ = ["Config", "PositionalEncoding", "generate_square_subsequent_mask", "VQAModel", "get_data_info",
           "load_model", "run_inference"]


class Config:
    """Mirrors the reference's mutable configuration class (IQAP:9-24); paths come from the environment."""
    PATH = os.environ.get("VQA_CODE_PATH", "./")
    FEATURES_H5 = os.environ.get("VQA_FEATURES_H5", PATH + "data/train_features.h5")
    QUESTIONS_H5 = os.environ.get("VQA_QUESTIONS_H5", PATH + "h5_files/train_questions.h5")
    MODELS_DIR = PATH + "models"
    MODEL_NAME = os.environ.get("VQA_IQAP_MODEL", PATH + "models/best_transformer_iqap.pth")
    EMBEDDING_DIM = 256
    HIDDEN_DIM = 256
    IMAGE_FEATURE_DIM = 1024
    NUM_CLASSES = None          # filled by load_model
    PROGRAM_SEQ_LEN = 27
    PROGRAM_VOCAB_SIZE = None   # filled by load_model
    MAX_QUESTION_LEN = 46
    NUM_IMAGE_TOKENS = 14 * 14
    SPECIAL_TOKEN_ID = 1        # <START>


def _sinusoid_table(max_len: int, d_model: int) -> torch.Tensor:
    """pe[p, 2i] = sin(p * w_i), pe[p, 2i+1] = cos(p * w_i), w_i = exp(-2i ln(1e4)/d)   (IQAP:30-36)."""
    pos = torch.arange(0, max_len).unsqueeze(1)
    freq = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
    table = torch.zeros(max_len, d_model)
    table[:, 0::2] = torch.sin(pos * freq)
    table[:, 1::2] = torch.cos(pos * freq)
    return table


class PositionalEncoding(nn.Module):
    """Seq-first sinusoidal table kept as the persistent buffer `pe` of shape (max_len, 1, d) (IQAP:27-45).

    `forward` exists for callers that use the module on its own; the fused kernels read `pe` directly.
    """

    def __init__(self, d_model, dropout=0.1, max_len=5000):
        super().__init__()
        self.register_buffer("pe", _sinusoid_table(max_len, d_model).unsqueeze(1))
        self.dropout = nn.Dropout(p=dropout)

    def forward(self, x):
        return self.dropout(x + self.pe[: x.size(0)])


def generate_square_subsequent_mask(sz):
    """Float causal mask, 0 on/below the diagonal and -inf above (IQAP:48-51).  The decode kernels use a KV
    cache instead of a mask; this stays for API compatibility."""
    keep = torch.ones(sz, sz).tril().bool()
    return torch.zeros(sz, sz).masked_fill(~keep, float("-inf"))


class VQAModel(nn.Module):
    """image features + question -> (answer logits, 27 program tokens); reference IQAP:95-241."""

    def __init__(self, vocab_size, embedding_dim, hidden_dim, num_classes, program_vocab_size, program_seq_len,
                 num_image_tokens, special_token_id=1):
        super().__init__()
        self.embedding_dim = embedding_dim
        self.hidden_dim = hidden_dim
        self.num_image_tokens = num_image_tokens
        self.special_token_id = special_token_id
        # construction order == the reference's, so a seeded default init yields identical parameters
        self.image_proj = nn.Linear(Config.IMAGE_FEATURE_DIM, embedding_dim)
        self.embedding = nn.Embedding(vocab_size, embedding_dim, padding_idx=0)
        self.cls_token = nn.Parameter(torch.randn(1, 1, embedding_dim))
        self.pos_encoder = PositionalEncoding(embedding_dim, dropout=0.1,
                                              max_len=num_image_tokens + Config.MAX_QUESTION_LEN + 1)
        enc_layer = nn.TransformerEncoderLayer(d_model=embedding_dim, nhead=4)
        self.transformer_encoder = nn.TransformerEncoder(enc_layer, num_layers=1)
        self.answer_classifier = nn.Sequential(nn.Linear(embedding_dim, hidden_dim), nn.ReLU(), nn.Dropout(0.1),
                                               nn.Linear(hidden_dim, num_classes))
        self.program_decoder_embedding = nn.Embedding(program_vocab_size, embedding_dim, padding_idx=0)
        self.pos_decoder = PositionalEncoding(embedding_dim, dropout=0.1, max_len=Config.PROGRAM_SEQ_LEN + 1)
        dec_layer = nn.TransformerDecoderLayer(d_model=embedding_dim, nhead=4)
        self.transformer_decoder = nn.TransformerDecoder(dec_layer, num_layers=2)
        self.program_output = nn.Linear(embedding_dim, program_vocab_size)
        self._nhead = 4
        self._pool = nat.HandlePool(self, self._build_desc)
        self._next_slot = 0
        self._host_inflight = []  # host tensors of submit_host calls that have not been drained yet

    # ------------------------------------------------------------------ native handle
    def _build_desc(self):
        keep = []
        enc_layers = list(self.transformer_encoder.layers)
        dec_layers = list(self.transformer_decoder.layers)
        for i, l in enumerate(enc_layers):
            nat.check_layer_contract(l, f"transformer_encoder.layers.{i}")
        for i, l in enumerate(dec_layers):
            nat.check_layer_contract(l, f"transformer_decoder.layers.{i}")
        enc_arr = (nat.EncoderLayerWeights * len(enc_layers))(*[nat.encoder_layer_weights(l, keep) for l in enc_layers])
        dec_arr = (nat.DecoderLayerWeights * len(dec_layers))(*[nat.decoder_layer_weights(l, keep) for l in dec_layers])
        keep += [enc_arr, dec_arr]
        d = nat.ModelDesc()
        d.kind = nat.MODEL_IQAP
        d.d_model = self.embedding_dim
        d.img_feat_dim = self.image_proj.in_features
        d.n_img_tokens = self.num_image_tokens
        d.nhead = enc_layers[0].self_attn.num_heads
        d.n_enc_layers = len(enc_layers)
        d.n_dec_layers = len(dec_layers)
        d.dim_ff = enc_layers[0].linear1.out_features
        d.enc_vocab = self.embedding.num_embeddings
        d.dec_vocab = self.program_output.out_features
        d.pe_enc_len = self.pos_encoder.pe.shape[0]
        d.pe_dec_len = self.pos_decoder.pe.shape[0]
        d.max_q_len = d.pe_enc_len - 1 - self.num_image_tokens
        d.answer_hidden = self.answer_classifier[0].out_features
        d.num_classes = self.answer_classifier[3].out_features
        d.layer_norm_eps = enc_layers[0].norm1.eps
        nat._set(d, keep,
                 image_proj_weight=self.image_proj.weight, image_proj_bias=self.image_proj.bias,
                 cls_token=self.cls_token.reshape(-1),
                 enc_embedding=self.embedding.weight, dec_embedding=self.program_decoder_embedding.weight,
                 pe_enc=self.pos_encoder.pe.reshape(d.pe_enc_len, -1),
                 pe_dec=self.pos_decoder.pe.reshape(d.pe_dec_len, -1),
                 head_weight=self.program_output.weight, head_bias=self.program_output.bias,
                 answer_w0=self.answer_classifier[0].weight, answer_b0=self.answer_classifier[0].bias,
                 answer_w1=self.answer_classifier[3].weight, answer_b1=self.answer_classifier[3].bias)
        d.enc_layers = enc_arr
        d.dec_layers = dec_arr
        return d, keep

    def _native(self, slot: int = 0) -> nat.Handle:
        h = self._pool.get(slot)
        h.set_start_token(Config.SPECIAL_TOKEN_ID)  # read at decode time like the reference (IQAP:205)
        return h

    @staticmethod
    def _check_input(t, what, dtype):
        if not t.is_cuda:
            raise nat.NativeError(f"{what} is on {t.device}; the B200 executor has no CPU path")
        return t.to(dtype).contiguous()

    # ------------------------------------------------------------------ reference surface
    def forward(self, image_features, questions, program_targets=None, max_program_length=None):
        """image_features (B, 196, 1024) f32, questions (B, 46) i64 -> (answer_output (B, C) f32,
        generated program tokens (B, 27) i64).  Like the reference (IQAP:181-188) the second result holds
        token ids in both the `program_targets is None` and the training-style call."""
        answer, programs, _, _ = self.forward_detailed(image_features, questions)
        return answer, programs

    @torch.no_grad()
    def forward_detailed(self, image_features, questions, forced_programs=None, want_logits=False,
                         want_memory=False, slot=0):
        """Extended entry used by the parity tests: optionally teacher-forces the decoder with
        `forced_programs` (B, 27) and returns the per-position program logits (B, 27, Vp) and the encoder
        memory (S, B, d) next to the reference's two outputs."""
        h = self._native(slot)
        # fp16 features (the half-size feature store) are consumed as they are; everything else is the reference's fp32
        f16 = image_features.dtype == torch.float16
        img = self._check_input(image_features, "image_features", torch.float16 if f16 else torch.float32)
        q = self._check_input(questions, "questions", torch.int64)
        B = img.shape[0]
        T = Config.PROGRAM_SEQ_LEN
        n_img, feat = self.num_image_tokens, self.image_proj.in_features
        if tuple(img.shape) != (B, n_img, feat):
            raise ValueError(f"image_features must be (B, {n_img}, {feat}), got {tuple(img.shape)}")
        q_len = self.pos_encoder.pe.shape[0] - 1 - n_img
        if tuple(q.shape) != (B, q_len):
            raise ValueError(f"questions must be (B, {q_len}), got {tuple(q.shape)}")
        dev = img.device
        C_, Vp, S = self.answer_classifier[3].out_features, self.program_output.out_features, 1 + n_img + q_len
        answer = torch.empty(B, C_, dtype=torch.float32, device=dev)
        programs = torch.empty(B, T, dtype=torch.int64, device=dev)
        logits = torch.empty(B, T, Vp, dtype=torch.float32, device=dev) if want_logits else None
        memory = torch.empty(S, B, self.embedding_dim, dtype=torch.float32, device=dev) if want_memory else None
        forced = None
        if forced_programs is not None:
            forced = self._check_input(forced_programs, "forced_programs", torch.int64)
            if tuple(forced.shape) != (B, T):
                raise ValueError(f"forced_programs must be (B, {T})")
        with torch.cuda.device(dev):
            fn = nat.lib().b200vqa_iqap_forward_f16 if f16 else nat.lib().b200vqa_iqap_forward
            nat.check(fn(h.raw, nat.ptr(img), nat.ptr(q), B, T, nat.ptr(answer), nat.ptr(programs), nat.ptr(logits),
                         nat.ptr(forced), nat.ptr(memory), nat.stream_ptr(dev)), "b200vqa_iqap_forward")
        return answer, programs, logits, memory

    @torch.no_grad()
    def autoregressive_program_generation(self, memory, program_seq_len):
        """memory (S, B, d) f32 -> greedy program tokens (B, program_seq_len) i64   (IQAP:190-241)."""
        h = self._native()
        mem = self._check_input(memory, "memory", torch.float32)
        S, B, d = mem.shape
        programs = torch.empty(B, program_seq_len, dtype=torch.int64, device=mem.device)
        with torch.cuda.device(mem.device):
            nat.check(nat.lib().b200vqa_iqap_decode(h.raw, nat.ptr(mem), S, B, int(program_seq_len), nat.ptr(programs),
                                                    None, None, nat.stream_ptr(mem.device)), "b200vqa_iqap_decode")
        return programs

    def _check_host_inputs(self, img, q):
        """The library reads exactly B x num_image_tokens x in_features features and B x q_len question tokens."""
        if img.is_cuda or q.is_cuda:
            raise ValueError("the host-buffer entry points take CPU tensors; use forward() for device tensors")
        q_len = self.pos_encoder.pe.shape[0] - 1 - self.num_image_tokens
        if img.dim() != 3 or tuple(img.shape[1:]) != (self.num_image_tokens, self.image_proj.in_features):
            raise ValueError(f"image_features must be (B, {self.num_image_tokens}, {self.image_proj.in_features}), "
                             f"got {tuple(img.shape)}")
        if tuple(q.shape) != (img.shape[0], q_len):
            raise ValueError(f"questions must be ({img.shape[0]}, {q_len}), got {tuple(q.shape)}")

    @staticmethod
    def resolve_upload(upload):
        """"fp32": the feature bytes cross PCIe as given (results bit-identical to `forward`).  "fp16": the library
        rounds them to fp16 on host threads while the previous chunk is on the wire - half the bytes (the upload, not
        the GPU, bounds the host-buffer calls: 803 KB per question), results bit-identical to passing `features.half()`.
        "auto": "fp16" when this process has at least 8 CPUs to itself (CPUs it may run on / ranks on the node)."""
        if upload == "auto":
            cpus = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
            ranks = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1"))))
            return "fp16" if cpus // ranks >= 8 else "fp32"
        if upload not in ("fp32", "fp16"):
            raise ValueError('upload must be "fp32", "fp16" or "auto"')
        return upload

    @torch.no_grad()
    def forward_host(self, image_features_cpu, questions_cpu, chunk=512, upload="fp32"):
        """End-to-end call with HOST tensors (pinned for full PCIe speed): upload, compute and download are
        pipelined inside the library; returns CPU tensors.  `upload`: see `resolve_upload`.  This is what bench.py
        times as `e2e`."""
        h = self._native()
        h.set_host_upload(self.resolve_upload(upload) == "fp16")
        f16 = image_features_cpu.dtype == torch.float16   # fp16 feature store: half the PCIe bytes
        img = image_features_cpu.to(torch.float16 if f16 else torch.float32).contiguous()
        q = questions_cpu.to(torch.int64).contiguous()
        self._check_host_inputs(img, q)
        B, T = img.shape[0], Config.PROGRAM_SEQ_LEN
        answer = torch.empty(B, self.answer_classifier[3].out_features, dtype=torch.float32).pin_memory()
        programs = torch.empty(B, T, dtype=torch.int64).pin_memory()
        dev = self.image_proj.weight.device
        with torch.cuda.device(dev):
            fn = nat.lib().b200vqa_iqap_forward_host_f16 if f16 else nat.lib().b200vqa_iqap_forward_host
            nat.check(fn(h.raw, nat.ptr(img), nat.ptr(q), B, T, nat.ptr(answer), nat.ptr(programs), int(chunk),
                         nat.stream_ptr(dev)), "b200vqa_iqap_forward_host")
        return answer, programs

    # ------------------------------------------------------------------ feature ingest with image de-duplication
    @torch.no_grad()
    def forward_indexed(self, image_features, image_idx, questions):
        """Several questions per image (CLEVR: ~10): image_features (n_img, 196, 1024) f32 holds every image once,
        image_idx (B,) names each question's image - the `image_idxs` dataset of the question file
        (VQADatasetSingleSample, IQAP:63-72).  image_proj runs once per image; the results equal
        `forward(image_features[image_idx], questions)`."""
        h = self._native()
        img = self._check_input(image_features, "image_features", torch.float32)
        q = self._check_input(questions, "questions", torch.int64)
        idx = self._check_input(image_idx.to(torch.int32), "image_idx", torch.int32)
        B, n_img, T = q.shape[0], img.shape[0], Config.PROGRAM_SEQ_LEN
        if tuple(img.shape[1:]) != (self.num_image_tokens, self.image_proj.in_features):
            raise ValueError(f"image_features must be (n_img, {self.num_image_tokens}, {self.image_proj.in_features})")
        if idx.shape != (B,):
            raise ValueError("image_idx must be (B,)")
        if B and (int(idx.min()) < 0 or int(idx.max()) >= n_img):
            raise IndexError("image_idx out of range")
        answer = torch.empty(B, self.answer_classifier[3].out_features, dtype=torch.float32, device=img.device)
        programs = torch.empty(B, T, dtype=torch.int64, device=img.device)
        with torch.cuda.device(img.device):
            nat.check(nat.lib().b200vqa_iqap_forward_indexed(h.raw, nat.ptr(img), n_img, nat.ptr(idx), nat.ptr(q), B, T,
                                                             nat.ptr(answer), nat.ptr(programs),
                                                             nat.stream_ptr(img.device)), "b200vqa_iqap_forward_indexed")
        return answer, programs

    @torch.no_grad()
    def forward_host_indexed(self, image_features_cpu, image_idx_cpu, questions_cpu, chunk=512):
        """`forward_indexed` with HOST tensors: each unique image crosses PCIe once."""
        h = self._native()
        img = image_features_cpu.to(torch.float32).contiguous()
        q = questions_cpu.to(torch.int64).contiguous()
        idx = image_idx_cpu.to(torch.int32).contiguous()
        if img.is_cuda or q.is_cuda or idx.is_cuda:
            raise ValueError("forward_host_indexed takes CPU tensors")
        B, n_img, T = q.shape[0], img.shape[0], Config.PROGRAM_SEQ_LEN
        if B and (int(idx.min()) < 0 or int(idx.max()) >= n_img):
            raise IndexError("image_idx out of range")
        answer = torch.empty(B, self.answer_classifier[3].out_features, dtype=torch.float32).pin_memory()
        programs = torch.empty(B, T, dtype=torch.int64).pin_memory()
        dev = self.image_proj.weight.device
        with torch.cuda.device(dev):
            nat.check(nat.lib().b200vqa_iqap_forward_host_indexed(h.raw, nat.ptr(img), n_img, nat.ptr(idx), nat.ptr(q),
                                                                  B, T, nat.ptr(answer), nat.ptr(programs), int(chunk),
                                                                  nat.stream_ptr(dev)),
                      "b200vqa_iqap_forward_host_indexed")
        return answer, programs

    # ------------------------------------------------------------------ evaluation tally on the device
    @torch.no_grad()
    def tally(self, answer_output, programs, gt_answers, gt_programs, counts=None):
        """The four-way tally of inference_transformer_iqap_tally.run_inference (TALLY:317-344) without a host loop:
        returns `counts` (4,) int64 on the device = {both correct, answer only, program only, neither}, accumulated
        into the tensor passed in, and the predicted answers (B,) int32."""
        h = self._native()
        a = self._check_input(answer_output, "answer_output", torch.float32)
        p = self._check_input(programs, "programs", torch.int64)
        ga = self._check_input(gt_answers.to(torch.int64), "gt_answers", torch.int64)
        gp = self._check_input(gt_programs.to(torch.int64), "gt_programs", torch.int64)
        B, T = p.shape
        if a.shape[0] != B or ga.shape != (B,) or tuple(gp.shape) != (B, T):
            raise ValueError("tally: shapes disagree")
        if counts is None:
            counts = torch.zeros(4, dtype=torch.int64, device=a.device)
        pred = torch.empty(B, dtype=torch.int32, device=a.device)
        with torch.cuda.device(a.device):
            nat.check(nat.lib().b200vqa_iqap_tally(h.raw, nat.ptr(a), nat.ptr(p), nat.ptr(ga), nat.ptr(gp), B, T,
                                                   nat.ptr(counts), nat.ptr(pred), nat.stream_ptr(a.device)),
                      "b200vqa_iqap_tally")
        return counts, pred

    def native_launch_count(self) -> int:
        return self._pool.launch_count()

    # ------------------------------------------------------------------ pipelined submission (no reference equivalent)
    @torch.no_grad()
    def submit(self, image_features, questions, depth=2):
        """Asynchronous forward of an independent batch on one of `depth` internal (handle, stream) slots, taken
        round-robin.  Returns (answer_output, programs) whose contents are valid only after `drain()` (or after the
        caller's stream waited for the slot stream).  The slot stream first waits for the caller's stream, so inputs
        produced there are complete before they are read."""
        slot = 1 + self._next_slot % depth
        self._next_slot += 1
        st = self._pool.stream(slot)
        self.last_submit_stream = st
        st.wait_stream(torch.cuda.current_stream(st.device))
        with torch.cuda.stream(st):
            answer, programs, _, _ = self.forward_detailed(image_features, questions, slot=slot)
        # inputs are read, and outputs written, on the slot stream: the caching allocator must not recycle their memory
        # for the caller's stream before that work has finished
        for t in (image_features, questions):
            if t.is_cuda:
                t.record_stream(st)
        cur = torch.cuda.current_stream(st.device)
        answer.record_stream(cur)
        programs.record_stream(cur)
        return answer, programs

    @torch.no_grad()
    def submit_host(self, image_features_cpu, questions_cpu, chunk=512, depth=2, upload="fp32", background=None):
        """`forward_host` without the final synchronisation, on a round-robin slot: the upload of this batch overlaps
        the decode tail of the previous one.  Returns pinned CPU tensors that are valid after `drain_host()`.
        `background` (default: on with the fp16 upload mode): the library call runs on the slot's own host thread, so the
        host-side rounding of this batch overlaps the enqueue work of the previous one and the caller is not held up."""
        slot = 1 + self._next_slot % depth
        self._next_slot += 1
        h = self._native(slot)
        fp16 = self.resolve_upload(upload) == "fp16"
        st = self._pool.stream(slot)
        img = image_features_cpu.to(torch.float32).contiguous()
        q = questions_cpu.to(torch.int64).contiguous()
        self._check_host_inputs(img, q)
        B, T = img.shape[0], Config.PROGRAM_SEQ_LEN
        answer = torch.empty(B, self.answer_classifier[3].out_features, dtype=torch.float32).pin_memory()
        programs = torch.empty(B, T, dtype=torch.int64).pin_memory()

        def call():
            with torch.cuda.device(st.device):
                h.set_host_upload(fp16)
                nat.check(nat.lib().b200vqa_iqap_forward_host_async(h.raw, nat.ptr(img), nat.ptr(q), B, T, nat.ptr(answer),
                                                                    nat.ptr(programs), int(chunk),
                                                                    C.c_void_p(st.cuda_stream)),
                          "b200vqa_iqap_forward_host_async")

        if background is None:
            background = fp16
        if background:
            self._pool.run_in_background(slot, call)
        else:
            call()
        # the asynchronous upload reads `img` / `q` (possibly temporaries made by .to() / .contiguous() above) and the
        # download writes `answer` / `programs` until the slot stream has drained: keep all four alive until drain_host()
        self._host_inflight.append((img, q, answer, programs))
        return answer, programs

    def drain(self):
        """Results of every `submit` become valid on the caller's current stream."""
        self._pool.drain()

    def drain_host(self):
        """Blocks the host until every `submit_host` has delivered its results."""
        self._pool.wait_background()  # re-raises a NativeError of a background call
        for st in list(self._pool._streams.values()):
            st.synchronize()
        self._host_inflight.clear()


def get_data_info(questions_h5_path):
    """(question vocab, answer classes, program vocab) = max id + 1 over the H5 arrays (IQAP:244-252)."""
    import h5py  # optional dependency: only the dataset helpers need it
    with h5py.File(questions_h5_path, "r") as f:
        return (int(np.max(f["questions"])) + 1, int(np.max(f["answers"])) + 1, int(np.max(f["programs"])) + 1)


def load_model(device):
    """Builds VQAModel from the H5 metadata and loads Config.MODEL_NAME (IQAP:255-276)."""
    vocab_size, num_classes, program_vocab_size = get_data_info(Config.QUESTIONS_H5)
    Config.NUM_CLASSES = num_classes
    Config.PROGRAM_VOCAB_SIZE = program_vocab_size
    model = VQAModel(vocab_size, Config.EMBEDDING_DIM, Config.HIDDEN_DIM, num_classes, program_vocab_size,
                     Config.PROGRAM_SEQ_LEN, Config.NUM_IMAGE_TOKENS, Config.SPECIAL_TOKEN_ID).to(device)
    model.load_state_dict(torch.load(Config.MODEL_NAME, map_location=device))
    model.eval()
    return model


def run_inference(idx=0, batch_size=256):
    """Runs samples [0, idx) like the reference driver (IQAP:279-319) but batched: features are read in
    (N,1024,14,14) layout, transposed to (N,196,1024) and pushed through the host-buffer entry point."""
    import h5py
    device = torch.device("cuda")
    model = load_model(device)
    results = []
    with h5py.File(Config.FEATURES_H5, "r") as ff, h5py.File(Config.QUESTIONS_H5, "r") as qf:
        for b0 in range(0, idx, batch_size):
            ids = list(range(b0, min(idx, b0 + batch_size)))
            img_ids = [int(i) for i in qf["image_idxs"][ids]]
            feats = torch.from_numpy(np.stack([ff["features"][i] for i in img_ids])).float()
            feats = feats.permute(0, 2, 3, 1).reshape(len(ids), -1, Config.IMAGE_FEATURE_DIM)
            questions = torch.from_numpy(np.asarray(qf["questions"][ids])).long()
            answer_output, programs = model.forward_host(feats, questions)
            for j, sample_idx in enumerate(ids):
                pred = int(answer_output[j].argmax())
                results.append((sample_idx, pred, programs[j].tolist()))
                print(f"sample {sample_idx}: predicted answer {pred} | ground truth {int(qf['answers'][sample_idx])}")
                print(f"  predicted program {programs[j].tolist()}")
                print(f"  ground truth      {qf['programs'][sample_idx].tolist()}")
    return results


if __name__ == "__main__":
    run_inference(idx=6)
