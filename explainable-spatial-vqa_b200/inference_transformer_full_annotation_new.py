"""Drop-in for the reference's `code/inference_transformer_full_annotation_new.py` ("FA") on B200:
the step-wise Program Executor with its inference cache.

Kept surface: `PositionalEncoding`, `MultiModalTransformer` (same constructor, attributes and state-dict,
SURVEY §8 a-5), `load_vocab`, `decode_tokens`, `tokenize_field`, `greedy_decode`, `run_inference_chain`,
`main_inference`.  New, batched entry points the reference never had: `project_images`,
`run_inference_chain_batched`, `chain_to_arrays`.

Underneath, libb200vqa.so (include/b200vqa.h) runs every program step on the GPU: the image tokens are
projected once per question (the reference redoes it every step, FA:130-131), the encoder input of step i
is gathered on the device from the HBM-resident cache through the step's dependency pointers, the 19-token
greedy decode runs with a KV cache and on-device argmax, and its 20 tokens land in `cache[b, i, :]` - no
host round trip between program steps (the reference does one H2D and one D2H per step, FA:118-120).
"""
from __future__ import annotations

import ctypes as C
import os
import json
import logging
import re

import numpy as np
import torch
import torch.nn as nn

from . import _native as nat

__all__ = ["PositionalEncoding", "MultiModalTransformer", "load_vocab", "decode_tokens", "tokenize_field",
           "run_inference_chain", "greedy_decode", "main_inference", "project_images", "chain_to_arrays",
           "run_inference_chain_batched"]

logger = logging.getLogger(__name__)
MAX_DEPS = 2  # CLEVR functions take at most two inputs


class PositionalEncoding(nn.Module):
    """Batch-first sinusoidal table, persistent buffer `pe` of shape (1, max_len, d)   (FA:14-27)."""

    def __init__(self, d_model, dropout=0.1, max_len=5000):
        super().__init__()
        self.dropout = nn.Dropout(p=dropout)
        pos = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
        freq = torch.exp(torch.arange(0, d_model, 2).float() * (-np.log(10000.0) / d_model))
        table = torch.zeros(max_len, d_model)
        table[:, 0::2] = torch.sin(pos * freq)
        table[:, 1::2] = torch.cos(pos * freq)
        self.register_buffer("pe", table.unsqueeze(0))

    def forward(self, x):
        return self.dropout(x + self.pe[:, : x.size(1)])


class MultiModalTransformer(nn.Module):
    """[196 image tokens | src tokens] -> encoder -> decoder -> vocabulary logits   (FA:32-58)."""

    def __init__(self, vocab_size, d_model=256, nhead=8, num_encoder_layers=3, num_decoder_layers=3,
                 dim_feedforward=512, dropout=0.1, max_text_len=50, max_img_tokens=196):
        super().__init__()
        self.d_model = d_model
        self.vocab_size = vocab_size
        self.max_img_tokens = max_img_tokens
        self.image_proj = nn.Linear(1024, d_model)
        self.text_embedding = nn.Embedding(vocab_size, d_model)
        self.pos_encoder = PositionalEncoding(d_model, dropout, max_len=max_text_len + max_img_tokens)
        self.pos_decoder = PositionalEncoding(d_model, dropout, max_len=max_text_len)
        # parameter container: gives the state-dict the reference's `transformer.encoder.layers.N...` names
        self.transformer = nn.Transformer(d_model, nhead, num_encoder_layers, num_decoder_layers, dim_feedforward,
                                          dropout, batch_first=True)
        self.output_linear = nn.Linear(d_model, vocab_size)
        self._pool = nat.HandlePool(self, self._build_desc)

    # ------------------------------------------------------------------ native handle
    def _build_desc(self):
        keep = []
        enc_layers = list(self.transformer.encoder.layers)
        dec_layers = list(self.transformer.decoder.layers)
        for i, l in enumerate(enc_layers):
            nat.check_layer_contract(l, f"transformer.encoder.layers.{i}")
        for i, l in enumerate(dec_layers):
            nat.check_layer_contract(l, f"transformer.decoder.layers.{i}")
        enc_arr = (nat.EncoderLayerWeights * len(enc_layers))(*[nat.encoder_layer_weights(l, keep) for l in enc_layers])
        dec_arr = (nat.DecoderLayerWeights * len(dec_layers))(*[nat.decoder_layer_weights(l, keep) for l in dec_layers])
        keep += [enc_arr, dec_arr]
        d = nat.ModelDesc()
        d.kind = nat.MODEL_FA
        d.d_model = self.d_model
        d.img_feat_dim = self.image_proj.in_features
        d.n_img_tokens = self.max_img_tokens
        d.nhead = enc_layers[0].self_attn.num_heads
        d.n_enc_layers = len(enc_layers)
        d.n_dec_layers = len(dec_layers)
        d.dim_ff = enc_layers[0].linear1.out_features
        d.enc_vocab = self.vocab_size
        d.dec_vocab = self.vocab_size
        d.pe_enc_len = self.pos_encoder.pe.shape[1]
        d.pe_dec_len = self.pos_decoder.pe.shape[1]
        d.max_q_len = d.pe_enc_len - self.max_img_tokens
        d.answer_hidden = 0
        d.num_classes = 0
        d.layer_norm_eps = enc_layers[0].norm1.eps
        enc_norm, dec_norm = self.transformer.encoder.norm, self.transformer.decoder.norm
        nat._set(d, keep,
                 image_proj_weight=self.image_proj.weight, image_proj_bias=self.image_proj.bias, cls_token=None,
                 enc_embedding=self.text_embedding.weight, dec_embedding=self.text_embedding.weight,
                 pe_enc=self.pos_encoder.pe.reshape(d.pe_enc_len, -1), pe_dec=self.pos_decoder.pe.reshape(d.pe_dec_len, -1),
                 enc_final_norm_weight=None if enc_norm is None else enc_norm.weight,
                 enc_final_norm_bias=None if enc_norm is None else enc_norm.bias,
                 dec_final_norm_weight=None if dec_norm is None else dec_norm.weight,
                 dec_final_norm_bias=None if dec_norm is None else dec_norm.bias,
                 head_weight=self.output_linear.weight, head_bias=self.output_linear.bias,
                 answer_w0=None, answer_b0=None, answer_w1=None, answer_b1=None)
        d.enc_layers = enc_arr
        d.dec_layers = dec_arr
        return d, keep

    def _native(self, slot: int = 0) -> nat.Handle:
        return self._pool.get(slot)

    def native_launch_count(self) -> int:
        return self._pool.launch_count()

    def drain(self):
        """Results of every pipelined call (`slot > 0`) become valid on the caller's current stream."""
        self._pool.drain()

    def drain_host(self):
        """Blocks the host until every `submit_inference_chain_host` has delivered its cache."""
        for st in list(self._pool._streams.values()):
            st.synchronize()
        if hasattr(self, "_host_inflight"):
            self._host_inflight.clear()

    # ------------------------------------------------------------------ reference surface
    @torch.no_grad()
    def forward(self, image_features, src_text, tgt_text, src_len=None):
        """Teacher-forced logits (B, T, V) for image_features (B,1024,14,14), src_text (B,S), tgt_text (B,T).
        `src_len` (B,) optionally marks per-question valid src lengths (keys beyond are masked)."""
        img_tokens = project_images(self, image_features)
        h = self._native()
        src = _dev(src_text, "src_text", torch.int64)
        tgt = _dev(tgt_text, "tgt_text", torch.int64)
        B, T = tgt.shape
        sl = None if src_len is None else _dev(src_len, "src_len", torch.int32)
        logits = torch.empty(B, T, self.vocab_size, dtype=torch.float32, device=tgt.device)
        with torch.cuda.device(tgt.device):
            nat.check(nat.lib().b200vqa_fa_forward(h.raw, nat.ptr(img_tokens), nat.ptr(src), nat.ptr(sl), src.shape[1], B,
                                                   nat.ptr(tgt), T, nat.ptr(logits), nat.stream_ptr(tgt.device)),
                      "b200vqa_fa_forward")
        return logits


def _dev(t, what, dtype):
    if not torch.is_tensor(t):
        raise TypeError(f"{what} must be a tensor")
    if not t.is_cuda:
        raise nat.NativeError(f"{what} is on {t.device}; the B200 executor has no CPU path")
    return t.to(dtype).contiguous()


# ---------------------------------------------------------------------------------------------------
# vocabulary helpers (FA:63-78)
# ---------------------------------------------------------------------------------------------------
def load_vocab(vocab_path):
    """Flat {token: index} JSON -> (vocab, {index: token})."""
    with open(vocab_path, "r") as f:
        vocab = json.load(f)
    return vocab, {int(idx): tok for tok, idx in vocab.items()}


def decode_tokens(token_indices, rev_vocab):
    return " ".join(rev_vocab.get(idx, "<unk>") for idx in token_indices)


def tokenize_field(text: str, field: str) -> list:
    """`function` fields are one token; everything else splits on whitespace with `[` / `]` as own tokens."""
    if field == "function":
        return [text] if text else []
    return re.findall(r"\[|\]|[^\[\]\s]+", text)


# ---------------------------------------------------------------------------------------------------
# batched executor
# ---------------------------------------------------------------------------------------------------
@torch.no_grad()
def project_images(model, image_features, slot=0):
    """(B,1024,14,14) or (B,1024,196) f32 -> opaque bf16 image tokens (B,196,d) with the positional rows
    0..195 folded in; computed once per question and reused by every program step."""
    h = model._native(slot)
    img = _dev(image_features, "image_features", torch.float32)
    B = img.shape[0]
    img = img.reshape(B, model.image_proj.in_features, -1)
    if img.shape[2] != model.max_img_tokens:
        raise ValueError(f"expected {model.max_img_tokens} spatial positions, got {img.shape[2]}")
    out = torch.empty(B, model.max_img_tokens, model.d_model, dtype=torch.bfloat16, device=img.device)
    with torch.cuda.device(img.device):
        nat.check(nat.lib().b200vqa_fa_project_images(h.raw, nat.ptr(img), B, nat.ptr(out), nat.stream_ptr(img.device)),
                  "b200vqa_fa_project_images")
    return out


@torch.no_grad()
def greedy_decode(model, image_features, src_text, start_token, max_len, device, src_len=None, forced=None,
                  want_logits=False, img_tokens=None):
    """Greedy decode of one program step: returns `ys` (B, max_len) i64 whose column 0 is `start_token`
    (FA:126-146).  Batched (the reference handles B = 1 only); `src_len` masks padded src columns.
    With `want_logits` also returns the (B, max_len-1, V) logits; `forced` (B, max_len-1) teacher-forces."""
    h = model._native()
    if img_tokens is None:
        img_tokens = project_images(model, image_features.to(device))
    src = _dev(src_text.to(device), "src_text", torch.int64)
    B = src.shape[0]
    sl = None if src_len is None else _dev(src_len.to(device), "src_len", torch.int32)
    fz = None if forced is None else _dev(forced.to(device), "forced", torch.int64)
    ys = torch.empty(B, max_len, dtype=torch.int64, device=src.device)
    logits = torch.empty(B, max_len - 1, model.vocab_size, dtype=torch.float32, device=src.device) if want_logits else None
    with torch.cuda.device(src.device):
        nat.check(nat.lib().b200vqa_fa_step(h.raw, nat.ptr(img_tokens), nat.ptr(src), nat.ptr(sl), src.shape[1], B,
                                            int(start_token), int(max_len), nat.ptr(ys), nat.ptr(logits), nat.ptr(fz),
                                            nat.stream_ptr(src.device)), "b200vqa_fa_step")
    return (ys, logits) if want_logits else ys


def chain_to_arrays(final_chain, rev_vocab, max_steps=None):
    """Parses one question's `final_chain_of_thought` (list of "func dep dep" strings of vocab indices) into
    (func [S] i32, deps [S, 2] i32 with -1 = none) following run_inference_chain's rules (FA:96-108): a token
    after the function is a dependency pointer iff its vocabulary entry is a digit string; anything else is
    skipped with a warning."""
    S = len(final_chain) if max_steps is None else max_steps
    func = np.zeros(S, dtype=np.int32)
    deps = np.full((S, MAX_DEPS), -1, dtype=np.int32)
    for i, elem in enumerate(final_chain):
        parts = elem.strip().split()
        func[i] = int(parts[0])
        k = 0
        for tok in parts[1:]:
            original = rev_vocab.get(int(tok), None) if rev_vocab is not None else None
            if original is not None and original.isdigit():
                if k >= MAX_DEPS:
                    raise ValueError(f"chain step {i} has more than {MAX_DEPS} dependency pointers")
                deps[i, k] = int(original)
                k += 1
            else:
                logger.warning("Token %s in chain step %d is not recognized as a digit; skipping.", tok, i)
    return func, deps


@torch.no_grad()
def run_inference_chain_batched(model, image_features, func, deps, n_steps, start_token=0, max_infer_len=20,
                                forced=None, want_logits=False, img_tokens=None, sort_by_steps=True, slot=0,
                                image_idx=None):
    """Executes B programs at once with the inference cache in HBM.

    func (B,S) i32, deps (B,S,2) i32 (-1 = none), n_steps (B,) i32  ->  cache (B,S,max_infer_len) i32 where
    cache[b,i] holds the max_infer_len tokens of step i (start token included, FA:120-121) and rows with
    i >= n_steps[b] stay -1.  Questions are processed longest-program-first so that finished questions drop
    out of later steps (`sort_by_steps`); results are returned in the caller's order.

    image_idx (B,) optional: several questions per image (CLEVR ~10) - `image_features` / `img_tokens` then hold every
    image ONCE (n_images rows) and question b uses image image_idx[b]; the projection runs once per image.
    """
    if slot > 0:
        # pipelined submission: the whole call runs on the slot's own handle + stream; results are valid after
        # model.drain() (independent batches overlap: one batch's small decode kernels fill the other's bubbles)
        st = model._pool.stream(slot)
        st.wait_stream(torch.cuda.current_stream(st.device))
        with torch.cuda.stream(st):
            return _chain_batched(model, image_features, func, deps, n_steps, start_token, max_infer_len, forced,
                                  want_logits, img_tokens, sort_by_steps, slot, image_idx)
    return _chain_batched(model, image_features, func, deps, n_steps, start_token, max_infer_len, forced, want_logits,
                          img_tokens, sort_by_steps, 0, image_idx)


def _chain_batched(model, image_features, func, deps, n_steps, start_token, max_infer_len, forced, want_logits,
                   img_tokens, sort_by_steps, slot, image_idx=None):
    h = model._native(slot)
    dev = model.image_proj.weight.device
    # program lengths on the host size the per-step launches (finished questions drop out): taken from the caller's
    # tensor when it already lives on the host - a device tensor costs one device->host read here
    ns_host_in = n_steps.detach().to(torch.int32).numpy() if not n_steps.is_cuda else None
    func = _dev(func.to(dev, non_blocking=True), "func", torch.int32)
    deps = _dev(deps.to(dev, non_blocking=True), "deps", torch.int32)
    n_steps = _dev(n_steps.to(dev, non_blocking=True), "n_steps", torch.int32)
    B, S = func.shape
    if tuple(deps.shape) != (B, S, MAX_DEPS) or tuple(n_steps.shape) != (B,):
        raise ValueError("deps must be (B,S,2) and n_steps (B,)")
    if img_tokens is None:
        img_tokens = project_images(model, image_features.to(dev), slot)
    if image_idx is not None:
        image_idx = _dev(image_idx.to(dev), "image_idx", torch.int32)
        if tuple(image_idx.shape) != (B,):
            raise ValueError("image_idx must be (B,)")
        if B and (int(image_idx.min()) < 0 or int(image_idx.max()) >= img_tokens.shape[0]):
            raise IndexError("image_idx out of range")
    elif img_tokens.shape[0] != B:
        raise ValueError("one image per question expected (or pass image_idx)")
    T = max_infer_len - 1
    order = None
    active = None
    if sort_by_steps and B > 1:
        if ns_host_in is not None:
            order_host = np.argsort(-ns_host_in.astype(np.int64), kind="stable")
            order = torch.from_numpy(order_host).to(dev, non_blocking=True)
        else:
            order = torch.argsort(n_steps, descending=True, stable=True)
        func, deps, n_steps = func[order].contiguous(), deps[order].contiguous(), n_steps[order].contiguous()
        if image_idx is not None:
            image_idx = image_idx[order].contiguous()     # the tokens themselves stay where they are
        else:
            img_tokens = img_tokens[order].contiguous()
        if forced is not None:
            forced = forced.to(dev)[order]
        ns_host = ns_host_in[order_host] if ns_host_in is not None else n_steps.cpu().numpy()
        active = np.ascontiguousarray([(ns_host > i).sum() for i in range(S)], dtype=np.int32)
    fz = None if forced is None else _dev(forced.to(dev), "forced", torch.int64)
    cache = torch.full((B, S, max_infer_len), -1, dtype=torch.int32, device=dev)
    logits = torch.zeros(B, S, T, model.vocab_size, dtype=torch.float32, device=dev) if want_logits else None
    active_ptr = None if active is None else active.ctypes.data
    with torch.cuda.device(dev):
        if image_idx is None:
            rc = nat.lib().b200vqa_fa_run_chain(h.raw, nat.ptr(img_tokens), nat.ptr(func), nat.ptr(deps),
                                                nat.ptr(n_steps), B, S, int(start_token), int(max_infer_len),
                                                nat.ptr(cache), active_ptr, nat.ptr(logits), nat.ptr(fz),
                                                nat.stream_ptr(dev))
        else:
            rc = nat.lib().b200vqa_fa_run_chain_indexed(h.raw, nat.ptr(img_tokens), int(img_tokens.shape[0]),
                                                        nat.ptr(image_idx), nat.ptr(func), nat.ptr(deps),
                                                        nat.ptr(n_steps), B, S, int(start_token), int(max_infer_len),
                                                        nat.ptr(cache), active_ptr, nat.ptr(logits), nat.ptr(fz),
                                                        nat.stream_ptr(dev))
        nat.check(rc, "b200vqa_fa_run_chain")
    if order is not None:
        inv = torch.empty_like(order)
        inv[order] = torch.arange(B, device=dev)
        cache = cache[inv]
        if logits is not None:
            logits = logits[inv]
    return (cache, logits) if want_logits else cache


def resolve_upload(upload):
    """"fp32": the feature bytes cross PCIe as given.  "bf16": the library rounds them to bf16 on host threads while the
    previous group of images is on the wire - half the bytes, and bit-identical results, because the device path rounds
    the features to bf16 anyway (the transpose + cast in front of image_proj).  "auto": "bf16" when this process has
    at least 8 CPUs to itself (CPUs it may run on / ranks on the node)."""
    if upload == "auto":
        cpus = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        ranks = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1"))))
        return "bf16" if cpus // ranks >= 8 else "fp32"
    if upload not in ("fp32", "bf16"):
        raise ValueError('upload must be "fp32", "bf16" or "auto"')
    return upload


@torch.no_grad()
def run_inference_chain_host(model, image_features_cpu, func, deps, n_steps, start_token=0, max_infer_len=20,
                             chunk=2048, parts=2, upload="fp32"):
    """`run_inference_chain_batched` with HOST tensors in and out - the library call that replaces the reference's
    driver loop (FA:193-206) with its per-step uploads and downloads (FA:109-121): image_features_cpu (B,1024,14,14) f32
    (pinned for full PCIe speed), func (B,S), deps (B,S,2), n_steps (B,) on the host -> cache (B,S,max_infer_len) i32
    in pinned host memory.  The batch is cut into `parts` contiguous ranges that run concurrently on separate
    (handle, stream) slots - a chain is a dependent sequence of small kernels per program step, two of them overlap well
    - and inside a part features are uploaded and projected in sub-batches of `chunk` questions while the previous
    sub-batch executes.  Results do not depend on `chunk` / `parts` (questions never interact).  This is what bench.py
    times as `e2e` for the FA workload."""
    img = image_features_cpu.to(torch.float32).contiguous()
    f = func.to(torch.int32).contiguous()
    d = deps.to(torch.int32).contiguous()
    n = n_steps.to(torch.int32).contiguous()
    if img.is_cuda or f.is_cuda or d.is_cuda or n.is_cuda:
        raise ValueError("run_inference_chain_host takes CPU tensors; use run_inference_chain_batched for device tensors")
    B, S = f.shape
    if tuple(d.shape) != (B, S, MAX_DEPS) or tuple(n.shape) != (B,) or img.shape[0] != B:
        raise ValueError("expected image_features (B,1024,14,14), func (B,S), deps (B,S,2), n_steps (B,)")
    if B and img[0].numel() != model.image_proj.in_features * model.max_img_tokens:
        raise ValueError(f"image_features must hold {model.image_proj.in_features} x {model.max_img_tokens} values per question")
    cache = torch.empty(B, S, max_infer_len, dtype=torch.int32).pin_memory()
    dev = model.image_proj.weight.device
    parts = max(1, min(int(parts), B // 256 if B >= 512 else 1))
    half = resolve_upload(upload) == "bf16"
    with torch.cuda.device(dev):
        if parts == 1:
            h = model._native(0)
            h.set_host_upload(half)
            nat.check(nat.lib().b200vqa_fa_run_chain_host(h.raw, nat.ptr(img), nat.ptr(f), nat.ptr(d), nat.ptr(n), B, S,
                                                          int(start_token), int(max_infer_len), nat.ptr(cache), int(chunk),
                                                          nat.stream_ptr(dev)), "b200vqa_fa_run_chain_host")
            return cache
        streams = []
        for i in range(parts):
            lo, hi = B * i // parts, B * (i + 1) // parts
            h = model._native(1 + i)
            h.set_host_upload(half)
            st = model._pool.stream(1 + i)
            streams.append(st)
            nat.check(nat.lib().b200vqa_fa_run_chain_host_async(
                h.raw, nat.ptr(img[lo:hi]), nat.ptr(f[lo:hi]), nat.ptr(d[lo:hi]), nat.ptr(n[lo:hi]), hi - lo, S,
                int(start_token), int(max_infer_len), nat.ptr(cache[lo:hi]), int(min(chunk, hi - lo)),
                C.c_void_p(st.cuda_stream)), "b200vqa_fa_run_chain_host_async")
        for st in streams:
            st.synchronize()   # img / f / d / n / cache are referenced by this frame until every part has finished
    return cache


@torch.no_grad()
def submit_inference_chain_host(model, image_features_cpu, func, deps, n_steps, start_token=0, max_infer_len=20,
                                chunk=1024, depth=2, out=None, upload="fp32"):
    """`run_inference_chain_host` without the final synchronisation, on a round-robin (handle, stream) slot: the upload
    of this batch runs under the chains of the previous one.  Returns the pinned host cache (`out` when given: a
    pinned int32 (B,S,max_infer_len) tensor the caller recycles), valid after `model.drain_host()`; the host inputs are
    kept alive until then."""
    img = image_features_cpu.to(torch.float32).contiguous()
    f = func.to(torch.int32).contiguous()
    d = deps.to(torch.int32).contiguous()
    n = n_steps.to(torch.int32).contiguous()
    if img.is_cuda or f.is_cuda or d.is_cuda or n.is_cuda:
        raise ValueError("submit_inference_chain_host takes CPU tensors")
    B, S = f.shape
    if tuple(d.shape) != (B, S, MAX_DEPS) or tuple(n.shape) != (B,) or img.shape[0] != B:
        raise ValueError("expected image_features (B,1024,14,14), func (B,S), deps (B,S,2), n_steps (B,)")
    if B and img[0].numel() != model.image_proj.in_features * model.max_img_tokens:
        raise ValueError(f"image_features must hold {model.image_proj.in_features} x {model.max_img_tokens} values per question")
    if not hasattr(model, "_host_inflight"):
        model._host_inflight, model._next_host_slot = [], 0
    slot = 1 + model._next_host_slot % max(1, int(depth))
    model._next_host_slot += 1
    h = model._native(slot)
    h.set_host_upload(resolve_upload(upload) == "bf16")
    st = model._pool.stream(slot)
    if out is None:
        cache = torch.empty(B, S, max_infer_len, dtype=torch.int32).pin_memory()
    else:
        cache = out
        if cache.dtype != torch.int32 or tuple(cache.shape) != (B, S, max_infer_len) or not cache.is_pinned() \
                or not cache.is_contiguous():
            raise ValueError("out must be a pinned contiguous int32 tensor of shape (B, S, max_infer_len)")
    with torch.cuda.device(st.device):
        nat.check(nat.lib().b200vqa_fa_run_chain_host_async(
            h.raw, nat.ptr(img), nat.ptr(f), nat.ptr(d), nat.ptr(n), B, S, int(start_token), int(max_infer_len),
            nat.ptr(cache), int(min(chunk, B)), C.c_void_p(st.cuda_stream)), "b200vqa_fa_run_chain_host_async")
    model._host_inflight.append((img, f, d, n, cache))
    return cache


def run_inference_chain(model, image_features, final_chain, device, start_token, max_infer_len=20, rev_vocab=None):
    """Reference-compatible single-question entry (FA:83-124): returns (final_output, cache) where cache maps
    step index -> space-joined predicted token ids and final_output is the last step's string."""
    model.eval()
    func, deps = chain_to_arrays(final_chain, rev_vocab)
    S = len(final_chain)
    for i in range(S):
        for d in deps[i]:
            if d >= 0 and (d >= i):
                logger.warning("Input index %d not found in cache for chain step %d. Using empty string.", d, i)
    cache_t = run_inference_chain_batched(
        model, image_features.to(device), torch.from_numpy(func)[None], torch.from_numpy(deps)[None],
        torch.tensor([S], dtype=torch.int32), start_token=start_token, max_infer_len=max_infer_len)
    rows = cache_t[0].cpu().tolist()
    cache = {i: " ".join(map(str, rows[i])) for i in range(S)}
    for i in range(S):
        logger.info("Chain step %d: function token %s, input steps %s, predicted output: %s", i, func[i],
                    [int(d) for d in deps[i] if d >= 0], cache[i])
    return cache[S - 1], cache


def main_inference(model_path="multimodal_transformer.pth", vocab_path="vocab.json",
                   annotated_h5_path="annotated_questions_with_vocab.h5", features_h5_path="train_features.h5",
                   num_examples=10):
    """The reference's demo driver (FA:151-206) with its hyper-parameters, batched over the examples."""
    import h5py
    vocab, rev_vocab = load_vocab(vocab_path)
    device = torch.device("cuda")
    model = MultiModalTransformer(len(vocab), 256, 2, 1, 1, 512, 0.1, 50, max_img_tokens=196).to(device)
    model.load_state_dict(torch.load(model_path, map_location=device))
    model.eval()
    with h5py.File(annotated_h5_path, "r") as hf:
        annotated = json.loads(hf["questions"][()].decode("utf-8"))["questions"][:num_examples]
    S = max(len(q["final_chain_of_thought"]) for q in annotated)
    parsed = [chain_to_arrays(q["final_chain_of_thought"], rev_vocab, S) for q in annotated]
    func = torch.from_numpy(np.stack([p[0] for p in parsed]))
    deps = torch.from_numpy(np.stack([p[1] for p in parsed]))
    n_steps = torch.tensor([len(q["final_chain_of_thought"]) for q in annotated], dtype=torch.int32)
    with h5py.File(features_h5_path, "r") as hf:
        feats = torch.from_numpy(np.stack([hf["features"][q["image_index"]] for q in annotated])).float()
    cache = run_inference_chain_batched(model, feats.to(device), func, deps, n_steps, 0, 20).cpu()
    for i, q in enumerate(annotated):
        final = cache[i, int(n_steps[i]) - 1].tolist()
        logger.info("Question %d actual answer: %s", i, q.get("answer", "Not provided"))
        logger.info("Question %d final predicted sentence: %s", i, decode_tokens(final, rev_vocab))
    return cache


if __name__ == "__main__":
    logging.basicConfig(level=logging.INFO, format="%(asctime)s - %(levelname)s: %(message)s")
    main_inference()
