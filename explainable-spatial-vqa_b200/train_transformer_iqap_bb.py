"""Inference mirror of the reference's bounding-box variant of the IQAP model
(`/root/reference/code/train_transformer_iqap_bb.py:222-356`): the same image + question encoder, a CONTINUOUS
bounding-box regression head on the mean of the image-token rows of the encoder memory (:272-276, :304-310 -
`bbox_regressor` = Linear(256, hidden) -> ReLU -> Linear(hidden, 40), viewed as 10 boxes x 4), and a one-layer decoder
that greedily emits the combined program + answer sequence of 28 tokens and returns its logits (:312-356).

Only the inference surface is mirrored (constructor, state-dict layout, `forward`); the training loop of the reference
file is out of scope.  Everything runs through the same C ABI as `inference_transformer_iqap` - the regression head is
the library's answer-head kernel in its pooled mode (`b200vqa_model_desc.answer_pool_rows`).  There is no CPU path.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _native as nat
from .inference_transformer_iqap import PositionalEncoding, generate_square_subsequent_mask  # noqa: F401


class Config:
    """Hyper-parameters of train_transformer_iqap_bb.py:23-60 that the model reads."""
    EMBEDDING_DIM = 256
    HIDDEN_DIM = 256
    IMAGE_FEATURE_DIM = 1024
    MAX_QUESTION_LEN = 46
    NUM_IMAGE_TOKENS = 196
    PROGRAM_SEQ_LEN = 27
    SPECIAL_TOKEN_ID = 1
    NUM_BOXES = 10


class VQAModel(nn.Module):
    """image features + question -> (sequence logits (B, 28, Vp), bounding boxes (B, 10, 4)); reference :222-356."""

    def __init__(self, vocab_size, embedding_dim, hidden_dim, program_vocab_size, program_seq_len, num_image_tokens,
                 special_token_id=1):
        super().__init__()
        self.embedding_dim = embedding_dim
        self.hidden_dim = hidden_dim
        self.num_image_tokens = num_image_tokens
        self.special_token_id = special_token_id
        self.seq_len = program_seq_len + 1
        # construction order == the reference's (:243-276), so a seeded default init yields identical parameters
        self.image_proj = nn.Linear(Config.IMAGE_FEATURE_DIM, embedding_dim)
        self.embedding = nn.Embedding(vocab_size, embedding_dim, padding_idx=0)
        self.cls_token = nn.Parameter(torch.randn(1, 1, embedding_dim))
        self.pos_encoder = PositionalEncoding(embedding_dim, dropout=0.1,
                                              max_len=num_image_tokens + Config.MAX_QUESTION_LEN + 1)
        enc_layer = nn.TransformerEncoderLayer(d_model=embedding_dim, nhead=4)
        self.transformer_encoder = nn.TransformerEncoder(enc_layer, num_layers=1)
        self.decoder_embedding = nn.Embedding(program_vocab_size, embedding_dim, padding_idx=0)
        self.pos_decoder = PositionalEncoding(embedding_dim, dropout=0.1, max_len=self.seq_len + 1)
        dec_layer = nn.TransformerDecoderLayer(d_model=embedding_dim, nhead=4)
        self.transformer_decoder = nn.TransformerDecoder(dec_layer, num_layers=1)
        self.output_layer = nn.Linear(embedding_dim, program_vocab_size)
        self.bbox_regressor = nn.Sequential(nn.Linear(embedding_dim, hidden_dim), nn.ReLU(),
                                            nn.Linear(hidden_dim, 4 * Config.NUM_BOXES))
        self._pool = nat.HandlePool(self, self._build_desc)

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
        d.dec_vocab = self.output_layer.out_features
        d.pe_enc_len = self.pos_encoder.pe.shape[0]
        d.pe_dec_len = self.pos_decoder.pe.shape[0]
        d.max_q_len = d.pe_enc_len - 1 - self.num_image_tokens
        d.answer_hidden = self.bbox_regressor[0].out_features
        d.num_classes = self.bbox_regressor[2].out_features
        d.answer_pool_rows = self.num_image_tokens  # the head reads mean(memory[1 : 1 + 196]) (:304-307)
        d.layer_norm_eps = enc_layers[0].norm1.eps
        nat._set(d, keep,
                 image_proj_weight=self.image_proj.weight, image_proj_bias=self.image_proj.bias,
                 cls_token=self.cls_token.reshape(-1),
                 enc_embedding=self.embedding.weight, dec_embedding=self.decoder_embedding.weight,
                 pe_enc=self.pos_encoder.pe.reshape(d.pe_enc_len, -1),
                 pe_dec=self.pos_decoder.pe.reshape(d.pe_dec_len, -1),
                 head_weight=self.output_layer.weight, head_bias=self.output_layer.bias,
                 answer_w0=self.bbox_regressor[0].weight, answer_b0=self.bbox_regressor[0].bias,
                 answer_w1=self.bbox_regressor[2].weight, answer_b1=self.bbox_regressor[2].bias)
        d.enc_layers = enc_arr
        d.dec_layers = dec_arr
        return d, keep

    def _native(self) -> nat.Handle:
        h = self._pool.get(0)
        h.set_start_token(Config.SPECIAL_TOKEN_ID)  # <SOS>, read at decode time (:323-328)
        return h

    def forward(self, image_features, questions):
        """image_features (B, 196, 1024) f32, questions (B, 46) i64 -> (seq_logits (B, 28, Vp), bbox_preds (B, 10, 4))."""
        seq_logits, bbox_preds, _ = self.forward_detailed(image_features, questions)
        return seq_logits, bbox_preds

    @torch.no_grad()
    def forward_detailed(self, image_features, questions, forced_tokens=None):
        """Also returns the greedy tokens (B, 28); `forced_tokens` (B, 28) teacher-forces the decoder (parity tests)."""
        h = self._native()
        if not image_features.is_cuda or not questions.is_cuda:
            raise nat.NativeError("inputs must live on the GPU; the B200 executor has no CPU path")
        img = image_features.to(torch.float32).contiguous()
        q = questions.to(torch.int64).contiguous()
        B, T = img.shape[0], self.seq_len
        n_img, feat = self.num_image_tokens, self.image_proj.in_features
        q_len = self.pos_encoder.pe.shape[0] - 1 - n_img
        if tuple(img.shape) != (B, n_img, feat) or tuple(q.shape) != (B, q_len):
            raise ValueError(f"expected image_features (B, {n_img}, {feat}) and questions (B, {q_len})")
        dev = img.device
        V = self.output_layer.out_features
        boxes = torch.empty(B, self.bbox_regressor[2].out_features, dtype=torch.float32, device=dev)
        tokens = torch.empty(B, T, dtype=torch.int64, device=dev)
        logits = torch.empty(B, T, V, dtype=torch.float32, device=dev)
        forced = None
        if forced_tokens is not None:
            forced = forced_tokens.to(device=dev, dtype=torch.int64).contiguous()
            if tuple(forced.shape) != (B, T):
                raise ValueError(f"forced_tokens must be (B, {T})")
        with torch.cuda.device(dev):
            nat.check(nat.lib().b200vqa_iqap_forward(h.raw, nat.ptr(img), nat.ptr(q), B, T, nat.ptr(boxes), nat.ptr(tokens),
                                                     nat.ptr(logits), nat.ptr(forced), None, nat.stream_ptr(dev)),
                      "b200vqa_iqap_forward")
        return logits, boxes.view(B, Config.NUM_BOXES, 4), tokens

    def native_launch_count(self) -> int:
        return self._pool.launch_count()
