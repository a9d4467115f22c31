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


# ----------------------------------------------------------------------------------------------------------
# Evaluation of the bounding-box variant (reference :104-150 IoU utilities, :423-538 `evaluate`).  The reference
# walks every box and every sample in Python with an `.item()` / `.tolist()` per element; here the IoUs, the
# four-way program x answer tally and the token accuracy are whole-batch tensor expressions on the device the
# model's outputs live on, accumulated there, and read back ONCE at the end of the data set.
# ----------------------------------------------------------------------------------------------------------
def bbox_iou_2d(pred_box, gt_box):
    """IoU of two [x_min, y_min, x_max, y_max] boxes (reference :104-124); 0.0 when the union is empty."""
    iw = max(min(pred_box[2], gt_box[2]) - max(pred_box[0], gt_box[0]), 0.0)
    ih = max(min(pred_box[3], gt_box[3]) - max(pred_box[1], gt_box[1]), 0.0)

    def area(b):
        return max(b[2] - b[0], 0.0) * max(b[3] - b[1], 0.0)

    inter = iw * ih
    union = area(pred_box) + area(gt_box) - inter
    return inter / union if union > 0.0 else 0.0


def batch_iou_sum_count(bbox_preds, bbox_gts):
    """(sum of IoUs, number of ground-truth boxes) of one batch as 0-d tensors on the inputs' device - no host
    synchronisation.  Semantics of the reference loop (:126-150): predictions clamped to [0, 1], ground-truth boxes
    whose four coordinates are all |x| < 1e-8 are padding and skipped, an empty union scores 0."""
    p = bbox_preds.detach().double().clamp(0.0, 1.0)
    g = bbox_gts.detach().to(p.device).double()
    real = ~(g.abs() < 1e-8).all(dim=-1)
    iw = (torch.minimum(p[..., 2], g[..., 2]) - torch.maximum(p[..., 0], g[..., 0])).clamp_min(0.0)
    ih = (torch.minimum(p[..., 3], g[..., 3]) - torch.maximum(p[..., 1], g[..., 1])).clamp_min(0.0)
    inter = iw * ih
    area_p = (p[..., 2] - p[..., 0]).clamp_min(0.0) * (p[..., 3] - p[..., 1]).clamp_min(0.0)
    area_g = (g[..., 2] - g[..., 0]).clamp_min(0.0) * (g[..., 3] - g[..., 1]).clamp_min(0.0)
    union = area_p + area_g - inter
    iou = torch.where(union > 0.0, inter / union.clamp_min(torch.finfo(torch.float64).tiny), torch.zeros_like(union))
    return (iou * real).sum(), real.sum()


def batch_mean_iou(bbox_preds, bbox_gts):
    """Mean IoU over the non-padding ground-truth boxes of a (B, 10, 4) batch, 0.0 when there are none (:126-150)."""
    s, n = batch_iou_sum_count(bbox_preds, bbox_gts)
    n = int(n)
    return float(s) / n if n else 0.0


def get_data_info(questions_h5_path):
    """(question vocab, program + answer vocab) = max id + 1 over the H5 arrays (reference :361-369)."""
    import h5py  # optional dependency: only the dataset helpers need it
    import numpy as np
    with h5py.File(questions_h5_path, "r") as f:
        return (int(np.max(f["questions"])) + 1, max(int(np.max(f["programs"])), int(np.max(f["answers"]))) + 1)


@torch.no_grad()
def evaluate(model, dataloader, criterion_seq, criterion_bbox, device):
    """The reference's evaluation loop (:423-538) over batches of (image_features, questions, combined_seq (B, 28),
    bboxes_gt (B, 10, 4)).  Returns the same tuple: (loss per sample, mean over batches of the batch-mean IoU,
    share of samples with correct program + correct answer, correct + incorrect, incorrect + correct,
    incorrect + incorrect, program token accuracy).  `criterion_bbox` must have reduction='none' (:616)."""
    model.eval()
    dev = torch.device(device)
    # [loss x batch size, samples, sum of batch-mean IoUs, batches, cpca, cpia, ipca, ipia, tokens right, tokens]
    acc = torch.zeros(10, dtype=torch.float64, device=dev)
    for image_features, questions, combined_seq, bboxes_gt in dataloader:
        image_features, questions = image_features.to(dev), questions.to(dev)
        combined_seq, bboxes_gt = combined_seq.to(dev), bboxes_gt.to(dev)
        n = image_features.size(0)
        seq_logits, bbox_preds = model(image_features, questions)
        loss_seq = criterion_seq(seq_logits.reshape(-1, seq_logits.size(-1)), combined_seq.reshape(-1))
        mask = bboxes_gt.sum(dim=2, keepdim=True) > 0                                     # :471
        loss_bbox = (criterion_bbox(bbox_preds, bboxes_gt) * mask).sum() / mask.sum()     # :472-474
        iou_sum, iou_n = batch_iou_sum_count(bbox_preds, bboxes_gt)
        batch_iou = torch.where(iou_n > 0, iou_sum / iou_n.clamp_min(1), torch.zeros_like(iou_sum))
        predicted = seq_logits.argmax(dim=2)       # lowest index wins ties, as torch.max(…, dim=2) does (:487)
        tok_ok = predicted[:, :-1] == combined_seq[:, :-1]
        prog_ok = tok_ok.all(dim=1)
        ans_ok = predicted[:, -1] == combined_seq[:, -1]
        step = torch.stack([
            (loss_seq + loss_bbox).double() * n, acc.new_tensor(n), batch_iou, acc.new_tensor(1.0),
            (prog_ok & ans_ok).sum().double(), (prog_ok & ~ans_ok).sum().double(),
            (~prog_ok & ans_ok).sum().double(), (~prog_ok & ~ans_ok).sum().double(),
            tok_ok.sum().double(), acc.new_tensor(tok_ok.numel())])
        acc += step
    a = acc.tolist()  # the only device -> host read of the evaluation
    total, batches, tokens = a[1], a[3], a[9]
    if total == 0:
        raise ZeroDivisionError("evaluate: empty dataloader")  # the reference divides by total == 0 as well (:519)
    return (a[0] / total, a[2] / batches if batches > 0 else 0.0, a[4] / total, a[5] / total, a[6] / total,
            a[7] / total, a[8] / tokens if tokens > 0 else 0.0)
