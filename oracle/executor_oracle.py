"""CPU restatement (plain fp32 torch ops on the host) of the reference's Program Executor inference path.

TEST INFRASTRUCTURE - see oracle/__init__.py for who may import this.

The arithmetic of the reference lives in PyTorch (third-party, unpinned in code/requirements.txt; torch
2.11.0 in this image): nn.TransformerEncoderLayer / DecoderLayer / nn.Transformer (post-norm, ReLU,
eps 1e-5), nn.MultiheadAttention (packed in_proj q|k|v), nn.Linear, nn.Embedding.  Every function below
restates one reference function from a *state-dict* (the reference's own parameter names) with
F.linear / softmax / layer_norm only and cites the lines it follows:

  IQAP = /root/reference/code/inference_transformer_iqap.py
  FA   = /root/reference/code/inference_transformer_full_annotation_new.py
  BB   = /root/reference/code/train_transformer_iqap_bb.py (the continuous bounding-box head, model class only)

Parity pin: the reference ships no tests, checkpoints or golden vectors for this path, so the
pins are outputs of the reference ITSELF, generated in the build container by oracle/make_golden.py
(imports the reference's modules) and committed under tests/golden/; tests/test_oracle_golden.py checks
this restatement against them.

`recompute=True` follows the reference's execution literally (whole decoder prefix re-run every step, cross
K/V re-projected every step: IQAP:208-236, FA:137-145) and is the mode bench.py times as the CPU
baseline; `recompute=False` is the algebraically identical KV-cached form the CUDA path implements.
"""
from __future__ import annotations

import math
import re

import torch
import torch.nn.functional as F

EPS = 1e-5


# ------------------------------------------------------------------------------------------------
# building blocks (torch/nn/functional.py multi_head_attention_forward; nn/modules/transformer.py)
# ------------------------------------------------------------------------------------------------
def mha(sd, prefix, x_q, x_kv, nhead, attn_mask=None, key_len=None):
    """nn.MultiheadAttention forward, batch-first tensors (B, Lq, d) / (B, Lk, d).
    attn_mask: additive (Lq, Lk) or None; key_len: (B,) valid key counts (keys beyond are masked) or None."""
    w, b = sd[prefix + "in_proj_weight"], sd[prefix + "in_proj_bias"]
    d = x_q.shape[-1]
    dh = d // nhead
    q = F.linear(x_q, w[:d], b[:d])
    k = F.linear(x_kv, w[d:2 * d], b[d:2 * d])
    v = F.linear(x_kv, w[2 * d:], b[2 * d:])
    B, Lq, Lk = x_q.shape[0], x_q.shape[1], x_kv.shape[1]
    q = q.view(B, Lq, nhead, dh).transpose(1, 2) * (1.0 / math.sqrt(dh))
    k = k.view(B, Lk, nhead, dh).transpose(1, 2)
    v = v.view(B, Lk, nhead, dh).transpose(1, 2)
    s = q @ k.transpose(-1, -2)
    if attn_mask is not None:
        s = s + attn_mask
    if key_len is not None:
        dead = torch.arange(Lk, device=s.device)[None, :] >= key_len[:, None]
        s = s.masked_fill(dead[:, None, None, :], float("-inf"))
    o = torch.softmax(s, dim=-1) @ v
    o = o.transpose(1, 2).reshape(B, Lq, d)
    return F.linear(o, sd[prefix + "out_proj.weight"], sd[prefix + "out_proj.bias"])


def layer_norm(sd, prefix, x):
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + "weight"], sd[prefix + "bias"], EPS)


def ffn(sd, prefix, x):
    return F.linear(torch.relu(F.linear(x, sd[prefix + "linear1.weight"], sd[prefix + "linear1.bias"])),
                    sd[prefix + "linear2.weight"], sd[prefix + "linear2.bias"])


def encoder_layer(sd, prefix, x, nhead, key_len=None):
    """Post-norm TransformerEncoderLayer: x = LN1(x + SA(x)); x = LN2(x + FFN(x))."""
    x = layer_norm(sd, prefix + "norm1.", x + mha(sd, prefix + "self_attn.", x, x, nhead, key_len=key_len))
    return layer_norm(sd, prefix + "norm2.", x + ffn(sd, prefix, x))


def causal_mask(n, device=None):
    return torch.full((n, n), float("-inf"), device=device).triu(1)


def decoder_layer(sd, prefix, x, memory, nhead, mem_len=None):
    """Post-norm TransformerDecoderLayer over the whole prefix x (B, t, d) with a causal mask."""
    x = layer_norm(sd, prefix + "norm1.", x + mha(sd, prefix + "self_attn.", x, x, nhead, causal_mask(x.shape[1], x.device)))
    x = layer_norm(sd, prefix + "norm2.", x + mha(sd, prefix + "multihead_attn.", x, memory, nhead, key_len=mem_len))
    return layer_norm(sd, prefix + "norm3.", x + ffn(sd, prefix, x))


def count_layers(sd, prefix):
    n = 0
    while f"{prefix}{n}.norm1.weight" in sd:
        n += 1
    return n


class _CachedDecoder:
    """KV-cached greedy decoder: identical arithmetic to re-running the causal prefix (each position only
    ever attends to earlier ones), with the cross-attention K/V of `memory` projected once."""

    def __init__(self, sd, layer_prefix, memory, nhead, mem_len=None):
        self.sd, self.pre, self.nhead, self.mem_len = sd, layer_prefix, nhead, mem_len
        self.n = count_layers(sd, layer_prefix)
        d = memory.shape[-1]
        self.d, self.dh = d, d // nhead
        B, L = memory.shape[0], memory.shape[1]
        self.ck, self.cv, self.sk, self.sv = [], [], [], []
        for l in range(self.n):
            w, b = sd[f"{self.pre}{l}.multihead_attn.in_proj_weight"], sd[f"{self.pre}{l}.multihead_attn.in_proj_bias"]
            self.ck.append(F.linear(memory, w[d:2 * d], b[d:2 * d]).view(B, L, nhead, self.dh).transpose(1, 2))
            self.cv.append(F.linear(memory, w[2 * d:], b[2 * d:]).view(B, L, nhead, self.dh).transpose(1, 2))
            self.sk.append(None)
            self.sv.append(None)

    def _attend(self, q, k, v, key_len=None):
        B = q.shape[0]
        qh = q.view(B, 1, self.nhead, self.dh).transpose(1, 2) * (1.0 / math.sqrt(self.dh))
        s = qh @ k.transpose(-1, -2)
        if key_len is not None:
            dead = torch.arange(k.shape[2], device=k.device)[None, :] >= key_len[:, None]
            s = s.masked_fill(dead[:, None, None, :], float("-inf"))
        return (torch.softmax(s, dim=-1) @ v).transpose(1, 2).reshape(B, 1, self.d)

    def step(self, x):
        """x (B, 1, d): embedding + positional row of the newest position -> decoder output (B, 1, d)."""
        sd, d, B = self.sd, self.d, x.shape[0]
        for l in range(self.n):
            p = f"{self.pre}{l}."
            w, b = sd[p + "self_attn.in_proj_weight"], sd[p + "self_attn.in_proj_bias"]
            qkv = F.linear(x, w, b)
            q, k, v = qkv[..., :d], qkv[..., d:2 * d], qkv[..., 2 * d:]
            k = k.view(B, 1, self.nhead, self.dh).transpose(1, 2)
            v = v.view(B, 1, self.nhead, self.dh).transpose(1, 2)
            self.sk[l] = k if self.sk[l] is None else torch.cat([self.sk[l], k], dim=2)
            self.sv[l] = v if self.sv[l] is None else torch.cat([self.sv[l], v], dim=2)
            a = self._attend(q, self.sk[l], self.sv[l])
            x = layer_norm(sd, p + "norm1.", x + F.linear(a, sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"]))
            w, b = sd[p + "multihead_attn.in_proj_weight"], sd[p + "multihead_attn.in_proj_bias"]
            q = F.linear(x, w[:d], b[:d])
            a = self._attend(q, self.ck[l], self.cv[l], self.mem_len)
            x = layer_norm(sd, p + "norm2.", x + F.linear(a, sd[p + "multihead_attn.out_proj.weight"], sd[p + "multihead_attn.out_proj.bias"]))
            x = layer_norm(sd, p + "norm3.", x + ffn(sd, p, x))
        return x


# ------------------------------------------------------------------------------------------------
# IQAP: VQAModel.forward (IQAP:136-188) + autoregressive_program_generation (IQAP:190-241)
# ------------------------------------------------------------------------------------------------
def iqap_encode(sd, image_features, questions, nhead=4):
    """-> memory (B, S, d) batch-first (the reference holds it seq-first; attention does not care)."""
    B = image_features.shape[0]
    img = F.linear(image_features, sd["image_proj.weight"], sd["image_proj.bias"])          # IQAP:152
    qe = F.embedding(questions, sd["embedding.weight"])                                      # IQAP:156 (row 0 is zero)
    cls = sd["cls_token"].expand(B, -1, -1)                                                  # IQAP:160
    x = torch.cat([cls, img, qe], dim=1)                                                     # IQAP:164
    x = x + sd["pos_encoder.pe"][: x.shape[1], 0][None]                                      # IQAP:170
    for l in range(count_layers(sd, "transformer_encoder.layers.")):                          # IQAP:173, no mask, no final norm
        x = encoder_layer(sd, f"transformer_encoder.layers.{l}.", x, nhead)
    return x


def iqap_answer(sd, memory):
    cls = memory[:, 0]                                                                       # IQAP:176
    h = torch.relu(F.linear(cls, sd["answer_classifier.0.weight"], sd["answer_classifier.0.bias"]))
    return F.linear(h, sd["answer_classifier.3.weight"], sd["answer_classifier.3.bias"])     # IQAP:179


def iqap_decode(sd, memory, T=27, start_token=1, nhead=4, forced=None, recompute=False):
    """Greedy program decode.  Returns (tokens (B,T) i64, logits (B,T,Vp)).  `forced` (B,T): position t+1 is
    fed forced[:, t] instead of the argmax (teacher forcing); tokens still hold the argmax."""
    B = memory.shape[0]
    emb, pe = sd["program_decoder_embedding.weight"], sd["pos_decoder.pe"][:, 0]
    prefix = torch.full((B, 1), start_token, dtype=torch.long, device=memory.device)         # IQAP:205
    dec = None if recompute else _CachedDecoder(sd, "transformer_decoder.layers.", memory, nhead)
    n_layers = count_layers(sd, "transformer_decoder.layers.")
    toks, logits = [], []
    for t in range(T):
        if recompute:
            x = F.embedding(prefix, emb) + pe[: prefix.shape[1]][None]                       # IQAP:210-216
            for l in range(n_layers):
                x = decoder_layer(sd, f"transformer_decoder.layers.{l}.", x, memory, nhead)  # IQAP:223-227
            last = x[:, -1]
        else:
            last = dec.step(F.embedding(prefix[:, -1:], emb) + pe[t][None, None])[:, 0]
        lg = F.linear(last, sd["program_output.weight"], sd["program_output.bias"])          # IQAP:230
        nxt = torch.max(lg, dim=1)[1]                                                        # IQAP:233
        toks.append(nxt)
        logits.append(lg)
        feed = forced[:, t] if forced is not None else nxt
        prefix = torch.cat([prefix, feed[:, None]], dim=1)                                   # IQAP:236
    return torch.stack(toks, dim=1), torch.stack(logits, dim=1)


@torch.no_grad()
def iqap_forward(sd, image_features, questions, T=27, forced=None, recompute=False):
    """-> dict(answer (B,C), programs (B,T), logits (B,T,Vp), memory (S,B,d) seq-first like the reference)."""
    memory = iqap_encode(sd, image_features.float(), questions.long())
    answer = iqap_answer(sd, memory)
    programs, logits = iqap_decode(sd, memory, T, 1, 4, forced, recompute)
    return {"answer": answer, "programs": programs, "logits": logits, "memory": memory.transpose(0, 1).contiguous()}


# ------------------------------------------------------------------------------------------------
# IQAP with the continuous bounding-box head: VQAModel.forward / autoregressive_decode of
# BB = /root/reference/code/train_transformer_iqap_bb.py (:287-356)
# ------------------------------------------------------------------------------------------------
def iqap_bb_boxes(sd, memory, n_img=196):
    """memory (B, S, d) batch-first -> boxes (B, 10, 4): bbox_regressor(mean(memory[1 : 1 + 196]))   (BB:304-310)."""
    pooled = memory[:, 1:1 + n_img].mean(dim=1)                                              # BB:305-307
    h = torch.relu(F.linear(pooled, sd["bbox_regressor.0.weight"], sd["bbox_regressor.0.bias"]))
    return F.linear(h, sd["bbox_regressor.2.weight"], sd["bbox_regressor.2.bias"]).view(memory.shape[0], 10, 4)


@torch.no_grad()
def iqap_bb_forward(sd, image_features, questions, seq_len=28, forced=None, recompute=False, nhead=4):
    """-> dict(seq_logits (B, 28, Vp), boxes (B, 10, 4), tokens (B, 28)).  The encoder is IQAP's (BB:287-302 ==
    IQAP:152-173); the decoder is one layer over program + answer tokens, start token <SOS> = 1 (BB:312-356)."""
    memory = iqap_encode(sd, image_features.float(), questions.long(), nhead)
    boxes = iqap_bb_boxes(sd, memory)
    B = memory.shape[0]
    emb, pe = sd["decoder_embedding.weight"], sd["pos_decoder.pe"][:, 0]
    prefix = torch.full((B, 1), 1, dtype=torch.long)                                         # BB:323-329
    dec = None if recompute else _CachedDecoder(sd, "transformer_decoder.layers.", memory, nhead)
    n_layers = count_layers(sd, "transformer_decoder.layers.")
    toks, logits = [], []
    for t in range(seq_len):                                                                 # BB:333
        if recompute:
            x = F.embedding(prefix, emb) + pe[: prefix.shape[1]][None]                       # BB:334-336
            for l in range(n_layers):
                x = decoder_layer(sd, f"transformer_decoder.layers.{l}.", x, memory, nhead)  # BB:341-345
            last = x[:, -1]
        else:
            last = dec.step(F.embedding(prefix[:, -1:], emb) + pe[t][None, None])[:, 0]
        lg = F.linear(last, sd["output_layer.weight"], sd["output_layer.bias"])              # BB:347
        nxt = torch.max(lg, dim=1)[1]                                                        # BB:351
        toks.append(nxt)
        logits.append(lg)
        feed = forced[:, t] if forced is not None else nxt
        prefix = torch.cat([prefix, feed[:, None]], dim=1)                                   # BB:353
    return {"seq_logits": torch.stack(logits, dim=1), "boxes": boxes, "tokens": torch.stack(toks, dim=1)}


# ------------------------------------------------------------------------------------------------
# FA: MultiModalTransformer.forward (FA:45-58), greedy_decode (FA:126-146), run_inference_chain (FA:83-124)
# ------------------------------------------------------------------------------------------------
def fa_encode(sd, image_features, src_text, nhead, src_len=None):
    """image_features (B,1024,14,14) or (B,1024,196); src_text (B,S) -> (memory (B,196+S,d), key_len or None)."""
    B = image_features.shape[0]
    img = image_features.reshape(B, 1024, -1).permute(0, 2, 1)                               # FA:47 / FA:130
    x = torch.cat([F.linear(img, sd["image_proj.weight"], sd["image_proj.bias"]),            # FA:48
                   F.embedding(src_text, sd["text_embedding.weight"])], dim=1)               # FA:49-50
    x = x + sd["pos_encoder.pe"][:, : x.shape[1]]                                            # FA:51
    key_len = None if src_len is None else img.shape[1] + src_len
    for l in range(count_layers(sd, "transformer.encoder.layers.")):
        x = encoder_layer(sd, f"transformer.encoder.layers.{l}.", x, nhead, key_len)
    if "transformer.encoder.norm.weight" in sd:                                              # nn.Transformer adds it
        x = layer_norm(sd, "transformer.encoder.norm.", x)
    return x, key_len


def fa_decode_logits(sd, memory, tgt, nhead, mem_len=None):
    """Teacher-forced decoder + head over tgt (B,T): logits (B,T,V)   (FA:53-57)."""
    x = F.embedding(tgt, sd["text_embedding.weight"]) + sd["pos_decoder.pe"][:, : tgt.shape[1]]
    for l in range(count_layers(sd, "transformer.decoder.layers.")):
        x = decoder_layer(sd, f"transformer.decoder.layers.{l}.", x, memory, nhead, mem_len)
    if "transformer.decoder.norm.weight" in sd:
        x = layer_norm(sd, "transformer.decoder.norm.", x)
    return F.linear(x, sd["output_linear.weight"], sd["output_linear.bias"])


@torch.no_grad()
def fa_forward(sd, image_features, src_text, tgt_text, nhead, src_len=None):
    memory, key_len = fa_encode(sd, image_features.float(), src_text.long(), nhead, src_len)
    return fa_decode_logits(sd, memory, tgt_text.long(), nhead, key_len)


@torch.no_grad()
def fa_greedy_decode(sd, image_features, src_text, start_token, max_len, nhead, src_len=None, forced=None,
                     recompute=False):
    """-> (ys (B,max_len) i64 with column 0 = start_token, logits (B,max_len-1,V)).  Batched with optional
    per-question src_len (the reference is batch-1 and never pads, FA:136)."""
    memory, key_len = fa_encode(sd, image_features.float(), src_text.long(), nhead, src_len)
    B = memory.shape[0]
    ys = torch.full((B, 1), start_token, dtype=torch.long, device=memory.device)             # FA:136
    fed = ys.clone()
    emb, pe = sd["text_embedding.weight"], sd["pos_decoder.pe"][0]
    dec = None if recompute else _CachedDecoder(sd, "transformer.decoder.layers.", memory, nhead, key_len)
    logits = []
    for t in range(max_len - 1):                                                             # FA:137
        if recompute:
            lg = fa_decode_logits(sd, memory, fed, nhead, key_len)[:, -1]                    # FA:138-143
        else:
            x = dec.step(F.embedding(fed[:, -1:], emb) + pe[t][None, None])
            if "transformer.decoder.norm.weight" in sd:
                x = layer_norm(sd, "transformer.decoder.norm.", x)
            lg = F.linear(x[:, 0], sd["output_linear.weight"], sd["output_linear.bias"])
        nxt = torch.argmax(lg, dim=1)                                                        # FA:144
        logits.append(lg)
        ys = torch.cat([ys, nxt[:, None]], dim=1)                                            # FA:145
        fed = torch.cat([fed, (forced[:, t] if forced is not None else nxt)[:, None]], dim=1)
    return ys, torch.stack(logits, dim=1)


def parse_chain(final_chain, rev_vocab):
    """run_inference_chain's parsing (FA:96-108): -> list of (func_token:int, [dependency step indices])."""
    steps = []
    for elem in final_chain:
        parts = elem.strip().split()
        deps = []
        for tok in parts[1:]:
            original = rev_vocab.get(int(tok), None)
            if original is not None and original.isdigit():
                deps.append(int(original))
        steps.append((int(parts[0]), deps))
    return steps


@torch.no_grad()
def fa_run_chain(sd, image_features, final_chain, rev_vocab, start_token, max_infer_len, nhead, forced=None,
                 recompute=False):
    """run_inference_chain (FA:83-124) for ONE question, cache kept as {step: [token ids]} (the reference keeps
    the same tokens space-joined in a string).  Returns (cache, logits {step: (max_len-1, V)}).
    `forced` {step: (max_len-1,) tokens}: teacher-forces the decode AND is what gets cached for that step."""
    cache, all_logits = {}, {}
    for i, (func, deps) in enumerate(parse_chain(final_chain, rev_vocab)):
        src = [func]
        for idx in deps:
            src += cache.get(idx, [])                                                        # FA:109-116, missing -> ""
        src_t = torch.tensor([src], dtype=torch.long)
        fz = None if forced is None else torch.as_tensor(forced[i], dtype=torch.long)[None]
        ys, lg = fa_greedy_decode(sd, image_features, src_t, start_token, max_infer_len, nhead, forced=fz,
                                  recompute=recompute)
        toks = ys[0].tolist()
        if fz is not None:
            toks = [start_token] + fz[0].tolist()
        cache[i] = toks                                                                      # FA:120-121 (all tokens)
        all_logits[i] = lg[0]
    return cache, all_logits


# ------------------------------------------------------------------------------------------------
# synthetic CLEVR-shaped workloads shared by the tests and bench.py
# ------------------------------------------------------------------------------------------------
def iqap_tally(answer_output, generated_programs, gt_answers, gt_programs):
    """The per-sample tally of inference_transformer_iqap_tally.run_inference (TALLY:317-344) as a plain loop:
    returns [both correct, answer only, program only, neither] and the predicted answers."""
    counts, preds = [0, 0, 0, 0], []
    for i in range(answer_output.shape[0]):
        _, predicted = torch.max(answer_output[i:i + 1], 1)                      # TALLY:317-318
        predicted = predicted.item()
        preds.append(predicted)
        answer_correct = predicted == int(gt_answers[i])                          # TALLY:327-330
        program_correct = generated_programs[i].tolist() == gt_programs[i].tolist()  # TALLY:333
        counts[(0 if program_correct else 1) if answer_correct else (2 if program_correct else 3)] += 1
    return counts, preds


# ------------------------------------------------------------------------------------------------
# Evaluation of the bounding-box variant: IoU utilities (BB:104-150) and the bookkeeping of `evaluate` (BB:423-538)
# ------------------------------------------------------------------------------------------------
def bb_iou(pred_box, gt_box):
    """Scalar IoU of two [x_min, y_min, x_max, y_max] lists (BB:104-124)."""
    ix0, iy0 = max(pred_box[0], gt_box[0]), max(pred_box[1], gt_box[1])             # BB:109-110
    ix1, iy1 = min(pred_box[2], gt_box[2]), min(pred_box[3], gt_box[3])             # BB:111-112
    inter = max(ix1 - ix0, 0.0) * max(iy1 - iy0, 0.0)                               # BB:114-116
    pa = max(pred_box[2] - pred_box[0], 0.0) * max(pred_box[3] - pred_box[1], 0.0)  # BB:118
    ga = max(gt_box[2] - gt_box[0], 0.0) * max(gt_box[3] - gt_box[1], 0.0)          # BB:119
    union = pa + ga - inter                                                         # BB:121
    return 0.0 if union <= 0.0 else inter / union                                   # BB:122-124


def bb_batch_mean_iou(bbox_preds, bbox_gts):
    """Mean IoU over the ground-truth boxes of a (B, 10, 4) batch that are not all-zero padding; predictions clamped
    to [0, 1]; 0.0 when nothing is left (BB:126-150).  Plain loops, as the reference."""
    ious = []
    for i in range(bbox_preds.shape[0]):
        for j in range(bbox_preds.shape[1]):
            gt = [float(x) for x in bbox_gts[i, j]]
            if all(abs(x) < 1e-8 for x in gt):                                      # BB:139-140
                continue
            pred = [min(max(float(c), 0.0), 1.0) for c in bbox_preds[i, j]]         # BB:142-144
            ious.append(bb_iou(pred, gt))
    return float(sum(ious) / len(ious)) if ious else 0.0                            # BB:148-150


def bb_evaluate(batches):
    """The bookkeeping of `evaluate` (BB:423-538) over `batches` = iterable of (seq_logits (B, 28, V), bbox_preds
    (B, 10, 4), combined_seq (B, 28), bboxes_gt (B, 10, 4)) with CrossEntropyLoss / masked SmoothL1Loss (BB:615-616):
    -> (loss, mean IoU, cpca, cpia, ipca, ipia, program token accuracy)."""
    running, total, iou_cum, iou_batches = 0.0, 0, 0.0, 0
    counts = [0, 0, 0, 0]
    tok_ok, tok_n = 0, 0
    for seq_logits, bbox_preds, combined_seq, bboxes_gt in batches:
        n = seq_logits.shape[0]
        total += n
        loss_seq = F.cross_entropy(seq_logits.reshape(-1, seq_logits.shape[-1]), combined_seq.reshape(-1))  # BB:466-468
        mask = bboxes_gt.sum(dim=2, keepdim=True) > 0                                                       # BB:471
        loss_bbox = (F.smooth_l1_loss(bbox_preds, bboxes_gt, reduction="none") * mask).sum() / mask.sum()   # BB:472-474
        running += (loss_seq + loss_bbox).item() * n                                                        # BB:477-478
        iou_cum += bb_batch_mean_iou(bbox_preds, bboxes_gt)                                                 # BB:481-483
        iou_batches += 1
        _, pred = torch.max(seq_logits, dim=2)                                                              # BB:487
        for i in range(n):
            prog = pred[i, :-1].tolist() == combined_seq[i, :-1].tolist()                                   # BB:490-494
            ans = int(pred[i, -1]) == int(combined_seq[i, -1])                                              # BB:503
            counts[(0 if ans else 1) if prog else (2 if ans else 3)] += 1                                   # BB:506-517
            tok_ok += sum(int(a == b) for a, b in zip(pred[i, :-1].tolist(), combined_seq[i, :-1].tolist()))
            tok_n += pred.shape[1] - 1
    return (running / total, iou_cum / iou_batches if iou_batches else 0.0, counts[0] / total, counts[1] / total,
            counts[2] / total, counts[3] / total, tok_ok / float(tok_n) if tok_n else 0.0)

# the seeded input generators live in the product package's neutral module (no model arithmetic there)
from explainable_spatial_vqa_b200.synthetic import chain_strings, fa_programs, fa_vocab, iqap_inputs  # noqa: E402,F401
