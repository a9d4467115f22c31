"""Drop-in for the model half of the reference's `code/run_model_lstm_qp.py`: the LSTM program generator
(question tokens -> 27 program tokens), SURVEY §8(f) next-1.

`Seq2SeqModel` keeps the reference's constructor, state-dict (`embedding`, `encoder`, `decoder`, `fc`) and
`forward(questions, program_targets=None)` semantics (reference lines 277-319); the arithmetic runs in
libb200vqa.so (`b200vqa_lstm_*`, csrc/lstm_api.cu): one tensor-core GEMM with a fused LSTM-cell epilogue per time
step, a tf32 head GEMM with fused argmax per decoder step, everything on the device, no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from . import _native as nat

__all__ = ["Seq2SeqModel", "get_data_info", "prefix_program_to_deps", "programs_to_chain"]


class Seq2SeqModel(nn.Module):
    def __init__(self, vocab_size, embedding_dim, lstm_hidden_dim, program_vocab_size, program_seq_len,
                 program_start_token_idx):
        super().__init__()
        # parameter containers in the reference's construction order (identical seeded init and state-dict keys)
        self.embedding = nn.Embedding(vocab_size, embedding_dim, padding_idx=0)
        self.encoder = nn.LSTM(embedding_dim, lstm_hidden_dim, batch_first=True)
        self.decoder = nn.LSTM(embedding_dim, lstm_hidden_dim, batch_first=True)
        self.fc = nn.Linear(lstm_hidden_dim, program_vocab_size)
        self.program_seq_len = program_seq_len
        self.program_vocab_size = program_vocab_size
        self.program_start_token_idx = program_start_token_idx
        self.relu = nn.ReLU()
        self.dropout = nn.Dropout(0.5)
        self._h = None
        self._version = None
        self._keep = None

    def _native(self):
        version = nat.weights_version(self)
        dev = self.embedding.weight.device
        if self._h is not None and version == self._version:
            return self._h
        if self._h is not None:
            nat.lib().b200vqa_lstm_destroy(self._h)
            self._h = None
        if dev.type != "cuda":
            raise nat.NativeError(f"parameters live on {dev}: the LSTM generator runs on an sm_100 GPU only")
        keep = []
        d = nat.LstmDesc()
        d.vocab = self.embedding.num_embeddings
        d.embedding_dim = self.embedding.embedding_dim
        d.hidden_dim = self.encoder.hidden_size
        d.prog_vocab = self.fc.out_features
        nat._set(d, keep, embedding=self.embedding.weight,
                 enc_w_ih=self.encoder.weight_ih_l0, enc_w_hh=self.encoder.weight_hh_l0,
                 enc_b_ih=self.encoder.bias_ih_l0, enc_b_hh=self.encoder.bias_hh_l0,
                 dec_w_ih=self.decoder.weight_ih_l0, dec_w_hh=self.decoder.weight_hh_l0,
                 dec_b_ih=self.decoder.bias_ih_l0, dec_b_hh=self.decoder.bias_hh_l0,
                 fc_w=self.fc.weight, fc_b=self.fc.bias)
        out = C.c_void_p(0)
        with torch.cuda.device(dev):
            torch.cuda.current_stream().synchronize()
            nat.check(nat.lib().b200vqa_lstm_create(C.byref(d), dev.index or 0, C.byref(out)), "b200vqa_lstm_create")
        self._h, self._version, self._keep = out, version, keep
        return self._h

    def __del__(self):
        try:
            if self._h is not None:
                nat.lib().b200vqa_lstm_destroy(self._h)
        except Exception:
            pass

    @torch.no_grad()
    def generate(self, questions, forced=None, want_logits=False):
        """questions (B, L) i64 -> greedy programs (B, program_seq_len) i64 [, logits (B, T, Vp)].
        `forced` (B, T): decoder step t+1 consumes forced[:, t] instead of its own argmax."""
        h = self._native()
        if not questions.is_cuda:
            raise nat.NativeError("questions must be a CUDA tensor (no CPU path)")
        q = questions.to(torch.int64).contiguous()
        B, L = q.shape
        T = self.program_seq_len
        programs = torch.empty(B, T, dtype=torch.int64, device=q.device)
        logits = torch.empty(B, T, self.program_vocab_size, dtype=torch.float32, device=q.device) if want_logits else None
        fz = None if forced is None else forced.to(q.device, torch.int64).contiguous()
        with torch.cuda.device(q.device):
            nat.check(nat.lib().b200vqa_lstm_generate(h, nat.ptr(q), L, B, T, int(self.program_start_token_idx),
                                                      nat.ptr(programs), nat.ptr(logits), nat.ptr(fz),
                                                      nat.stream_ptr(q.device)), "b200vqa_lstm_generate")
        return (programs, logits) if want_logits else programs

    def forward(self, questions, program_targets=None):
        """Inference: (B, 27) greedy tokens.  With `program_targets` (B, T): the teacher-forced logits (B, T, Vp) of
        the reference's training branch - the decoder consumes program_targets[:, t] at step t (column 0 is the
        start token and must be the same for the whole batch)."""
        if program_targets is None:
            return self.generate(questions)
        tgt = program_targets.to(torch.int64)
        if tgt.shape[1] != self.program_seq_len:
            raise ValueError(f"program_targets must have {self.program_seq_len} columns")
        if bool((tgt[:, 0] != tgt[0, 0]).any()):
            raise ValueError("program_targets[:, 0] (the start token) must be the same for the whole batch")
        saved = self.program_start_token_idx
        try:
            self.program_start_token_idx = int(tgt[0, 0])
            # step t+1 consumes forced[:, t] = targets[:, t+1]; the last column is never consumed
            forced = torch.cat([tgt[:, 1:], tgt[:, -1:]], dim=1).contiguous()
            _, logits = self.generate(questions, forced=forced, want_logits=True)
        finally:
            self.program_start_token_idx = saved
        return logits


def get_data_info(questions_h5_path):
    """(question vocab, program vocab) = max id + 1 over the H5 arrays (reference lines 322-329)."""
    import h5py
    with h5py.File(questions_h5_path, "r") as f:
        return int(np.max(f["questions"])) + 1, int(np.max(f["programs"])) + 1


def prefix_program_to_deps(arity):
    """Program -> chain glue (reference preprocess_questions/utils_programs.py:100-156, `prefix_to_list`): for a
    program given in PREFIX order as per-token input counts (`get_num_inputs`: scene 0, equal_* / union / intersect /
    less_than / greater_than 2, everything else 1) returns (order, deps): `order[i]` = prefix position of the node
    executed at step i (inputs before consumers, as `tree_to_list` numbers them) and `deps[i]` = the step indices
    feeding step i (-1 = none) - the `deps[B,S,2]` layout of `run_inference_chain_batched`."""
    arity = [int(a) for a in arity]
    pos = 0

    def parse():
        nonlocal pos
        me = pos
        pos += 1
        return (me, [parse() for _ in range(arity[me])])

    tree = parse()

    def count(node):
        return 1 + sum(count(ch) for ch in node[1])

    n = count(tree)
    order = [-1] * n
    deps = [[-1, -1] for _ in range(n)]

    def place(node, idx):
        order[idx] = node[0]
        nxt = idx - 1
        ins = []
        for ch in reversed(node[1]):
            ins.insert(0, nxt)
            nxt = place(ch, nxt)
        for k, v in enumerate(ins[:2]):
            deps[idx][k] = v
        return nxt

    place(tree, n - 1)
    return order, deps


@torch.no_grad()
def programs_to_chain(programs, arity, func_map, max_steps=25):
    """Device version of `prefix_program_to_deps` for a batch: programs (B, T) i64 CUDA tensor in prefix order,
    arity / func_map (Vp,) i32 -> (func (B,S) i32, deps (B,S,2) i32, n_steps (B,) i32) ready for
    `run_inference_chain_batched`.  No host round trip between the generator and the executor."""
    if not programs.is_cuda:
        raise nat.NativeError("programs must be a CUDA tensor (no CPU path)")
    dev = programs.device
    prog = programs.to(torch.int64).contiguous()
    ar = arity.to(dev, torch.int32).contiguous()
    fm = func_map.to(dev, torch.int32).contiguous()
    B, T = prog.shape
    func = torch.empty(B, max_steps, dtype=torch.int32, device=dev)
    deps = torch.empty(B, max_steps, 2, dtype=torch.int32, device=dev)
    n_steps = torch.empty(B, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        nat.check(nat.lib().b200vqa_programs_to_chain(nat.ptr(prog), B, T, nat.ptr(ar), nat.ptr(fm), ar.numel(),
                                                      int(max_steps), nat.ptr(func), nat.ptr(deps), nat.ptr(n_steps),
                                                      nat.stream_ptr(dev)), "b200vqa_programs_to_chain")
    return func, deps, n_steps
