"""CPU restatement (plain fp32 torch ops) of the reference's LSTM program generator, Seq2SeqModel.forward in
/root/reference/code/run_model_lstm_qp.py:291-319.  TEST INFRASTRUCTURE - see oracle/__init__.py.

torch.nn.LSTM semantics (gate order i|f|g|o):  gates = x W_ih^T + b_ih + h W_hh^T + b_hh;
c' = sigmoid(f) c + sigmoid(i) tanh(g);  h' = sigmoid(o) tanh(c').  Pinned by tests/golden/lstm_qp.npz (outputs of the
reference itself, oracle/make_golden.py)."""
import torch
import torch.nn.functional as F


def lstm_cell(sd, prefix, x, h, c):
    gates = F.linear(x, sd[prefix + "weight_ih_l0"], sd[prefix + "bias_ih_l0"]) + \
        F.linear(h, sd[prefix + "weight_hh_l0"], sd[prefix + "bias_hh_l0"])
    i, f, g, o = gates.chunk(4, dim=-1)
    c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
    return torch.sigmoid(o) * torch.tanh(c), c


@torch.no_grad()
def generate(sd, questions, program_seq_len=27, start_token=1, forced=None):
    """-> (programs (B,T) i64, logits (B,T,Vp)).  forced (B,T): step t+1 consumes forced[:, t]."""
    emb = sd["embedding.weight"]
    B = questions.shape[0]
    H = sd["encoder.weight_hh_l0"].shape[1]
    h = torch.zeros(B, H)
    c = torch.zeros(B, H)
    for t in range(questions.shape[1]):                                   # :294  encoder over every position
        h, c = lstm_cell(sd, "encoder.", F.embedding(questions[:, t], emb), h, c)
    tok = torch.full((B,), start_token, dtype=torch.long)                 # :305
    toks, logits = [], []
    for t in range(program_seq_len):                                      # :311-317
        h, c = lstm_cell(sd, "decoder.", F.embedding(tok, emb), h, c)
        lg = F.linear(h, sd["fc.weight"], sd["fc.bias"])
        nxt = torch.max(lg, dim=1)[1]
        toks.append(nxt)
        logits.append(lg)
        tok = forced[:, t] if forced is not None else nxt
    return torch.stack(toks, 1), torch.stack(logits, 1)


# seeded input generators: the product package's neutral module (no model arithmetic there)
from explainable_spatial_vqa_b200.synthetic import (lstm_questions as questions, prefix_programs, program_arity,  # noqa: E402,F401
                                                    program_func_map)
