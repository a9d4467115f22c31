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


def questions(B, seed=4242):
    g = torch.Generator().manual_seed(seed)
    q = torch.zeros(B, 46, dtype=torch.long)
    for b in range(B):
        n = int(torch.randint(8, 47, (1,), generator=g))
        q[b, 0] = 1
        q[b, 1:n - 1] = torch.randint(4, 85, (n - 2,), generator=g)
        q[b, n - 1] = 2
    return q


# ---- synthetic CLEVR-shaped PREFIX programs in the generator's token space (for the glue tests and bench config 4)
# token ids: 0 <NULL>, 1 <START>, 2 <END>, 3 <UNK> (all end a program), 4 = scene (no input), 5..9 binary functions
# (equal_* / union / intersect / less_than / greater_than in the reference's get_num_inputs), 10..43 unary.
def program_arity(prog_vocab=44):
    a = torch.ones(prog_vocab, dtype=torch.int32)
    a[:4] = -1
    a[4] = 0
    a[5:10] = 2
    return a


def program_func_map(prog_vocab=44):
    return (27 + torch.arange(prog_vocab, dtype=torch.int32) % 40).to(torch.int32)  # FA function ids 27..66


def prefix_programs(B, T=27, seed=777, max_nodes=25):
    """-> (programs (B,T) i64 padded with <END> then <NULL>, node counts (B,))."""
    g = torch.Generator().manual_seed(seed)
    out = torch.zeros(B, T, dtype=torch.long)
    counts = torch.zeros(B, dtype=torch.long)
    for b in range(B):
        budget = int(torch.randint(2, max_nodes + 1, (1,), generator=g))

        def build(n):  # a subtree with exactly n nodes, prefix order
            if n == 1:
                return [4]
            if n >= 3 and float(torch.rand(1, generator=g)) < 0.2:
                left = int(torch.randint(1, n - 1, (1,), generator=g))
                return [int(torch.randint(5, 10, (1,), generator=g))] + build(left) + build(n - 1 - left)
            return [int(torch.randint(10, 44, (1,), generator=g))] + build(n - 1)

        toks = build(budget)
        counts[b] = len(toks)
        out[b, : len(toks)] = torch.tensor(toks)
        if len(toks) < T:
            out[b, len(toks)] = 2
    return out, counts
