"""Seeded synthetic CLEVR-shaped workloads shared by bench.py, the tools and the test-suite.

Pure input generators: no model arithmetic lives here.  They live in the product package so that the product arm of bench.py needs
nothing from the CPU checker; the checker's modules re-export these names for the tests.

  iqap_inputs      (B,196,1024) conv4-like features (post-ReLU) + (B,46) questions [<START>, tokens.., <END>, pad]
  fa_programs      CLEVR-shaped program DAGs in the chain format of the step-wise executor (func / deps / n_steps)
  fa_vocab, chain_strings   the reference's `final_chain_of_thought` string form of those arrays (FA:99-108)
  lstm_questions, prefix_programs, program_arity, program_func_map   inputs of the program generator + glue (config 4)
"""
from __future__ import annotations

import torch


def iqap_inputs(B, seed=1234, relu=True):
    g = torch.Generator().manual_seed(seed)
    img = torch.randn(B, 196, 1024, generator=g)
    if relu:
        img.relu_()  # conv4 features are post-ReLU
    q = torch.zeros(B, 46, dtype=torch.long)
    for b in range(B):
        n = int(torch.randint(8, 47, (1,), generator=g))
        q[b, 0] = 1
        q[b, 1:n - 1] = torch.randint(4, 85, (n - 2,), generator=g)
        q[b, n - 1] = 2
    return img, q


def fa_vocab(V=170):
    """ids 0..26 are the digit strings "0".."26" (dependency pointers), the rest are opaque tokens."""
    return {i: (str(i) if i < 27 else f"tok{i}") for i in range(V)}


def fa_programs(B, seed=4321, max_steps=25, V=170):
    """CLEVR-shaped DAGs: n_steps~U[2,25]; step 0 is a root; later steps: 15 % new root, 75 % unary on the
    previous step, 10 % binary on two distinct earlier steps.  -> func (B,S) i32, deps (B,S,2) i32, n_steps (B,)"""
    g = torch.Generator().manual_seed(seed)
    func = torch.zeros(B, max_steps, dtype=torch.int32)
    deps = torch.full((B, max_steps, 2), -1, dtype=torch.int32)
    n_steps = torch.randint(2, max_steps + 1, (B,), generator=g, dtype=torch.int32)
    for b in range(B):
        for i in range(int(n_steps[b])):
            func[b, i] = int(torch.randint(27, min(67, V), (1,), generator=g))
            if i == 0:
                continue
            r = float(torch.rand(1, generator=g))
            if r < 0.15:
                continue
            if r < 0.90 or i < 2:
                deps[b, i, 0] = i - 1
            else:
                a = int(torch.randint(0, i - 1, (1,), generator=g))
                deps[b, i, 0] = a
                deps[b, i, 1] = i - 1
    return func, deps, n_steps


def chain_strings(func_row, deps_row, n):
    """arrays -> the reference's `final_chain_of_thought` strings (dependency k is written as vocab id k)."""
    out = []
    for i in range(int(n)):
        toks = [str(int(func_row[i]))] + [str(int(d)) for d in deps_row[i] if int(d) >= 0]
        out.append(" ".join(toks))
    return out


def lstm_questions(B, seed=4242):
    g = torch.Generator().manual_seed(seed)
    q = torch.zeros(B, 46, dtype=torch.long)
    for b in range(B):
        n = int(torch.randint(8, 47, (1,), generator=g))
        q[b, 0] = 1
        q[b, 1:n - 1] = torch.randint(4, 85, (n - 2,), generator=g)
        q[b, n - 1] = 2
    return q


# token ids of the generator's program vocabulary: 0 <NULL>, 1 <START>, 2 <END>, 3 <UNK> (all end a program), 4 = scene
# (no input), 5..9 binary functions, 10..43 unary (the reference's get_num_inputs, utils_programs.py:100-110)
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
