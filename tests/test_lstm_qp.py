"""LSTM program generator: CPU tests of the oracle against the reference's golden outputs and
of the program->chain glue; GPU parity of the tcgen05 path through the C ABI."""
import numpy as np
import pytest
import torch

import common
from explainable_spatial_vqa_b200 import run_model_lstm_qp as qp
from oracle import lstm_oracle


def seeded(seed=0):
    torch.manual_seed(seed)
    return qp.Seq2SeqModel(85, 256, 512, 44, 27, 1).eval()


def test_state_dict_matches_reference_layout():
    g = common.load_golden("lstm_qp.npz")
    sd = seeded().state_dict()
    assert sorted(sd) == [str(k) for k in g["sd_keys"]]
    assert tuple(sd["encoder.weight_hh_l0"].shape) == (2048, 512) and tuple(sd["fc.weight"].shape) == (44, 512)
    assert float(sd["embedding.weight"][0].abs().sum()) == 0.0


def test_oracle_matches_reference_golden():
    g = common.load_golden("lstm_qp.npz")
    sd = seeded().state_dict()
    if not common.weights_match_golden(sd, g):
        pytest.skip("seeded init differs from the golden run")
    q = lstm_oracle.questions(6, seed=4242)
    assert np.array_equal(q.numpy(), g["questions"])
    prog, logits = lstm_oracle.generate(sd, q)
    assert np.array_equal(prog.numpy(), g["programs"])
    assert common.rel_err(logits, g["logits"]) < 1e-4
    # teacher forcing with the greedy tokens reproduces the same logits (and the reference's training branch)
    _, tf = lstm_oracle.generate(sd, q, forced=torch.from_numpy(g["programs"]))
    assert common.rel_err(tf, g["tf_logits"]) < 1e-4


def test_prefix_program_to_deps_follows_reference_numbering():
    # equal_color(query_color(unique(filter(scene))), query_color(unique(scene)))  in prefix order
    arity = [2, 1, 1, 1, 0, 1, 1, 0]
    order, deps = qp.prefix_program_to_deps(arity)
    assert order == [4, 3, 2, 1, 7, 6, 5, 0]          # inputs before consumers, root last (tree_to_list)
    assert deps == [[-1, -1], [0, -1], [1, -1], [2, -1], [-1, -1], [4, -1], [5, -1], [3, 6]]
    order, deps = qp.prefix_program_to_deps([1, 1, 0])
    assert order == [2, 1, 0] and deps == [[-1, -1], [0, -1], [1, -1]]


@pytest.mark.gpu
def test_gpu_teacher_forced_logits_and_tokens():
    m = seeded()
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    m = m.cuda()
    q = lstm_oracle.questions(64, seed=7)
    ref_prog, ref_logits = lstm_oracle.generate(sd, q)
    prog, logits = m.generate(q.cuda(), forced=ref_prog.cuda(), want_logits=True)
    assert common.rel_err(logits, ref_logits) < common.LOGIT_REL_TOL
    common.check_tokens_where_decisive(prog, ref_prog, ref_logits, logits, "lstm programs")
    g = common.load_golden("lstm_qp.npz")
    if common.weights_match_golden(sd, g):
        gq = torch.from_numpy(g["questions"]).cuda()
        p2, l2 = m.generate(gq, forced=torch.from_numpy(g["programs"]).cuda(), want_logits=True)
        assert common.rel_err(l2, g["logits"]) < common.LOGIT_REL_TOL
        tgt = torch.cat([torch.ones(6, 1, dtype=torch.long), torch.from_numpy(g["programs"])[:, :-1]], dim=1).cuda()
        assert common.rel_err(m(gq, tgt), g["tf_logits"]) < common.LOGIT_REL_TOL


@pytest.mark.gpu
def test_gpu_free_running_graph_path_is_deterministic_and_split_invariant():
    m = seeded().cuda()
    q = lstm_oracle.questions(700, seed=9).cuda()
    p1 = m(q)
    p2 = m(q)
    assert p1.shape == (700, 27) and p1.dtype == torch.int64 and torch.equal(p1, p2)
    assert torch.equal(torch.cat([m(q[:300]), m(q[300:])]), p1)
    eager, _ = m.generate(q, want_logits=True)  # eager (logit-returning) path == graph-replayed path
    assert torch.equal(eager, p1)
    assert int(p1.min()) >= 0 and int(p1.max()) < 44


@pytest.mark.gpu
def test_gpu_decisive_head_free_running_exact():
    m = seeded()
    with torch.no_grad():
        m.fc.weight.mul_(8.0)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    m = m.cuda()
    q = lstm_oracle.questions(16, seed=3)
    ref_prog, ref_logits = lstm_oracle.generate(sd, q)
    prog = m(q.cuda()).cpu()
    mg = common.margins(ref_logits)
    exact = 0
    for b in range(16):
        t_end = next((t for t in range(27) if float(mg[b, t]) < 0.05), 27)  # up to the first near-tie
        assert torch.equal(prog[b, :t_end], ref_prog[b, :t_end]), b
        exact += t_end
    assert exact >= 27, exact


@pytest.mark.gpu
def test_gpu_programs_to_chain_matches_host_glue():
    progs, counts = lstm_oracle.prefix_programs(300, seed=5)
    arity, fmap = lstm_oracle.program_arity(), lstm_oracle.program_func_map()
    func, deps, n_steps = qp.programs_to_chain(progs.cuda(), arity, fmap, max_steps=25)
    func, deps, n_steps = func.cpu(), deps.cpu(), n_steps.cpu()
    assert torch.equal(n_steps.long(), counts)
    for b in range(300):
        n = int(counts[b])
        toks = progs[b, :n].tolist()
        order, ref_deps = qp.prefix_program_to_deps([int(arity[t]) for t in toks])
        assert func[b, :n].tolist() == [int(fmap[toks[p]]) for p in order], b
        assert deps[b, :n].tolist() == ref_deps, b
        assert bool((deps[b, n:] == -1).all())
        for i in range(n):
            assert all(int(d) < i for d in deps[b, i]), (b, i)  # inputs always precede their consumer
    # malformed programs never write out of bounds: no terminator / unknown ids / more nodes than max_steps
    bad = torch.randint(0, 60, (64, 27))
    f2, d2, n2 = qp.programs_to_chain(bad.cuda(), arity, fmap, max_steps=8)
    assert int(n2.max()) <= 8 and int(n2.min()) >= 0 and int(d2.max()) < 8
