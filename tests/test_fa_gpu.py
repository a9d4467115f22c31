"""GPU parity of the FA step-wise executor (greedy_decode / forward / run_inference_chain with the HBM
cache) against the CPU oracle and the reference's golden outputs."""
import numpy as np
import pytest
import torch

import common
from explainable_spatial_vqa_b200 import inference_transformer_full_annotation_new as fa
from oracle import executor_oracle as orc

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda")


@pytest.fixture(scope="module")
def model():
    return common.seeded_fa().cuda()


def cpu_sd(m):
    return {k: v.detach().cpu() for k, v in m.state_dict().items()}


def golden_img():
    gen = torch.Generator().manual_seed(77)
    return torch.randn(1, 1024, 14, 14, generator=gen).relu_()


def test_greedy_decode_golden_teacher_forced(model):
    g = common.load_golden("fa_nhead2.npz")
    if not common.weights_match_golden(cpu_sd(model), g):
        pytest.skip("seeded init differs from the golden run")
    img = golden_img()
    for s in (1, 21, 41):
        src = torch.from_numpy(g[f"gd_src_{s}"])
        ref_ys = torch.from_numpy(g[f"gd_ys_{s}"])
        ys, logits = fa.greedy_decode(model, img, src, 0, 20, DEV, forced=ref_ys[:, 1:], want_logits=True)
        assert int(ys[0, 0]) == 0
        assert common.rel_err(logits, g[f"gd_logits_{s}"]) < common.LOGIT_REL_TOL, s
        common.check_tokens_where_decisive(ys[:, 1:], ref_ys[:, 1:], g[f"gd_logits_{s}"], logits, f"fa gd s={s}")


def test_batched_ragged_src_matches_oracle(model):
    """One batch with src lengths 1/21/41/11 (padded, masked by src_len) == per-question oracle runs."""
    sd = cpu_sd(model)
    gen = torch.Generator().manual_seed(5)
    B = 4
    img = torch.randn(B, 1024, 14, 14, generator=gen).relu_()
    lens = [1, 21, 41, 11]
    src = torch.zeros(B, 41, dtype=torch.long)
    for b, n in enumerate(lens):
        src[b, :n] = torch.randint(0, 170, (n,), generator=gen)
    src_len = torch.tensor(lens, dtype=torch.int32)
    ref_ys, ref_lg = orc.fa_greedy_decode(sd, img, src, 0, 20, 2, src_len=torch.tensor(lens))
    ys, lg = fa.greedy_decode(model, img, src, 0, 20, DEV, src_len=src_len, forced=ref_ys[:, 1:], want_logits=True)
    assert common.rel_err(lg, ref_lg) < common.LOGIT_REL_TOL
    common.check_tokens_where_decisive(ys[:, 1:], ref_ys[:, 1:], ref_lg, lg, "fa ragged")


def test_forward_teacher_forced_matches_golden(model):
    g = common.load_golden("fa_nhead2.npz")
    if not common.weights_match_golden(cpu_sd(model), g):
        pytest.skip("seeded init differs from the golden run")
    gen = torch.Generator().manual_seed(77)
    torch.randn(1, 1024, 14, 14, generator=gen)
    for s in (1, 21, 41):
        torch.randint(0, 170, (1, s), generator=gen)
    img2 = torch.randn(2, 1024, 14, 14, generator=gen).relu_()
    logits = model(img2.cuda(), torch.from_numpy(g["fw_src"]).cuda(), torch.from_numpy(g["fw_tgt"]).cuda())
    assert logits.shape == (2, 20, 170)
    assert common.rel_err(logits, g["fw_logits"]) < common.LOGIT_REL_TOL


@pytest.mark.parametrize("switch", ["B200VQA_NO_FUSED_ENC_FFN", "B200VQA_NO_FUSED_FINAL_LN"])
def test_forward_golden_with_the_unfused_encoder_kernels(monkeypatch, switch):
    """The encoder feed-forward block as two GEMMs through HBM / the final encoder norm as its own kernel: the same
    golden gate as the fused default."""
    monkeypatch.setenv(switch, "1")
    m = common.seeded_fa().cuda()
    g = common.load_golden("fa_nhead2.npz")
    if not common.weights_match_golden(cpu_sd(m), g):
        pytest.skip("seeded init differs from the golden run")
    gen = torch.Generator().manual_seed(77)
    torch.randn(1, 1024, 14, 14, generator=gen)
    for s in (1, 21, 41):
        torch.randint(0, 170, (1, s), generator=gen)
    img2 = torch.randn(2, 1024, 14, 14, generator=gen).relu_()
    logits = m(img2.cuda(), torch.from_numpy(g["fw_src"]).cuda(), torch.from_numpy(g["fw_tgt"]).cuda())
    assert common.rel_err(logits, g["fw_logits"]) < common.LOGIT_REL_TOL


def test_chain_golden_teacher_forced_cache(model):
    """run_inference_chain with the HBM cache, fed the reference's tokens: every step's logits match and the
    cache holds exactly the reference's cache (dependency pointers gather the right rows)."""
    g = common.load_golden("fa_nhead2.npz")
    if not common.weights_match_golden(cpu_sd(model), g):
        pytest.skip("seeded init differs from the golden run")
    func = torch.from_numpy(g["chain_func"])[None]
    deps = torch.from_numpy(g["chain_deps"])[None]
    S = func.shape[1]
    ref_cache = torch.from_numpy(g["chain_cache"])
    forced = ref_cache[None, :, 1:].long()
    cache, logits = fa.run_inference_chain_batched(model, golden_img(), func, deps, torch.tensor([S], dtype=torch.int32),
                                                   0, 20, forced=forced, want_logits=True)
    assert torch.equal(cache[0].cpu(), ref_cache)
    assert common.rel_err(logits[0], g["chain_logits"]) < common.LOGIT_REL_TOL
    common.check_tokens_where_decisive(logits[0].argmax(-1), ref_cache[:, 1:], g["chain_logits"], logits[0], "fa chain")


def test_chain_batched_ragged_programs_match_oracle(model):
    """B ragged CLEVR-shaped programs at once (sorted, finished questions dropped) == per-question oracle
    chains; rows beyond n_steps stay -1."""
    sd = cpu_sd(model)
    B = 6
    func, deps, n_steps = orc.fa_programs(B, seed=4321, max_steps=9)
    gen = torch.Generator().manual_seed(8)
    img = torch.randn(B, 1024, 14, 14, generator=gen).relu_()
    rev = orc.fa_vocab(170)
    S = func.shape[1]
    ref_cache = torch.full((B, S, 20), -1, dtype=torch.int32)
    ref_logits = torch.zeros(B, S, 19, 170)
    for b in range(B):
        chain = orc.chain_strings(func[b], deps[b], n_steps[b])
        c, lg = orc.fa_run_chain(sd, img[b:b + 1], chain, rev, 0, 20, 2)
        for i in range(int(n_steps[b])):
            ref_cache[b, i] = torch.tensor(c[i], dtype=torch.int32)
            ref_logits[b, i] = lg[i]
    forced = ref_cache[:, :, 1:].clamp_min(0).long()
    cache, logits = fa.run_inference_chain_batched(model, img, func, deps, n_steps, 0, 20, forced=forced, want_logits=True)
    cache = cache.cpu()
    for b in range(B):
        n = int(n_steps[b])
        assert torch.equal(cache[b, :n], ref_cache[b, :n]), b
        assert bool((cache[b, n:] == -1).all()), b
        assert common.rel_err(logits[b, :n], ref_logits[b, :n]) < common.LOGIT_REL_TOL, b
    # unsorted execution path gives the same answer
    cache2 = fa.run_inference_chain_batched(model, img, func, deps, n_steps, 0, 20, forced=forced, sort_by_steps=False)
    assert torch.equal(cache2.cpu(), cache)


def test_chain_free_running_decisive_weights():
    """Sharpened head (x8): free-running chains reproduce the oracle's cache exactly up to each question's
    first non-decisive decision (a flipped token propagates through the cache)."""
    m = common.seeded_fa()
    with torch.no_grad():
        m.output_linear.weight.mul_(8.0)
    sd = cpu_sd(m)
    m = m.cuda()
    B = 4
    func, deps, n_steps = orc.fa_programs(B, seed=11, max_steps=6)
    gen = torch.Generator().manual_seed(2)
    img = torch.randn(B, 1024, 14, 14, generator=gen).relu_()
    cache = fa.run_inference_chain_batched(m, img, func, deps, n_steps, 0, 20).cpu()
    rev = orc.fa_vocab(170)
    exact = 0
    for b in range(B):
        chain = orc.chain_strings(func[b], deps[b], n_steps[b])
        c, lg = orc.fa_run_chain(sd, img[b:b + 1], chain, rev, 0, 20, 2)
        diverged = False
        for i in range(int(n_steps[b])):
            mg = common.margins(lg[i])
            assert int(cache[b, i, 0]) == 0
            for t in range(19):
                if float(mg[t]) < 0.05:
                    diverged = True  # a near-tie: bf16 may legitimately flip it, later inputs then differ
                    break
                assert int(cache[b, i, t + 1]) == c[i][t + 1], (b, i, t)
                exact += 1
            if diverged:
                break
    assert exact >= 19, exact


def test_reference_compatible_run_inference_chain(model):
    """String-cache surface: dict[int -> 'tok tok ...'] with 20 tokens per executed step."""
    func = [30, 31, 32, 30, 40]
    deps = [[], [0], [1], [], [2, 3]]
    chain = [" ".join([str(f)] + [str(d) for d in dd]) for f, dd in zip(func, deps)]
    final, cache = fa.run_inference_chain(model, golden_img(), chain, DEV, 0, 20, orc.fa_vocab(170))
    assert sorted(cache) == [0, 1, 2, 3, 4]
    assert all(len(v.split()) == 20 and v.split()[0] == "0" for v in cache.values())
    assert final == cache[4]


def test_missing_dependency_contributes_nothing(model):
    """A pointer to a step that has not run (>= current index) is the reference's `cache.get(idx, "")`."""
    img = golden_img()
    func = torch.tensor([[30, 31]], dtype=torch.int32)
    deps_bad = torch.tensor([[[-1, -1], [5, -1]]], dtype=torch.int32)
    deps_none = torch.tensor([[[-1, -1], [-1, -1]]], dtype=torch.int32)
    n = torch.tensor([2], dtype=torch.int32)
    a = fa.run_inference_chain_batched(model, img, func, deps_bad, n)
    b = fa.run_inference_chain_batched(model, img, func, deps_none, n)
    assert torch.equal(a, b)


def test_nhead4_two_layers_long_text():
    """Training-time shape of the reference (nhead 4, 2+2 layers) and the L=260 sweep point (max_text_len 64)."""
    m = common.seeded_fa(seed=3, nhead=4, enc_layers=2, dec_layers=2, max_text_len=64)
    sd = cpu_sd(m)
    m = m.cuda()
    gen = torch.Generator().manual_seed(4)
    img = torch.randn(2, 1024, 14, 14, generator=gen).relu_()
    src = torch.randint(0, 170, (2, 60), generator=gen)
    src_len = torch.tensor([60, 33], dtype=torch.int32)
    ref_ys, ref_lg = orc.fa_greedy_decode(sd, img, src, 0, 20, 4, src_len=src_len.long())
    ys, lg = fa.greedy_decode(m, img, src, 0, 20, DEV, src_len=src_len, forced=ref_ys[:, 1:], want_logits=True)
    assert common.rel_err(lg, ref_lg) < common.LOGIT_REL_TOL


def test_full_size_chain_properties(model):
    """BASELINE config 3 shape at a size the test can afford (512 ragged programs, up to 25 steps): determinism,
    split invariance, untouched rows beyond n_steps, and dependency gathering checked against the oracle on
    sampled questions (teacher-forced so every step sees oracle-identical inputs)."""
    B = 512
    func, deps, n_steps = orc.fa_programs(B, seed=99)
    g = torch.Generator(device="cuda").manual_seed(5)
    img = torch.randn(B, 1024, 14, 14, device="cuda", generator=g).relu_()
    c1 = fa.run_inference_chain_batched(model, img, func, deps, n_steps, 0, 20)
    c2 = fa.run_inference_chain_batched(model, img, func, deps, n_steps, 0, 20)
    assert torch.equal(c1, c2)
    lo = fa.run_inference_chain_batched(model, img[:200], func[:200], deps[:200], n_steps[:200], 0, 20)
    hi = fa.run_inference_chain_batched(model, img[200:], func[200:], deps[200:], n_steps[200:], 0, 20)
    assert torch.equal(torch.cat([lo, hi]), c1)
    c1 = c1.cpu()
    for b in range(B):
        n = int(n_steps[b])
        assert bool((c1[b, n:] == -1).all()) and bool((c1[b, :n, 0] == 0).all())
        assert int(c1[b, :n].min()) >= 0 and int(c1[b, :n].max()) < 170
    sd = cpu_sd(model)
    rev = orc.fa_vocab(170)
    rows = [3, 200, 511]
    S = func.shape[1]
    forced = torch.zeros(B, S, 19, dtype=torch.long)
    ref_lg = {}
    for r in rows:
        c, lg = orc.fa_run_chain(sd, img[r:r + 1].cpu(), orc.chain_strings(func[r], deps[r], n_steps[r]), rev, 0, 20, 2)
        for i in range(int(n_steps[r])):
            forced[r, i] = torch.tensor(c[i][1:])
        ref_lg[r] = torch.stack([lg[i] for i in range(int(n_steps[r]))])
    cache, logits = fa.run_inference_chain_batched(model, img, func, deps, n_steps, 0, 20, forced=forced, want_logits=True)
    for r in rows:
        n = int(n_steps[r])
        assert torch.equal(cache[r, :n, 1:].cpu().long(), forced[r, :n])
        assert common.rel_err(logits[r, :n], ref_lg[r]) < common.LOGIT_REL_TOL, r


def test_pipelined_chain_slots_match_serial(model):
    B = 48
    func, deps, n_steps = orc.fa_programs(B, seed=31, max_steps=7)
    g = torch.Generator(device="cuda").manual_seed(6)
    img = torch.randn(B, 1024, 14, 14, device="cuda", generator=g).relu_()
    ref = fa.run_inference_chain_batched(model, img, func, deps, n_steps, 0, 20)
    a = fa.run_inference_chain_batched(model, img[:24], func[:24], deps[:24], n_steps[:24], 0, 20, slot=1)
    b = fa.run_inference_chain_batched(model, img[24:], func[24:], deps[24:], n_steps[24:], 0, 20, slot=2)
    model.drain()
    assert torch.equal(torch.cat([a, b]), ref)


def test_chain_with_shared_images_equals_expanded(model):
    """SURVEY 8f next-2 for the FA path: several questions per image.  The chain over `img_tokens` of the unique images
    + image_idx must fill the same cache, bit for bit, as the chain over one (repeated) image per question."""
    B, n_img = 40, 6
    func, deps, n_steps = orc.fa_programs(B, seed=17, max_steps=6)
    g = torch.Generator(device="cuda").manual_seed(8)
    img = torch.randn(n_img, 1024, 14, 14, device="cuda", generator=g).relu_()
    idx = torch.randint(0, n_img, (B,), generator=torch.Generator().manual_seed(1), dtype=torch.int32)
    ref = fa.run_inference_chain_batched(model, img[idx.long().cuda()], func, deps, n_steps, 0, 20)
    got = fa.run_inference_chain_batched(model, img, func, deps, n_steps, 0, 20, image_idx=idx)
    assert torch.equal(got, ref)
    toks = fa.project_images(model, img)
    got2 = fa.run_inference_chain_batched(model, None, func, deps, n_steps, 0, 20, img_tokens=toks, image_idx=idx,
                                          sort_by_steps=False)
    assert torch.equal(got2, ref)
    with pytest.raises(IndexError):
        fa.run_inference_chain_batched(model, img, func, deps, n_steps, 0, 20, image_idx=idx + n_img)
    with pytest.raises(ValueError):
        fa.run_inference_chain_batched(model, img, func, deps, n_steps, 0, 20)      # 6 images, 40 questions, no index
