"""The persistent decode kernel (csrc/decode_persist.cu, opt-in with B200VQA_DECODE=persist): all positions x layers of
the greedy decode in one cluster launch.  Compared with the ORACLE (not with the default path): teacher-forced logits
within north_star's 1e-2, decisive tokens exact, for the IQAP model (nhead 4, 2 layers, ff 2048, fused 44-token head)
and the FA chain (nhead 2, 1 layer, ff 512, 170-token head, ragged memory lengths, final decoder norm)."""
import pytest
import torch

import common
from explainable_spatial_vqa_b200 import inference_transformer_full_annotation_new as fa
from oracle import executor_oracle as orc

pytestmark = pytest.mark.gpu


def test_iqap_persistent_decode_matches_oracle(monkeypatch):
    monkeypatch.setenv("B200VQA_DECODE", "persist")
    model = common.seeded_iqap()
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    img, q = orc.iqap_inputs(70, seed=77)  # two question tiles, the second one partial
    ref = orc.iqap_forward(sd, img, q)
    m = model.cuda()
    ans, prog, logits, _ = m.forward_detailed(img.cuda(), q.cuda(), forced_programs=ref["programs"].cuda(), want_logits=True)
    torch.cuda.synchronize()
    assert m.native_launch_count() <= 16, "the persistent path is a handful of launches per forward"
    assert common.rel_err(ans, ref["answer"]) < common.LOGIT_REL_TOL
    assert common.rel_err(logits, ref["logits"]) < common.LOGIT_REL_TOL
    common.check_tokens_where_decisive(prog, ref["programs"], ref["logits"], logits, "persistent decode tokens")
    # the plain call (no logits, no forcing) takes the same kernel: identical tokens to the instrumented run
    _, prog_free, lg_free, _ = m.forward_detailed(img.cuda(), q.cuda(), want_logits=True)
    _, prog_plain = m(img.cuda(), q.cuda())
    assert torch.equal(prog_plain, prog_free)


def test_fa_chain_persistent_decode_matches_oracle(monkeypatch):
    monkeypatch.setenv("B200VQA_DECODE", "persist")
    f = common.seeded_fa()
    fsd = {k: v.clone() for k, v in f.state_dict().items()}
    f = f.cuda()
    B = 3
    func, deps, n_steps = orc.fa_programs(B, seed=4321, max_steps=4)
    g = torch.Generator().manual_seed(5)
    fimg = torch.randn(B, 1024, 14, 14, generator=g).relu_()
    rev = orc.fa_vocab(170)
    S = func.shape[1]
    forced = torch.zeros(B, S, 19, dtype=torch.long)
    ref_lg = torch.zeros(B, S, 19, 170)
    for b in range(B):
        c, lg = orc.fa_run_chain(fsd, fimg[b:b + 1], orc.chain_strings(func[b], deps[b], n_steps[b]), rev, 0, 20, 2)
        for i in range(int(n_steps[b])):
            forced[b, i] = torch.tensor(c[i][1:])
            ref_lg[b, i] = lg[i]
    cache, lg = fa.run_inference_chain_batched(f, fimg, func, deps, n_steps, 0, 20, forced=forced, want_logits=True)
    torch.cuda.synchronize()
    for b in range(B):
        n = int(n_steps[b])
        assert common.rel_err(lg[b, :n], ref_lg[b, :n]) < common.LOGIT_REL_TOL
        assert torch.equal(cache[b, :n, 1:].cpu().long(), forced[b, :n])
