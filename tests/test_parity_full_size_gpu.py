"""Parity at the BASELINE sizes, against the ORACLE.

* IQAP, 1024 questions per call (configs[1]): teacher-forced per-position logits and answers of 128 sampled questions
  against the oracle run on exactly those questions, decisive tokens exact.
* FA, 4096 ragged programs per call (configs[2]): teacher-forced per-step logits of 64 sampled questions against the
  oracle's batch-1 chain (the reference's only mode), cache rows exact under forcing.
* Every environment switch that selects another kernel sequence is compared with the oracle, not with the default
  CUDA path.
* The host-buffer entry of the FA chain equals the device-tensor entry bit for bit, also across sub-batches.
"""
import pytest
import torch

import common
from explainable_spatial_vqa_b200 import inference_transformer_full_annotation_new as fa
from oracle import executor_oracle as orc

pytestmark = pytest.mark.gpu


def cpu_sd(m):
    return {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}


def test_iqap_1024_questions_128_sampled_against_oracle():
    model = common.seeded_iqap()
    sd = cpu_sd(model)
    model = model.cuda()
    B = 1024
    img, q = orc.iqap_inputs(B, seed=2025)
    rows = torch.randperm(B, generator=torch.Generator().manual_seed(1))[:128].sort().values
    ref = orc.iqap_forward(sd, img[rows], q[rows])
    forced = torch.zeros(B, 27, dtype=torch.long)
    forced[rows] = ref["programs"]
    ans, prog, logits, _ = model.forward_detailed(img.cuda(), q.cuda(), forced_programs=forced.cuda(), want_logits=True)
    torch.cuda.synchronize()
    worst_ans = max(common.rel_err(ans[r], ref["answer"][k]) for k, r in enumerate(rows.tolist()))
    worst_log = max(common.rel_err(logits[r], ref["logits"][k]) for k, r in enumerate(rows.tolist()))
    assert worst_ans < common.LOGIT_REL_TOL and worst_log < common.LOGIT_REL_TOL, (worst_ans, worst_log)
    frac, err = common.check_tokens_where_decisive(prog[rows.cuda()], ref["programs"], ref["logits"], logits[rows.cuda()],
                                                   "iqap 1024/128")
    print(f"iqap B=1024: 128 sampled questions, worst per-question logits rel err {worst_log:.2e}, answers {worst_ans:.2e}, "
          f"decisive decisions {frac:.3f}")


def test_fa_4096_programs_64_sampled_against_oracle():
    model = common.seeded_fa()
    sd = cpu_sd(model)
    model = model.cuda()
    B = 4096
    func, deps, n_steps = orc.fa_programs(B, seed=4321)
    g = torch.Generator(device="cuda").manual_seed(4321)
    img = torch.randn(B, 1024, 14, 14, device="cuda", generator=g).relu_()
    rows = torch.randperm(B, generator=torch.Generator().manual_seed(2))[:64].sort().values.tolist()
    rev = orc.fa_vocab(170)
    S = func.shape[1]
    forced = torch.zeros(B, S, 19, dtype=torch.long)
    ref_lg = {}
    for r in rows:
        c, lg = orc.fa_run_chain(sd, img[r:r + 1].cpu(), orc.chain_strings(func[r], deps[r], n_steps[r]), rev, 0, 20, 2)
        for i in range(int(n_steps[r])):
            forced[r, i] = torch.tensor(c[i][1:])
        ref_lg[r] = torch.stack([lg[i] for i in range(int(n_steps[r]))])
    cache, logits = fa.run_inference_chain_batched(model, img, func, deps, n_steps, 0, 20, forced=forced, want_logits=True)
    torch.cuda.synchronize()
    worst = 0.0
    for r in rows:
        n = int(n_steps[r])
        assert torch.equal(cache[r, :n, 1:].cpu().long(), forced[r, :n]), r
        assert bool((cache[r, n:] == -1).all())
        worst = max(worst, common.rel_err(logits[r, :n], ref_lg[r]))
    assert worst < common.LOGIT_REL_TOL, worst
    print(f"fa B=4096: 64 sampled questions ({sum(int(n_steps[r]) for r in rows)} program steps), worst logits rel err {worst:.2e}")


@pytest.mark.parametrize("switch", ["B200VQA_NO_ABSORB", "B200VQA_ABSORB_OV", "B200VQA_NO_FUSED_HEAD",
                                    "B200VQA_NO_WARP_SELF_ATTN", "B200VQA_NO_LN_CLUSTER", "B200VQA_NO_GRAPH",
                                    "B200VQA_NO_PDL", "B200VQA_NO_FUSED_ENC_FFN", "B200VQA_ENC_ATTN_WHOLE_HEAD",
                                    "B200VQA_NO_IMG_PROJ_PAIR", "B200VQA_NO_L2_HINTS"])
def test_every_kernel_switch_against_the_oracle(monkeypatch, switch):
    """Each switch selects a different kernel sequence for the same mathematics: all of them must meet the oracle gate
    themselves (comparing them with the default CUDA path would leave no margin: 5e-3 + 4e-3)."""
    monkeypatch.setenv(switch, "1")
    model = common.seeded_iqap()
    sd = cpu_sd(model)
    model = model.cuda()
    img, q = orc.iqap_inputs(16, seed=31)
    ref = orc.iqap_forward(sd, img, q)
    ans, prog, logits, _ = model.forward_detailed(img.cuda(), q.cuda(), forced_programs=ref["programs"].cuda(), want_logits=True)
    torch.cuda.synchronize()
    assert common.rel_err(ans, ref["answer"]) < common.LOGIT_REL_TOL
    assert common.rel_err(logits, ref["logits"]) < common.LOGIT_REL_TOL
    common.check_tokens_where_decisive(prog, ref["programs"], ref["logits"], logits, switch)
    _, p_free = model(img.cuda(), q.cuda())  # the plain (graph-replayed unless disabled) call runs too
    assert int(p_free.min()) >= 0 and int(p_free.max()) < 44
    if switch == "B200VQA_NO_PDL":
        monkeypatch.delenv(switch)
        # process-wide in the library (re-evaluated on every create): a handle created without it restores the default
        common.seeded_iqap().cuda()(img[:2].cuda(), q[:2].cuda())


def test_fa_host_entry_equals_device_entry():
    model = common.seeded_fa().cuda()
    B = 11
    func, deps, n_steps = orc.fa_programs(B, seed=17, max_steps=5)
    g = torch.Generator().manual_seed(9)
    img = torch.randn(B, 1024, 14, 14, generator=g).relu_()
    want = fa.run_inference_chain_batched(model, img.cuda(), func, deps, n_steps, 0, 20).cpu()
    for chunk in (4, 2048):  # three sub-batches (double-buffered staging reused) / one
        got = fa.run_inference_chain_host(model, img.pin_memory(), func, deps, n_steps, 0, 20, chunk=chunk)
        assert got.dtype == torch.int32 and tuple(got.shape) == tuple(want.shape)
        assert torch.equal(got, want), chunk
    # a second call re-uses the staging while nothing is in flight; results stay identical
    assert torch.equal(fa.run_inference_chain_host(model, img.pin_memory(), func, deps, n_steps, 0, 20, chunk=4), want)
    with pytest.raises(ValueError):
        fa.run_inference_chain_host(model, img.cuda(), func, deps, n_steps)
    # two concurrent parts on separate (handle, stream) slots, several sub-batches each
    B = 600
    func, deps, n_steps = orc.fa_programs(B, seed=18, max_steps=3)
    img = torch.randn(B, 1024, 14, 14, generator=g).relu_()
    want = fa.run_inference_chain_batched(model, img.cuda(), func, deps, n_steps, 0, 20).cpu()
    got = fa.run_inference_chain_host(model, img.pin_memory(), func, deps, n_steps, 0, 20, chunk=128, parts=2)
    assert torch.equal(got, want)
    # asynchronous submission on round-robin slots: three batches in flight on two slots, valid after drain_host()
    outs = [fa.submit_inference_chain_host(model, img[i * 200:(i + 1) * 200].pin_memory(), func[i * 200:(i + 1) * 200],
                                           deps[i * 200:(i + 1) * 200], n_steps[i * 200:(i + 1) * 200], 0, 20, chunk=64,
                                           depth=2) for i in range(3)]
    model.drain_host()
    assert torch.equal(torch.cat(outs), want)
    # upload="bf16": features rounded on host threads before the upload - what the device path does to them anyway, so
    # the caches are identical; several image groups per sub-batch (staging reused), synchronous and pipelined entries
    got = fa.run_inference_chain_host(model, img.pin_memory(), func, deps, n_steps, 0, 20, chunk=300, parts=1, upload="bf16")
    assert torch.equal(got, want)
    outs = [fa.submit_inference_chain_host(model, img[i * 200:(i + 1) * 200].pin_memory(), func[i * 200:(i + 1) * 200],
                                           deps[i * 200:(i + 1) * 200], n_steps[i * 200:(i + 1) * 200], 0, 20, chunk=64,
                                           depth=2, upload="bf16") for i in range(3)]
    model.drain_host()
    assert torch.equal(torch.cat(outs), want)
    assert fa.resolve_upload("auto") in ("bf16", "fp32")
