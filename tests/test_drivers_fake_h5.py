"""The reference's driver functions - load_model / run_inference (IQAP:255-319), the tally driver and
main_inference (FA:151-206) - executed end to end on a FAKE `h5py` (the image has none, and the reference's data files
are not shipped): datasets are numpy arrays behind the h5py.File / dataset[...] surface the drivers use."""
import json
import sys
import types

import numpy as np
import pytest
import torch

import common
from oracle import executor_oracle as orc

pytestmark = pytest.mark.gpu


class _FakeDataset:
    def __init__(self, arr):
        self.arr = arr
        self.shape = getattr(arr, "shape", ())

    def __getitem__(self, key):
        if isinstance(key, tuple) and key == ():
            return self.arr
        if isinstance(key, list):
            return self.arr[np.asarray(key)]
        return self.arr[key]

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.arr, dtype=dtype)

    def __len__(self):
        return len(self.arr)


class _FakeFile(dict):
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


@pytest.fixture
def fake_h5py(monkeypatch):
    store = {}
    mod = types.ModuleType("h5py")
    mod.File = lambda path, mode="r": _FakeFile({k: _FakeDataset(v) for k, v in store[path].items()})
    monkeypatch.setitem(sys.modules, "h5py", mod)
    return store


def test_iqap_load_model_and_run_inference(fake_h5py, tmp_path, capsys):
    from explainable_spatial_vqa_b200 import inference_transformer_iqap as iqap
    N = 7
    img, q = orc.iqap_inputs(N, seed=5)
    feats = img.view(N, 14, 14, 1024).permute(0, 3, 1, 2).contiguous().numpy()   # the H5 layout (N, 1024, 14, 14)
    q = q.clone()
    q[0, 1] = 84  # make the max id + 1 equal the constructor's vocab
    programs = torch.randint(0, 44, (N, 27))
    programs[0, 0] = 43
    answers = torch.randint(0, 32, (N,))
    answers[0] = 31
    fake_h5py["feat.h5"] = {"features": feats}
    fake_h5py["q.h5"] = {"questions": q.numpy(), "answers": answers.numpy(), "programs": programs.numpy(),
                         "image_idxs": np.arange(N)}
    model = common.seeded_iqap()
    ckpt = tmp_path / "best.pth"
    torch.save(model.state_dict(), ckpt)
    saved = {k: getattr(iqap.Config, k, None) for k in ("FEATURES_H5", "QUESTIONS_H5", "MODEL_NAME")}
    try:
        iqap.Config.FEATURES_H5, iqap.Config.QUESTIONS_H5, iqap.Config.MODEL_NAME = "feat.h5", "q.h5", str(ckpt)
        assert iqap.get_data_info("q.h5") == (85, 32, 44)
        loaded = iqap.load_model(torch.device("cuda"))
        assert not loaded.training and next(loaded.parameters()).is_cuda
        results = iqap.run_inference(idx=6, batch_size=4)   # two batches: 4 + 2 samples
    finally:
        for k, v in saved.items():
            setattr(iqap.Config, k, v)
    assert [r[0] for r in results] == list(range(6))
    ans, prog = model.cuda()(img[:6].cuda(), q[:6].cuda())
    assert [r[1] for r in results] == ans.argmax(1).tolist()
    assert [r[2] for r in results] == prog.tolist()
    out = capsys.readouterr().out
    assert "sample 5: predicted answer" in out and "ground truth" in out


def test_fa_main_inference(fake_h5py, tmp_path, caplog):
    from explainable_spatial_vqa_b200 import inference_transformer_full_annotation_new as fa
    rev = orc.fa_vocab(170)
    vocab = {v: k for k, v in rev.items()}
    vocab_path = tmp_path / "vocab.json"
    vocab_path.write_text(json.dumps(vocab))
    func, deps, n_steps = orc.fa_programs(3, seed=8, max_steps=4)
    g = torch.Generator().manual_seed(3)
    feats = torch.randn(5, 1024, 14, 14, generator=g).relu_()
    questions = [{"final_chain_of_thought": orc.chain_strings(func[b], deps[b], n_steps[b]), "image_index": 4 - b,
                  "answer": "yes"} for b in range(3)]
    fake_h5py["ann.h5"] = {"questions": json.dumps({"questions": questions}).encode("utf-8")}
    fake_h5py["feat.h5"] = {"features": feats.numpy()}
    model = common.seeded_fa()
    ckpt = tmp_path / "mm.pth"
    torch.save(model.state_dict(), ckpt)
    import logging
    with caplog.at_level(logging.INFO):
        cache = fa.main_inference(str(ckpt), str(vocab_path), "ann.h5", "feat.h5", num_examples=3)
    want = fa.run_inference_chain_batched(model.cuda(), feats[[4, 3, 2]].cuda(), func, deps, n_steps, 0, 20).cpu()
    S = int(n_steps.max())
    assert torch.equal(cache[:, :S], want[:, :S])
    assert any("final predicted sentence" in r.message for r in caplog.records)


def test_tally_run_inference_driver(fake_h5py, tmp_path, capsys):
    """inference_transformer_iqap_tally.run_inference (TALLY:279-357) end to end: checkpoint + H5 files -> four tallies
    that equal the reference's per-sample loop (oracle) on the library's own outputs."""
    from explainable_spatial_vqa_b200 import inference_transformer_iqap as iqap
    from explainable_spatial_vqa_b200 import inference_transformer_iqap_tally as tally
    N = 12
    img, q = orc.iqap_inputs(4, seed=6)                      # four images, three questions each
    feats = img.view(4, 14, 14, 1024).permute(0, 3, 1, 2).contiguous().numpy()
    _, qs = orc.iqap_inputs(N, seed=7)
    qs[0, 1] = 84
    image_idxs = np.arange(N) % 4
    model = common.seeded_iqap()
    ckpt = tmp_path / "best.pth"
    torch.save(model.state_dict(), ckpt)
    m = model.cuda()
    ans, prog = m(img[image_idxs].cuda(), qs.cuda())
    gt_prog = prog.cpu().clone()
    gt_prog[::2, 3] = (gt_prog[::2, 3] + 1) % 44             # every other program is "wrong"
    gt_prog[0, 0] = 43
    gt_ans = ans.argmax(1).cpu().clone()
    gt_ans[::3] = (gt_ans[::3] + 1) % 32                     # every third answer is "wrong"
    gt_ans[1] = 31
    fake_h5py["feat.h5"] = {"features": feats}
    fake_h5py["q.h5"] = {"questions": qs.numpy(), "answers": gt_ans.numpy(), "programs": gt_prog.numpy(),
                         "image_idxs": image_idxs}
    saved = {k: getattr(iqap.Config, k, None) for k in ("FEATURES_H5", "QUESTIONS_H5", "MODEL_NAME")}
    try:
        iqap.Config.FEATURES_H5, iqap.Config.QUESTIONS_H5, iqap.Config.MODEL_NAME = "feat.h5", "q.h5", str(ckpt)
        got = tally.run_inference(tally=True, batch_size=5)   # three batches
    finally:
        for k, v in saved.items():
            setattr(iqap.Config, k, v)
    want, _ = orc.iqap_tally(ans.cpu(), prog.cpu(), gt_ans, gt_prog)
    assert list(got) == want and sum(got) == N
    assert "=== Inference Tally Results ===" in capsys.readouterr().out
