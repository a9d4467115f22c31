"""The continuous bounding-box head: the reference's bounding-box variant of the IQAP model,
/root/reference/code/train_transformer_iqap_bb.py:222-356 (`bbox_regressor` on the mean of the image-token rows of
the encoder memory, one-layer decoder over program + answer tokens).

CPU: the oracle restatement against the reference's own outputs (tests/golden/iqap_bb_b4.npz, written by
oracle/make_golden.py iqap_bb) and the state-dict layout of the mirror class.  GPU: boxes within north_star's 1e-2
relative of the oracle / golden, teacher-forced sequence logits likewise, decisive tokens exact."""
import numpy as np
import pytest
import torch

import common
from explainable_spatial_vqa_b200 import train_transformer_iqap_bb as bb
from oracle import executor_oracle as orc


def seeded_bb():
    torch.manual_seed(0)
    return bb.VQAModel(85, 256, 256, 44, 27, 196).eval()


@pytest.fixture(scope="module")
def case():
    g = common.load_golden("iqap_bb_b4.npz")
    model = seeded_bb()
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    if not common.weights_match_golden(sd, g):
        pytest.skip(f"seeded init differs from the golden run (torch {torch.__version__} vs {g['torch_version']})")
    img, q = orc.iqap_inputs(4, seed=4321)
    assert np.isclose(img.double().sum().item(), float(g["img_sum"]), rtol=1e-12)
    assert np.array_equal(q.numpy(), g["questions"])
    return g, sd, img, q, model


def test_state_dict_layout_matches_reference():
    sd = seeded_bb().state_dict()
    # names and shapes of train_transformer_iqap_bb.py:243-276 (one encoder layer, ONE decoder layer, 40 box outputs)
    assert sd["bbox_regressor.0.weight"].shape == (256, 256) and sd["bbox_regressor.2.weight"].shape == (40, 256)
    assert sd["decoder_embedding.weight"].shape == (44, 256) and sd["output_layer.weight"].shape == (44, 256)
    assert sd["pos_decoder.pe"].shape == (29, 1, 256) and sd["pos_encoder.pe"].shape == (243, 1, 256)
    assert "transformer_decoder.layers.0.multihead_attn.in_proj_weight" in sd
    assert not any(k.startswith("transformer_decoder.layers.1.") for k in sd)
    assert not any(k.startswith("answer_classifier") for k in sd)


@pytest.mark.parametrize("recompute", [False, True])
def test_oracle_matches_reference(case, recompute):
    g, sd, img, q, _ = case
    out = orc.iqap_bb_forward(sd, img, q, recompute=recompute)
    assert common.rel_err(out["boxes"], g["boxes"]) < 1e-5
    assert common.rel_err(out["seq_logits"], g["seq_logits"]) < 1e-4
    assert np.array_equal(out["tokens"].numpy(), g["tokens"])


@pytest.mark.gpu
def test_boxes_and_logits_match_reference_golden(case):
    g, sd, img, q, model = case
    m = model.cuda()
    forced = torch.from_numpy(g["tokens"]).cuda()
    logits, boxes, tokens = m.forward_detailed(img.cuda(), q.cuda(), forced_tokens=forced)
    torch.cuda.synchronize()
    assert boxes.shape == (4, 10, 4) and logits.shape == (4, 28, 44)
    assert common.rel_err(boxes, g["boxes"]) < common.LOGIT_REL_TOL       # north_star: boxes within 1e-2 relative
    assert common.rel_err(logits, g["seq_logits"]) < common.LOGIT_REL_TOL
    common.check_tokens_where_decisive(tokens, g["tokens"], g["seq_logits"], logits, "iqap_bb tokens")
    # the reference surface: forward() -> (seq_logits, bbox_preds)
    lg2, bx2 = m(img.cuda(), q.cuda())
    assert torch.equal(bx2, boxes) and lg2.shape == logits.shape


@pytest.mark.gpu
def test_boxes_match_oracle_on_a_larger_batch():
    model = seeded_bb()
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    img, q = orc.iqap_inputs(96, seed=99)
    ref = orc.iqap_bb_forward(sd, img, q)
    m = model.cuda()
    logits, boxes, _ = m.forward_detailed(img.cuda(), q.cuda(), forced_tokens=ref["tokens"].cuda())
    torch.cuda.synchronize()
    assert common.rel_err(boxes, ref["boxes"]) < common.LOGIT_REL_TOL
    assert common.rel_err(logits, ref["seq_logits"]) < common.LOGIT_REL_TOL


# ----------------------------------------------------------------------------------------------------------
# Evaluation of the bounding-box variant (IoU utilities train_transformer_iqap_bb.py:104-150,
# `evaluate` :423-538).  tests/golden/iqap_bb_eval.npz holds what the reference's own functions returned
# (oracle/make_golden.py iqap_bb_eval).  The product versions are whole-batch tensor expressions with one host read.
# ----------------------------------------------------------------------------------------------------------
PARTS = [slice(0, 4), slice(4, 8), slice(8, 11)]


@pytest.fixture(scope="module")
def eval_case():
    return common.load_golden("iqap_bb_eval.npz")


def golden_batches(g):
    t = {k: torch.from_numpy(g[k]) for k in ("seq_logits", "boxes", "combined_seq", "gt_boxes")}
    return [(t["seq_logits"][p], t["boxes"][p], t["combined_seq"][p], t["gt_boxes"][p]) for p in PARTS]


def test_oracle_iou_and_evaluate_match_the_reference(eval_case):
    g = eval_case
    preds, gts = torch.from_numpy(g["iou_preds"]), torch.from_numpy(g["iou_gts"])
    assert abs(orc.bb_batch_mean_iou(preds, gts) - float(g["iou"])) < 1e-12
    assert orc.bb_batch_mean_iou(preds[3:4], gts[3:4]) == float(g["iou_none"]) == 0.0
    got = orc.bb_evaluate(golden_batches(g))
    assert np.allclose(got, g["evaluate"], rtol=1e-6, atol=1e-12), (got, g["evaluate"])


def test_iou_utilities_match_the_reference(eval_case):
    g = eval_case
    preds, gts = torch.from_numpy(g["iou_preds"]), torch.from_numpy(g["iou_gts"])
    assert abs(bb.batch_mean_iou(preds, gts) - float(g["iou"])) < 1e-12
    assert bb.batch_mean_iou(preds[3:4], gts[3:4]) == 0.0            # no ground-truth box at all
    s, n = bb.batch_iou_sum_count(preds, gts)
    assert int(n) == int((np.abs(g["iou_gts"]) >= 1e-8).any(-1).sum()) and s.dtype == torch.float64
    # the scalar helper against the oracle's, including empty unions and disjoint / nested / identical boxes
    boxes = [[0.1, 0.1, 0.4, 0.5], [0.2, 0.2, 0.3, 0.3], [0.6, 0.6, 0.9, 0.9], [0.5, 0.5, 0.5, 0.5], [0.4, 0.5, 0.1, 0.1]]
    for a in boxes:
        for b in boxes:
            assert bb.bbox_iou_2d(a, b) == orc.bb_iou(a, b)
    assert bb.bbox_iou_2d(boxes[0], boxes[0]) == 1.0 and bb.bbox_iou_2d(boxes[3], boxes[3]) == 0.0


def test_batch_mean_iou_equals_the_loop_form_on_random_boxes():
    g = torch.Generator().manual_seed(5)
    for trial in range(20):
        B = int(torch.randint(1, 6, (1,), generator=g))
        preds = torch.rand(B, 10, 4, generator=g) * 1.6 - 0.3
        lo = torch.rand(B, 10, 2, generator=g)
        gts = torch.cat([lo, lo + torch.rand(B, 10, 2, generator=g) * 0.5 - 0.05], dim=-1)   # some inverted boxes
        gts = gts * (torch.rand(B, 10, 1, generator=g) > 0.3)
        assert abs(bb.batch_mean_iou(preds, gts) - orc.bb_batch_mean_iou(preds, gts)) < 1e-12, trial


class ReplayModel:
    """Stands in for the model in `evaluate`: returns the reference's own outputs for the batch it is shown."""

    def __init__(self, g):
        self.q = torch.from_numpy(g["questions"])
        self.logits, self.boxes = torch.from_numpy(g["seq_logits"]), torch.from_numpy(g["boxes"])

    def eval(self):
        return self

    def __call__(self, image_features, questions):
        rows = [int((self.q == row).all(dim=1).nonzero()[0]) for row in questions.cpu()]
        return self.logits[rows].to(questions.device), self.boxes[rows].to(questions.device)


def eval_loader(g):
    img, q = orc.iqap_inputs(11, seed=777)
    assert np.isclose(img.double().sum().item(), float(g["img_sum"]), rtol=1e-12) and np.array_equal(q.numpy(), g["questions"])
    seq, gt = torch.from_numpy(g["combined_seq"]), torch.from_numpy(g["gt_boxes"])
    return [(img[p], q[p], seq[p], gt[p]) for p in PARTS]


def test_evaluate_bookkeeping_matches_the_reference(eval_case):
    g = eval_case
    got = bb.evaluate(ReplayModel(g), eval_loader(g), torch.nn.CrossEntropyLoss(), torch.nn.SmoothL1Loss(reduction="none"), "cpu")
    assert len(got) == 7 and all(isinstance(v, float) for v in got)
    assert np.allclose(got, g["evaluate"], rtol=1e-6, atol=1e-12), (got, g["evaluate"])
    assert abs(sum(got[2:6]) - 1.0) < 1e-12 and min(got[2:6]) > 0     # all four tally classes occur in the fixture
    with pytest.raises(ZeroDivisionError):
        bb.evaluate(ReplayModel(g), [], torch.nn.CrossEntropyLoss(), torch.nn.SmoothL1Loss(reduction="none"), "cpu")


@pytest.mark.gpu
def test_evaluate_on_the_gpu_model(eval_case):
    """evaluate() around the CUDA model: its bookkeeping equals the oracle's loops fed the SAME device outputs (exact),
    and the IoU / loss it reports are the reference's within what 1e-2-relative boxes and logits allow."""
    g = eval_case
    model = seeded_bb()
    with torch.no_grad():
        model.bbox_regressor[2].bias.copy_(torch.from_numpy(g["box_bias"]).repeat(10))
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    if not common.weights_match_golden(sd, g):
        pytest.skip(f"seeded init differs from the golden run (torch {torch.__version__} vs {g['torch_version']})")
    m = model.cuda()
    loader = eval_loader(g)
    crit = (torch.nn.CrossEntropyLoss(), torch.nn.SmoothL1Loss(reduction="none"))
    got = bb.evaluate(m, loader, *crit, "cuda")
    outs = []
    for img, q, seq, gt in loader:
        lg, bx = m(img.cuda(), q.cuda())
        outs.append((lg.cpu(), bx.cpu(), seq, gt))
    want = orc.bb_evaluate(outs)
    assert np.allclose(got, want, rtol=1e-5, atol=1e-9), (got, want)
    ref = g["evaluate"]
    assert abs(got[1] - ref[1]) < 1e-2 and abs(got[0] - ref[0]) < 2e-2 * abs(ref[0]), (got, ref)
    assert abs(sum(got[2:6]) - 1.0) < 1e-12
