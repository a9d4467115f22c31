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
