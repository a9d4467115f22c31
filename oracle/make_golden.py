"""Generates tests/golden/*.npz from the REFERENCE ITSELF (run in the build container, where
/root/reference exists; the GPU box only sees the committed fixtures).

    python oracle/make_golden.py

Imports the reference's two inference modules (with a stub `h5py`, which only their dataset helpers
need), builds them under torch.manual_seed(0), runs them on the seeded synthetic inputs of
oracle/executor_oracle.py and stores: weight checksums (so a test can tell whether re-seeding reproduces
the same parameters), inputs' checksums and the reference's outputs.  Per-position logits are captured
with forward hooks on the reference's output heads (the reference only returns token ids).
"""
import os
import sys
import types

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.modules.setdefault("h5py", types.ModuleType("h5py"))
sys.path.insert(0, "/root/reference/code")

import inference_transformer_iqap as ref_iqap  # noqa: E402
import inference_transformer_full_annotation_new as ref_fa  # noqa: E402
from oracle import executor_oracle as orc  # noqa: E402

OUT = os.path.join(REPO, "tests", "golden")


def checksums(sd):
    return {k: np.float64(v.double().sum().item()) for k, v in sd.items()}


def pack_checksums(sd):
    keys = sorted(sd)
    return np.array(keys), np.array([sd[k].double().sum().item() for k in keys], dtype=np.float64), \
        np.array([sd[k].double().abs().sum().item() for k in keys], dtype=np.float64)


def golden_iqap():
    torch.manual_seed(0)
    model = ref_iqap.VQAModel(85, 256, 256, 32, 44, 27, 196).eval()
    img, q = orc.iqap_inputs(4, seed=1234)
    captured = []
    hook = model.program_output.register_forward_hook(lambda m, i, o: captured.append(o.detach().clone()))
    mem_cap = []
    hook2 = model.transformer_encoder.register_forward_hook(lambda m, i, o: mem_cap.append(o.detach().clone()))
    with torch.no_grad():
        answer, programs = model(img, q)
    hook.remove()
    hook2.remove()
    logits = torch.stack(captured, dim=1)  # (B, 27, 44)
    keys, sums, asums = pack_checksums(model.state_dict())
    np.savez_compressed(os.path.join(OUT, "iqap_b4.npz"), torch_version=torch.__version__, sd_keys=keys, sd_sums=sums,
                        sd_abs_sums=asums, img_sum=img.double().sum().item(), questions=q.numpy(),
                        answer=answer.numpy(), programs=programs.numpy(), logits=logits.numpy(),
                        memory_cls=mem_cap[0][0].numpy(), memory_last=mem_cap[0][-1].numpy())
    print("iqap: programs[0] =", programs[0].tolist())


def golden_fa():
    torch.manual_seed(0)
    model = ref_fa.MultiModalTransformer(170, 256, 2, 1, 1, 512, 0.1, 50, 196).eval()
    rev_vocab = orc.fa_vocab(170)
    g = torch.Generator().manual_seed(77)
    img = torch.randn(1, 1024, 14, 14, generator=g).relu_()
    out = {}
    keys, sums, asums = pack_checksums(model.state_dict())
    out.update(torch_version=torch.__version__, sd_keys=keys, sd_sums=sums, sd_abs_sums=asums,
               img_sum=img.double().sum().item())
    # (1) greedy_decode at the three src lengths the chain produces (1, 21, 41 tokens)
    captured = []
    hook = model.output_linear.register_forward_hook(lambda m, i, o: captured.append(o.detach()[:, -1].clone()))
    for s in (1, 21, 41):
        src = torch.randint(0, 170, (1, s), generator=g)
        captured.clear()
        ys = ref_fa.greedy_decode(model, img, src, 0, 20, torch.device("cpu"))
        out[f"gd_src_{s}"] = src.numpy()
        out[f"gd_ys_{s}"] = ys.numpy()
        out[f"gd_logits_{s}"] = torch.stack(captured, dim=1).numpy()  # (1, 19, V)
    # (2) teacher-forced forward, batch 2
    img2 = torch.randn(2, 1024, 14, 14, generator=g).relu_()
    src2 = torch.randint(0, 170, (2, 21), generator=g)
    tgt2 = torch.randint(0, 170, (2, 20), generator=g)
    hook.remove()
    with torch.no_grad():
        out["fw_logits"] = model(img2, src2, tgt2).numpy()
    out["fw_img_sum"] = img2.double().sum().item()
    out["fw_src"] = src2.numpy()
    out["fw_tgt"] = tgt2.numpy()
    # (3) run_inference_chain on an 11-step CLEVR-shaped program: two roots, unary chains, one binary op
    #     (shape of examples/CLEVR_train_questions_first.json: scene, filter x3, unique, same_size-ish relate,
    #      second branch, final binary comparison)
    func = [30, 31, 32, 33, 34, 30, 35, 36, 37, 38, 39]
    deps = [[], [0], [1], [2], [3], [], [5], [6], [7], [8], [4, 9]]
    chain = [" ".join([str(f)] + [str(d) for d in dd]) for f, dd in zip(func, deps)]
    captured.clear()
    logits_per_step = []
    hook = model.output_linear.register_forward_hook(lambda m, i, o: captured.append(o.detach()[:, -1].clone()))
    final, cache = ref_fa.run_inference_chain(model, img, chain, torch.device("cpu"), 0, 20, rev_vocab)
    hook.remove()
    steps = len(chain)
    out["chain_func"] = np.array(func, dtype=np.int32)
    out["chain_deps"] = np.array([dd + [-1] * (2 - len(dd)) for dd in deps], dtype=np.int32)
    out["chain_cache"] = np.array([[int(t) for t in cache[i].split()] for i in range(steps)], dtype=np.int32)
    out["chain_logits"] = torch.stack(captured, dim=0).view(steps, 19, -1).numpy()
    out["chain_final"] = np.array([int(t) for t in final.split()], dtype=np.int32)
    np.savez_compressed(os.path.join(OUT, "fa_nhead2.npz"), **out)
    print("fa: chain final =", final)


def golden_iqap_bb():
    """The reference's bounding-box variant (train_transformer_iqap_bb.py:222-356): boxes + 28 x Vp logits."""
    import train_transformer_iqap_bb as ref_bb
    torch.manual_seed(0)
    model = ref_bb.VQAModel(85, 256, 256, 44, 27, 196).eval()
    img, q = orc.iqap_inputs(4, seed=4321)
    with torch.no_grad():
        seq_logits, boxes = model(img, q)
    keys, sums, asums = pack_checksums(model.state_dict())
    np.savez_compressed(os.path.join(OUT, "iqap_bb_b4.npz"), torch_version=torch.__version__, sd_keys=keys, sd_sums=sums,
                        sd_abs_sums=asums, img_sum=img.double().sum().item(), questions=q.numpy(),
                        seq_logits=seq_logits.numpy(), boxes=boxes.numpy(),
                        tokens=seq_logits.argmax(-1).numpy())
    print("iqap_bb: boxes[0,0] =", boxes[0, 0].tolist())


BB_EVAL_BOX_BIAS = torch.tensor([0.15, 0.2, 0.55, 0.65])  # planted into bbox_regressor.2.bias by the evaluation golden


def bb_eval_case():
    """Seeded evaluation data of the bounding-box variant: 11 samples in batches of 4, 4, 3; ground-truth boxes with
    zero-padded tails; (the sequences are made from the model's own greedy output by the caller)."""
    img, q = orc.iqap_inputs(11, seed=777)
    g = torch.Generator().manual_seed(778)
    xy = torch.rand(11, 10, 2, generator=g) * 0.6
    wh = torch.rand(11, 10, 2, generator=g) * 0.3 + 0.1
    gt = torch.cat([xy, xy + wh], dim=-1)
    n_real = torch.randint(1, 11, (11,), generator=g)
    gt = gt * (torch.arange(10)[None, :] < n_real[:, None])[..., None]
    return img, q, gt, [slice(0, 4), slice(4, 8), slice(8, 11)]


def golden_iqap_bb_eval():
    """IoU utilities and `evaluate` of the bounding-box variant (train_transformer_iqap_bb.py:104-150, :423-538), run
    by the reference itself: (a) batch_mean_iou on random boxes, (b) evaluate() on three batches whose ground-truth
    sequences are the model's own greedy output with planted errors, so that all four tally classes occur."""
    import train_transformer_iqap_bb as ref_bb
    g = torch.Generator().manual_seed(99)
    preds = torch.rand(6, 10, 4, generator=g) * 1.4 - 0.2          # some coordinates outside [0, 1]: clamped
    lo = torch.rand(6, 10, 2, generator=g) * 0.7
    gts = torch.cat([lo, lo + torch.rand(6, 10, 2, generator=g) * 0.3], dim=-1)
    gts[0, 6:] = 0
    gts[3] = 0                                                      # a sample without boxes
    preds[1, :3] = gts[1, :3]                                       # exact hits
    preds[2, 0] = torch.tensor([0.5, 0.5, 0.4, 0.4])                # degenerate prediction: zero area
    iou = ref_bb.batch_mean_iou(preds, gts)
    iou_none = ref_bb.batch_mean_iou(preds[3:4], gts[3:4])

    torch.manual_seed(0)
    model = ref_bb.VQAModel(85, 256, 256, 44, 27, 196).eval()
    with torch.no_grad():  # a random-init head predicts boxes around 0: move them into the image so that IoUs are not all 0
        model.bbox_regressor[2].bias.copy_(BB_EVAL_BOX_BIAS.repeat(10))
    img, q, gt_boxes, parts = bb_eval_case()
    with torch.no_grad():
        seq_logits, boxes = model(img, q)
    seq = seq_logits.argmax(-1).clone()
    seq[3, -1] = (seq[3, -1] + 1) % 44                             # correct program, wrong answer
    seq[4, 5] = (seq[4, 5] + 1) % 44                               # wrong program, correct answer
    seq[5, 0] = (seq[5, 0] + 3) % 44
    for i in (6, 7, 9):                                            # both wrong
        seq[i, 2] = (seq[i, 2] + 1) % 44
        seq[i, -1] = (seq[i, -1] + 2) % 44
    loader = [(img[p], q[p], seq[p], gt_boxes[p]) for p in parts]
    import torch.nn as nn
    import tqdm as _tqdm  # noqa: F401  (the reference wraps the loader in tqdm)
    res = ref_bb.evaluate(model, loader, nn.CrossEntropyLoss(), nn.SmoothL1Loss(reduction="none"), torch.device("cpu"))
    keys, sums, asums = pack_checksums(model.state_dict())
    np.savez_compressed(os.path.join(OUT, "iqap_bb_eval.npz"), torch_version=torch.__version__, sd_keys=keys, sd_sums=sums,
                        sd_abs_sums=asums, iou_preds=preds.numpy(), iou_gts=gts.numpy(), iou=np.float64(iou),
                        iou_none=np.float64(iou_none), img_sum=img.double().sum().item(), questions=q.numpy(),
                        gt_boxes=gt_boxes.numpy(), combined_seq=seq.numpy(), seq_logits=seq_logits.numpy(),
                        boxes=boxes.numpy(), evaluate=np.array(res, dtype=np.float64), box_bias=BB_EVAL_BOX_BIAS.numpy())
    print("iqap_bb_eval: batch_mean_iou =", iou, "evaluate =", res)


def golden_lstm():
    import run_model_lstm_qp as ref_qp
    from oracle import lstm_oracle
    torch.manual_seed(0)
    model = ref_qp.Seq2SeqModel(85, 256, 512, 44, 27, 1).eval()
    q = lstm_oracle.questions(6, seed=4242)
    captured = []
    hook = model.fc.register_forward_hook(lambda m, i, o: captured.append(o.detach()[:, 0].clone()))
    with torch.no_grad():
        programs = model(q)
    hook.remove()
    logits = torch.stack(captured, dim=1)  # (B, 27, 44)
    with torch.no_grad():
        tgt = torch.cat([torch.ones(6, 1, dtype=torch.long), programs[:, :-1]], dim=1)
        tf_logits = model(q, tgt)  # teacher-forced branch fed [<START>, p0..p25] reproduces the greedy logits
    keys, sums, asums = pack_checksums(model.state_dict())
    np.savez_compressed(os.path.join(OUT, "lstm_qp.npz"), torch_version=torch.__version__, sd_keys=keys, sd_sums=sums,
                        sd_abs_sums=asums, questions=q.numpy(), programs=programs.numpy(), logits=logits.numpy(),
                        tf_logits=tf_logits.numpy())
    print("lstm: programs[0] =", programs[0].tolist())


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    only = sys.argv[1:]
    for name, fn in (("iqap", golden_iqap), ("fa", golden_fa), ("lstm", golden_lstm), ("iqap_bb", golden_iqap_bb),
                     ("iqap_bb_eval", golden_iqap_bb_eval)):
        if not only or name in only:
            fn()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
