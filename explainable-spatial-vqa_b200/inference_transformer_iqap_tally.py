"""Drop-in for the driver half of the reference's `code/inference_transformer_iqap_tally.py`.

The reference loops over the validation set one sample at a time (TALLY:306-344): H5 read, `.to(device)`, batch-1
forward, `.item()` / `.tolist()`, four Python counters.  Here the same tally is a batched device pipeline: the unique
images of a batch cross PCIe once (`VQAModel.forward_host_indexed` semantics on device tensors), the model runs on
the whole batch, and the four counters live in HBM (`b200vqa_iqap_tally`) - one device->host read at the very end.

The model classes are the ones of `inference_transformer_iqap` (the reference duplicates them verbatim, TALLY:95-241).
"""
from __future__ import annotations

import numpy as np
import torch

from .inference_transformer_iqap import Config, VQAModel, PositionalEncoding, generate_square_subsequent_mask  # noqa: F401

TALLY_NAMES = ("Both Answer and Program Correct", "Answer Correct but Program Incorrect",
               "Answer Incorrect but Program Correct", "Both Answer and Program Incorrect")


@torch.no_grad()
def tally_dataset(model: VQAModel, features, image_idxs, questions, answers, programs, max_samples=None,
                  batch_size=1024):
    """features (n_images, 1024, 14, 14) or (n_images, 196, 1024) f32 (array-like: numpy, torch or an h5py dataset),
    image_idxs / answers (N,), questions (N, 46), programs (N, 27).  Returns the four tallies as a tuple of ints in
    the reference's order (TALLY:339-344)."""
    dev = model.image_proj.weight.device
    n = len(questions) if max_samples is None else min(len(questions), int(max_samples))
    counts = torch.zeros(4, dtype=torch.int64, device=dev)
    for b0 in range(0, n, batch_size):
        b1 = min(n, b0 + batch_size)
        idx = np.asarray(image_idxs[b0:b1]).astype(np.int64)
        uniq, inverse = np.unique(idx, return_inverse=True)          # h5py wants sorted, unique fancy indices
        feats = torch.as_tensor(np.asarray(features[uniq.tolist()]), dtype=torch.float32)
        if feats.dim() == 4:                                           # (n, 1024, 14, 14) -> (n, 196, 1024)  (IQAP:72)
            feats = feats.flatten(2).transpose(1, 2).contiguous()
        q = torch.as_tensor(np.asarray(questions[b0:b1]), dtype=torch.int64)
        answer_output, generated = model.forward_indexed(feats.to(dev, non_blocking=True),
                                                         torch.as_tensor(inverse, dtype=torch.int32).to(dev),
                                                         q.to(dev))
        gt_a = torch.as_tensor(np.asarray(answers[b0:b1]), dtype=torch.int64).to(dev)
        gt_p = torch.as_tensor(np.asarray(programs[b0:b1]), dtype=torch.int64).to(dev)
        model.tally(answer_output, generated, gt_a, gt_p, counts)
    return tuple(int(c) for c in counts.cpu())


def run_inference(tally=True, max_samples=None, model=None, batch_size=1024):
    """Same entry point as the reference (TALLY:279): opens Config.FEATURES_H5 / Config.QUESTIONS_H5, tallies the four
    cases and prints the reference's report.  Needs h5py and the checkpoint like the reference does."""
    import h5py  # noqa: PLC0415 - optional dependency, exactly as in the reference

    from .inference_transformer_iqap import load_model
    device = torch.device("cuda")
    print(f"Using device: {device}")
    model = model or load_model(device)
    print("Model loaded successfully.")
    with h5py.File(Config.QUESTIONS_H5, "r") as fq, h5py.File(Config.FEATURES_H5, "r") as ff:
        total = fq["questions"].shape[0] if max_samples is None else min(fq["questions"].shape[0], max_samples)
        print(f"Total samples to process: {total}")
        t = tally_dataset(model, ff["features"], fq["image_idxs"], fq["questions"], fq["answers"], fq["programs"],
                          total, batch_size)
    if tally:
        print("\n=== Inference Tally Results ===")
        print(f"Total Samples Processed: {total}")
        for i, (name, c) in enumerate(zip(TALLY_NAMES, t), 1):
            print(f"{i}. {name}: {c}")
        print("================================\n")
    return t
