"""Shared fixtures of the test-suite: seeded models (our drop-in classes), golden files, tolerances."""
import os
import warnings

import numpy as np
import torch

warnings.filterwarnings("ignore", message="enable_nested_tensor")

from explainable_spatial_vqa_b200 import inference_transformer_full_annotation_new as fa  # noqa: E402
from explainable_spatial_vqa_b200 import inference_transformer_iqap as iqap  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# north_star: "bounding boxes and logits within 1e-2 relative (bf16 against the fp32 reference)".
# Relative = max |ours - oracle| / max |oracle| over the compared tensor.
LOGIT_REL_TOL = 1e-2


def rel_err(ours, ref):
    ours = torch.as_tensor(ours).double().cpu()
    ref = torch.as_tensor(ref).double().cpu()
    return float((ours - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def seeded_iqap(seed=0):
    torch.manual_seed(seed)
    return iqap.VQAModel(85, 256, 256, 32, 44, 27, 196).eval()


def seeded_fa(seed=0, nhead=2, enc_layers=1, dec_layers=1, ff=512, max_text_len=50, vocab=170):
    torch.manual_seed(seed)
    return fa.MultiModalTransformer(vocab, 256, nhead, enc_layers, dec_layers, ff, 0.1, max_text_len, 196).eval()


def weights_match_golden(sd, g):
    """True when re-seeding reproduced the parameters the golden file was generated with."""
    keys = [str(k) for k in g["sd_keys"]]
    if sorted(sd) != keys:
        return False
    sums = np.array([sd[k].double().sum().item() for k in keys])
    asums = np.array([sd[k].double().abs().sum().item() for k in keys])
    return bool(np.allclose(sums, g["sd_sums"], rtol=1e-9, atol=1e-9) and np.allclose(asums, g["sd_abs_sums"], rtol=1e-9))


def margins(logits):
    """top-1 minus top-2 of every decision, same leading shape as logits[..., 0]."""
    top = torch.topk(torch.as_tensor(logits).float(), 2, dim=-1).values
    return top[..., 0] - top[..., 1]


def check_tokens_where_decisive(our_tokens, oracle_tokens, oracle_logits, our_logits, what=""):
    """Protocol H1: bf16 cannot reproduce near-tie argmaxes.  Every decision whose oracle
    top-1/top-2 margin exceeds 4x the measured logit error must match exactly; the rest are reported."""
    our_tokens = torch.as_tensor(our_tokens).cpu().long()
    oracle_tokens = torch.as_tensor(oracle_tokens).cpu().long()
    err = float((torch.as_tensor(our_logits).float().cpu() - torch.as_tensor(oracle_logits).float().cpu()).abs().max())
    m = margins(oracle_logits).cpu()
    decisive = m > 4.0 * err
    wrong = (our_tokens != oracle_tokens) & decisive
    assert not bool(wrong.any()), f"{what}: {int(wrong.sum())} decisive tokens differ (logit err {err:.3e})"
    frac = float(decisive.float().mean())
    assert frac > 0.5, f"{what}: only {frac:.2%} of decisions are decisive at logit error {err:.3e}"
    return frac, err
