"""Stage-by-stage comparison of the persistent decode kernel (csrc/decode_persist.cu) with the per-kernel chain.

B200 only.  `--stop K` runs ONE decode position of a one-decoder-layer IQAP model through both paths, the persistent
kernel returning after its K-th cluster rendezvous (0 = the whole position), and prints the difference of every scratch
buffer that is final at that point; `--full` runs the 27-position, two-layer model teacher-forced through both and
compares logits and tokens.  Each configuration runs in its own process (tools/debug_persist.sh) so that a device trap
in one does not take the others with it.
"""
import argparse
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

BUF = {"dx": 0, "dqkv": 1, "dattn": 2, "dx1": 3, "dq": 4, "du": 5, "dx2": 6, "dxo0": 7, "dxo1": 8, "dpre": 9, "tok": 10}
# buffers that hold their final value once the persistent kernel has passed rendezvous K of the first stage
AFTER = {1: ["dqkv"], 2: ["dqkv", "self"], 3: ["dpre"], 4: ["dx1"], 5: ["dx1", "dq"], 6: ["dq", "du"],
         7: ["du", "dattn"], 8: ["dattn"], 9: ["dx2"], 10: ["dx2"], 11: ["dxo0", "tok"],
         0: ["dqkv", "dx1", "dq", "du", "dattn", "dx2", "dxo0", "tok"]}


def make_model(dec_layers, fa_model=False):
    from explainable_spatial_vqa_b200 import inference_transformer_iqap as iqap
    from explainable_spatial_vqa_b200 import inference_transformer_full_annotation_new as fa
    torch.manual_seed(0)
    if fa_model:
        return fa.MultiModalTransformer(170, 256, 2, 1, dec_layers, 512, 0.1, 50, 196).eval().cuda(), fa
    m = iqap.VQAModel(85, 256, 256, 32, 44, 27, 196).eval()
    if dec_layers != 2:
        m.transformer_decoder = nn.TransformerDecoder(nn.TransformerDecoderLayer(d_model=256, nhead=4), num_layers=dec_layers)
        m = m.eval()
    return m.cuda(), iqap


def grab(m, names, B, nhead):
    h = m._native(0)
    out = {}
    for n in names:
        if n == "self":
            continue
        if n == "tok":
            out[n] = h.dbg_workspace(BUF[n], torch.int64).view(-1, 65)[:B, :2].clone()
        elif n == "dpre":
            out[n] = h.dbg_workspace(BUF[n], torch.float32).view(-1, 256)[:B].clone()
        else:
            width = {"dqkv": 768, "dq": nhead * 256, "du": nhead * 256}.get(n, 256)
            out[n] = h.dbg_workspace(BUF[n], torch.bfloat16).view(-1, width)[:B].float().clone()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stop", type=int, default=0)
    ap.add_argument("--full", action="store_true")
    ap.add_argument("--batch", type=int, default=70)
    ap.add_argument("--layers", type=int, default=1)
    args = ap.parse_args()
    from oracle import executor_oracle as orc  # input generators only
    B = args.batch
    img, q = orc.iqap_inputs(B, seed=1234)
    img, q = img.cuda(), q.cuda()

    if args.full:
        os.environ.pop("B200VQA_DECODE", None)
        m0, _ = make_model(2)
        ans0, prog0, lg0, _ = m0.forward_detailed(img, q, want_logits=True)
        _, _, lgf0, _ = m0.forward_detailed(img, q, forced_programs=prog0, want_logits=True)
        torch.cuda.synchronize()
        os.environ["B200VQA_DECODE"] = "persist"
        m1, _ = make_model(2)
        ans1, prog1, lg1, _ = m1.forward_detailed(img, q, want_logits=True)
        _, progf, lgf1, _ = m1.forward_detailed(img, q, forced_programs=prog0, want_logits=True)
        _, prog2 = m1(img, q)
        torch.cuda.synchronize()
        err = float((lgf1 - lgf0).abs().max() / lgf0.abs().max())
        print(f"full: teacher-forced logits rel diff persist vs chain {err:.3e}; forced-run tokens equal "
              f"{float((progf == prog0).float().mean()):.4f}; free-running tokens equal {float((prog1 == prog0).float().mean()):.4f}; "
              f"plain call == logits call {bool((prog2 == prog1).all())}; launches chain {m0.native_launch_count()} persist {m1.native_launch_count()}")
        bad = (lgf1 - lgf0).abs().amax(dim=(1, 2))
        print("  worst questions:", torch.topk(bad, min(5, B)).indices.tolist(), [f"{v:.3e}" for v in torch.topk(bad, min(5, B)).values.tolist()])
        print("  per-position max diff:", [f"{v:.2e}" for v in (lgf1 - lgf0).abs().amax(dim=(0, 2)).tolist()])
        return

    names = AFTER[args.stop]
    os.environ.pop("B200VQA_DECODE", None)
    os.environ.pop("B200VQA_PERSIST_DBG_STOP", None)
    m0, mod = make_model(args.layers)
    mod.Config.PROGRAM_SEQ_LEN = 1
    m0(img, q)
    torch.cuda.synchronize()
    ref = grab(m0, [n for n in AFTER[0]] + ["dpre"], B, 4)
    os.environ["B200VQA_DECODE"] = "persist"
    if args.stop:
        os.environ["B200VQA_PERSIST_DBG_STOP"] = str(args.stop)
    m1, _ = make_model(args.layers)
    m1(img, q)
    torch.cuda.synchronize()
    got = grab(m1, names, B, 4)
    for n in names:
        if n == "self":  # one key at position 0: the self-attention output is the value row
            h = m1._native(0)
            a = h.dbg_workspace(BUF["dattn"], torch.bfloat16).view(-1, 256)[:B].float()
            v = got["dqkv"][:, 512:768]
            print(f"stop {args.stop} self-attn(t=0) vs v: max diff {float((a - v).abs().max()):.3e} (max |v| {float(v.abs().max()):.3e})")
            continue
        if n == "dpre" and args.layers == 1 and args.stop != 3:
            continue
        r = ref[n]
        if n == "dpre":
            print(f"stop {args.stop} dpre: finite {bool(torch.isfinite(got[n]).all())} max |x| {float(got[n].abs().max()):.3e}")
            continue
        d = (got[n].double() - r.double()).abs()
        rows = d.view(B, -1).amax(dim=1)
        print(f"stop {args.stop} {n}: max diff {float(d.max()):.3e} (max |ref| {float(r.double().abs().max()):.3e}); "
              f"rows with diff > 0.1: {int((rows > 0.1).sum())} of {B}; worst rows {torch.topk(rows, 4).indices.tolist()}")
        if float(d.max()) > 0.1:
            cols = d.view(B, -1).amax(dim=0)
            print("   worst columns:", torch.topk(cols, 8).indices.tolist())


if __name__ == "__main__":
    main()
