#!/usr/bin/env python
"""BASELINE config 5: CoGenT-shaped sweep of the FA executor step - batch 256..16k x sequence 197..260 tokens
(src length 1 / 21 / 41 / 64) on one GPU.  Prints a markdown table (program-steps/s, ms per step, achieved
TFLOP/s from the algorithmic FLOPs of SURVEY §8d).  Multi-GPU scaling is `bench.py --gpus N`.

    python tools/sweep.py > profiles/r1_sweep_fa_step.md
"""
import os
import sys
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import torch

from explainable_spatial_vqa_b200 import inference_transformer_full_annotation_new as fa

MFLOP = {1: 328.5, 21: 363.6, 41: 399.5, 64: 441.8}  # per program-step, SURVEY §8d


def main():
    dev = torch.device("cuda")
    print("| batch | src tokens | seq L | ms / step | program-steps/s | algorithmic TFLOP/s |")
    print("|---:|---:|---:|---:|---:|---:|")
    for s, text_len in ((1, 50), (21, 50), (41, 50), (64, 64)):
        torch.manual_seed(0)
        model = fa.MultiModalTransformer(170, 256, 2, 1, 1, 512, 0.1, text_len, 196).eval().to(dev)
        for B in (256, 1024, 4096, 16384):
            g = torch.Generator(device=dev).manual_seed(B + s)
            img = torch.randn(min(B, 4096), 1024, 14, 14, device=dev, generator=g).relu_()
            tokens = fa.project_images(model, img)
            if B > tokens.shape[0]:
                tokens = tokens.repeat(B // tokens.shape[0], 1, 1)
            src = torch.randint(0, 170, (B, min(s, 60)), device=dev, generator=g)
            for _ in range(2):
                fa.greedy_decode(model, None, src, 0, 20, dev, img_tokens=tokens)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            iters = 3
            e0.record()
            for _ in range(iters):
                fa.greedy_decode(model, None, src, 0, 20, dev, img_tokens=tokens)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            print(f"| {B} | {min(s, 60)} | {196 + min(s, 60)} | {ms:.2f} | {B / ms * 1e3:,.0f} | "
                  f"{B * MFLOP[s] * 1e6 / (ms * 1e-3) / 1e12:.1f} |")
        del model


if __name__ == "__main__":
    main()
