#!/usr/bin/env python
"""BASELINE config 5: CoGenT-shaped sweep of the FA executor step - batch 256..16k x sequence 197..256 tokens
(src length 1 / 21 / 41 / 60) on 1 / 2 / 4 / 8 GPUs, with the host-CPU reference (oracle, the reference's batch-1 loop)
timed at batch 256.  Prints a markdown table (program-steps/s over all ranks, ms per step, achieved TFLOP/s from the
algorithmic FLOPs of SURVEY §8d).

    python tools/sweep.py                                   # one GPU, CPU column included
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/sweep.py   # N GPUs (weak scaling)

Every rank runs the same (batch, length) point on its own questions (no collective in the data path); a point's time is
the max over ranks (CUDA events), its throughput the sum of the ranks' batches over that time.
"""
import os
import sys
import time
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import torch  # noqa: E402

from explainable_spatial_vqa_b200 import inference_transformer_full_annotation_new as fa  # noqa: E402
from explainable_spatial_vqa_b200 import sharding  # noqa: E402

MFLOP = {1: 328.5, 21: 363.6, 41: 399.5, 60: 434.6}  # per program-step, SURVEY §8d (60 source tokens: L = 256, the row cap)


def cpu_steps_per_s(s, n_questions=3):
    """The reference's only mode (batch 1): oracle greedy decode of `n_questions` questions, all host threads."""
    from oracle import executor_oracle as orc
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    sd = fa.MultiModalTransformer(170, 256, 2, 1, 1, 512, 0.1, 64, 196).eval().state_dict()
    g = torch.Generator().manual_seed(s)
    img = torch.randn(n_questions, 1024, 14, 14, generator=g).relu_()
    src = torch.randint(0, 170, (n_questions, s), generator=g)
    orc.fa_greedy_decode(sd, img[:1], src[:1], 0, 20, 2, recompute=True)
    t0 = time.perf_counter()
    for b in range(n_questions):
        orc.fa_greedy_decode(sd, img[b:b + 1], src[b:b + 1], 0, 20, 2, recompute=True)
    return n_questions / (time.perf_counter() - t0)


def main():
    rank, local, world = sharding.init_from_env("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if rank == 0:
        print(f"# FA executor step sweep (BASELINE config 5 shape), {world} x B200, tools/sweep.py\n")
        print("One program step = encoder over [196 image tokens + src tokens] + 19-position greedy decode; image tokens "
              "cached; batch = questions per GPU (weak scaling).  Sequence rows are capped at 256 per question, so the "
              "survey's L = 260 point runs at 60 source tokens (L = 256).\n")
        print("| GPUs | batch / GPU | src tokens | seq L | ms / step | program-steps/s (all GPUs) | algorithmic TFLOP/s | host-CPU reference steps/s (batch 1) |")
        print("|---:|---:|---:|---:|---:|---:|---:|---:|")
    for s in (1, 21, 41, 60):
        torch.manual_seed(0)
        model = fa.MultiModalTransformer(170, 256, 2, 1, 1, 512, 0.1, 64, 196).eval().to(dev)
        cpu = cpu_steps_per_s(s) if (rank == 0 and world == 1 and "--no-cpu" not in sys.argv) else None
        for B in (256, 1024, 4096, 16384):
            g = torch.Generator(device=dev).manual_seed(B + s + 1000 * rank)
            img = torch.randn(min(B, 4096), 1024, 14, 14, device=dev, generator=g).relu_()
            tokens = fa.project_images(model, img)
            if B > tokens.shape[0]:
                tokens = tokens.repeat(B // tokens.shape[0], 1, 1)
            src = torch.randint(0, 170, (B, s), device=dev, generator=g)
            for _ in range(2):
                fa.greedy_decode(model, None, src, 0, 20, dev, img_tokens=tokens)
            torch.cuda.synchronize()
            if world > 1:
                torch.distributed.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            iters = 3
            e0.record()
            for _ in range(iters):
                fa.greedy_decode(model, None, src, 0, 20, dev, img_tokens=tokens)
            e1.record()
            torch.cuda.synchronize()
            ms = sharding.max_over_ranks(e0.elapsed_time(e1) / iters, dev)
            if rank == 0:
                cpu_col = f"{cpu:,.1f}" if (cpu is not None and B == 256) else ""
                print(f"| {world} | {B} | {s} | {196 + s} | {ms:.2f} | {world * B / ms * 1e3:,.0f} | "
                      f"{world * B * MFLOP[s] * 1e6 / (ms * 1e-3) / 1e12:.1f} | {cpu_col} |", flush=True)
            del img, tokens, src
        del model
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
