"""Times the absorbed cross-attention kernel (b200vqa_dbg_mem_attn) alone on cuda:0: CUDA events around `reps` back-to-back launches over `nbuf` rotating memory buffers (> L2 in total).

    python tools/microbench_mem_attn.py [B ...]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from explainable_spatial_vqa_b200 import _native as nat  # noqa: E402


def run(B, impl, nhead=4, length=243, reps=40):
    nbuf = max(2, (400 << 20) // (B * 256 * 512) + 1)
    mems = [torch.randn(B * 256, 256, device="cuda").bfloat16() for _ in range(nbuf)]
    qp = (torch.randn(B, nhead * 256, device="cuda") * 0.25).bfloat16()
    out = torch.empty(B, nhead * 256, dtype=torch.bfloat16, device="cuda")
    lib, st = nat.lib(), nat.stream_ptr()

    def launch(i):
        nat.check(lib.b200vqa_dbg_mem_attn(qp.data_ptr(), mems[i % nbuf].data_ptr(), None, length, B, nhead, impl,
                                           out.data_ptr(), None, st), "dbg_mem_attn")
    for i in range(5):
        launch(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps):
        launch(i)
    b.record()
    torch.cuda.synchronize()
    us = a.elapsed_time(b) * 1e3 / reps
    gbs = B * length * 512 / us / 1e3
    return us, gbs


if __name__ == "__main__":
    sizes = [int(x) for x in sys.argv[1:]] or [128, 256, 512, 1024, 2048, 4096]
    for B in sizes:
        for length in (243, 217):
            us, gbs = run(B, 0, length=length)
            print(f"B {B:5d} len {length} {us:8.2f} us/launch {gbs:8.1f} GB/s algorithmic "
                  f"({gbs / 6541.1:.3f} of the measured HBM copy peak)", flush=True)
