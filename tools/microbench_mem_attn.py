"""Times the two absorbed cross-attention kernels (b200vqa_dbg_mem_attn: impl 0 = warp-MMA ring, 1 = tcgen05 cluster kernel with persistent clusters, 2 = one cluster per question,
3 = tcgen05 with one persistent CTA per SM and a three-stage tile ring)
alone on cuda:0: CUDA events around `reps` back-to-back launches over `nbuf` rotating memory buffers (> L2 in total).

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


def stamps(B, impl, nhead=4, length=243):
    """Stage stamps of one launch of the tcgen05 kernel: mean cycles between the stages of a (question, half) work item."""
    mem = torch.randn(B * 256, 256, device="cuda").bfloat16()
    qp = (torch.randn(B, nhead * 256, device="cuda") * 0.25).bfloat16()
    out = torch.empty(B, nhead * 256, dtype=torch.bfloat16, device="cuda")
    st = torch.zeros(2 * B, 16, dtype=torch.int64, device="cuda")
    lib = nat.lib()
    for _ in range(3):
        nat.check(lib.b200vqa_dbg_mem_attn(qp.data_ptr(), mem.data_ptr(), None, length, B, nhead, impl, out.data_ptr(),
                                           st.data_ptr(), nat.stream_ptr()), "dbg_mem_attn")
    torch.cuda.synchronize()
    s = st.cpu().double()
    names = ["top->tma_issued", "tma_issued->S_committed", "S_committed->S_seen", "S_seen->P_arrived", "P_arrived->U_seen",
             "U_seen->exchanged", "exchanged->stored"]
    cols = [(2, 3), (3, 4), (4, 5), (5, 6), (6, 7), (7, 8), (8, 9)]
    for it in sorted(set(s[:, 11].tolist()))[:3]:
        rows = s[s[:, 11] == it]
        print(f"  impl {impl} B {B} iteration {int(it)} ({len(rows)} work items): "
              + ", ".join(f"{n} {float((rows[:, b] - rows[:, a]).mean()):.0f}" for n, (a, b) in zip(names, cols))
              + f" | whole item {float((rows[:, 10] - rows[:, 1]).mean()):.0f} ns", flush=True)
    t0, t1 = s[:, 1].min(), s[:, 10].max()
    per_sm = {}
    for row in s.tolist():
        per_sm.setdefault(int(row[0]), []).append((row[1], row[10]))
    conc = []
    for sm, iv in per_sm.items():
        busy = sum(b - a for a, b in iv)
        conc.append(busy / (t1 - t0))
    print(f"  impl {impl} B {B}: span {float(t1 - t0) / 1e3:.1f} us, SMs used {len(per_sm)}, mean resident work items per SM "
          f"{sum(conc) / len(conc):.2f}, items per SM {2 * B / len(per_sm):.1f}", flush=True)


if __name__ == "__main__":
    sizes = [int(x) for x in sys.argv[1:]] or [128, 256, 512, 1024, 2048, 4096]
    for B in sizes:
        for impl in (0, 1, 2, 3):
            us, gbs = run(B, impl)
            print(f"B {B:5d} impl {('mma', 'tc ', 'tc1', 'tcr')[impl]} {us:8.2f} us/launch {gbs:8.1f} GB/s algorithmic", flush=True)
    for impl in (1, 2):
        stamps(1024, impl)
