"""Host fp32 -> fp16 conversion rate (b200vqa_host_f32_to_f16) against the pinned host->device copy rate: decides whether
rounding the features on the host before the upload pays on this machine.

    python tools/microbench_host_convert.py
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from explainable_spatial_vqa_b200 import _native as nat  # noqa: E402

n = 512 * 196 * 1024  # one 512-question chunk of IQAP features
pin = (lambda t: t.pin_memory()) if torch.cuda.is_available() else (lambda t: t)
src = pin(torch.randn(n).relu_())
dst = pin(torch.empty(n, dtype=torch.float16))
lib = nat.lib()
for threads in (1, 2, 4, 8, 12, 16, 0):
    best = 1e9
    for _ in range(4):
        t0 = time.perf_counter()
        nat.check(lib.b200vqa_host_f32_to_f16(src.data_ptr(), dst.data_ptr(), n, threads), "convert")
        best = min(best, time.perf_counter() - t0)
    print(f"threads {threads:2d}: {1e3 * best:7.2f} ms per 512-question chunk = {n * 4 / best / 1e9:6.1f} GB/s of fp32 in", flush=True)
print("bit-exact with torch .half():", bool(torch.equal(dst, src.half())))
if torch.cuda.is_available():
    d32 = torch.empty(n, device="cuda")
    d16 = torch.empty(n, dtype=torch.float16, device="cuda")
    for name, h, d in (("fp32", src, d32), ("fp16", dst, d16)):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(4):
            d.copy_(h, non_blocking=True)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 4
        print(f"H2D {name}: {1e3 * dt:.2f} ms per chunk = {h.numel() * h.element_size() / dt / 1e9:.1f} GB/s")

# ---- the two together: does the conversion starve the DMA engine of host memory bandwidth?
if torch.cuda.is_available():
    import threading

    for threads in (4, 8, 12, 16):
        stop = False
        done = [0]

        def conv_loop():
            while not stop:
                nat.check(lib.b200vqa_host_f32_to_f16(src.data_ptr(), dst.data_ptr(), n, threads), "convert")
                done[0] += 1

        h16 = pin(torch.empty(n, dtype=torch.float16))
        th = threading.Thread(target=conv_loop)
        th.start()
        time.sleep(0.05)
        torch.cuda.synchronize()
        c0 = done[0]
        t0 = time.perf_counter()
        for _ in range(8):
            d16.copy_(h16, non_blocking=True)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 8
        c1 = done[0]
        stop = True
        th.join()
        print(f"conversion on {threads:2d} threads running: H2D fp16 {1e3 * dt:.2f} ms per chunk = {h16.numel() * 2 / dt / 1e9:.1f} GB/s; "
              f"conversions completed meanwhile: {c1 - c0} ({(c1 - c0) * n * 4 / (8 * dt) / 1e9:.0f} GB/s of fp32 in)", flush=True)
