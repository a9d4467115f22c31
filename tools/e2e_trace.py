"""Host time of every `submit_host` call against the wall time of the pipelined end-to-end job (IQAP, 1024 questions per
step, fp32 host features): python tools/e2e_trace.py [upload] [chunk] [steps] [depth]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from explainable_spatial_vqa_b200 import inference_transformer_iqap as iq  # noqa: E402
from explainable_spatial_vqa_b200 import synthetic as syn  # noqa: E402

upload = sys.argv[1] if len(sys.argv) > 1 else "fp16"
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
depth = int(sys.argv[4]) if len(sys.argv) > 4 else 2
background = None if len(sys.argv) <= 5 else sys.argv[5] == "bg"
torch.manual_seed(0)
m = iq.VQAModel(85, 256, 256, 32, 44, 27, 196).eval().cuda()
img, q = syn.iqap_inputs(1024, seed=1)
img, q = img.pin_memory(), q.pin_memory()
for rep in range(3):
    for _ in range(3):
        m.submit_host(img, q, chunk=chunk, depth=depth, upload=upload, background=background)
    m.drain_host()
    torch.cuda.synchronize()
    host = []
    t0 = time.perf_counter()
    for _ in range(steps):
        a = time.perf_counter()
        m.submit_host(img, q, chunk=chunk, depth=depth, upload=upload, background=background)
        host.append(round(1e3 * (time.perf_counter() - a), 1))
    t1 = time.perf_counter()
    m.drain_host()
    t2 = time.perf_counter()
    print(f"upload {upload} chunk {chunk} depth {depth}: host ms per submit {host}; submitted after {1e3 * (t1 - t0):.1f} ms, drained "
          f"after {1e3 * (t2 - t0):.1f} ms = {1e3 * (t2 - t0) / steps:.1f} ms per step", flush=True)
