"""Host time of every `submit_inference_chain_host` call (BASELINE config-3 size) against the wall time of the whole
pipelined job: says whether the host (launch enqueue) or the device bounds the pipelined host-buffer path.

    python tools/fa_submit_trace.py [depth] [chunk] [steps]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from explainable_spatial_vqa_b200 import inference_transformer_full_annotation_new as fa  # noqa: E402
from explainable_spatial_vqa_b200 import synthetic as syn  # noqa: E402

depth = int(sys.argv[1]) if len(sys.argv) > 1 else 3
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
B = 4096
torch.manual_seed(0)
m = fa.MultiModalTransformer(170, 256, 2, 1, 1, 512, 0.1, 50, 196).eval().cuda()
func, deps, n_steps = syn.fa_programs(B, seed=4321)
img = torch.randn(B, 1024, 14, 14).relu_().pin_memory()
f, d, n = func.pin_memory(), deps.pin_memory(), n_steps.pin_memory()
outs = [torch.empty(B, f.shape[1], 20, dtype=torch.int32).pin_memory() for _ in range(depth)]
for rep in range(2):
    torch.cuda.synchronize()
    host = []
    t0 = time.perf_counter()
    for k in range(steps):
        a = time.perf_counter()
        fa.submit_inference_chain_host(m, img, f, d, n, 0, 20, chunk=chunk, depth=depth, out=outs[k % depth])
        host.append(1e3 * (time.perf_counter() - a))
    t1 = time.perf_counter()
    m.drain_host()
    t2 = time.perf_counter()
print(f"depth {depth} chunk {chunk} (CUDA_DEVICE_MAX_CONNECTIONS={os.environ.get('CUDA_DEVICE_MAX_CONNECTIONS')}): "
      f"host ms per submit {[round(x, 1) for x in host]}; all submitted after {1e3 * (t1 - t0):.1f} ms, drained after "
      f"{1e3 * (t2 - t0):.1f} ms = {1e3 * (t2 - t0) / steps:.1f} ms per step", flush=True)
