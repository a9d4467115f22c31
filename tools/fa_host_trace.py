"""b200vqa_fa_run_chain_host at the BASELINE config-3 size: wall time per (parts, chunk) and, with `--trace`, the GPU
timeline of one call (B200VQA_FA_HOST_TRACE=1)."""
import os
import sys
import time

if "--trace" in sys.argv:
    os.environ["B200VQA_FA_HOST_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from explainable_spatial_vqa_b200 import inference_transformer_full_annotation_new as fa  # noqa: E402
from explainable_spatial_vqa_b200 import synthetic as syn  # noqa: E402

B = 4096
torch.manual_seed(0)
m = fa.MultiModalTransformer(170, 256, 2, 1, 1, 512, 0.1, 50, 196).eval().cuda()
func, deps, n_steps = syn.fa_programs(B, seed=4321)
img = torch.randn(B, 1024, 14, 14).relu_().pin_memory()
f, d, n = func.pin_memory(), deps.pin_memory(), n_steps.pin_memory()
steps = int(n_steps.sum())
ref = None
configs = [(1, 4096), (1, 2048), (2, 2048), (2, 1024), (2, 512), (4, 1024), (4, 512)] if "--trace" not in sys.argv else [(2, 1024)]
for parts, chunk in configs:
    best = 1e9
    for rep in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fa.run_inference_chain_host(m, img, f, d, n, 0, 20, chunk=chunk, parts=parts)
        dt = time.perf_counter() - t0
        if rep:
            best = min(best, dt)
    if ref is None:
        ref = out.clone()
    print(f"parts {parts} chunk {chunk}: {1e3 * best:.1f} ms -> {steps / best / 1e3:.0f} k program-steps/s; identical to the first config: {bool(torch.equal(out, ref))}",
          flush=True)
