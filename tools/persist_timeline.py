"""Per-stage timeline of the persistent decode kernel (CTA (0,0), worker thread 0): B200VQA_PERSIST_STAMPS=1."""
import os
import sys

os.environ["B200VQA_DECODE"] = "persist"
os.environ["B200VQA_PERSIST_STAMPS"] = "1"
os.environ["B200VQA_PERSIST_VERBOSE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from explainable_spatial_vqa_b200 import inference_transformer_iqap as iqap  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
torch.manual_seed(0)
m = iqap.VQAModel(85, 256, 256, 32, 44, 27, 196).eval().cuda()
g = torch.Generator(device="cuda").manual_seed(1)
img = torch.randn(B, 196, 1024, device="cuda", generator=g).relu_()
q = torch.randint(1, 85, (B, 46), device="cuda")
for _ in range(3):
    m(img, q)
torch.cuda.synchronize()
clk = m._native(0).dbg_workspace(11, torch.int64).view(8, 24).cpu()
names = ["start", "A copied", "acc G1", "epi G1", "rv1", "rv2 (self-attn)", "rv3 (G3)", "rv4 (LN1)", "acc G5", "rv5",
         "R6 done", "rv6", "rv7 (G7)", "rv9 (G8+LN2)", "rv10 (FFN)", "rv11 (reduce+head)"]
khz = torch.cuda.get_device_properties(0).clock_rate if hasattr(torch.cuda.get_device_properties(0), "clock_rate") else 1965000
print(f"batch {B}; cycles -> us at {khz / 1e3:.0f} MHz")
for st in range(8):
    row = clk[st]
    t0 = int(row[0])
    parts = []
    prev = t0
    for i in range(1, 16):
        parts.append(f"{names[i]} +{(int(row[i]) - prev) / khz * 1e3:.1f}")
        prev = int(row[i])
    nxt = int(clk[st + 1][0]) if st + 1 < 8 else prev
    print(f"stage {st}: total {(prev - t0) / khz * 1e3:.1f} us | " + " | ".join(parts))
    print(f"    R6: queries {(int(row[18]) - int(row[17])) / khz * 1e3:.1f} us, tiles {(int(row[10]) - int(row[18])) / khz * 1e3:.1f} us "
          f"of which waiting for memory {int(row[16]) / khz * 1e3:.1f} us")
