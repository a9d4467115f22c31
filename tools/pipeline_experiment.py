import sys, os, warnings
sys.path.insert(0, "/root/repo"); warnings.filterwarnings("ignore")
import torch
from explainable_spatial_vqa_b200 import inference_transformer_iqap as iqap
from oracle import executor_oracle as orc
dev = torch.device("cuda")
B = 1024
g = torch.Generator(device=dev).manual_seed(1234)
img = torch.randn(B, 196, 1024, device=dev, generator=g).relu_()
_, q = orc.iqap_inputs(B, seed=1234, relu=False); q = q.to(dev)
for depth in (1, 2, 3):
    models, streams = [], []
    for i in range(depth):
        torch.manual_seed(0)
        models.append(iqap.VQAModel(85, 256, 256, 32, 44, 27, 196).eval().to(dev))
        streams.append(torch.cuda.Stream())
    torch.cuda.synchronize()
    def run(K):
        outs = []
        for k in range(K):
            s = streams[k % depth]
            with torch.cuda.stream(s):
                outs.append(models[k % depth](img, q))
        return outs
    run(2 * depth); torch.cuda.synchronize()
    K = 12
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in streams: s.wait_stream(torch.cuda.current_stream())
    outs = run(K)
    for s in streams: torch.cuda.current_stream().wait_stream(s)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    print(f"depth {depth}: {ms:.3f} ms/step -> {B*27/ms*1e3/1e6:.2f} M program-steps/s")
    ref = models[0](img, q)
    assert all(torch.equal(o[1], ref[1]) and torch.equal(o[0], ref[0]) for o in outs)
    del models
