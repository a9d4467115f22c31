"""sharding by question must not change a single bit.  Two ranks under torchrun on two GPUs
run their contiguous halves of one question set; the gathered answers + programs (and FA caches) must equal the
single-GPU run of the whole set bitwise.  Skipped on a one-GPU box (the gloo test in test_host_logic.py covers the
host-side logic there)."""
import os
import subprocess
import sys

import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys, torch
sys.path.insert(0, {repo!r})
sys.path.insert(0, os.path.join({repo!r}, "tests"))
import torch.distributed as dist
import common
from explainable_spatial_vqa_b200 import sharding, synthetic as syn
from explainable_spatial_vqa_b200 import inference_transformer_full_annotation_new as fa
rank, local, world = sharding.init_from_env("nccl")
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
N = 300
img, q = syn.iqap_inputs(N, seed=41)
model = common.seeded_iqap().to(dev)
lo, hi = sharding.shard_range(N, rank, world)
ans, prog = model(img[lo:hi].to(dev), q[lo:hi].to(dev))
counts = [sharding.shard_range(N, r, world)[1] - sharding.shard_range(N, r, world)[0] for r in range(world)]
both = sharding.gather_varlen(torch.cat([ans.view(torch.int32).long(), prog], dim=1), counts)   # logits bit patterns + tokens
func, deps, n_steps = syn.fa_programs(64, seed=42, max_steps=6)
g = torch.Generator().manual_seed(43)
fimg = torch.randn(64, 1024, 14, 14, generator=g).relu_()
fmodel = common.seeded_fa().to(dev)
flo, fhi = sharding.shard_range(64, rank, world)
cache = fa.run_inference_chain_batched(fmodel, fimg[flo:fhi].to(dev), func[flo:fhi], deps[flo:fhi], n_steps[flo:fhi], 0, 20)
fcounts = [sharding.shard_range(64, r, world)[1] - sharding.shard_range(64, r, world)[0] for r in range(world)]
fall = sharding.gather_varlen(cache.reshape(cache.shape[0], -1), fcounts)
if rank == 0:
    a1, p1 = model(img.to(dev), q.to(dev))          # the whole set on one GPU
    want = torch.cat([a1.view(torch.int32).long(), p1], dim=1)
    assert torch.equal(both, want), "IQAP: gathered 2-GPU results differ from the 1-GPU run"
    c1 = fa.run_inference_chain_batched(fmodel, fimg.to(dev), func, deps, n_steps, 0, 20)
    assert torch.equal(fall, c1.reshape(64, -1)), "FA: gathered 2-GPU caches differ from the 1-GPU run"
dist.barrier()
dist.destroy_process_group()
os.write(1, ("rank %d ok\n" % rank).encode())
"""


@pytest.mark.gpu
def test_two_gpu_results_are_bitwise_equal_to_one_gpu(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(repo=REPO))
    port = 29600 + (os.getpid() % 300)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "rank 0 ok" in out.stdout and "rank 1 ok" in out.stdout


@pytest.mark.gpu
def test_two_handles_on_two_devices_in_one_process():
    """kernel attributes (dynamic shared memory opt-in) and the SM count are per device - a second handle on
    another GPU of the SAME process must work, and both devices must give identical results."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    sys.path.insert(0, os.path.join(REPO, "tests"))
    import common
    from explainable_spatial_vqa_b200 import synthetic as syn
    img, q = syn.iqap_inputs(40, seed=3)
    m0 = common.seeded_iqap().to("cuda:0")
    m1 = common.seeded_iqap().to("cuda:1")
    a0, p0 = m0(img.to("cuda:0"), q.to("cuda:0"))
    a1, p1 = m1(img.to("cuda:1"), q.to("cuda:1"))
    a0b, p0b = m0(img.to("cuda:0"), q.to("cuda:0"))   # and back on the first device
    torch.cuda.synchronize("cuda:0")
    torch.cuda.synchronize("cuda:1")
    assert torch.equal(a0.cpu(), a1.cpu()) and torch.equal(p0.cpu(), p1.cpu())
    assert torch.equal(a0, a0b) and torch.equal(p0, p0b)
    f0 = common.seeded_fa().to("cuda:1")
    func, deps, n_steps = syn.fa_programs(8, seed=5, max_steps=3)
    fimg = torch.randn(8, 1024, 14, 14).relu_()
    from explainable_spatial_vqa_b200 import inference_transformer_full_annotation_new as fa
    c1 = fa.run_inference_chain_batched(f0, fimg.to("cuda:1"), func, deps, n_steps, 0, 20)
    assert int(c1[:, 0, 0].max()) == 0
