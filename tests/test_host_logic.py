"""CPU: host-side logic of the drop-in surface - state-dict compatibility, vocabulary / chain parsing, the
sharding helpers, and the world_size-2 gather over gloo."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import common
from explainable_spatial_vqa_b200 import inference_transformer_full_annotation_new as fa
from explainable_spatial_vqa_b200 import inference_transformer_iqap as iqap
from explainable_spatial_vqa_b200 import sharding
from oracle import executor_oracle as orc

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_iqap_state_dict_layout():
    """names/shapes of the reference's VQAModel(85,256,256,32,44,27,196)."""
    sd = common.seeded_iqap().state_dict()
    assert len(sd) == 61 and sum(p.numel() for p in common.seeded_iqap().parameters()) == 4853580
    want = {"cls_token": (1, 1, 256), "image_proj.weight": (256, 1024), "embedding.weight": (85, 256),
            "pos_encoder.pe": (243, 1, 256), "pos_decoder.pe": (28, 1, 256),
            "transformer_encoder.layers.0.self_attn.in_proj_weight": (768, 256),
            "transformer_encoder.layers.0.linear1.weight": (2048, 256),
            "transformer_decoder.layers.1.multihead_attn.out_proj.weight": (256, 256),
            "transformer_decoder.layers.0.norm3.bias": (256,),
            "answer_classifier.0.weight": (256, 256), "answer_classifier.3.weight": (32, 256),
            "program_decoder_embedding.weight": (44, 256), "program_output.weight": (44, 256)}
    for k, shape in want.items():
        assert tuple(sd[k].shape) == shape, k
    assert float(sd["embedding.weight"][0].abs().sum()) == 0.0  # padding_idx row
    g = common.load_golden("iqap_b4.npz")
    assert sorted(sd) == [str(k) for k in g["sd_keys"]]  # the reference's own key set


def test_fa_state_dict_layout():
    m = common.seeded_fa()
    sd = m.state_dict()
    assert len(sd) == 41 and sum(p.numel() for p in m.parameters()) == 1668522
    for k, shape in {"pos_encoder.pe": (1, 246, 256), "pos_decoder.pe": (1, 50, 256),
                     "transformer.encoder.norm.weight": (256,), "transformer.decoder.norm.bias": (256,),
                     "transformer.decoder.layers.0.linear1.weight": (512, 256), "output_linear.weight": (170, 256),
                     "text_embedding.weight": (170, 256)}.items():
        assert tuple(sd[k].shape) == shape, k
    g = common.load_golden("fa_nhead2.npz")
    assert sorted(sd) == [str(k) for k in g["sd_keys"]]


def test_positional_encoding_and_mask():
    pe = iqap.PositionalEncoding(256, max_len=30).pe
    assert pe.shape == (30, 1, 256)
    p, i = 7, 5
    w = np.exp(-(2 * i) * np.log(10000.0) / 256)
    assert abs(float(pe[p, 0, 2 * i]) - np.sin(p * w)) < 1e-6 and abs(float(pe[p, 0, 2 * i + 1]) - np.cos(p * w)) < 1e-6
    assert torch.allclose(fa.PositionalEncoding(256, max_len=30).pe[0], pe[:, 0], atol=1e-6)
    m = iqap.generate_square_subsequent_mask(4)
    assert m[2, 3] == float("-inf") and m[3, 2] == 0 and m[1, 1] == 0
    x = torch.zeros(5, 2, 256)
    assert torch.equal(iqap.PositionalEncoding(256, max_len=30).eval()(x)[:, 0], pe[:5, 0])


def test_vocab_helpers(tmp_path):
    vocab = {"<PAD>": 0, "<UNK>": 1, "0": 2, "1": 3, "[": 4, "]": 5, "scene": 6}
    p = tmp_path / "vocab.json"
    p.write_text(json.dumps(vocab))
    v, rev = fa.load_vocab(str(p))
    assert v == vocab and rev[6] == "scene" and rev[2] == "0"
    assert fa.decode_tokens([6, 4, 99], rev) == "scene [ <unk>"
    assert fa.tokenize_field("filter_color", "function") == ["filter_color"]
    assert fa.tokenize_field("", "function") == []
    assert fa.tokenize_field("[0.3 0.1] red [ 0.5]", "output") == ["[", "0.3", "0.1", "]", "red", "[", "0.5", "]"]


def test_chain_parsing_follows_reference_rules(caplog):
    rev = {2: "0", 3: "1", 4: "2", 50: "scene", 51: "filter", 60: "red"}
    chain = ["50", "51 2", "51 3 60", "52 2 4"]
    func, deps = fa.chain_to_arrays(chain, rev)
    assert func.tolist() == [50, 51, 51, 52]
    assert deps.tolist() == [[-1, -1], [0, -1], [1, -1], [0, 2]]  # non-digit token 60 skipped with a warning
    assert any("not recognized as a digit" in r.message for r in caplog.records)
    with pytest.raises(ValueError):
        fa.chain_to_arrays(["50 2 3 4"], rev)
    # same rule set as the oracle's parser
    assert [d for _, d in orc.parse_chain(chain, rev)] == [[], [0], [1], [0, 2]]


def test_synthetic_programs_are_clevr_shaped():
    func, deps, n = orc.fa_programs(64, seed=4321)
    assert func.shape == (64, 25) and deps.shape == (64, 25, 2)
    assert int(n.min()) >= 2 and int(n.max()) <= 25
    for b in range(64):
        assert deps[b, 0].tolist() == [-1, -1]
        for i in range(int(n[b])):
            assert all(int(d) < i for d in deps[b, i])
            assert 27 <= int(func[b, i]) < 67
    strings = orc.chain_strings(func[0], deps[0], n[0])
    f2, d2 = fa.chain_to_arrays(strings, orc.fa_vocab(170))
    assert f2.tolist() == func[0, : int(n[0])].tolist() and d2.tolist() == deps[0, : int(n[0])].tolist()


def test_shard_ranges_partition_the_batch():
    for n in (0, 1, 7, 1024, 150000):
        for w in (1, 2, 4, 8):
            parts = [sharding.shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1
    costs = [int(c) for c in orc.fa_programs(200, seed=1)[2]]
    parts = sharding.balanced_ranges(costs, 8)
    assert parts[0][0] == 0 and parts[-1][1] == 200 and all(parts[i][1] == parts[i + 1][0] for i in range(7))
    loads = [sum(costs[lo:hi]) for lo, hi in parts]
    assert max(loads) <= 1.25 * (sum(costs) / 8) + 25
    with pytest.raises(ValueError):
        sharding.shard_range(10, 4, 4)


GLOO_WORKER = r"""
import os, sys, torch
sys.path.insert(0, {repo!r})
import torch.distributed as dist
from explainable_spatial_vqa_b200 import sharding
rank, local, world = sharding.init_from_env("gloo")
N = 11
lo, hi = sharding.shard_range(N, rank, world)
full = torch.arange(N * 3, dtype=torch.int64).view(N, 3)
mine = full[lo:hi].clone()                       # this rank's "programs"
counts = [sharding.shard_range(N, r, world)[1] - sharding.shard_range(N, r, world)[0] for r in range(world)]
got = sharding.gather_varlen(mine, counts)
assert torch.equal(got, full), (rank, got)
# the job-level gather: every step's rows stay local, ONE collective at the end
jg = sharding.JobGather(steps=3, rows=4, width=2, dtype=torch.int64, device="cpu")
for st in range(3):
    jg.put(st, torch.full((4, 2), 100 * rank + st, dtype=torch.int64))
allr = jg.finish()
assert tuple(allr.shape) == (world, 3, 4, 2)
for r in range(world):
    for st in range(3):
        assert bool((allr[r, st] == 100 * r + st).all()), (rank, r, st)
assert sharding.max_over_ranks(float(rank + 1)) == float(world)
assert sharding.sum_over_ranks(1.0) == float(world)
dist.barrier()
dist.destroy_process_group()
os.write(1, ("rank %d ok\n" % rank).encode())  # one write() per rank: the two ranks share the pipe
"""


def test_world_size_2_gather_over_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER.format(repo=REPO))
    port = 29500 + (os.getpid() % 400)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "rank 0 ok" in out.stdout and "rank 1 ok" in out.stdout


def test_bench_reference_arm_runs_on_cpu():
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-sample", "2"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["unit"] == "program-steps/s"
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0


def test_modules_survive_deepcopy_and_pickle():
    """Native handles never travel through copy / pickle; the copy rebuilds its own lazily."""
    import copy
    import io
    m = common.seeded_iqap()
    m2 = copy.deepcopy(m)
    assert m2._pool._module is m2 and m2._pool._handles == {}
    buf = io.BytesIO()
    torch.save(m.state_dict(), buf)
    buf.seek(0)
    m2.load_state_dict(torch.load(buf))
    f2 = copy.deepcopy(common.seeded_fa())
    assert f2._pool._module is f2


def test_oracle_tally_follows_reference_loop():
    """oracle.iqap_tally restates inference_transformer_iqap_tally.py:317-344: torch.max keeps the FIRST maximum, a
    program is correct only if all 27 tokens match, and the four cases partition the samples."""
    import torch
    from oracle import executor_oracle as orc
    logits = torch.tensor([[0.0, 2.0, 2.0], [1.0, 0.0, 0.0], [0.0, 0.0, 3.0], [5.0, 5.0, 5.0]])
    progs = torch.arange(4 * 27).view(4, 27)
    gt_p = progs.clone()
    gt_p[1, 26] += 1          # last token differs
    gt_p[3, 0] += 1
    gt_a = torch.tensor([1, 0, 0, 1])   # sample 0: tie -> class 1 (first max) correct; 2: wrong; 3: tie -> class 0, wrong
    counts, preds = orc.iqap_tally(logits, progs, gt_a, gt_p)
    assert preds == [1, 0, 2, 0]
    assert counts == [1, 1, 1, 1]


def test_cpulist_parsing_and_numa_binding_is_harmless_without_a_gpu():
    assert sharding._parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert sharding._parse_cpulist("") == []
    info = sharding.bind_to_gpu_numa(0)   # no GPU here: reports the failure instead of raising
    assert info["bound"] is False


def test_host_fp32_to_fp16_conversion_is_numpy_exact():
    """b200vqa_host_f32_to_f16 (the "fp16" upload mode of the host entry points): round to nearest even like numpy /
    torch .half(), including ties, subnormals, overflow to inf and the scalar tail; any thread count."""
    import ctypes as C

    import numpy as np

    from explainable_spatial_vqa_b200 import _native as nat

    rng = np.random.default_rng(5)
    n = (1 << 19) + 37  # several pool tasks + a tail that is not a multiple of the vector width
    x = (rng.standard_normal(n) * np.exp(rng.uniform(-14, 12, n))).astype(np.float32)
    x[:8] = [0.0, -0.0, 65504.0, 65520.0, 1e-8, 5.96e-8, 1.0009765625, 2049.0]  # max, overflow tie, subnormals, ties
    want = x.astype(np.float16)
    lib = nat.lib()
    for threads in (1, 3, 0):
        raw = np.empty(n + 16, dtype=np.uint16)
        off = (-raw.ctypes.data % 32) // 2  # 32-byte aligned start inside the buffer
        out = raw[off:off + n]
        assert out.ctypes.data % 32 == 0
        nat.check(lib.b200vqa_host_f32_to_f16(x.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), n, threads), "convert")
        assert np.array_equal(out.view(np.float16).view(np.uint16), want.view(np.uint16)), threads
    assert lib.b200vqa_host_f32_to_f16(x.ctypes.data_as(C.c_void_p), C.c_void_p(raw.ctypes.data + 2), 4, 1) < 0  # misaligned
    # bf16 (the FA host entry's upload mode): the same rounding as torch / the device conversion
    import torch
    want16 = torch.from_numpy(x).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    for threads in (1, 0):
        raw = np.empty(n + 16, dtype=np.uint16)
        off = (-raw.ctypes.data % 32) // 2
        out = raw[off:off + n]
        nat.check(lib.b200vqa_host_f32_to_bf16(x.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), n, threads), "convert")
        assert np.array_equal(out, want16), threads
