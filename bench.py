#!/usr/bin/env python
"""Benchmark of the batched Program Executor inference path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload iqap|fa|e2e]

A "step" is one pass of the hot path over one batch of synthetic input on every GPU:
  iqap (default, BASELINE configs[1]): VQAModel.forward on 1024 questions per GPU = encoder + answer head +
       27 greedy program positions; 1 question = 27 program-steps.
  fa   (configs[2]): run_inference_chain_batched on 4096 questions per GPU with CLEVR-shaped ragged programs;
       1 chain element = 1 program-step.
  e2e  (configs[3]): LSTM program generator decode on the questions + device program->chain glue + the FA executor
       on 4096 questions per GPU (synthetic CLEVR-shaped prefix programs drive the executor: a random-init generator
       does not emit valid programs; its decode is executed and timed all the same).
`value` is whole-job program-steps/s with inputs resident in HBM (CUDA events, max over ranks); `e2e` is the
same metric through the public host-buffer call (pinned host inputs uploaded and results downloaded inside
the timed region).  N > 1: one process per GPU under torchrun, questions sharded, no per-step collective in
the data path; every step copies its results into a preallocated device buffer and ONE all_gather_into_tensor of the
whole job's results runs at the end, inside the timed region.

`--impl reference` times the reference's own CPU algorithm (oracle/executor_oracle.py in `recompute` mode: the
decoder prefix and the cross K/V are recomputed every step exactly as the reference's PyTorch modules do) on
the host cores, on a bounded sample of the same workload per step.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
import warnings

import torch

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
warnings.filterwarnings("ignore", message="enable_nested_tensor")

T_PROG = 27
S_IQAP = 243
D = 256


def load_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


# ----------------------------------------------------------------------------------------------------------
# algorithmic work per kernel class and per question (S=243 valid rows, d=256, ff=2048, T=27)
# ----------------------------------------------------------------------------------------------------------
def iqap_work_per_question(ff=2048, n_dec=2, S=S_IQAP, T=T_PROG, Vp=44, C=32):
    d = D
    w = {
        "image_proj_gemm": ("tensor", 2 * 196 * 1024 * d),
        "enc_qkv_gemm": ("tensor", 2 * S * d * 3 * d),
        "enc_attention": ("tensor", 4 * S * S * d),
        "enc_outproj_ln_gemm": ("tensor", 2 * S * d * d),
        "enc_ffn1_gemm": ("tensor", 2 * S * d * ff),
        "enc_ffn2_ln_gemm": ("tensor", 2 * S * d * ff),
        "enc_ffn_fused": ("tensor", 4 * S * d * ff),            # linear1 + linear2 in one kernel (the default)
        "dec_cross_kv_gemm": ("tensor", n_dec * 2 * S * d * 2 * d),
        # the three plain GEMMs of a decode position, one class per call site (each is one kernel at one shape)
        "dec_self_qkv_gemm": ("tensor", n_dec * T * 2 * d * 3 * d),               # self-attention in_proj
        "dec_cross_q_gemm": ("tensor", n_dec * T * 2 * d * d),                    # cross-attention query projection
        # per-head value projection of the attention-weighted memory (absorbed form only; replaces the K|V projection
        # of the memory, which the algorithmic count lists under dec_cross_kv_gemm)
        "dec_cross_v_gemm": ("tensor", n_dec * T * 2 * d * d),
        "dec_outproj_ln_gemm": ("tensor", n_dec * T * 2 * 2 * d * d),            # two out-proj + LayerNorm
        "dec_ffn_split": ("tensor", n_dec * T * 4 * d * ff),
        # HBM-bound: every decode position re-reads the encoder memory rows (bf16, 512 B per row) - the absorbed
        # form of SURVEY H2; with B200VQA_NO_ABSORB=1 it reads a projected K and a V row instead (twice the bytes)
        "dec_cross_attention": ("hbm", n_dec * T * S * d * 2 * (1 if ABSORB else 2)),
        "dec_self_attention": ("hbm", n_dec * sum((t + 1) * 2 * d * 2 for t in range(T))),
        "dec_head_argmax": ("tensor", T * 2 * d * Vp),
        "answer_head": ("hbm", d * 2 + C * 4),
        "embed_gather": ("hbm", (256 - 196) * d * 2),
    }
    return w


IQAP_FLOPS_PER_QUESTION = 1.0983e9
ABSORB = os.environ.get("B200VQA_NO_ABSORB", "0") in ("", "0")  # library default: absorbed cross-attention


class ClockSampler:
    """nvidia-smi style clock / throttle sampling DURING the timed region (NVML, 100 ms period)."""

    def __init__(self, index):
        self.samples, self.reasons, self.stop_flag = [], set(), threading.Event()
        self.max_mhz = None
        self.thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.dev = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover - NVML missing
            self.nv, self.err = None, str(e)

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                 "hw_power_brake": 0x80, "sync_boost": 0x10, "applications_clocks_setting": 0x2}
        while not self.stop_flag.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.dev, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.dev) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.dev)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def start(self):
        if self.nv is not None:
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()

    def stop(self):
        self.stop_flag.set()
        if self.thread is not None:
            self.thread.join()
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def cpu_model_name():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


# ----------------------------------------------------------------------------------------------------------
# CPU legs (oracle; rank 0 only)
# ----------------------------------------------------------------------------------------------------------
def cpu_iqap(sample_b, repeats, recompute=True):
    """program-steps/s of the oracle on `sample_b` questions of the workload, best of `repeats`."""
    from oracle import executor_oracle as orc
    from explainable_spatial_vqa_b200 import inference_transformer_iqap as iqap
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    sd = iqap.VQAModel(85, 256, 256, 32, 44, T_PROG, 196).eval().state_dict()
    img, q = orc.iqap_inputs(sample_b, seed=1234)
    orc.iqap_forward(sd, img[:4], q[:4], recompute=recompute)  # warm-up
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        orc.iqap_forward(sd, img, q, recompute=recompute)
        best = min(best, time.perf_counter() - t0)
    return sample_b * T_PROG / best, best


def cpu_fa(n_questions, recompute=True, with_generator=False):
    """program-steps/s of the oracle chain (batch-1 loop, the reference's only mode); `with_generator` adds the LSTM
    program generator's decode of the same questions (config 4)."""
    from oracle import executor_oracle as orc
    from explainable_spatial_vqa_b200 import inference_transformer_full_annotation_new as fa
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    sd = fa.MultiModalTransformer(170, 256, 2, 1, 1, 512, 0.1, 50, 196).eval().state_dict()
    func, deps, n_steps = orc.fa_programs(n_questions, seed=4321)
    g = torch.Generator().manual_seed(4321)
    img = torch.randn(n_questions, 1024, 14, 14, generator=g).relu_()
    rev = orc.fa_vocab(170)
    orc.fa_run_chain(sd, img[:1], orc.chain_strings(func[0], deps[0], 2), rev, 0, 20, 2, recompute=recompute)
    t0 = time.perf_counter()
    if with_generator:
        from oracle import lstm_oracle
        from explainable_spatial_vqa_b200 import run_model_lstm_qp as qp
        torch.manual_seed(1)
        gsd = qp.Seq2SeqModel(85, 256, 512, 44, T_PROG, 1).eval().state_dict()
        t0 = time.perf_counter()
        lstm_oracle.generate(gsd, lstm_oracle.questions(n_questions))
    for b in range(n_questions):
        orc.fa_run_chain(sd, img[b:b + 1], orc.chain_strings(func[b], deps[b], n_steps[b]), rev, 0, 20, 2,
                         recompute=recompute)
    dt = time.perf_counter() - t0
    return float(n_steps.sum()) / dt, dt


def gpu_eager_iqap(dev, sample_b):
    """Context only: the reference's algorithm (oracle, recompute-every-step form, fp32) run with PyTorch's own
    CUDA kernels on this GPU - what the unmodified reference would do with device='cuda'.  program-steps/s."""
    from oracle import executor_oracle as orc
    from explainable_spatial_vqa_b200 import inference_transformer_iqap as iqap
    torch.manual_seed(0)
    sd = {k: v.to(dev) for k, v in iqap.VQAModel(85, 256, 256, 32, 44, T_PROG, 196).eval().state_dict().items()}
    img, q = orc.iqap_inputs(sample_b, seed=1234)
    img, q = img.to(dev), q.to(dev)
    orc.iqap_forward(sd, img[:8], q[:8], recompute=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    orc.iqap_forward(sd, img, q, recompute=True)
    e1.record()
    torch.cuda.synchronize()
    return sample_b * T_PROG / (e0.elapsed_time(e1) * 1e-3)


def h2d_probe(dev, nbytes=512 << 20, barrier=None, repeats=3):
    """Pinned host -> device copy bandwidth of this box (GB/s): the ceiling of the e2e number.  With `barrier` every
    rank starts its copies together, so the result is the per-rank share of the host's PCIe / memory bandwidth."""
    src = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    dst = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    if barrier is not None:
        barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(repeats):
        dst.copy_(src, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    return repeats * nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9


def gpu_kvcached_bf16_iqap(dev, sample_b):
    """Context only: the reference's algorithm in its KV-cached form (cross K/V projected once, one new position per
    step) with PyTorch's own CUDA kernels under bf16 autocast on this GPU - the 'competent PyTorch' bar.  program-steps/s."""
    from oracle import executor_oracle as orc
    from explainable_spatial_vqa_b200 import inference_transformer_iqap as iqap
    torch.manual_seed(0)
    sd = {k: v.to(dev) for k, v in iqap.VQAModel(85, 256, 256, 32, 44, T_PROG, 196).eval().state_dict().items()}
    img, q = orc.iqap_inputs(sample_b, seed=1234)
    img, q = img.to(dev), q.to(dev)
    with torch.autocast(device_type="cuda", dtype=torch.bfloat16):
        orc.iqap_forward(sd, img[:8], q[:8], recompute=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orc.iqap_forward(sd, img, q, recompute=False)
        e1.record()
        torch.cuda.synchronize()
    return sample_b * T_PROG / (e0.elapsed_time(e1) * 1e-3)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = args.cpu_sample or (64 if args.workload == "iqap" else 2)
    cores = os.cpu_count() or 1
    times = []
    units = 0.0
    for i in range(args.warmup + args.steps):
        if args.workload == "iqap":
            v, dt = cpu_iqap(sample, 1)
            u = sample * T_PROG
        else:
            v, dt = cpu_fa(sample, with_generator=args.workload == "e2e")
            u = v * dt
        if i >= args.warmup:
            times.append(dt)
            units += u
    total = sum(times)
    value = units / total
    line = {
        "impl": "reference", "metric": "program_steps_per_s", "value": value, "unit": "program-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / max(1, len(times)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": "program-steps/s", "cores": cores, "kind": "port",
                         "cpu": cpu_model_name(),
                         "sample": f"{sample} questions of the workload per step, oracle in the reference's "
                                   "recompute-every-step form (torch CPU fp32, all host threads)"},
        "e2e": {"value": value, "unit": "program-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args):
    if args.workload == "iqap":
        return {"workload": f"IQAP executor (VQAModel 85/256/256/32/44, 1 enc + 2 dec layers, ff 2048), batch {args.batch} "
                            "questions per GPU, 27 program positions each, seq 243",
                "batch_per_gpu": args.batch, "program_len": T_PROG,
                "l2": "inputs are 822 MB of fp32 features per step (> 126 MB L2), activations 2.4 GB",
                "gather": "results stay on the device; one all_gather_into_tensor of the job's answers + programs (NCCL) at "
                          "the end of the timed region when n_gpus > 1"}
    if args.workload == "e2e":
        return {"workload": f"end to end: LSTM program generator (85/256/512/44, 46 -> 27 tokens) + device program->chain "
                            f"glue + FA executor, batch {args.batch} questions per GPU, synthetic CLEVR-shaped prefix "
                            "programs of 2..25 nodes drive the executor",
                "batch_per_gpu": args.batch,
                "l2": "per-step activations exceed L2 (4096 questions x 128 KB encoder rows)",
                "gather": "results stay on the device; one all_gather_into_tensor of the job's final-step tokens (NCCL) at "
                          "the end of the timed region when n_gpus > 1"}
    return {"workload": f"FA executor (MultiModalTransformer V=170 nhead 2, 1+1 layers, ff 512), batch {args.batch} "
                        "questions per GPU, CLEVR-shaped ragged programs of 2..25 steps, 20 tokens per step",
            "batch_per_gpu": args.batch,
            "l2": "per-step activations exceed L2 (4096 questions x 128 KB encoder rows)",
            "gather": "per step: all_gather of the final-step tokens (NCCL) when n_gpus > 1"}


# ----------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------
def run_ours(args):
    from explainable_spatial_vqa_b200 import sharding
    from explainable_spatial_vqa_b200 import synthetic as syn

    rank, local, world = sharding.init_from_env("nccl")
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # host side of the e2e path: every rank's pinned buffers on the NUMA node of its GPU (before anything is allocated)
    numa = sharding.bind_to_gpu_numa(local) if not args.no_numa_bind else {"bound": False, "disabled": True}
    peaks = load_peaks()
    B = args.batch
    depth = max(1, args.pipeline_depth)

    if args.workload == "iqap":
        from explainable_spatial_vqa_b200 import inference_transformer_iqap as iqap
        torch.manual_seed(0)
        model = iqap.VQAModel(85, 256, 256, 32, 44, T_PROG, 196).eval().to(dev)
        # synthetic conv4-like features (post-ReLU) generated on the device; questions from the CPU generator
        g = torch.Generator(device=dev).manual_seed(1234 + rank)
        img = torch.randn(B, 196, 1024, device=dev, generator=g).relu_()
        _, q_cpu = syn.iqap_inputs(B, seed=1234 + rank, relu=False)
        q = q_cpu.to(dev)
        units_per_step = B * T_PROG
        counts = [B] * world

        def step_local():
            return model(img, q)

        gath = sharding.JobGather(args.steps, B, 1 + T_PROG, torch.int64, dev)  # answer + 27 program tokens per question
        step_no = [0]

        def step():
            if depth <= 1:
                ans, prog = model(img, q)
                st = torch.cuda.current_stream()
            else:  # independent batches in flight on `depth` (handle, stream) slots; drained before the clock stops
                ans, prog = model.submit(img, q, depth)
                st = model.last_submit_stream
            with torch.cuda.stream(st):  # results stay on the device; no collective per step
                gath.put(step_no[0], torch.cat([ans.argmax(1, keepdim=True), prog], dim=1))
            step_no[0] += 1
            return ans, prog

        img_host = q_host = None
        if not args.skip_host_e2e:
            img_host = torch.empty(B, 196, 1024, dtype=torch.float32).pin_memory()
            img_host.copy_(img)
            q_host = q_cpu.pin_memory()

        # fp32 host features; with enough host cores per rank the library rounds them to fp16 on host threads before
        # they cross PCIe ("auto", VQAModel.resolve_upload) - conversion inside the timed region
        e2e_upload = iqap.VQAModel.resolve_upload(args.e2e_upload)
        if args.e2e_chunk is None:
            args.e2e_chunk = 512 if e2e_upload == "fp16" else 1024

        def step_e2e():
            if depth <= 1:
                return model.forward_host(img_host, q_host, chunk=args.e2e_chunk, upload=e2e_upload)
            return model.submit_host(img_host, q_host, chunk=args.e2e_chunk, depth=depth, upload=e2e_upload)

        drain_host = model.drain_host

        h2d = B * 196 * 1024 * (2 if e2e_upload == "fp16" else 4) + B * 46 * 8
        d2h = B * 32 * 4 + B * T_PROG * 8
    else:
        from explainable_spatial_vqa_b200 import inference_transformer_full_annotation_new as fa
        torch.manual_seed(0)
        model = fa.MultiModalTransformer(170, 256, 2, 1, 1, 512, 0.1, 50, 196).eval().to(dev)
        g = torch.Generator(device=dev).manual_seed(4321 + rank)
        img = torch.randn(B, 1024, 14, 14, device=dev, generator=g).relu_()
        generator = None
        if args.workload == "e2e":
            from explainable_spatial_vqa_b200 import run_model_lstm_qp as qp
            torch.manual_seed(1)
            generator = qp.Seq2SeqModel(85, 256, 512, 44, T_PROG, 1).eval().to(dev)
            questions = syn.lstm_questions(B, seed=4242 + rank).to(dev)
            synth_programs, node_counts = syn.prefix_programs(B, seed=777 + rank)
            synth_programs = synth_programs.to(dev)
            arity, fmap = syn.program_arity().to(dev), syn.program_func_map().to(dev)
            func, deps, n_steps = qp.programs_to_chain(synth_programs, arity, fmap)
        else:
            func, deps, n_steps = syn.fa_programs(B, seed=4321 + rank)
            n_steps_host = n_steps.clone()  # program lengths known on the host: no device->host read per call
            func, deps, n_steps = func.to(dev), deps.to(dev), n_steps.to(dev)
        units_per_step = int(n_steps.sum())
        counts = [B] * world

        slot_counter = [0]

        gath = sharding.JobGather(args.steps, B, 20, torch.int32, dev)  # the final step's 20 tokens per question
        step_no = [0]
        last_idx = (n_steps - 1).long()
        rows_idx = torch.arange(B, device=dev)

        def chain(slot=0):
            if generator is None:
                return fa.run_inference_chain_batched(model, img, func, deps, n_steps_host, 0, 20, slot=slot)
            generator(questions)  # greedy program decode, all on the device
            f, d, n = qp.programs_to_chain(synth_programs, arity, fmap)  # device glue: prefix program -> chain
            return fa.run_inference_chain_batched(model, img, f, d, n, 0, 20, slot=slot)

        def step_local():
            return chain()

        def step():
            slot = 0
            if depth > 1:
                slot = 1 + slot_counter[0] % depth
                slot_counter[0] += 1
            cache = chain(slot)
            st = model._pool.stream(slot) if slot else torch.cuda.current_stream()
            with torch.cuda.stream(st):  # results stay on the device; no collective per step
                gath.put(step_no[0], cache[rows_idx, last_idx])
            step_no[0] += 1
            return cache

        img_host = None
        if not args.skip_host_e2e:
            img_host = torch.empty(B, 1024, 14, 14, dtype=torch.float32).pin_memory()
            img_host.copy_(img)
        f_h, d_h, n_h = func.cpu().pin_memory(), deps.cpu().pin_memory(), n_steps.cpu().pin_memory()

        if generator is not None:
            q_h, p_h = questions.cpu().pin_memory(), synth_programs.cpu().pin_memory()

        # fp32 host features; with enough host cores per rank they are rounded to bf16 on host threads before the upload
        # (what the device does to them anyway: identical results, half the PCIe bytes)
        fa_upload = fa.resolve_upload({"fp16": "bf16"}.get(args.e2e_upload, args.e2e_upload))
        host_out_no = [0]
        host_out = [torch.empty(B, func.shape[1], 20, dtype=torch.int32).pin_memory() for _ in range(depth)] \
            if generator is None and depth > 1 else []

        def step_e2e():
            if generator is None:
                # the library's host-buffer entry: sub-batches uploaded + projected while the previous one executes;
                # with several steps in flight the upload of a step also runs under the chains of the previous step
                if depth > 1:
                    k = host_out_no[0] % depth
                    host_out_no[0] += 1
                    return fa.submit_inference_chain_host(model, img_host, f_h, d_h, n_h, 0, 20, chunk=args.fa_host_chunk,
                                                          depth=depth, out=host_out[k], upload=fa_upload)
                return fa.run_inference_chain_host(model, img_host, f_h, d_h, n_h, 0, 20, chunk=args.fa_host_chunk,
                                                   parts=args.fa_host_parts, upload=fa_upload)
            else:
                generator(q_h.to(dev, non_blocking=True))
                f, d, n = qp.programs_to_chain(p_h.to(dev, non_blocking=True), arity, fmap)
                cache = fa.run_inference_chain_batched(model, img_host.to(dev, non_blocking=True), f, d, n, 0, 20)
            return cache.cpu()

        h2d = B * 1024 * 196 * (2 if generator is None and fa_upload == "bf16" else 4) + (
            f_h.numel() * 4 + d_h.numel() * 4 + n_h.numel() * 4 if generator is None
                                      else q_h.numel() * 8 + p_h.numel() * 8)
        drain_host = model.drain_host if generator is None else torch.cuda.synchronize
        d2h = B * func.shape[1] * 20 * 4

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, drain, finish=None):
        for _ in range(warmup):
            fn()
        drain()
        barrier()
        launches0 = model.native_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        drain()  # every in-flight batch of the pipeline completes inside the timed region
        if finish is not None:
            finish()  # the job's single result gather (one all_gather_into_tensor over NCCL), also inside
        e1.record()
        barrier()
        ms = sharding.max_over_ranks(e0.elapsed_time(e1), dev)
        return ms, model.native_launch_count() - launches0

    sampler = ClockSampler(local)
    sampler.start()
    ms, launches = timed(step, args.steps, args.warmup, model.drain, gath.finish)
    # sustained behaviour: the same K-step block repeated, median reported next to the first block's value
    block_ms = [ms]
    for _ in range(max(0, args.blocks - 1)):
        m_, _ = timed(step, args.steps, 0, model.drain, gath.finish)
        block_ms.append(m_)
    clocks = sampler.stop()
    # the same K steps strictly one after the other (no batches in flight concurrently), for the record
    serial_ms = None
    if depth > 1:
        depth_saved, depth = depth, 1
        serial_ms, _ = timed(step, args.steps, 1, model.drain, gath.finish)
        depth = depth_saved
    value = world * units_per_step * args.steps / (ms * 1e-3)

    # end-to-end: host buffers, H2D + D2H inside the timed region (wall clock around a synchronous call ==
    # device time here; still reported from CUDA events for consistency)
    # warm-up covers every pipeline slot (each slot allocates its staging buffers on first use)
    if args.skip_host_e2e:  # a job too large to hold a second, pinned copy of its features on the host of an 8-GPU box
        ms_e2e, e2e_value = float("nan"), None
    else:
        # as many warm-up steps as timed ones: every result buffer of the timed loop (pinned host tensors, kept alive until
        # the drain) then comes out of torch's pinned-memory cache - a cudaHostAlloc inside the loop costs the host
        # thread 5-15 ms, and with the fp16 upload mode the host thread is the critical path
        ms_e2e, _ = timed(step_e2e, max(1, args.steps // 2), max(3, depth, args.steps // 2), drain_host)
        e2e_value = world * units_per_step * max(1, args.steps // 2) / (ms_e2e * 1e-3)

    # live per-kernel-class timing (CUDA events on the launch stream) over a few extra steps -> roofline
    roofline, kernels = None, None
    if rank == 0 or world == 1:
        h = model._native()
        prof_steps = 3
        torch.cuda.synchronize()
        h.profile_begin()
        for _ in range(prof_steps):
            h.profile_delay(12.0)  # the host enqueues the step while the GPU is held: events see back-to-back kernels
            step_local()  # rank-local: no collective outside the lock-step timed loop
            torch.cuda.synchronize()
        prof = h.profile_end()
        total = sum(v[0] for v in prof.values())
        kernels = {k: {"ms_per_step": v[0] / prof_steps, "launches_per_step": v[1] / prof_steps,
                       "share": v[0] / total} for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])}
        # algorithmic work of every kernel class per bench step (whole batch of this rank)
        if args.workload == "iqap":
            work = {k: (kind, per_q * B) for k, (kind, per_q) in iqap_work_per_question().items()}
        else:
            d_, ff_, V_, T_ = D, 512, 170, 19
            ns, dp = n_steps.cpu(), deps.cpu()
            S_ = dp.shape[1]
            valid = ((dp >= 0) & (dp < torch.arange(S_)[None, :, None])).sum(-1)      # consumed dependencies per step
            active = (torch.arange(S_)[None, :] < ns[:, None])
            L = (196 + 1 + 20 * valid) * active                                       # sequence rows of every executed step
            Lsum, L2sum, nstep = float(L.sum()), float((L.double() ** 2).sum()), float(active.sum())
            work = {
                "enc_qkv_gemm": ("tensor", 2 * d_ * 3 * d_ * Lsum), "enc_attention": ("tensor", 4 * d_ * L2sum),
                "enc_outproj_ln_gemm": ("tensor", 2 * d_ * d_ * Lsum), "enc_ffn1_gemm": ("tensor", 2 * d_ * ff_ * Lsum),
                "enc_ffn2_ln_gemm": ("tensor", 2 * d_ * ff_ * Lsum), "enc_ffn_fused": ("tensor", 4 * d_ * ff_ * Lsum),
                "dec_cross_kv_gemm": ("tensor", 2 * d_ * 2 * d_ * Lsum),
                "enc_final_ln": ("hbm", Lsum * d_ * 2 * 2), "embed_gather": ("hbm", Lsum * d_ * 2 * 2),
                "dec_self_qkv_gemm": ("tensor", nstep * T_ * 2 * d_ * 3 * d_),
                "dec_cross_q_gemm": ("tensor", nstep * T_ * 2 * d_ * d_),
                "dec_cross_v_gemm": ("tensor", nstep * T_ * 2 * d_ * d_),
                "dec_outproj_ln_gemm": ("tensor", nstep * T_ * 4 * d_ * d_),
                "dec_ffn_split": ("tensor", nstep * T_ * 4 * d_ * ff_),
                "dec_head_argmax": ("tensor", nstep * T_ * 2 * d_ * V_),
                "dec_self_attention": ("hbm", nstep * sum((t + 1) * 2 * d_ * 2 for t in range(T_))),
                # HBM-bound: every decode position re-reads the memory rows (512 B each; K|V rows = 1 KB without absorption)
                "dec_cross_attention": ("hbm", Lsum * (512 if ABSORB else 1024) * T_),
            }

        def roof(name):
            kind, total = work[name]
            per_launch = total / kernels[name]["launches_per_step"]
            dur = kernels[name]["ms_per_step"] / kernels[name]["launches_per_step"] * 1e-3
            if kind == "tensor":
                ach = per_launch / dur / 1e12
                r = {"kernel": name, "bound": "tensor", "achieved": ach, "peak": peaks["bf16_tflops_sustained"],
                     "unit": "TFLOP/s", "frac": ach / peaks["bf16_tflops_sustained"], "traffic": None,
                     "peak_source": peaks["source"] + " (sustained cuBLAS bf16)"}
            else:
                ach = per_launch / dur / 1e9
                r = {"kernel": name, "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": ach / peaks["hbm_gbs"], "traffic": None, "peak_source": peaks["source"] + " (copy)"}
            return r, per_launch, dur

        top = next(k for k in kernels if k in work)  # kernel class with the largest share of the step
        roofline, per_launch, dur = roof(top)
        for k in kernels:
            if k in work:
                rk, _, _ = roof(k)
                kernels[k].update(bound=rk["bound"], achieved=rk["achieved"], unit=rk["unit"], frac=rk["frac"])
        # DRAM bytes per launch of this kernel class from the committed `ncu --set full` capture, if any
        tpath = os.path.join(REPO, "profiles", "traffic.json")
        if args.workload == "iqap" and os.path.exists(tpath):
            with open(tpath) as f:
                tr = json.load(f).get(top)
            if tr:
                roofline["traffic"] = tr["dram_bytes_per_launch"]
                roofline["traffic_source"] = "profiles/" + tr["source"]
        roofline["algorithmic_per_launch"] = per_launch
        roofline["avg_launch_ms"] = dur * 1e3
        if args.workload == "iqap":
            # whole-step tensor-core utilisation: algorithmic FLOPs of the model / step time / peak
            roofline["model_flops_frac_of_bf16_peak"] = (IQAP_FLOPS_PER_QUESTION * B / (ms / args.steps * 1e-3)) / (
                peaks["bf16_tflops_sustained"] * 1e12)

    extra = {}
    # host -> device ceiling with EVERY rank copying at once: the bound of the e2e number at N GPUs (all ranks take part)
    conc = h2d_probe(dev, barrier=barrier)
    conc_min = -sharding.max_over_ranks(-conc, dev)
    conc_sum = sharding.sum_over_ranks(conc, dev)
    if rank == 0:
        extra["numa"] = numa
        extra["h2d_gbs_concurrent"] = {"per_rank_min": conc_min, "sum_over_ranks": conc_sum, "ranks": world,
                                       "e2e_h2d_gbs_per_rank_achieved": None if e2e_value is None else
                                       h2d / (ms_e2e / max(1, args.steps // 2) * 1e-3) / 1e9,
                                       "what": "pinned host -> device copies of 512 MiB started together on all ranks"}
    if rank == 0:
        try:
            extra["h2d_gbs_measured"] = h2d_probe(dev)
            if args.workload == "iqap" and world == 1 and not args.no_cpu_baseline:
                extra["torch_eager_gpu_fp32"] = {
                    "value": gpu_eager_iqap(dev, 256), "unit": "program-steps/s",
                    "what": "oracle (reference algorithm, recompute-every-step, fp32) on this GPU with PyTorch's CUDA "
                            "kernels, 256 questions - context only"}
                extra["torch_kvcached_bf16_gpu"] = {
                    "value": gpu_kvcached_bf16_iqap(dev, B), "unit": "program-steps/s",
                    "what": f"the same algorithm KV-cached (cross K/V once, one position per step) under bf16 autocast "
                            f"with PyTorch's CUDA kernels on this GPU, {B} questions - context only"}
            if args.workload == "iqap" and world == 1 and not args.skip_host_e2e:
                # same questions with CLEVR's ~10 questions per image: every unique image crosses PCIe once
                # (forward_host_indexed); context only - the headline e2e uses one distinct image per question
                n_img = max(1, B // 10)
                idx_host = (torch.arange(B, dtype=torch.int32) % n_img).pin_memory()
                img_u = img_host[:n_img]
                model.forward_host_indexed(img_u, idx_host, q_host, chunk=args.e2e_chunk)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(3):
                    model.forward_host_indexed(img_u, idx_host, q_host, chunk=args.e2e_chunk)
                e1.record()
                torch.cuda.synchronize()
                extra["e2e_indexed_10_questions_per_image"] = {
                    "value": 3 * units_per_step / (e0.elapsed_time(e1) * 1e-3), "unit": "program-steps/s",
                    "h2d_bytes_per_step": img_u.numel() * 4 + q_host.numel() * 8 + idx_host.numel() * 4,
                    "what": "forward_host_indexed: host buffers, unique images uploaded and projected once"}
                # the same questions with the features held as fp16 on the host (the half-size feature store)
                img16 = img_host.half().pin_memory()
                model.forward_host(img16, q_host, chunk=args.e2e_chunk)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(3):
                    model.forward_host(img16, q_host, chunk=args.e2e_chunk)
                e1.record()
                torch.cuda.synchronize()
                extra["e2e_fp16_feature_store"] = {
                    "value": 3 * units_per_step / (e0.elapsed_time(e1) * 1e-3), "unit": "program-steps/s",
                    "h2d_bytes_per_step": img16.numel() * 2 + q_host.numel() * 8,
                    "what": "forward_host on fp16 features (same mantissa width as the tf32 path); serial calls"}
                del img16
        except Exception as e:  # pragma: no cover - context numbers must never break the bench line
            extra["context_error"] = repr(e)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        if args.workload == "iqap":
            sample = args.cpu_sample or 64
            v, dt = cpu_iqap(sample, 2)
            what = f"{sample} of the {B} questions, best of 2 ({dt:.1f} s each)"
        else:
            sample = args.cpu_sample or 4
            v, dt = cpu_fa(sample, with_generator=args.workload == "e2e")
            what = f"{sample} of the {B} questions as a batch-1 loop ({dt:.1f} s)"
        cpu_baseline = {"value": v, "unit": "program-steps/s", "cores": os.cpu_count() or 1, "kind": "port",
                        "cpu": cpu_model_name(),
                        "sample": what + "; oracle in the reference's recompute-every-step form, torch CPU fp32"}

    if rank == 0:
        line = {
            "metric": "program_steps_per_s", "value": value, "unit": "program-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(args),
            "questions_per_s": value / T_PROG if args.workload == "iqap" else world * B * args.steps / (ms * 1e-3),
            "blocks": {"n": len(block_ms), "steps_per_block": args.steps,
                       "ms_per_step_median": sorted(block_ms)[len(block_ms) // 2] / args.steps,
                       "ms_per_step_all": [b / args.steps for b in block_ms],
                       "value_median": world * units_per_step * args.steps / (sorted(block_ms)[len(block_ms) // 2] * 1e-3)},
            "pipeline": {"depth": depth, "what": "independent batches (steps) in flight on separate handle+stream "
                                                 "slots; all drained inside the timed region",
                         "serial_ms_per_step": None if serial_ms is None else serial_ms / args.steps},
            "clocks": clocks,
            "e2e": None if e2e_value is None else
                   {"value": e2e_value, "unit": "program-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / max(1, args.steps // 2),
                    **({"upload": e2e_upload, "host_input_bytes_per_step": B * 196 * 1024 * 4 + B * 46 * 8}
                       if args.workload == "iqap" else ({"upload": fa_upload} if args.workload == "fa" else {}))},
            "gpu_launches": launches,
            "roofline": roofline, "cpu_baseline": cpu_baseline, "kernels": kernels, "context": extra,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="iqap", choices=["iqap", "fa", "e2e"])
    ap.add_argument("--batch", type=int, default=None, help="questions per GPU per step (default 1024 iqap / 4096 fa)")
    ap.add_argument("--e2e-chunk", type=int, default=None,
                    help="questions per upload chunk of the IQAP host-buffer call (default 512 with the fp16 upload mode: the "
                         "conversion of a chunk runs under the upload of the previous one; else 1024)")
    ap.add_argument("--pipeline-depth", type=int, default=None,
                    help="independent batches in flight (1 = strictly serial steps); default 2 (iqap) / 3 (fa, e2e)")
    ap.add_argument("--blocks", type=int, default=5, help="timed K-step blocks (the first gives `value`; the median is reported too)")
    ap.add_argument("--e2e-upload", default="auto", choices=["auto", "fp32", "fp16"],
                    help="IQAP e2e: how the fp32 host features cross PCIe (auto: fp16 rounded on host threads when the rank "
                         "has >= 8 CPUs to itself)")
    ap.add_argument("--fa-host-chunk", type=int, default=None,
                    help="questions per sub-batch of the FA host-buffer call (default: the whole batch when steps are "
                         "pipelined - uploads of consecutive steps queue on one ingest stream - else 1024)")
    ap.add_argument("--fa-host-parts", type=int, default=4, help="concurrent parts (handle + stream slots) of the FA host-buffer call")
    ap.add_argument("--cpu-sample", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-host-e2e", action="store_true",
                    help="device-resident measurement only (no pinned host copy of the features: very large batches)")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not pin ranks to their GPU's NUMA node")
    args = ap.parse_args()
    if args.batch is None:
        args.batch = 1024 if args.workload == "iqap" else 4096
    if args.pipeline_depth is None:
        args.pipeline_depth = 2 if args.workload == "iqap" else 3
    if args.fa_host_chunk is None:
        args.fa_host_chunk = args.batch if args.pipeline_depth > 1 else 1024
    # every pipeline slot allocates its workspace and captures its graphs on first use: all of them warm up
    args.warmup = max(args.warmup, 3, args.pipeline_depth) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
